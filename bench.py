#!/usr/bin/env python
"""Benchmark of the phasing hot path (BASELINE.json metric) -- see the contract in DESIGN.md.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

A step is one pass of the whole hot path (window search, site classification, read scan, read-site
allele lookup, chaining, tally, final call) over one batch: BASELINE config[1], a synthetic trio
with ``--dnms`` (10 000) SNV DNMs, 5 kb search distance, 30x 150 bp reads, extended read-backed
phasing.  With N ranks every rank phases its own trio of that size (DNMs shard by kid: weak scaling,
no collective on the data path).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DNMs phased/sec"
UNIT = "DNM/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--dnms", type=int, default=10000, help="DNMs per trio (per rank)")
    ap.add_argument("--search-dist", type=int, default=5000)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=0, help="DNMs in the CPU-baseline sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args):
    return {
        "workload": "BASELINE config[1]: synthetic trio, %d SNV DNMs, %d bp search-dist, %gx 150bp reads, "
                    "extended read-backed phasing" % (args.dnms, args.search_dist, args.coverage),
        "dnms_per_gpu": args.dnms, "search_dist": args.search_dist, "coverage": args.coverage,
        "readlen": 150, "site_spacing_bp": 300, "noise": True,
        "parallelism": "shard-by-kid x%d, no collective" % args.gpus,
        "l2_policy": "inputs (GBs of site/read columns) exceed the 126 MB L2; no flush needed",
    }


def make_ds(args, rank, n_dnms=None):
    from unfazed_b200.synth import SynthConfig, make_dataset
    return make_dataset(SynthConfig(dnms_per_trio=n_dnms or args.dnms, search_dist=args.search_dist,
                                    coverage=args.coverage, noise=True, seed=args.seed + 1000 * rank))


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every 2 ms from a
    thread (the timed region of this benchmark is tens of milliseconds, too short for nvidia-smi -lms)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)): "sw_power_cap",
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self.stop_flag = True
        if self.thread is not None:
            self.thread.join(timeout=1)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "samples": 0, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "samples": len(self.samples),
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------
# CPU baseline (oracle port)
# ----------------------------------------------------------------------------------------------
def _port_run(ds, dnms):
    import copy
    from oracle import port
    ph = port.Phaser(ds.sites, ds.reads, ds.pedigrees, port.Params(threads=1, readlen=151))
    t0 = time.perf_counter()
    recs = ph.phase(copy.deepcopy(dnms))
    return time.perf_counter() - t0, len(recs)


_SHARED = {}


def _port_worker(idx):
    ds = _SHARED["ds"]
    return _port_run(ds, _SHARED["shards"][idx])


def cpu_baseline_scalar(ds, n_sample):
    dn = ds.dnms[:n_sample]
    dt, nrec = _port_run(ds, dn)
    return {"value": len(dn) / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "first %d DNMs of the same workload, oracle/port.py (pure Python restatement of the reference), "
                      "%.1f s, %d records" % (len(dn), dt, nrec)}


def run_reference(args):
    """The reference's CPU implementation of the path on all host cores: the oracle port (the
    reference itself is Python + cyvcf2/pysam and is not installable here), process-sharded."""
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores * 60, 240)      # a few seconds of all-core CPU work per step
    n_gen = min(args.dnms, per_step)
    ds = make_ds(args, 0, n_dnms=n_gen)
    shards = [ds.dnms[i::cores] for i in range(cores)]
    shards = [s for s in shards if s]
    _SHARED["ds"], _SHARED["shards"] = ds, shards
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(len(shards)) as pool:
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_port_worker, range(len(shards)))
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
            if sum(times) > 240:
                break
    n = len(ds.dnms)
    ms = 1000.0 * float(np.mean(times))
    value = n / (ms / 1000.0)
    sample = ("%d DNMs per step (same generator and per-DNM shape as the %d-DNM workload), oracle/port.py "
              "sharded over %d processes" % (n, args.dnms, len(shards)))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64",
        "data": "synthetic", "config": workload_config(args), "impl": "reference",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": len(shards), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# the CUDA path
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from unfazed_b200.engine import Engine, make_params
    from unfazed_b200.phaser import BatchPhaser
    from unfazed_b200.plan import plan_find_fast as plan_find

    t_gen = time.perf_counter()
    ds = make_ds(args, rank)
    t_gen = time.perf_counter() - t_gen
    eng = Engine(local_rank)
    dev = eng.device
    bp = BatchPhaser(eng, ds.sites, ds.reads, ds.pedigrees)
    params = make_params(readlen=151)
    cul = bp.cul(151, 1000000, 3)
    plan_kw = dict(search_dist=args.search_dist, whole_region=False, build="38", multiread_proc_min=10 ** 9,
                   threads=1, with_reads=True)
    plan = plan_find(ds.dnms, ds.pedigrees, bp.sidx, ds.reads, **plan_kw)
    n_dnms = len(ds.dnms)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        return eng.run(bp.dsites, bp.dreads, plan, params, blk_cul=cul, download=True, keep_device=False)

    # The synthetic dataset is millions of long-lived Python objects; a generation-2 garbage collection in
    # the middle of a step costs tens of milliseconds of host time.  Collect once, then park the survivors.
    import gc
    gc.collect()
    gc.freeze()
    res = None
    for _ in range(max(args.warmup, 3)):
        res = step_resident()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        res = step_resident()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches_per_step = res.launches
    n_pairs, n_hits, n_reads = res.n_pairs, res.n_hits, ds.reads.n_reads
    phased = int((res.calls_strict["emitted"] == 1).sum())

    # per-kernel timings (CUDA events on the launching stream), averaged over the same K steps
    stage = {}
    for _ in range(args.steps):
        r = eng.run(bp.dsites, bp.dreads, plan, params, blk_cul=cul, time_stages=True, keep_device=False)
        for k, v in r.timings_ms.items():
            stage[k] = stage.get(k, 0.0) + v / args.steps

    # end to end: host (pinned) tables -> device -> kernels -> results on the host, every step
    pinned_sites = ds.sites
    pinned_reads = ds.reads
    h2d_bytes = bp.dsites.nbytes + bp.dreads.nbytes + plan.dnm.nbytes + plan.seg.nbytes + plan.alleles.nbytes
    d2h_bytes = n_dnms * (32 + 16 + 16 + 4 * 4)
    del bp.dsites, bp.dreads
    torch.cuda.empty_cache()
    host_site_t = _pin_table(ds.sites)
    host_read_t = eng.pack_reads(ds.reads, min_gt_qual=20, pin=True)

    e2e_parts = {}

    def step_e2e(probe=False):
        # the H2D copies are asynchronous (pinned source): the host plans the windows while they fly
        if probe:
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
        h0 = time.perf_counter()
        dsites = eng.upload_sites(host_site_t, pin=False)
        dreads = eng.upload_reads(host_read_t, pin=False)
        h1 = time.perf_counter()
        if probe:
            eb.record()
        pl = plan_find(ds.dnms, ds.pedigrees, bp.sidx, ds.reads, **plan_kw)
        h2 = time.perf_counter()
        out = eng.run(dsites, dreads, pl, params, blk_cul=cul, download=True, keep_device=False)
        h3 = time.perf_counter()
        if probe:       # one untimed step: where the end-to-end time goes
            e2e_parts.update(h2d_copies_gpu_ms=ea.elapsed_time(eb), issue_uploads_host_ms=(h1 - h0) * 1e3,
                             plan_host_ms=(h2 - h1) * 1e3, run_until_results_host_ms=(h3 - h2) * 1e3)
        return out

    for _ in range(2):
        step_e2e()
    step_e2e(probe=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps

    ms_step = ms_total / args.steps
    t_max = torch.tensor([ms_step, e2e_s * 1000.0], device=dev, dtype=torch.float64)
    tot = torch.tensor([float(n_dnms), float(n_pairs), float(n_reads), float(phased)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_max, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_step, e2e_ms = float(t_max[0]), float(t_max[1])
    all_dnms, all_pairs, all_reads, all_phased = (float(x) for x in tot)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        hd = ds.reads.hdr
        rs_bytes = float(32 * n_reads + 4 * int(hd["n_cigar"].sum()) + int(hd["l_seq"].sum()) / 8.0 + 20 * n_reads)
        survey_read_bytes = float(24 * n_reads + 4 * int(hd["n_cigar"].sum()) + int(np.ceil(hd["l_seq"] / 4).sum())
                                  + int(hd["l_seq"].sum()) + 4 * n_hits)
        kern = {
            "classify_sites": {"ms": stage.get("classify_sites", 0.0), "bytes": 45.0 * n_pairs},
            "read_scan": {"ms": stage.get("read_scan", 0.0), "bytes": rs_bytes},
            "read_site_alleles": {"ms": stage.get("read_site_alleles", 0.0), "bytes": None},
            "chain_tally": {"ms": stage.get("chain_tally", 0.0), "bytes": None},
            "compact_sites+scan": {"ms": stage.get("compact_sites", 0.0), "bytes": None},
            "window_search+scan": {"ms": stage.get("window_search", 0.0), "bytes": None},
            "chain_size+scans": {"ms": stage.get("chain_size", 0.0), "bytes": None},
        }
        for k, v in kern.items():
            if v["bytes"] and v["ms"] > 0:
                v["gbps"] = v["bytes"] / (v["ms"] * 1e6)
                v["frac"] = v["gbps"] / peak
        try:
            sat = saturating_classify(eng, max(args.steps, 3))
            sat["gbps"] = sat["bytes"] / (sat["classify_ms"] * 1e6)
            sat["frac"] = sat["gbps"] / peak
            sat["pairs_per_s"] = sat["pairs"] / (sat["classify_ms"] / 1000.0)
        except Exception as e:  # keep the headline line even if the extra measurement cannot run
            sat = {"error": repr(e)}
        dom = max(kern, key=lambda k: kern[k]["ms"])
        domk = kern[dom]
        lookup_ms = kern["read_scan"]["ms"] + kern["read_site_alleles"]["ms"]
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tj.get(dom)
            if ent and ent.get("dnms") == args.dnms:
                traffic = ent["dram_bytes_per_launch"]
        except Exception:
            pass
        roof = {
            "kernel": dom, "bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
            "achieved": domk.get("gbps"), "frac": domk.get("frac"), "traffic": traffic,
            "ms_per_launch": domk["ms"],
            "kernels": kern, "stages_ms": {k: round(v, 4) for k, v in stage.items()},
            "classify_sites_saturating": sat,
            "read_allele_lookup_survey_bytes": {
                "bytes": survey_read_bytes, "ms": lookup_ms,
                "gbps": survey_read_bytes / (lookup_ms * 1e6) if lookup_ms > 0 else None,
                "frac": survey_read_bytes / (lookup_ms * 1e6) / peak if lookup_ms > 0 else None},
        }
        if domk.get("gbps") is None:
            roof["note"] = ("dominant kernel %s is latency-bound graph work (SURVEY 8(d): not a bandwidth target); "
                            "bandwidth kernels are listed under kernels" % dom)
        line = {
            "metric": METRIC, "value": all_dnms / (ms_step / 1000.0), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32/u8 + f64 allele balance", "data": "synthetic",
            "config": workload_config(args),
            "secondary": {
                "informative_site_pairs_classified_per_s_kernel": all_pairs / world / (kern["classify_sites"]["ms"] / 1000.0) * world if kern["classify_sites"]["ms"] > 0 else None,
                "reads_examined_per_s_kernel": all_reads / world / (lookup_ms / 1000.0) * world if lookup_ms > 0 else None,
                "pairs": all_pairs, "reads": all_reads, "hits": n_hits, "dnms_with_call": all_phased,
                "dataset_gen_s": t_gen,
            },
            "roofline": roof,
            "e2e": {"value": all_dnms / (e2e_ms / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                    "d2h_bytes_per_step": int(d2h_bytes), "ms_per_step": e2e_ms,
                    "includes": "host window planning + H2D of all site/read columns + kernels + D2H of tallies/calls",
                    "breakdown_ms": {k: round(v, 3) for k, v in e2e_parts.items()}},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            n_s = args.cpu_sample or 150
            line["cpu_baseline"] = cpu_baseline_scalar(ds, min(n_s, n_dnms))
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def saturating_classify(eng, steps, n_rows=1 << 25, rows_per_window=336):
    """Bandwidth measurement of the site classifier alone (SURVEY 8(d) "saturating variant"): the
    named configs move only 15-150 MB through it, i.e. they are launch-bound.  2^25 rows (1.5 GB of
    site columns, generated on the device) are covered once by non-overlapping DNM windows."""
    import ctypes as C
    import torch
    from unfazed_b200 import _lib as L
    from unfazed_b200.engine import make_params
    from unfazed_b200.plan import Plan
    dev = eng.device
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    V = n_rows
    pos = (torch.arange(V, device=dev, dtype=torch.int32) * 3)
    flag = (torch.rand(V, device=dev, generator=g) < 0.98).to(torch.uint8)
    gt_codes = torch.tensor([0, 1, 3, 2], device=dev, dtype=torch.uint8)
    gt = gt_codes[torch.multinomial(torch.tensor([0.45, 0.35, 0.19, 0.01], device=dev), 3 * V, replacement=True, generator=g)].reshape(3, V).contiguous()
    gq = torch.where(torch.rand(3, V, device=dev, generator=g) < 0.05, torch.rand(3, V, device=dev, generator=g) * 40, torch.full((3, V), 99.0, device=dev)).to(torch.float32).contiguous()
    depth = torch.poisson(torch.full((3, V), 30.0, device=dev), generator=g).to(torch.int32)
    frac = torch.where(gt == 1, 0.5, torch.where(gt == 3, 0.99, 0.01))
    ad = torch.round(depth * frac + torch.randn(3, V, device=dev, generator=g) * 2).clamp(min=0).to(torch.int32)
    ad = torch.minimum(ad, depth).contiguous()
    rd = (depth - ad).contiguous()
    ref = torch.full((V,), 65, device=dev, dtype=torch.uint8)
    alt = torch.full((V,), 67, device=dev, dtype=torch.uint8)
    blk_off = torch.tensor([0, V], device=dev, dtype=torch.int64)

    from unfazed_b200.engine import make_site_cols, pack_site_rows

    class Raw:
        pass
    ds = Raw()
    ds.n_rows = V
    meta, rec, dep = pack_site_rows(pos, flag, gt, gq, rd, ad)
    del flag, gt, gq, rd, ad, depth, frac
    ds.keep = (pos, ref, alt, blk_off, meta, rec, dep)
    ds.cols = make_site_cols(V, 1, blk_off, pos, ref, alt, meta, rec, dep)
    n_win = V // rows_per_window
    seg = np.zeros(n_win, dtype=L.SEG_DTYPE)
    seg["sblk"] = 0
    seg["lo_pos"] = np.arange(n_win, dtype=np.int64) * rows_per_window * 3
    seg["hi_pos"] = seg["lo_pos"] + rows_per_window * 3 - 1
    seg["mult"] = 1
    seg["dnm"] = np.arange(n_win)
    seg["mode"] = L.MODE_READ
    dnm = np.zeros(n_win, dtype=L.DNM_DTYPE)
    dnm["seg_lo"] = np.arange(n_win)
    dnm["seg_hi"] = dnm["seg_lo"] + 1
    dnm["rblk"], dnm["cnv_entry"] = -1, -1
    dnm["pos"] = seg["lo_pos"] + 1
    dnm["end"] = dnm["pos"] + 1
    plan = Plan(dnm=dnm, seg=seg, alleles=np.zeros(0, np.uint8), entries=[{}] * n_win, trio=np.zeros(n_win, np.int32),
                found=np.ones(n_win, bool))
    params = make_params()
    for _ in range(3):
        eng.run(ds, None, plan, params, keep_device=False)
    ms = {}
    for _ in range(steps):
        r = eng.run(ds, None, plan, params, time_stages=True, keep_device=False)
        for k, v in r.timings_ms.items():
            ms[k] = ms.get(k, 0.0) + v / steps
    return {"pairs": int(r.n_pairs), "rows": V, "windows": n_win, "classify_ms": ms.get("classify_sites"),
            "compact_ms": ms.get("compact_sites"), "window_search_ms": ms.get("window_search"),
            "bytes": 45.0 * r.n_pairs}


def _pin_table(t):
    """Copy of a table whose big arrays live in pinned host memory (for the end-to-end leg)."""
    import copy
    import torch
    out = copy.copy(t)
    for name in ("pos", "flag", "ref", "alt", "gt", "gq", "rd", "ad"):
        if hasattr(t, name):
            a = getattr(t, name)
            p = torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
            setattr(out, name, p)
    if hasattr(t, "hdr"):
        raw = torch.from_numpy(t.hdr.view(np.uint8).reshape(-1).copy()).pin_memory().numpy()
        out.hdr = raw.view(t.hdr.dtype)
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
