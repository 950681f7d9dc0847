"""Batch engine: columnar tables + a window plan -> device pipeline -> per-DNM results.

PyTorch is used only for device memory, streams and events; every kernel is in
``libunfazed_sm100.so`` and is reached through the C ABI (``_lib.py``).  There is no CPU fallback:
constructing an ``Engine`` without the library or without a GPU raises.

Pipeline (one stream, in order):

  K0 window_search -> scan -> K1 classify_sites -> compact_sites -> scan(row marks)
  -> K2 read_scan -> scan(hit counts) -> K3 read_site_alleles
  -> chain_size -> 6 scans -> K4 chain_tally -> K5 summarize
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from .plan import Plan
from .schema import ReadTable, SiteTable


def make_params(ab_homref=(0.0, 0.2), ab_homalt=(0.8, 1.0), ab_het=(0.2, 0.8), min_gt_qual=20, min_depth=10,
                min_map_qual=1, readlen=151, insert_size_max_sample=1000000, no_extended=False,
                evidence_min_ratio=10, split_error_margin=5) -> L.Params:
    p = L.Params()
    p.ab_homref[0], p.ab_homref[1] = float(ab_homref[0]), float(ab_homref[1])
    p.ab_homalt[0], p.ab_homalt[1] = float(ab_homalt[0]), float(ab_homalt[1])
    p.ab_het[0], p.ab_het[1] = float(ab_het[0]), float(ab_het[1])
    p.min_gt_qual = float(min_gt_qual)
    p.min_depth = int(min_depth)
    p.min_map_qual = int(min_map_qual)
    p.readlen = int(readlen)
    p.ext_read_goal = int(min(insert_size_max_sample, 2**31 - 1))
    p.no_extended = int(bool(no_extended))
    p.evidence_min_ratio = int(evidence_min_ratio)
    p.split_error_margin = int(split_error_margin)
    return p


def pack_site_rows(pos, flag, gt, gq, rd, ad):
    """Device-side repack of the SoA genotype columns into the classifier's row layout
    (meta u32, rec 4xf32, dep 6xi32 -- see include/unfazed_sm100.h).  All torch ops on the device."""
    g = gt.to(torch.int32)
    meta = (flag.to(torch.int32) | (g[0] << 8) | (g[1] << 16) | (g[2] << 24)).contiguous()
    rec = torch.stack([pos.view(torch.float32), gq[0], gq[1], gq[2]], dim=1).contiguous()
    dep = torch.stack([rd[0], ad[0], rd[1], ad[1], rd[2], ad[2]], dim=1).contiguous()
    return meta, rec, dep


class DeviceSites:
    def __init__(self, table: SiteTable, device: torch.device, pin: bool = False):
        self.table = table
        self.n_rows = table.n_rows
        up = lambda a: _to_device(np.ascontiguousarray(a), device, pin)
        self.blk_off = up(table.blk_off.astype(np.int64))
        self.pos, self.ref, self.alt = up(table.pos), up(table.ref), up(table.alt)
        flag, gt, gq, rd, ad = up(table.flag), up(table.gt), up(table.gq), up(table.rd), up(table.ad)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in (self.pos, self.ref, self.alt, flag, gt, gq, rd, ad))
        self.meta, self.rec, self.dep = pack_site_rows(self.pos, flag, gt, gq, rd, ad)
        self.cols = make_site_cols(table.n_rows, table.n_blocks, self.blk_off, self.pos, self.ref, self.alt,
                                   self.meta, self.rec, self.dep)

    @property
    def nbytes(self) -> int:
        return self.h2d_bytes


def make_site_cols(n_rows, n_blocks, blk_off, pos, ref, alt, meta, rec, dep) -> L.SiteCols:
    c = L.SiteCols()
    c.n_rows, c.n_blocks = n_rows, n_blocks
    c.blk_off, c.pos, c.ref, c.alt = blk_off.data_ptr(), pos.data_ptr(), ref.data_ptr(), alt.data_ptr()
    c.meta, c.rec, c.dep = meta.data_ptr(), rec.data_ptr(), dep.data_ptr()
    return c


class DeviceReads:
    def __init__(self, table: ReadTable, device: torch.device, pin: bool = False):
        self.table = table
        self.n_reads = table.n_reads
        self.max_l_seq = int(table.hdr["l_seq"].max()) if table.n_reads else 0
        up = lambda a, pad=0: _to_device(np.ascontiguousarray(a), device, pin, pad)
        self.blk_off = up(table.blk_off.astype(np.int64))
        self.hdr = up(table.hdr.view(np.uint8).reshape(-1))
        self.cigar = up(table.cigar)
        # tail padding: the TMA bulk copies round the staged span up to 16 B
        self.qual = up(table.qual, 32)
        self.seq2 = up(table.seq2, 16)
        self.blk_sblk = torch.full((max(table.n_blocks, 1),), -1, dtype=torch.int32, device=device)
        self.blk_cul = torch.zeros((max(table.n_blocks, 1),), dtype=torch.float64, device=device)
        c = L.ReadCols()
        c.n_reads, c.n_blocks = table.n_reads, table.n_blocks
        c.blk_off, c.blk_sblk, c.blk_cul = self.blk_off.data_ptr(), self.blk_sblk.data_ptr(), self.blk_cul.data_ptr()
        c.hdr, c.cigar, c.qual, c.seq2 = self.hdr.data_ptr(), self.cigar.data_ptr(), self.qual.data_ptr(), self.seq2.data_ptr()
        c.n_qual, c.n_cigar = int(table.qual.shape[0]), int(table.cigar.shape[0])
        self.cols = c

    @property
    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.hdr, self.cigar, self.qual, self.seq2))


def _to_device(a: np.ndarray, device, pin: bool, pad: int = 0) -> torch.Tensor:
    """Host array -> device tensor (+ ``pad`` zero elements).  Arrays that already live in pinned
    memory are copied asynchronously; ``pin`` stages pageable arrays through a pinned copy."""
    t = torch.from_numpy(a.view(np.uint8).reshape(-1)) if a.dtype.fields is not None else torch.from_numpy(a)
    if pin and not t.is_pinned():
        t = t.pin_memory()
    out = torch.empty(t.shape[:-1] + (t.shape[-1] + pad,), dtype=t.dtype, device=device) if pad else torch.empty_like(t, device=device)
    if pad:
        out[..., t.shape[-1]:].zero_()
        out[..., : t.shape[-1]].copy_(t, non_blocking=True)
    else:
        out.copy_(t, non_blocking=True)
    return out


@dataclass
class BatchResult:
    plan: Plan
    n_pairs: int
    n_hits: int
    seg_row_lo: np.ndarray
    seg_pair_off: np.ndarray
    n_het: np.ndarray
    n_cand: np.ndarray
    cnv_dad: np.ndarray
    cnv_mom: np.ndarray
    tally: Optional[np.ndarray] = None
    calls_strict: Optional[np.ndarray] = None
    calls_ambiguous: Optional[np.ndarray] = None
    win: Optional[np.ndarray] = None          # [4, n]: read-index ranges [a_lo,a_hi) U [b_lo,b_hi) of every entry
    slot_off: Optional[np.ndarray] = None
    timings_ms: Dict[str, float] = field(default_factory=dict)
    launches: int = 0
    _dev: dict = field(default_factory=dict)

    # lazily downloaded bulk outputs ------------------------------------------------------
    def _np(self, name):
        key = "_np_" + name
        if key not in self._dev:
            self._dev[key] = self._dev[name].cpu().numpy()
        return self._dev[key]

    def class_codes(self) -> np.ndarray:
        return self._np("cls")

    def het_rows(self, d: int) -> np.ndarray:
        base = int(self.seg_pair_off[self.plan.dnm["seg_lo"][d]]) if self.plan.dnm["seg_hi"][d] > self.plan.dnm["seg_lo"][d] else 0
        return self._np("het_list")[base: base + int(self.n_het[d])]

    def cand_words(self, d: int) -> np.ndarray:
        base = int(self.seg_pair_off[self.plan.dnm["seg_lo"][d]]) if self.plan.dnm["seg_hi"][d] > self.plan.dnm["seg_lo"][d] else 0
        return self._np("cand_list")[base: base + int(self.n_cand[d])]

    def cand_evidence(self, d: int) -> np.ndarray:
        base = int(self.seg_pair_off[self.plan.dnm["seg_lo"][d]]) if self.plan.dnm["seg_hi"][d] > self.plan.dnm["seg_lo"][d] else 0
        return self._np("cand_evid")[base: base + int(self.n_cand[d])]

    def slot_labels(self, d: int) -> np.ndarray:
        a, b = int(self.slot_off[d]), int(self.slot_off[d + 1])
        return self._np("slot_label")[a:b]

    def slot_evidence(self, d: int) -> np.ndarray:
        a, b = int(self.slot_off[d]), int(self.slot_off[d + 1])
        return self._np("slot_evid")[a:b]

    def slot_reads(self, d: int, slots: np.ndarray) -> np.ndarray:
        """Read index behind window slots of entry d."""
        a_lo, a_hi, b_lo, _b_hi = (int(x) for x in self.win[:, d])
        na = a_hi - a_lo
        slots = np.asarray(slots, dtype=np.int64)
        return np.where(slots < na, a_lo + slots, b_lo + (slots - na))

    def read_summaries(self) -> np.ndarray:
        return self._np("rsum").view(L.RSUM_DTYPE)


class Engine:
    """One engine per (process, GPU)."""

    def __init__(self, device: int = 0):
        if not torch.cuda.is_available():
            raise RuntimeError("unfazed_b200 needs a CUDA device (sm_100a); none is visible -- there is no CPU path")
        self.lib = L.load()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        ctx = C.c_void_p()
        rc = self.lib.unfz_ctx_create(device, C.byref(ctx))
        if rc != 0:
            raise RuntimeError("unfz_ctx_create failed (%d): libunfazed_sm100.so targets sm_100a only" % rc)
        self.ctx = ctx

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.unfz_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------
    def upload_sites(self, table: SiteTable, pin: bool = False) -> DeviceSites:
        return DeviceSites(table, self.device, pin)

    def upload_reads(self, table: ReadTable, pin: bool = False) -> DeviceReads:
        return DeviceReads(table, self.device, pin)

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, self.lib.unfz_last_error(self.ctx).decode()))

    def _empty(self, n, dtype):
        return torch.empty((max(int(n), 1),), dtype=dtype, device=self.device)

    def _zeros(self, n, dtype):
        return torch.zeros((max(int(n), 1),), dtype=dtype, device=self.device)

    def _scan_work(self, n):
        return self._empty(self.lib.unfz_scan_work_bytes(int(n)), torch.uint8)

    # ------------------------------------------------------------------------------------
    def run(self, dsites: DeviceSites, dreads: Optional[DeviceReads], plan: Plan, params: L.Params,
            blk_cul: Optional[np.ndarray] = None, time_stages: bool = False, download: bool = True,
            keep_device: bool = True) -> BatchResult:
        lib, ctx, dev = self.lib, self.ctx, self.device
        st = torch.cuda.current_stream(dev)
        s = C.c_void_p(st.cuda_stream)
        n_dnms, n_segs = int(plan.dnm.shape[0]), int(plan.seg.shape[0])
        ev: List = []
        launches = 0

        def mark(name):
            if time_stages:
                e = torch.cuda.Event(enable_timing=True)
                e.record(st)
                ev.append((name, e))

        up = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).to(dev) if a.size else self._empty(1, torch.uint8)
        d_dnm, d_seg, d_all = up(plan.dnm), up(plan.seg), up(plan.alleles)
        sc = C.byref(dsites.cols)
        mark("start")

        # ---- K0 + scan ---------------------------------------------------------------------
        seg_row_lo = self._empty(n_segs, torch.int32)
        seg_count = self._empty(n_segs, torch.int64)
        seg_pair_off = self._zeros(n_segs + 1, torch.int64)
        self._check(lib.unfz_window_search(ctx, sc, d_seg.data_ptr(), n_segs, seg_row_lo.data_ptr(), seg_count.data_ptr(), s), "window_search")
        work = self._scan_work(max(n_segs, dsites.n_rows, dreads.n_reads if dreads else 0, n_dnms) + 1)
        self._check(lib.unfz_exclusive_scan_i64(ctx, seg_count.data_ptr(), seg_pair_off.data_ptr(), n_segs, work.data_ptr(), s), "scan(pairs)")
        launches += 4
        mark("window_search")
        h_pair_off = seg_pair_off.cpu().numpy()[: n_segs + 1]
        n_pairs = int(h_pair_off[n_segs]) if n_segs > 0 else 0

        # ---- K1 + compaction + marks ---------------------------------------------------------
        cls = self._empty(n_pairs, torch.uint8)
        het_list = self._empty(n_pairs, torch.int32)
        cand_list = self._empty(n_pairs, torch.int32)
        n_het = self._zeros(n_dnms, torch.int32)
        n_cand = self._zeros(n_dnms, torch.int32)
        cnv_dad = self._zeros(n_dnms, torch.int32)
        cnv_mom = self._zeros(n_dnms, torch.int32)
        row_mark = self._zeros(dsites.n_rows, torch.uint8)
        mark_prefix = self._zeros(dsites.n_rows + 1, torch.int32)
        mark("alloc1")
        self._check(lib.unfz_classify_sites(ctx, sc, d_seg.data_ptr(), seg_row_lo.data_ptr(), seg_pair_off.data_ptr(),
                                            n_segs, n_pairs, C.byref(params), cls.data_ptr(), s), "classify_sites")
        launches += 1 if n_pairs else 0
        mark("classify_sites")
        self._check(lib.unfz_compact_sites(ctx, d_dnm.data_ptr(), n_dnms, d_seg.data_ptr(), seg_row_lo.data_ptr(),
                                           seg_pair_off.data_ptr(), cls.data_ptr(), het_list.data_ptr(), n_het.data_ptr(),
                                           cand_list.data_ptr(), n_cand.data_ptr(), cnv_dad.data_ptr(), cnv_mom.data_ptr(),
                                           row_mark.data_ptr(), s), "compact_sites")
        self._check(lib.unfz_exclusive_scan_u8_i32(ctx, row_mark.data_ptr(), mark_prefix.data_ptr(), dsites.n_rows,
                                                   work.data_ptr(), s), "scan(marks)")
        launches += 4
        mark("compact_sites")

        res = BatchResult(plan=plan, n_pairs=n_pairs, n_hits=0, seg_row_lo=None, seg_pair_off=h_pair_off,
                          n_het=None, n_cand=None, cnv_dad=None, cnv_mom=None)
        dv = res._dev
        dv.update(cls=cls, het_list=het_list, cand_list=cand_list)
        tally = self._zeros(n_dnms * 8, torch.int32)
        calls_s = self._zeros(n_dnms * 4, torch.int32)
        calls_a = self._zeros(n_dnms * 4, torch.int32)

        has_reads = dreads is not None and dreads.n_reads > 0 and bool((plan.dnm["rblk"] >= 0).any())
        if has_reads:
            rc_ = C.byref(dreads.cols)
            sb = np.full(max(dreads.table.n_blocks, 1), -1, dtype=np.int32)
            for rb, sblk in plan.rblk_sblk.items():
                sb[rb] = sblk
            dreads.blk_sblk.copy_(torch.from_numpy(sb))
            if blk_cul is not None:
                dreads.blk_cul[: blk_cul.shape[0]].copy_(torch.from_numpy(np.ascontiguousarray(blk_cul, dtype=np.float64)))
            # ---- K2 + scan + K3 ----------------------------------------------------------------
            rsum = self._empty(dreads.n_reads * 16, torch.uint8)
            blk_maxspan = self._zeros(dreads.table.n_blocks, torch.int32)
            total_hits = self._zeros(1, torch.int64)
            mark("alloc2")
            self._check(lib.unfz_read_scan(ctx, rc_, sc, mark_prefix.data_ptr(), C.byref(params), dreads.max_l_seq, rsum.data_ptr(),
                                           blk_maxspan.data_ptr(), s), "read_scan")
            launches += 1
            mark("read_scan")
            self._check(lib.unfz_exclusive_scan_u16_u32(ctx, rsum.data_ptr() + 10, 16, rsum.data_ptr() + 12, 16,
                                                        dreads.n_reads, total_hits.data_ptr(), work.data_ptr(), s), "scan(hits)")
            launches += 3
            n_hits = int(total_hits.item())
            if n_hits >= 2**32:
                raise RuntimeError("more than 2^32 read x site hits in one batch")
            hits = self._empty(n_hits, torch.int32)
            mark("scan_hits")
            self._check(lib.unfz_read_site_alleles(ctx, rc_, sc, row_mark.data_ptr(), mark_prefix.data_ptr(),
                                                   rsum.data_ptr(), hits.data_ptr(), s), "read_site_alleles")
            launches += 1
            mark("read_site_alleles")
            res.n_hits = n_hits
            # ---- chain sizing + scans ------------------------------------------------------------
            win = self._zeros(4 * n_dnms, torch.int32)
            need = self._zeros(6 * n_dnms, torch.int64)
            off = self._zeros(6 * (n_dnms + 1), torch.int64)
            self._check(lib.unfz_chain_size(ctx, d_dnm.data_ptr(), n_dnms, d_seg.data_ptr(), seg_pair_off.data_ptr(), sc, rc_,
                                            rsum.data_ptr(), blk_maxspan.data_ptr(), het_list.data_ptr(), n_het.data_ptr(),
                                            cand_list.data_ptr(), n_cand.data_ptr(), win.data_ptr(),
                                            need.data_ptr(), s), "chain_size")
            self._check(lib.unfz_exclusive_scan_rows_i64(ctx, need.data_ptr(), off.data_ptr(), 6, n_dnms, s), "scan(need)")
            launches += 2
            h_off = off.cpu().numpy().reshape(6, n_dnms + 1)
            totals = np.ascontiguousarray(h_off[:, n_dnms]).astype(np.int64)
            mark("chain_size")
            nbytes = lib.unfz_chain_scratch_bytes(*(int(x) for x in totals), n_dnms)
            scratch = self._empty(nbytes, torch.uint8)
            slot_label = self._zeros(int(totals[0]) + 4, torch.uint8)
            slot_evid = self._zeros(int(totals[0]) + 4, torch.uint8)
            cand_evid = self._zeros(n_pairs + 8, torch.uint8)
            mark("alloc3")
            self._check(lib.unfz_chain_tally(ctx, d_dnm.data_ptr(), n_dnms, d_seg.data_ptr(), seg_pair_off.data_ptr(), sc, rc_,
                                             rsum.data_ptr(), blk_maxspan.data_ptr(), hits.data_ptr(), mark_prefix.data_ptr(),
                                             het_list.data_ptr(), n_het.data_ptr(), cand_list.data_ptr(), n_cand.data_ptr(),
                                             d_all.data_ptr(), win.data_ptr(), off.data_ptr(),
                                             totals.ctypes.data, C.byref(params), scratch.data_ptr(), nbytes,
                                             slot_label.data_ptr(), slot_evid.data_ptr(), cand_evid.data_ptr(),
                                             tally.data_ptr(), s), "chain_tally")
            launches += 1
            mark("chain_tally")
            dv.update(rsum=rsum, hits=hits, slot_label=slot_label, slot_evid=slot_evid, cand_evid=cand_evid,
                      blk_maxspan=blk_maxspan, row_mark=row_mark, mark_prefix=mark_prefix)
            res.slot_off = h_off[0].copy()
            if download:
                res.win = win.cpu().numpy()[: 4 * n_dnms].reshape(4, n_dnms)
        self._check(lib.unfz_summarize(ctx, d_dnm.data_ptr(), n_dnms, tally.data_ptr(), cnv_dad.data_ptr(), cnv_mom.data_ptr(),
                                       n_cand.data_ptr(), C.byref(params), calls_s.data_ptr(), calls_a.data_ptr(), s), "summarize")
        launches += 1
        mark("summarize")
        # ---- results (a few bytes per DNM) ---------------------------------------------------
        if download:
            res.seg_row_lo = seg_row_lo.cpu().numpy()[:n_segs]
            res.n_het, res.n_cand = n_het.cpu().numpy()[:n_dnms], n_cand.cpu().numpy()[:n_dnms]
            res.cnv_dad, res.cnv_mom = cnv_dad.cpu().numpy()[:n_dnms], cnv_mom.cpu().numpy()[:n_dnms]
            res.tally = tally.cpu().numpy()[: n_dnms * 8].view(L.TALLY_DTYPE)
            res.calls_strict = calls_s.cpu().numpy()[: n_dnms * 4].view(L.CALL_DTYPE)
            res.calls_ambiguous = calls_a.cpu().numpy()[: n_dnms * 4].view(L.CALL_DTYPE)
        else:
            dv.update(n_het=n_het, n_cand=n_cand, tally=tally, calls_s=calls_s, calls_a=calls_a)
        mark("download")
        res.launches = launches
        if time_stages:
            torch.cuda.synchronize(dev)
            for (n0, e0), (n1, e1) in zip(ev[:-1], ev[1:]):
                res.timings_ms[n1] = res.timings_ms.get(n1, 0.0) + e0.elapsed_time(e1)
        if not keep_device:
            dv.clear()
        return res
