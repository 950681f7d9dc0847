"""CPU: the host reshaping after the kernels, driven by a hand-made batch result (no device): BatchPhaser.records
(native builder over the device's evidence lists) against BatchPhaser.compact(...).to_records() (what a rank ships
to rank 0 in a multi-GPU run), with synthesized and with explicit read names, autophased entries and entries
without a record in between; plus the bench's whole-workload oracle run."""
import copy
import pickle
import types

import numpy as np
import pytest

from unfazed_b200 import _lib as L
from unfazed_b200.phaser import BatchPhaser, CompactRecords, _cut
from unfazed_b200.synth import SynthConfig, make_dataset


def fake_result(ds, rng, n):
    dnms = ds.dnms[:n]
    flags = np.zeros(n, dtype=np.int32)
    auto = rng.random(n) < 0.15
    flags[auto] = L.DNM_AUTOPHASE
    has = (rng.random(n) < 0.6) & ~auto
    cnt = np.stack([rng.integers(0, 25, n), rng.integers(0, 25, n), rng.integers(0, 4, n), rng.integers(0, 4, n)])
    cnt[:, ~has] = 0                                   # entries without a record own no evidence
    off = np.zeros((4, n + 1), dtype=np.int64)
    off[:, 1:] = np.cumsum(cnt, axis=1)
    ev = {"off": off,
          "read_dad": rng.integers(0, ds.reads.n_reads, off[0, n]).astype(np.int32),
          "read_mom": rng.integers(0, ds.reads.n_reads, off[1, n]).astype(np.int32),
          "pos_dad": rng.integers(1, 10 ** 8, off[2, n]).astype(np.int32),
          "pos_mom": rng.integers(1, 10 ** 8, off[3, n]).astype(np.int32)}
    dnm = np.zeros(n, dtype=L.DNM_DTYPE)
    dnm["flags"] = flags
    tally = np.zeros(n, dtype=L.TALLY_DTYPE)
    tally["has_record"] = has
    plan = types.SimpleNamespace(dnm=dnm, entries=dnms)
    res = types.SimpleNamespace(plan=plan, ev=ev, tally=tally)
    layout = {"sv_cnv": (0, 0), "sv_read": (0, 0), "snv": (0, n)}
    return res, layout, int(has.sum() + auto.sum())


@pytest.mark.parametrize("with_names", [False, True])
def test_compact_equals_records_on_a_hand_made_result(with_names):
    ds = make_dataset(SynthConfig(dnms_per_trio=60, n_trios=2, seed=21, coverage=3.0))
    if with_names:
        ds.reads.names = ["r%d/%d" % (i // 2, i) for i in range(ds.reads.n_reads)]
    rng = np.random.default_rng(5)
    res, layout, n_live = fake_result(ds, rng, len(ds.dnms))
    bp = BatchPhaser(None, ds.sites, ds.reads, ds.pedigrees)
    want = bp.records(res, layout)
    c = bp.compact(res, layout)
    assert isinstance(c, CompactRecords) and len(c) == n_live == len(want)
    got = pickle.loads(pickle.dumps(c)).to_records()
    assert got == want and list(got) == list(want)
    auto = [r for r in want.values() if r["evidence_type"] == "SEX-CHROM"]
    assert auto and all(r["dad_reads"] == [] and r["cnv_dad_sites"] == "NA" for r in auto)
    # a batch with SV entries is not compacted
    assert bp.compact(res, dict(layout, sv_read=(0, 1))) is None
    res.ev = None
    assert bp.compact(res, layout) is None


def test_cut_keeps_the_slices_of_the_live_entries():
    off = np.array([0, 3, 3, 7, 9], dtype=np.int64)
    flat = np.arange(9)
    sub, new_off = _cut(flat, off, np.array([0, 2, 3]))
    assert sub.tolist() == [0, 1, 2, 3, 4, 5, 6, 7, 8] and new_off.tolist() == [0, 3, 7, 9]     # entry 1 is empty
    sub, new_off = _cut(flat, off, np.array([2]))
    assert sub.tolist() == [3, 4, 5, 6] and new_off.tolist() == [0, 4]
    sub, new_off = _cut(flat, off, np.array([], dtype=np.int64))
    assert sub.tolist() == [] and new_off.tolist() == [0]


def test_bench_whole_workload_oracle_equals_one_pass():
    """bench.py --full-parity: the port sharded over the host cores returns what one Phaser returns (find and find_many)."""
    import sys
    from oracle import port
    argv = sys.argv
    try:
        sys.argv = ["bench.py", "--config", "c2_many"]
        import bench
        args = bench.parse()
    finally:
        sys.argv = argv
    ds = make_dataset(SynthConfig(noise=True, seed=3, dnms_per_trio=16, n_trios=2, search_dist=5000, coverage=20.0))
    saved = bench.CONFIGS["c2_many"]["run"]["multiread_proc_min"]
    try:
        for mpm in (8, 10 ** 9):
            bench.CONFIGS["c2_many"]["run"]["multiread_proc_min"] = mpm
            got = bench.port_all(ds, args)
            want = port.Phaser(ds.sites, ds.reads, ds.pedigrees, bench.port_params(args)).phase(copy.deepcopy(ds.dnms))
            assert got == want and len(want) > 5
    finally:
        bench.CONFIGS["c2_many"]["run"]["multiread_proc_min"] = saved
