"""CPU: host window planner.  (1) the vectorised planner is identical to the per-DNM one;
(2) the planned windows + the oracle's per-row classifier reproduce the oracle's ``find`` (which has
its own, literal membership logic) -- i.e. the planner encodes find()/find_many() membership."""
import copy

import numpy as np
import pytest

from oracle import port
from unfazed_b200 import _lib as L
from unfazed_b200.plan import SiteIndex, plan_find, plan_find_fast
from unfazed_b200.synth import SynthConfig, make_dataset

CFGS = [
    SynthConfig(dnms_per_trio=40, seed=301, coverage=4.0),
    SynthConfig(dnms_per_trio=40, seed=302, coverage=4.0, cluster_frac=0.4, indel_frac=0.3, sv_frac=0.3, sv_max_len=30000),
    SynthConfig(dnms_per_trio=25, seed=303, coverage=4.0, n_trios=3, sex_chrom_frac=0.3, male_frac=1.0, chr_prefix="chr", dnm_chr_prefix=""),
]


@pytest.mark.parametrize("cfg", CFGS, ids=[str(c.seed) for c in CFGS])
@pytest.mark.parametrize("whole_region,mpm", [(False, 10 ** 9), (True, 10 ** 9), (False, 1), (True, 1)],
                         ids=["reads", "cnv", "reads_many", "cnv_many"])
def test_fast_planner_equals_generic(cfg, whole_region, mpm):
    ds = make_dataset(cfg)
    sidx = SiteIndex(ds.sites)
    kw = dict(search_dist=0 if whole_region else 5000, whole_region=whole_region, build="38", multiread_proc_min=mpm,
              threads=2, with_reads=not whole_region, first_entry=7, alleles_base=11, sv_quirk=not whole_region)
    a = plan_find(ds.dnms, ds.pedigrees, sidx, ds.reads, **kw)
    b = plan_find_fast(ds.dnms, ds.pedigrees, sidx, ds.reads, **kw)
    assert np.array_equal(a.seg, b.seg)
    for f in ("pos", "end", "rblk", "kind", "seg_lo", "seg_hi", "ref_len", "alt_len", "cnv_entry", "flags"):
        assert np.array_equal(a.dnm[f], b.dnm[f]), f
    assert np.array_equal(a.trio, b.trio) and np.array_equal(a.found, b.found) and a.rblk_sblk == b.rblk_sblk
    for i in range(len(ds.dnms)):           # the allele bytes, wherever they were put in the blob
        for off, ln in (("ref_off", "ref_len"), ("alt_off", "alt_len")):
            xa = a.alleles[a.dnm[off][i] - 11: a.dnm[off][i] - 11 + a.dnm[ln][i]]
            xb = b.alleles[b.dnm[off][i] - 11: b.dnm[off][i] - 11 + b.dnm[ln][i]]
            assert xa.tobytes() == xb.tobytes()


@pytest.mark.parametrize("cfg", CFGS, ids=[str(c.seed) for c in CFGS])
@pytest.mark.parametrize("whole_region,mpm", [(False, 10 ** 9), (False, 1), (True, 10 ** 9)])
def test_planned_windows_reproduce_find_membership(cfg, whole_region, mpm):
    ds = make_dataset(cfg)
    sd = 0 if whole_region else 5000
    p = port.Params(multiread_proc_min=mpm, threads=2)
    want = port.find(copy.deepcopy(ds.dnms), ds.pedigrees, ds.sites, p, sd, whole_region)
    by_key = {port._key(d): d for d in want}
    sidx = SiteIndex(ds.sites)
    plan = plan_find(ds.dnms, ds.pedigrees, sidx, None, search_dist=sd, whole_region=whole_region, build="38",
                     multiread_proc_min=mpm, threads=2, with_reads=False)
    s = ds.sites
    for d, dn in enumerate(ds.dnms):
        cands, hets = [], []
        for k in range(plan.dnm["seg_lo"][d], plan.dnm["seg_hi"][d]):
            sg = plan.seg[k]
            if sg["sblk"] < 0:
                continue
            a, b = int(s.blk_off[sg["sblk"]]), int(s.blk_off[sg["sblk"] + 1])
            lo = a + int(np.searchsorted(s.pos[a:b], sg["lo_pos"], "left"))
            hi = a + int(np.searchsorted(s.pos[a:b], sg["hi_pos"], "right"))
            for row in range(lo, hi):
                if not (s.flag[row] & 1):
                    continue
                dad, mom = ds.pedigrees[dn["kid"]]["dad"], ds.pedigrees[dn["kid"]]["mom"]
                het, cand = port.classify_row(s, row, dn, dad, mom, p, whole_region)
                hets += [het] * int(sg["mult"]) if het else []
                cands += [cand] * int(sg["mult"]) if cand else []
        w = by_key[port._key(dn)]
        assert cands == w.get("candidate_sites", []), port._key(dn)
        assert hets == w.get("het_sites", []), port._key(dn)
