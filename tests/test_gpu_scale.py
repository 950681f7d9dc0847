"""GPU: the other BASELINE configs at reduced scale against the oracle, and size-independent
properties at full scale (ground truth of the synthetic trio, determinism, shard invariance,
sortedness of the site lists, --no-extended monotonicity)."""
import copy

import numpy as np
import pytest

from oracle import port
from tests.util import gpu_kwargs, norm_record, run_port
from unfazed_b200 import _lib as L
from unfazed_b200.synth import SynthConfig, make_dataset

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from unfazed_b200.engine import Engine
    return Engine(0)


def _compare(engine, cfg, params):
    from unfazed_b200.phaser import BatchPhaser
    ds = make_dataset(cfg)
    want, _ph = run_port(ds, **params)
    bp = BatchPhaser(engine, ds.sites, ds.reads, ds.pedigrees)
    got = bp.phase(copy.deepcopy(ds.dnms), **gpu_kwargs(**params))
    assert set(got) == set(want)
    for k in want:
        assert norm_record(got[k]) == norm_record(want[k]), k
    return len(want)


def test_config3_sv_set_up_to_1mb(engine):
    """BASELINE config[2]: DEL/DUP/INV up to 1 Mb, allele-balance CNV phasing over interior sites."""
    n = _compare(engine, SynthConfig(dnms_per_trio=24, seed=401, sv_frac=1.0, sv_max_len=1_000_000, coverage=20.0), {})
    assert n > 5


def test_config4_long_range_chaining(engine):
    """BASELINE config[3]: --search-dist 50000, deep coverage (scaled: 4 DNMs, 45x)."""
    n = _compare(engine, SynthConfig(dnms_per_trio=4, seed=402, search_dist=50000, coverage=45.0), dict(search_dist=50000))
    assert n >= 2


def test_config5_cohort_with_sex_chromosomes(engine):
    """BASELINE config[4]: many trios, build 38, chrX/Y autophasing incl. PAR boundaries (scaled)."""
    cfg = SynthConfig(n_trios=12, dnms_per_trio=10, seed=403, sex_chrom_frac=0.3, male_frac=0.5, coverage=16.0)
    ds = make_dataset(cfg)
    # put a few DNMs right on the PAR boundaries the reference uses for --build 38 (Q6)
    kid = ds.sites.trios[0][0]
    ds.pedigrees[kid]["sex"] = "1"
    for start in (60000, 60001, 2699520, 2699521, 154931044, 155260561):
        ds.dnms.append({"chrom": "X", "start": start, "end": start + 1, "kid": kid, "vartype": "POINT",
                        "bam": ds.bam_name(kid), "cram_ref": None})
    from unfazed_b200.phaser import BatchPhaser
    want, _ = run_port(ds, build="38")
    bp = BatchPhaser(engine, ds.sites, ds.reads, ds.pedigrees)
    got = bp.phase(copy.deepcopy(ds.dnms), build="38")
    assert set(got) == set(want)
    for k in want:
        assert norm_record(got[k]) == norm_record(want[k]), k
    auto = [k for k, r in want.items() if r["evidence_type"] == "SEX-CHROM"]
    assert len(auto) >= 3


@pytest.fixture(scope="module")
def big(engine):
    """2 000 DNMs of the headline workload (the bench uses 10 000 of the same shape)."""
    from unfazed_b200.phaser import BatchPhaser
    ds = make_dataset(SynthConfig(dnms_per_trio=2000, seed=404))
    bp = BatchPhaser(engine, ds.sites, ds.reads, ds.pedigrees)
    res, layout = bp.run(ds.dnms, [])
    return ds, bp, res, layout


def test_full_size_calls_agree_with_ground_truth(big):
    ds, bp, res, layout = big
    from unfazed_b200.phaser import dnm_key
    calls = res.calls_strict
    n_called = n_right = 0
    for d in range(*layout["snv"]):
        if calls["emitted"][d] and calls["origin"][d] in (L.ORIGIN_DAD, L.ORIGIN_MOM):
            n_called += 1
            truth = ds.truth[dnm_key(res.plan.entries[d])]
            n_right += (truth == "dad") == (calls["origin"][d] == L.ORIGIN_DAD)
    assert n_called > 0.5 * len(ds.dnms)          # most DNMs have an informative site within reach
    assert n_right >= 0.995 * n_called            # 1 % base errors: wrong calls are essentially absent


def test_full_size_site_lists_sorted_and_consistent(big):
    ds, bp, res, layout = big
    pos = ds.sites.pos
    for d in range(0, len(ds.dnms), 37):
        h = pos[res.het_rows(d)]
        c = pos[res.cand_words(d).view(np.uint32) & 0x3FFFFFFF]
        assert np.all(np.diff(h) >= 0) and np.all(np.diff(c) >= 0)
        assert set(c.tolist()) <= set(h.tolist())            # read mode: every candidate is a het site
        dn = ds.dnms[d]
        assert np.all(np.abs(h.astype(np.int64) - dn["start"]) <= 5001)


def test_full_size_deterministic_and_shard_invariant(big, engine):
    ds, bp, res, layout = big
    res2, _ = bp.run(ds.dnms, [])
    assert np.array_equal(res.tally, res2.tally) and np.array_equal(res.calls_strict, res2.calls_strict)
    half = len(ds.dnms) // 2
    ra, _ = bp.run(ds.dnms[:half], [])
    rb, _ = bp.run(ds.dnms[half:], [])
    assert np.array_equal(np.concatenate([ra.tally, rb.tally]), res.tally)
    assert np.array_equal(np.concatenate([ra.calls_ambiguous, rb.calls_ambiguous]), res.calls_ambiguous)


def test_no_extended_labels_are_a_subset(big):
    ds, bp, res, layout = big
    res_ne, _ = bp.run(ds.dnms[:200], [], no_extended=True)
    for d in range(0, 200, 9):
        seeds = bp.labels(res_ne, d)
        full = bp.labels(res, d)
        for name, hap in seeds.items():
            assert full.get(name) == hap
        assert len(full) >= len(seeds)


def test_more_het_sites_than_the_shared_memory_budget(engine):
    """> 256 het sites per DNM: the chain kernel's per-site state falls back to global scratch."""
    n = _compare(engine, SynthConfig(dnms_per_trio=3, seed=502, search_dist=30000, site_spacing=100, coverage=16.0),
                 dict(search_dist=30000))
    assert n >= 2


def test_long_reads_take_the_chunked_scan_kernel(engine):
    """40 kb reads do not fit a TMA stage: read_scan falls back to the chunked kernel; read summaries
    (reference_end, goodread, filters) and records must still match the oracle."""
    from unfazed_b200.phaser import BatchPhaser
    cfg = SynthConfig(dnms_per_trio=4, seed=501, readlen=40000, frag_mean=90000.0, frag_sd=3000.0, search_dist=3000,
                      coverage=40.0, read_margin=100000, noise=False)
    ds = make_dataset(cfg)
    # sprinkle low qualities / bad flags so that the filters have something to reject
    rng = np.random.default_rng(5)
    q = ds.reads.qual
    for r in range(0, ds.reads.n_reads, 7):
        q0 = ds.reads.qoff(r)
        q[q0 + rng.integers(0, 40000, size=int(rng.integers(0, 25)))] = 5
    ds.reads.hdr["mapq"][::11] = 0
    params = dict(readlen=40001)
    want, _ = run_port(ds, **params)
    bp = BatchPhaser(engine, ds.sites, ds.reads, ds.pedigrees)
    res, layout = bp.run(ds.dnms, [], **gpu_kwargs(**params))
    got = bp.records(res, layout)
    assert set(got) == set(want)
    for k in want:
        assert norm_record(got[k]) == norm_record(want[k]), k
    rs = res.read_summaries()
    bam = port.Bam(ds.reads, 0)
    p = port.Params(readlen=40001)
    assert np.array_equal(rs["end"], ds.reads.ref_ends().astype(np.int32))
    for r in range(0, ds.reads.n_reads, 3):
        assert bool(rs["flags"][r] & L.RS_GOOD_CONC) == port.goodread(bam, r, p), r
        assert bool(rs["flags"][r] & L.RS_GOOD_DISC) == port.goodread(bam, r, p, True), r
