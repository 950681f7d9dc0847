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


def test_long_reads_are_counted_out_of_global_memory(engine):
    """The quality bits of 32 reads of 40 kb do not fit a warp's shared-memory slice: read_scan counts them
    straight out of global memory; read summaries (reference_end, goodread, filters) and records must
    still match the oracle."""
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


def test_speculative_sizing_equals_the_exact_path_and_survives_an_overflow():
    """Engine.run sizes its variable buffers from the previous batch (device-side capacity check, guarded
    kernels).  A batch that fits, a batch that overflows every capacity (re-run on the exact path) and a
    batch after the capacities grew must all give the results of the exact, two-sync path."""
    from unfazed_b200.engine import Engine
    from unfazed_b200.phaser import BatchPhaser
    eng = Engine(0)
    ds = make_dataset(SynthConfig(dnms_per_trio=120, seed=909, coverage=20.0))
    bp = BatchPhaser(eng, ds.sites, ds.reads, ds.pedigrees)
    small, large = ds.dnms[:10], ds.dnms

    def fields(res):
        return (res.n_pairs, res.n_hits, res.tally.copy(), res.calls_strict.copy(), res.calls_ambiguous.copy(),
                res.seg_pair_off.copy(), res.slot_off.copy(), res.n_het.copy(), res.n_cand.copy())

    def same(a, b):
        return all(np.array_equal(x, y) for x, y in zip(a, b))

    eng._caps = None
    exact_small = fields(bp.run(small, [])[0])            # first batch: exact path, capacities recorded
    caps_small = dict(eng._caps)
    exact_large_ref = None
    spec_small = fields(bp.run(small, [])[0])             # fits: speculative
    assert same(exact_small, spec_small)
    over = fields(bp.run(large, [])[0])                   # exceeds every capacity: guard + exact re-run
    assert eng._caps["pairs"] > caps_small["pairs"] and eng._caps["hits"] > caps_small["hits"]
    eng._caps = None
    exact_large_ref = fields(bp.run(large, [])[0])
    assert same(over, exact_large_ref)
    spec_large = fields(bp.run(large, [])[0])             # now speculative with the grown capacities
    assert same(spec_large, exact_large_ref)
    spec_small2 = fields(bp.run(small, [])[0])            # a smaller batch inside large capacities
    assert same(spec_small2, exact_small)
    # labels / evidence accessors work on speculatively sized buffers
    res, _ = bp.run(large, [])
    res_exact = None
    eng._caps = None
    res_exact, _ = bp.run(large, [])
    for d in range(0, len(large), 17):
        assert bp.labels(res, d) == bp.labels(res_exact, d)


def test_phase_stream_equals_one_call_per_batch(engine):
    """BatchPhaser.phase_stream (device work of batch k+1 queued before the records of batch k are built, host
    columns re-uploaded for every batch) yields exactly what phase() returns for each batch, in order."""
    from unfazed_b200.phaser import BatchPhaser
    ds = make_dataset(SynthConfig(dnms_per_trio=90, seed=606, indel_frac=0.1, sv_frac=0.2, sv_max_len=30000, coverage=20.0))
    batches = [ds.dnms[:30], ds.dnms[30:45], ds.dnms[45:]]
    packed = engine.pack_reads(ds.reads, min_gt_qual=20, pin=True)
    bp = BatchPhaser(engine, ds.sites, packed, ds.pedigrees, resident=False)
    want = [bp.phase(copy.deepcopy(b)) for b in batches]
    got = list(bp.phase_stream([copy.deepcopy(b) for b in batches]))
    assert len(got) == 3 and sum(len(w) for w in want) > 10
    for g, w in zip(got, want):
        assert set(g) == set(w)
        for k in w:
            assert norm_record(g[k]) == norm_record(w[k]), k
    # and both equal the oracle
    ref, _ = run_port(ds)
    merged = {}
    for g in got:
        merged.update(g)
    assert set(merged) == set(ref)
    for k in ref:
        assert norm_record(merged[k]) == norm_record(ref[k]), k


def test_compact_records_equal_the_record_dicts(engine):
    """BatchPhaser.phase(compact=True) (what a rank ships to rank 0 in a multi-GPU run) survives a pickle round trip and
    ``to_records()`` gives exactly the dict ``phase()`` returns, autophased entries included; a batch with SVs falls
    back to the dict itself."""
    import pickle
    from unfazed_b200.phaser import BatchPhaser, CompactRecords
    cfg = SynthConfig(n_trios=6, dnms_per_trio=12, seed=411, sex_chrom_frac=0.3, male_frac=0.5, coverage=16.0, indel_frac=0.2)
    ds = make_dataset(cfg)
    bp = BatchPhaser(engine, ds.sites, ds.reads, ds.pedigrees)
    want = bp.phase(copy.deepcopy(ds.dnms), build="38")
    c = bp.phase(copy.deepcopy(ds.dnms), build="38", compact=True)
    assert isinstance(c, CompactRecords) and len(c) == len(want) > 10
    assert any(r["evidence_type"] == "SEX-CHROM" for r in want.values())
    got = pickle.loads(pickle.dumps(c)).to_records()
    assert got == want and list(got) == list(want)
    for got_s in bp.phase_stream([copy.deepcopy(ds.dnms)] * 2, build="38", compact=True):
        assert got_s.to_records() == want
    ds2 = make_dataset(SynthConfig(dnms_per_trio=20, seed=412, sv_frac=0.5, sv_max_len=20000, coverage=16.0))
    bp2 = BatchPhaser(engine, ds2.sites, ds2.reads, ds2.pedigrees)
    assert isinstance(bp2.phase(copy.deepcopy(ds2.dnms), compact=True), dict)


@pytest.mark.parametrize("shape", ["wide", "small"])
def test_both_chaining_cta_shapes_give_the_oracle_records(shape, monkeypatch):
    """The chaining kernels exist in two CTA shapes (128 threads for ordinary windows, 512 for deep / wide ones,
    chosen per batch from the incidence count).  Forced either way, a small-window batch with indels and SVs, a
    > 128-het-site window and a 50 kb window all reproduce the oracle's records -- first call (exact sizes) and
    second call (speculative sizes, CUDA graph)."""
    from unfazed_b200.engine import Engine
    from unfazed_b200.phaser import BatchPhaser
    monkeypatch.setenv("UNFZ_CHAIN_SHAPE", shape)
    eng = Engine(0)                       # its own context: the batch graphs of the shared engine were captured with the default choice
    cases = [
        (SynthConfig(dnms_per_trio=40, seed=701, indel_frac=0.2, sv_frac=0.2, sv_max_len=20000, coverage=20.0), {}),
        (SynthConfig(dnms_per_trio=4, seed=702, search_dist=50000, coverage=45.0), dict(search_dist=50000)),
        (SynthConfig(dnms_per_trio=3, seed=703, search_dist=30000, coverage=12.0, site_spacing=150), dict(search_dist=30000)),
    ]
    for cfg, params in cases:
        ds = make_dataset(cfg)
        want, _ = run_port(ds, **params)
        bp = BatchPhaser(eng, ds.sites, ds.reads, ds.pedigrees)
        for _ in range(2):
            got = bp.phase(copy.deepcopy(ds.dnms), **gpu_kwargs(**params))
            assert set(got) == set(want) and len(want) > 0
            for k in want:
                assert norm_record(got[k]) == norm_record(want[k]), (shape, cfg.seed, k)
