"""GPU parity: the CUDA path (through the C ABI) against the oracle port on seeded synthetic trios.

Bit-exact bar: candidate/het site lists incl. duplicates, read->haplotype labels, per-parent site
and read-name sets, evidence counts and calls (SURVEY.md 8(c) parity definition)."""
import copy

import numpy as np
import pytest

from oracle import port
from unfazed_b200 import _lib as L
from unfazed_b200.synth import SynthConfig, make_dataset
from tests.util import gpu_kwargs, norm_record, port_params, run_port, summarize_all

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from unfazed_b200.engine import Engine
    return Engine(0)


CASES = [
    ("clean", SynthConfig(dnms_per_trio=40, seed=21, noise=False), {}),
    ("noisy", SynthConfig(dnms_per_trio=60, seed=22), {}),
    ("find_many", SynthConfig(dnms_per_trio=40, seed=23), dict(multiread_proc_min=1)),
    ("clustered_indel", SynthConfig(dnms_per_trio=50, seed=24, cluster_frac=0.4, indel_frac=0.25), {}),
    ("two_trios_many", SynthConfig(dnms_per_trio=30, seed=25, n_trios=2, cluster_frac=0.3, indel_frac=0.2),
     dict(multiread_proc_min=1, threads=2)),
    ("no_extended", SynthConfig(dnms_per_trio=40, seed=26), dict(no_extended=True)),
    ("sex_chr_prefix", SynthConfig(dnms_per_trio=40, seed=27, sex_chrom_frac=0.4, male_frac=1.0, chr_prefix="chr"), {}),
    ("prefix_mismatch_many", SynthConfig(dnms_per_trio=30, seed=28, chr_prefix="chr", dnm_chr_prefix=""),
     dict(multiread_proc_min=1)),
    ("prefix_mismatch_find", SynthConfig(dnms_per_trio=30, seed=29, chr_prefix="chr", dnm_chr_prefix=""), {}),
    ("thresholds", SynthConfig(dnms_per_trio=40, seed=30),
     dict(min_gt_qual=30, min_depth=25, ab_het=(0.3, 0.7), ab_homref=(0.0, 0.1), min_map_qual=20, readlen=150)),
    ("deep_long_range", SynthConfig(dnms_per_trio=6, seed=31, search_dist=20000, coverage=45.0), dict(search_dist=20000)),
    ("svs_cnv", SynthConfig(dnms_per_trio=40, seed=32, sv_frac=0.6, sv_max_len=100000, indel_frac=0.1), {}),
]


@pytest.mark.parametrize("name,cfg,params", CASES, ids=[c[0] for c in CASES])
def test_records_labels_and_sites(engine, name, cfg, params):
    from unfazed_b200.phaser import BatchPhaser, dnm_key
    ds = make_dataset(cfg)
    want, ph = run_port(ds, **params)
    bp = BatchPhaser(engine, ds.sites, ds.reads, ds.pedigrees)
    kids = set(ds.pedigrees)
    svs = [d for d in ds.dnms if d["vartype"].upper() in port.SV_TYPES and d["kid"] in kids]
    snvs = [d for d in ds.dnms if d["vartype"].upper() in port.SNV_TYPES and d["kid"] in kids]
    res, layout = bp.run(snvs, svs, **gpu_kwargs(**params))
    got = bp.records(res, layout)
    assert not res.tally["status"].any(), "capacity problems reported by the chain kernel"

    # site lists of the SNV read-mode entries against port.find
    p = port_params(**params)
    if snvs:
        ann = port.find(copy.deepcopy(snvs), ds.pedigrees, ds.sites, p, p.search_dist, whole_region=False)
        by_key = {port._key(d): d for d in ann}
        a, b = layout["snv"]
        for d in range(a, b):
            dn = res.plan.entries[d]
            w = by_key[dnm_key(dn)]
            if res.plan.trio[d] < 0:
                assert "candidate_sites" not in w
                continue
            cands, hets = bp.site_dicts(res, d, int(res.plan.trio[d]), False)
            assert cands == w.get("candidate_sites", []), (name, dnm_key(dn))
            assert hets == w.get("het_sites", []), (name, dnm_key(dn))

    # read -> haplotype labels
    for key, lab in ph.labels.items():
        ds_ = [d for d in range(*layout["snv"]) if dnm_key(res.plan.entries[d]) == key]
        if not ds_:
            continue
        mine = bp.labels(res, ds_[0])
        assert mine == lab, (name, key)

    # records and calls
    assert set(got) == set(want), (name, sorted(set(got) ^ set(want))[:5])
    for k in want:
        assert norm_record(got[k]) == norm_record(want[k]), (name, k)
    for amb in (True, False):
        sw = summarize_all(want, amb)
        sg = summarize_all(got, amb)
        for k in sw:
            a_, b_ = sw[k], sg[k]
            if a_ is None or b_ is None:
                assert a_ is None and b_ is None
                continue
            for f in ("origin_parent", "other_parent", "evidence_count", "evidence_types", "origin_parent_sites", "other_parent_sites"):
                assert a_[f] == b_[f], (name, k, f)


def test_device_calls_match_summarize_record(engine):
    """unfz_summarize (device) == summarize_record on the records, strict and ambiguous."""
    from unfazed_b200.phaser import BatchPhaser, dnm_key
    ds = make_dataset(SynthConfig(dnms_per_trio=60, seed=41, sv_frac=0.4, sv_max_len=60000, sex_chrom_frac=0.2, male_frac=1.0))
    bp = BatchPhaser(engine, ds.sites, ds.reads, ds.pedigrees)
    kids = set(ds.pedigrees)
    svs = [d for d in ds.dnms if d["vartype"].upper() in port.SV_TYPES]
    snvs = [d for d in ds.dnms if d["vartype"].upper() in port.SNV_TYPES]
    res, layout = bp.run(snvs, svs)
    recs = bp.records(res, layout)
    names = {L.ORIGIN_NONE: None}
    for rng_name in ("sv_read", "snv"):
        for d in range(*layout[rng_name]):
            dn = res.plan.entries[d]
            key = dnm_key(dn)
            dad, mom = ds.pedigrees[dn["kid"]]["dad"], ds.pedigrees[dn["kid"]]["mom"]
            for amb, calls in ((False, res.calls_strict), (True, res.calls_ambiguous)):
                want = port.summarize_record(copy.deepcopy(recs[key]), amb, False, 10) if key in recs else None
                c = calls[d]
                if want is None:
                    assert c["emitted"] == 0, (key, amb)
                    continue
                assert c["emitted"] == 1, (key, amb)
                origin = {L.ORIGIN_NONE: None, L.ORIGIN_DAD: dad, L.ORIGIN_MOM: mom, L.ORIGIN_BOTH: dad + "|" + mom}[int(c["origin"])]
                assert origin == want["origin_parent"], (key, amb)
                assert int(c["evidence_count"]) == want["evidence_count"], (key, amb)
                types = set(want["evidence_types"])
                bits = {"READBACKED": 1, "ALLELE-BALANCE": 2, "AMBIGUOUS_READBACKED": 4, "AMBIGUOUS_ALLELE-BALANCE": 8,
                        "AMBIGUOUS_BOTH": 16, "SEX-CHROM": 32}
                assert int(c["evidence_types"]) == sum(bits[t] for t in types), (key, amb)


def test_hit_words_match_the_oracle_lookup(engine):
    """Hit words written by unfz_read_site_alleles == get_reference_positions().index(pos) + base +
    (quality < --min-gt-qual) computed by the oracle's decoder, for a sample of (read, marked site) overlaps."""
    from unfazed_b200.phaser import BatchPhaser
    ds = make_dataset(SynthConfig(dnms_per_trio=40, seed=51, indel_frac=0.2))
    bp = BatchPhaser(engine, ds.sites, ds.reads, ds.pedigrees)
    res, _ = bp.run(ds.dnms, [])
    rs = res.read_summaries()
    hits = res._np("hits").view(np.uint32)
    tile_base = res._np("tile_base").view(np.uint32)
    tile_reads = res._dev["tile_reads"]
    mark = res._np("row_mark")
    pos = ds.sites.pos
    bam = port.Bam(ds.reads, 0)
    idx = np.nonzero(rs["cnt"])[0]
    assert idx.size > 500
    checked = 0
    for r in idx[:: max(1, idx.size // 400)]:
        cig, rp, seq, quals = bam.decoded(int(r))
        st, en = int(ds.reads.hdr["start"][r]), int(rs["end"][r])
        rb = int(np.searchsorted(ds.reads.blk_off, r, side="right")) - 1
        sb = ds.sites.block_of(0, ds.reads.contigs[int(ds.reads.blk_contig[rb])])
        a, b = int(ds.sites.blk_off[sb]), int(ds.sites.blk_off[sb + 1])
        rows = [j for j in range(a + int(np.searchsorted(pos[a:b], st)), a + int(np.searchsorted(pos[a:b], en))) if mark[j]]
        assert len(rows) == rs["cnt"][r]
        base = int(tile_base[r // tile_reads]) + int(rs["hoff"][r])
        for k, j in enumerate(rows):
            w = int(hits[base + k])
            p = int(pos[j])
            if p in rp:
                q = rp.index(p)
                assert (w & 0xFFFF) == q + 1
                assert bool(w & (1 << 16)) == (quals[q] < 20)          # the device holds the comparison, not the byte
                ch = "N" if (w >> 16) & 0x80 and ((w >> 24) & 3) == 0 else ("?" if (w >> 16) & 0x80 else "ACGT"[(w >> 24) & 3])
                assert ch == seq[q]
                assert bool(w & (1 << 26)) == (q + 1 < len(seq))
            else:
                assert (w & 0xFFFF) == 0
            checked += 1
    assert checked > 300


@pytest.mark.parametrize("no_extended", [False, True])
def test_collect_reads_drop_in_matches_oracle(engine, no_extended):
    """read_collector.collect_reads_snv / collect_reads_sv (reference signatures, GPU inside) return the
    same haplotype-grouped reads as the oracle's per-variant functions."""
    from unfazed_b200 import datasource, read_collector
    ds = make_dataset(SynthConfig(dnms_per_trio=24, seed=61, indel_frac=0.2, sv_frac=0.3, sv_max_len=20000))
    p = port.Params(no_extended=no_extended)
    bam_name = ds.bam_name("kid0")
    datasource._registry.clear()
    datasource.register_tables(bam_name, reads=ds.reads)
    ann = port.find(copy.deepcopy(ds.dnms), ds.pedigrees, ds.sites, p, p.search_dist, whole_region=False)
    bam = port.Bam(ds.reads, 0)
    n_checked = 0
    for dn in ann:
        if "het_sites" not in dn:
            continue
        region = {"chrom": dn["chrom"], "start": dn["start"], "end": dn["end"]}
        if dn["vartype"] in port.SV_TYPES:
            want, cul, _ = port.collect_reads_sv(bam, region, dn["het_sites"], None, p)
            got, cul2 = read_collector.collect_reads_sv(bam_name, region, dn["het_sites"], None, no_extended, None,
                                                        p.insert_size_max_sample, p.stdevs, p.min_map_qual, p.min_gt_qual,
                                                        p.readlen, p.split_error_margin)
        else:
            ref, alts = port.get_refalt(ds.sites, 0, dn["chrom"], dn["start"], "")
            if len(alts) != 1:
                continue
            want, cul, _ = port.collect_reads_snv(bam, region, dn["het_sites"], ref, alts[0], None, p)
            got, cul2 = read_collector.collect_reads_snv(bam_name, region, dn["het_sites"], ref, alts[0], None, no_extended,
                                                         None, p.insert_size_max_sample, p.stdevs, p.min_map_qual,
                                                         p.min_gt_qual, p.readlen, p.split_error_margin)
        assert float(cul) == float(cul2)
        for hap in ("ref", "alt"):
            w = sorted({(bam.name(r), int(ds.reads.hdr["start"][r])) for r in want[hap]})
            g = sorted({(r.query_name, r.reference_start) for r in got[hap]})
            assert g == w, (port._key(dn), hap)
        n_checked += 1
    assert n_checked >= 12


def test_device_insert_size_estimate_is_bit_identical(engine):
    """estimate_concordant_insert_len via unfz_insert_size_order_stats == the host numpy percentile the
    reference computes, for several sample caps (incl. caps that cut inside a read block)."""
    from unfazed_b200.plan import concordant_upper_lens
    ds = make_dataset(SynthConfig(n_trios=2, dnms_per_trio=12, seed=71, coverage=12.0))
    dreads = engine.upload_reads(ds.reads)
    for cap in (1000000, 5000, 777, 10, 0):
        want = concordant_upper_lens(ds.reads, 151, cap, 3)
        got = engine.concordant_upper_lens(dreads, 151, cap, 3)
        assert np.array_equal(want, got), cap
    bam = port.Bam(ds.reads, 1)
    assert float(port.estimate_concordant_insert_len(bam, port.Params())) == float(engine.concordant_upper_lens(dreads, 151, 1000000, 3)[-1])
