"""CPU, build container only: the per-variant interface mirror `unfazed_b200.site_searcher` against the
reference's `unfazed/site_searcher.py` on random inputs (the order pivot / right run / left run of
`binary_search`, Q16, and the parent-consistency filter of `match_informative_sites`), and
`informative_site_finder.autophaseable` against the reference's PAR tables."""
import random
import types

import pytest

from oracle import ref_driver
from unfazed_b200 import site_searcher as new

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_driver.available(), reason="reference checkout not mounted")]


def _sites(rng, n):
    pos = sorted(rng.sample(range(0, 3000), n)) if n else []
    if n > 3 and rng.random() < 0.5:                       # duplicate positions (overlapping windows, Q9)
        pos = sorted(pos + [pos[n // 2], pos[n // 3]])
    return [{"pos": p, "ref_parent": rng.choice(["dad", "mom"]), "alt_parent": rng.choice(["dad", "mom"]),
             "ref_allele": "A", "alt_allele": "C"} for p in pos]


def test_binary_search_random():
    ref = ref_driver.modules()["site_searcher"]
    rng = random.Random(7)
    for _ in range(3000):
        sites = _sites(rng, rng.randint(0, 40))
        a = rng.randint(-50, 3050)
        b = a + rng.choice([0, 1, 2, 30, 151, 400, 2000])
        assert new.binary_search(a, b, sites) == ref.binary_search(a, b, sites), (a, b, [s["pos"] for s in sites])


def test_match_informative_sites_random():
    ref = ref_driver.modules()["site_searcher"]
    rng = random.Random(11)
    for _ in range(300):
        sites = _sites(rng, rng.randint(1, 30))
        reads = {}
        for hap in ("ref", "alt"):
            lst = []
            for _r in range(rng.randint(0, 12)):
                st = rng.randint(0, 2900)
                lst.append(types.SimpleNamespace(reference_start=st, reference_end=st + rng.choice([1, 50, 151, 600])))
            reads[hap] = lst
        assert new.match_informative_sites(reads, sites) == ref.match_informative_sites(reads, sites)


def test_autophaseable_matches_reference():
    ref = ref_driver.modules()["informative_site_finder"]
    from unfazed_b200.informative_site_finder import autophaseable
    peds = {"boy": {"kid": "boy", "dad": "d", "mom": "m", "sex": "1"}, "girl": {"kid": "girl", "dad": "d", "mom": "m", "sex": "2"}}
    rng = random.Random(5)
    edges = [10000, 10001, 60001, 2699520, 2699521, 2781479, 2781480, 154931044, 155260560, 155701383, 156030895,
             57227415, 59034050, 59363566, 56887903]
    for build in ("37", "38", 37, 38, "19"):
        for chrom in ("X", "Y", "chrX", "chrY", "x", "7", "chr7"):
            for kid in ("boy", "girl"):
                for start in edges + [rng.randint(1, 160000000) for _ in range(6)]:
                    dn = {"chrom": chrom, "start": start, "end": start + 1, "kid": kid, "vartype": "POINT"}
                    try:
                        want = ref.autophaseable(dict(dn), peds, build)
                    except Exception as e:                # e.g. an unknown build in the reference's tables
                        with pytest.raises(type(e)):
                            autophaseable(dict(dn), peds, build)
                        continue
                    assert bool(autophaseable(dict(dn), peds, build)) == bool(want), (build, chrom, kid, start)


def test_phase_by_reads_and_phase_by_snvs_random():
    """The per-variant evidence functions of the drop-in phasers against the reference's (truth table
    snv_phaser.py:52-69, CNV votes sv_phaser.py:71-85) on random matches."""
    mods = ref_driver.modules()
    from unfazed_b200 import snv_phaser as new_snv, sv_phaser as new_sv
    rng = random.Random(13)

    class Read:
        def __init__(self, start, seq, gaps):
            self.reference_start, self.query_sequence, self._gaps = start, seq, gaps

        def get_reference_positions(self, full_length=False):
            out, p = [], self.reference_start
            for i in range(len(self.query_sequence)):
                if i in self._gaps:
                    out.append(None)
                else:
                    out.append(p)
                    p += 1
            return out if full_length else [x for x in out if x is not None]

    for _ in range(200):
        sites = _sites(rng, rng.randint(1, 12))
        for s in sites:
            s["ref_allele"], s["alt_allele"] = rng.sample("ACGT", 2)
            s["ref_parent"], s["alt_parent"] = rng.sample(["dad", "mom"], 2)      # a trio: the two parents, either way round
            s["kid_allele"] = rng.choice(["ref_parent", "alt_parent"])
        matches = {}
        for hap in rng.choice([("ref", "alt"), ("alt", "ref")]):
            lst = []
            for _r in range(rng.randint(0, 6)):
                st = rng.randint(0, 2900)
                rd = Read(st, "".join(rng.choice("ACGTN") for _ in range(120)), set(rng.sample(range(120), rng.randint(0, 6))))
                lst.append({"read": rd, "matches": rng.sample(sites, rng.randint(0, len(sites)))})
            matches[hap] = lst
        for ref_mod in (mods["snv_phaser"], mods["sv_phaser"]):
            assert new_snv.phase_by_reads(matches) == ref_mod.phase_by_reads(matches)
        assert new_sv.phase_by_snvs(sites) == mods["sv_phaser"].phase_by_snvs(sites)


def test_autophase_and_get_refalt_match_the_reference():
    """The remaining per-variant callables of the phaser modules: ``autophase`` (snv_phaser.py:302-352 returns True and
    writes the SEX-CHROM record, sv_phaser.py:304-354 writes it and returns None) and ``get_refalt``
    (snv_phaser.py:73-84) over the in-memory cyvcf2 stand-in."""
    import copy
    mods = ref_driver.modules()
    from oracle import fakes
    from unfazed_b200 import snv_phaser as new_snv, sv_phaser as new_sv
    from unfazed_b200.synth import SynthConfig, make_dataset
    peds = {"boy": {"kid": "boy", "dad": "d", "mom": "m", "sex": "1"}, "girl": {"kid": "girl", "dad": "d", "mom": "m", "sex": "2"}}
    rng = random.Random(17)
    starts = [10000, 10001, 60001, 2699520, 2699521, 2781479, 154931044, 155260560, 156030895, 59363566] + \
             [rng.randint(1, 160000000) for _ in range(10)]
    for build in ("37", "38", "19", 38):
        for chrom in ("X", "chrY", "y", "7"):
            for kid in ("boy", "girl"):
                for start in starts:
                    dn = {"chrom": chrom, "start": start, "end": start + 1, "kid": kid, "vartype": rng.choice(["POINT", "DEL"])}
                    for new_mod, ref_mod in ((new_snv, mods["snv_phaser"]), (new_sv, mods["sv_phaser"])):
                        want_rec, got_rec = {}, {}
                        want = ref_mod.autophase(copy.deepcopy(dn), peds, want_rec, "d", "m", build)
                        got = new_mod.autophase(copy.deepcopy(dn), peds, got_rec, "d", "m", build)
                        assert got == want and got_rec == want_rec, (build, chrom, kid, start, new_mod.__name__)
                        if want_rec:
                            assert list(got_rec.values())[0].keys() == list(want_rec.values())[0].keys()
    ds = make_dataset(SynthConfig(dnms_per_trio=30, n_trios=2, seed=23, coverage=1.0, chr_prefix="chr"))
    fakes.install()
    fakes.register_vcf("mem://refalt.vcf", ds.sites)
    import cyvcf2
    s = ds.sites
    probes = [(s.contigs[int(s.blk_contig[b])], int(s.pos[int(s.blk_off[b]) + k])) for b in range(s.n_blocks)
              for k in range(0, int(s.blk_off[b + 1] - s.blk_off[b]), 97)][:60]
    for contig, pos in probes:
        for chrom in (contig, contig.replace("chr", "")):
            for p in (pos, pos + 1, pos - 1, str(pos + 1)):
                want = mods["snv_phaser"].get_refalt(chrom, p, cyvcf2.VCF("mem://refalt.vcf"), 0)
                got = new_snv.get_refalt(chrom, p, cyvcf2.VCF("mem://refalt.vcf"), 0)
                assert got == want, (chrom, p)
    assert any(new_snv.get_refalt(c.replace("chr", ""), p + 1, cyvcf2.VCF("mem://refalt.vcf"), 0)[0] is not None for c, p in probes)
