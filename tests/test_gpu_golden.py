"""GPU: the CUDA path against the golden vectors produced by the UNMODIFIED reference
(tests/golden, oracle/make_golden.py): records, site lists, unit-level classification vectors and
the command line's BED output."""
import copy
import os

import numpy as np
import pytest

from oracle import port
from tests.golden_util import CASES, load_case, one_trio_table, unit_vectors
from tests.util import gpu_kwargs, norm_record, port_params
from unfazed_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from unfazed_b200.engine import Engine
    return Engine(0)


@pytest.mark.parametrize("name", CASES)
def test_records_match_reference(engine, name):
    from unfazed_b200.phaser import BatchPhaser
    ds, want = load_case(name)
    bp = BatchPhaser(engine, ds.sites, ds.reads, ds.pedigrees)
    got = bp.phase(copy.deepcopy(ds.dnms), **gpu_kwargs(**ds.params))
    assert set(got) == set(want["records"])
    for k, rec in want["records"].items():
        assert norm_record(got[k]) == rec, k


@pytest.mark.parametrize("name", CASES)
def test_find_matches_reference(engine, name):
    """Drop-in informative_site_finder.find (GPU) == the reference's find on the same DNMs."""
    from unfazed_b200 import datasource, informative_site_finder as isf
    ds, want = load_case(name)
    p = port_params(**ds.params)
    datasource.register_tables("mem://golden.vcf", sites=ds.sites)
    snvs = [d for d in ds.dnms if d["vartype"].upper() in port.SNV_TYPES]
    svs = [d for d in ds.dnms if d["vartype"].upper() in port.SV_TYPES]
    for label, dn, wr in (("snv_read", snvs, False), ("sv_read", svs, False), ("sv_cnv", svs, True)):
        if label not in want["find"]:
            continue
        w = want["find"][label]
        args = (copy.deepcopy(dn), ds.pedigrees, "mem://golden.vcf", 0 if wr else p.search_dist, p.threads, p.build,
                p.multiread_proc_min, True, p.ab_homref, p.ab_homalt, p.ab_het, p.min_gt_qual, p.min_depth)
        if isinstance(w, dict) and "raises" in w:
            with pytest.raises(KeyError):
                isf.find(*args, whole_region=wr)
            continue
        got = isf.find(*args, whole_region=wr)
        assert [port._key(d) for d in got] == [port._key(d) for d in w], label      # same order, too
        for g, d in zip(got, w):
            for key in ("candidate_sites", "het_sites"):
                assert g.get(key) == d.get(key), (label, port._key(d), key)


def test_unit_vectors_through_classify_kernel(engine):
    """is_high_quality_site / get_kid_allele known answers, evaluated by unfz_classify_sites.

    The vector sits in the dad (resp. kid) slot of a row whose other members are clean, so the class
    code exposes exactly the reference function's answer."""
    from unfazed_b200.engine import make_params
    from unfazed_b200.plan import Plan
    uv = unit_vectors()
    rows, expect = [], []
    for rd, ad, gt, gq, want in uv["is_high_quality_site"]:
        rows.append(dict(pos=len(rows), gt=[1, gt, 0], gq=[99, gq, 99], rd=[15, rd, 30], ad=[15, ad, 0]))
        expect.append(("het", want))
    n_hq = len(rows)
    modes = {"DEL": L.MODE_CNV_DEL, "DUP": L.MODE_CNV_DUP, "INV": L.MODE_CNV_NA}
    ka_rows = {m: [] for m in modes}
    for vt, rd, ad, gts, want in uv["get_kid_allele"]:
        ka_rows[vt].append((len(rows), rd, ad, gts, want))
        rows.append(dict(pos=len(rows), gt=gts, gq=[99, 99, 99], rd=rd, ad=ad))
    table = one_trio_table(rows)
    n = len(rows)
    dnm = np.zeros(4, dtype=L.DNM_DTYPE)
    dnm["rblk"], dnm["cnv_entry"] = -1, -1
    seg = np.zeros(4, dtype=L.SEG_DTYPE)
    for i, mode in enumerate([L.MODE_READ, L.MODE_CNV_DEL, L.MODE_CNV_DUP, L.MODE_CNV_NA]):
        seg[i] = (0, 0, n, 1, i, 0, 0, mode)
        dnm[i]["seg_lo"], dnm[i]["seg_hi"], dnm[i]["pos"], dnm[i]["end"] = i, i + 1, 10 ** 6, 10 ** 6 + 100
    plan = Plan(dnm=dnm, seg=seg, alleles=np.zeros(0, np.uint8), entries=[{}] * 4, trio=np.zeros(4, np.int32),
                found=np.ones(4, bool))
    res = engine.run(engine.upload_sites(table), None, plan, make_params())
    cls = res.class_codes()[: 4 * n].reshape(4, n)
    for r, (_kind, want) in enumerate(expect):
        assert bool(cls[0, r] & L.CLS_HET) == want, uv["is_high_quality_site"][r]
    p = port.Params()
    for vt, mode_idx in (("DEL", 1), ("DUP", 2), ("INV", 3)):
        for r, rd, ad, gts, want in ka_rows[vt]:
            het, cand = port.classify_row(table, r, {"start": 10 ** 6, "end": 10 ** 6 + 100, "vartype": vt}, "dad", "mom", p, True)
            code = int(cls[mode_idx, r])
            assert bool(code & L.CLS_HET) == (het is not None), (vt, r)
            assert bool(code & L.CLS_CAND) == (cand is not None), (vt, r, rd, ad, gts)
            if cand is not None:
                assert want is not None
                assert bool(code & L.CLS_KID_ALT) == (want == "alt_parent")
                assert bool(code & L.CLS_ALT_IS_DAD) == (cand["alt_parent"] == "dad")


def _bed_rows(text):
    rows = []
    for line in text.strip().split("\n"):
        f = line.split("\t")
        if len(f) == 13 and not line.startswith("#"):
            f[10] = ",".join(sorted(f[10].split(",")))       # read names: set order in the reference (Q23)
            f[12] = ",".join(sorted(f[12].split(",")))
        rows.append(f)
    return rows


@pytest.mark.parametrize("decode", ["tables", "cyvcf2_pysam_api"])
@pytest.mark.parametrize("name", CASES)
def test_cli_bed_output_matches_reference(engine, tmp_path, name, decode):
    """`python -m unfazed_b200` with the reference's flags writes the BED the reference wrote.
    decode=cyvcf2_pysam_api additionally goes through datasource -> packers over the (fake)
    cyvcf2 / pysam modules, i.e. the path real files take."""
    import sys
    from oracle import fakes
    from oracle.make_golden import cli_files
    from unfazed_b200 import datasource
    from unfazed_b200.__main__ import main
    ds, want = load_case(name)
    bed, ped, pairs = cli_files(ds, str(tmp_path))
    vcf_name = "mem://golden.vcf"
    datasource._registry.clear()
    if decode == "tables":
        datasource.register_tables(vcf_name, sites=ds.sites)
        for kid, path in pairs:
            datasource.register_tables(path, reads=ds.reads)
    else:
        fakes.install()
        fakes.register_vcf(vcf_name, ds.sites)
        for kid, path in pairs:
            fakes.register_bam(path, ds.reads, ds.reads.kids.index(kid))
    p = dict(threads=1, build="38", multiread_proc_min=1000, no_extended=False)
    p.update(ds.params)
    for amb in (False, True):
        out = tmp_path / ("out_%s.bed" % amb)
        argv = ["-d", bed, "-s", vcf_name, "-p", ped, "--bam-pairs"] + ["%s:%s" % (k, b) for k, b in pairs] + \
               ["-t", str(p["threads"]), "-g", p["build"], "--multiread-proc-min", str(p["multiread_proc_min"]),
                "--quiet", "--verbose", "-o", "bed", "--outfile", str(out)]
        if p["no_extended"]:
            argv.append("--no-extended")
        if amb:
            argv.append("--include-ambiguous")
        main(argv)
        assert _bed_rows(out.read_text()) == _bed_rows(want["cli"]["ambiguous" if amb else "strict"]), (name, amb)
    for m in ("cyvcf2", "pysam"):
        if decode != "tables" and getattr(sys.modules.get(m), "__fake__", False):
            del sys.modules[m]
