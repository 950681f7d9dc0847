"""CPU: the oracle port against the golden vectors the UNMODIFIED reference produced
(oracle/make_golden.py).  This is what pins the oracle."""
import copy

import numpy as np
import pytest

from oracle import port
from tests.golden_util import CASES, load_case, one_trio_table, unit_vectors
from tests.util import norm_record, port_params


@pytest.mark.parametrize("name", CASES)
def test_port_records_match_reference(name):
    ds, want = load_case(name)
    p = port_params(**ds.params)
    ph = port.Phaser(ds.sites, ds.reads, ds.pedigrees, p)
    got = ph.phase(copy.deepcopy(ds.dnms))
    assert set(got) == set(want["records"])
    for k, rec in want["records"].items():
        assert norm_record(got[k]) == rec, k
    for amb, key in ((True, "ambiguous"), (False, "strict")):
        for k, w in want["summary"][key].items():
            g = port.summarize_record(copy.deepcopy(got[k]), amb, False, 10)
            assert g == w, (k, key)


@pytest.mark.parametrize("name", CASES)
def test_port_find_matches_reference(name):
    ds, want = load_case(name)
    p = port_params(**ds.params)
    snvs = [d for d in ds.dnms if d["vartype"].upper() in port.SNV_TYPES]
    svs = [d for d in ds.dnms if d["vartype"].upper() in port.SV_TYPES]
    for label, dn, wr in (("snv_read", snvs, False), ("sv_read", svs, False), ("sv_cnv", svs, True)):
        if label not in want["find"]:
            continue
        w = want["find"][label]
        if isinstance(w, dict) and "raises" in w:
            with pytest.raises(port.ReferenceUndefined):
                port.find(copy.deepcopy(dn), ds.pedigrees, ds.sites, p, 0 if wr else p.search_dist, wr)
            continue
        got = port.find(copy.deepcopy(dn), ds.pedigrees, ds.sites, p, 0 if wr else p.search_dist, wr)
        gk = {port._key(d): d for d in got}
        for d in w:
            g = gk[port._key(d)]
            for key in ("candidate_sites", "het_sites"):
                assert g.get(key) == d.get(key), (label, port._key(d), key)


def test_unit_is_high_quality_site():
    p = port.Params()
    for rd, ad, gt, gq, want in unit_vectors()["is_high_quality_site"]:
        t = one_trio_table([dict(pos=1, gt=[gt, 0, 0], gq=[gq, 99, 99], rd=[rd, 30, 30], ad=[ad, 0, 0])])
        assert port.high_quality(t, 0, 0, p) == want, (rd, ad, gt, gq)


def test_unit_get_kid_allele():
    p = port.Params()
    for vt, rd, ad, gts, want in unit_vectors()["get_kid_allele"]:
        t = one_trio_table([dict(pos=1, gt=gts, gq=[99, 99, 99], rd=rd, ad=ad)])
        assert port.kid_allele(t, 0, vt, p) == want, (vt, rd, ad, gts)


def test_unit_binary_search():
    for pos, start, end, want in unit_vectors()["binary_search"]:
        got = [s["pos"] for s in port.binary_search(start, end, [{"pos": x} for x in pos])]
        assert got == want, (pos, start, end)


def test_unit_autophaseable():
    ped = {"m": {"sex": "1"}, "f": {"sex": "2"}}
    for build, chrom, start, kid, want in unit_vectors()["autophaseable"]:
        assert port.autophaseable({"chrom": chrom, "start": start, "kid": kid}, ped, build) == want


def test_unit_summarize_record():
    base = {"region": {"chrom": "1", "start": 5, "end": 6}, "vartype": "POINT", "kid": "k", "dad": "D", "mom": "M"}
    for nd, nm, cd, cm, amb, want in unit_vectors()["summarize_record"]:
        rec = dict(base)
        rec.update(dad_reads=["r%d" % i for i in range(nd)], mom_reads=["s%d" % i for i in range(nm)],
                   dad_sites=[str(100 + i) for i in range(min(nd, 3))], mom_sites=[str(200 + i) for i in range(min(nm, 2))],
                   cnv_dad_sites=[str(300 + i) for i in range(cd)], cnv_mom_sites=[str(400 + i) for i in range(cm)],
                   evidence_type="readbacked" if (nd or nm) else "", cnv_evidence_type="ALLELE-BALANCE" if (cd or cm) else "")
        assert port.summarize_record(rec, amb, True, 10) == want, (nd, nm, cd, cm, amb)
