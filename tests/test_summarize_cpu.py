"""CPU, build container only: `summarize_record` (the final call, `unfazed/unfazed.py:162-334`) of the
drop-in module and of the oracle port against the reference's, over random evidence records that cover
every branch of the decision table (read-backed / allele-balance / contradictory / ambiguous /
autophased), for several `evidence_min_ratio` values and both flags."""
import copy
import random

import pytest

from oracle import port, ref_driver

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_driver.available(), reason="reference checkout not mounted")]


def _record(rng):
    n = lambda hi: rng.choice([0, 0, 0, 1, 2, 3, hi])
    sites = lambda k, tag: [str(rng.randint(1, 99999)) for _ in range(k)]
    reads = lambda k, tag: ["%s%d" % (tag, i) for i in range(k)]
    sex = rng.random() < 0.1
    rec = {
        "region": {"chrom": rng.choice(["1", "chrX", "22"]), "start": rng.randint(1, 10 ** 6), "end": rng.randint(1, 10 ** 6)},
        "vartype": rng.choice(["POINT", "DEL", "DUP", "INV"]),
        "kid": "kid", "dad": "dad", "mom": "mom",
        "dad_sites": sites(n(12), "d"), "mom_sites": sites(n(12), "m"),
        "dad_reads": reads(n(40), "dr"), "mom_reads": reads(n(40), "mr"),
        "cnv_dad_sites": sites(n(25), "cd"), "cnv_mom_sites": sites(n(25), "cm"),
        "evidence_type": "SEX-CHROM" if sex else rng.choice(["READBACKED", "NA", "READBACKED,NA"]),
    }
    if sex:
        rec["origin_parent"] = rng.choice(["dad", "mom"])
        rec["other_parent"] = "mom" if rec["origin_parent"] == "dad" else "dad"
    return rec


def test_summarize_record_random():
    ref = ref_driver.modules()["unfazed"]
    from unfazed_b200 import unfazed as new
    rng = random.Random(2024)
    seen = set()
    for _ in range(4000):
        rec = _record(rng)
        for ratio in (1, 3, 10):
            for amb in (False, True):
                for verbose in (False, True):
                    try:
                        want = ref.summarize_record(copy.deepcopy(rec), amb, verbose, ratio)
                    except Exception as e:
                        with pytest.raises(type(e)):
                            new.summarize_record(copy.deepcopy(rec), amb, verbose, ratio)
                        continue
                    got = new.summarize_record(copy.deepcopy(rec), amb, verbose, ratio)
                    assert got == want, (rec, ratio, amb, verbose)
                    got_p = port.summarize_record(copy.deepcopy(rec), amb, verbose, ratio)
                    assert got_p == want, (rec, ratio, amb, verbose)
                    if want:
                        seen.add(tuple(want["evidence_types"]))
    # the sweep reached every kind of call the reference can make (AMBIGUOUS_BOTH is unreachable there: the
    # origin is only ever a single parent when READBACKED was appended with it)
    for kind in (("READBACKED",), ("ALLELE-BALANCE",), ("READBACKED", "ALLELE-BALANCE"), ("AMBIGUOUS_READBACKED",),
                 ("AMBIGUOUS_ALLELE-BALANCE",), ("AMBIGUOUS_READBACKED", "AMBIGUOUS_ALLELE-BALANCE")):
        assert kind in seen, kind
