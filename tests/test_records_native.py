"""CPU: the native record builder (unfazed_b200/csrc/records.c) against the plain-Python statement of the
same reshaping (snv_phaser.py:169-203 record layout; sorted(set(str(pos))) site lists; names in list order)."""
import numpy as np
import pytest

from unfazed_b200 import _records


def python_records(entries, ped, live, auto, names, pair_ids, rd, rm, pd, pm, off):
    out = {}
    for d, is_auto in zip(live.tolist(), auto.tolist()):
        dn = entries[d]
        p = ped[dn["kid"]]
        key = "_".join([str(dn["chrom"]), str(dn["start"]), str(dn["end"]), dn["kid"], dn["vartype"]])
        base = {"region": {"chrom": dn["chrom"], "start": dn["start"], "end": dn["end"]},
                "vartype": dn["vartype"], "kid": dn["kid"], "dad": p["dad"], "mom": p["mom"]}
        if is_auto:
            base.update({"cnv_dad_sites": "NA", "cnv_mom_sites": "NA", "cnv_evidence_type": "SEX-CHROM", "dad_sites": "",
                         "mom_sites": "", "evidence_type": "SEX-CHROM", "dad_reads": [], "mom_reads": []})
        else:
            nm = (lambda i: names[i]) if names is not None else (lambda i: "q%d" % pair_ids[i])
            base.update({
                "dad_sites": sorted({str(int(x)) for x in pd[off[2, d]:off[2, d + 1]]}),
                "mom_sites": sorted({str(int(x)) for x in pm[off[3, d]:off[3, d + 1]]}),
                "evidence_type": "readbacked",
                "dad_reads": [nm(int(i)) for i in rd[off[0, d]:off[0, d + 1]]],
                "mom_reads": [nm(int(i)) for i in rm[off[1, d]:off[1, d + 1]]],
                "cnv_dad_sites": "", "cnv_mom_sites": "", "cnv_evidence_type": ""})
        out[key] = base
    return out


def make_case(seed, n, with_names):
    rng = np.random.default_rng(seed)
    n_reads = 5000
    entries = [{"chrom": rng.choice(["1", "chrX", 7]), "start": int(rng.integers(1, 10 ** 9)), "end": int(rng.integers(1, 10 ** 9)),
                "kid": "kid%d" % rng.integers(0, 3), "vartype": rng.choice(["POINT", "DEL", "INS"])} for _ in range(n)]
    for e in entries:                                   # numpy scalars would not be what the readers produce
        e["chrom"] = e["chrom"].item() if hasattr(e["chrom"], "item") else e["chrom"]
        e["vartype"] = str(e["vartype"])
    ped = {"kid%d" % k: {"dad": "dad%d" % k, "mom": "mom%d" % k} for k in range(3)}
    cnt = rng.integers(0, 90, size=(4, n))
    cnt[:, rng.random(n) < 0.2] = 0
    cnt[2:, rng.random(n) < 0.2] = 1                     # single-site lists are returned as they come
    off = np.zeros((4, n + 1), dtype=np.int64)
    off[:, 1:] = np.cumsum(cnt, axis=1)
    rd = rng.integers(0, n_reads, size=off[0, n]).astype(np.int32)
    rm = rng.integers(0, n_reads, size=off[1, n]).astype(np.int32)
    # positions with repeats and with different digit counts: str order differs from numeric order
    pool = np.concatenate([rng.integers(1, 100, 20), rng.integers(900, 1100, 20), rng.integers(99990, 100010, 20),
                           rng.integers(1, 2 ** 31 - 1, 20)])
    pd = rng.choice(pool, size=off[2, n]).astype(np.int32)
    pm = rng.choice(pool, size=off[3, n]).astype(np.int32)
    live = np.flatnonzero(rng.random(n) < 0.7).astype(np.int64)
    auto = (rng.random(live.shape[0]) < 0.15).astype(np.uint8)
    names = ["read:%d/x" % i for i in range(n_reads)] if with_names else None
    pair_ids = None if with_names else rng.integers(0, 10 ** 11, size=n_reads).astype(np.int64)
    return entries, ped, live, auto, names, pair_ids, rd, rm, pd, pm, off


@pytest.mark.parametrize("with_names", [True, False])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_native_records_equal_python(seed, with_names):
    case = make_case(seed, 300, with_names)
    want = python_records(*case)
    got = {}
    added = _records.read_records(got, *case)
    assert added == len(case[2])
    assert got == want
    for k in want:                                       # same key order too (writers iterate the dicts)
        assert list(got[k]) == list(want[k]), k
    assert list(got) == list(want)


def test_native_records_empty_and_errors():
    entries, ped, live, auto, names, pair_ids, rd, rm, pd, pm, off = make_case(5, 10, True)
    got = {}
    z = np.zeros(0, dtype=np.int64)
    assert _records.read_records(got, entries, ped, z, z.astype(np.uint8), names, None, rd, rm, pd, pm, off) == 0 and got == {}
    with pytest.raises(IndexError):
        _records.read_records({}, entries, ped, np.array([10], dtype=np.int64), np.zeros(1, np.uint8), names, None, rd, rm, pd, pm, off)
    with pytest.raises(IndexError):                      # a read index beyond the name table
        _records.read_records({}, entries, ped, np.arange(10, dtype=np.int64), np.zeros(10, np.uint8), names[:5], None, rd, rm, pd, pm, off)
    with pytest.raises(KeyError):                        # unknown kid: the same KeyError the Python loop raised
        _records.read_records({}, entries, {}, np.arange(10, dtype=np.int64), np.zeros(10, np.uint8), names, None, rd, rm, pd, pm, off)
    with pytest.raises(TypeError):
        _records.read_records({}, entries, ped, np.arange(10, dtype=np.int32), np.zeros(10, np.uint8), names, None, rd, rm, pd, pm, off)
