"""CPU: the multi-GPU sharding logic with world_size 2 over gloo.  Each rank phases its shard (the
per-rank worker here is the oracle port, because this box has no GPU) and rank 0 must end up with
exactly the single-process result."""
import copy
import os
import socket

import pytest
import torch.multiprocessing as mp

from unfazed_b200.shard import assign_kids, shard_dnms
from unfazed_b200.synth import SynthConfig, make_dataset


def test_lpt_assignment_balances_and_is_deterministic():
    work = {"k%d" % i: float(w) for i, w in enumerate([9, 7, 6, 5, 5, 4, 2, 1])}
    a = assign_kids(work, 3)
    assert a == assign_kids(dict(reversed(list(work.items()))), 3)
    loads = [sum(work[k] for k in work if a[k] == r) for r in range(3)]
    assert max(loads) - min(loads) <= 2
    assert set(a.values()) == {0, 1, 2}


def test_every_dnm_lands_on_its_kids_rank():
    dnms = [{"kid": "k%d" % (i % 5), "start": i} for i in range(40)]
    shards = shard_dnms(dnms, 4)
    assert sum(len(s) for s in shards) == 40
    for s in shards:
        kids = {d["kid"] for d in s}
        for other in shards:
            if other is not s:
                assert not kids & {d["kid"] for d in other}


def test_single_heavy_family_is_cut_into_genomic_slices():
    # one trio with 90 DNMs next to two light ones, 4 GPUs: the heavy family must spread out, in
    # contiguous (chrom, start) slices; light kids stay whole
    dnms = [{"kid": "big", "chrom": "chr%d" % (1 + i % 3), "start": 1000 * (90 - i)} for i in range(90)]
    dnms += [{"kid": "s1", "chrom": "1", "start": i} for i in range(5)]
    dnms += [{"kid": "s2", "chrom": "X", "start": i} for i in range(5)]
    shards = shard_dnms(dnms, 4)
    assert sorted(id(d) for s in shards for d in s) == sorted(id(d) for d in dnms)
    sizes = [len(s) for s in shards]
    assert max(sizes) <= 35 and min(sizes) >= 15
    for kid in ("s1", "s2"):
        assert sum(1 for s in shards if any(d["kid"] == kid for d in s)) == 1
    spans = []
    for s in shards:
        big = sorted((int(d["chrom"][3:]), d["start"]) for d in s if d["kid"] == "big")
        if big:
            spans.append((big[0], big[-1]))
    spans.sort()
    assert len(spans) >= 3
    for a, b in zip(spans, spans[1:]):
        assert a[1] < b[0]                      # slices do not interleave
    # opting out keeps the family on one GPU
    whole = shard_dnms(dnms, 4, split_heavy=False)
    assert sum(1 for s in whole if any(d["kid"] == "big" for d in s)) == 1
    assert shard_dnms(dnms, 4) == shards        # deterministic


def test_cross_kid_coupling_detector():
    """Q12: in find_many CNV mode one kid's KeyError empties the chromosome for everybody; only a
    kid that would not trip it alone makes a by-kid shard differ from the single-process run."""
    from unfazed_b200.shard import cross_kid_coupling
    ped = {k: {"dad": "d", "mom": "m", "sex": 2} for k in ("k1", "k2")}
    sv = lambda k, c, s, e: {"kid": k, "chrom": c, "start": s, "end": e, "vartype": "DEL"}
    assert not cross_kid_coupling([sv("k1", "1", 100, 500), sv("k2", "1", 900, 1500)], ped, "38", 2)
    assert cross_kid_coupling([sv("k1", "1", 100, 500), sv("k2", "1", 900, 902)], ped, "38", 2)
    assert not cross_kid_coupling([sv("k1", "1", 100, 500), sv("k2", "1", 900, 902)], ped, "38", 3)   # per-DNM find
    assert not cross_kid_coupling([sv("k1", "1", 100, 500), sv("k2", "2", 900, 902)], ped, "38", 2)
    # a DNM of the same kid starting at the end location defuses the KeyError
    assert not cross_kid_coupling([sv("k1", "1", 100, 500), sv("k1", "1", 500, 502), sv("k2", "1", 900, 902)], ped, "38", 2)


def _compact_from_records(recs):
    """Test helper: finished read-backed records -> phaser.CompactRecords (what a GPU rank ships to rank 0)."""
    import numpy as np
    from unfazed_b200.phaser import CompactRecords
    entries, auto, names_d, names_m, pd, pm = [], [], [], [], [], []
    off = [[0], [0], [0], [0]]
    ped = {}
    for r in recs.values():
        entries.append({"chrom": r["region"]["chrom"], "start": r["region"]["start"], "end": r["region"]["end"], "kid": r["kid"],
                        "vartype": r["vartype"]})
        ped[r["kid"]] = {"dad": r["dad"], "mom": r["mom"]}
        is_auto = r["evidence_type"] == "SEX-CHROM"
        auto.append(1 if is_auto else 0)
        if not is_auto:
            names_d += r["dad_reads"]
            names_m += r["mom_reads"]
            pd += [int(x) for x in r["dad_sites"]]
            pm += [int(x) for x in r["mom_sites"]]
        for q, lst in enumerate((names_d, names_m, pd, pm)):
            off[q].append(len(lst))
    return CompactRecords(entries, ped, np.array(auto, dtype=np.uint8), names_d + names_m, None, len(names_d),
                          np.array(pd, dtype=np.int32), np.array(pm, dtype=np.int32), np.array(off, dtype=np.int64))


def _worker(rank, world, port_no, out_path, n_trios, per_trio, compact_rank=-1):
    import pickle
    import torch.distributed as dist
    from oracle import port
    from unfazed_b200.shard import phase_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = make_dataset(SynthConfig(n_trios=n_trios, dnms_per_trio=per_trio, seed=77, coverage=16.0))

    def phase_fn(dnms):
        ph = port.Phaser(ds.sites, ds.reads, ds.pedigrees, port.Params())
        recs = ph.phase(copy.deepcopy(dnms))
        return _compact_from_records(recs) if rank == compact_rank else recs

    if n_trios == 1:                            # the family is split: both ranks must get work
        from unfazed_b200.shard import shard_dnms as _sd
        assert all(len(x) > 0 for x in _sd(ds.dnms, world))
    recs = phase_sharded(phase_fn, ds.dnms)
    if rank == 0:
        with open(out_path, "wb") as f:
            pickle.dump(recs, f)
    else:
        assert recs is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_trios,per_trio,compact_rank", [(3, 4, -1), (1, 10, -1), (3, 4, 1), (3, 4, 0)],
                         ids=["kids", "slices", "compact_from_rank1", "compact_from_rank0"])
def test_sharded_equals_single_process(tmp_path, n_trios, per_trio, compact_rank):
    import pickle
    from oracle import port
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "recs.pkl")
    mp.spawn(_worker, args=(2, port_no, out, n_trios, per_trio, compact_rank), nprocs=2, join=True)
    got = pickle.load(open(out, "rb"))
    ds = make_dataset(SynthConfig(n_trios=n_trios, dnms_per_trio=per_trio, seed=77, coverage=16.0))
    want = port.Phaser(ds.sites, ds.reads, ds.pedigrees, port.Params()).phase(copy.deepcopy(ds.dnms))
    assert set(got) == set(want) and len(want) > 0
    for k in want:
        for f in ("dad_sites", "mom_sites", "dad_reads", "mom_reads"):
            assert sorted(got[k][f]) == sorted(want[k][f])
        for f in ("region", "vartype", "kid", "dad", "mom", "evidence_type", "cnv_dad_sites", "cnv_mom_sites", "cnv_evidence_type"):
            assert got[k][f] == want[k][f], (k, f)
