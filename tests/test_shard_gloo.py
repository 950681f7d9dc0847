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


def _worker(rank, world, port_no, out_path):
    import pickle
    import torch.distributed as dist
    from oracle import port
    from unfazed_b200.shard import phase_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ds = make_dataset(SynthConfig(n_trios=3, dnms_per_trio=4, seed=77, coverage=16.0))

    def phase_fn(dnms):
        ph = port.Phaser(ds.sites, ds.reads, ds.pedigrees, port.Params())
        return ph.phase(copy.deepcopy(dnms))

    recs = phase_sharded(phase_fn, ds.dnms)
    if rank == 0:
        with open(out_path, "wb") as f:
            pickle.dump(recs, f)
    else:
        assert recs is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_equals_single_process(tmp_path):
    import pickle
    from oracle import port
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "recs.pkl")
    mp.spawn(_worker, args=(2, port_no, out), nprocs=2, join=True)
    got = pickle.load(open(out, "rb"))
    ds = make_dataset(SynthConfig(n_trios=3, dnms_per_trio=4, seed=77, coverage=16.0))
    want = port.Phaser(ds.sites, ds.reads, ds.pedigrees, port.Params()).phase(copy.deepcopy(ds.dnms))
    assert set(got) == set(want) and len(want) > 0
    for k in want:
        for f in ("dad_sites", "mom_sites", "dad_reads", "mom_reads"):
            assert sorted(got[k][f]) == sorted(want[k][f])
