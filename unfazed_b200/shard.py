"""Multi-GPU sharding of the DNM list: by kid, no collective on the data path.

Every trio's site and read columns live on exactly one GPU; a DNM goes where its kid is.  Kids are
assigned longest-processing-time-first on an estimate of their work (reads + site pairs), the
classic greedy for makespan.  Results are a few bytes per DNM and are gathered on rank 0 with
``torch.distributed.gather_object`` -- the only communication of a sharded run.
"""
from __future__ import annotations

import heapq
from typing import Callable, Dict, List, Optional, Sequence


def assign_kids(work: Dict[str, float], n_ranks: int) -> Dict[str, int]:
    """LPT greedy: heaviest kid first onto the least loaded rank; ties broken by kid id (stable)."""
    loads = [(0.0, r) for r in range(n_ranks)]
    heapq.heapify(loads)
    out: Dict[str, int] = {}
    for kid in sorted(work, key=lambda k: (-work[k], k)):
        load, r = heapq.heappop(loads)
        out[kid] = r
        heapq.heappush(loads, (load + work[kid], r))
    return out


def estimate_work(dnms: Sequence[dict], reads_per_kid: Optional[Dict[str, int]] = None) -> Dict[str, float]:
    work: Dict[str, float] = {}
    for d in dnms:
        work[d["kid"]] = work.get(d["kid"], 0.0) + 1.0
    if reads_per_kid:
        for k in work:
            work[k] += reads_per_kid.get(k, 0) / 2000.0      # ~2k reads per DNM window at 30x
    return work


def shard_dnms(dnms: Sequence[dict], n_ranks: int, reads_per_kid: Optional[Dict[str, int]] = None) -> List[List[dict]]:
    owner = assign_kids(estimate_work(dnms, reads_per_kid), n_ranks)
    shards: List[List[dict]] = [[] for _ in range(n_ranks)]
    for d in dnms:
        shards[owner[d["kid"]]].append(d)
    return shards


def phase_sharded(phase_fn: Callable[[List[dict]], Dict[str, dict]], dnms: Sequence[dict],
                  reads_per_kid: Optional[Dict[str, int]] = None) -> Optional[Dict[str, dict]]:
    """Run ``phase_fn`` on this rank's shard and gather the record dicts on rank 0 (None elsewhere).
    Works without an initialised process group (single process)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return phase_fn(list(dnms))
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = shard_dnms(dnms, world, reads_per_kid)[rank]
    local = phase_fn(mine) if mine else {}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        return None
    out: Dict[str, dict] = {}
    for part in gathered:
        out.update(part)
    return out
