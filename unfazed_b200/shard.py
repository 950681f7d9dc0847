"""Multi-GPU sharding of the DNM list: by kid, no collective on the data path.

Every trio's site and read columns live on exactly one GPU; a DNM goes where its kid is.  The one
exception is a family heavier than a GPU's fair share, which is cut into genomic slices.  Kids are
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


def _chrom_key(d: dict):
    c = str(d["chrom"])
    c = c[3:] if c.lower().startswith("chr") else c
    return (0, int(c), "") if c.isdigit() else (1, 0, c)


def split_heavy_kids(dnms: Sequence[dict], n_ranks: int, work: Dict[str, float],
                     slack: float = 1.25) -> Dict[int, str]:
    """A family with more work than one GPU's fair share (the single-trio stress configs) is cut
    into contiguous genomic slices, each treated as a kid of its own by the LPT assignment; the
    family's site columns are then loaded by every GPU that owns a slice, its reads only for the
    windows of the slice.  Returns {index into dnms: pseudo-kid}; DNMs of light kids keep theirs."""
    total = sum(work.values())
    fair = total / max(n_ranks, 1)
    part: Dict[int, str] = {}
    if n_ranks <= 1 or total <= 0:
        return part
    by_kid: Dict[str, List[int]] = {}
    for i, d in enumerate(dnms):
        by_kid.setdefault(d["kid"], []).append(i)
    for kid, idx in by_kid.items():
        if work[kid] <= fair * slack or len(idx) < 2:
            continue
        pieces = min(len(idx), max(2, int(-(-work[kid] // fair))))
        idx = sorted(idx, key=lambda i: (_chrom_key(dnms[i]), int(dnms[i]["start"]), i))
        for j, i in enumerate(idx):
            part[i] = "%s#%d" % (kid, j * pieces // len(idx))
    return part


def shard_dnms(dnms: Sequence[dict], n_ranks: int, reads_per_kid: Optional[Dict[str, int]] = None,
               split_heavy: bool = True) -> List[List[dict]]:
    work = estimate_work(dnms, reads_per_kid)
    part = split_heavy_kids(dnms, n_ranks, work) if split_heavy else {}
    if part:
        units: Dict[str, float] = {}
        sizes: Dict[str, int] = {}
        for i, d in enumerate(dnms):
            sizes[d["kid"]] = sizes.get(d["kid"], 0) + 1
        for i, d in enumerate(dnms):
            u = part.get(i, d["kid"])
            units[u] = units.get(u, 0.0) + work[d["kid"]] / sizes[d["kid"]]
        owner = assign_kids(units, n_ranks)
    else:
        owner = assign_kids(work, n_ranks)
    shards: List[List[dict]] = [[] for _ in range(n_ranks)]
    for i, d in enumerate(dnms):
        shards[owner[part.get(i, d["kid"])]].append(d)
    return shards


def cross_kid_coupling(svs: Sequence[dict], pedigrees: dict, build, multiread_proc_min: int) -> bool:
    """True when one kid's DNMs change another kid's result, which is the single case where a
    by-kid shard would not reproduce the single-process run: ``find_many`` in whole-region (CNV)
    mode raises KeyError for a chromosome as soon as some kid has an event longer than 2 bp with no
    DNM of its own starting at the end (informative_site_finder.py:392-395,:415; quirk Q12), and the
    exception takes the whole chromosome down for every kid.  Kids that trip it themselves lose the
    chromosome on any rank; the coupling only matters for a kid that would NOT trip it alone."""
    from .plan import is_autophaseable
    live = [d for d in svs if not is_autophaseable(d, pedigrees, build)]
    if len(svs) < multiread_proc_min or not live:
        return False
    starts = {(d["kid"], d["chrom"], int(d["start"])) for d in live}
    trips = {}
    for d in live:
        s, e = int(d["start"]), int(d["end"])
        bad = (e - s) > 2 and (d["kid"], d["chrom"], e) not in starts
        key = (d["chrom"], d["kid"])
        trips[key] = trips.get(key, False) or bad
    by_chrom: Dict[str, List[bool]] = {}
    for (c, _), bad in trips.items():
        by_chrom.setdefault(c, []).append(bad)
    return any(any(v) and not all(v) for v in by_chrom.values())


def phase_sharded(phase_fn: Callable[[List[dict]], object], dnms: Sequence[dict],
                  reads_per_kid: Optional[Dict[str, int]] = None,
                  split_heavy: bool = True, all_on_rank0: bool = False) -> Optional[Dict[str, dict]]:
    """Run ``phase_fn`` on this rank's shard and gather the record dicts on rank 0 (None elsewhere).
    Works without an initialised process group (single process).  ``all_on_rank0`` keeps the whole list
    on rank 0 (runs whose kids are coupled) while every rank still takes part in the gather."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        out = {}
        merge_part(out, phase_fn(list(dnms)))
        return out
    rank, world = dist.get_rank(), dist.get_world_size()
    if all_on_rank0:
        mine = list(dnms) if rank == 0 else []
    else:
        mine = shard_dnms(dnms, world, reads_per_kid, split_heavy)[rank]
    try:
        local = (True, phase_fn(mine) if mine else {})
    except BaseException as exc:                    # every rank must reach the gather
        local = (False, exc if isinstance(exc, Exception) else RuntimeError(repr(exc)))
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        if not local[0]:
            raise local[1]
        return None
    out: Dict[str, dict] = {}
    for ok, part in gathered:
        if not ok:
            raise part
    for _, part in gathered:
        merge_part(out, part)
    return out


def merge_part(out: Dict[str, dict], part) -> None:
    """A rank's contribution is either the record dict itself or its compact form (phaser.CompactRecords: arrays that
    pickle in microseconds); the compact form becomes dicts here, on rank 0, through the native builder."""
    if hasattr(part, "to_records"):
        part.to_records(out)
    else:
        out.update(part)
