"""Interface mirror of the reference's ``unfazed/site_searcher.py`` for callers that use the
per-variant API: ``binary_search`` (:6-47) and ``match_informative_sites`` (:50-78) over plain lists.

The batched product path does NOT call these: site matching, the parent-consistency filter and the
per-read base lookup run inside ``chain_kernel`` (phase 5, csrc/chain.cu).  They are list lookups
with no arithmetic; they are kept so that code written against the reference's module still imports.
"""
from __future__ import annotations


def binary_search(start, end, informative_sites):
    """Sites with start <= pos <= end around a pivot with start <= pos < end, in the reference's
    order: pivot, the run to its right, the run to its left (Q16)."""
    lo, hi = 0, len(informative_sites) - 1
    seen = (-1, -1)
    while hi > -1 and lo <= hi and (lo, hi) != seen:
        seen = (lo, hi)
        mid = (lo + hi) // 2
        p = informative_sites[mid]["pos"]
        if start <= p < end:
            right = []
            for s in informative_sites[mid + 1:]:
                if not (start <= s["pos"] <= end):
                    break
                right.append(s)
            left = []
            for s in reversed(informative_sites[:mid]):
                if not (start <= s["pos"] <= end):
                    break
                left.append(s)
            return [informative_sites[mid]] + right + left
        if p > start:
            hi = mid - 1
        elif p < start:
            lo = mid + 1
    return []


def match_informative_sites(reads, informative_sites):
    """{"ref"|"alt": [{"matches": [...], "read": read}]}; reads whose matched sites disagree on the
    parents are dropped (:69-75)."""
    out = {}
    for hap, lst in reads.items():
        out[hap] = []
        for read in lst:
            ms = binary_search(read.reference_start, read.reference_end, informative_sites)
            if ms and len({m["ref_parent"] for m in ms}) == 1 and len({m["alt_parent"] for m in ms}) == 1:
                out[hap].append({"matches": ms, "read": read})
    return out
