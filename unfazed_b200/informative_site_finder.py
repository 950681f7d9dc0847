"""Drop-in for the reference's ``unfazed/informative_site_finder.py``.

``find`` keeps the reference signature (:167-182) and its return conventions -- the same DNM dicts
gain ``candidate_sites`` / ``het_sites`` (sorted by pos, duplicates preserved); ``find_many``
ordering and key-presence rules (:647-661) are reproduced -- but the window search, the trio
classification of every (DNM x site) pair and the compaction run on the GPU
(unfz_window_search / unfz_classify_sites / unfz_compact_sites).
"""
from __future__ import annotations

import sys
from typing import List

from . import datasource
from .plan import FindManyKeyError, SiteIndex, is_autophaseable, plan_find_fast as plan_find

_engine = None


def get_engine():
    global _engine
    if _engine is None:
        from .engine import Engine
        _engine = Engine(0)
    return _engine


def autophaseable(denovo, pedigrees, build):
    """Reference :137-164."""
    return is_autophaseable(denovo, pedigrees, build)


def find(dnms, pedigrees, vcf_name, search_dist, threads, build, multithread_proc_min, quiet_mode,
         ab_homref, ab_homalt, ab_het, min_gt_qual, min_depth, whole_region=True):
    if len(dnms) <= 0 and not (len(dnms) >= multithread_proc_min):
        return None
    from .engine import make_params
    from .phaser import BatchPhaser
    eng = get_engine()
    sites = datasource.load_sites(vcf_name, dnms, pedigrees, search_dist)
    bp = BatchPhaser(eng, sites, None, pedigrees)
    try:
        plan = plan_find(dnms, pedigrees, bp.sidx, None, search_dist=search_dist, whole_region=whole_region,
                         build=build, multiread_proc_min=multithread_proc_min, threads=threads, with_reads=False)
    except FindManyKeyError as e:
        raise KeyError(str(e))
    params = make_params(ab_homref, ab_homalt, ab_het, min_gt_qual, min_depth)
    res = eng.run(bp.dsites, None, plan, params)
    use_many = len(dnms) >= multithread_proc_min
    for d, dn in enumerate(dnms):
        if not plan.found[d]:
            continue
        cands, hets = bp.site_dicts(res, d, int(plan.trio[d]), whole_region and ("vartype" in dn))
        if use_many:
            # find_many appends to whatever the dict already holds and leaves absent keys absent (Q13)
            if hets:
                dn["het_sites"] = sorted(dn.get("het_sites", []) + hets, key=lambda x: x["pos"])
            if cands:
                dn["candidate_sites"] = sorted(dn.get("candidate_sites", []) + cands, key=lambda x: x["pos"])
        else:
            dn["candidate_sites"], dn["het_sites"] = cands, hets
    if not use_many:
        return dnms
    # find_many returns the DNMs grouped by sample / chrom / start, autophased ones last (:647-661)
    auto = [dn for dn in dnms if is_autophaseable(dn, pedigrees, build)]
    rest = [dn for dn in dnms if not is_autophaseable(dn, pedigrees, build)]
    groups: dict = {}
    for dn in rest:
        groups.setdefault(dn["kid"], {}).setdefault(dn["chrom"], {}).setdefault(int(dn["start"]), []).append(dn)
    ordered: List[dict] = []
    for k in groups:
        for c in groups[k]:
            for s in groups[k][c]:
                ordered += groups[k][c][s]
    return ordered + auto


if __name__ == "__main__":
    sys.exit("Import this as a module")
