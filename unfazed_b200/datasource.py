"""Where the drop-in modules get their columnar tables from.

* ``*.npz`` written by ``tableio.save_tables`` (our native cache of decoded inputs), or
* the sites VCF/BCF through ``cyvcf2`` and the kids' BAM/CRAMs through ``pysam`` -- imported
  lazily, so the package works without them as long as tables are supplied -- decoded only around
  the DNMs and packed by ``packers.py``.

Tables are cached per (file, request) for the lifetime of the process, like the reference's
``concordant_upper_lens`` cache (snv_phaser.py:14).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

from . import packers
from .plan import strip_chr
from .schema import ReadTable, SiteTable
from .tableio import load_tables

_npz_cache: Dict[str, tuple] = {}
_registry: Dict[str, object] = {}


def register_tables(name: str, sites: Optional[SiteTable] = None, reads: Optional[ReadTable] = None) -> None:
    """Bind an in-memory table to a file name (``--sites`` / bam path) -- used by tests and by
    callers that decode their inputs themselves."""
    _registry[name] = (sites, reads)


def is_registered(name: str) -> bool:
    """True for a name bound to in-memory tables with register_tables() (such a name needs no file)."""
    return name in _registry


def _npz(name):
    if name not in _npz_cache:
        _npz_cache[name] = load_tables(name)
    return _npz_cache[name]


def _toggle(chrom: str) -> str:
    return strip_chr(chrom) if "chr" in chrom else "chr" + chrom


def load_sites(vcf_name: str, dnms: List[dict], pedigrees: dict, search_dist: int) -> SiteTable:
    if vcf_name in _registry and _registry[vcf_name][0] is not None:
        return _registry[vcf_name][0]
    if vcf_name.endswith(".npz"):
        return _npz(vcf_name)[0]
    from cyvcf2 import VCF
    from .utils import get_prefix
    prefix = get_prefix(VCF(vcf_name))
    trios, seen = [], set()
    for d in dnms:
        ped = pedigrees.get(d["kid"])
        if ped and d["kid"] not in seen:
            seen.add(d["kid"])
            trios.append((d["kid"], ped["dad"], ped["mom"]))
    regions: Dict[str, List[Tuple[int, int]]] = {}
    for d in dnms:
        contig = prefix + strip_chr(d["chrom"])
        s, e = int(d["start"]), int(d["end"])
        regions.setdefault(contig, []).append((s - search_dist - 2, e + search_dist + 2))
    return packers.pack_sites(VCF(vcf_name), trios, regions)


def load_reads(dnms: List[dict], search_dist: int, readlen: int, insert_size_max_sample: int) -> Optional[ReadTable]:
    """One ReadTable for all kids of the DNM list (their ``bam`` entries name the files)."""
    bam_of: Dict[str, str] = {}
    for d in dnms:
        bam_of.setdefault(d["kid"], d["bam"])
    if not bam_of:
        return None
    names = set(bam_of.values())
    if len(names) == 1:
        name = next(iter(names))
        if name in _registry and _registry[name][1] is not None:
            return _registry[name][1]
        if name.endswith(".npz"):
            return _npz(name)[1]
    reg = [_registry[n][1] for n in names if n in _registry and _registry[n][1] is not None]
    if reg and all(r is reg[0] for r in reg) and len(reg) == len(names):
        return reg[0]
    import pysam
    bams, regions = {}, {}
    for kid, name in bam_of.items():
        cram_ref = next((d.get("cram_ref") for d in dnms if d["kid"] == kid), None)
        bam = pysam.AlignmentFile(name, "rc", reference_filename=cram_ref) if name[-4:] == "cram" else pysam.AlignmentFile(name, "rb")
        bams[kid] = bam
    table = None
    # the insert-size estimate needs the head of every file first; then fetch around the DNMs
    heads = {}
    for kid, bam in bams.items():
        tl = []
        for i, r in enumerate(bam):
            tl.append(r.tlen)
            if i >= insert_size_max_sample:
                break
        heads[kid] = np.array(tl, dtype=np.int64)
    for d in dnms:
        kid = d["kid"]
        h = heads[kid]
        cul = int(np.percentile(np.abs(h - 2 * readlen), 99.5)) if h.shape[0] else 0
        margin = max(search_dist, cul) + 2 * readlen + 2
        chrom = d["chrom"]
        try:
            bams[kid].fetch(chrom, 0, 1)
        except ValueError:
            chrom = _toggle(chrom)
        for p in {int(d["start"]), int(d["end"])}:
            regions.setdefault(kid, {}).setdefault(chrom, []).append((p - margin, p + margin))
    table = packers.pack_reads(bams, regions)
    table.head_tlen = heads
    return table
