"""Drop-in for the reference's ``unfazed/read_collector.py`` entry points
``collect_reads_snv`` (:339-354) and ``collect_reads_sv`` (:435-448): same positional arguments, same
return value ``({"alt": [...], "ref": [...]}, concordant_upper_len)``.

Seed selection (``goodread`` / ``snv_match_alleles`` / ``indel_match_alleles`` / the SV breakpoint
heuristics) and the extended chaining (``group_reads_by_haplotype`` / ``connect_reads``) run on the
GPU: the caller's ``het_sites`` become a one-trio site block whose rows classify as heterozygous
informative sites, the region becomes one DNM entry, and the batch pipeline does the rest
(``unfz_read_scan`` -> ``unfz_read_site_alleles`` -> ``unfz_chain_tally``).  The batched
``snv_phaser.phase_snvs`` / ``sv_phaser.phase_svs`` never go through here; these functions exist for
callers that use the reference's per-variant interface.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import datasource
from .informative_site_finder import get_engine
from .plan import Plan, concordant_upper_lens
from .readview import ReadView
from .schema import SiteTable


def _site_block(het_sites, contig: str) -> SiteTable:
    """Rows that every threshold set classifies as het + candidate: kid 0/1 20:20, dad 1/1, mom 0/0."""
    n = len(het_sites)
    pos = np.array([s["pos"] for s in het_sites], dtype=np.int32)
    one = lambda k, d, m, dt: np.ascontiguousarray(np.tile(np.array([[k], [d], [m]], dtype=dt), (1, n)))
    return SiteTable(
        trios=[("kid", "dad", "mom")], contigs=[contig], blk_trio=np.zeros(1, np.int32), blk_contig=np.zeros(1, np.int32),
        blk_off=np.array([0, n], np.int64), pos=pos, flag=np.ones(n, np.uint8),
        ref=np.array([ord(s["ref_allele"][0]) for s in het_sites], dtype=np.uint8),
        alt=np.array([ord(s["alt_allele"][0]) for s in het_sites], dtype=np.uint8),
        gt=one(1, 3, 0, np.uint8), gq=one(99, 99, 99, np.float32), rd=one(20, 0, 40, np.int32), ad=one(20, 40, 0, np.int32))


def _collect(bam_name, region, het_sites, ref, alt, cram_ref, no_extended, concordant_upper_len,
             insert_size_max_sample, stdevs, min_map_qual, min_gt_qual, readlen, split_error_margin, kind):
    from .engine import make_params
    eng = get_engine()
    het_sites = sorted(het_sites, key=lambda s: s["pos"]) if any(
        het_sites[i]["pos"] > het_sites[i + 1]["pos"] for i in range(len(het_sites) - 1)) else list(het_sites)
    start, end = int(region["start"]), int(region["end"])
    span = max([abs(s["pos"] - start) for s in het_sites] + [abs(s["pos"] - end) for s in het_sites] + [0])
    dn = {"chrom": region["chrom"], "start": start, "end": end, "kid": "kid", "vartype": "POINT", "bam": bam_name,
          "cram_ref": cram_ref}
    reads = datasource.load_reads([dn], max(span, 1), readlen, insert_size_max_sample)
    kid_idx = 0
    if reads is not None and len(reads.kids) > 1:
        raise ValueError("collect_reads_*: %s holds more than one kid" % bam_name)
    chrom = region["chrom"]
    flags = 0
    if reads is not None and chrom not in reads.contigs:
        chrom = chrom.strip("chr") if "chr" in chrom else "chr" + chrom
        flags |= L.DNM_FALLBACK_FETCH
        if chrom not in reads.contigs:
            raise ValueError("invalid contig `%s`" % region["chrom"])       # what pysam's fetch raises
    rb = reads.block_of(kid_idx, chrom) if reads is not None else -1
    sites = _site_block(het_sites, "c")
    dnm = np.zeros(1, dtype=L.DNM_DTYPE)
    blob = (ref or "").encode("ascii") + (alt or "").encode("ascii")
    dnm[0] = (start, end, rb, kind, 0, 1, 0, len(ref or ""), len(ref or ""), len(alt or ""), -1, flags)
    if kind != L.KIND_SV:
        dnm["kind"] = L.KIND_SNV if len(ref) == len(alt) else L.KIND_INDEL
    n = len(het_sites)
    seg = np.zeros(1, dtype=L.SEG_DTYPE)
    lo = int(sites.pos.min()) if n else 0
    hi = int(sites.pos.max()) if n else -1
    seg[0] = (0, lo, hi, 1, 0, 0, 0, L.MODE_READ)
    # the chain kernel needs at least one candidate to run; with no het sites add a far-away dummy row
    if n == 0:
        sites = _site_block([{"pos": -(1 << 30), "ref_allele": "A", "alt_allele": "C"}], "c")
        seg[0] = (0, -(1 << 30), -(1 << 30), 1, 0, 0, 0, L.MODE_READ)
    plan = Plan(dnm=dnm, seg=seg, alleles=np.frombuffer(blob + b"\0", dtype=np.uint8).copy(), entries=[dn],
                trio=np.zeros(1, np.int32), found=np.ones(1, bool), rblk_sblk={rb: 0} if rb >= 0 else {})
    if reads is None or rb < 0:
        return {"alt": [], "ref": []}, concordant_upper_len
    cul = concordant_upper_lens(reads, readlen, insert_size_max_sample, stdevs)
    if concordant_upper_len:
        cul = np.full_like(cul, float(concordant_upper_len))
    params = make_params(min_gt_qual=min_gt_qual, min_map_qual=min_map_qual, readlen=readlen,
                         insert_size_max_sample=insert_size_max_sample, no_extended=no_extended,
                         split_error_margin=split_error_margin)
    # min_gt_qual is also the genotype-quality gate of the synthetic rows (GQ 99): keep them passing.
    # Base qualities never exceed 93, so clamping at 99 does not change any base-quality decision.
    params.min_gt_qual = min(float(min_gt_qual), 99.0)
    # the device copy of a registered table is kept with the table (one upload per base-quality threshold)
    from .schema import min_base_qual
    key = (id(eng), min_base_qual(params.min_gt_qual))
    cache = reads.__dict__.setdefault("_device_reads", {})
    if key not in cache:
        cache.clear()
        cache[key] = eng.upload_reads(reads, min_gt_qual=params.min_gt_qual)
    res = eng.run(eng.upload_sites(sites), cache[key], plan, params, blk_cul=cul)
    lab = res.slot_labels(0)
    out = {"alt": [], "ref": []} if no_extended else {"ref": [], "alt": []}
    xs = np.nonzero(lab)[0]
    for x, r in zip(xs, res.slot_reads(0, xs)):
        r = int(r)
        m = int(reads.hdr["mate"][r])
        pair = [ReadView(reads, r)] + ([ReadView(reads, m)] if m >= 0 else [])
        for hap, bit in (("ref", 1), ("alt", 2)):
            if lab[x] & bit:
                out[hap] += pair
    return out, float(cul[rb])


def collect_reads_snv(bam_name, region, het_sites, ref, alt, cram_ref, no_extended, concordant_upper_len,
                      insert_size_max_sample, stdevs, min_map_qual, min_gt_qual, readlen, split_error_margin):
    return _collect(bam_name, region, het_sites, ref, alt, cram_ref, no_extended, concordant_upper_len,
                    insert_size_max_sample, stdevs, min_map_qual, min_gt_qual, readlen, split_error_margin, L.KIND_SNV)


def collect_reads_sv(bam_name, region, het_sites, cram_ref, no_extended, concordant_upper_len,
                     insert_size_max_sample, stdevs, min_map_qual, min_gt_qual, readlen, split_error_margin):
    return _collect(bam_name, region, het_sites, None, None, cram_ref, no_extended, concordant_upper_len,
                    insert_size_max_sample, stdevs, min_map_qual, min_gt_qual, readlen, split_error_margin, L.KIND_SV)
