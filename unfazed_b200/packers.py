"""Decoders -> columnar tables.  cyvcf2 / pysam (or anything exposing the same API, e.g. the
in-memory fakes the tests install) decode the sites VCF/BCF and the kids' BAM/CRAMs; this module
packs what they yield into the pinned-friendly SoA tables of ``schema.py``.

This is host-side I/O glue ("mate resolution + columnar packer", "streaming site
table"); no phasing arithmetic happens here.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .schema import (AUX_HAS_SA, AUX_SAME_REF, QUAL_ESCAPE, READ_HDR, SITE_FLAG_SIMPLE, ReadTable,
                     SiteTable, encode_bases, pack_seq)


def _merge(intervals: List[Tuple[int, int]]) -> List[Tuple[int, int]]:
    out: List[List[int]] = []
    for a, b in sorted(intervals):
        if out and a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return [(a, b) for a, b in out]


def pack_sites(vcf, trios: Sequence[Tuple[str, str, str]], regions: Optional[Dict[str, List[Tuple[int, int]]]] = None) -> SiteTable:
    """Trio-major extraction of a joint VCF.

    ``vcf``: a cyvcf2.VCF-like handle.  ``regions``: contig -> [(start0, end0)] 0-based half-open
    intervals to decode (None: the whole file).  Every record becomes one row per trio block,
    tagged with the same ``rec_id``.
    """
    samples = list(vcf.samples)
    sidx = {s: i for i, s in enumerate(samples)}
    cols = [(t, [sidx[m] for m in trio]) for t, trio in enumerate(trios) if all(m in sidx for m in trio)]
    contigs: List[str] = []
    cindex: Dict[str, int] = {}
    per: Dict[Tuple[int, int], dict] = {}
    extras_rows: Dict[Tuple[int, int], Dict[int, Tuple[str, List[str]]]] = {}
    rec_id = 0

    def records():
        if regions is None:
            yield from vcf
        else:
            seen = set()
            for contig, ivs in regions.items():
                for a, b in _merge(ivs):
                    for v in vcf("%s:%d-%d" % (contig, max(a, 0) + 1, max(b, 1))):
                        key = (v.CHROM, v.start, v.REF, tuple(v.ALT))
                        if key in seen:
                            continue
                        seen.add(key)
                        yield v

    for v in records():
        c = cindex.get(v.CHROM)
        if c is None:
            c = cindex[v.CHROM] = len(contigs)
            contigs.append(v.CHROM)
        alts = list(v.ALT)
        simple = len(alts) == 1 and len(v.REF) == 1 and len(alts[0]) == 1 and alts[0] != "*"
        gt, gq, rd, ad = v.gt_types, v.gt_quals, v.gt_ref_depths, v.gt_alt_depths
        for t, ix in cols:
            blk = per.setdefault((t, c), {k: [] for k in ("pos", "flag", "ref", "alt", "gt", "gq", "rd", "ad", "rid")})
            if not simple:
                extras_rows.setdefault((t, c), {})[len(blk["pos"])] = (v.REF, alts)
            blk["pos"].append(v.start)
            blk["flag"].append(SITE_FLAG_SIMPLE if simple else 0)
            blk["ref"].append(ord(v.REF[0]) if v.REF else 0)
            blk["alt"].append(ord(alts[0][0]) if alts and alts[0] else 0)
            blk["gt"].append([int(gt[i]) for i in ix])
            blk["gq"].append([float(gq[i]) for i in ix])
            blk["rd"].append([int(rd[i]) for i in ix])
            blk["ad"].append([int(ad[i]) for i in ix])
            blk["rid"].append(rec_id)
        rec_id += 1

    keys = sorted(per)
    offs = [0]
    parts = {k: [] for k in ("pos", "flag", "ref", "alt", "gt", "gq", "rd", "ad", "rid")}
    extras: Dict[int, Tuple[str, List[str]]] = {}
    for key in keys:
        blk = per[key]
        order = np.argsort(np.array(blk["pos"], dtype=np.int64), kind="stable")
        inv = {int(o): i for i, o in enumerate(order)}
        for local, val in extras_rows.get(key, {}).items():
            extras[offs[-1] + inv[local]] = val
        for name in parts:
            a = np.array(blk[name])
            parts[name].append(a[order])
        offs.append(offs[-1] + len(order))
    cat = lambda name, dt: (np.concatenate(parts[name]).astype(dt) if parts[name] else np.zeros(0, dtype=dt))
    V = offs[-1]
    tri = lambda name, dt: np.ascontiguousarray(cat(name, dt).reshape(V, 3).T) if V else np.zeros((3, 0), dtype=dt)
    table = SiteTable(
        trios=[tuple(t) for t in trios], contigs=contigs,
        blk_trio=np.array([k[0] for k in keys], dtype=np.int32), blk_contig=np.array([k[1] for k in keys], dtype=np.int32),
        blk_off=np.array(offs, dtype=np.int64), pos=cat("pos", np.int32), flag=cat("flag", np.uint8),
        ref=cat("ref", np.uint8), alt=cat("alt", np.uint8), gt=tri("gt", np.uint8), gq=tri("gq", np.float32),
        rd=tri("rd", np.int32), ad=tri("ad", np.int32), extras=extras, rec_id=cat("rid", np.int64))
    table.validate()
    return table


def _cigar_words(cigartuples) -> List[int]:
    return [(int(ln) << 4) | int(op) for op, ln in (cigartuples or [])]


def pack_reads(bams: Dict[str, object], regions: Dict[str, Dict[str, List[Tuple[int, int]]]],
               head_reads: int = 0) -> ReadTable:
    """Reads of every kid overlapping the requested regions, plus their mates, in file order.

    ``bams``: kid -> pysam.AlignmentFile-like handle.  ``regions``: kid -> contig -> [(start0, end0)].
    ``head_reads``: also keep the template lengths of the first N reads of the file (the reference
    estimates the concordant insert size from the head of the BAM, read_collector.py:11-25);
    they are returned in ``table.head_tlen[kid]``.
    """
    kids = list(bams)
    contigs: List[str] = []
    cindex: Dict[str, int] = {}
    blocks = []
    head_tlen: Dict[str, np.ndarray] = {}
    for k, kid in enumerate(kids):
        bam = bams[kid]
        if head_reads:
            tl = []
            for i, r in enumerate(bam):
                tl.append(r.tlen)
                if i >= head_reads:
                    break
            head_tlen[kid] = np.array(tl, dtype=np.int64)
        for contig, ivs in regions.get(kid, {}).items():
            recs: Dict[tuple, object] = {}

            def ident(r):
                return (r.query_name, r.flag & 0xC0, r.reference_start, r.flag & 0x900)

            for a, b in _merge(ivs):
                try:
                    it = bam.fetch(contig, max(a, 0), max(b, 1))
                except ValueError:
                    continue
                for r in it:
                    recs.setdefault(ident(r), r)
            # mates that lie outside the fetched intervals (bamfile.mate(), read_collector.py:185,400)
            by_name: Dict[str, list] = {}
            for key, r in recs.items():
                by_name.setdefault(r.query_name, []).append(r)
            for r in list(recs.values()):
                if not (r.flag & 0x1) or (r.flag & 0x8) or (r.flag & 0x900):
                    continue
                want = 0x80 if (r.flag & 0x40) else 0x40
                if any((m.flag & want) and not (m.flag & 0x900) for m in by_name[r.query_name]):
                    continue
                try:
                    m = bam.mate(r)
                except ValueError:
                    continue
                if getattr(m, "reference_id", 0) != getattr(r, "reference_id", 0):
                    continue
                recs.setdefault(ident(m), m)
                by_name[r.query_name].append(m)
            rl = sorted(recs.values(), key=lambda r: (r.reference_start,))
            if contig not in cindex:
                cindex[contig] = len(contigs)
                contigs.append(contig)
            blocks.append((k, cindex[contig], rl))
    n = sum(len(b[2]) for b in blocks)
    hdr = np.zeros(n, dtype=READ_HDR)
    names: List[str] = []
    cig: List[int] = []
    quals: List[np.ndarray] = []
    codes: List[np.ndarray] = []
    offs = [0]
    qoff = 0
    i = 0
    for k, c, rl in blocks:
        base = i
        index = {}
        for j, r in enumerate(rl):
            index.setdefault(r.query_name, []).append(base + j)
        for r in rl:
            h = hdr[i]
            h["start"], h["tlen"], h["flag"], h["mapq"] = r.reference_start, r.tlen, r.flag, r.mapping_quality
            words = _cigar_words(r.cigartuples)
            h["cigar_off"], h["n_cigar"] = len(cig), len(words)
            cig += words
            seq = r.query_sequence or ""
            q = np.array(r.query_qualities if r.query_qualities is not None else [], dtype=np.uint8)
            code, esc = encode_bases(seq)
            if q.shape[0] != code.shape[0]:
                q = np.zeros(code.shape[0], dtype=np.uint8)
            q = np.where(esc, q | QUAL_ESCAPE, q & 0x7F).astype(np.uint8)
            h["qoff_lo"], h["qoff_hi"], h["l_seq"] = qoff & 0xFFFFFFFF, qoff >> 32, code.shape[0]
            qoff += code.shape[0]
            quals.append(q)
            codes.append(code)
            aux = AUX_SAME_REF if r.next_reference_id == r.reference_id else 0
            if r.has_tag("SA"):
                aux |= AUX_HAS_SA
            h["aux"] = aux
            names.append(r.query_name)
            i += 1
        # mate pointers: same name, the other segment, primary alignment
        for j, r in enumerate(rl):
            m = -1
            if (r.flag & 0x1) and not (r.flag & 0x8):
                want = 0x80 if (r.flag & 0x40) else 0x40
                for cand in index[r.query_name]:
                    f = int(hdr["flag"][cand])
                    if cand != base + j and (f & want) and not (f & 0x900):
                        m = cand
                        break
            hdr["mate"][base + j] = m
        offs.append(i)
    table = ReadTable(
        kids=kids, contigs=contigs, blk_kid=np.array([b[0] for b in blocks], dtype=np.int32),
        blk_contig=np.array([b[1] for b in blocks], dtype=np.int32), blk_off=np.array(offs, dtype=np.int64),
        hdr=hdr, cigar=np.array(cig, dtype=np.uint32),
        qual=np.concatenate(quals) if quals else np.zeros(0, dtype=np.uint8),
        seq2=pack_seq(np.concatenate(codes) if codes else np.zeros(0, dtype=np.uint8)), names=names)
    table.head_tlen = head_tlen
    table.validate()
    return table
