"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference's phasing hot path.

Pure Python over the columnar tables of ``unfazed_b200.schema``.  It exists because the reference
itself cannot travel to the GPU box (``/root/reference`` is only mounted in the build container):
this file is the checker the ``-m gpu`` parity tests, ``smoke()`` and ``bench.py``'s
``cpu_baseline`` leg use there.  It is pinned against the real reference here -- see
``tests/test_port_vs_reference.py`` (runs the unmodified reference over ``oracle/fakes.py``) and the
committed fixtures in ``tests/golden`` produced by ``oracle/make_golden.py``.

It is NOT part of the product: nothing under ``unfazed_b200/`` imports it.

Every function cites the reference lines it restates (paths relative to the reference repo).
Reads are integers (rows of the ReadTable); sites are small dicts exactly like the reference's.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from unfazed_b200.schema import (
    AUX_HAS_SA, AUX_SAME_REF, BASE_CHARS, HET, HOM_ALT, HOM_REF, QUAL_ESCAPE, ReadTable,
    SITE_FLAG_SIMPLE, SiteTable,
)

SV_TYPES = ["DEL", "DUP", "INV", "CNV", "DUP:TANDEM", "DEL:ME", "CPX", "CTX"]   # utils.py:8
SNV_TYPES = ["POINT", "SNV", "INDEL"]                                            # utils.py:9
CIGAR_LETTERS = "MIDNSHP=XB"                                                     # utils.py:13-24

# utils.py:26-43 -- reproduced as written (the names are swapped w.r.t. the assemblies, Q6)
PAR = {
    "37": ({"x": (10001, 2781479), "y": (10001, 2781479)},
           {"x": (155701383, 156030895), "y": (56887903, 57217415)}),
    "38": ({"x": (60001, 2699520), "y": (10001, 2649520)},
           {"x": (154931044, 155260560), "y": (59034050, 59363566)}),
}


@dataclass
class Params:
    """CLI defaults of the reference, __main__.py:75-223."""
    threads: int = 1
    build: str = "38"
    no_extended: bool = False
    multiread_proc_min: int = 1000
    ab_homref: Tuple[float, float] = (0.0, 0.2)
    ab_homalt: Tuple[float, float] = (0.8, 1.0)
    ab_het: Tuple[float, float] = (0.2, 0.8)
    min_gt_qual: int = 20
    min_depth: int = 10
    search_dist: int = 5000
    insert_size_max_sample: int = 1000000
    stdevs: int = 3
    min_map_qual: int = 1
    readlen: int = 151
    split_error_margin: int = 5
    evidence_min_ratio: int = 10


class ReferenceUndefined(Exception):
    """The reference raises on this input (e.g. Q12 KeyError in find_many CNV mode)."""


# ==============================================================================================
# sites
# ==============================================================================================

def strip_chr(s: str) -> str:
    return s.strip("chr")


def vcf_prefix(sites: SiteTable) -> str:
    """utils.py:46-52 applied to the first record of the joint VCF (assumed stable)."""
    if sites.n_rows == 0:
        return ""
    # first record in (contig, pos) order
    best = None
    for b in range(sites.n_blocks):
        lo, hi = int(sites.blk_off[b]), int(sites.blk_off[b + 1])
        if hi > lo:
            k = (int(sites.blk_contig[b]), int(sites.pos[lo]), b)
            if best is None or k < best:
                best = k
    name = sites.contigs[best[0]]
    return name[:3] if "chr" in name.lower() else ""


def autophaseable(dn: dict, pedigrees: dict, build: str) -> bool:
    """informative_site_finder.py:137-164 (same test in snv_phaser.py:302-328)."""
    chrom = strip_chr(dn["chrom"].lower())
    if chrom not in ("x", "y"):
        return False
    if int(pedigrees[dn["kid"]]["sex"]) != 1:
        return False
    if build not in PAR:
        return False
    par1, par2 = PAR[build]
    s = dn["start"]
    if par1[chrom][0] <= s <= par1[chrom][1] or par2[chrom][0] <= s <= par2[chrom][1]:
        return False
    return True


def high_quality(sites: SiteTable, m: int, row: int, p: Params) -> bool:
    """is_high_quality_site, informative_site_finder.py:46-73."""
    g = int(sites.gt[m, row])
    if g == HOM_REF:
        lo, hi = p.ab_homref
    elif g == HOM_ALT:
        lo, hi = p.ab_homalt
    elif g == HET:
        lo, hi = p.ab_het
    else:
        return False
    if float(sites.gq[m, row]) < p.min_gt_qual:
        return False
    rd, ad = sites.rd[m, row], sites.ad[m, row]            # numpy int32 scalars
    with np.errstate(all="ignore"):
        tot = np.int32(rd + ad)
        if tot < p.min_depth:
            return False
        ab = float(np.float64(ad) / np.float64(tot))
    return bool(lo <= ab <= hi)


def kid_allele(sites: SiteTable, row: int, vartype: str, p: Params) -> Optional[str]:
    """get_kid_allele, informative_site_finder.py:76-134."""
    rd = [int(sites.rd[m, row]) for m in range(3)]
    ad = [int(sites.ad[m, row]) for m in range(3)]
    g = int(sites.gt[0, row])
    if vartype == "DEL" and (rd[0] + ad[0]) > 4:
        if g == HOM_ALT:
            return "ref_parent"
        if g == HOM_REF:
            return "alt_parent"
        return None
    if vartype == "DUP" and rd[0] > 2 and ad[0] > 2 and (rd[0] + ad[0]) > p.min_depth:
        if g != HET:
            return None
        with np.errstate(all="ignore"):
            bal = [float(np.float64(ad[m]) / np.float64(rd[m] + ad[m])) for m in range(3)]
        k, d, m_ = bal
        if ((d + m_) < 1 and k > 0.5) or ((d + m_) > 1 and k < 0.5):
            return None
        if k >= 0.67:
            return "alt_parent"
        if k <= 0.33:
            return "ref_parent"
        return None
    return None


def classify_row(sites: SiteTable, row: int, dn: dict, dad: str, mom: str, p: Params,
                 whole_region: bool):
    """The per (DNM, record) body shared by ``find`` (informative_site_finder.py:252-339) and
    ``add_good_candidate_variant`` (:457-542).  Returns (het_site or None, candidate or None)."""
    pos = int(sites.pos[row])
    if (dn["end"] - dn["start"]) < 20 and dn["start"] <= pos < dn["end"]:
        return None, None
    ref, alts = sites.ref_alts(row)
    gk, gd, gm = (int(sites.gt[m, row]) for m in range(3))
    dad_ok = None

    def parents_ok():
        nonlocal dad_ok
        if dad_ok is None:
            dad_ok = high_quality(sites, 1, row, p) and high_quality(sites, 2, row, p)
        return dad_ok

    het = None
    if gk == HET and parents_ok():
        het = {"pos": pos, "ref_allele": ref, "alt_allele": alts[0]}
    cand = {"pos": pos, "ref_allele": ref, "alt_allele": alts[0]}
    if whole_region and ("vartype" in dn):
        cand["kid_allele"] = kid_allele(sites, row, dn["vartype"], p)
        if not cand["kid_allele"]:
            return het, None
    elif gk != HET or not high_quality(sites, 0, row, p):
        return het, None
    if not parents_ok():
        return het, None
    if gd in (HET, HOM_ALT) and gm == HOM_REF:
        cand["alt_parent"], cand["ref_parent"] = dad, mom
    elif gm in (HET, HOM_ALT) and gd == HOM_REF:
        cand["alt_parent"], cand["ref_parent"] = mom, dad
    elif gm == HET and gd == HOM_ALT:
        cand["alt_parent"], cand["ref_parent"] = dad, mom
    elif gd == HET and gm == HOM_ALT:
        cand["alt_parent"], cand["ref_parent"] = mom, dad
    else:
        return het, None
    if gk in (HOM_ALT, HOM_REF):
        pg = (gd, gm)
        if HET in pg and (HOM_ALT in pg or HOM_REF in pg):
            for g in pg:
                if g in (HOM_ALT, HOM_REF) and gk == g:
                    return het, None
    return het, cand


def _trio_index(sites: SiteTable, pedigrees: dict, kid: str) -> int:
    """Trio whose three sample columns are all present (the reference's sample_dict test)."""
    ped = pedigrees[kid]
    for t, trio in enumerate(sites.trios):
        if trio == (kid, ped["dad"], ped["mom"]):
            return t
    return -1


def _simple_rows(sites: SiteTable, blk: int, lo: int, hi: int) -> List[int]:
    """Rows of a block passing the prefilter (:239-244) with lo <= pos <= hi, in file order."""
    if blk < 0:
        return []
    a, b = int(sites.blk_off[blk]), int(sites.blk_off[blk + 1])
    pos = sites.pos[a:b]
    i0 = int(np.searchsorted(pos, lo, side="left"))
    i1 = int(np.searchsorted(pos, hi, side="right"))
    return [a + i for i in range(i0, i1) if sites.flag[a + i] & SITE_FLAG_SIMPLE]


def find(dnms: List[dict], pedigrees: dict, sites: SiteTable, p: Params, search_dist: int,
         whole_region: bool = True):
    """informative_site_finder.py:167-344."""
    if len(dnms) >= p.multiread_proc_min:
        return find_many(dnms, pedigrees, sites, p, search_dist, whole_region)
    if len(dnms) <= 0:
        return None
    prefix = vcf_prefix(sites)
    for dn in dnms:
        if autophaseable(dn, pedigrees, p.build):
            continue
        trio = _trio_index(sites, pedigrees, dn["kid"])
        if trio < 0:
            continue
        dad, mom = pedigrees[dn["kid"]]["dad"], pedigrees[dn["kid"]]["mom"]
        blk = sites.block_of(trio, prefix + strip_chr(dn["chrom"]))
        s, e = int(dn["start"]), int(dn["end"])
        if whole_region:                                   # get_position :10-43
            regions = [(s - search_dist, e + search_dist)]
        else:
            regions = [(s - search_dist, s + search_dist)]
            if (e - s) > search_dist:
                regions.append((e - search_dist, e + search_dist))
        cands, hets = [], []
        for B, E in regions:
            for row in _simple_rows(sites, blk, B - 1, E - 1):
                het, cand = classify_row(sites, row, dn, dad, mom, p, whole_region)
                if het:
                    hets.append(het)
                if cand:
                    cands.append(cand)
        dn["candidate_sites"] = sorted(cands, key=lambda x: x["pos"])
        dn["het_sites"] = sorted(hets, key=lambda x: x["pos"])
    return dnms


def find_many(dnms: List[dict], pedigrees: dict, sites: SiteTable, p: Params, search_dist: int,
              whole_region: bool = True):
    """informative_site_finder.py:347-661 (create_lookups, get_close_vars,
    add_good_candidate_variant, multithread_find_many, find_many)."""
    by_loc: Dict[str, Dict[int, List[str]]] = {}
    by_sample: Dict[str, Dict[str, Dict[int, List[dict]]]] = {}
    ranges: Dict[str, List[int]] = {}
    auto, rest = [], []
    for dn in dnms:                                        # create_lookups :347-396
        if autophaseable(dn, pedigrees, p.build):
            auto.append(dn)
            continue
        rest.append(dn)
        c, s, e, k = dn["chrom"], int(dn["start"]), int(dn["end"]), dn["kid"]
        r = ranges.setdefault(c, [s, e])
        r[0], r[1] = min(r[0], s), max(r[1], e)
        by_sample.setdefault(k, {}).setdefault(c, {}).setdefault(s, []).append(dn)
        by_loc.setdefault(c, {}).setdefault(s, []).append(k)
        if (e - s) > 2:
            by_loc[c].setdefault(e, []).append(k)
    prefix = vcf_prefix(sites)
    for chrom in {dn["chrom"] for dn in rest}:
        try:
            _find_many_chrom(chrom, ranges[chrom], by_loc, by_sample, sites, pedigrees, p,
                             search_dist, whole_region, prefix)
        except KeyError:
            if p.threads == 1:                             # Q12 / Q28
                raise ReferenceUndefined("find_many KeyError on chromosome %s" % chrom)
    out = []
    for k in by_sample:
        for c in by_sample[k]:
            for s in by_sample[k][c]:
                for dn in by_sample[k][c][s]:
                    for key in ("candidate_sites", "het_sites"):
                        if key in dn:
                            dn[key] = sorted(dn[key], key=lambda x: x["pos"])
                    out.append(dn)
    return out + auto


def _find_many_chrom(chrom, rng, by_loc, by_sample, sites, pedigrees, p, sd, whole_region, prefix):
    contig = prefix + strip_chr(chrom)
    if contig not in sites.contigs:
        return
    B, E = rng[0] - sd, rng[1] + sd
    # joint-VCF record order: (pos, block)
    recs = []
    for blk in range(sites.n_blocks):
        if sites.contigs[int(sites.blk_contig[blk])] == contig:
            recs += [(int(sites.pos[r]), blk, r) for r in _simple_rows(sites, blk, B - 1, E - 1)]
    recs.sort()
    for pos0, blk, row in recs:
        POS = pos0 + 1
        keys = []                                          # get_close_vars :399-420
        if contig in by_loc:
            if not whole_region:
                for loc, kids in by_loc[contig].items():
                    if (loc - sd) <= POS <= (loc + sd):
                        keys += [(k, contig, loc) for k in kids]
            else:
                for loc, kids in by_loc[contig].items():
                    for k in kids:
                        for dn in by_sample[k][contig][loc]:        # KeyError: Q12
                            if (int(dn["start"]) - sd) <= POS <= (int(dn["end"]) + sd):
                                keys.append((k, contig, int(dn["start"])))
        for kid, c, loc in keys:                           # add_good_candidate_variant :442-543
            trio = _trio_index(sites, pedigrees, kid)
            if trio < 0 or loc not in by_sample[kid][c]:
                continue
            if int(sites.blk_trio[blk]) != trio:
                continue                                   # other trio's columns: all UNKNOWN
            dad, mom = pedigrees[kid]["dad"], pedigrees[kid]["mom"]
            for dn in by_sample[kid][c][loc]:
                if autophaseable(dn, pedigrees, p.build):
                    continue
                het, cand = classify_row(sites, row, dn, dad, mom, p, whole_region)
                if het:
                    dn.setdefault("het_sites", []).append(het)
                if cand:
                    dn.setdefault("candidate_sites", []).append(cand)


def get_refalt(sites: SiteTable, trio: int, chrom: str, pos: int, prefix: str):
    """snv_phaser.py:73-84: records overlapping 1-based [pos, pos+1] -> (REF of first, all ALTs).
    The joint VCF returns every record there, whichever trio its genotypes belong to."""
    contig = prefix + strip_chr(chrom)
    rid = sites.record_ids()
    rows = {}
    for b in range(sites.n_blocks):
        if sites.contigs[int(sites.blk_contig[b])] != contig:
            continue
        a, z = int(sites.blk_off[b]), int(sites.blk_off[b + 1])
        for r in range(a, z):
            q = int(sites.pos[r])
            if q >= pos + 1:
                break
            ref, _ = sites.ref_alts(r)
            if q + len(ref) > pos - 1:
                rows.setdefault((q, int(rid[r])), r)
    ref, alts = None, []
    for key in sorted(rows):
        rr, aa = sites.ref_alts(rows[key])
        if ref is None:
            ref = rr
        alts += list(aa)
    return ref, alts


# ==============================================================================================
# reads
# ==============================================================================================

class Bam:
    """The kid's alignments: what pysam gives the reference, from the ReadTable."""

    def __init__(self, reads: ReadTable, kid: int):
        self.t = reads
        self.kid = kid
        h = reads.hdr
        self.start = h["start"]
        self.end = reads.ref_ends()
        self._cache: Dict[int, tuple] = {}
        self._max_span: Dict[int, int] = {}

    def all_reads(self):
        t = self.t
        for b in range(t.n_blocks):
            if int(t.blk_kid[b]) == self.kid:
                yield from range(int(t.blk_off[b]), int(t.blk_off[b + 1]))

    def fetch(self, contig: str, start, stop) -> List[int]:
        """pysam fetch(contig, start, stop): reads overlapping the 0-based half-open interval, in file order.  Like
        an indexed BAM it does not walk the contig from its beginning: no read is longer than the longest span
        of its block, so the candidates start in [start - max_span, stop)."""
        t = self.t
        if contig not in t.contigs:
            raise ValueError(contig)
        start, stop = int(start), int(stop)
        b = t.block_of(self.kid, contig)
        if b < 0:
            return []
        lo, hi = int(t.blk_off[b]), int(t.blk_off[b + 1])
        span = self._max_span.get(b)
        if span is None:
            span = self._max_span[b] = int((self.end[lo:hi] - self.start[lo:hi]).max()) if hi > lo else 0
        j0 = int(np.searchsorted(self.start[lo:hi], start - span, side="left"))
        j1 = int(np.searchsorted(self.start[lo:hi], stop, side="left"))
        sel = np.nonzero(self.end[lo + j0:lo + j1] > start)[0]
        return [lo + j0 + int(i) for i in sel]

    def mate(self, r: int) -> int:
        h = self.t.hdr[r]
        if not (int(h["flag"]) & 0x1) or (int(h["flag"]) & 0x8) or int(h["mate"]) < 0:
            raise ValueError("no mate")
        return int(h["mate"])

    def decoded(self, r: int):
        """(cigartuples, reference positions full_length=True, query_sequence, qualities)."""
        d = self._cache.get(r)
        if d is None:
            t = self.t
            h = t.hdr[r]
            o, n = int(h["cigar_off"]), int(h["n_cigar"])
            cig = [(int(w) & 15, int(w) >> 4) for w in t.cigar[o:o + n]]
            pos, cur = [], int(h["start"])
            for op, ln in cig:
                if op in (0, 7, 8):
                    pos.extend(range(cur, cur + ln))
                    cur += ln
                elif op in (1, 4):
                    pos.extend([None] * ln)
                elif op in (2, 3):
                    cur += ln
            q0, L = t.qoff(r), int(h["l_seq"])
            qb = t.qual[q0:q0 + L]
            g = np.arange(q0, q0 + L)
            code = (t.seq2[g >> 2] >> ((g & 3) << 1).astype(np.uint8)) & 3
            ch = np.frombuffer(BASE_CHARS.encode(), dtype=np.uint8)[code]
            esc = (qb & QUAL_ESCAPE) != 0
            ch = np.where(esc, np.where(code == 0, ord("N"), ord("?")), ch).astype(np.uint8)
            d = (cig, pos, ch.tobytes().decode("ascii"), (qb & 0x7F).tolist())
            self._cache[r] = d
        return d

    def name(self, r: int) -> str:
        return self.t.name_of(r)


def estimate_concordant_insert_len(bam: Bam, p: Params):
    """read_collector.py:11-25 (Q14: the percentile collapses to a scalar)."""
    vals = []
    tl = bam.t.hdr["tlen"]
    for i, r in enumerate(bam.all_reads()):
        vals.append(abs(int(tl[r]) - p.readlen * 2))
        if i >= p.insert_size_max_sample:
            break
    pct = np.percentile(np.array(vals), 99.5)
    return int(np.mean(pct)) + (np.std(pct) * p.stdevs)


def goodread(bam: Bam, r: int, p: Params, discordant: bool = False) -> bool:
    """read_collector.py:28-53 (Q3: every CIGAR op counts as a 'mismatch')."""
    h = bam.t.hdr[r]
    f = int(h["flag"])
    if (f & (0x200 | 0x4 | 0x400 | 0x100 | 0x800 | 0x8)) or int(h["mapq"]) < p.min_map_qual \
            or not (int(h["aux"]) & AUX_SAME_REF):
        return False
    if not discordant:
        cig, _pos, _seq, quals = bam.decoded(r)
        low = sum(1 for q in quals if q < p.min_gt_qual)
        if low > 10 or len(cig) > 10:
            return False
    return True


def get_allele_at(bam: Bam, read: int, mate: Optional[int], pos: int, n: int, p: Params):
    """read_collector.py:56-73."""
    _c, rp, seq, _q = bam.decoded(read)
    if pos in rp:
        q = rp.index(pos)
        if q < 4 or q > (p.readlen - 4):
            return False
        if len(seq) > q + n:
            return seq[q:q + n]
        return False
    if mate is not None:
        _c, mp, mseq, _q = bam.decoded(mate)
        if pos in mp:
            q = mp.index(pos)
            if q < 4 or q > (p.readlen - 4):
                return False
            if len(mseq) > q + n:
                return mseq[q:q + n]
    return False


def _pair_filters(bam: Bam, read: int, mate: int) -> bool:
    """None-count and mate-overlap tests, read_collector.py:198-214 == :405-418."""
    rp = bam.decoded(read)[1]
    mp = bam.decoded(mate)[1]
    if rp.count(None) > 5 or mp.count(None) > 5:
        return False
    r0, r1 = int(bam.start[read]), int(bam.end[read])
    m0, m1 = int(bam.start[mate]), int(bam.end[mate])
    if m0 <= r0 <= m1 or m0 <= r1 <= m1:
        return False
    return True


def binary_search(start: int, end: int, site_list: List[dict]) -> List[dict]:
    """site_searcher.py:6-47, including its pivot choice and neighbour rule (Q16)."""
    lo, hi = 0, len(site_list) - 1
    plo = phi = -1
    out: List[dict] = []
    while not out and hi > -1:
        if lo > hi or (lo == plo and hi == phi):
            break
        plo, phi = lo, hi
        mid = int((hi + lo) / 2)
        x = site_list[mid]["pos"]
        if start <= x < end:
            out.append(site_list[mid])
            j = mid + 1
            while j < len(site_list) and start <= site_list[j]["pos"] <= end:
                out.append(site_list[j])
                j += 1
            j = mid - 1
            while j >= 0 and start <= site_list[j]["pos"] <= end:
                out.append(site_list[j])
                j -= 1
            break
        elif x > start:
            hi = mid - 1
        elif x < start:
            lo = mid + 1
    return out


def connect_reads(bam, grouped, read_sites, site_reads, frontier, fetched, p: Params, level=0):
    """read_collector.py:76-152.  ``frontier`` is {"alt": [...], "ref": [...]} in the dict order
    the reference uses at this depth (Q19)."""
    nxt = {"ref": [], "alt": []}
    for hap in frontier:
        other = "ref" if hap == "alt" else "alt"
        for name, found_pos in frontier[hap]:
            if name not in read_sites:
                continue
            for site in read_sites[name]:
                if site["pos"] == found_pos:
                    continue
                fa = get_allele_at(bam, fetched[name][0], fetched[name][1], site["pos"], 1, p)
                nfa = None
                if fa:
                    if fa == site["ref_allele"]:
                        nfa = site["alt_allele"]
                    elif fa == site["alt_allele"]:
                        nfa = site["ref_allele"]
                if not (fa and nfa):
                    continue
                for other_name in site_reads[site["pos"]]:
                    if other_name in grouped["ref"] or other_name in grouped["alt"]:
                        continue
                    rd, mt = fetched[other_name]
                    na = get_allele_at(bam, rd, mt, site["pos"], 1, p)
                    if not na:
                        continue
                    _c, rp, _s, quals = bam.decoded(rd)
                    if site["pos"] not in rp:
                        continue
                    if quals[rp.index(site["pos"])] < p.min_gt_qual:
                        continue
                    if na == fa:
                        nxt[hap].append([other_name, site["pos"]])
                        grouped[hap].add(other_name)
                    elif na == nfa:
                        nxt[other].append([other_name, site["pos"]])
                        grouped[other].add(other_name)
    if nxt["alt"] or nxt["ref"]:
        return connect_reads(bam, grouped, read_sites, site_reads, nxt, fetched, p, level + 1)
    return grouped


def group_reads_by_haplotype(bam: Bam, region: dict, seeds: Dict[str, List[int]],
                             het_sites: List[dict], cul, p: Params):
    """read_collector.py:155-263.  Returns ({"ref": [reads], "alt": [reads]}, labels) where
    labels maps read name -> haplotype for every name in the connected sets."""
    fetched: Dict[str, List[int]] = {}
    read_sites: Dict[str, List[dict]] = {}
    site_reads: Dict[int, List[str]] = {}
    het_site = None
    for het_site in het_sites:
        try:
            it = bam.fetch(region["chrom"], het_site["pos"], het_site["pos"] + 1)
        except ValueError:
            c = region["chrom"]
            c = strip_chr(c) if "chr" in c else "chr" + c
            it = bam.fetch(c, het_site["pos"], het_site["pos"] + 1)
        for i, r in enumerate(it):
            if i > p.insert_size_max_sample:                # EXTENDED_RB_READ_GOAL, Q2
                continue
            ins = abs(int(bam.t.hdr["tlen"][r]) - p.readlen * 2)
            if not (goodread(bam, r, p) and ins <= cul):
                continue
            try:
                m = bam.mate(r)
            except ValueError:
                continue
            if not goodread(bam, m, p):
                continue
            if sum(1 for op, _l in bam.decoded(r)[0] if op not in (0, 7)) > 5:
                continue
            if not _pair_filters(bam, r, m):
                continue
            nm = bam.name(r)
            read_sites.setdefault(nm, []).append(het_site)
            site_reads.setdefault(het_site["pos"], []).append(nm)
            fetched[nm] = [r, m]
    grouped = {"ref": set(), "alt": set()}
    frontier = {"alt": [], "ref": []}
    for hap in ("ref", "alt"):
        for r in seeds[hap]:
            nm = bam.name(r)
            grouped[hap].add(nm)
            frontier[hap].append([nm, -1])
            try:
                m = bam.mate(r)
            except ValueError:
                continue
            fetched[nm] = [r, m]
            hits = binary_search(int(bam.start[r]), int(bam.end[r]), het_sites)
            if not hits:
                continue
            read_sites.setdefault(nm, [])
            site_reads.setdefault(het_site["pos"], [])       # stale loop variable, Q17
            for h in hits:
                read_sites[nm].append(h)
                site_reads[het_site["pos"]].append(nm)
    grouped = connect_reads(bam, grouped, read_sites, site_reads, frontier, fetched, p)
    out = {"ref": [], "alt": []}
    labels = {}
    for hap in grouped:
        for nm in grouped[hap]:
            labels[nm] = hap
            if nm in fetched:
                out[hap] += fetched[nm]
    return out, labels


def _seed_snv(bam, out, read, mate, ref, alt, position, p):
    """snv_match_alleles, read_collector.py:296-336."""
    n = max(len(ref), len(alt))
    a = get_allele_at(bam, read, mate, position, n, p)
    if not a:
        return
    if len(ref) >= len(alt):
        if a == ref:
            out["ref"] += [read, mate]
        elif a[:len(alt)] == alt:
            out["alt"] += [read, mate]
    else:
        if a == alt:
            out["alt"] += [read, mate]
        elif a[:len(ref) + 1] == ref:
            out["ref"] += [read, mate]


def _seed_indel(bam, out, read, mate, ref, alt, position, p):
    """indel_match_alleles, read_collector.py:266-293 (Q21: op expansion of *all* ops)."""
    n = max(len(ref), len(alt))
    cig, rp, _seq, quals = bam.decoded(read)
    if position not in rp:
        return
    q = rp.index(position)
    ops = []
    for op, ln in cig:
        ops += [CIGAR_LETTERS[op]] * ln
    for ql in quals[q:q + n]:
        if ql < p.min_gt_qual:
            return
    window = ops[q:q + n]
    if "I" in window or "D" in window:
        out["alt"] += [read, mate]
    elif 7 < q < (len(rp) - 7):
        out["ref"] += [read, mate]


def collect_reads_snv(bam: Bam, region: dict, het_sites, ref: str, alt: str, cul, p: Params):
    """read_collector.py:339-432."""
    if not cul:
        cul = estimate_concordant_insert_len(bam, p)
    position = int(region["start"])
    try:
        it = bam.fetch(region["chrom"], position - 1, position + 1)
    except ValueError:
        c = region["chrom"]
        c = strip_chr(c) if "chr" in c else "chr" + c
        it = bam.fetch(c, position, position + 1)           # Q24: different window
    seeds = {"alt": [], "ref": []}
    for r in it:
        ins = abs(int(bam.t.hdr["tlen"][r]) - p.readlen * 2)
        if not goodread(bam, r, p) or ins > cul:
            continue
        try:
            m = bam.mate(r)
        except ValueError:
            continue
        if not goodread(bam, m, p):
            continue
        if not _pair_filters(bam, r, m):
            continue
        if len(ref) == len(alt):
            _seed_snv(bam, seeds, r, m, ref, alt, position, p)
        else:
            _seed_indel(bam, seeds, r, m, ref, alt, position, p)
    if p.no_extended:
        labels = {}
        for hap in ("alt", "ref"):
            for r in seeds[hap]:
                labels[bam.name(r)] = hap
        return seeds, cul, labels
    grouped, labels = group_reads_by_haplotype(bam, region, seeds, het_sites, cul, p)
    return grouped, cul, labels


def collect_reads_sv(bam: Bam, region: dict, het_sites, cul, p: Params):
    """read_collector.py:435-602."""
    if not cul:
        cul = estimate_concordant_insert_len(bam, p)
    support: List[int] = []
    var_len = abs(float(region["end"]) - float(region["start"]))
    banned: List[str] = []
    for position in (region["start"], region["end"]):
        position = int(position)
        try:
            it = bam.fetch(region["chrom"], max(0, position - cul), position + cul)
        except ValueError:
            c = region["chrom"]
            c = c.replace("chr", "") if "chr" in c else "chr" + c
            it = bam.fetch(c, max(0, position - cul), position + cul)
        banned = []
        for r in it:
            nm = bam.name(r)
            if nm in banned:
                continue
            if not goodread(bam, r, p, True):
                continue
            try:
                m = bam.mate(r)
            except ValueError:
                continue
            ins = abs(int(bam.t.hdr["tlen"][r]) - p.readlen * 2)
            if not goodread(bam, m, p, True):
                continue
            cig, rp, _s, _q = bam.decoded(r)
            ops = []
            for op, ln in cig:
                ops += [CIGAR_LETTERS[op]] * ln
            s_m = ops[:10].count("M") + ops[:10].count("=")
            e_m = ops[-10:].count("M") + ops[-10:].count("=")
            if e_m < 7 and s_m < 7:
                banned.append(nm)
                continue
            r0, r1 = int(bam.start[r]), int(bam.end[r])
            em = p.split_error_margin
            if int(bam.t.hdr["aux"][r]) & AUX_HAS_SA:
                if (position - em) <= r0 <= (position + em) or (position - em) <= r1 <= (position + em):
                    support += [r, m]
            elif ins > cul and 0.7 < abs(var_len / ins) < 1.3:
                m0, m1 = int(bam.start[m]), int(bam.end[m])
                left0 = min(m0, r0)
                right0 = max(m0, r0)
                w = int(cul)
                if not ((region["start"] - w) < left0 < (region["start"] + w)
                        and (region["end"] - w) < right0 < (region["end"] + w)):
                    continue
                support += [m, r]
            else:
                if position in rp:
                    k = rp.index(position)
                elif position - 1 in rp:
                    k = rp.index(position - 1)
                elif position + 1 in rp:
                    k = rp.index(position + 1)
                else:
                    continue
                if k < 2 or k > (len(rp) - 4):
                    continue
                before = list(set(rp[:k - 1]))
                after = list(set(rp[k + 1:]))
                if (len(before) == 1 and before[0] is None) or (len(after) == 1 and after[0] is None):
                    support += [m, r]
    kept = [r for r in support if bam.name(r) not in banned]
    if len(kept) < 2:
        return {"alt": [], "ref": []}, cul, {}
    seeds = {"alt": kept, "ref": []}
    if p.no_extended:
        return seeds, cul, {bam.name(r): "alt" for r in kept}
    grouped, labels = group_reads_by_haplotype(bam, region, seeds, het_sites, cul, p)
    return grouped, cul, labels


def match_informative_sites(bam: Bam, reads: Dict[str, List[int]], cand: List[dict]):
    """site_searcher.py:50-78."""
    out = {}
    for hap in reads:
        out[hap] = []
        for r in reads[hap]:
            ms = binary_search(int(bam.start[r]), int(bam.end[r]), cand)
            if ms:
                if len({m["ref_parent"] for m in ms}) != 1 or len({m["alt_parent"] for m in ms}) != 1:
                    continue
                out[hap].append({"matches": ms, "read": r})
    return out


def phase_by_reads(bam: Bam, matches):
    """snv_phaser.py:16-70 == sv_phaser.py:14-68.  Returns parent id -> [(read, pos)]."""
    ev: Dict[str, list] = {}
    for hap in matches:
        for mi in matches[hap]:
            r = mi["read"]
            _c, rp, seq, _q = bam.decoded(r)
            for m in mi["matches"]:
                if not ev:
                    ev[m["ref_parent"]] = []
                    ev[m["alt_parent"]] = []
                if m["pos"] not in rp:
                    continue
                base = seq[rp.index(m["pos"])]
                if base == m["ref_allele"]:
                    origin_is_ref = True
                elif base == m["alt_allele"]:
                    origin_is_ref = False
                else:
                    continue
                # credited to alt_parent iff (origin is ref_parent) XOR (read haplotype is "alt")
                to_alt = origin_is_ref == (hap == "ref")
                ev[m["alt_parent"] if to_alt else m["ref_parent"]].append((r, m["pos"]))
    return ev


# ==============================================================================================
# drivers (snv_phaser.py / sv_phaser.py) and the final call (unfazed.py:190-334)
# ==============================================================================================

def _key(dn: dict) -> str:
    return "_".join([str(dn["chrom"]), str(dn["start"]), str(dn["end"]), dn["kid"], dn["vartype"]])


def _auto_record(dn, dad, mom):
    return {
        "region": {"chrom": dn["chrom"], "start": dn["start"], "end": dn["end"]},
        "vartype": dn["vartype"], "kid": dn["kid"], "dad": dad, "mom": mom,
        "cnv_dad_sites": "NA", "cnv_mom_sites": "NA", "cnv_evidence_type": "SEX-CHROM",
        "dad_sites": "", "mom_sites": "", "evidence_type": "SEX-CHROM",
        "dad_reads": [], "mom_reads": [],
    }


def _read_record(bam, dn, dad, mom, ev):
    def uniq(parent):
        items = ev.get(parent, [])
        return sorted({str(pos) for _r, pos in items}), sorted({bam.name(r) for r, _pos in items})
    ds, dr = uniq(dad)
    ms, mr = uniq(mom)
    return {
        "region": {"chrom": dn["chrom"], "start": dn["start"], "end": dn["end"]},
        "vartype": dn["vartype"], "kid": dn["kid"], "dad": dad, "mom": mom,
        "dad_sites": ds, "mom_sites": ms, "evidence_type": "readbacked",
        "dad_reads": dr, "mom_reads": mr,
        "cnv_dad_sites": "", "cnv_mom_sites": "", "cnv_evidence_type": "",
    }


class Phaser:
    """State shared across DNMs in one run: per-kid Bam handles and insert-size cache
    (the reference's module global ``concordant_upper_lens``, snv_phaser.py:14)."""

    def __init__(self, sites: SiteTable, reads: ReadTable, pedigrees: dict, p: Params):
        self.sites, self.reads, self.ped, self.p = sites, reads, pedigrees, p
        self.bams: Dict[str, Bam] = {}
        self.cul: Dict[str, float] = {}
        self.labels: Dict[str, Dict[str, str]] = {}     # DNM key -> read name -> haplotype
        self.prefix = vcf_prefix(sites)

    def bam(self, kid: str) -> Bam:
        if kid not in self.bams:
            self.bams[kid] = Bam(self.reads, self.reads.kids.index(kid))
        return self.bams[kid]

    # ---- snv_phaser.run_read_phasing :206-299 + multithread_read_phasing :87-203
    def phase_snvs(self, dnms: List[dict]) -> Dict[str, dict]:
        p = self.p
        annotated = find(dnms, self.ped, self.sites, p, p.search_dist, whole_region=False)
        records: Dict[str, dict] = {}
        for dn in annotated or []:
            dad, mom = self.ped[dn["kid"]]["dad"], self.ped[dn["kid"]]["mom"]
            if autophaseable(dn, self.ped, p.build):
                records[_key(dn)] = _auto_record(dn, dad, mom)
                continue
            if not dn.get("candidate_sites"):
                continue
            trio = _trio_index(self.sites, self.ped, dn["kid"])
            if trio < 0:
                continue
            ref, alts = get_refalt(self.sites, trio, dn["chrom"], dn["start"], self.prefix)
            if len(alts) != 1:
                continue
            if "het_sites" not in dn:
                continue                                   # the reference raises KeyError (swallowed)
            bam = self.bam(dn["kid"])
            region = {"chrom": dn["chrom"], "start": dn["start"], "end": dn["end"]}
            grouped, cul, labels = collect_reads_snv(
                bam, region, dn["het_sites"], ref, alts[0], self.cul.get(dn["kid"]), p)
            self.cul[dn["kid"]] = cul
            self.labels[_key(dn)] = labels
            matches = match_informative_sites(bam, grouped, dn["candidate_sites"])
            if not matches["alt"] and not matches["ref"]:
                continue
            ev = phase_by_reads(bam, matches)
            records[_key(dn)] = _read_record(bam, dn, dad, mom, ev)
        return records

    # ---- sv_phaser.phase_svs :427-493
    def phase_svs(self, dnms: List[dict]) -> Dict[str, dict]:
        p = self.p
        cnv: Dict[str, dict] = {}
        annotated = find(dnms, self.ped, self.sites, p, 0, whole_region=True)   # run_cnv_phasing
        for dn in annotated or []:
            dad, mom = self.ped[dn["kid"]]["dad"], self.ped[dn["kid"]]["mom"]
            if autophaseable(dn, self.ped, p.build):
                cnv[_key(dn)] = _auto_record(dn, dad, mom)      # Q8: flow continues
            if dn["vartype"] not in ("DEL", "DUP"):
                continue
            cs = dn.get("candidate_sites")
            if not cs:
                continue
            votes = {cs[0]["ref_parent"]: [], cs[0]["alt_parent"]: []}          # phase_by_snvs
            for s in cs:
                votes[s[s["kid_allele"]]].append(str(s["pos"]))
            cnv[_key(dn)] = {
                "region": {"chrom": dn["chrom"], "start": dn["start"], "end": dn["end"]},
                "vartype": dn["vartype"], "kid": dn["kid"], "dad": dad, "mom": mom,
                "cnv_dad_sites": votes.get(dad, []), "cnv_mom_sites": votes.get(mom, []),
                "cnv_evidence_type": "ALLELE-BALANCE", "dad_sites": "", "mom_sites": "",
                "evidence_type": "", "dad_reads": [], "mom_reads": [],
            }
        recs: Dict[str, dict] = {}
        annotated = find(dnms, self.ped, self.sites, p, p.search_dist, whole_region=False)
        for dn in annotated or []:
            dad, mom = self.ped[dn["kid"]]["dad"], self.ped[dn["kid"]]["mom"]
            if autophaseable(dn, self.ped, p.build):
                recs[_key(dn)] = _auto_record(dn, dad, mom)     # Q8: no early exit
            if not dn.get("candidate_sites"):
                continue
            if "het_sites" not in dn:
                continue
            bam = self.bam(dn["kid"])
            region = {"chrom": dn["chrom"], "start": dn["start"], "end": dn["end"]}
            grouped, cul, labels = collect_reads_sv(bam, region, dn["het_sites"], self.cul.get(dn["kid"]), p)
            self.cul[dn["kid"]] = cul
            self.labels[_key(dn)] = labels
            matches = match_informative_sites(bam, grouped, dn["candidate_sites"])
            if not matches["alt"] and not matches["ref"]:
                continue
            ev = phase_by_reads(bam, matches)
            recs[_key(dn)] = _read_record(bam, dn, dad, mom, ev)
        for k, c in cnv.items():                                                # merge :484-492
            if k not in recs:
                recs[k] = c
            else:
                recs[k]["cnv_dad_sites"] = c["cnv_dad_sites"]
                recs[k]["cnv_mom_sites"] = c["cnv_mom_sites"]
                recs[k]["evidence_type"] += "," + c["cnv_evidence_type"]
        return recs

    def phase(self, dnms: List[dict]) -> Dict[str, dict]:
        """unfazed.py:585-649: SVs first, then SNVs; SV records win on key clashes."""
        kids = set(self.ped)
        svs = [d for d in dnms if d["vartype"].upper() in SV_TYPES and d["kid"] in kids]
        snvs = [d for d in dnms if d["vartype"].upper() in SNV_TYPES and d["kid"] in kids]
        out_sv = self.phase_svs(svs) if svs else {}
        out = self.phase_snvs(snvs) if snvs else {}
        out.update(out_sv)
        return out


def summarize_record(rec: dict, include_ambiguous: bool, verbose: bool, ratio: int):
    """unfazed.py:162-334."""
    if rec["evidence_type"] == "SEX-CHROM":
        is_y = strip_chr(rec["region"]["chrom"].lower()) == "y"
        out = {
            "chrom": rec["region"]["chrom"], "start": int(rec["region"]["start"]),
            "end": int(rec["region"]["end"]), "vartype": rec["vartype"], "kid": rec["kid"],
            "origin_parent": rec["dad"] if is_y else rec["mom"],
            "other_parent": rec["mom"] if is_y else rec["dad"],
            "evidence_count": 1, "evidence_types": ["SEX-CHROM"],
        }
        if verbose:
            for k in ("origin_parent_sites", "origin_parent_reads", "other_parent_sites", "other_parent_reads"):
                out[k] = "NA"
        return out
    nd, nm = len(rec["dad_reads"]), len(rec["mom_reads"])
    origin = other = None
    o_sites, o_reads, x_sites, x_reads = [], [], [], []
    count, types, ambig = 0, [], False
    if nd > 0 and nd >= ratio * nm:
        origin, other, count = rec["dad"], rec["mom"], len(rec["dad_sites"])
        o_sites += rec["dad_sites"]; o_reads += rec["dad_reads"]
        x_sites += rec["mom_sites"]; x_reads += rec["mom_reads"]
        types.append("READBACKED")
    elif nm > 0 and nm >= ratio * nd:
        origin, other, count = rec["mom"], rec["dad"], len(rec["mom_sites"])
        o_sites += rec["mom_sites"]; o_reads += rec["mom_reads"]
        x_sites += rec["dad_sites"]; x_reads += rec["dad_reads"]
        types.append("READBACKED")
    elif nd > 0 and nm > 0:
        origin, count = rec["dad"] + "|" + rec["mom"], nd + nm
        o_sites += rec["dad_sites"]; o_reads += rec["dad_reads"]
        x_sites += rec["mom_sites"]; x_reads += rec["mom_reads"]
        types.append("AMBIGUOUS_READBACKED")
        ambig = True
    cd, cm = len(rec["cnv_dad_sites"]), len(rec["cnv_mom_sites"])
    if cd > 0 and cd >= ratio * cm:
        if origin == rec["mom"] and "READBACKED" not in types:
            origin = None
            count += cd + cm
            o_sites += rec["cnv_dad_sites"]
            x_sites = rec["cnv_mom_sites"]
            types, ambig = ["AMBIGUOUS_BOTH"], True
        else:
            origin, other, count = rec["dad"], rec["mom"], cd
            o_sites += rec["cnv_dad_sites"]; o_reads += rec["dad_reads"]
            x_sites += rec["mom_sites"]; x_reads += rec["mom_reads"]
            if "AMBIGUOUS_READBACKED" in types:
                types.remove("AMBIGUOUS_READBACKED")
                ambig = False
            types.append("ALLELE-BALANCE")
    elif cm > 0 and cm >= ratio * cd:
        if origin == rec["dad"] and "READBACKED" not in types:
            origin = None
            count += cd + cm
            o_sites += rec["cnv_dad_sites"]
            x_sites += rec["cnv_mom_sites"]
            types, ambig = ["AMBIGUOUS_BOTH"], True
        else:
            origin, other, count = rec["mom"], rec["dad"], cm
            o_sites += rec["cnv_mom_sites"]; o_reads += rec["mom_reads"]
            x_sites += rec["dad_sites"]; x_reads += rec["dad_reads"]
            if "AMBIGUOUS_READBACKED" in types:
                types.remove("AMBIGUOUS_READBACKED")
            types.append("ALLELE-BALANCE")
    elif (cd + cm) > 0 and "READBACKED" not in types:
        origin = None
        count += cd + cm
        o_sites += rec["cnv_dad_sites"]
        x_sites = rec["cnv_mom_sites"]
        types.append("AMBIGUOUS_ALLELE-BALANCE")
        ambig = True
    if (origin is None or ambig) and not include_ambiguous:
        return None
    o_sites, x_sites = sorted(o_sites), sorted(x_sites)
    out = {
        "chrom": rec["region"]["chrom"], "start": int(rec["region"]["start"]),
        "end": int(rec["region"]["end"]), "vartype": rec["vartype"], "kid": rec["kid"],
        "origin_parent": origin, "other_parent": other, "evidence_count": count,
        "evidence_types": types,
    }
    if verbose:
        out["origin_parent_sites"] = ",".join(o_sites) if o_sites else "-"
        out["origin_parent_reads"] = ",".join(o_reads) if o_reads else "-"
        out["other_parent_sites"] = ",".join(x_sites) if x_sites else "-"
        out["other_parent_reads"] = ",".join(x_reads) if x_reads else "-"
    return out
