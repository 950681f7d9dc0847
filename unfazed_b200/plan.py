"""Host-side window planner: turns the reference's site-finding *membership* rules into segments.

``informative_site_finder.find`` (reference :167-344) and ``find_many`` (:601-661) differ in which
sites a DNM sees and how often.  Neither difference involves genotype arithmetic,
so it is resolved here, per DNM, into position windows with a multiplicity; the GPU then does the
binary searches, classifies every (DNM x site) pair and compacts the lists.

The planner also resolves, per DNM: autophasing (:137-164), the DNM's own REF/ALT as seen in the
sites table (snv_phaser.get_refalt :73-84), the read block to fetch from (with the reference's
``chr`` toggling on a failed fetch, read_collector.py:384-392) and the kid's concordant insert
size (read_collector.py:11-25).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib as L
from .schema import ReadTable, SITE_FLAG_SIMPLE, SiteTable

SV_TYPES = ["DEL", "DUP", "INV", "CNV", "DUP:TANDEM", "DEL:ME", "CPX", "CTX"]
SNV_TYPES = ["POINT", "SNV", "INDEL"]

# PAR tables exactly as the reference spells them (utils.py:26-43; the build labels are swapped, Q6)
_PAR = {
    "37": ({"x": (10001, 2781479), "y": (10001, 2781479)},
           {"x": (155701383, 156030895), "y": (56887903, 57217415)}),
    "38": ({"x": (60001, 2699520), "y": (10001, 2649520)},
           {"x": (154931044, 155260560), "y": (59034050, 59363566)}),
}


class FindManyKeyError(KeyError):
    """find_many in CNV mode hits the reference's KeyError (Q12) and threads == 1."""


def strip_chr(s: str) -> str:
    return s.strip("chr")


def is_autophaseable(dn: dict, pedigrees: dict, build: str) -> bool:
    chrom = strip_chr(dn["chrom"].lower())
    if chrom not in ("x", "y"):
        return False
    if int(pedigrees[dn["kid"]]["sex"]) != 1 or build not in _PAR:
        return False
    par1, par2 = _PAR[build]
    s = dn["start"]
    return not (par1[chrom][0] <= s <= par1[chrom][1] or par2[chrom][0] <= s <= par2[chrom][1])


class SiteIndex:
    """Host lookups over a SiteTable that the planner needs (no genotype math)."""

    def __init__(self, sites: SiteTable):
        self.sites = sites
        self.trio_of = {trio: t for t, trio in enumerate(sites.trios)}
        self._contig_rows: Dict[str, tuple] = {}
        self.prefix = self._prefix()

    def _prefix(self) -> str:
        """utils.get_prefix on the first record of the joint VCF (Q26: treated as constant)."""
        s = self.sites
        best = None
        for b in range(s.n_blocks):
            lo, hi = int(s.blk_off[b]), int(s.blk_off[b + 1])
            if hi > lo:
                k = (int(s.blk_contig[b]), int(s.pos[lo]), b)
                if best is None or k < best:
                    best = k
        if best is None:
            return ""
        name = s.contigs[best[0]]
        return name[:3] if "chr" in name.lower() else ""

    def trio(self, pedigrees: dict, kid: str) -> int:
        ped = pedigrees.get(kid)
        if ped is None:
            return -1
        return self.trio_of.get((kid, ped["dad"], ped["mom"]), -1)

    def contig_rows(self, contig: str):
        """(pos, rec_id, row) of every row on a contig over all trio blocks, sorted by pos."""
        c = self._contig_rows.get(contig)
        if c is None:
            s = self.sites
            rows = [np.arange(int(s.blk_off[b]), int(s.blk_off[b + 1]))
                    for b in range(s.n_blocks) if s.contigs[int(s.blk_contig[b])] == contig]
            rows = np.concatenate(rows) if rows else np.zeros(0, dtype=np.int64)
            order = np.argsort(s.pos[rows], kind="stable")
            rows = rows[order]
            ext = sorted((int(s.pos[r]), len(s.extras[r][0]), r) for r in rows.tolist() if r in s.extras)
            c = (s.pos[rows].astype(np.int64), s.record_ids()[rows], rows, ext)
            self._contig_rows[contig] = c
        return c

    def any_simple_row(self, contig: str, lo: int, hi: int) -> bool:
        pos, _rid, rows, _ext = self.contig_rows(contig)
        a, b = np.searchsorted(pos, lo, "left"), np.searchsorted(pos, hi, "right")
        return bool(np.any(self.sites.flag[rows[a:b]] & SITE_FLAG_SIMPLE))

    def refalt(self, chrom: str, start: int):
        """snv_phaser.get_refalt :73-84: records overlapping 0-based [start-1, start+1)."""
        contig = self.prefix + strip_chr(chrom)
        pos, rid, rows, ext = self.contig_rows(contig)
        a, b = np.searchsorted(pos, start - 1, "left"), np.searchsorted(pos, start, "right")
        found = {}
        for i in range(a, b):
            found.setdefault((int(pos[i]), int(rid[i])), int(rows[i]))
        for p, ln, r in ext:                      # long REFs reaching into the interval
            if p >= start - 1:
                break
            if p + ln > start - 1:
                found.setdefault((p, int(self.sites.record_ids()[r])), r)
        ref, alts = None, []
        for key in sorted(found):
            rr, aa = self.sites.ref_alts(found[key])
            if ref is None:
                ref = rr
            alts += list(aa)
        return ref, alts


@dataclass
class Plan:
    """Everything the engine uploads for one batch of DNM entries."""
    dnm: np.ndarray                  # L.DNM_DTYPE[n]
    seg: np.ndarray                  # L.SEG_DTYPE[s]
    alleles: np.ndarray              # u8 blob
    entries: List[dict]              # the DNM dict behind every entry (shared objects)
    trio: np.ndarray                 # i32[n] trio index or -1
    found: np.ndarray                # bool[n]: False -> find() leaves the DNM untouched
    rblk_sblk: Dict[int, int] = field(default_factory=dict)


def _find_windows(dn, sd, whole_region):
    """Segments (lo_pos, hi_pos, mult) of find(): get_position :10-43 as 0-based site positions."""
    s, e = int(dn["start"]), int(dn["end"])
    if whole_region:
        return [(s - sd - 1, e + sd - 1, 1)]
    w1 = (s - sd - 1, s + sd - 1)
    if (e - s) <= sd:
        return [(w1[0], w1[1], 1)]
    w2 = (e - sd - 1, e + sd - 1)
    if w2[0] > w1[1]:
        return [(w1[0], w1[1], 1), (w2[0], w2[1], 1)]
    # overlapping tabix regions return the shared records twice (Q9)
    out = []
    if w2[0] - 1 >= w1[0]:
        out.append((w1[0], w2[0] - 1, 1))
    out.append((w2[0], w1[1], 2))
    out.append((w1[1] + 1, w2[1], 1))
    return out


def plan_find(dnms: List[dict], pedigrees: dict, sidx: SiteIndex, reads: Optional[ReadTable], *,
              search_dist: int, whole_region: bool, build: str, multiread_proc_min: int, threads: int,
              with_reads: bool, first_entry: int = 0, alleles_base: int = 0,
              sv_quirk: bool = False) -> Plan:
    """One entry per DNM, in input order, with the windows ``find``/``find_many`` would search."""
    sites = sidx.sites
    n = len(dnms)
    dnm = np.zeros(n, dtype=L.DNM_DTYPE)
    dnm["rblk"] = -1
    dnm["cnv_entry"] = -1
    trio = np.full(n, -1, dtype=np.int32)
    found = np.zeros(n, dtype=bool)
    segs: List[tuple] = []
    blob = bytearray()
    rblk_sblk: Dict[int, int] = {}
    use_many = n >= multiread_proc_min
    auto = [is_autophaseable(d, pedigrees, build) for d in dnms]

    # ---- find_many lookups (create_lookups :347-396) -------------------------------------
    dead_chroms = set()
    if use_many:
        by_loc: Dict[str, Dict[int, List[str]]] = {}
        starts: Dict[tuple, List[int]] = {}
        ranges: Dict[str, List[int]] = {}
        for i, d in enumerate(dnms):
            if auto[i]:
                continue
            c, s, e, k = d["chrom"], int(d["start"]), int(d["end"]), d["kid"]
            r = ranges.setdefault(c, [s, e])
            r[0], r[1] = min(r[0], s), max(r[1], e)
            by_loc.setdefault(c, {}).setdefault(s, []).append(k)
            starts.setdefault((k, c, s), []).append(i)
            if (e - s) > 2:
                by_loc[c].setdefault(e, []).append(k)
        if whole_region:
            for c, locs in by_loc.items():                      # Q12
                contig = sidx.prefix + strip_chr(c)
                broken = any((k, c, loc) not in starts for loc, kids in locs.items() for k in kids)
                if broken and contig == c and contig in sites.contigs and \
                        sidx.any_simple_row(contig, ranges[c][0] - search_dist - 1, ranges[c][1] + search_dist - 1):
                    if threads == 1:
                        raise FindManyKeyError("find_many: end location without a DNM starting there on %s" % c)
                    dead_chroms.add(c)

    for i, d in enumerate(dnms):
        s, e = int(d["start"]), int(d["end"])
        rec = dnm[i]
        rec["pos"], rec["end"] = s, e
        rec["seg_lo"] = len(segs)
        if auto[i]:
            rec["flags"] = L.DNM_AUTOPHASE | (L.DNM_AUTOPHASE_Y if strip_chr(d["chrom"].lower()) == "y" else 0) \
                | (L.DNM_SV_QUIRK if sv_quirk else 0)
            rec["seg_hi"] = len(segs)
            continue
        t = sidx.trio(pedigrees, d["kid"])
        trio[i] = t
        if t < 0:
            rec["seg_hi"] = len(segs)
            continue
        found[i] = True
        contig = sidx.prefix + strip_chr(d["chrom"])
        sblk = sites.block_of(t, contig)
        vt = d.get("vartype")
        if whole_region and vt is not None:
            mode = L.MODE_CNV_DEL if vt == "DEL" else (L.MODE_CNV_DUP if vt == "DUP" else L.MODE_CNV_NA)
        else:
            mode = L.MODE_READ
        excl = (s, e) if (e - s) < 20 else (0, 0)
        if not use_many:
            wins = _find_windows(d, search_dist, whole_region)
        else:
            c, k = d["chrom"], d["kid"]
            wins = []
            if contig == c and c not in dead_chroms:            # Q11: CHROM must equal the DNM's spelling
                m = by_loc[c][s].count(k)
                if not whole_region:
                    wins = [(s - search_dist - 1, s + search_dist - 1, m)]
                else:
                    ends = sorted(int(dnms[j]["end"]) for j in starts[(k, c, s)])
                    lo = s - search_dist - 1
                    nleft = len(ends)
                    for ee in ends:
                        hi = ee + search_dist - 1
                        if hi >= lo:
                            wins.append((lo, hi, m * nleft))
                            lo = hi + 1
                        nleft -= 1
        for lo, hi, mult in wins:
            if hi >= lo and mult > 0:
                segs.append((sblk, lo, hi, mult, first_entry + i, excl[0], excl[1], mode))
        rec["seg_hi"] = len(segs)

        if with_reads and reads is not None:
            kid_idx = reads.kids.index(d["kid"]) if d["kid"] in reads.kids else -1
            chrom = d["chrom"]
            flags = 0
            if kid_idx >= 0:
                if chrom not in reads.contigs:                    # fetch() raised -> toggled spelling (Q24)
                    chrom = strip_chr(chrom) if "chr" in chrom else "chr" + chrom
                    flags |= L.DNM_FALLBACK_FETCH
                rb = reads.block_of(kid_idx, chrom) if chrom in reads.contigs else -1
                rec["rblk"] = rb
                rec["flags"] |= flags
                if rb >= 0 and sblk >= 0:
                    rblk_sblk[rb] = sblk
            if vt is not None and vt.upper() in SV_TYPES:
                rec["kind"] = L.KIND_SV
            else:
                ref, alts = sidx.refalt(d["chrom"], s)
                if len(alts) == 1 and ref is not None:
                    alt = alts[0]
                    rec["ref_off"], rec["ref_len"] = alleles_base + len(blob), len(ref)
                    blob += ref.encode("ascii")
                    rec["alt_off"], rec["alt_len"] = alleles_base + len(blob), len(alt)
                    blob += alt.encode("ascii")
                    rec["kind"] = L.KIND_SNV if len(ref) == len(alt) else L.KIND_INDEL
    seg = np.array(segs, dtype=L.SEG_DTYPE) if segs else np.zeros(0, dtype=L.SEG_DTYPE)
    return Plan(dnm=dnm, seg=seg, alleles=np.frombuffer(bytes(blob), dtype=np.uint8).copy(),
                entries=list(dnms), trio=trio, found=found, rblk_sblk=rblk_sblk)


def plan_find_fast(dnms: List[dict], pedigrees: dict, sidx: SiteIndex, reads: Optional[ReadTable], *,
                   search_dist: int, whole_region: bool, build: str, multiread_proc_min: int, threads: int,
                   with_reads: bool, first_entry: int = 0, alleles_base: int = 0, sv_quirk: bool = False) -> Plan:
    """Vectorised ``plan_find`` for the per-DNM ``find`` path (len(dnms) < multiread_proc_min) and for
    ``find_many`` in read mode (whole_region=False: one window per DNM around its start, repeated once per
    DNM location of the kid that coincides with it, :392-395,:412-419,:451-452): identical output, numpy
    instead of a Python loop per DNM (the planner would otherwise dominate the end-to-end time of a
    10 k-DNM batch).  ``find_many`` in CNV mode (Q12's KeyError logic) uses the generic planner."""
    n = len(dnms)
    use_many = n >= multiread_proc_min
    if (use_many and whole_region) or n == 0:
        return plan_find(dnms, pedigrees, sidx, reads, search_dist=search_dist, whole_region=whole_region, build=build,
                         multiread_proc_min=multiread_proc_min, threads=threads, with_reads=with_reads,
                         first_entry=first_entry, alleles_base=alleles_base, sv_quirk=sv_quirk)
    sites = sidx.sites
    start = np.fromiter((d["start"] for d in dnms), dtype=np.int64, count=n)
    end = np.fromiter((d["end"] for d in dnms), dtype=np.int64, count=n)
    # ---- per (kid, chrom) group facts -------------------------------------------------------
    gid_of: Dict[tuple, int] = {}
    gids = np.fromiter((gid_of.setdefault((d["kid"], d["chrom"]), len(gid_of)) for d in dnms), dtype=np.int64, count=n)
    groups: List[tuple] = list(gid_of)
    G = len(groups)
    g_trio = np.full(G, -1, dtype=np.int32)
    g_sblk = np.full(G, -1, dtype=np.int32)
    g_rblk = np.full(G, -1, dtype=np.int32)
    g_flag = np.zeros(G, dtype=np.int32)
    g_sex_chrom = np.zeros(G, dtype=np.int8)        # 0 none, 1 x, 2 y (male, known build)
    g_contig: List[str] = []
    rblk_sblk: Dict[int, int] = {}
    for g, (kid, chrom) in enumerate(groups):
        norm = strip_chr(chrom.lower())
        if norm in ("x", "y") and build in _PAR and int(pedigrees[kid]["sex"]) == 1:
            g_sex_chrom[g] = 1 if norm == "x" else 2
        t = sidx.trio(pedigrees, kid)
        g_trio[g] = t
        contig = sidx.prefix + strip_chr(chrom)
        g_contig.append(contig)
        if t >= 0:
            g_sblk[g] = sites.block_of(t, contig)
        if with_reads and reads is not None and kid in reads.kids:
            k = reads.kids.index(kid)
            c2 = chrom
            if c2 not in reads.contigs:
                c2 = strip_chr(c2) if "chr" in c2 else "chr" + c2
                g_flag[g] |= L.DNM_FALLBACK_FETCH
            rb = reads.block_of(k, c2) if c2 in reads.contigs else -1
            g_rblk[g] = rb
    # ---- autophase (:137-164) -----------------------------------------------------------------
    auto = np.zeros(n, dtype=bool)
    sx = g_sex_chrom[gids]
    if sx.any():
        par1, par2 = _PAR[build]
        for code, name in ((1, "x"), (2, "y")):
            m = sx == code
            in_par = ((start >= par1[name][0]) & (start <= par1[name][1])) | ((start >= par2[name][0]) & (start <= par2[name][1]))
            auto |= m & ~in_par
    trio = np.where(auto, -1, g_trio[gids]).astype(np.int32)
    found = (~auto) & (trio >= 0)
    for g in np.unique(gids[found]):
        if g_rblk[g] >= 0 and g_sblk[g] >= 0:
            rblk_sblk[int(g_rblk[g])] = int(g_sblk[g])
    dnm = np.zeros(n, dtype=L.DNM_DTYPE)
    dnm["pos"], dnm["end"] = start, end
    dnm["rblk"] = -1
    dnm["cnv_entry"] = -1
    fl = np.zeros(n, dtype=np.int32)
    fl[auto] = L.DNM_AUTOPHASE | (L.DNM_SV_QUIRK if sv_quirk else 0)
    fl[auto & (sx == 2)] |= L.DNM_AUTOPHASE_Y
    # ---- windows (get_position :10-43) ----------------------------------------------------------
    sd = search_dist
    vt = [d.get("vartype") for d in dnms]
    if whole_region:
        mode = np.fromiter((L.MODE_READ if v is None else (L.MODE_CNV_DEL if v == "DEL" else (L.MODE_CNV_DUP if v == "DUP" else L.MODE_CNV_NA))
                            for v in vt), dtype=np.int32, count=n)
    else:
        mode = np.full(n, L.MODE_READ, dtype=np.int32)
    small = (end - start) < 20
    ex_lo = np.where(small, start, 0)
    ex_hi = np.where(small, end, 0)
    single = whole_region | ((end - start) <= sd)
    has_win = found
    mult1 = np.ones(n, dtype=np.int64)
    if use_many:
        # find_many: only the start window; the VCF's contig name must equal the DNM's spelling (Q11); a site is appended
        # once per occurrence of the kid in the location list of the DNM's start -- DNM starts, and ends of events
        # longer than 2 bp (:392-395) -- of every non-autophased DNM
        single = np.ones(n, dtype=bool)
        g_same = np.array([g_contig[g] == groups[g][1] for g in range(G)], dtype=bool)
        has_win = found & g_same[gids]
        # multiset of (group, location) keys over the non-autophased DNMs, counted with one sort
        live = ~np.asarray(auto, dtype=bool)
        key_s = (gids.astype(np.int64) << 33) + start.astype(np.int64)
        key_e = (gids.astype(np.int64) << 33) + end.astype(np.int64)
        locs = np.concatenate([key_s[live], key_e[live & ((end - start) > 2)]])
        uniq, counts = np.unique(locs, return_counts=True)
        at_ = np.searchsorted(uniq, key_s)
        at_c = np.minimum(at_, max(len(uniq) - 1, 0))
        mult1 = np.where((at_ < len(uniq)) & (uniq[at_c] == key_s) if len(uniq) else np.zeros(n, dtype=bool), counts[at_c] if len(uniq) else 0, 0).astype(np.int64)
        has_win &= mult1 > 0
    nseg = np.where(has_win, np.where(single, 1, 0), 0).astype(np.int64)
    multi_idx = np.nonzero(has_win & ~single)[0]
    multi_wins = {}
    for i in multi_idx:
        w = [x for x in _find_windows(dnms[i], sd, whole_region) if x[1] >= x[0] and x[2] > 0]
        multi_wins[int(i)] = w
        nseg[i] = len(w)
    seg_lo_idx = np.cumsum(nseg) - nseg
    dnm["seg_lo"] = seg_lo_idx
    dnm["seg_hi"] = seg_lo_idx + nseg
    S = int(nseg.sum())
    seg = np.zeros(S, dtype=L.SEG_DTYPE)
    one = np.nonzero(has_win & single)[0]
    at = seg_lo_idx[one]
    seg["sblk"][at] = g_sblk[gids[one]]
    seg["lo_pos"][at] = start[one] - sd - 1
    seg["hi_pos"][at] = (end[one] if whole_region else start[one]) + sd - 1
    seg["mult"][at] = mult1[one]
    seg["dnm"][at] = first_entry + one
    seg["excl_lo"][at], seg["excl_hi"][at], seg["mode"][at] = ex_lo[one], ex_hi[one], mode[one]
    for i, wins in multi_wins.items():
        for k, (lo, hi, mult) in enumerate(wins):
            seg[int(seg_lo_idx[i]) + k] = (int(g_sblk[gids[i]]), lo, hi, mult, first_entry + i, int(ex_lo[i]), int(ex_hi[i]), int(mode[i]))
    # ---- reads: block, kind, REF/ALT of the DNM (get_refalt :73-84) --------------------------------
    blob = bytearray()
    if with_reads and reads is not None:
        dnm["rblk"] = np.where(found, g_rblk[gids], -1)
        fl |= np.where(found, g_flag[gids], 0)
        is_sv = np.fromiter((v is not None and v.upper() in SV_TYPES for v in vt), dtype=bool, count=n)
        kind = np.zeros(n, dtype=np.int32)
        kind[found & is_sv] = L.KIND_SV
        need = np.nonzero(found & ~is_sv)[0]
        ref_off = np.zeros(n, dtype=np.int32); ref_len = np.zeros(n, dtype=np.int32)
        alt_off = np.zeros(n, dtype=np.int32); alt_len = np.zeros(n, dtype=np.int32)
        fast_blob = np.zeros(2 * n, dtype=np.uint8)           # [ref, alt] of entry i at 2i, 2i+1
        slow: List[int] = []
        # group the DNMs that need REF/ALT by contig (a handful of contigs: one vectorised pass each)
        cid_of: Dict[str, int] = {}
        g_cid = np.array([cid_of.setdefault(c, len(cid_of)) for c in g_contig], dtype=np.int64)
        need_cid = g_cid[gids[need]]
        by_contig = {c: need[need_cid == k] for c, k in cid_of.items() if np.any(need_cid == k)}
        rid_all = sites.record_ids()
        for contig, idxs in by_contig.items():
            pos, rid, rows, ext = sidx.contig_rows(contig)
            s0 = start[idxs]
            a = np.searchsorted(pos, s0 - 1, "left")
            b = np.searchsorted(pos, s0, "right")
            cnt = b - a
            # fast case: exactly one row in [start-1, start], it is a simple SNV record and no long REF reaches in
            ok = cnt == 1
            row1 = rows[np.minimum(a, max(len(rows) - 1, 0))] if len(rows) else np.zeros(len(idxs), dtype=np.int64)
            ok &= (sites.flag[row1] & SITE_FLAG_SIMPLE) != 0 if len(rows) else False
            if ext:
                # any extra record with p < start-1 < p+len -> generic path (ext is sorted by p: the furthest end of
                # the records before start-1 is a running maximum)
                cache = sidx.__dict__.setdefault("_ext_cache", {})
                if contig not in cache:
                    ep = np.array([e[0] for e in ext], dtype=np.int64)
                    cache[contig] = (ep, np.maximum.accumulate(ep + np.array([e[1] for e in ext], dtype=np.int64)))
                ep, reach = cache[contig]
                k = np.searchsorted(ep, s0 - 1, "left")
                ok &= ~((k > 0) & (reach[np.maximum(k - 1, 0)] > s0 - 1))
            # two rows that are the same record seen through two trio blocks also go the generic way
            fi = idxs[ok]
            fast_blob[2 * fi] = sites.ref[row1[ok]]
            fast_blob[2 * fi + 1] = sites.alt[row1[ok]]
            ref_off[fi] = alleles_base + 2 * fi; ref_len[fi] = 1
            alt_off[fi] = alleles_base + 2 * fi + 1; alt_len[fi] = 1
            kind[fi] = L.KIND_SNV
            slow += idxs[~ok].tolist()
        blob = bytearray(fast_blob.tobytes())
        for i in slow:
            ref, alts = sidx.refalt(dnms[i]["chrom"], int(start[i]))
            if len(alts) == 1 and ref is not None:
                alt = alts[0]
                ref_off[i], ref_len[i] = alleles_base + len(blob), len(ref)
                blob += ref.encode("ascii")
                alt_off[i], alt_len[i] = alleles_base + len(blob), len(alt)
                blob += alt.encode("ascii")
                kind[i] = L.KIND_SNV if len(ref) == len(alt) else L.KIND_INDEL
        dnm["kind"], dnm["ref_off"], dnm["ref_len"], dnm["alt_off"], dnm["alt_len"] = kind, ref_off, ref_len, alt_off, alt_len
    dnm["flags"] = fl
    return Plan(dnm=dnm, seg=seg, alleles=np.frombuffer(bytes(blob), dtype=np.uint8).copy(), entries=list(dnms),
                trio=trio, found=found, rblk_sblk=rblk_sblk)


def concat_plans(plans: List[Plan]) -> Plan:
    """Entries of several plans back to back (segment / entry indices were made global by the
    caller through first_entry / alleles_base)."""
    off_seg = 0
    dn_parts, seg_parts = [], []
    rs: Dict[int, int] = {}
    for p in plans:
        d = p.dnm.copy()
        d["seg_lo"] += off_seg
        d["seg_hi"] += off_seg
        dn_parts.append(d)
        seg_parts.append(p.seg)
        off_seg += p.seg.shape[0]
        rs.update(p.rblk_sblk)
    return Plan(dnm=np.concatenate(dn_parts), seg=np.concatenate(seg_parts),
                alleles=np.concatenate([p.alleles for p in plans]),
                entries=[e for p in plans for e in p.entries],
                trio=np.concatenate([p.trio for p in plans]),
                found=np.concatenate([p.found for p in plans]), rblk_sblk=rs)


def concordant_upper_lens(reads: ReadTable, readlen: int, insert_size_max_sample: int, stdevs: int) -> np.ndarray:
    """Per read block: estimate_concordant_insert_len of the block's kid (read_collector.py:11-25).
    np.percentile returns a scalar, so the result is int(p99.5) + 0.0 whatever ``stdevs`` is (Q14)."""
    out = np.zeros(reads.n_blocks, dtype=np.float64)
    per_kid: Dict[int, float] = {}
    tl = reads.hdr["tlen"]
    head = getattr(reads, "head_tlen", None) or {}
    for k in range(len(reads.kids)):
        if reads.kids[k] in head:                      # packed from a real file: the head of the BAM
            h = head[reads.kids[k]][: insert_size_max_sample + 1]
            if h.shape[0] == 0:
                per_kid[k] = 0.0
                continue
            pct = np.percentile(np.abs(h.astype(np.int64) - 2 * readlen), 99.5)
            per_kid[k] = float(int(np.mean(pct)) + (np.std(pct) * stdevs))
            continue
        blocks = [b for b in range(reads.n_blocks) if int(reads.blk_kid[b]) == k]
        parts, left = [], insert_size_max_sample + 1
        for b in blocks:
            if left <= 0:
                break
            lo, hi = int(reads.blk_off[b]), int(reads.blk_off[b + 1])
            take = min(left, hi - lo)
            parts.append(tl[lo:lo + take])
            left -= take
        if not parts or sum(p.shape[0] for p in parts) == 0:
            per_kid[k] = 0.0
            continue
        ins = np.abs(np.concatenate(parts).astype(np.int64) - 2 * readlen)
        pct = np.percentile(ins, 99.5)
        per_kid[k] = float(int(np.mean(pct)) + (np.std(pct) * stdevs))
    for b in range(reads.n_blocks):
        out[b] = per_kid[int(reads.blk_kid[b])]
    return out
