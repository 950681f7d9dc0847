"""Seeded synthetic trio data in the columnar layout of ``schema.py``.

The generator is vectorised (numpy) so the BASELINE configs can be produced at full size:
sites on a jittered 1-per-``site_spacing`` grid around every DNM, trio haplotypes Bernoulli(0.5),
kid = first haplotype of each parent, Poisson depths, a few low-GQ / unknown / non-simple records,
allele-balance boundary cases; 150 bp read pairs sampled from the kid's two haplotypes with soft
clips, 2-bp indels, base errors, MAPQ-0 / duplicate / secondary flags and lost mates.

Everything is a pure function of ``SynthConfig`` (including ``seed``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .schema import (
    AUX_SAME_REF, CIG_D, CIG_I, CIG_M, CIG_S, HET, HOM_ALT, HOM_REF, GT_UNKNOWN,
    QUAL_ESCAPE, READ_HDR, SITE_FLAG_SIMPLE, ReadTable, SiteTable, pack_seq,
)

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class SynthConfig:
    n_trios: int = 1
    dnms_per_trio: int = 50
    search_dist: int = 5000
    coverage: float = 30.0
    readlen: int = 150
    site_spacing: int = 300
    frag_mean: float = 450.0
    frag_sd: float = 30.0
    noise: bool = True            # base errors, clips, indels, bad flags, low GQ, unknown GTs
    seed: int = 1
    contigs: Tuple[str, ...] = tuple(str(i) for i in range(1, 23))
    cluster_frac: float = 0.0     # fraction of DNMs placed 1-3 kb after the previous one
    indel_frac: float = 0.0       # fraction of DNMs that are 2-bp deletions / insertions
    sv_frac: float = 0.0          # fraction of DNMs that are DEL/DUP/INV intervals
    sv_max_len: int = 1_000_000
    sex_chrom_frac: float = 0.0   # fraction of DNMs on X / Y (contig names get the same prefix)
    male_frac: float = 0.5
    chr_prefix: str = ""          # prefix of VCF/BAM contig names
    dnm_chr_prefix: Optional[str] = None   # prefix used in the DNM list (default: same)
    read_margin: int = 700        # reads are simulated this far beyond the site window
    # cohort mode: the GLOBAL indices of the trios to generate.  Every trio then draws from its own
    # generator seeded (seed, trio id), so any subset of a cohort -- one rank's shard -- is made of
    # exactly the trios the whole cohort would contain.  None: trios 0..n_trios-1 off one generator.
    trio_ids: Optional[Tuple[int, ...]] = None
    chunk_frags: int = 200_000


@dataclass
class Dataset:
    cfg: SynthConfig
    sites: SiteTable
    reads: ReadTable
    dnms: List[dict]
    pedigrees: Dict[str, dict]
    truth: Dict[str, str]          # DNM key -> "dad" | "mom"
    vcf_name: str = "mem://sites.vcf.gz"

    def bam_name(self, kid: str) -> str:
        return "mem://%s.bam" % kid


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser, vectorised."""
    x = x.astype(np.uint64, copy=True)
    with np.errstate(over="ignore"):
        x += np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def genome_code(contig_id, pos) -> np.ndarray:
    """Reference base code (0..3) at (contig id, 0-based pos); any shape."""
    key = (np.asarray(contig_id).astype(np.uint64) << np.uint64(40)) ^ np.asarray(pos).astype(np.uint64)
    return (_mix64(key) & np.uint64(3)).astype(np.uint8)


# ----------------------------------------------------------------------------------------------
# DNM placement
# ----------------------------------------------------------------------------------------------

def _place_dnms(cfg: SynthConfig, rng: np.random.Generator, trio: int):
    """Return per-DNM arrays (contig index, start, end, kind) for one trio.

    kind: 0 SNV, 1 2-bp deletion, 2 2-bp insertion, 3 DEL, 4 DUP, 5 INV.
    Starts sit on grid offset ``site_spacing-1`` so they never collide with a site row.
    """
    n = cfg.dnms_per_trio
    sp = cfg.site_spacing
    ncont = len(cfg.contigs)
    kind = np.zeros(n, dtype=np.int8)
    u = rng.random(n)
    kind[u < cfg.indel_frac] = 1 + (rng.random(n)[u < cfg.indel_frac] < 0.5)
    svm = (u >= cfg.indel_frac) & (u < cfg.indel_frac + cfg.sv_frac)
    kind[svm] = 3 + rng.integers(0, 3, size=int(svm.sum()))
    # SV lengths log-uniform 50 bp .. sv_max_len
    svlen = np.exp(rng.uniform(np.log(50), np.log(cfg.sv_max_len), size=n)).astype(np.int64)
    svlen[~svm] = 0
    contig = (np.arange(n) + trio) % ncont
    sexm = rng.random(n) < cfg.sex_chrom_frac
    # two extra contigs X, Y appended after the autosomes
    contig = np.where(sexm, ncont + (rng.random(n) < 0.3), contig).astype(np.int32)
    order = np.argsort(contig, kind="stable")
    contig, kind, svlen = contig[order], kind[order], svlen[order]
    clustered = rng.random(n) < cfg.cluster_frac
    start = np.zeros(n, dtype=np.int64)
    base_stride = 2 * cfg.search_dist + 2 * cfg.read_margin + 2 * sp
    # cohort mode: every trio gets its own stretch of each contig.  The site table stands in for a joint VCF, where a
    # position holds ONE record; trios that shared coordinates would see each other's records in get_refalt (Q22)
    trio_shift = 0
    if cfg.trio_ids is not None:
        per_contig = -(-n // max(ncont, 1)) + 2
        trio_shift = trio * ((per_contig * (base_stride + 1000 + cfg.sv_max_len * (cfg.sv_frac > 0)) // sp + 4) * sp)
    cur = 0
    prev_c = -1
    for i in range(n):
        if contig[i] != prev_c:
            # sex chromosomes: start beyond PAR1 of either table (utils.py:26-43)
            cur = 3_000_000 if contig[i] >= ncont else 100 * sp
            cur += trio_shift
            cur += cfg.search_dist + cfg.read_margin + int(rng.integers(0, 64)) * sp
            prev_c = contig[i]
            first = True
        else:
            first = False
        if clustered[i] and not first:
            gap = int(rng.integers(1000, 3000))
        elif first:
            gap = 0
        else:
            gap = base_stride + int(rng.integers(0, 1000))
        cur += gap
        s = (cur // sp) * sp + (sp - 1)
        start[i] = s
        cur = s + int(svlen[i])
    end = start + 1
    end = np.where(kind == 1, start + 3, end)       # REF of 3 bases, 2 deleted
    end = np.where(kind == 2, start + 1, end)
    end = np.where(kind >= 3, start + svlen, end)
    return contig, start.astype(np.int64), end.astype(np.int64), kind


def _merge_intervals(lo: np.ndarray, hi: np.ndarray):
    order = np.argsort(lo, kind="stable")
    lo, hi = lo[order], hi[order]
    out_lo, out_hi = [], []
    for a, b in zip(lo, hi):
        if out_lo and a <= out_hi[-1]:
            out_hi[-1] = max(out_hi[-1], b)
        else:
            out_lo.append(a)
            out_hi.append(b)
    return np.array(out_lo, dtype=np.int64), np.array(out_hi, dtype=np.int64)


def _ranges(lo: np.ndarray, hi: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenated aranges [lo_i, hi_i) and the owning interval index of each element."""
    n = (hi - lo).astype(np.int64)
    tot = int(n.sum())
    if tot == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    owner = np.repeat(np.arange(len(lo)), n)
    offs = np.arange(tot) - np.repeat(np.cumsum(n) - n, n)
    return lo[owner] + offs, owner


# ----------------------------------------------------------------------------------------------
# reads
# ----------------------------------------------------------------------------------------------

def _read_geometry(L: int, start: np.ndarray, ctype: np.ndarray, split: np.ndarray):
    """Reference position of every query base (-1 = soft clip / insertion) and reference end.

    ctype 0: L M | 1: cS (L-c)M | 2: (L-c)M cS | 3: aM 2D (L-a)M | 4: aM 2I (L-a-2)M
    with c = a = ``split``.  ``start`` is reference_start (first aligned base).
    """
    n = start.shape[0]
    k = np.arange(L, dtype=np.int64)[None, :]
    s = start[:, None].astype(np.int64)
    a = split[:, None].astype(np.int64)
    ct = ctype[:, None]
    ref = s + k
    ref = np.where(ct == 1, np.where(k < a, -1, s + k - a), ref)
    ref = np.where(ct == 2, np.where(k < L - a, s + k, -1), ref)
    ref = np.where(ct == 3, np.where(k < a, s + k, s + k + 2), ref)
    ref = np.where(ct == 4, np.where(k < a, s + k, np.where(k < a + 2, -1, s + k - 2)), ref)
    span = np.full(n, L, dtype=np.int64)
    span = np.where((ctype == 1) | (ctype == 2), L - split, span)
    span = np.where(ctype == 3, L + 2, span)
    span = np.where(ctype == 4, L - 2, span)
    return ref, start + span


def _cigar_words(L: int, ctype: np.ndarray, split: np.ndarray):
    """BAM CIGAR words (len<<4|op) per read as (n_cigar, flat words)."""
    n = ctype.shape[0]
    ncig = np.where(ctype == 0, 1, np.where(ctype <= 2, 2, 3)).astype(np.int64)
    w = np.zeros((n, 3), dtype=np.uint32)
    a = split.astype(np.uint32)
    Lu = np.uint32(L)
    m0 = ctype == 0
    w[m0, 0] = (Lu << 4) | CIG_M
    m1 = ctype == 1
    w[m1, 0] = (a[m1] << 4) | CIG_S
    w[m1, 1] = ((Lu - a[m1]) << 4) | CIG_M
    m2 = ctype == 2
    w[m2, 0] = ((Lu - a[m2]) << 4) | CIG_M
    w[m2, 1] = (a[m2] << 4) | CIG_S
    m3 = ctype == 3
    w[m3, 0] = (a[m3] << 4) | CIG_M
    w[m3, 1] = (np.uint32(2) << 4) | CIG_D
    w[m3, 2] = ((Lu - a[m3]) << 4) | CIG_M
    m4 = ctype == 4
    w[m4, 0] = (a[m4] << 4) | CIG_M
    w[m4, 1] = (np.uint32(2) << 4) | CIG_I
    w[m4, 2] = ((Lu - a[m4] - 2) << 4) | CIG_M
    mask = np.arange(3)[None, :] < ncig[:, None]
    return ncig, w[mask]


def make_dataset(cfg: SynthConfig) -> Dataset:
    rng = np.random.default_rng(cfg.seed)
    L = cfg.readlen
    sp = cfg.site_spacing
    sd = cfg.search_dist
    contig_names = [cfg.chr_prefix + c for c in cfg.contigs] + [cfg.chr_prefix + "X", cfg.chr_prefix + "Y"]
    dpref = cfg.chr_prefix if cfg.dnm_chr_prefix is None else cfg.dnm_chr_prefix
    dnm_contig_names = [dpref + c for c in cfg.contigs] + [dpref + "X", dpref + "Y"]

    trios, pedigrees, dnms, truth = [], {}, [], {}
    s_blk_trio, s_blk_contig, s_blk_n = [], [], []
    S = {k: [] for k in ("pos", "flag", "ref", "alt", "gt", "gq", "rd", "ad")}
    extras: Dict[int, Tuple[str, List[str]]] = {}
    r_blk_kid, r_blk_contig, r_blk_n = [], [], []
    R_hdr, R_cig, R_qual, R_code = [], [], [], []
    n_site_rows = 0
    n_reads_tot = 0
    n_cig_tot = 0
    n_q_tot = 0

    for t, gid in enumerate(cfg.trio_ids if cfg.trio_ids is not None else range(cfg.n_trios)):
        if cfg.trio_ids is not None:
            rng = np.random.default_rng([cfg.seed, int(gid)])
        kid, dad, mom = "kid%d" % gid, "dad%d" % gid, "mom%d" % gid
        trios.append((kid, dad, mom))
        sex = "1" if rng.random() < cfg.male_frac else "2"
        pedigrees[kid] = {"kid": kid, "dad": dad, "mom": mom, "sex": sex}
        d_contig, d_start, d_end, d_kind = _place_dnms(cfg, rng, int(gid))
        d_hap = rng.integers(0, 2, size=len(d_start))     # 0: paternal haplotype, 1: maternal

        for c in np.unique(d_contig):
            sel = np.nonzero(d_contig == c)[0]
            cs, ce, ck, ch = d_start[sel], d_end[sel], d_kind[sel], d_hap[sel]
            # ------------------------------------------------------------------ sites
            # window intervals for sites: around start and end (SVs: also the whole interior,
            # capped for very long events to keep fixtures small: interior sampled fully)
            lo = np.concatenate([cs - sd - 2 * sp, ce - sd - 2 * sp])
            hi = np.concatenate([cs + sd + 2 * sp, ce + sd + 2 * sp])
            svm = ck >= 3
            if svm.any():
                lo = np.concatenate([lo, cs[svm]])
                hi = np.concatenate([hi, ce[svm]])
            lo = np.maximum(lo, sp)
            ilo, ihi = _merge_intervals(lo, hi)
            cells, _ = _ranges(ilo // sp, ihi // sp + 1)
            pos = cells * sp + rng.integers(0, sp - 1, size=cells.shape[0])
            nsite = pos.shape[0]
            refc = genome_code(c, pos)
            altc = (refc + rng.integers(1, 4, size=nsite).astype(np.uint8)) & 3
            hap = (rng.random((4, nsite)) < 0.5)              # d0 d1 m0 m1 carry ALT?
            # append the DNM's own records (Q22: REF/ALT of the DNM come from the sites VCF)
            small = ck <= 2
            dn_pos = cs[small]
            dn_ref = genome_code(c, dn_pos)
            dn_alt = (dn_ref + rng.integers(1, 4, size=dn_pos.shape[0]).astype(np.uint8)) & 3
            pos_all = np.concatenate([pos, dn_pos])
            is_dnm = np.concatenate([np.zeros(nsite, bool), np.ones(dn_pos.shape[0], bool)])
            refc = np.concatenate([refc, dn_ref])
            altc = np.concatenate([altc, dn_alt])
            hap = np.concatenate([hap, np.zeros((4, dn_pos.shape[0]), bool)], axis=1)
            order = np.argsort(pos_all, kind="stable")
            pos_all, is_dnm, refc, altc, hap = pos_all[order], is_dnm[order], refc[order], altc[order], hap[:, order]
            V = pos_all.shape[0]
            n_alt = np.stack([hap[0] + 0 + hap[2], hap[0] + 0 + hap[1], hap[2] + 0 + hap[3]])  # kid dad mom
            n_alt[0, is_dnm] = 1
            gt = np.where(n_alt == 0, HOM_REF, np.where(n_alt == 1, HET, HOM_ALT)).astype(np.uint8)
            depth = rng.poisson(30, size=(3, V)).astype(np.int32)
            p_alt = np.where(n_alt == 0, 0.01, np.where(n_alt == 1, 0.5, 0.99))
            ad = rng.binomial(depth, p_alt).astype(np.int32)
            rd = depth - ad
            gq = np.full((3, V), 99.0, dtype=np.float32)
            flag = np.full(V, SITE_FLAG_SIMPLE, dtype=np.uint8)
            refa = _BASES[refc].copy()
            alta = _BASES[altc].copy()
            # kid-side truth for read simulation: ALT on paternal / maternal haplotype
            kid_alt = np.stack([hap[0], hap[2]])
            dn_rows = np.nonzero(is_dnm)[0]
            dn_small_idx = np.nonzero(small)[0]
            # rows sorted by pos; DNM rows map back to DNMs in the same (sorted) order
            dn_sorted = np.argsort(dn_pos, kind="stable")
            for row, di in zip(dn_rows, dn_small_idx[dn_sorted]):
                kid_alt[:, row] = False
                kid_alt[ch[di], row] = True
                if ck[di] == 1:      # 2-bp deletion: REF = 3 bases, ALT = first base
                    r = "".join(chr(_BASES[x]) for x in genome_code(c, np.arange(cs[di], cs[di] + 3)))
                    extras[n_site_rows + row] = (r, [r[0]])
                    flag[row] = 0
                elif ck[di] == 2:    # 2-bp insertion
                    r = chr(refa[row])
                    extras[n_site_rows + row] = (r, [r + "GT"])
                    flag[row] = 0
            if cfg.noise and V:
                m = rng.random((3, V)) < 0.05
                gq[m] = rng.uniform(0, 40, size=int(m.sum())).astype(np.float32)
                m = (rng.random((3, V)) < 0.01) & ~is_dnm[None, :]
                gt[m] = GT_UNKNOWN
                rd[m] = -1
                ad[m] = -1
                m = (rng.random(V) < 0.02) & ~is_dnm
                for row in np.nonzero(m)[0]:
                    r = chr(refa[row])
                    if rng.random() < 0.5:
                        extras[n_site_rows + row] = (r, [chr(alta[row]), "ACGT"[(int(altc[row]) + 1) & 3]])
                    else:
                        extras[n_site_rows + row] = (r + "A", [r])
                    flag[row] = 0
                # allele-balance boundary cases (Q4, Q5): 3/15, 12/15, 33/100, 67/100, 2/3
                bnd = np.array([[12, 3], [3, 12], [67, 33], [33, 67], [5, 10], [10, 20], [0, 30], [30, 0]], dtype=np.int32)
                m = np.nonzero((rng.random((3, V)) < 0.03) & ~is_dnm[None, :])
                pick = rng.integers(0, bnd.shape[0], size=m[0].shape[0])
                rd[m] = bnd[pick, 0]
                ad[m] = bnd[pick, 1]
            # CNV interiors: shift kid depths (DEL hemizygous, DUP 2:1) so get_kid_allele fires
            for di in np.nonzero(svm)[0]:
                inside = (pos_all >= cs[di]) & (pos_all < ce[di]) & ~is_dnm
                rows = np.nonzero(inside)[0]
                if rows.size == 0:
                    continue
                keep = kid_alt[1 - ch[di], rows]       # allele on the haplotype NOT carrying the DEL
                dup = kid_alt[ch[di], rows]            # allele on the duplicated haplotype
                if ck[di] == 3:       # DEL: kid hemizygous -> looks homozygous for the kept allele
                    gt[0, rows] = np.where(keep, HOM_ALT, HOM_REF)
                    d = rng.poisson(15, size=rows.size).astype(np.int32)
                    ad[0, rows] = np.where(keep, d, 0)
                    rd[0, rows] = np.where(keep, 0, d)
                elif ck[di] == 4:     # DUP: 2:1 towards the duplicated haplotype's allele
                    het = gt[0, rows] == HET
                    d = rng.poisson(45, size=rows.size).astype(np.int32)
                    a = rng.binomial(d, np.where(dup, 0.70, 0.30)).astype(np.int32)
                    ad[0, rows] = np.where(het, a, ad[0, rows])
                    rd[0, rows] = np.where(het, d - a, rd[0, rows])
            s_blk_trio.append(t)
            s_blk_contig.append(int(c))
            s_blk_n.append(V)
            S["pos"].append(pos_all.astype(np.int32))
            S["flag"].append(flag)
            S["ref"].append(refa)
            S["alt"].append(alta)
            S["gt"].append(gt)
            S["gq"].append(gq)
            S["rd"].append(rd)
            S["ad"].append(ad)
            n_site_rows += V

            for j in range(len(cs)):
                vt = ["POINT", "POINT", "POINT", "DEL", "DUP", "INV"][int(ck[j])]
                d = {
                    "chrom": dnm_contig_names[int(c)], "start": int(cs[j]), "end": int(ce[j]),
                    "kid": kid, "vartype": vt, "bam": "mem://%s.bam" % kid, "cram_ref": None,
                }
                dnms.append(d)
                key = "{}_{}_{}_{}_{}".format(d["chrom"], d["start"], d["end"], kid, vt)
                truth[key] = "dad" if ch[j] == 0 else "mom"

            # ------------------------------------------------------------------ reads
            # simulate around every breakpoint only (not SV interiors)
            lo = np.concatenate([cs, ce]) - sd - cfg.read_margin
            hi = np.concatenate([cs, ce]) + sd + cfg.read_margin
            lo = np.maximum(lo, 0)
            rlo, rhi = _merge_intervals(lo, hi)
            nfr = np.maximum(((rhi - rlo) * cfg.coverage / (2 * L)).astype(np.int64), 1)
            owner = np.repeat(np.arange(len(rlo)), nfr)
            fstart = rlo[owner] + (rng.random(owner.shape[0]) * (rhi - rlo)[owner]).astype(np.int64)
            frag = np.maximum(np.rint(rng.normal(cfg.frag_mean, cfg.frag_sd, size=owner.shape[0])).astype(np.int64), 2 * L + 20)
            fhap = rng.integers(0, 2, size=owner.shape[0])
            nf = owner.shape[0]
            # per-read arrays, r1 then r2
            start = np.concatenate([fstart, fstart + frag - L])
            rhap = np.concatenate([fhap, fhap])
            n = 2 * nf
            ctype = np.zeros(n, dtype=np.int8)
            split = np.zeros(n, dtype=np.int64)
            if cfg.noise:
                u = rng.random(n)
                ctype[u < 0.03] = 1
                ctype[(u >= 0.03) & (u < 0.06)] = 2
                ctype[(u >= 0.06) & (u < 0.08)] = 3
                ctype[(u >= 0.08) & (u < 0.10)] = 4
                split = np.where(ctype <= 2, 10, rng.integers(20, L - 20, size=n))
                split = np.where(ctype == 0, 0, split)
            # small indel DNMs: reads on the carrying haplotype show the event in their CIGAR
            for di in np.nonzero((ck == 1) | (ck == 2))[0]:
                s0 = cs[di]
                cov = (rhap == ch[di]) & (start <= s0 - 1) & (start + L >= s0 + 6)
                ctype[cov] = 3 if ck[di] == 1 else 4
                split[cov] = s0 + 1 - start[cov]
                # reads starting inside the affected bases: push them past the event
                ins = (rhap == ch[di]) & (start > s0 - 1) & (start <= s0 + 3)
                start[ins] = s0 + 4
            ncig, cigw = _cigar_words(L, ctype, split)
            # haplotype sequences over a virtual concatenation of the simulated intervals
            vlo = np.maximum(rlo - 16, 0)
            vhi = rhi + int(cfg.frag_mean + 12 * cfg.frag_sd) + 2 * L + 64
            vlen = vhi - vlo
            voff = np.cumsum(vlen) - vlen
            vpos, _ = _ranges(vlo, vhi)
            hapseq = np.empty((2, vpos.shape[0]), dtype=np.uint8)
            hapseq[0] = genome_code(c, vpos)
            hapseq[1] = hapseq[0]
            if V:
                iv = np.searchsorted(vlo, pos_all, side="right") - 1
                ok = (iv >= 0) & (pos_all < vhi[np.maximum(iv, 0)]) & (flag == SITE_FLAG_SIMPLE)
                vidx = voff[np.maximum(iv, 0)] + pos_all - vlo[np.maximum(iv, 0)]
                for h in (0, 1):
                    mm = ok & kid_alt[h]
                    hapseq[h, vidx[mm]] = altc[mm]
            rown = np.concatenate([owner, owner])
            vstart = voff[rown] + start - vlo[rown]
            code = np.empty((n, L), dtype=np.uint8)
            simple_rd = ctype == 0
            for h in (0, 1):
                win = np.lib.stride_tricks.sliding_window_view(hapseq[h], L)
                sel = np.nonzero(simple_rd & (rhap == h))[0]
                for a0 in range(0, sel.shape[0], cfg.chunk_frags):
                    ss = sel[a0:a0 + cfg.chunk_frags]
                    code[ss] = win[vstart[ss]]
            sel = np.nonzero(~simple_rd)[0]
            if sel.size:
                rp, _ = _read_geometry(L, start[sel], ctype[sel], split[sel])
                vi = voff[rown[sel], None] + np.where(rp < 0, start[sel, None], rp) - vlo[rown[sel], None]
                cc = hapseq[rhap[sel, None], vi]
                code[sel] = np.where(rp < 0, rng.integers(0, 4, size=rp.shape).astype(np.uint8), cc)
            qual = np.full((n, L), 37, dtype=np.uint8)
            flagr = np.concatenate([np.full(nf, 0x1 | 0x2 | 0x40 | 0x20), np.full(nf, 0x1 | 0x2 | 0x80 | 0x10)]).astype(np.uint16)
            mapq = np.full(n, 60, dtype=np.uint8)
            tlen = np.concatenate([frag, -frag]).astype(np.int32)
            mate = np.concatenate([np.arange(nf, n), np.arange(0, nf)]).astype(np.int64)
            keep = np.ones(n, dtype=bool)
            if cfg.noise:
                cf, qf = code.reshape(-1), qual.reshape(-1)
                ne = rng.binomial(n * L, 0.01)
                e = rng.integers(0, n * L, size=ne)
                cf[e] = (cf[e] + rng.integers(1, 4, size=ne).astype(np.uint8)) & 3
                lowe = e[rng.random(ne) < 0.5]
                qf[lowe] = 12
                nn = rng.binomial(n * L, 0.001)
                e = rng.integers(0, n * L, size=nn)
                cf[e] = 0
                qf[e] = 2 | QUAL_ESCAPE
                lowq = rng.random(n) < 0.02          # reads with many low-quality bases
                qual[lowq, : L // 3] = 8
                mapq[rng.random(n) < 0.02] = 0
                u = rng.random(n)
                flagr[u < 0.005] |= 0x400
                flagr[(u >= 0.005) & (u < 0.01)] |= 0x100
                lost = rng.random(nf) < 0.005         # r2 missing from the file
                keep[nf:][lost] = False
            # sort by start (file order), drop lost mates, remap mate pointers
            order = np.argsort(start, kind="stable")
            order = order[keep[order]]
            newidx = np.full(n, -1, dtype=np.int64)
            newidx[order] = np.arange(order.shape[0])
            mate_new = newidx[mate[order]]
            m = order.shape[0]
            hdr = np.zeros(m, dtype=READ_HDR)
            hdr["start"] = start[order]
            hdr["tlen"] = tlen[order]
            hdr["mate"] = np.where(mate_new >= 0, mate_new + n_reads_tot, -1)
            nc = ncig[order]
            coff = np.cumsum(nc) - nc
            hdr["cigar_off"] = coff + n_cig_tot
            q0 = np.arange(m, dtype=np.int64) * L + n_q_tot
            hdr["qoff_lo"] = (q0 & 0xFFFFFFFF).astype(np.uint32)
            hdr["qoff_hi"] = (q0 >> 32).astype(np.uint8)
            hdr["l_seq"] = L
            hdr["flag"] = flagr[order]
            hdr["n_cigar"] = nc
            hdr["mapq"] = mapq[order]
            hdr["aux"] = AUX_SAME_REF
            # CIGAR words in the new order
            cstart = np.cumsum(ncig) - ncig
            widx, _ = _ranges(cstart[order], cstart[order] + nc)
            R_hdr.append(hdr)
            R_cig.append(cigw[widx])
            R_qual.append(qual[order].reshape(-1))
            R_code.append(code[order].reshape(-1))
            r_blk_kid.append(t)
            r_blk_contig.append(int(c))
            r_blk_n.append(m)
            n_reads_tot += m
            n_cig_tot += int(nc.sum())
            n_q_tot += m * L

    def cat(lst, dtype, axis=0):
        return np.concatenate(lst, axis=axis).astype(dtype, copy=False) if lst else np.zeros(0, dtype=dtype)

    sites = SiteTable(
        trios=trios, contigs=contig_names,
        blk_trio=np.array(s_blk_trio, dtype=np.int32), blk_contig=np.array(s_blk_contig, dtype=np.int32),
        blk_off=np.concatenate([[0], np.cumsum(s_blk_n)]).astype(np.int64),
        pos=cat(S["pos"], np.int32), flag=cat(S["flag"], np.uint8), ref=cat(S["ref"], np.uint8),
        alt=cat(S["alt"], np.uint8), gt=np.ascontiguousarray(cat(S["gt"], np.uint8, axis=1)),
        gq=np.ascontiguousarray(cat(S["gq"], np.float32, axis=1)),
        rd=np.ascontiguousarray(cat(S["rd"], np.int32, axis=1)),
        ad=np.ascontiguousarray(cat(S["ad"], np.int32, axis=1)), extras=extras,
    )
    codes = cat(R_code, np.uint8)
    reads = ReadTable(
        kids=[t[0] for t in trios], contigs=contig_names,
        blk_kid=np.array(r_blk_kid, dtype=np.int32), blk_contig=np.array(r_blk_contig, dtype=np.int32),
        blk_off=np.concatenate([[0], np.cumsum(r_blk_n)]).astype(np.int64),
        hdr=np.concatenate(R_hdr) if R_hdr else np.zeros(0, dtype=READ_HDR),
        cigar=cat(R_cig, np.uint32), qual=cat(R_qual, np.uint8), seq2=pack_seq(codes),
    )
    sites.validate()
    reads.validate()
    return Dataset(cfg=cfg, sites=sites, reads=reads, dnms=dnms, pedigrees=pedigrees, truth=truth)
