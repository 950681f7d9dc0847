// Read path, part 2: seed reads, extended read-backed chaining, site matching and the evidence tally.
// One CTA per DNM.
//
// The reference chains reads with a recursive, order-dependent breadth-first 2-colouring
// (read_collector.py connect_reads :76-152).  It is reproduced exactly as a level-synchronous BFS:
// inside one level a still-unlabelled read pair is claimed by the FIRST (finder, site) visit of the
// sequential order, and that order is a total order on keys (haplotype order, position of the
// finder in its list, index of the site in the finder's site list).  Because the first finder that
// visits a site labels every eligible pair registered there, only the minimum key per site matters:
//   best[site]  = min key over this level's finders that can read an allele at the site
//   claim[pair] = min best[site] over the sites where the pair is an eligible target
// The order of the next level's lists is (claiming key, position in the site's read list), which is
// recovered with per-site ballot ranks and a rank of the sites by key -- no global sort.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace {

constexpr unsigned long long KEY_NONE = ~0ull;
#ifndef UNFZ_CHAIN_WIDE_MIN_INC
#define UNFZ_CHAIN_WIDE_MIN_INC 2048     // incidences per DNM from which the 512-thread CTA shape is launched
#endif
constexpr int CH_SMEM_SITES = 128;         // per-site state lives in shared memory up to this many het sites

// what a window slot carries: the key that elects a pair's last writer and the pair's dense id
struct __align__(16) SlotRec {
    unsigned long long key;
    int32_t dense;
    int32_t pad;
};

struct Scratch {
    SlotRec* slot;                                                  // per window slot
    // per het-site incidence (site-major); inc_x holds the window slot during set-up, then the dense pair
    int32_t* inc_r; int32_t* inc_x; int32_t* inc_site; uint8_t* inc_al;
    // seed entries
    int32_t* seed_e; uint8_t* seed_hap; uint32_t* seed_reg;
    // seed incidences
    int32_t* sinc_px; int32_t* sinc_site; uint32_t* sinc_sidx; uint8_t* sinc_al;
    // per dense pair (at most one per incidence + one per seed entry): slot, primary read, chaining state
    int32_t* x_of; int32_t* prim;
    uint32_t* p_ord; int32_t* p_fpos; int32_t* p_tmp; unsigned long long* p_minkey; uint8_t* p_label;
    int32_t* adj; int32_t* adj_off; int32_t* front0; int32_t* front1;
    // per het site (site_off has one extra entry per DNM)
    int32_t* spos; uint8_t* sref; uint8_t* salt; int32_t* site_off; int32_t* cand_off;
    unsigned long long* bestkey; int32_t* site_cnt; int32_t* site_base; int32_t* active;
    // per candidate site
    int32_t* cpos;
    // per DNM: hand-over between the three chaining kernels
    int32_t* meta;
};

struct ChainArgs {
    const UnfzDnm* dnms; int32_t n_dnms;
    const UnfzSegIn* segs; const int64_t* seg_pair_off;
    UnfzSiteCols sites; UnfzReadCols reads;
    const UnfzReadSum* rsum; const int32_t* blk_maxspan; const uint32_t* hits; const int32_t* mp;
    const int32_t* het_list; const int32_t* n_het; const uint32_t* cand_list; const int32_t* n_cand;
    const uint8_t* alleles; const int32_t* win; const int64_t* off;  // win[4][n], off[6][n+1]
    int32_t split_margin;
    const uint32_t* hit_tile_base; int32_t hit_tile_reads;
    const int32_t* site_lo; const int32_t* site_n; const int32_t* seed_win;   // fetch ranges found by chain_size
    int32_t readlen, ext_goal, no_extended;
    uint8_t* slot_label; uint8_t* slot_evid; uint8_t* cand_evid; UnfzTally* tally;
    int64_t* ev_need;             // [4][n_dnms] or null: dad pairs, mom pairs, dad site entries, mom site entries
    const int32_t* guard;
    Scratch S;
};

#ifdef CH_PREFETCH
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

// start and mate of a read come out of its 32-byte summary (same sector as end / flags / hit slots), not the header
__device__ __forceinline__ int32_t rs_start(const UnfzReadSum* __restrict__ S, int64_t r) { return __ldg(&S[r].start); }
__device__ __forceinline__ int32_t rs_mate(const UnfzReadSum* __restrict__ S, int64_t r) { return __ldg(&S[r].mate); }

// first read index in [lo,hi) with start >= v
__device__ __forceinline__ int64_t lb_start(const int32_t* __restrict__ R, int64_t lo, int64_t hi, int64_t v) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)__ldg(R + mid) < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int cigar_qpos2(const uint32_t* __restrict__ cg, int n_cigar, int32_t start, int32_t p) {
    int32_t cur = start;
    int q = 0;
    for (int k = 0; k < n_cigar; ++k) {
        const uint32_t w = __ldg(cg + k);
        const uint32_t op = w & 15u;
        const int32_t ln = (int32_t)(w >> 4);
        if (op == 0 || op == 7 || op == 8) {
            if (p < cur + ln) return p >= cur ? q + (p - cur) : -1;
            cur += ln; q += ln;
        } else if (op == 1 || op == 4) {
            q += ln;
        } else if (op == 2 || op == 3) {
            if (p < cur + ln) return -1;
            cur += ln;
        }
    }
    return -1;
}

__device__ __forceinline__ uint32_t plane_bit(const uint32_t* __restrict__ plane, int64_t g) {
    return (__ldg(plane + (g >> 5)) >> (g & 31)) & 1u;
}
__device__ __forceinline__ char base_char(const UnfzReadCols& R, int64_t g) {
    const uint32_t code = (__ldg(R.seq2 + (g >> 2)) >> ((g & 3) << 1)) & 3u;
    if (plane_bit(R.nmask, g)) return code == 0 ? 'N' : '?';
    return "ACGT"[code];
}
__device__ __forceinline__ char hit_char(uint32_t w) {
    const uint32_t code = (w >> 24) & 3u;
    if (w & (0x80u << 16)) return code == 0 ? 'N' : '?';
    return "ACGT"[code];
}

// hit word of (read e, site row) or 0
// The lookups below are chains of dependent gathers; everything that does not depend on an earlier
// result is loaded up front (no early-out between the loads), so a lookup costs two round trips, not four.
__device__ __forceinline__ uint32_t hit_lookup(const ChainArgs& A, int64_t e, int64_t row) {
    const UnfzReadSum s = load_rsum(A.rsum + e);
    const int mp = __ldg(A.mp + row);
    const uint32_t tb = __ldg(A.hit_tile_base + (uint32_t)e / (uint32_t)A.hit_tile_reads);
    const int k = mp - s.fmark;
    if (k < 0 || k >= (int)s.cnt) return 0;
    return __ldg(A.hits + (int64_t)tb + s.hoff + k);
}

// goodread + insert + mate + None-count + mate-overlap (read_collector.py:181-214, :395-418)
__device__ bool pair_ok(const ChainArgs& A, int64_t r, bool ext) {
    const UnfzReadSum s = load_rsum(A.rsum + r);                                // one sector: span, mate, flags
    uint32_t need = UNFZ_RS_GOOD_CONC | UNFZ_RS_INS_OK | UNFZ_RS_HAS_MATE | UNFZ_RS_NONE_OK;
    if (ext) need |= UNFZ_RS_EXT_OK;
    const bool ok_r = (s.flags & need) == need;                                  // HAS_MATE: mate >= 0
    const int64_t m = ok_r ? (int64_t)s.mate : r;
    const UnfzReadSum sm = load_rsum(A.rsum + m);                               // one sector for the mate
    const uint32_t needm = UNFZ_RS_GOOD_CONC | UNFZ_RS_NONE_OK;
    const int32_t r0 = s.start, r1 = s.end, m0 = sm.start, m1 = sm.end;
    return ok_r && (sm.flags & needm) == needm && !((m0 <= r0 && r0 <= m1) || (m0 <= r1 && r1 <= m1));
}

// query index of reference position p in read e, the offset of the read's first base and its length.  A single-M read
// (aux bit6, set at upload: nine reads in ten) answers out of its header; the CIGAR gather is skipped
struct QPos { int q; int l_seq; int64_t g0; int32_t mate; };
__device__ __forceinline__ QPos qpos_of(const ChainArgs& A, int64_t e, int32_t p) {
    const UnfzRead h = load_read(A.reads.hdr + e);
    QPos o;
    if (h.aux & 0x40u) o.q = (p >= h.start && p < h.start + h.l_seq) ? (int)(p - h.start) : -1;
    else o.q = cigar_qpos2(A.reads.cigar + h.cigar_off, h.n_cigar, h.start, p);
    o.l_seq = h.l_seq;
    o.g0 = read_qoff(h);
    o.mate = h.mate;
    return o;
}

// snv_match_alleles (:296-336) through get_allele_at (:56-73): 0 none, 1 ref, 2 alt
__device__ int seed_snv(const ChainArgs& A, const UnfzDnm& dn, int64_t r) {
    QPos h = qpos_of(A, r, dn.pos);
    int q = h.q;
    if (q < 0) {
        h = qpos_of(A, h.mate, dn.pos);
        q = h.q;
        if (q < 0) return 0;
    }
    if (q < 4 || q > A.readlen - 4) return 0;
    const int n = dn.ref_len > dn.alt_len ? dn.ref_len : dn.alt_len;
    if (!(h.l_seq > q + n)) return 0;
    const int64_t g = h.g0 + q;
    bool is_ref = true;
    for (int i = 0; i < dn.ref_len; ++i)
        if (base_char(A.reads, g + i) != (char)A.alleles[dn.ref_off + i]) { is_ref = false; break; }
    if (is_ref) return 1;
    bool is_alt = true;
    for (int i = 0; i < dn.alt_len; ++i)
        if (base_char(A.reads, g + i) != (char)A.alleles[dn.alt_off + i]) { is_alt = false; break; }
    return is_alt ? 2 : 0;
}

// indel_match_alleles (:266-293), Q21: the CIGAR is expanded over ALL operations
__device__ int seed_indel(const ChainArgs& A, const UnfzDnm& dn, int64_t r) {
    const UnfzRead h = load_read(A.reads.hdr + r);
    const uint32_t* cg = A.reads.cigar + h.cigar_off;
    const bool simple = (h.aux & 0x40u) != 0;            // one M operation: no I/D, only the "ref" outcome is possible
    const int q = simple ? ((dn.pos >= h.start && dn.pos < h.start + h.l_seq) ? (int)(dn.pos - h.start) : -1)
                         : cigar_qpos2(cg, h.n_cigar, h.start, dn.pos);
    if (q < 0) return 0;
    const int n = dn.ref_len > dn.alt_len ? dn.ref_len : dn.alt_len;
    const int64_t g0 = read_qoff(h);
    for (int i = q; i < q + n && i < h.l_seq; ++i)
        if (plane_bit(A.reads.lowq, g0 + i)) return 0;
    if (simple) return (7 < q && q < h.l_seq - 7) ? 1 : 0;
    bool has_id = false;
    int64_t eo = 0, npos = 0;
    for (int k = 0; k < h.n_cigar; ++k) {
        const uint32_t w = __ldg(cg + k);
        const uint32_t op = w & 15u;
        const int64_t ln = w >> 4;
        if ((op == 1 || op == 2) && eo < (int64_t)q + n && eo + ln > q) has_id = true;
        eo += ln;
        if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) npos += ln;
    }
    if (has_id) return 2;
    if (7 < q && q < npos - 7) return 1;
    return 0;
}


// collect_reads_sv (:515-522): fewer than 7 of the first 10 AND of the last 10 expanded CIGAR
// operations are M/= -> the read name is banned
__device__ bool sv_bad_ends(const uint32_t* __restrict__ cg, int n) {
    int s_m = 0, left = 10;
    for (int k = 0; k < n && left > 0; ++k) {
        const uint32_t w = __ldg(cg + k);
        const int take = min(left, (int)(w >> 4));
        if ((w & 15u) == 0 || (w & 15u) == 7) s_m += take;
        left -= take;
    }
    int e_m = 0;
    left = 10;
    for (int k = n - 1; k >= 0 && left > 0; --k) {
        const uint32_t w = __ldg(cg + k);
        const int take = min(left, (int)(w >> 4));
        if ((w & 15u) == 0 || (w & 15u) == 7) e_m += take;
        left -= take;
    }
    return e_m < 7 && s_m < 7;
}

// goodread(read, True), mate found, goodread(mate, True)  (:503-513)
__device__ bool sv_goodok(const ChainArgs& A, int64_t r) {
    const UnfzReadSum s = load_rsum(A.rsum + r);
    const uint32_t need = UNFZ_RS_GOOD_DISC | UNFZ_RS_HAS_MATE;
    if ((s.flags & need) != need) return false;
    const int64_t m = rs_mate(A.rsum, r);
    return (load_rsum(A.rsum + m).flags & UNFZ_RS_GOOD_DISC) != 0;
}

__device__ bool sv_banned(const ChainArgs& A, int64_t r) {
    if (!sv_goodok(A, r)) return false;
    const UnfzRead h = load_read(A.reads.hdr + r);
    return sv_bad_ends(A.reads.cigar + h.cigar_off, h.n_cigar);
}

// split / discordant / clipped support of an SV breakpoint (:524-586): 0 none, 1 [read, mate], 2 [mate, read]
__device__ int sv_support(const ChainArgs& A, const UnfzDnm& dn, int64_t r, int64_t position, double cul) {
    const UnfzRead h = load_read(A.reads.hdr + r);
    const UnfzReadSum s = load_rsum(A.rsum + r);
    const int64_t r0 = h.start, r1 = s.end;
    if (h.aux & 2u) {
        const int64_t em = A.split_margin;
        if ((position - em <= r0 && r0 <= position + em) || (position - em <= r1 && r1 <= position + em)) return 1;
        return 0;
    }
    long long ins = (long long)h.tlen - 2ll * A.readlen;
    if (ins < 0) ins = -ins;
    const double var_len = fabs((double)dn.end - (double)dn.pos);
    if ((double)ins > cul) {
        const double ratio = fabs(var_len / (double)ins);
        if (0.7 < ratio && ratio < 1.3) {
            const int64_t m0 = rs_start(A.rsum, h.mate);
            const int64_t left0 = min(m0, r0), right0 = max(m0, r0);
            const int64_t w = (int64_t)cul;
            if ((dn.pos - w) < left0 && left0 < (dn.pos + w) && (dn.end - w) < right0 && right0 < (dn.end + w)) return 2;
            return 0;
        }
    }
    // clipped read that is not a split alignment
    const uint32_t* cg = A.reads.cigar + h.cigar_off;
    int k = cigar_qpos2(cg, h.n_cigar, h.start, (int32_t)position);
    if (k < 0) k = cigar_qpos2(cg, h.n_cigar, h.start, (int32_t)position - 1);
    if (k < 0) k = cigar_qpos2(cg, h.n_cigar, h.start, (int32_t)position + 1);
    if (k < 0) return 0;
    int64_t npos = 0, lead = 0, trail = 0;
    bool seen = false;
    for (int i = 0; i < h.n_cigar; ++i) {
        const uint32_t w = __ldg(cg + i);
        const uint32_t op = w & 15u;
        const int64_t ln = w >> 4;
        if (op == 0 || op == 7 || op == 8) { npos += ln; seen = true; trail = 0; }
        else if (op == 1 || op == 4) { npos += ln; if (!seen) lead += ln; else trail += ln; }
    }
    if (k < 2 || k > npos - 4) return 0;
    const int64_t first_al = lead, last_al = npos - 1 - trail;
    const bool before_none = first_al >= k - 1;                 // positions[:k-1] are all None (k-1 >= 1)
    const bool after_none = (k + 1 < npos) && last_al <= k;     // positions[k+1:] non-empty and all None
    return (before_none || after_none) ? 2 : 0;
}


// site_searcher.binary_search pivot (:6-47); -1 when no site has start <= pos < end
__device__ int bisect_pivot(const int32_t* __restrict__ spos, int n, int32_t start, int32_t end) {
    int lo = 0, hi = n - 1, plo = -1, phi = -1;
    while (hi > -1) {
        if (lo > hi) break;
        if (lo == plo && hi == phi) break;
        plo = lo; phi = hi;
        const int mid = (hi + lo) / 2;
        const int32_t x = spos[mid];
        if (start <= x && x < end) return mid;
        else if (x > start) hi = mid - 1;
        else if (x < start) lo = mid + 1;
    }
    return -1;
}

// allele code of pair x at het site i relative to the site: bits0-1 finder (0 none,1 ref,2 alt),
// bits2-3 target (needs the position in the PRIMARY read and base quality >= min)
__device__ uint8_t allele_info(const ChainArgs& A, const Scratch& S, int64_t e0, int64_t row, char ref, char alt) {
    if (e0 < 0) return 0;
    const int64_t e1 = rs_mate(A.rsum, e0);          // in flight with the first lookup
    const uint32_t h0 = hit_lookup(A, e0, row);
    uint32_t h = h0;
    if (!(h0 & 0xffffu)) {
        if (e1 < 0) return 0;
        h = hit_lookup(A, e1, row);
        if (!(h & 0xffffu)) return 0;
    }
    const int q = (int)(h & 0xffffu) - 1;
    if (q < 4 || q > A.readlen - 4) return 0;
    if (!(h & (1u << 26))) return 0;                 // len(seq) > q + 1
    const char c = hit_char(h);
    const uint8_t code = c == ref ? 1 : (c == alt ? 2 : 0);
    if (!code) return 0;
    uint8_t out = code;
    if ((h0 & 0xffffu) && !(h0 & (1u << 16))) out |= code << 2;       // base quality >= min (hit word bit16: below)
    return out;
}

#ifdef CH_DEBUG
__device__ unsigned long long g_ch_dbg[16];
#define CH_MARK(i) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&g_ch_dbg[i], (unsigned long long)(t_ - ch_t0)); ch_t0 = t_; } } while (0)
#else
#define CH_MARK(i)
#endif

// ------------------------------------------------------------------------------------------------
// the three chaining kernels, instantiated for two CTA shapes (see chain_cta.inc)
// ------------------------------------------------------------------------------------------------
#ifndef CH_THREADS_N            // profiling builds may override the small shape
#define CH_THREADS_N 128
#endif
#ifndef CH_MINB
#define CH_MINB 16
#endif
#ifndef CH_BFS_MINB
#define CH_BFS_MINB 8
#endif
#define CH_NS cta_small
#include "chain_cta.inc"
#undef CH_NS
#undef CH_THREADS_N
#undef CH_MINB
#undef CH_BFS_MINB
#define CH_NS cta_wide
#define CH_THREADS_N 512
#define CH_MINB 4
#define CH_BFS_MINB 2
#include "chain_cta.inc"
#undef CH_NS
#undef CH_THREADS_N
#undef CH_MINB
#undef CH_BFS_MINB


// ------------------------------------------------------------------------------------------------
// evidence lists: the pairs and informative sites behind the tallies, per DNM and per parent, in
// slot / list order -- what snv_phaser.py:169-203 turns into the record's dad_reads / mom_reads /
// dad_sites / mom_sites.  One warp per DNM; ev_off[4][n+1] = dad pairs, mom pairs, dad sites, mom sites.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
evidence_lists_kernel(const UnfzDnm* __restrict__ dnms, int32_t n_dnms, const int64_t* __restrict__ seg_pair_off,
                      const int32_t* __restrict__ site_pos, const uint32_t* __restrict__ cand_list,
                      const int32_t* __restrict__ n_cand, const uint8_t* __restrict__ cand_evid,
                      const int32_t* __restrict__ win, const int64_t* __restrict__ slot_off,
                      const uint8_t* __restrict__ slot_evid, const int64_t* __restrict__ ev_off,
                      int32_t* __restrict__ ev_read_dad, int32_t* __restrict__ ev_read_mom,
                      int32_t* __restrict__ ev_pos_dad, int32_t* __restrict__ ev_pos_mom,
                      const int32_t* __restrict__ guard) {
    UNFZ_GUARD(guard);
    const int lane = threadIdx.x & 31;
    const int d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (d >= n_dnms) return;
    const unsigned below = (1u << lane) - 1u;
    const int64_t n1 = (int64_t)n_dnms + 1;
    int64_t kd = ev_off[d], km = ev_off[n1 + d];
    if (ev_off[d + 1] > kd || ev_off[n1 + d + 1] > km) {
        const int64_t nd = n_dnms;
        const int64_t a_lo = win[d], a_hi = win[nd + d], b_lo = win[2 * nd + d], b_hi = win[3 * nd + d];
        const int na = (int)(a_hi - a_lo), W = na + (int)(b_hi - b_lo);
        const uint8_t* ev = slot_evid + slot_off[d];
        for (int x0 = 0; x0 < W; x0 += 32) {
            const int x = x0 + lane;
            const uint8_t e = x < W ? ev[x] : 0;
            const int32_t r = (int32_t)(x < na ? a_lo + x : b_lo + (x - na));
            const unsigned bd = __ballot_sync(0xffffffffu, e & 1), bm = __ballot_sync(0xffffffffu, e & 2);
            if (e & 1) ev_read_dad[kd + __popc(bd & below)] = r;
            if (e & 2) ev_read_mom[km + __popc(bm & below)] = r;
            kd += __popc(bd);
            km += __popc(bm);
        }
    }
    kd = ev_off[2 * n1 + d];
    km = ev_off[3 * n1 + d];
    if (ev_off[2 * n1 + d + 1] > kd || ev_off[3 * n1 + d + 1] > km) {
        const UnfzDnm dn = dnms[d];
        const int64_t lbase = seg_pair_off[dn.seg_lo];
        const int nc = n_cand[d];
        for (int j0 = 0; j0 < nc; j0 += 32) {
            const int j = j0 + lane;
            const uint8_t e = j < nc ? cand_evid[lbase + j] : 0;
            const int32_t p = e ? __ldg(site_pos + (int64_t)(cand_list[lbase + j] & 0x3fffffffu)) : 0;
            const unsigned bd = __ballot_sync(0xffffffffu, e & 1), bm = __ballot_sync(0xffffffffu, e & 2);
            if (e & 1) ev_pos_dad[kd + __popc(bd & below)] = p;
            if (e & 2) ev_pos_mom[km + __popc(bm & below)] = p;
            kd += __popc(bd);
            km += __popc(bm);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// sizing: read window (union of two index ranges) and scratch needs per DNM, one warp per DNM
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t warp_sum64(int64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int64_t warp_min64(int64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int64_t warp_max64(int64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// lb_start by a whole warp: 32 probes per round trip (all lanes get the result)
__device__ __forceinline__ int64_t warp_lb_start(const int32_t* __restrict__ R, int64_t lo, int64_t hi, int64_t v, int lane) {
    while (hi - lo > 32) {
        const int64_t step = (hi - lo + 32) / 33;
        const int64_t i = lo + (int64_t)(lane + 1) * step - 1;
        const bool lt = i < hi && (int64_t)__ldg(R + i) < v;
        const int k = __popc(__ballot_sync(0xffffffffu, lt));
        const int64_t nhi = k < 32 ? min(hi, lo + (int64_t)(k + 1) * step - 1) : hi;
        lo += (int64_t)k * step;
        hi = nhi;
    }
    const bool lt = lo + lane < hi && (int64_t)__ldg(R + lo + lane) < v;
    return lo + __popc(__ballot_sync(0xffffffffu, lt));
}
// lb_start when the answer is expected a few dozen reads after lo: gallop, then bisect
__device__ __forceinline__ int64_t lb_start_near(const int32_t* __restrict__ R, int64_t lo, int64_t hi, int64_t v) {
    int64_t step = 32, top = lo;
    while (top + step < hi && (int64_t)__ldg(R + top + step - 1) < v) { top += step; step <<= 1; }
    return lb_start(R, top, min(hi, top + step), v);
}

constexpr int CS_HP = 128;                 // het positions staged per warp

__global__ void __launch_bounds__(128)
chain_size_kernel(const UnfzDnm* __restrict__ dnms, int32_t n_dnms, const int64_t* __restrict__ seg_pair_off,
                  UnfzSiteCols sites, UnfzReadCols reads, const UnfzReadSum* __restrict__ rsum,
                  const int32_t* __restrict__ blk_maxspan, const int32_t* __restrict__ het_list,
                  const int32_t* __restrict__ n_het, const uint32_t* __restrict__ cand_list,
                  const int32_t* __restrict__ n_cand, int32_t* __restrict__ win, int64_t* __restrict__ need,
                  int32_t* __restrict__ site_lo, int32_t* __restrict__ site_n, int32_t* __restrict__ seed_win,
                  const int32_t* __restrict__ guard) {
    UNFZ_GUARD(guard);
    __shared__ int32_t s_hp[4][CS_HP];
    const int lane = threadIdx.x & 31;
    const int d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (d >= n_dnms) return;
    int32_t* hp = s_hp[threadIdx.x >> 5];
    const UnfzDnm dn = dnms[d];
    int64_t nd[6] = {0, 0, 0, 0, 0, 0};
    int64_t a_lo = 0, a_hi = 0, b_lo = 0, b_hi = 0;
    const int nh = n_het[d], nc = n_cand[d];
    if (!(dn.kind == UNFZ_KIND_SKIP || dn.rblk < 0 || nc <= 0 || dn.seg_hi <= dn.seg_lo)) {
        const int64_t lbase = seg_pair_off[dn.seg_lo];
        const int32_t* H = het_list + lbase;
        const int64_t blk_lo = reads.blk_off[dn.rblk], blk_hi = reads.blk_off[dn.rblk + 1];
        const int64_t maxspan = blk_maxspan[dn.rblk];
        const double cul = reads.blk_cul[dn.rblk];
        const bool sv = dn.kind == UNFZ_KIND_SV;
        // seed fetch windows in position space
        int64_t sa_lo, sa_hi, sb_lo = 0, sb_hi = 0;
        if (sv) {
            const double l0 = (double)dn.pos - cul, l1 = (double)dn.end - cul;
            sa_lo = l0 > 0.0 ? (int64_t)l0 : 0; sa_hi = (int64_t)((double)dn.pos + cul);
            sb_lo = l1 > 0.0 ? (int64_t)l1 : 0; sb_hi = (int64_t)((double)dn.end + cul);
        } else {
            sa_lo = (dn.flags & 8) ? (int64_t)dn.pos : (int64_t)dn.pos - 1;
            sa_hi = (int64_t)dn.pos + 1;
        }
        const int64_t mid = sv ? ((int64_t)dn.pos + (int64_t)dn.end) / 2 : (int64_t)1 << 40;
        int64_t minA = sa_lo, maxA = sa_hi, minB = sv ? sb_lo : ((int64_t)1 << 40), maxB = sv ? sb_hi : -1;
        int64_t incs = 0;
        for (int i = lane; i < nh; i += 32) {
            const int64_t p = sites.pos[H[i]];
            if (i < CS_HP) hp[i] = (int32_t)p;
            if (p < mid) { minA = min(minA, p); maxA = max(maxA, p + 1); }
            else { minB = min(minB, p); maxB = max(maxB, p + 1); }
            const int64_t a = lb_start(reads.start, blk_lo, blk_hi, p - maxspan + 1);
            const int64_t b = lb_start_near(reads.start, a, blk_hi, p + 1);
            site_lo[lbase + i] = (int32_t)(a - blk_lo);
            site_n[lbase + i] = (int32_t)(b - a);
            incs += b - a;
        }
        minA = warp_min64(minA); maxA = warp_max64(maxA); minB = warp_min64(minB); maxB = warp_max64(maxB);
        nd[1] = warp_sum64(incs);
        __syncwarp();
        a_lo = warp_lb_start(reads.start, blk_lo, blk_hi, minA - maxspan + 1, lane);
        a_hi = warp_lb_start(reads.start, a_lo, blk_hi, maxA, lane);
        if (maxB > minB) {
            b_lo = warp_lb_start(reads.start, blk_lo, blk_hi, minB - maxspan + 1, lane);
            b_hi = warp_lb_start(reads.start, b_lo, blk_hi, maxB, lane);
            if (b_lo <= a_hi) { a_hi = max(a_hi, b_hi); a_lo = min(a_lo, b_lo); b_lo = b_hi = 0; }
        }
        nd[0] = ((a_hi - a_lo) + (b_hi - b_lo) + 3) & ~(int64_t)3;   // word-aligned label/evidence regions
        // seed entries and the het sites each entry registers (start <= pos <= end)
        int64_t seeds = 0, sincs = 0;
        for (int wdx = 0; wdx < (sv ? 2 : 1); ++wdx) {
            const int64_t flo = wdx ? sb_lo : sa_lo, fhi = wdx ? sb_hi : sa_hi;
            const int64_t a = warp_lb_start(reads.start, blk_lo, blk_hi, flo - maxspan + 1, lane);
            const int64_t b = warp_lb_start(reads.start, a, blk_hi, fhi, lane);
            if (lane == 0) { seed_win[4 * (int64_t)d + 2 * wdx] = (int32_t)a; seed_win[4 * (int64_t)d + 2 * wdx + 1] = (int32_t)b; }
            seeds += 2 * (b - a);
            for (int64_t r = a + lane; r < b; r += 32) {
                const int64_t ents[2] = {r, (int64_t)rsum[r].mate};
                for (int t = 0; t < 2; ++t) {
                    const int64_t e = ents[t];
                    if (e < 0) continue;
                    const int64_t st = rsum[e].start, en = rsum[e].end;
                    int l = 0, h = nh;
                    int u;
                    if (nh <= CS_HP) {                               // het positions staged per warp
                        while (l < h) { const int m2 = (l + h) >> 1; if (hp[m2] < st) l = m2 + 1; else h = m2; }
                        u = l; h = nh;
                        while (u < h) { const int m2 = (u + h) >> 1; if (hp[m2] <= en) u = m2 + 1; else h = m2; }
                    } else {
                        while (l < h) { const int m2 = (l + h) >> 1; if (sites.pos[H[m2]] < st) l = m2 + 1; else h = m2; }
                        u = l; h = nh;
                        while (u < h) { const int m2 = (u + h) >> 1; if (sites.pos[H[m2]] <= en) u = m2 + 1; else h = m2; }
                    }
                    sincs += u - l;
                }
            }
        }
        nd[2] = seeds;
        nd[3] = warp_sum64(sincs);
        nd[4] = nh;
        nd[5] = nc;
    }
    if (lane == 0) {
        win[d] = (int32_t)a_lo; win[(int64_t)n_dnms + d] = (int32_t)a_hi;
        win[2 * (int64_t)n_dnms + d] = (int32_t)b_lo; win[3 * (int64_t)n_dnms + d] = (int32_t)b_hi;
        for (int k = 0; k < 6; ++k) need[(int64_t)k * n_dnms + d] = nd[k];
    }
}

// ------------------------------------------------------------------------------------------------
// final call: unfazed.py summarize_record :190-334 on counts
// ------------------------------------------------------------------------------------------------
__device__ UnfzCall summarize_one(const UnfzDnm& dn, int nd_reads, int nm_reads, int nd_sites, int nm_sites,
                                  int cd, int cm, bool has_read_rec, bool has_cnv_rec, int ratio, bool include_ambiguous) {
    UnfzCall c;
    c.origin = UNFZ_ORIGIN_NONE; c.evidence_count = 0; c.evidence_types = 0; c.emitted = 0;
    const bool autoph = dn.flags & 1;
    const bool sv_quirk = dn.flags & 4;      // Q8: "SEX-CHROM,SEX-CHROM", cnv sites "NA" (len 2)
    if (autoph && !sv_quirk) {
        c.origin = (dn.flags & 2) ? UNFZ_ORIGIN_DAD : UNFZ_ORIGIN_MOM;
        c.evidence_count = 1;
        c.evidence_types = UNFZ_EV_SEX_CHROM;
        c.emitted = 1;
        return c;
    }
    if (!autoph && !has_read_rec && !has_cnv_rec) return c;
    if (sv_quirk) { nd_reads = nm_reads = nd_sites = nm_sites = 0; cd = cm = 2; }
    if (!has_read_rec && !sv_quirk) nd_reads = nm_reads = nd_sites = nm_sites = 0;
    if (!has_cnv_rec && !sv_quirk) cd = cm = 0;
    int origin = UNFZ_ORIGIN_NONE, count = 0, types = 0;
    bool ambig = false;
    if (nd_reads > 0 && nd_reads >= ratio * nm_reads) { origin = UNFZ_ORIGIN_DAD; count = nd_sites; types |= UNFZ_EV_READBACKED; }
    else if (nm_reads > 0 && nm_reads >= ratio * nd_reads) { origin = UNFZ_ORIGIN_MOM; count = nm_sites; types |= UNFZ_EV_READBACKED; }
    else if (nd_reads > 0 && nm_reads > 0) { origin = UNFZ_ORIGIN_BOTH; count = nd_reads + nm_reads; types |= UNFZ_EV_AMBIG_READBACKED; ambig = true; }
    if (cd > 0 && cd >= ratio * cm) {
        if (origin == UNFZ_ORIGIN_MOM && !(types & UNFZ_EV_READBACKED)) {
            origin = UNFZ_ORIGIN_NONE; count += cd + cm; types = UNFZ_EV_AMBIG_BOTH; ambig = true;
        } else {
            origin = UNFZ_ORIGIN_DAD; count = cd;
            if (types & UNFZ_EV_AMBIG_READBACKED) { types &= ~UNFZ_EV_AMBIG_READBACKED; ambig = false; }
            types |= UNFZ_EV_ALLELE_BALANCE;
        }
    } else if (cm > 0 && cm >= ratio * cd) {
        if (origin == UNFZ_ORIGIN_DAD && !(types & UNFZ_EV_READBACKED)) {
            origin = UNFZ_ORIGIN_NONE; count += cd + cm; types = UNFZ_EV_AMBIG_BOTH; ambig = true;
        } else {
            origin = UNFZ_ORIGIN_MOM; count = cm;
            if (types & UNFZ_EV_AMBIG_READBACKED) types &= ~UNFZ_EV_AMBIG_READBACKED;
            types |= UNFZ_EV_ALLELE_BALANCE;
        }
    } else if ((cd + cm) > 0 && !(types & UNFZ_EV_READBACKED)) {
        origin = UNFZ_ORIGIN_NONE; count += cd + cm; types |= UNFZ_EV_AMBIG_ALLELE_BAL; ambig = true;
    }
    c.origin = origin; c.evidence_count = count; c.evidence_types = types;
    c.emitted = ((origin == UNFZ_ORIGIN_NONE || ambig) && !include_ambiguous) ? 0 : 1;
    return c;
}

__global__ void summarize_kernel(const UnfzDnm* __restrict__ dnms, int32_t n_dnms, const UnfzTally* __restrict__ tally,
                                 const int32_t* __restrict__ cnv_dad, const int32_t* __restrict__ cnv_mom,
                                 const int32_t* __restrict__ n_cand, int ratio, UnfzCall* __restrict__ strict,
                                 UnfzCall* __restrict__ ambiguous) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n_dnms) return;
    const UnfzDnm dn = dnms[d];
    const UnfzTally t = tally[d];
    int cd = 0, cm = 0;
    bool has_cnv = false;
    if (dn.cnv_entry >= 0) {
        cd = cnv_dad[dn.cnv_entry];
        cm = cnv_mom[dn.cnv_entry];
        has_cnv = n_cand[dn.cnv_entry] > 0;
    }
    strict[d] = summarize_one(dn, t.n_dad_reads, t.n_mom_reads, t.n_dad_sites, t.n_mom_sites, cd, cm,
                              t.has_record != 0, has_cnv, ratio, false);
    ambiguous[d] = summarize_one(dn, t.n_dad_reads, t.n_mom_reads, t.n_dad_sites, t.n_mom_sites, cd, cm,
                                 t.has_record != 0, has_cnv, ratio, true);
}

template <typename T>
char* carve(char*& p, int64_t n) {
    char* out = p;
    p += ((n * (int64_t)sizeof(T)) + 15) & ~(int64_t)15;
    return out;
}

Scratch carve_all(char* base, int64_t slots, int64_t incs, int64_t seeds, int64_t sincs, int64_t hets, int64_t cands,
                  int64_t n_dnms, int64_t* total) {
    Scratch S;
    char* p = base;
    const int64_t pairs = incs + seeds, adjs = incs + sincs;
    S.slot = (SlotRec*)carve<SlotRec>(p, slots);                  // first: the base is 256-byte aligned
    S.p_minkey = (unsigned long long*)carve<unsigned long long>(p, pairs);
    S.inc_r = (int32_t*)carve<int32_t>(p, incs); S.inc_x = (int32_t*)carve<int32_t>(p, incs);
    S.inc_site = (int32_t*)carve<int32_t>(p, incs); S.inc_al = (uint8_t*)carve<uint8_t>(p, incs);
    S.seed_e = (int32_t*)carve<int32_t>(p, seeds); S.seed_hap = (uint8_t*)carve<uint8_t>(p, seeds);
    S.seed_reg = (uint32_t*)carve<uint32_t>(p, seeds);
    S.sinc_px = (int32_t*)carve<int32_t>(p, sincs); S.sinc_site = (int32_t*)carve<int32_t>(p, sincs);
    S.sinc_sidx = (uint32_t*)carve<uint32_t>(p, sincs); S.sinc_al = (uint8_t*)carve<uint8_t>(p, sincs);
    S.x_of = (int32_t*)carve<int32_t>(p, pairs); S.prim = (int32_t*)carve<int32_t>(p, pairs);
    S.p_ord = (uint32_t*)carve<uint32_t>(p, pairs); S.p_fpos = (int32_t*)carve<int32_t>(p, pairs);
    S.p_tmp = (int32_t*)carve<int32_t>(p, pairs); S.p_label = (uint8_t*)carve<uint8_t>(p, pairs);
    S.adj = (int32_t*)carve<int32_t>(p, adjs); S.adj_off = (int32_t*)carve<int32_t>(p, pairs + 2 * n_dnms + 2);
    S.front0 = (int32_t*)carve<int32_t>(p, pairs); S.front1 = (int32_t*)carve<int32_t>(p, pairs);
    S.spos = (int32_t*)carve<int32_t>(p, hets); S.sref = (uint8_t*)carve<uint8_t>(p, hets);
    S.salt = (uint8_t*)carve<uint8_t>(p, hets); S.site_off = (int32_t*)carve<int32_t>(p, hets + n_dnms + 1);
    S.cand_off = (int32_t*)carve<int32_t>(p, hets + n_dnms + 1);
    S.bestkey = (unsigned long long*)carve<unsigned long long>(p, hets);
    S.site_cnt = (int32_t*)carve<int32_t>(p, hets); S.site_base = (int32_t*)carve<int32_t>(p, hets);
    S.active = (int32_t*)carve<int32_t>(p, hets);
    S.cpos = (int32_t*)carve<int32_t>(p, cands);
    S.meta = (int32_t*)carve<int32_t>(p, n_dnms * 8);
    *total = (int64_t)(p - base);
    return S;
}

}  // namespace

extern "C" int64_t unfz_chain_scratch_bytes(int64_t slots, int64_t incs, int64_t seeds, int64_t seed_incs,
                                            int64_t het_sites, int64_t cand_sites, int64_t n_dnms) {
    int64_t total = 0;
    carve_all(nullptr, slots, incs, seeds, seed_incs, het_sites, cand_sites, n_dnms, &total);
    return total + 256;
}

extern "C" int unfz_chain_size(UnfzCtx* ctx, const UnfzDnm* dnms, int32_t n_dnms, const UnfzSegIn* segs,
                               const int64_t* seg_pair_off, const UnfzSiteCols* sites, const UnfzReadCols* reads,
                               const UnfzReadSum* rsum, const int32_t* blk_maxspan, const int32_t* het_list,
                               const int32_t* n_het, const uint32_t* cand_list, const int32_t* n_cand,
                               int32_t* win, int64_t* need, int32_t* site_lo, int32_t* site_n, int32_t* seed_win,
                               void* stream) {
    (void)segs;
    if (n_dnms <= 0) return 0;
    chain_size_kernel<<<(n_dnms + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
        dnms, n_dnms, seg_pair_off, *sites, *reads, rsum, blk_maxspan, het_list, n_het, cand_list, n_cand, win, need,
        site_lo, site_n, seed_win, ctx->guard);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

#ifdef CH_DEBUG
extern "C" int unfz_debug_chain(unsigned long long* out16) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_ch_dbg, sizeof(g_ch_dbg));
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_ch_dbg, z, sizeof(z));
    return 0;
}
#endif

extern "C" int unfz_chain_tally(UnfzCtx* ctx, const UnfzDnm* dnms, int32_t n_dnms, const UnfzSegIn* segs,
                                const int64_t* seg_pair_off, const UnfzSiteCols* sites, const UnfzReadCols* reads,
                                const UnfzReadSum* rsum, const int32_t* blk_maxspan, const uint32_t* hits,
                                const uint32_t* hit_tile_base, int32_t hit_tile_reads, const int32_t* mark_prefix, const int32_t* het_list, const int32_t* n_het,
                                const uint32_t* cand_list, const int32_t* n_cand, const uint8_t* alleles,
                                const int32_t* win, const int32_t* site_lo, const int32_t* site_n, const int32_t* seed_win,
                                const int64_t* off, const int64_t* h_totals, const UnfzParams* hp, void* scratch, int64_t scratch_bytes,
                                uint8_t* slot_label, uint8_t* slot_evid, uint8_t* cand_evid, UnfzTally* tally,
                                int64_t* ev_need, void* stream) {
    if (n_dnms <= 0) return 0;
    ChainArgs A;
    A.ev_need = ev_need;
    A.dnms = dnms; A.n_dnms = n_dnms; A.segs = segs; A.seg_pair_off = seg_pair_off;
    A.sites = *sites; A.reads = *reads; A.rsum = rsum; A.blk_maxspan = blk_maxspan; A.hits = hits; A.mp = mark_prefix;
    A.het_list = het_list; A.n_het = n_het; A.cand_list = cand_list; A.n_cand = n_cand; A.alleles = alleles;
    A.win = win; A.off = off; A.split_margin = hp->split_error_margin;
    A.hit_tile_base = hit_tile_base; A.hit_tile_reads = hit_tile_reads;
    A.site_lo = site_lo; A.site_n = site_n; A.seed_win = seed_win;
    A.readlen = hp->readlen;
    A.ext_goal = hp->ext_read_goal;
    A.no_extended = hp->no_extended;
    A.slot_label = slot_label; A.slot_evid = slot_evid; A.cand_evid = cand_evid; A.tally = tally;
    int64_t total = 0;
    uintptr_t base = ((uintptr_t)scratch + 255) & ~(uintptr_t)255;
    A.S = carve_all((char*)base, h_totals[0], h_totals[1], h_totals[2], h_totals[3], h_totals[4], h_totals[5], n_dnms, &total);
    if ((int64_t)(base - (uintptr_t)scratch) + total > scratch_bytes) return unfz_fail(ctx, -20, "chain scratch too small");
    A.guard = ctx->guard;
    if (!ctx->chain_carveout_set) {           // 8 CTAs x 25 KB of static shared memory per SM need the large carve-out
        UNFZ_CHECK(ctx, cudaFuncSetAttribute(cta_small::chain_bfs_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        UNFZ_CHECK(ctx, cudaFuncSetAttribute(cta_wide::chain_bfs_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ctx->chain_carveout_set = true;
    }
    // CTA shape: h_totals[1] is the incidence capacity of the batch (exact, or the engine's speculative bound).  A DNM with
    // thousands of (het site x read) incidences keeps 512 threads busy; measured on the 50 kb / 60x stress (about 20 k
    // incidences per DNM, 500 DNMs): 1.03 ms with 128-thread CTAs, 0.58 ms with 512; on 5 kb / 30x windows (450 per DNM,
    // 10 k DNMs) the small shape wins (0.56 vs 0.63 ms at 256 threads).
    bool wide = h_totals[1] / n_dnms >= UNFZ_CHAIN_WIDE_MIN_INC;
    if (const char* force = getenv("UNFZ_CHAIN_SHAPE")) wide = force[0] == 'w' ? true : (force[0] == 's' ? false : wide);   // tests
    if (wide) {
        cta_wide::chain_setup_kernel<<<n_dnms, cta_wide::CH_THREADS, 0, (cudaStream_t)stream>>>(A);
        UNFZ_LAUNCH_CHECK(ctx);
        cta_wide::chain_bfs_kernel<<<n_dnms, cta_wide::CH_THREADS, 0, (cudaStream_t)stream>>>(A);
        UNFZ_LAUNCH_CHECK(ctx);
        cta_wide::chain_evidence_kernel<<<n_dnms, cta_wide::CH_THREADS, 0, (cudaStream_t)stream>>>(A);
        UNFZ_LAUNCH_CHECK(ctx);
        return 0;
    }
    cta_small::chain_setup_kernel<<<n_dnms, cta_small::CH_THREADS, 0, (cudaStream_t)stream>>>(A);
    UNFZ_LAUNCH_CHECK(ctx);
    cta_small::chain_bfs_kernel<<<n_dnms, cta_small::CH_THREADS, 0, (cudaStream_t)stream>>>(A);
    UNFZ_LAUNCH_CHECK(ctx);
    cta_small::chain_evidence_kernel<<<n_dnms, cta_small::CH_THREADS, 0, (cudaStream_t)stream>>>(A);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int unfz_evidence_lists(UnfzCtx* ctx, const UnfzDnm* dnms, int32_t n_dnms, const int64_t* seg_pair_off,
                                   const UnfzSiteCols* sites, const uint32_t* cand_list, const int32_t* n_cand,
                                   const uint8_t* cand_evid, const int32_t* win, const int64_t* slot_off,
                                   const uint8_t* slot_evid, const int64_t* ev_off, int32_t* ev_read_dad, int32_t* ev_read_mom,
                                   int32_t* ev_pos_dad, int32_t* ev_pos_mom, void* stream) {
    if (n_dnms <= 0) return 0;
    evidence_lists_kernel<<<(n_dnms + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
        dnms, n_dnms, seg_pair_off, sites->pos, cand_list, n_cand, cand_evid, win, slot_off, slot_evid, ev_off,
        ev_read_dad, ev_read_mom, ev_pos_dad, ev_pos_mom, ctx->guard);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int unfz_summarize(UnfzCtx* ctx, const UnfzDnm* dnms, int32_t n_dnms, const UnfzTally* tally,
                              const int32_t* cnv_dad, const int32_t* cnv_mom, const int32_t* n_cand,
                              const UnfzParams* hp, UnfzCall* calls_strict, UnfzCall* calls_ambiguous, void* stream) {
    if (n_dnms <= 0) return 0;
    summarize_kernel<<<(n_dnms + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        dnms, n_dnms, tally, cnv_dad, cnv_mom, n_cand, hp->evidence_min_ratio, calls_strict, calls_ambiguous);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}
