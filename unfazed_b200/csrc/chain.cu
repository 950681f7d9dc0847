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
#include <type_traits>

#include "common.cuh"

namespace {

#ifndef CH_THREADS_N
#define CH_THREADS_N 128
#endif
constexpr int CH_THREADS = CH_THREADS_N;
constexpr int CH_WARPS = CH_THREADS / 32;
constexpr unsigned long long KEY_NONE = ~0ull;
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

// snv_match_alleles (:296-336) through get_allele_at (:56-73): 0 none, 1 ref, 2 alt
__device__ int seed_snv(const ChainArgs& A, const UnfzDnm& dn, int64_t r) {
    int64_t e = r;
    UnfzRead h = load_read(A.reads.hdr + e);
    int q = cigar_qpos2(A.reads.cigar + h.cigar_off, h.n_cigar, h.start, dn.pos);
    if (q < 0) {
        e = h.mate;
        h = load_read(A.reads.hdr + e);
        q = cigar_qpos2(A.reads.cigar + h.cigar_off, h.n_cigar, h.start, dn.pos);
        if (q < 0) return 0;
    }
    if (q < 4 || q > A.readlen - 4) return 0;
    const int n = dn.ref_len > dn.alt_len ? dn.ref_len : dn.alt_len;
    if (!(h.l_seq > q + n)) return 0;
    const int64_t g = read_qoff(h) + q;
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
    const int q = cigar_qpos2(cg, h.n_cigar, h.start, dn.pos);
    if (q < 0) return 0;
    const int n = dn.ref_len > dn.alt_len ? dn.ref_len : dn.alt_len;
    const int64_t g0 = read_qoff(h);
    for (int i = q; i < q + n && i < h.l_seq; ++i)
        if (plane_bit(A.reads.lowq, g0 + i)) return 0;
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

// exclusive prefix of `flag` over the CTA in thread order + total
__device__ __forceinline__ int block_prefix(bool flag, int* total) {
    __shared__ int wsum[CH_WARPS];
    const unsigned b = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) wsum[w] = __popc(b);
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < CH_WARPS; ++i) { if (i < w) base += wsum[i]; tot += wsum[i]; }
    __syncthreads();
    *total = tot;
    return base + __popc(b & ((1u << lane) - 1u));
}

// exclusive prefix SUM of v over the CTA in thread order + total
__device__ __forceinline__ int block_prefix_sum(int v, int* total) {
    __shared__ int wsum3[CH_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum3[w] = x;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < CH_WARPS; ++i) { if (i < w) base += wsum3[i]; tot += wsum3[i]; }
    __syncthreads();
    *total = tot;
    return base + x - v;
}

__device__ __forceinline__ int block_sum(int v) {
    __shared__ int wsum2[CH_WARPS];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) wsum2[threadIdx.x >> 5] = v;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int i = 0; i < CH_WARPS; ++i) tot += wsum2[i];
    __syncthreads();
    return tot;
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
// The breadth-first 2-colouring over DENSE pair ids, one instantiation per storage class:
//   <uint16_t, uint8_t>  everything in shared memory (the common case: a 5 kb window at 30x has a few
//                        hundred pairs and incidences) -- a level is a handful of shared-memory passes
//   <int32_t, int32_t>   everything in global scratch (deep / wide windows)
// A level touches only its FRONTIER (the pairs labelled by the previous level, through a pair-major
// adjacency of the incidences where the pair can read an allele) and the ACTIVE sites (those some
// frontier pair reaches), so a level costs what it does, not a rescan of every incidence.
// ------------------------------------------------------------------------------------------------
template <typename PT, typename ST>
struct BfsView {
    // per dense pair
    uint32_t* ord; int32_t* fpos; int32_t* tmp; unsigned long long* minkey; uint8_t* label;
    // per het-site incidence (site-major): pair, site, allele codes (bits0-1 finder, bits2-3 target)
    PT* inc_px; ST* inc_site; uint8_t* inc_al;
    // per seed incidence: pair, site, position in the pair's site list, allele codes
    PT* sinc_px; ST* sinc_site; uint32_t* sinc_sidx; uint8_t* sinc_al;
    // pair-major adjacency over finder-capable entries (entry e < n_inc: incidence e, else seed incidence e - n_inc)
    PT* adj; PT* adj_off;
    PT* front[2];
    // per het site
    const int32_t* spos; const int32_t* site_off; unsigned long long* bestkey; int32_t* site_cnt; int32_t* site_base;
    ST* active;
};

template <typename PT, typename ST>
__device__ void bfs_levels(const BfsView<PT, ST>& V, int nh, int n_inc, int n_front0) {
    __shared__ int s_nact, s_nnext;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int n_front = n_front0;
    PT* fcur = V.front[0];
    PT* fnext = V.front[1];
    for (;;) {
        for (int i = tid; i < nh; i += CH_THREADS) V.bestkey[i] = KEY_NONE;
        if (tid == 0) { s_nact = 0; s_nnext = 0; }
        __syncthreads();
        // a. best finder per site, over the frontier only; the finder's haplotype and allele ride in the low key bits
        for (int f = tid; f < n_front; f += CH_THREADS) {
            const int p = (int)fcur[f];
            const uint8_t fh = (V.label[p] & 2) ? 2 : 1;            // level 0: the "alt" visit comes first
            const unsigned long long okey = (unsigned long long)V.ord[p] << 36;
            const int32_t fp = V.fpos[p];
            const int q1 = (int)V.adj_off[p + 1];
            for (int q = (int)V.adj_off[p]; q < q1; ++q) {
                const int e = (int)V.adj[q];
                const bool sd = e >= n_inc;
                const int kk = sd ? e - n_inc : e;
                const int i = sd ? (int)V.sinc_site[kk] : (int)V.inc_site[kk];
                if (V.spos[i] == fp) continue;
                const uint8_t al = sd ? V.sinc_al[kk] : V.inc_al[kk];
                const uint32_t sidx = sd ? V.sinc_sidx[kk] : (uint32_t)i;   // registered sites come in site order
                atomicMin(V.bestkey + i, okey | ((unsigned long long)sidx << 4) | (unsigned long long)((fh << 2) | (al & 3)));
            }
        }
        __syncthreads();
        for (int i = tid; i < nh; i += CH_THREADS)
            if (V.bestkey[i] != KEY_NONE) V.active[atomicAdd(&s_nact, 1)] = (ST)i;
        __syncthreads();
        const int n_act = s_nact;
        if (n_act == 0) break;
        // b. earliest claiming key per unlabelled pair, over the active sites' incidences
        for (int a = warp; a < n_act; a += CH_WARPS) {
            const int i = (int)V.active[a];
            const unsigned long long bk = V.bestkey[i];
            const int k1 = V.site_off[i + 1];
            for (int k = V.site_off[i] + lane; k < k1; k += 32) {
                const int p = (int)V.inc_px[k];
                if (V.label[p] == 0 && (V.inc_al[k] >> 2)) atomicMin(V.minkey + p, bk);
            }
        }
        __syncthreads();
        // c. claim, one warp per active site, in the order of the site's read list
        for (int a = warp; a < n_act; a += CH_WARPS) {
            const int i = (int)V.active[a];
            const unsigned long long bk = V.bestkey[i];
            const int k0 = V.site_off[i], k1 = V.site_off[i + 1];
            int cnt = 0;
            for (int kb = k0; kb < k1; kb += 32) {
                const int k = kb + lane;
                bool claim = false;
                int p = 0;
                uint8_t ta = 0;
                if (k < k1) {
                    p = (int)V.inc_px[k];
                    ta = V.inc_al[k] >> 2;
                    claim = V.label[p] == 0 && ta && V.minkey[p] == bk;
                }
                const unsigned b = __ballot_sync(0xffffffffu, claim);
                if (claim) {
                    const uint8_t fh = (uint8_t)((bk >> 2) & 3), fa = (uint8_t)(bk & 3);
                    const uint8_t nh_ = (ta == fa) ? fh : (uint8_t)(3 - fh);
                    V.tmp[p] = (i << 8) | (nh_ << 4) | 1;
                    V.ord[p] = (uint32_t)(cnt + __popc(b & ((1u << lane) - 1u)));   // rank inside the site
                }
                cnt += __popc(b);
            }
            if (lane == 0) V.site_cnt[i] = cnt;
        }
        __syncthreads();
        // d. order of the active sites by key -> base offsets
        int assigned = 0;
        for (int a = tid; a < n_act; a += CH_THREADS) {
            const int i = (int)V.active[a];
            const int c = V.site_cnt[i];
            int base = 0;
            if (c > 0) {
                const unsigned long long bk = V.bestkey[i];
                for (int b = 0; b < n_act; ++b) {
                    const int j = (int)V.active[b];
                    if (V.site_cnt[j] > 0 && V.bestkey[j] < bk) base += V.site_cnt[j];
                }
            }
            V.site_base[i] = base;
            assigned += c;
        }
        assigned = block_sum(assigned);
        if (assigned == 0) break;
        // e. commit the new level; the newly labelled pairs are the next frontier
        for (int a = warp; a < n_act; a += CH_WARPS) {
            const int i = (int)V.active[a];
            const unsigned long long bk = V.bestkey[i];
            const int k1 = V.site_off[i + 1];
            for (int k = V.site_off[i] + lane; k < k1; k += 32) {
                const int p = (int)V.inc_px[k];
                if (V.label[p] != 0) continue;
                const int t = V.tmp[p];
                if (!(t & 1)) { V.minkey[p] = KEY_NONE; continue; }      // still unlabelled: re-arm for the next level
                if ((t >> 8) != i || V.minkey[p] != bk || !(V.inc_al[k] >> 2)) continue;
                const uint8_t nhap = (t >> 4) & 3;
                V.ord[p] = ((nhap == 1 ? 0u : 1u) << 24) | ((uint32_t)V.site_base[i] + V.ord[p]);   // deeper levels: "ref" list first
                V.fpos[p] = V.spos[i];
                V.tmp[p] = 0;
                V.label[p] = nhap;
                fnext[atomicAdd(&s_nnext, 1)] = (PT)p;
            }
        }
        __syncthreads();
        n_front = s_nnext;
        { PT* t_ = fcur; fcur = fnext; fnext = t_; }
        __syncthreads();                                   // s_nnext is cleared at the top of the next level
    }
    __syncthreads();
}

// The chaining of one DNM runs as three kernels, one CTA per DNM each, because its parts want opposite things:
//   chain_setup_kernel     seeds, het-site incidences, dense pair ids, allele codes, adjacency: chains of dependent
//                          gathers -> as many resident CTAs as possible (16 per SM, almost no shared memory)
//   chain_bfs_kernel       the level-synchronous colouring: no gathers at all once its few KB of state sit in shared
//                          memory -> few CTAs per SM, each fast
//   chain_evidence_kernel  matching against the informative sites + tallies: gathers again
// The hand-over between them is the per-DNM global scratch the wide-window mode uses anyway.
#ifndef CH_MINB
#define CH_MINB 16
#endif
constexpr int META_INTS = 8;              // per DNM: pairs (-1: nothing to do), incidences, seed incidences, level-0 frontier, status

__global__ void __launch_bounds__(CH_THREADS, CH_MINB)
chain_setup_kernel(ChainArgs A) {
    UNFZ_GUARD(A.guard);
#ifdef CH_DEBUG
    long long ch_t0 = clock64();
#endif
    const int d = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const UnfzDnm dn = A.dnms[d];
    const Scratch& S0 = A.S;

    UnfzTally T;
    T.n_dad_sites = T.n_mom_sites = T.n_dad_reads = T.n_mom_reads = 0;
    T.cnv_dad = T.cnv_mom = 0;
    T.has_record = 0;
    T.status = 0;
    const int nh = A.n_het[d], nc = A.n_cand[d];
    if (dn.kind == UNFZ_KIND_SKIP || dn.rblk < 0 || nc <= 0 || dn.seg_hi <= dn.seg_lo) {
        if (tid == 0) {
            A.tally[d] = T;
            if (A.ev_need) for (int q = 0; q < 4; ++q) A.ev_need[(int64_t)q * A.n_dnms + d] = 0;
            S0.meta[(int64_t)d * META_INTS] = -1;
        }
        return;
    }
    const int64_t lbase = A.seg_pair_off[dn.seg_lo];
    const int32_t* H = A.het_list + lbase;
    const uint32_t* C = A.cand_list + lbase;
    const int64_t* off = A.off;
    const int64_t n1 = (int64_t)A.n_dnms + 1;
    const int64_t o_slot = off[0 * n1 + d], o_inc = off[1 * n1 + d], o_seed = off[2 * n1 + d];
    const int64_t o_sinc = off[3 * n1 + d], o_het = off[4 * n1 + d], o_cand = off[5 * n1 + d];
    const int64_t cap_inc = off[1 * n1 + d + 1] - o_inc, cap_seed = off[2 * n1 + d + 1] - o_seed;
    const int64_t cap_sinc = off[3 * n1 + d + 1] - o_sinc;
    const int64_t o_pair = o_inc + o_seed, o_adj = o_inc + o_sinc;      // pairs <= incidences + seed entries

    // per-DNM views of the global scratch
    SlotRec* rec = S0.slot + o_slot;
    int32_t* inc_r = S0.inc_r + o_inc; int32_t* inc_x = S0.inc_x + o_inc; int32_t* inc_site = S0.inc_site + o_inc;
    int32_t* seed_e = S0.seed_e + o_seed; uint8_t* seed_hap = S0.seed_hap + o_seed; uint32_t* seed_reg = S0.seed_reg + o_seed;
    int32_t* x_of = S0.x_of + o_pair; int32_t* prim = S0.prim + o_pair;
    __shared__ int32_t sh_spos[CH_SMEM_SITES], sh_site_off[CH_SMEM_SITES + 1], sh_cand_off[CH_SMEM_SITES + 1];
    __shared__ int32_t sh_site_cnt[CH_SMEM_SITES], sh_site_base[CH_SMEM_SITES];
    __shared__ uint8_t sh_sref[CH_SMEM_SITES], sh_salt[CH_SMEM_SITES];
    const bool small_sites = nh <= CH_SMEM_SITES;
    int32_t* spos = small_sites ? sh_spos : S0.spos + o_het;
    uint8_t* sref = small_sites ? sh_sref : S0.sref + o_het;
    uint8_t* salt = small_sites ? sh_salt : S0.salt + o_het;
    int32_t* site_off = small_sites ? sh_site_off : S0.site_off + o_het + d;
    int32_t* cand_off = small_sites ? sh_cand_off : S0.cand_off + o_het + d;
    int32_t* site_cnt = small_sites ? sh_site_cnt : S0.site_cnt + o_het;
    int32_t* site_base = small_sites ? sh_site_base : S0.site_base + o_het;
    int32_t* cpos = S0.cpos + o_cand;

    const UnfzReadCols& R = A.reads;
    const int64_t blk_lo = R.blk_off[dn.rblk];
    // the read window is the union of two index ranges (second one only for far-apart SV breakpoints)
    const int64_t nd_ = A.n_dnms;
    const int64_t a_lo = A.win[d], a_hi = A.win[nd_ + d], b_lo = A.win[2 * nd_ + d], b_hi = A.win[3 * nd_ + d];
    auto slot_of = [&](int64_t r) -> int {
        if (r >= a_lo && r < a_hi) return (int)(r - a_lo);
        if (r >= b_lo && r < b_hi) return (int)((a_hi - a_lo) + (r - b_lo));
        return -1;
    };
    // canonical window slot of the pair a read belongs to (the lower read index inside the window)
    auto canon = [&](int64_t r) -> int {
        const int64_t m = rs_mate(A.rsum, r);
        const int sr = slot_of(r), sm = m >= 0 ? slot_of(m) : -1;
        if (sr < 0) return sm;
        if (sm < 0) return sr;
        return m < r ? sm : sr;
    };
    const double cul = R.blk_cul[dn.rblk];

    // a window slot only carries what maps a pair to its DENSE id (16 bytes, initialised lazily for the pairs
    // that get touched: seeds + registered reads); all chaining state is indexed by the dense id
    auto init_slot = [&](int x) { *reinterpret_cast<int4*>(rec + x) = make_int4(0, 0, -1, 0); };   // key 0, dense -1
#ifdef CH_PREFETCH   // measured on B200: SLOWER (0.638 vs 0.595 ms at 10 k DNMs) -- the set-up is bound by sector throughput, not latency
    // Everything below is chains of dependent gathers into the summaries and headers of the window's reads (and their
    // mates, which lie in the same window).  Their addresses are known now: ask L2 for the lines up front, so that the
    // gathers that follow find them there instead of paying a DRAM round trip each.
    {
        const char* p0 = reinterpret_cast<const char*>(A.rsum + a_lo);
        const char* p1 = reinterpret_cast<const char*>(A.rsum + a_hi);
        for (const char* q = p0 + (size_t)tid * 128; q < p1; q += (size_t)CH_THREADS * 128) prefetch_l2(q);
        p0 = reinterpret_cast<const char*>(R.hdr + a_lo);
        p1 = reinterpret_cast<const char*>(R.hdr + a_hi);
        for (const char* q = p0 + (size_t)tid * 128; q < p1; q += (size_t)CH_THREADS * 128) prefetch_l2(q);
        if (b_hi > b_lo) {
            p0 = reinterpret_cast<const char*>(A.rsum + b_lo);
            p1 = reinterpret_cast<const char*>(A.rsum + b_hi);
            for (const char* q = p0 + (size_t)tid * 128; q < p1; q += (size_t)CH_THREADS * 128) prefetch_l2(q);
            p0 = reinterpret_cast<const char*>(R.hdr + b_lo);
            p1 = reinterpret_cast<const char*>(R.hdr + b_hi);
            for (const char* q = p0 + (size_t)tid * 128; q < p1; q += (size_t)CH_THREADS * 128) prefetch_l2(q);
        }
    }
#endif
    for (int i = tid; i < nh; i += CH_THREADS) {
        const int64_t row = H[i];
        spos[i] = __ldg(A.sites.pos + row);
        sref[i] = __ldg(A.sites.ref + row);
        salt[i] = __ldg(A.sites.alt + row);
    }
    for (int j = tid; j < nc; j += CH_THREADS) {
        cpos[j] = __ldg(A.sites.pos + (int64_t)(C[j] & 0x3fffffffu));   // cand_evid is zeroed by the caller
    }
    __syncthreads();

    CH_MARK(0);
    // ---------------------------------------------------------------- phase 1: seed reads
    // seeds are kept as ENTRIES (one read per entry) in the order the reference appends them
    int n_seed = 0;
    if (dn.kind == UNFZ_KIND_SNV || dn.kind == UNFZ_KIND_INDEL) {
        // fetch(chrom, pos-1, pos+1); after a failed fetch the reference retries with (pos, pos+1) (Q24)
        const int64_t flo = (dn.flags & 8) ? (int64_t)dn.pos : (int64_t)dn.pos - 1;
        const int64_t lo = A.seed_win[4 * (int64_t)d], hi = A.seed_win[4 * (int64_t)d + 1];   // from chain_size
        for (int64_t base = lo; base < hi; base += CH_THREADS) {
            const int64_t r = base + tid;
            int hap = 0;
            if (r < hi && (int64_t)A.rsum[r].end > flo && pair_ok(A, r, false))
                hap = dn.kind == UNFZ_KIND_SNV ? seed_snv(A, dn, r) : seed_indel(A, dn, r);
            int tot;
            const int k = n_seed + 2 * block_prefix(hap != 0, &tot);
            if (hap && k + 1 < cap_seed) {
                seed_e[k] = (int32_t)r; seed_hap[k] = (uint8_t)hap;
                seed_e[k + 1] = rs_mate(A.rsum, r); seed_hap[k + 1] = (uint8_t)hap;
            }
            n_seed += 2 * tot;
        }
        if (n_seed > cap_seed) { n_seed = (int)cap_seed & ~1; T.status |= 1; }
    } else if (dn.kind == UNFZ_KIND_SV) {
        int64_t e_lo = 0, e_hi = 0, e_flo = 0, e_fhi = 0;
        for (int wdx = 0; wdx < 2; ++wdx) {
            const int64_t position = wdx == 0 ? dn.pos : dn.end;
            const double dlo = (double)position - cul;
            const int64_t flo = dlo > 0.0 ? (int64_t)dlo : 0;
            const int64_t fhi = (int64_t)((double)position + cul);
            const int64_t lo = A.seed_win[4 * (int64_t)d + 2 * wdx], hi = A.seed_win[4 * (int64_t)d + 2 * wdx + 1];
            if (wdx == 1) { e_lo = lo; e_hi = hi; e_flo = flo; e_fhi = fhi; }
            for (int64_t base = lo; base < hi; base += CH_THREADS) {
                const int64_t r = base + tid;
                int sup = 0;
                if (r < hi && (int64_t)A.rsum[r].end > flo && sv_goodok(A, r)) {
                    const UnfzRead h = load_read(R.hdr + r);
                    const int64_t m = h.mate;
                    // name already banned by the mate earlier in this fetch?
                    const bool skip = m < r && m >= lo && (int64_t)A.rsum[m].end > flo && sv_banned(A, m);
                    if (!skip && !sv_bad_ends(R.cigar + h.cigar_off, h.n_cigar)) sup = sv_support(A, dn, r, position, cul);
                }
                int tot;
                const int k = n_seed + 2 * block_prefix(sup != 0, &tot);
                if (sup && k + 1 < cap_seed) {
                    const int32_t m = rs_mate(A.rsum, r);
                    seed_e[k] = sup == 1 ? (int32_t)r : m; seed_hap[k] = 2;
                    seed_e[k + 1] = sup == 1 ? m : (int32_t)r; seed_hap[k + 1] = 2;
                }
                n_seed += 2 * tot;
            }
        }
        if (n_seed > cap_seed) { n_seed = (int)cap_seed & ~1; T.status |= 1; }
        __syncthreads();
        // drop entries whose name was banned while scanning the END breakpoint (:588-591), in place
        int kept = 0;
        for (int base = 0; base < n_seed; base += CH_THREADS) {
            const int k = base + tid;
            bool keep = false;
            int32_t e = 0;
            if (k < n_seed) {
                e = seed_e[k];
                const int64_t m = rs_mate(A.rsum, e);
                auto in_end_fetch = [&](int64_t z) {
                    return z >= e_lo && z < e_hi && (int64_t)A.rsum[z].end > e_flo && (int64_t)rs_start(A.rsum, z) < e_fhi;
                };
                const bool banned = (in_end_fetch(e) && sv_banned(A, e)) || (m >= 0 && in_end_fetch(m) && sv_banned(A, m));
                keep = !banned;
            }
            int tot;
            const int kk = kept + block_prefix(keep, &tot);     // block_prefix syncs: all reads precede the writes
            if (keep) { seed_e[kk] = e; seed_hap[kk] = 2; }
            kept += tot;
            __syncthreads();
        }
        n_seed = kept < 2 ? 0 : kept;
    }
    __syncthreads();

    CH_MARK(1);
    for (int k = tid; k < n_seed; k += CH_THREADS) { const int x = canon(seed_e[k]); if (x >= 0) init_slot(x); }
    __syncthreads();
    int n_inc = 0, n_sinc = 0;
    int P = 0;                                     // dense pairs
    const int nh_reg = A.no_extended ? 0 : nh;     // --no-extended: seeds are the haplotype lists
    if (!A.no_extended) {
        // ------------------------------------------------------------ phase 2: het-site incidences
        // (a) candidate read range of every het site: fetch(chrom, pos, pos+1) as index range
        for (int i = tid; i < nh; i += CH_THREADS) {          // found by chain_size, no search here
            site_base[i] = A.site_lo[lbase + i];
            site_cnt[i] = A.site_n[lbase + i];
        }
        __syncthreads();
        // (b) exclusive scan of the range sizes (warp 0), candidates are flattened site-major
        if (warp == 0) {
            int run = 0;
            for (int b0 = 0; b0 < nh; b0 += 32) {
                const int i = b0 + lane;
                const int v = i < nh ? site_cnt[i] : 0;
                int x = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                if (i < nh) cand_off[i] = run + x - v;
                run += __shfl_sync(0xffffffffu, x, 31);
            }
            if (lane == 0) cand_off[nh] = run;
        }
        for (int i = tid; i <= nh; i += CH_THREADS) site_off[i] = 0x7fffffff;
        __syncthreads();
        const int n_candidates = cand_off[nh];
    CH_MARK(2);
        // (c) one ordered pass over all (site, read) candidates.  The filter (three dependent gathers per
        // candidate) runs without any barrier: each batch of CH_THREADS candidates leaves only its four warp
        // ballots in shared memory; one scan over the ballots then gives every survivor its slot, and a second,
        // load-free pass writes them in candidate order.
        constexpr int CH_SB = 32;                                  // batches per round (ballots kept in smem)
        __shared__ uint32_t s_bal[CH_SB * CH_WARPS];
        __shared__ int32_t s_bpre[CH_SB * CH_WARPS];
        __shared__ int32_t s_btot;
        for (int c0 = 0; c0 < n_candidates; c0 += CH_SB * CH_THREADS) {
            const int nb = min(CH_SB, (n_candidates - c0 + CH_THREADS - 1) / CH_THREADS);
            int i0 = 0;                                            // site of this thread's first candidate
            if (c0 + tid < n_candidates) {
                int lo = 0, hi = nh;                               // last i with cand_off[i] <= c
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cand_off[mid] <= c0 + tid) lo = mid; else hi = mid; }
                i0 = lo;
            }
            int i = i0;
#pragma unroll 4
            for (int b = 0; b < nb; ++b) {
                const int c = c0 + b * CH_THREADS + tid;
                bool ok = false;
                if (c < n_candidates) {
                    while (cand_off[i + 1] <= c) ++i;              // candidates are site-major: the cursor only moves forward
                    const int64_t r = blk_lo + site_base[i] + (c - cand_off[i]);
                    const int32_t p = spos[i];
                    bool ov = A.rsum[r].end > p;
                    if (ov && site_cnt[i] > A.ext_goal) {          // Q2: `i > EXTENDED_RB_READ_GOAL` (never in practice)
                        int idx = 0;
                        for (int64_t q = blk_lo + site_base[i]; q < r; ++q) idx += A.rsum[q].end > p;
                        ov = idx <= A.ext_goal;
                    }
                    ok = ov && pair_ok(A, r, true);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, ok);
                if (lane == 0) s_bal[b * CH_WARPS + warp] = bal;
            }
            __syncthreads();
            if (warp == 0) {                                       // exclusive prefix of the ballot populations
                int run = 0;
                for (int j0 = 0; j0 < nb * CH_WARPS; j0 += 32) {
                    const int j = j0 + lane;
                    const int v = j < nb * CH_WARPS ? __popc(s_bal[j]) : 0;
                    int x = v;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
                    if (j < nb * CH_WARPS) s_bpre[j] = run + x - v;
                    run += __shfl_sync(0xffffffffu, x, 31);
                }
                if (lane == 0) s_btot = run;
            }
            __syncthreads();
            i = i0;
            for (int b = 0; b < nb; ++b) {
                const int c = c0 + b * CH_THREADS + tid;
                const unsigned bal = s_bal[b * CH_WARPS + warp];
                if (c < n_candidates && ((bal >> lane) & 1u)) {
                    while (cand_off[i + 1] <= c) ++i;
                    const int64_t r = blk_lo + site_base[i] + (c - cand_off[i]);
                    const int k = n_inc + s_bpre[b * CH_WARPS + warp] + __popc(bal & ((1u << lane) - 1u));
                    if (k < cap_inc) {
                        inc_r[k] = (int32_t)r;
                        inc_x[k] = canon(r);
                        inc_site[k] = i;
                    }
                }
            }
            n_inc += s_btot;
            __syncthreads();                                       // the ballots are reused by the next round
        }
    CH_MARK(3);
        if (n_inc > cap_inc) { n_inc = (int)cap_inc; T.status |= 2; }
        __syncthreads();
        for (int k = tid; k < n_inc; k += CH_THREADS) {
            if (k == 0 || inc_site[k - 1] != inc_site[k]) site_off[inc_site[k]] = k;
            init_slot(inc_x[k]);                                    // idempotent
        }
        __syncthreads();
        // fetched_reads[name] = [read, mate]: the last writer (highest site) wins (Q18)
        for (int k = tid; k < n_inc; k += CH_THREADS)
            atomicMax(&rec[inc_x[k]].key, ((unsigned long long)(uint32_t)(inc_site[k] + 1) << 32) | (uint32_t)inc_r[k]);
        __syncthreads();
        // the winning incidence of a pair gives it its dense id (ids in incidence order) and its primary read
        for (int base = 0; base < n_inc; base += CH_THREADS) {
            const int k = base + tid;
            bool rep = false;
            int x = 0;
            if (k < n_inc) {
                x = inc_x[k];
                rep = rec[x].key == (((unsigned long long)(uint32_t)(inc_site[k] + 1) << 32) | (uint32_t)inc_r[k]);
            }
            int tot;
            const int id = P + block_prefix(rep, &tot);
            if (rep) { rec[x].dense = id; x_of[id] = x; prim[id] = inc_r[k]; }
            P += tot;
        }
        if (tid == 0) {                                             // empty sites inherit the next offset
            int nxt = n_inc;
            site_off[nh] = n_inc;
            for (int i = nh - 1; i >= 0; --i) { if (site_off[i] == 0x7fffffff) site_off[i] = nxt; else nxt = site_off[i]; }
        }
        __syncthreads();
    }
    CH_MARK(4);
    // ------------------------------------------------------------ phase 3: seed registration
    // Entries register in the order "ref" list then "alt" list (:226-249); the level-0 visit order
    // is "alt" list first (Q19).  Everything is resolved with order keys instead of a serial loop:
    //   reg(k)   = position of entry k in the registration order
    //   prim[p]  = entry with the largest reg  (last writer wins)
    //   ord[p]   = smallest visit position of the pair
    int n_ref_entries = 0;
    for (int base = 0; base < n_seed; base += CH_THREADS) {
        const int k = base + tid;
        int tot;
        block_prefix(k < n_seed && seed_hap[k] == 1, &tot);
        n_ref_entries += tot;
    }
    for (int k = tid; k < n_inc; k += CH_THREADS) rec[inc_x[k]].key = 0ull;
    for (int k = tid; k < n_seed; k += CH_THREADS) { const int x = canon(seed_e[k]); if (x >= 0) rec[x].key = 0ull; }
    __syncthreads();
    {
        // pass A: registration keys; the entry with the largest one is the pair's last writer
        int ref_seen = 0, alt_seen = 0;
        for (int base = 0; base < n_seed; base += CH_THREADS) {
            const int k = base + tid;
            const bool live = k < n_seed;
            const int hap = live ? seed_hap[k] : 0;
            int tot_ref, tot_alt;
            const int pr = block_prefix(hap == 1, &tot_ref);
            const int pa = block_prefix(hap == 2, &tot_alt);
            if (live) {
                const int32_t e = seed_e[k];
                const int x = canon(e);
                const unsigned reg = hap == 1 ? (unsigned)(ref_seen + pr) : (unsigned)(n_ref_entries + alt_seen + pa);
                // visit order at level 0: alt entries first, then ref entries
                const unsigned visit = hap == 2 ? (unsigned)(alt_seen + pa) : (unsigned)(ref_seen + pr);
                seed_reg[k] = reg;
                seed_e[k] = x >= 0 ? e : -1 - e;                       // entries outside the window are inert
                if (x >= 0) atomicMax(&rec[x].key, ((unsigned long long)(reg + 1u) << 32) | (uint32_t)e);
                (void)visit;
            }
            ref_seen += tot_ref;
            alt_seen += tot_alt;
        }
        __syncthreads();
        // pass B: dense ids of the pairs only seeds touch; the last writer's read becomes the primary read
        for (int base = 0; base < n_seed; base += CH_THREADS) {
            const int k = base + tid;
            bool rep = false, fresh = false;
            int x = -1;
            int32_t e = -1;
            if (k < n_seed && (e = seed_e[k]) >= 0) {
                x = canon(e);
                rep = rec[x].key == (((unsigned long long)(seed_reg[k] + 1u) << 32) | (uint32_t)e);
                fresh = rep && rec[x].dense < 0;
            }
            int tot;
            const int id = P + block_prefix(fresh, &tot);
            if (fresh) { rec[x].dense = id; x_of[id] = x; }
            if (rep) prim[fresh ? id : rec[x].dense] = e;
            P += tot;
        }
        __syncthreads();
    }
    if ((int64_t)P > cap_inc + cap_seed) { P = (int)(cap_inc + cap_seed); T.status |= 8; }     // cannot happen: one id per entry at most

    // the rest of the set-up is the same code over either view
    int n_front0 = 0;
    auto setup = [&](auto& V) {
        using PT = typename std::remove_reference<decltype(V.inc_px[0])>::type;
        using ST = typename std::remove_reference<decltype(V.inc_site[0])>::type;
        __shared__ int s_front;
        for (int p = tid; p < P; p += CH_THREADS) {
            V.ord[p] = 0xffffffffu; V.fpos[p] = -1; V.tmp[p] = 0; V.minkey[p] = KEY_NONE; V.label[p] = 0;
        }
        if (tid == 0) s_front = 0;
        // incidences: window slot -> dense pair (in global mode the column is rewritten in place)
        for (int k = tid; k < n_inc; k += CH_THREADS) {
            const int px = rec[inc_x[k]].dense;
            const int st = inc_site[k];
            V.inc_px[k] = (PT)px;
            V.inc_site[k] = (ST)st;
        }
        __syncthreads();
        // pass C over the seed entries: labels, visit order, level-0 frontier, seed incidences
        int ref_seen = 0, alt_seen = 0, ns_total = 0;
        for (int base = 0; base < n_seed; base += CH_THREADS) {
            const int k = base + tid;
            const bool live = k < n_seed;
            const int hap = live ? seed_hap[k] : 0;
            int tot_ref, tot_alt;
            const int pr = block_prefix(hap == 1, &tot_ref);
            const int pa = block_prefix(hap == 2, &tot_alt);
            int n_match = 0, piv = -1, px = -1;
            int32_t st = 0, en = 0;
            unsigned reg = 0;
            if (live && seed_e[k] >= 0) {
                const int32_t e = seed_e[k];
                const int x = canon(e);
                px = rec[x].dense;
                reg = seed_reg[k];
                const unsigned visit = hap == 2 ? (unsigned)(alt_seen + pa) : (unsigned)(ref_seen + pr);
                atomicMin(V.ord + px, (hap == 2 ? 0u : (1u << 24)) | (visit & 0xffffffu));
                atomicOr(V.tmp + px, hap);                          // haplotype bits gather in tmp (word atomics), see below
                if (rec[x].key == (((unsigned long long)(reg + 1u) << 32) | (uint32_t)e))     // once per pair
                    V.front[0][atomicAdd(&s_front, 1)] = (PT)px;
                if (nh_reg > 0) {
                    st = rs_start(A.rsum, e);
                    en = A.rsum[e].end;
                    piv = bisect_pivot(spos, nh_reg, st, en);
                    if (piv >= 0) {
                        n_match = 1;
                        for (int j = piv + 1; j < nh_reg && st <= spos[j] && spos[j] <= en; ++j) ++n_match;
                        for (int j = piv - 1; j >= 0 && st <= spos[j] && spos[j] <= en; --j) ++n_match;
                    }
                }
            }
            // seed incidences may be stored in any order: their position in read_sites[x] is carried by
            // the order key (after all registered sites, then by registration order, then by the
            // pivot / right / left order of binary_search, Q16)
            int tm;
            const int off_m = block_prefix_sum(n_match, &tm);
            if (n_match > 0) {
                int w = 0;
                const unsigned keybase = (unsigned)nh + reg * (unsigned)(nh + 1);
                const int dst0 = ns_total + off_m;
                auto push = [&](int i) {
                    const int dst = dst0 + w;
                    if (dst < cap_sinc) { V.sinc_px[dst] = (PT)px; V.sinc_site[dst] = (ST)i; V.sinc_sidx[dst] = keybase + (unsigned)w; }
                    ++w;
                };
                push(piv);
                for (int j = piv + 1; j < nh_reg && st <= spos[j] && spos[j] <= en; ++j) push(j);
                for (int j = piv - 1; j >= 0 && st <= spos[j] && spos[j] <= en; --j) push(j);
            }
            ref_seen += tot_ref;
            alt_seen += tot_alt;
            ns_total += tm;
        }
        n_sinc = ns_total;
        if (n_sinc > cap_sinc) { n_sinc = (int)cap_sinc; T.status |= 4; }
        __syncthreads();
        n_front0 = s_front;
        for (int p = tid; p < P; p += CH_THREADS) { V.label[p] = (uint8_t)V.tmp[p]; V.tmp[p] = 0; }
        __syncthreads();
        if (A.no_extended) return;
    CH_MARK(5);
        // -------------------------------------------------------- phase 3.5: allele codes + pair-major adjacency
        // tmp[p] counts the entries where pair p can read an allele (finder role), then serves as the fill cursor
        for (int k = tid; k < n_inc; k += CH_THREADS) {
            const int i = (int)V.inc_site[k], px = (int)V.inc_px[k];
            const uint8_t al = allele_info(A, S0, (int64_t)prim[px], (int64_t)H[i], (char)sref[i], (char)salt[i]);
            V.inc_al[k] = al;
            if (al & 3) atomicAdd(V.tmp + px, 1);
        }
        for (int k = tid; k < n_sinc; k += CH_THREADS) {
            const int i = (int)V.sinc_site[k], px = (int)V.sinc_px[k];
            const uint8_t al = allele_info(A, S0, (int64_t)prim[px], (int64_t)H[i], (char)sref[i], (char)salt[i]);
            V.sinc_al[k] = al;
            if (al & 3) atomicAdd(V.tmp + px, 1);
        }
        __syncthreads();
        int run = 0;
        for (int base = 0; base < P; base += CH_THREADS) {
            const int p = base + tid;
            const int c = p < P ? V.tmp[p] : 0;
            int tot;
            const int o = run + block_prefix_sum(c, &tot);
            if (p < P) { V.adj_off[p] = (PT)o; V.tmp[p] = 0; }
            run += tot;
        }
        if (tid == 0) V.adj_off[P] = (PT)run;
        __syncthreads();
        for (int k = tid; k < n_inc; k += CH_THREADS)
            if (V.inc_al[k] & 3) { const int px = (int)V.inc_px[k]; V.adj[(int)V.adj_off[px] + atomicAdd(V.tmp + px, 1)] = (PT)k; }
        for (int k = tid; k < n_sinc; k += CH_THREADS)
            if (V.sinc_al[k] & 3) { const int px = (int)V.sinc_px[k]; V.adj[(int)V.adj_off[px] + atomicAdd(V.tmp + px, 1)] = (PT)(n_inc + k); }
        __syncthreads();
        for (int p = tid; p < P; p += CH_THREADS) V.tmp[p] = 0;
        __syncthreads();
    };
    BfsView<int32_t, int32_t> VG;
    VG.ord = S0.p_ord + o_pair; VG.fpos = S0.p_fpos + o_pair; VG.tmp = S0.p_tmp + o_pair; VG.minkey = S0.p_minkey + o_pair;
    VG.label = S0.p_label + o_pair;
    VG.inc_px = inc_x; VG.inc_site = inc_site; VG.inc_al = S0.inc_al + o_inc;          // the slot column becomes the pair column
    VG.sinc_px = S0.sinc_px + o_sinc; VG.sinc_site = S0.sinc_site + o_sinc;
    VG.sinc_sidx = S0.sinc_sidx + o_sinc; VG.sinc_al = S0.sinc_al + o_sinc;
    VG.adj = S0.adj + o_adj; VG.adj_off = S0.adj_off + o_pair + 2 * (int64_t)d;
    VG.front[0] = S0.front0 + o_pair; VG.front[1] = S0.front1 + o_pair;
    VG.spos = spos; VG.site_off = site_off; VG.bestkey = nullptr; VG.site_cnt = site_cnt; VG.site_base = site_base;
    VG.active = S0.active + o_het;

    setup(VG);
    __syncthreads();
    // hand-over: what the next two kernels need and only shared memory holds
    if (small_sites) {
        int32_t* g_spos = S0.spos + o_het;
        int32_t* g_off = S0.site_off + o_het + d;
        for (int i = tid; i < nh; i += CH_THREADS) g_spos[i] = spos[i];
        for (int i = tid; i <= nh; i += CH_THREADS) g_off[i] = A.no_extended ? 0 : site_off[i];
    }
    if (tid == 0) {
        int32_t* m = S0.meta + (int64_t)d * META_INTS;
        m[0] = P; m[1] = n_inc; m[2] = n_sinc; m[3] = n_front0; m[4] = T.status;
    }
    CH_MARK(6);
}

// per-pair state, incidence views and adjacency of the shared-memory mode
constexpr int SM_P = 512;                 // dense pairs
constexpr int SM_I = 768;                 // het-site incidences
constexpr int SM_SI = 256;                // seed incidences
#ifndef CH_BFS_MINB
#define CH_BFS_MINB 8
#endif

__global__ void __launch_bounds__(CH_THREADS, CH_BFS_MINB)
chain_bfs_kernel(ChainArgs A) {
    UNFZ_GUARD(A.guard);
    const int d = blockIdx.x;
    const int tid = threadIdx.x;
    const Scratch& S0 = A.S;
    const int32_t* meta = S0.meta + (int64_t)d * META_INTS;
    const int P = meta[0], n_inc = meta[1], n_sinc = meta[2], n_front0 = meta[3];
    if (P <= 0 || n_front0 <= 0 || A.no_extended) return;          // nothing can spread (labels of the seeds are final)
    const int nh = A.n_het[d];
    const int64_t* off = A.off;
    const int64_t n1 = (int64_t)A.n_dnms + 1;
    const int64_t o_inc = off[1 * n1 + d], o_seed = off[2 * n1 + d], o_sinc = off[3 * n1 + d], o_het = off[4 * n1 + d];
    const int64_t o_pair = o_inc + o_seed, o_adj = o_inc + o_sinc;
    BfsView<int32_t, int32_t> G;
    G.ord = S0.p_ord + o_pair; G.fpos = S0.p_fpos + o_pair; G.tmp = S0.p_tmp + o_pair; G.minkey = S0.p_minkey + o_pair;
    G.label = S0.p_label + o_pair;
    G.inc_px = S0.inc_x + o_inc; G.inc_site = S0.inc_site + o_inc; G.inc_al = S0.inc_al + o_inc;
    G.sinc_px = S0.sinc_px + o_sinc; G.sinc_site = S0.sinc_site + o_sinc;
    G.sinc_sidx = S0.sinc_sidx + o_sinc; G.sinc_al = S0.sinc_al + o_sinc;
    G.adj = S0.adj + o_adj; G.adj_off = S0.adj_off + o_pair + 2 * (int64_t)d;
    G.front[0] = S0.front0 + o_pair; G.front[1] = S0.front1 + o_pair;
    G.spos = S0.spos + o_het; G.site_off = S0.site_off + o_het + d; G.bestkey = S0.bestkey + o_het;
    G.site_cnt = S0.site_cnt + o_het; G.site_base = S0.site_base + o_het; G.active = S0.active + o_het;
    const int n_adj = G.adj_off[P];
    if (!(nh <= CH_SMEM_SITES && P <= SM_P && n_inc <= SM_I && n_sinc <= SM_SI)) {
        bfs_levels(G, nh, n_inc, n_front0);                        // wide window: the state stays in global scratch
        return;
    }
    // stage the DNM's chaining state (a few KB, contiguous arrays) into shared memory, colour there, write the labels back
    __shared__ __align__(8) unsigned long long sm_minkey[SM_P], sm_bestkey[CH_SMEM_SITES];
    __shared__ uint32_t sm_ord[SM_P], sm_sinc_sidx[SM_SI];
    __shared__ int32_t sm_fpos[SM_P], sm_tmp[SM_P], sm_spos[CH_SMEM_SITES], sm_site_off[CH_SMEM_SITES + 1];
    __shared__ int32_t sm_site_cnt[CH_SMEM_SITES], sm_site_base[CH_SMEM_SITES];
    __shared__ uint16_t sm_inc_px[SM_I], sm_sinc_px[SM_SI], sm_adj[SM_I + SM_SI], sm_adj_off[SM_P + 2], sm_front[2][SM_P];
    __shared__ uint8_t sm_label[SM_P], sm_inc_site[SM_I], sm_inc_al[SM_I], sm_sinc_site[SM_SI], sm_sinc_al[SM_SI];
    __shared__ uint8_t sm_active[CH_SMEM_SITES];
    for (int p = tid; p < P; p += CH_THREADS) {
        sm_ord[p] = G.ord[p]; sm_label[p] = G.label[p]; sm_adj_off[p] = (uint16_t)G.adj_off[p];
        sm_fpos[p] = -1; sm_tmp[p] = 0; sm_minkey[p] = KEY_NONE;
    }
    if (tid == 0) sm_adj_off[P] = (uint16_t)n_adj;
    for (int k = tid; k < n_inc; k += CH_THREADS) {
        sm_inc_px[k] = (uint16_t)G.inc_px[k]; sm_inc_site[k] = (uint8_t)G.inc_site[k]; sm_inc_al[k] = G.inc_al[k];
    }
    for (int k = tid; k < n_sinc; k += CH_THREADS) {
        sm_sinc_px[k] = (uint16_t)G.sinc_px[k]; sm_sinc_site[k] = (uint8_t)G.sinc_site[k];
        sm_sinc_sidx[k] = G.sinc_sidx[k]; sm_sinc_al[k] = G.sinc_al[k];
    }
    for (int q = tid; q < n_adj; q += CH_THREADS) sm_adj[q] = (uint16_t)G.adj[q];
    for (int f = tid; f < n_front0; f += CH_THREADS) sm_front[0][f] = (uint16_t)G.front[0][f];
    for (int i = tid; i < nh; i += CH_THREADS) sm_spos[i] = G.spos[i];
    for (int i = tid; i <= nh; i += CH_THREADS) sm_site_off[i] = G.site_off[i];
    __syncthreads();
    BfsView<uint16_t, uint8_t> V;
    V.ord = sm_ord; V.fpos = sm_fpos; V.tmp = sm_tmp; V.minkey = sm_minkey; V.label = sm_label;
    V.inc_px = sm_inc_px; V.inc_site = sm_inc_site; V.inc_al = sm_inc_al;
    V.sinc_px = sm_sinc_px; V.sinc_site = sm_sinc_site; V.sinc_sidx = sm_sinc_sidx; V.sinc_al = sm_sinc_al;
    V.adj = sm_adj; V.adj_off = sm_adj_off; V.front[0] = sm_front[0]; V.front[1] = sm_front[1];
    V.spos = sm_spos; V.site_off = sm_site_off; V.bestkey = sm_bestkey; V.site_cnt = sm_site_cnt; V.site_base = sm_site_base;
    V.active = sm_active;
    bfs_levels(V, nh, n_inc, n_front0);
    for (int p = tid; p < P; p += CH_THREADS) G.label[p] = sm_label[p];
}

__global__ void __launch_bounds__(CH_THREADS, CH_MINB)
chain_evidence_kernel(ChainArgs A) {
    UNFZ_GUARD(A.guard);
#ifdef CH_DEBUG
    long long ch_t0 = clock64();
#endif
    const int d = blockIdx.x;
    const int tid = threadIdx.x;
    const Scratch& S0 = A.S;
    const int32_t* meta = S0.meta + (int64_t)d * META_INTS;
    const int P = meta[0];
    if (P < 0) return;                                 // skipped by the set-up kernel (tally already written)
    const UnfzDnm dn = A.dnms[d];
    UnfzTally T;
    T.n_dad_sites = T.n_mom_sites = T.n_dad_reads = T.n_mom_reads = 0;
    T.cnv_dad = T.cnv_mom = 0;
    T.has_record = 0;
    T.status = meta[4];
    const int nc = A.n_cand[d];
    const int64_t lbase = A.seg_pair_off[dn.seg_lo];
    const uint32_t* C = A.cand_list + lbase;
    const int64_t* off = A.off;
    const int64_t n1 = (int64_t)A.n_dnms + 1;
    const int64_t o_slot = off[0 * n1 + d], o_inc = off[1 * n1 + d], o_seed = off[2 * n1 + d], o_cand = off[5 * n1 + d];
    const int64_t o_pair = o_inc + o_seed;
    uint8_t* label_out = A.slot_label + o_slot; uint8_t* evid_out = A.slot_evid + o_slot;
    const int32_t* x_of = S0.x_of + o_pair; const int32_t* prim = S0.prim + o_pair;
    const int32_t* cpos = S0.cpos + o_cand;
    uint8_t* cev = A.cand_evid + lbase;

    CH_MARK(7);
    // ---------------------------------------------------------------- phase 5: matching + evidence
    int has_rec = 0, dr = 0, mr = 0;
    const uint8_t* final_label = S0.p_label + o_pair;
    for (int p = tid; p < P; p += CH_THREADS) {
        const uint8_t lab = final_label[p];
        if (!lab) continue;
        const int64_t e0 = prim[p];
        const int64_t ents[2] = {e0, (int64_t)rs_mate(A.rsum, e0)};
        uint8_t ev = 0;
        for (int t = 0; t < 2; ++t) {
            const int64_t e = ents[t];
            if (e < 0) continue;
            const int32_t st = rs_start(A.rsum, e), en = A.rsum[e].end;
            int lb = 0, hi = nc;
            while (lb < hi) { const int mid = (lb + hi) >> 1; if (cpos[mid] < st) lb = mid + 1; else hi = mid; }
            if (lb >= nc || cpos[lb] >= en) continue;           // binary_search finds nothing
            int ub = lb;
            while (ub < nc && cpos[ub] <= en) ++ub;              // neighbour rule (Q16)
            bool consistent = true;
            for (int j = lb + 1; j < ub; ++j)
                if ((C[j] ^ C[lb]) & 0x80000000u) { consistent = false; break; }
            if (!consistent) continue;
            has_rec = 1;
            for (int j = lb; j < ub; ++j) {
                const uint32_t cw = C[j];
                const int64_t row = cw & 0x3fffffffu;
                const uint32_t h = hit_lookup(A, e, row);
                if (!(h & 0xffffu)) continue;
                const char c = hit_char(h);
                bool origin_ref;
                if (c == (char)__ldg(A.sites.ref + row)) origin_ref = true;
                else if (c == (char)__ldg(A.sites.alt + row)) origin_ref = false;
                else continue;
                const bool alt_is_dad = (cw & 0x80000000u) != 0;
                uint8_t bits = 0;
                for (int hp = 1; hp <= 2; ++hp) {
                    if (!(lab & hp)) continue;
                    const bool to_alt = origin_ref == (hp == 1);
                    bits |= (to_alt == alt_is_dad) ? 1 : 2;       // 1: dad, 2: mom
                }
                ev |= bits;
                unsigned* wp = reinterpret_cast<unsigned*>(reinterpret_cast<uintptr_t>(cev + j) & ~(uintptr_t)3);
                atomicOr(wp, (unsigned)bits << (8 * (reinterpret_cast<uintptr_t>(cev + j) & 3)));
            }
        }
        const int x = x_of[p];
        label_out[x] = lab;                                      // zero-filled by the caller: only labelled pairs are written
        evid_out[x] = ev;
        dr += ev & 1;
        mr += (ev >> 1) & 1;
    }
    __syncthreads();

    CH_MARK(8);
    // ---------------------------------------------------------------- phase 6: tally
    int ds = 0, ms = 0, esd = 0, esm = 0;
    for (int j = tid; j < nc; j += CH_THREADS) {
        const uint8_t b = cev[j];
        esd += b & 1;
        esm += (b >> 1) & 1;
        for (int bit = 1; bit <= 2; ++bit) {
            if (!(b & bit)) continue;
            bool first = true;                      // unique str(pos): count the first duplicate only
            for (int j2 = j - 1; j2 >= 0 && cpos[j2] == cpos[j]; --j2)
                if (cev[j2] & bit) { first = false; break; }
            if (first) { if (bit == 1) ++ds; else ++ms; }
        }
    }
    if (A.ev_need) { esd = block_sum(esd); esm = block_sum(esm); }
    ds = block_sum(ds); ms = block_sum(ms); dr = block_sum(dr); mr = block_sum(mr);
    has_rec = block_sum(has_rec);
    if (tid == 0) {
        T.n_dad_sites = ds; T.n_mom_sites = ms; T.n_dad_reads = dr; T.n_mom_reads = mr;
        T.has_record = has_rec > 0;
        A.tally[d] = T;
        if (A.ev_need) {                        // list lengths of unfz_evidence_lists (site entries keep duplicates)
            const int64_t n = A.n_dnms;
            A.ev_need[d] = dr; A.ev_need[n + d] = mr; A.ev_need[2 * n + d] = esd; A.ev_need[3 * n + d] = esm;
        }
    }
    CH_MARK(9);
}

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
        UNFZ_CHECK(ctx, cudaFuncSetAttribute(chain_bfs_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        ctx->chain_carveout_set = true;
    }
    chain_setup_kernel<<<n_dnms, CH_THREADS, 0, (cudaStream_t)stream>>>(A);
    UNFZ_LAUNCH_CHECK(ctx);
    chain_bfs_kernel<<<n_dnms, CH_THREADS, 0, (cudaStream_t)stream>>>(A);
    UNFZ_LAUNCH_CHECK(ctx);
    chain_evidence_kernel<<<n_dnms, CH_THREADS, 0, (cudaStream_t)stream>>>(A);
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
