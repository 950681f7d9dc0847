// Informative-site path: window membership, per-pair trio classification, stable compaction.
//
// Data movement is the whole cost here (45 algorithmic bytes per (DNM x site) pair, no reuse), so
// the classifier is a flat, persistent, grid-strided stream over pairs: consecutive threads take
// consecutive pairs, which map to consecutive rows of the SoA site columns inside one DNM window
// -> every column load of a warp is one contiguous 32..128 B span.  Allele balance is IEEE double
// division exactly like the reference's np.int32 / float (informative_site_finder.py:69).
#include <math.h>

#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// K0: window search
// ------------------------------------------------------------------------------------------------
__global__ void window_search_kernel(UnfzSiteCols sites, const UnfzSegIn* __restrict__ segs, int32_t n_segs,
                                     int32_t* __restrict__ seg_row_lo, int64_t* __restrict__ seg_count) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_segs) return;
    const UnfzSegIn sg = segs[s];
    int64_t lo = 0, cnt = 0;
    if (sg.sblk >= 0 && sg.hi_pos >= sg.lo_pos) {
        const int64_t a = sites.blk_off[sg.sblk], b = sites.blk_off[sg.sblk + 1];
        lo = lower_bound_dev(sites.pos, a, b, sg.lo_pos);
        const int64_t hi = upper_bound_dev(sites.pos, lo, b, sg.hi_pos);
        cnt = (hi - lo) * (int64_t)sg.mult;
    }
    seg_row_lo[s] = (int32_t)lo;
    seg_count[s] = cnt;
}

// ------------------------------------------------------------------------------------------------
// K1: classification
// ------------------------------------------------------------------------------------------------
// Allele balance without a division.  For a member with genotype g and total depth tot the reference asks
// min_ab <= float(ad / float(tot)) <= max_ab (informative_site_finder.py:69-71).  The quotient is monotone in ad, so
// for every (genotype class, tot) the passing ad form one interval [min_ad, max_ad]; the host works the intervals
// out ONCE with the very IEEE double division the reference performs (unfz_classify_sites) for every tot below
// CLS_TAB_TOT, folds "genotype is known" and "tot >= min_depth" into them (empty interval) and hands the table to the
// kernel (16 KB in device memory, copied into shared memory by every CTA with coalesced 16-byte loads): one LDS + two integer compares per member instead of a reciprocal, a
// multiply, six float compares and a guard band.  Depths of CLS_TAB_TOT and more, and negative counts, take the exact
// fp64 division (ab_exact) as before.
constexpr int CLS_TAB_TOT = 1024;
constexpr int CLS_TAB_ROWS = 4;           // homref, het, homalt, unknown/invalid (always empty)

struct ClsParams {
    const uint32_t* tab;                  // device, [CLS_TAB_ROWS][CLS_TAB_TOT]: min_ad | max_ad << 16; empty: min 1, max 0
    double lo[4], hi[4];                  // fp64 thresholds per table row for the exact path (row 3: NaN, never true)
    float min_gq_f;       // smallest float >= min_gq: for a float gq, (double)gq < min_gq <=> gq < min_gq_f
    int32_t min_depth;
};
struct ClsHost {                          // what the context keeps per set of thresholds
    ClsParams P;
    uint32_t tab[CLS_TAB_ROWS * CLS_TAB_TOT];
};

// exact: min_ab <= float(ad / float(rd+ad)) <= max_ab, informative_site_finder.py:69-71
__device__ __noinline__ bool ab_exact(double lo, double hi, int32_t ad, int32_t tot) {
    const double ab = (double)ad / (double)tot;                     // IEEE division, NaN/inf as numpy
    return lo <= ab && ab <= hi;
}

// DUP branch of get_kid_allele :89-130 (exact fp64); returns 0 none, 1 ref_parent, 2 alt_parent
__device__ __noinline__ int kid_allele_dup(int32_t rd0, int32_t ad0, int32_t rd1, int32_t ad1, int32_t rd2, int32_t ad2) {
    const double k = (double)ad0 / (double)(int32_t)((uint32_t)rd0 + (uint32_t)ad0);
    const double d = (double)ad1 / (double)(int32_t)((uint32_t)rd1 + (uint32_t)ad1);
    const double m = (double)ad2 / (double)(int32_t)((uint32_t)rd2 + (uint32_t)ad2);
    const double dm = d + m;
    if ((dm < 1.0 && k > 0.5) || (dm > 1.0 && k < 0.5)) return 0;
    if (k >= 0.67) return 2;
    if (k <= 0.33) return 1;
    return 0;
}

// One (DNM, row) pair: is_high_quality_site :46-73, get_kid_allele :76-134 and the classifier
// find :252-339 == add_good_candidate_variant :457-542.
// Written branch-free (a warp covers 32 different rows, so every early exit of the reference would
// be taken by some lane anyway); all predicates are pure, so the evaluation order is irrelevant.
// Only two rare events branch: an allele balance within 1e-6 of a threshold (exact fp64 division)
// and the DUP allele-balance rule.
__device__ __forceinline__ uint8_t classify_pair(const uint32_t* __restrict__ tab, const double* __restrict__ lohi,
                                                 const uint8_t* __restrict__ gt_lut,
                                                 float min_gq_f, int32_t min_depth, int mode,
                                                 int32_t pos, int32_t excl_lo, int32_t excl_hi, uint32_t meta,
                                                 const float gq[3], const int32_t rd[3], const int32_t ad[3]) {
    const bool base = (meta & 1u) && !(pos >= excl_lo && pos < excl_hi);  // prefilter :239-244, small event :253-256
    const int gk = (meta >> 8) & 0xff, gd = (meta >> 16) & 0xff, gm = meta >> 24;
    const bool read_mode = mode == UNFZ_MODE_READ;
    bool hq[3], need[3];
    int32_t tot[3];
    int trow[3];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const int g = m == 0 ? gk : (m == 1 ? gd : gm);
        // table row of the cyvcf2 gt code: 0 -> 0, 1 -> 1, 3 -> 2, 2 (unknown) and anything else -> 3 (empty)
        const int row = g > 3 ? 3 : (int)((0x2310u >> (g << 2)) & 3u);
        trow[m] = row;
        tot[m] = (int32_t)((uint32_t)rd[m] + (uint32_t)ad[m]);             // numpy int32 add wraps
        const bool gq_ok = !(gq[m] < min_gq_f);
        // inside the table: 0 <= tot < CLS_TAB_TOT and 0 <= ad <= tot, as two unsigned compares
        const bool in_tab = ((uint32_t)tot[m] < (uint32_t)CLS_TAB_TOT) & ((uint32_t)ad[m] <= (uint32_t)tot[m]);
        const uint32_t e = tab[row * CLS_TAB_TOT + (in_tab ? tot[m] : 0)];
        hq[m] = gq_ok & in_tab & ((uint32_t)ad[m] >= (e & 0xffffu)) & ((uint32_t)ad[m] <= (e >> 16));
        need[m] = gq_ok & !in_tab & (row != 3) & (tot[m] >= min_depth);
    }
    if (base && (need[0] | need[1] | need[2])) {                           // rare: depth >= CLS_TAB_TOT or negative counts
#pragma unroll
        for (int m = 0; m < 3; ++m)
            if (need[m]) hq[m] = ab_exact(lohi[trow[m]], lohi[4 + trow[m]], ad[m], tot[m]);
    }
    const bool parents = hq[1] & hq[2];
    int ka = 0;                                                            // get_kid_allele
    if (!read_mode) {
        if (mode == UNFZ_MODE_CNV_DEL && tot[0] > 4) ka = gk == 3 ? 1 : (gk == 0 ? 2 : 0);
        else if (mode == UNFZ_MODE_CNV_DUP && base && parents && rd[0] > 2 && ad[0] > 2 && tot[0] > min_depth && gk == 1)
            ka = kid_allele_dup(rd[0], ad[0], rd[1], ad[1], rd[2], ad[2]);
    }
    const bool het = base & (gk == 1) & parents;
    // parental pattern :307-320 and hemizygous-kid uniqueness :322-337 depend on the three genotype
    // codes only: 64-entry table (bit0 informative pattern, bit1 alt_parent is dad, bit2 kid shares the
    // homozygous parent's allele -> reject)
    const uint32_t lut = gt_lut[((gk & 3) << 4) | ((gd & 3) << 2) | (gm & 3)];
    const bool alt_is_dad = (lut & 2u) != 0;
    const bool kid_ok = read_mode ? ((gk == 1) & hq[0]) : (ka != 0);
    const bool cand = base & parents & kid_ok & ((lut & 1u) != 0) & !(lut & 4u);
    uint8_t code = 0;
    if (het) code |= UNFZ_CLS_HET;
    if (cand) code |= UNFZ_CLS_CAND;
    if (cand & alt_is_dad) code |= UNFZ_CLS_ALT_IS_DAD;
    if (cand & (ka == 2)) code |= UNFZ_CLS_KID_ALT;
    return code;
}

// 4 CTAs per SM (64 registers; the software pipeline spills 64 B, still the fastest: 3 CTAs x 80 registers
// 0.461 ms, 4 x 64 0.454 ms, 5 x 48 0.487 ms on 2^25 pairs)
#ifndef CLS_MINB
#define CLS_MINB 4
#endif
constexpr int CLS_THREADS = 256;
constexpr int CLS_PER_THREAD = 8;
constexpr int CLS_TILE = CLS_THREADS * CLS_PER_THREAD;
constexpr int CLS_SMEM_SEGS = 512;

// Flat stream over pairs.  Every CTA owns a contiguous range of 1024-pair tiles, so the first
// segment of a tile is carried over from the previous tile (one global binary search per CTA);
// the descriptors of the segments a tile touches are staged in shared memory and a pair -> segment
// map is built with one block scan per tile (segment starts are counted per pair slot, the
// inclusive prefix is the index of the last segment starting at or before the pair).
__global__ void __launch_bounds__(CLS_THREADS, CLS_MINB)
classify_kernel(UnfzSiteCols sites, const UnfzSegIn* __restrict__ segs, const int32_t* __restrict__ seg_row_lo,
                const int64_t* __restrict__ seg_pair_off, int32_t n_segs, int64_t n_pairs_cap, ClsParams P,
                uint8_t* __restrict__ out, const int32_t* __restrict__ guard) {
    UNFZ_GUARD(guard);
    // the launch is sized for n_pairs_cap (the caller's buffer); the batch's own total is on the device
    const int64_t n_pairs = min(n_pairs_cap, seg_pair_off[n_segs]);
    __shared__ int32_t s_off[CLS_SMEM_SEGS + 1];     // pair offset of the segment relative to the tile
    __shared__ int4 s_seg[CLS_SMEM_SEGS];            // row_lo, mult | mode << 24, excl_lo, excl_hi
    __shared__ int32_t s_map[CLS_TILE];
    __shared__ int32_t s_warp[CLS_THREADS / 32];
    __shared__ __align__(16) uint32_t s_tab[CLS_TAB_ROWS * CLS_TAB_TOT];
    __shared__ double s_lohi[8];
    __shared__ uint8_t s_lut[64];
    __shared__ int32_t s_seg0;
    const int64_t n_tiles = (n_pairs + CLS_TILE - 1) / CLS_TILE;
    const int64_t tpc = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * tpc, t1 = min(t0 + tpc, n_tiles);
    if (t0 >= t1) return;
    for (int i = threadIdx.x; i < CLS_TAB_ROWS * CLS_TAB_TOT / 4; i += CLS_THREADS)
        reinterpret_cast<uint4*>(s_tab)[i] = __ldg(reinterpret_cast<const uint4*>(P.tab) + i);
    if (threadIdx.x < 8) s_lohi[threadIdx.x] = threadIdx.x < 4 ? P.lo[threadIdx.x] : P.hi[threadIdx.x - 4];
    if (threadIdx.x < 64) {
        const int gk = threadIdx.x >> 4, gd = (threadIdx.x >> 2) & 3, gm = threadIdx.x & 3;
        const bool p1 = ((gd == 1) | (gd == 3)) & (gm == 0);
        const bool p2 = ((gm == 1) | (gm == 3)) & (gd == 0);
        const bool p3 = (gm == 1) & (gd == 3);
        const bool p4 = (gd == 1) & (gm == 3);
        const bool kid_hom = (gk == 3) | (gk == 0);
        const bool any_het = (gd == 1) | (gm == 1);
        const bool any_hom = (gd == 3) | (gm == 3) | (gd == 0) | (gm == 0);
        const bool shared = (((gd == 3) | (gd == 0)) & (gk == gd)) | (((gm == 3) | (gm == 0)) & (gk == gm));
        s_lut[threadIdx.x] = (uint8_t)(((p1 | p2 | p3 | p4) ? 1 : 0) | ((p1 | p3) ? 2 : 0) | ((kid_hom & any_het & any_hom & shared) ? 4 : 0));
    }
    if (threadIdx.x == 0)
        s_seg0 = (int32_t)(upper_bound_dev(seg_pair_off, 0, (int64_t)n_segs + 1, t0 * CLS_TILE) - 1);
    __syncthreads();
    int seg0 = s_seg0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4* __restrict__ rec4 = reinterpret_cast<const float4*>(sites.rec);
    const int2* __restrict__ dep2 = reinterpret_cast<const int2*>(sites.dep);
    for (int64_t tile = t0; tile < t1; ++tile) {
        const int64_t p0 = tile * CLS_TILE;
        const int npair = (int)min((int64_t)CLS_TILE, n_pairs - p0);
#pragma unroll
        for (int j = 0; j < CLS_PER_THREAD; ++j) s_map[j * CLS_THREADS + threadIdx.x] = 0;
        // stage the descriptors of the segments this tile touches, starting at the carried one, in
        // rounds of 128 (typically one round)
        int staged_n = 0;
        for (int round = 0; round < CLS_SMEM_SEGS / 128; ++round) {
            const int i = round * 128 + threadIdx.x;
            if (threadIdx.x < 128) {
                const int sidx = seg0 + i;
                int64_t rel = (int64_t)1 << 30;
                if (sidx <= n_segs) rel = min(max(seg_pair_off[sidx] - p0, -((int64_t)1 << 30)), (int64_t)1 << 30);
                s_off[i] = (int32_t)rel;
                if (sidx < n_segs && rel < npair) {
                    const UnfzSegIn sg = segs[sidx];
                    s_seg[i] = make_int4(seg_row_lo[sidx], (sg.mult & 0xffffff) | (sg.mode << 24), sg.excl_lo, sg.excl_hi);
                }
            }
            if (threadIdx.x == 0 && round == CLS_SMEM_SEGS / 128 - 1) s_off[CLS_SMEM_SEGS] = 1 << 30;
            __syncthreads();
            staged_n = round * 128 + 128;
            if (s_off[staged_n - 1] >= npair) break;              // the next segment starts beyond the tile
        }
        const bool staged = !(s_off[staged_n - 1] < npair);       // false: more than 512 segments in this tile
        int last_started = 0;                                     // largest staged i with s_off[i] < npair
        if (staged) {
            // count segment starts per pair slot (segments i >= 1 that start inside the tile)
            for (int i = 1 + threadIdx.x; i < staged_n; i += CLS_THREADS) {
                const int q = s_off[i];
                if (q < npair) atomicAdd(&s_map[max(q, 0)], 1);
            }
            __syncthreads();
            // inclusive scan over the 1024 slots: thread t owns slots 4t..4t+3
            int c[CLS_PER_THREAD];
            int sum = 0;
#pragma unroll
            for (int j = 0; j < CLS_PER_THREAD; ++j) { sum += s_map[threadIdx.x * CLS_PER_THREAD + j]; c[j] = sum; }
            int x = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) s_warp[warp] = x;
            __syncthreads();
            int basew = 0, total = 0;
#pragma unroll
            for (int w = 0; w < CLS_THREADS / 32; ++w) { if (w < warp) basew += s_warp[w]; total += s_warp[w]; }
            const int excl = basew + x - sum;
#pragma unroll
            for (int j = 0; j < CLS_PER_THREAD; ++j) s_map[threadIdx.x * CLS_PER_THREAD + j] = excl + c[j];
            last_started = total;
            __syncthreads();
        }
        // Software pipeline over the thread's CLS_PER_THREAD pairs: the five row loads of pair j+1 are issued
        // before pair j is classified.  No branch around a pair -- the tail of the last tile is clamped and
        // its store predicated -- so the unrolled loop is straight-line code.
        struct PairIn { uint32_t meta; float4 rc; int2 d0, d1, d2; int32_t mode, exlo, exhi; };
        auto fetch = [&](int j) -> PairIn {
            const int q_raw = j * CLS_THREADS + threadIdx.x;      // pair index inside the tile
            const int q = q_raw < npair ? q_raw : npair - 1;
            int32_t row_lo, mult, within;
            PairIn in;
            if (staged) {
                const int lo = s_map[q];
                const int4 sg = s_seg[lo];
                row_lo = sg.x; mult = sg.y & 0xffffff; in.mode = (uint32_t)sg.y >> 24; in.exlo = sg.z; in.exhi = sg.w;
                within = q - s_off[lo];
            } else {
                const int sidx = (int)(upper_bound_dev(seg_pair_off, (int64_t)seg0, (int64_t)n_segs + 1, p0 + q) - 1);
                const UnfzSegIn sg = segs[sidx];
                row_lo = seg_row_lo[sidx]; mult = sg.mult; in.exlo = sg.excl_lo; in.exhi = sg.excl_hi; in.mode = sg.mode;
                within = (int32_t)(p0 + q - seg_pair_off[sidx]);
            }
            const int64_t row = (int64_t)row_lo + (mult == 1 ? within : within / mult);
            in.meta = __ldg(sites.meta + row);
            in.rc = __ldg(rec4 + row);
            in.d0 = __ldg(dep2 + row * 3); in.d1 = __ldg(dep2 + row * 3 + 1); in.d2 = __ldg(dep2 + row * 3 + 2);
            return in;
        };
        PairIn cur = fetch(0);
#pragma unroll
        for (int j = 0; j < CLS_PER_THREAD; ++j) {
            PairIn nxt = cur;
            if (j + 1 < CLS_PER_THREAD) nxt = fetch(j + 1);
            const float gq[3] = {cur.rc.y, cur.rc.z, cur.rc.w};
            const int32_t rd[3] = {cur.d0.x, cur.d1.x, cur.d2.x}, ad[3] = {cur.d0.y, cur.d1.y, cur.d2.y};
            const uint8_t code = classify_pair(s_tab, s_lohi, s_lut, P.min_gq_f, P.min_depth, cur.mode, __float_as_int(cur.rc.x), cur.exlo,
                                               cur.exhi, cur.meta, gq, rd, ad);
            const int q_raw = j * CLS_THREADS + threadIdx.x;
            if (q_raw < npair) out[p0 + q_raw] = code;
            cur = nxt;
        }
        __syncthreads();
        if (staged) seg0 += last_started;                         // the last segment may continue into the next tile
        else seg0 = (int)(upper_bound_dev(seg_pair_off, (int64_t)seg0, (int64_t)n_segs + 1, p0 + npair - 1) - 1);
    }
}

// ------------------------------------------------------------------------------------------------
// compaction: one warp per DNM entry, segments in order, ballot-ordered appends
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
compact_kernel(const UnfzDnm* __restrict__ dnms, int32_t n_dnms, const UnfzSegIn* __restrict__ segs,
               const int32_t* __restrict__ seg_row_lo, const int64_t* __restrict__ seg_pair_off,
               const uint8_t* __restrict__ cls, int32_t* __restrict__ het_list, int32_t* __restrict__ n_het,
               uint32_t* __restrict__ cand_list, int32_t* __restrict__ n_cand, int32_t* __restrict__ cnv_dad,
               int32_t* __restrict__ cnv_mom, uint8_t* __restrict__ row_mark, const int32_t* __restrict__ guard) {
    UNFZ_GUARD(guard);
    const int lane = threadIdx.x & 31;
    const int d = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (d >= n_dnms) return;
    const UnfzDnm dn = dnms[d];
    int nh = 0, nc = 0, vd = 0, vm = 0;
    if (dn.seg_hi > dn.seg_lo) {
        const int64_t base = seg_pair_off[dn.seg_lo];
        for (int s = dn.seg_lo; s < dn.seg_hi; ++s) {
            const UnfzSegIn sg = segs[s];
            const int64_t p0 = seg_pair_off[s], p1 = seg_pair_off[s + 1];
            const int64_t row0 = seg_row_lo[s];
            for (int64_t p = p0 + lane; p - lane < p1; p += 32) {
                uint8_t c = 0;
                int32_t row = 0;
                if (p < p1) {
                    c = cls[p];
                    row = (int32_t)(row0 + (p - p0) / sg.mult);
                }
                const unsigned mh = __ballot_sync(0xffffffffu, c & UNFZ_CLS_HET);
                const unsigned mc = __ballot_sync(0xffffffffu, c & UNFZ_CLS_CAND);
                const unsigned below = (1u << lane) - 1u;
                if (c & UNFZ_CLS_HET) het_list[base + nh + __popc(mh & below)] = row;
                if (c & UNFZ_CLS_CAND) {
                    uint32_t v = (uint32_t)row;
                    if (c & UNFZ_CLS_ALT_IS_DAD) v |= 0x80000000u;
                    if (c & UNFZ_CLS_KID_ALT) v |= 0x40000000u;
                    cand_list[base + nc + __popc(mc & below)] = v;
                }
                if (sg.mode == UNFZ_MODE_READ) {
                    if (c & (UNFZ_CLS_HET | UNFZ_CLS_CAND)) row_mark[row] = 1;
                } else {
                    // phase_by_snvs: the site votes for site[site["kid_allele"]]
                    const bool dad = ((c & UNFZ_CLS_ALT_IS_DAD) != 0) == ((c & UNFZ_CLS_KID_ALT) != 0);
                    const unsigned md = __ballot_sync(0xffffffffu, (c & UNFZ_CLS_CAND) && dad);
                    vd += __popc(md);
                    vm += __popc(mc) - __popc(md);
                }
                nh += __popc(mh);
                nc += __popc(mc);
            }
        }
    }
    if (lane == 0) {
        n_het[d] = nh;
        n_cand[d] = nc;
        cnv_dad[d] = vd;
        cnv_mom[d] = vm;
    }
}

// ------------------------------------------------------------------------------------------------
// upload completion: SoA genotype columns -> the classifier's 44-byte row layout
// ------------------------------------------------------------------------------------------------
__global__ void pack_site_rows_kernel(int64_t n, const int32_t* __restrict__ pos, const uint8_t* __restrict__ flag,
                                      const uint8_t* __restrict__ gt, const float* __restrict__ gq,
                                      const int32_t* __restrict__ rd, const int32_t* __restrict__ ad,
                                      uint32_t* __restrict__ meta, float4* __restrict__ rec, int2* __restrict__ dep) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    meta[i] = (uint32_t)flag[i] | ((uint32_t)gt[i] << 8) | ((uint32_t)gt[n + i] << 16) | ((uint32_t)gt[2 * n + i] << 24);
    rec[i] = make_float4(__int_as_float(pos[i]), gq[i], gq[n + i], gq[2 * n + i]);
    dep[3 * i] = make_int2(rd[i], ad[i]);
    dep[3 * i + 1] = make_int2(rd[n + i], ad[n + i]);
    dep[3 * i + 2] = make_int2(rd[2 * n + i], ad[2 * n + i]);
}

}  // namespace

extern "C" int unfz_pack_site_rows(UnfzCtx* ctx, int64_t n_rows, const int32_t* pos, const uint8_t* flag, const uint8_t* gt,
                                   const float* gq, const int32_t* rd, const int32_t* ad, uint32_t* meta, float* rec,
                                   int32_t* dep, void* stream) {
    if (n_rows <= 0) return 0;
    pack_site_rows_kernel<<<(unsigned)((n_rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n_rows, pos, flag, gt, gq, rd, ad, meta, reinterpret_cast<float4*>(rec), reinterpret_cast<int2*>(dep));
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int unfz_window_search(UnfzCtx* ctx, const UnfzSiteCols* sites, const UnfzSegIn* segs, int32_t n_segs,
                                  int32_t* seg_row_lo, int64_t* seg_count, void* stream) {
    if (n_segs <= 0) return 0;
    window_search_kernel<<<(n_segs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*sites, segs, n_segs, seg_row_lo, seg_count);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

// The classifier's parameter block for the thresholds of this call; rebuilt only when they change (the engine passes
// the same ones batch after batch).  Interval ends are found with the reference's own arithmetic -- IEEE double
// division of the two counts, inclusive compares -- starting next to lo*tot / hi*tot, so building all 3 x 1024
// intervals is a few thousand divisions.
static ClsParams* cls_params_for(UnfzCtx* ctx, const UnfzParams* hp) {
    const double key[8] = {hp->ab_homref[0], hp->ab_homref[1], hp->ab_het[0], hp->ab_het[1], hp->ab_homalt[0], hp->ab_homalt[1],
                           hp->min_gt_qual, (double)hp->min_depth};
    if (ctx->cls_params && memcmp(key, ctx->cls_key, sizeof(key)) == 0) return &static_cast<ClsHost*>(ctx->cls_params)->P;
    if (!ctx->cls_params) ctx->cls_params = malloc(sizeof(ClsHost));
    if (!ctx->cls_tab_dev && cudaMalloc(&ctx->cls_tab_dev, sizeof(uint32_t) * CLS_TAB_ROWS * CLS_TAB_TOT) != cudaSuccess) return nullptr;
    ClsHost& H = *static_cast<ClsHost*>(ctx->cls_params);
    ClsParams& P = H.P;
    const double ab[3][2] = {{hp->ab_homref[0], hp->ab_homref[1]}, {hp->ab_het[0], hp->ab_het[1]}, {hp->ab_homalt[0], hp->ab_homalt[1]}};
    for (int r = 0; r < CLS_TAB_ROWS; ++r) {
        P.lo[r] = r < 3 ? ab[r][0] : NAN;
        P.hi[r] = r < 3 ? ab[r][1] : NAN;
        for (int tot = 0; tot < CLS_TAB_TOT; ++tot) {
            uint32_t e = 1u;                                               // empty: min 1 > max 0
            if (r < 3 && tot >= 1 && tot >= hp->min_depth) {
                const double lo = ab[r][0], hi = ab[r][1], t = (double)tot;
                auto pass = [&](int a) { const double q = (double)a / t; return lo <= q && q <= hi; };
                // smallest passing ad: start two below the real-valued bound, walk up; largest: two above, walk down
                double g0 = floor(lo * t) - 2.0, g1 = ceil(hi * t) + 2.0;
                int a0 = !(g0 > 0.0) ? 0 : (g0 > t ? tot + 1 : (int)g0);
                int a1 = !(g1 < t) ? tot : (g1 < 0.0 ? -1 : (int)g1);
                while (a0 <= tot && !pass(a0)) ++a0;
                while (a1 >= 0 && !pass(a1)) --a1;
                if (a0 <= a1) e = (uint32_t)a0 | ((uint32_t)a1 << 16);
            }
            H.tab[r * CLS_TAB_TOT + tot] = e;
        }
    }
    P.tab = ctx->cls_tab_dev;
    P.min_gq_f = (float)hp->min_gt_qual;
    if ((double)P.min_gq_f < hp->min_gt_qual) P.min_gq_f = nextafterf(P.min_gq_f, INFINITY);
    P.min_depth = hp->min_depth;
    // synchronous on purpose: the table is read by launches on any stream from now on (thresholds change once per run)
    memset(ctx->cls_key, 0xff, sizeof(ctx->cls_key));
    if (cudaMemcpy(ctx->cls_tab_dev, H.tab, sizeof(H.tab), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    memcpy(ctx->cls_key, key, sizeof(key));
    return &P;
}

extern "C" int unfz_classify_prepare(UnfzCtx* ctx, const UnfzParams* hp) {
    if (!cls_params_for(ctx, hp)) { UNFZ_CHECK(ctx, cudaGetLastError()); return unfz_fail(ctx, -22, "classifier table upload failed"); }
    return 0;
}

extern "C" int unfz_classify_sites(UnfzCtx* ctx, const UnfzSiteCols* sites, const UnfzSegIn* segs,
                                   const int32_t* seg_row_lo, const int64_t* seg_pair_off, int32_t n_segs,
                                   int64_t n_pairs, const UnfzParams* hp, uint8_t* out_class, void* stream) {
    if (n_pairs <= 0) return 0;
    ClsParams* Pp = cls_params_for(ctx, hp);
    if (!Pp) { UNFZ_CHECK(ctx, cudaGetLastError()); return unfz_fail(ctx, -22, "classifier table upload failed (thresholds changed under stream capture?)"); }
    const ClsParams& P = *Pp;
    const int64_t n_tiles = (n_pairs + CLS_TILE - 1) / CLS_TILE;
    // persistent grid: a multiple of the SM count, 8 resident CTAs of 256 threads per SM
    int64_t grid = (int64_t)ctx->sm_count * 8;
    if (grid > n_tiles) grid = n_tiles;
    classify_kernel<<<(unsigned)grid, CLS_THREADS, 0, (cudaStream_t)stream>>>(
        *sites, segs, seg_row_lo, seg_pair_off, n_segs, n_pairs, P, out_class, ctx->guard);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int unfz_compact_sites(UnfzCtx* ctx, const UnfzDnm* dnms, int32_t n_dnms, const UnfzSegIn* segs,
                                  const int32_t* seg_row_lo, const int64_t* seg_pair_off, const uint8_t* cls,
                                  int32_t* het_list, int32_t* n_het, uint32_t* cand_list, int32_t* n_cand,
                                  int32_t* cnv_dad, int32_t* cnv_mom, uint8_t* row_mark, void* stream) {
    if (n_dnms <= 0) return 0;
    const int warps_per_block = 8;
    compact_kernel<<<(n_dnms + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0, (cudaStream_t)stream>>>(
        dnms, n_dnms, segs, seg_row_lo, seg_pair_off, cls, het_list, n_het, cand_list, n_cand, cnv_dad, cnv_mom, row_mark,
        ctx->guard);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}
