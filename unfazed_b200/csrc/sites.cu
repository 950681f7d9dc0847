// Informative-site path: window membership, per-pair trio classification, stable compaction.
//
// Data movement is the whole cost here (45 algorithmic bytes per (DNM x site) pair, no reuse), so
// the classifier is a flat, persistent, grid-strided stream over pairs: consecutive threads take
// consecutive pairs, which map to consecutive rows of the SoA site columns inside one DNM window
// -> every column load of a warp is one contiguous 32..128 B span.  Allele balance is IEEE double
// division exactly like the reference's np.int32 / float (informative_site_finder.py:69).
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
struct ClsParams {
    double ab[3][2];      // indexed by range id: 0 homref, 1 het, 2 homalt
    double min_gq;
    int32_t min_depth;
};

// is_high_quality_site, informative_site_finder.py:46-73
__device__ __forceinline__ bool high_quality(const ClsParams& P, int gt, float gq, int32_t rd, int32_t ad) {
    int r;
    if (gt == 0) r = 0; else if (gt == 3) r = 2; else if (gt == 1) r = 1; else return false;
    if ((double)gq < P.min_gq) return false;
    const int32_t tot = (int32_t)((uint32_t)rd + (uint32_t)ad);    // numpy int32 add wraps
    if (tot < P.min_depth) return false;
    const double ab = (double)ad / (double)tot;                     // IEEE division, NaN/inf as numpy
    return P.ab[r][0] <= ab && ab <= P.ab[r][1];
}

// get_kid_allele, informative_site_finder.py:76-134.  returns 0 none, 1 ref_parent, 2 alt_parent
__device__ __forceinline__ int kid_allele(const ClsParams& P, int mode, int gk, const int32_t rd[3], const int32_t ad[3]) {
    const int32_t tk = (int32_t)((uint32_t)rd[0] + (uint32_t)ad[0]);
    if (mode == UNFZ_MODE_CNV_DEL && tk > 4) {
        if (gk == 3) return 1;
        if (gk == 0) return 2;
        return 0;
    }
    if (mode == UNFZ_MODE_CNV_DUP && rd[0] > 2 && ad[0] > 2 && tk > P.min_depth) {
        if (gk != 1) return 0;
        const double k = (double)ad[0] / (double)tk;
        const double d = (double)ad[1] / (double)(int32_t)((uint32_t)rd[1] + (uint32_t)ad[1]);
        const double m = (double)ad[2] / (double)(int32_t)((uint32_t)rd[2] + (uint32_t)ad[2]);
        const double dm = d + m;
        if ((dm < 1.0 && k > 0.5) || (dm > 1.0 && k < 0.5)) return 0;
        if (k >= 0.67) return 2;
        if (k <= 0.33) return 1;
        return 0;
    }
    return 0;
}

// find :252-339 == add_good_candidate_variant :457-542 for one (DNM, row)
__device__ __forceinline__ uint8_t classify_pair(const ClsParams& P, int mode, int32_t pos, int32_t excl_lo,
                                                 int32_t excl_hi, uint8_t flag, const uint8_t gt[3],
                                                 const float gq[3], const int32_t rd[3], const int32_t ad[3]) {
    if (!(flag & 1)) return 0;                                   // prefilter :239-244
    if (pos >= excl_lo && pos < excl_hi) return 0;                // small-event rule :253-256
    const int gk = gt[0], gd = gt[1], gm = gt[2];
    const bool parents = high_quality(P, gd, gq[1], rd[1], ad[1]) && high_quality(P, gm, gq[2], rd[2], ad[2]);
    uint8_t code = 0;
    if (gk == 1 && parents) code |= UNFZ_CLS_HET;
    int ka = 0;
    if (mode != UNFZ_MODE_READ) {
        ka = kid_allele(P, mode, gk, rd, ad);
        if (!ka) return code;
    } else if (gk != 1 || !high_quality(P, gk, gq[0], rd[0], ad[0])) {
        return code;
    }
    if (!parents) return code;
    bool alt_is_dad;
    if ((gd == 1 || gd == 3) && gm == 0) alt_is_dad = true;
    else if ((gm == 1 || gm == 3) && gd == 0) alt_is_dad = false;
    else if (gm == 1 && gd == 3) alt_is_dad = true;
    else if (gd == 1 && gm == 3) alt_is_dad = false;
    else return code;
    if (gk == 3 || gk == 0) {                                     // hemizygous-kid uniqueness :322-337
        const bool any_het = (gd == 1) || (gm == 1);
        const bool any_hom = (gd == 3) || (gm == 3) || (gd == 0) || (gm == 0);
        if (any_het && any_hom) {
            if (((gd == 3 || gd == 0) && gk == gd) || ((gm == 3 || gm == 0) && gk == gm)) return code;
        }
    }
    code |= UNFZ_CLS_CAND;
    if (alt_is_dad) code |= UNFZ_CLS_ALT_IS_DAD;
    if (ka == 2) code |= UNFZ_CLS_KID_ALT;
    return code;
}

constexpr int CLS_THREADS = 256;
constexpr int CLS_PER_THREAD = 4;
constexpr int CLS_TILE = CLS_THREADS * CLS_PER_THREAD;
constexpr int CLS_SMEM_SEGS = 1024;

__global__ void __launch_bounds__(CLS_THREADS)
classify_kernel(UnfzSiteCols sites, const UnfzSegIn* __restrict__ segs, const int32_t* __restrict__ seg_row_lo,
                const int64_t* __restrict__ seg_pair_off, int32_t n_segs, int64_t n_pairs, ClsParams P,
                uint8_t* __restrict__ out) {
    __shared__ int64_t s_off[CLS_SMEM_SEGS + 1];
    __shared__ int32_t s_seg0, s_nseg;
    const int64_t n_tiles = (n_pairs + CLS_TILE - 1) / CLS_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t p0 = tile * CLS_TILE;
        const int64_t p1 = min(p0 + (int64_t)CLS_TILE, n_pairs);
        if (threadIdx.x == 0) {
            // segment holding pair p0: last s with off[s] <= p0; segment holding pair p1-1
            const int64_t a = upper_bound_dev(seg_pair_off, 0, (int64_t)n_segs + 1, p0) - 1;
            const int64_t b = upper_bound_dev(seg_pair_off, a, (int64_t)n_segs + 1, p1 - 1) - 1;
            s_seg0 = (int32_t)a;
            s_nseg = (int32_t)(b - a + 1);
        }
        __syncthreads();
        const int seg0 = s_seg0, nseg = s_nseg;
        const bool staged = nseg <= CLS_SMEM_SEGS;
        if (staged)
            for (int i = threadIdx.x; i <= nseg; i += CLS_THREADS) s_off[i] = seg_pair_off[seg0 + i];
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CLS_PER_THREAD; ++j) {
            const int64_t p = p0 + (int64_t)j * CLS_THREADS + threadIdx.x;
            if (p >= p1) continue;
            int s;
            if (staged) {
                int lo = 0, hi = nseg;                   // last i with s_off[i] <= p
                while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (s_off[mid] <= p) lo = mid; else hi = mid; }
                s = seg0 + lo;
            } else {
                s = (int)(upper_bound_dev(seg_pair_off, (int64_t)seg0, (int64_t)n_segs + 1, p) - 1);
            }
            const UnfzSegIn sg = segs[s];
            const int64_t within = p - (staged ? s_off[s - seg0] : seg_pair_off[s]);
            const int64_t row = (int64_t)seg_row_lo[s] + (sg.mult == 1 ? within : within / sg.mult);
            uint8_t gt[3]; float gq[3]; int32_t rd[3], ad[3];
            const int32_t pos = __ldg(sites.pos + row);
            const uint8_t flag = __ldg(sites.flag + row);
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                gt[m] = __ldg(sites.gt[m] + row);
                gq[m] = __ldg(sites.gq[m] + row);
                rd[m] = __ldg(sites.rd[m] + row);
                ad[m] = __ldg(sites.ad[m] + row);
            }
            out[p] = classify_pair(P, sg.mode, pos, sg.excl_lo, sg.excl_hi, flag, gt, gq, rd, ad);
        }
        __syncthreads();
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
               int32_t* __restrict__ cnv_mom, uint8_t* __restrict__ row_mark) {
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

}  // namespace

extern "C" int unfz_window_search(UnfzCtx* ctx, const UnfzSiteCols* sites, const UnfzSegIn* segs, int32_t n_segs,
                                  int32_t* seg_row_lo, int64_t* seg_count, void* stream) {
    if (n_segs <= 0) return 0;
    window_search_kernel<<<(n_segs + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*sites, segs, n_segs, seg_row_lo, seg_count);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int unfz_classify_sites(UnfzCtx* ctx, const UnfzSiteCols* sites, const UnfzSegIn* segs,
                                   const int32_t* seg_row_lo, const int64_t* seg_pair_off, int32_t n_segs,
                                   int64_t n_pairs, const UnfzParams* hp, uint8_t* out_class, void* stream) {
    if (n_pairs <= 0) return 0;
    ClsParams P;
    P.ab[0][0] = hp->ab_homref[0]; P.ab[0][1] = hp->ab_homref[1];
    P.ab[1][0] = hp->ab_het[0];    P.ab[1][1] = hp->ab_het[1];
    P.ab[2][0] = hp->ab_homalt[0]; P.ab[2][1] = hp->ab_homalt[1];
    P.min_gq = hp->min_gt_qual;
    P.min_depth = hp->min_depth;
    const int64_t n_tiles = (n_pairs + CLS_TILE - 1) / CLS_TILE;
    // persistent grid: a multiple of the SM count, 8 resident CTAs of 256 threads per SM
    int64_t grid = (int64_t)ctx->sm_count * 8;
    if (grid > n_tiles) grid = n_tiles;
    classify_kernel<<<(unsigned)grid, CLS_THREADS, 0, (cudaStream_t)stream>>>(
        *sites, segs, seg_row_lo, seg_pair_off, n_segs, n_pairs, P, out_class);
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
        dnms, n_dnms, segs, seg_row_lo, seg_pair_off, cls, het_list, n_het, cand_list, n_cand, cnv_dad, cnv_mom, row_mark);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}
