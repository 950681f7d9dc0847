// Read path, part 1: the streaming read scan (goodread + pair-independent filters + overlap counts)
// and the read-by-site allele lookup.
//
// read_scan is the bandwidth kernel of the read path: per read it must see the 32 B header, the
// CIGAR words and every quality byte.  Reads are stored in file order, so the quality bytes of a
// tile of consecutive reads are one contiguous span: a single elected thread moves that span into
// shared memory with TMA bulk copies (cp.async.bulk, completion on an mbarrier) while the other
// threads walk their CIGARs; then every thread counts the low-quality bases of its own read out of
// shared memory with 4-byte SIMD compares.  Several CTAs are resident per SM, so one CTA's bulk
// copy overlaps the others' compute.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------
// mbarrier / TMA bulk-copy wrappers (sm_90+; on sm_100a these lower to UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------
// K2: read scan
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;                 // reads per tile
constexpr int RS_QBUF = 48 * 1024;              // staged quality bytes per round
constexpr int RS_SPOS = 128;                    // site positions staged per tile

struct ScanParams {
    int32_t min_mapq;
    int32_t min_bq;      // clamped to [0,128]
    int32_t readlen;
};

__device__ __forceinline__ int count_low_quals(const uint8_t* __restrict__ s, int beg, int end, uint32_t minq) {
    // number of bytes b in s[beg,end) with (b & 0x7f) < minq; minq in [0,128].
    // Per aligned 32-bit word: t = (x | 0x80808080) - m4 never borrows across bytes (the OR also drops
    // the escape bit), and bit 7 of a byte of t is clear exactly when (x & 0x7f) < m.  The four flag
    // bytes (0x80 or 0) are summed with one DP4A; the first and last word are masked to the range.
    // No POPC (quarter-rate XU pipe, it was the top pipe in the first ncu capture) and no byte loops.
    if (end <= beg) return 0;
    const uint32_t m4 = minq * 0x01010101u;
    const int w0 = beg & ~3, w1 = (end + 3) & ~3;
    const uint32_t head = 0x80808080u << (8 * (beg & 3));                 // bytes >= beg inside the first word
    const uint32_t tail = 0x80808080u >> (8 * ((4 - (end & 3)) & 3));     // bytes <  end inside the last word
    unsigned acc = 0;
    // first word (masked), unmasked middle words, last word (masked): the masks stay out of the loop
    {
        const uint32_t x = *reinterpret_cast<const uint32_t*>(s + w0);
        uint32_t f = ~((x | 0x80808080u) - m4) & 0x80808080u & head;
        if (w1 - w0 == 4) f &= tail;
        acc = __dp4a(f, 0x01010101u, acc);
    }
    const int wl = w1 - 4;
#pragma unroll 8
    for (int i = w0 + 4; i < wl; i += 4) {
        const uint32_t x = *reinterpret_cast<const uint32_t*>(s + i);
        acc = __dp4a(~((x | 0x80808080u) - m4) & 0x80808080u, 0x01010101u, acc);
    }
    if (wl > w0) {
        const uint32_t x = *reinterpret_cast<const uint32_t*>(s + wl);
        acc = __dp4a(~((x | 0x80808080u) - m4) & 0x80808080u & tail, 0x01010101u, acc);
    }
    return (int)(acc >> 7);
}

__global__ void __launch_bounds__(RS_THREADS)
read_scan_kernel(UnfzReadCols reads, UnfzSiteCols sites, const int32_t* __restrict__ mark_prefix, ScanParams P,
                 UnfzReadSum* __restrict__ out, int32_t* __restrict__ row_lb, int32_t* __restrict__ blk_maxspan,
                 uint32_t* __restrict__ tile_tot) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* qbuf = smem;                                             // RS_QBUF + 16
    int32_t* spos = reinterpret_cast<int32_t*>(smem + RS_QBUF + 16);  // RS_SPOS
    __shared__ __align__(8) uint64_t bar;
    __shared__ int64_t s_qa, s_qb, s_row_base, s_row_end;
    __shared__ int32_t s_rb0;
    __shared__ int32_t s_maxspan;
    __shared__ int32_t s_wsum[RS_THREADS / 32];

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    uint32_t phase = 0;

    const int64_t n = reads.n_reads;
    const int64_t n_tiles = (n + RS_THREADS - 1) / RS_THREADS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r0 = tile * RS_THREADS;
        const int64_t r = r0 + threadIdx.x;
        const bool live = r < n;
        UnfzRead h;
        if (live) h = load_read(reads.hdr + r);
        const int64_t q0 = live ? read_qoff(h) : 0;
        const int L = live ? h.l_seq : 0;
        if (threadIdx.x == 0) {
            s_qa = q0;
            s_rb0 = (int32_t)(upper_bound_dev(reads.blk_off, 0, (int64_t)reads.n_blocks + 1, r0) - 1);
            s_maxspan = 0;
        }
        const int64_t last = min(r0 + RS_THREADS, n) - 1;
        if (r == last) s_qb = q0 + L;
        __syncthreads();
        const int64_t qa = s_qa, qb = s_qb;
        const int64_t ga = qa & ~(int64_t)15;
        // round 0 of the quality staging is issued now so it overlaps the CIGAR walk below
        int64_t chunk_lo = qb > qa ? ga : qb;
        if (threadIdx.x == 0 && qb > qa) {
            const uint32_t bytes = (uint32_t)min((int64_t)RS_QBUF, ((qb - chunk_lo) + 15) & ~(int64_t)15);
            fence_proxy_async();
            mbar_expect_tx(&bar, bytes);
            tma_bulk_g2s(qbuf, reads.qual + chunk_lo, bytes, &bar);
        }

        // ---- block of this read, site rows of the tile ------------------------------------
        int rb = s_rb0;
        if (live && r >= reads.blk_off[rb + 1])
            rb = (int)(upper_bound_dev(reads.blk_off, (int64_t)rb, (int64_t)reads.n_blocks + 1, r) - 1);
        const int sb = live ? reads.blk_sblk[rb] : -1;
        if (threadIdx.x == 0) {
            int64_t rowb = 0, rowe = 0;
            if (sb >= 0) {
                rowe = sites.blk_off[sb + 1];
                rowb = lower_bound_dev(sites.pos, sites.blk_off[sb], rowe, h.start);
            }
            s_row_base = rowb;
            s_row_end = rowe;
        }
        __syncthreads();
        const int64_t row_base = s_row_base, row_end = s_row_end;
        for (int i = threadIdx.x; i < RS_SPOS; i += RS_THREADS)
            spos[i] = (row_base + i < row_end) ? __ldg(sites.pos + row_base + i) : 0x7fffffff;

        // ---- header flags + CIGAR walk -----------------------------------------------------
        int32_t end = 0;
        uint32_t flags = 0;
        int none_cnt = 0, non_m = 0;
        if (live) {
            end = h.start;
            const uint32_t* cg = reads.cigar + h.cigar_off;
            for (int k = 0; k < h.n_cigar; ++k) {
                const uint32_t w = __ldg(cg + k);
                const uint32_t op = w & 15u, ln = w >> 4;
                if (op == 0 || op == 7 || op == 8 || op == 2 || op == 3) end += (int32_t)ln;
                if (op == 1 || op == 4) none_cnt += (int)ln;
                if (op != 0 && op != 7) ++non_m;
            }
            const uint32_t f = h.flag;
            const bool base_ok = !(f & (0x200u | 0x4u | 0x400u | 0x100u | 0x800u | 0x8u)) &&
                                 (int)h.mapq >= P.min_mapq && (h.aux & 1u);
            if (base_ok) flags |= UNFZ_RS_GOOD_DISC;
            if (none_cnt <= 5) flags |= UNFZ_RS_NONE_OK;
            if (non_m <= 5) flags |= UNFZ_RS_EXT_OK;
            long long ins = (long long)h.tlen - 2ll * P.readlen;
            if (ins < 0) ins = -ins;
            if ((double)ins <= reads.blk_cul[rb]) flags |= UNFZ_RS_INS_OK;
            if ((f & 1u) && !(f & 8u) && h.mate >= 0) flags |= UNFZ_RS_HAS_MATE;
            atomicMax(&s_maxspan, end - h.start);
        }
        __syncthreads();   // spos visible

        // ---- marked-site overlap count -------------------------------------------------------
        int32_t fmark = 0, cnt = 0;
        int64_t lbs = 0, lbe = 0;
        if (live && sb >= 0) {
            if (rb == s_rb0) {
                int lo = 0, hi = RS_SPOS;
                while (lo < hi) { int mid = (lo + hi) >> 1; if (spos[mid] < h.start) lo = mid + 1; else hi = mid; }
                int lo2 = lo; hi = RS_SPOS;
                while (lo2 < hi) { int mid = (lo2 + hi) >> 1; if (spos[mid] < end) lo2 = mid + 1; else hi = mid; }
                lbs = row_base + lo;
                lbe = row_base + lo2;
                if (lo2 == RS_SPOS) {     // ran off the staged window: finish in global memory
                    if (lo == RS_SPOS) lbs = lower_bound_dev(sites.pos, row_base + RS_SPOS - 1, row_end, h.start);
                    lbe = lower_bound_dev(sites.pos, lbs, row_end, end);
                }
            } else {
                const int64_t a = sites.blk_off[sb], b = sites.blk_off[sb + 1];
                lbs = lower_bound_dev(sites.pos, a, b, h.start);
                lbe = lower_bound_dev(sites.pos, lbs, b, end);
            }
            fmark = __ldg(mark_prefix + lbs);
            const int32_t c = __ldg(mark_prefix + lbe) - fmark;
            cnt = c > 0xffff ? 0xffff : c;
        }

        uint32_t hoff = 0;                       // offset inside the tile; see the pipelined kernel
        {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            int x = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) s_wsum[warp] = x;
            __syncthreads();
            int basew = 0, total = 0;
#pragma unroll
            for (int w = 0; w < RS_THREADS / 32; ++w) { if (w < warp) basew += s_wsum[w]; total += s_wsum[w]; }
            hoff = (uint32_t)(basew + x - cnt);
            if (threadIdx.x == 0) tile_tot[tile] = (uint32_t)total;
        }
        // ---- low-quality bases out of the staged span ---------------------------------------
        int low = 0;
        while (chunk_lo < qb) {
            mbar_wait(&bar, phase);
            phase ^= 1u;
            const int64_t chunk_hi = chunk_lo + RS_QBUF;
            if (live && L > 0) {
                const int64_t a = max(q0, chunk_lo), b = min(q0 + (int64_t)L, min(chunk_hi, qb));
                if (a < b) low += count_low_quals(qbuf, (int)(a - chunk_lo), (int)(b - chunk_lo), (uint32_t)P.min_bq);
            }
            chunk_lo = chunk_hi;
            __syncthreads();     // everyone is done with qbuf
            if (threadIdx.x == 0 && chunk_lo < qb) {
                const uint32_t bytes = (uint32_t)min((int64_t)RS_QBUF, ((qb - chunk_lo) + 15) & ~(int64_t)15);
                fence_proxy_async();
                mbar_expect_tx(&bar, bytes);
                tma_bulk_g2s(qbuf, reads.qual + chunk_lo, bytes, &bar);
            }
        }
        if (live) {
            // goodread(read): <= 10 low-quality bases and (Q3) <= 10 CIGAR operations
            if ((flags & UNFZ_RS_GOOD_DISC) && low <= 10 && h.n_cigar <= 10) flags |= UNFZ_RS_GOOD_CONC;
            UnfzReadSum o;
            o.end = end;
            o.fmark = fmark;
            o.flags = (uint16_t)flags;
            o.cnt = (uint16_t)cnt;
            o.hoff = hoff;
            *reinterpret_cast<int4*>(out + r) = *reinterpret_cast<const int4*>(&o);
            row_lb[r] = (int32_t)lbs;
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_maxspan > 0) {
            // a tile can straddle blocks; attribute the span to every block it touches (conservative)
            const int rb_last = (int)(upper_bound_dev(reads.blk_off, (int64_t)s_rb0, (int64_t)reads.n_blocks + 1, last) - 1);
            for (int b = s_rb0; b <= rb_last; ++b) atomicMax(blk_maxspan + b, s_maxspan);
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------------------------------------
// K2, pipelined variant: used when a whole tile's quality span fits one stage
// (tile_reads * max l_seq <= RS_STAGE - 32).
//  * every CTA owns a CONTIGUOUS range of tiles, so the site-row window and the read block are
//    carried from tile to tile instead of being re-searched (one binary search per CTA);
//  * two quality stages per CTA: the TMA bulk copy of tile j+1 is issued at the top of iteration j;
//  * headers are prefetched two tiles ahead, so the spans TMA needs are already in shared memory.
// ------------------------------------------------------------------------------------------------
constexpr int RS_STAGE = 36 * 1024;

// query index of reference position p, or -1 (pysam get_reference_positions(full_length=True).index)
__device__ __forceinline__ int cigar_qpos(const uint32_t* __restrict__ cg, int n_cigar, int32_t start, int32_t p) {
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



__global__ void __launch_bounds__(RS_THREADS)
read_scan_pipe_kernel(UnfzReadCols reads, UnfzSiteCols sites, const int32_t* __restrict__ mark_prefix, ScanParams P,
                      int tile_reads, UnfzReadSum* __restrict__ out, int32_t* __restrict__ row_lb,
                      int32_t* __restrict__ blk_maxspan, uint32_t* __restrict__ tile_tot) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* stage[2] = {smem, smem + RS_STAGE};
    int32_t* spos = reinterpret_cast<int32_t*>(smem + 2 * RS_STAGE);
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int64_t s_qa[3], s_qb[3];
    __shared__ int32_t s_wsum[RS_THREADS / 32];
    __shared__ int64_t s_base, s_row_end;
    __shared__ int32_t s_rb0, s_maxspan;

    const int64_t n = reads.n_reads;
    const int64_t n_tiles = (n + tile_reads - 1) / tile_reads;
    const int64_t tpc = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * tpc;
    const int64_t t1 = min(t0 + tpc, n_tiles);
    if (t0 >= t1) return;

    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    uint32_t phase[2] = {0, 0};
    const bool lane_ok = (int)threadIdx.x < tile_reads;

    auto hdr_of = [&](int64_t tile, bool& live) -> UnfzRead {
        const int64_t r = tile * tile_reads + threadIdx.x;
        live = lane_ok && tile < t1 && r < n;
        UnfzRead h;
        h.start = 0; h.l_seq = 0; h.n_cigar = 0; h.qoff_lo = 0; h.qoff_hi = 0;
        if (live) h = load_read(reads.hdr + r);
        return h;
    };
    auto publish_span = [&](int64_t tile, const UnfzRead& h, bool live) {     // slot tile % 3
        if (tile >= t1) return;
        const int64_t r0 = tile * tile_reads;
        const int64_t last = min(r0 + tile_reads, n) - 1;
        const int64_t r = r0 + threadIdx.x;
        const int slot = (int)((tile - t0) % 3);
        if (threadIdx.x == 0) s_qa[slot] = read_qoff(h);
        if (live && r == last) s_qb[slot] = read_qoff(h) + h.l_seq;
    };
    auto issue = [&](int64_t tile) {                                          // thread 0 only
        const int slot = (int)((tile - t0) % 3), st = (int)((tile - t0) & 1);
        const int64_t qa = s_qa[slot], qb = s_qb[slot];
        if (qb > qa) {
            const int64_t ga = qa & ~(int64_t)15;
            const uint32_t bytes = (uint32_t)(((qb - ga) + 15) & ~(int64_t)15);
            fence_proxy_async();
            mbar_expect_tx(&bar[st], bytes);
            tma_bulk_g2s(stage[st], reads.qual + ga, bytes, &bar[st]);
        }
    };

    bool live, live1, live2 = false;
    UnfzRead h = hdr_of(t0, live);
    UnfzRead h1 = hdr_of(t0 + 1, live1);
    UnfzRead h2 = h1;
    publish_span(t0, h, live);
    publish_span(t0 + 1, h1, live1);
    if (threadIdx.x == 0) {
        const int64_t r0 = t0 * tile_reads;
        const int rb = (int)(upper_bound_dev(reads.blk_off, 0, (int64_t)reads.n_blocks + 1, r0) - 1);
        s_rb0 = rb;
        const int sb0 = reads.blk_sblk[rb];
        int64_t rowb = 0, rowe = 0;
        if (sb0 >= 0) {
            rowe = sites.blk_off[sb0 + 1];
            rowb = lower_bound_dev(sites.pos, sites.blk_off[sb0], rowe, h.start);
        }
        s_base = rowb;
        s_row_end = rowe;
        s_maxspan = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) issue(t0);

    struct Pending { bool live; int64_t r; int64_t tile; UnfzReadSum o; int32_t lbs; } pend;
    pend.live = false; pend.tile = -1;
    auto flush = [&]() {                                       // after a barrier that follows the s_wsum writes
        if (pend.tile < 0) return;
        const int warp = threadIdx.x >> 5;
        int basew = 0, total = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) { const int v = s_wsum[w]; if (w < warp) basew += v; total += v; }
        if (threadIdx.x == 0) tile_tot[pend.tile] = (uint32_t)total;
        if (pend.live) {
            pend.o.hoff += (uint32_t)basew;
            *reinterpret_cast<int4*>(out + pend.r) = *reinterpret_cast<const int4*>(&pend.o);
            if (row_lb) row_lb[pend.r] = pend.lbs;
        }
    };
    for (int64_t tile = t0; tile < t1; ++tile) {
        const int st = (int)((tile - t0) & 1);
        const int slot = (int)((tile - t0) % 3);
        flush();                                               // results of the previous tile
        h2 = hdr_of(tile + 2, live2);                        // in flight during this iteration
        if (threadIdx.x == 0 && tile + 1 < t1) issue(tile + 1);
        const int64_t r0 = tile * tile_reads;
        const int64_t r = r0 + threadIdx.x;
        const int rb0 = s_rb0;
        const int64_t row_base = s_base, row_end = s_row_end;
        for (int i = threadIdx.x; i < RS_SPOS; i += RS_THREADS)
            spos[i] = (row_base + i < row_end) ? __ldg(sites.pos + row_base + i) : 0x7fffffff;
        int rb = rb0;
        if (live && r >= reads.blk_off[rb + 1])
            rb = (int)(upper_bound_dev(reads.blk_off, (int64_t)rb, (int64_t)reads.n_blocks + 1, r) - 1);
        const int sb = live ? reads.blk_sblk[rb] : -1;

        int32_t end = 0;
        uint32_t flags = 0;
        if (live) {
            int none_cnt = 0, non_m = 0;
            end = h.start;
            const uint32_t* cg = reads.cigar + h.cigar_off;
            for (int k = 0; k < h.n_cigar; ++k) {
                const uint32_t w = __ldg(cg + k);
                const uint32_t op = w & 15u, ln = w >> 4;
                if (op == 0 || op == 7 || op == 8 || op == 2 || op == 3) end += (int32_t)ln;
                if (op == 1 || op == 4) none_cnt += (int)ln;
                if (op != 0 && op != 7) ++non_m;
            }
            const uint32_t f = h.flag;
            const bool base_ok = !(f & (0x200u | 0x4u | 0x400u | 0x100u | 0x800u | 0x8u)) &&
                                 (int)h.mapq >= P.min_mapq && (h.aux & 1u);
            if (base_ok) flags |= UNFZ_RS_GOOD_DISC;
            if (none_cnt <= 5) flags |= UNFZ_RS_NONE_OK;
            if (non_m <= 5) flags |= UNFZ_RS_EXT_OK;
            long long ins = (long long)h.tlen - 2ll * P.readlen;
            if (ins < 0) ins = -ins;
            if ((double)ins <= reads.blk_cul[rb]) flags |= UNFZ_RS_INS_OK;
            if ((f & 1u) && !(f & 8u) && h.mate >= 0) flags |= UNFZ_RS_HAS_MATE;
            atomicMax(&s_maxspan, end - h.start);
        }
        __syncthreads();   // spos visible

        int32_t fmark = 0, cnt = 0;
        int64_t lbs = 0, lbe = 0;
        if (live && sb >= 0) {
            if (rb == rb0) {
                // a tile spans a few site rows only: gallop from the front of the staged window
                int lo = 0, hi = 8;
                while (hi < RS_SPOS && spos[hi - 1] < h.start) { lo = hi; hi = min(hi * 2, RS_SPOS); }
                while (lo < hi) { int mid = (lo + hi) >> 1; if (spos[mid] < h.start) lo = mid + 1; else hi = mid; }
                int lo2 = lo; hi = min(lo + 4, RS_SPOS);
                while (hi < RS_SPOS && spos[hi - 1] < end) { lo2 = hi; hi = min(hi + 8, RS_SPOS); }
                while (lo2 < hi) { int mid = (lo2 + hi) >> 1; if (spos[mid] < end) lo2 = mid + 1; else hi = mid; }
                lbs = row_base + lo;
                lbe = row_base + lo2;
                if (lo2 == RS_SPOS) {
                    if (lo == RS_SPOS) lbs = lower_bound_dev(sites.pos, row_base + RS_SPOS - 1, row_end, h.start);
                    lbe = lower_bound_dev(sites.pos, lbs, row_end, end);
                }
            } else {
                const int64_t a = sites.blk_off[sb], b = sites.blk_off[sb + 1];
                lbs = lower_bound_dev(sites.pos, a, b, h.start);
                lbe = lower_bound_dev(sites.pos, lbs, b, end);
            }
            fmark = __ldg(mark_prefix + lbs);
            const int32_t c = __ldg(mark_prefix + lbe) - fmark;
            cnt = c > 0xffff ? 0xffff : c;
        }

        // ---- hit slots: offset of the read inside its tile (block scan of cnt) + the tile's total; the
        // pipeline scans the tile totals, so hoff(read) = tile_base[tile] + local offset.  Only the warp
        // part of the scan runs here; it is finished after the tile's closing barrier (see `pending`),
        // so the scan costs no barrier of its own.
        int warp_incl = cnt;
        {
            const int lane = threadIdx.x & 31;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, warp_incl, o); if (lane >= o) warp_incl += y; }
        }

        int low = 0;
        const int64_t qa = s_qa[slot], qb = s_qb[slot];
        if (qb > qa) {
            mbar_wait(&bar[st], phase[st]);
            phase[st] ^= 1u;
            if (live && h.l_seq > 0) {
                const int off = (int)(read_qoff(h) - (qa & ~(int64_t)15));
                low = count_low_quals(stage[st], off, off + h.l_seq, (uint32_t)P.min_bq);
            }
        }
        if (live && (flags & UNFZ_RS_GOOD_DISC) && low <= 10 && h.n_cigar <= 10) flags |= UNFZ_RS_GOOD_CONC;
        // stash this tile's results; they are written once the block scan is complete
        pend.live = live; pend.r = r; pend.tile = tile;
        pend.o.end = end; pend.o.fmark = fmark; pend.o.flags = (uint16_t)flags; pend.o.cnt = (uint16_t)cnt;
        pend.o.hoff = (uint32_t)(warp_incl - cnt);
        pend.lbs = (int32_t)lbs;
        if ((threadIdx.x & 31) == 31) s_wsum[threadIdx.x >> 5] = warp_incl;
        // ---- hand-over to the next tile ---------------------------------------------------------
        publish_span(tile + 2, h2, live2);                   // slot (tile+2)%3 is not in use
        if (threadIdx.x == 0) {
            const int64_t rn = (tile + 1) * tile_reads;
            // flush the span of this tile to every block it touches, then carry the window forward
            const int64_t last = min(r0 + tile_reads, n) - 1;
            int rb_last = rb0;
            if (last >= reads.blk_off[rb0 + 1])
                rb_last = (int)(upper_bound_dev(reads.blk_off, (int64_t)rb0, (int64_t)reads.n_blocks + 1, last) - 1);
            if (s_maxspan > 0)
                for (int b = rb0; b <= rb_last; ++b) atomicMax(blk_maxspan + b, s_maxspan);
            s_maxspan = 0;
            if (tile + 1 < t1 && rn < n) {
                if (rn >= reads.blk_off[rb0 + 1]) {           // next tile starts in another read block
                    const int rbn = (int)(upper_bound_dev(reads.blk_off, (int64_t)rb0, (int64_t)reads.n_blocks + 1, rn) - 1);
                    s_rb0 = rbn;
                    const int sbn = reads.blk_sblk[rbn];
                    int64_t rowb = 0, rowe = 0;
                    if (sbn >= 0) {
                        rowe = sites.blk_off[sbn + 1];
                        rowb = lower_bound_dev(sites.pos, sites.blk_off[sbn], rowe, h1.start);
                    }
                    s_base = rowb;
                    s_row_end = rowe;
                } else {
                    int lo = 0, hi = RS_SPOS;                 // first staged row with pos >= next tile's first start
                    while (lo < hi) { int mid = (lo + hi) >> 1; if (spos[mid] < h1.start) lo = mid + 1; else hi = mid; }
                    int64_t nb = row_base + lo;
                    if (lo == RS_SPOS) nb = lower_bound_dev(sites.pos, row_base + RS_SPOS - 1, row_end, h1.start);
                    s_base = nb;
                }
            }
        }
        h = h1; live = live1;
        h1 = h2; live1 = live2;
        __syncthreads();   // stage st, spos, s_wsum and the carried state are consistent for the next tile
    }
    flush();
}

// ------------------------------------------------------------------------------------------------
// K3: read x marked-site allele lookup
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
read_site_alleles_kernel(UnfzReadCols reads, UnfzSiteCols sites, const uint8_t* __restrict__ row_mark,
                         const int32_t* __restrict__ mark_prefix, const UnfzReadSum* __restrict__ rsum,
                         const int32_t* __restrict__ row_lb, const uint32_t* __restrict__ tile_base, int32_t tile_reads,
                         uint32_t* __restrict__ hits) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= reads.n_reads) return;
    const UnfzReadSum s = load_rsum(rsum + r);
    if (s.cnt == 0) return;
    const UnfzRead h = load_read(reads.hdr + r);
    const int rb = (int)(upper_bound_dev(reads.blk_off, 0, (int64_t)reads.n_blocks + 1, r) - 1);
    const int sb = reads.blk_sblk[rb];
    if (sb < 0) return;
    const int64_t b = sites.blk_off[sb + 1];
    int64_t row = __ldg(row_lb + r);                       // first site row with pos >= start (from read_scan)
    const uint32_t* cg = reads.cigar + h.cigar_off;
    const int64_t q0 = read_qoff(h);
    const int64_t hbase = (int64_t)__ldg(tile_base + (uint32_t)r / (uint32_t)tile_reads) + s.hoff;
    int written = 0;
    for (; row < b && written < s.cnt; ++row) {
        const int32_t p = __ldg(sites.pos + row);
        if (p >= s.end) break;
        const int mp0 = __ldg(mark_prefix + row);
        if (__ldg(mark_prefix + row + 1) == mp0) continue;      // not a marked row (same sectors as mp0: no row_mark load)
        const int k = mp0 - s.fmark;
        uint32_t word = 0;
        const int q = cigar_qpos(cg, h.n_cigar, h.start, p);
        if (q >= 0 && q < 0xffff) {
            const int64_t g = q0 + q;
            const uint32_t qb = __ldg(reads.qual + g);
            const uint32_t code = (__ldg(reads.seq2 + (g >> 2)) >> ((g & 3) << 1)) & 3u;
            word = (uint32_t)(q + 1) | (qb << 16) | (code << 24) | ((q + 1 < h.l_seq) ? (1u << 26) : 0u);
        }
        if (k >= 0 && k < s.cnt) hits[hbase + k] = word;
        ++written;
    }
}

}  // namespace

static int scan_tile_reads(int32_t max_l_seq) {
    if (max_l_seq > 0 && (int64_t)max_l_seq + 32 <= RS_STAGE) {
        int t = (RS_STAGE - 32) / max_l_seq;
        return t > RS_THREADS ? RS_THREADS : t;
    }
    return RS_THREADS;
}

extern "C" int32_t unfz_read_scan_tile_reads(int32_t max_l_seq) { return scan_tile_reads(max_l_seq); }

extern "C" int unfz_read_scan(UnfzCtx* ctx, const UnfzReadCols* reads, const UnfzSiteCols* sites,
                              const int32_t* mark_prefix, const UnfzParams* hp, int32_t max_l_seq, UnfzReadSum* out,
                              int32_t* row_lb, int32_t* blk_maxspan, uint32_t* tile_tot, void* stream) {
    if (reads->n_reads <= 0) return 0;
    ScanParams P;
    P.min_mapq = hp->min_map_qual;
    double bq = hp->min_gt_qual;
    P.min_bq = bq <= 0 ? 0 : (bq >= 128 ? 128 : (int32_t)ceil(bq));
    P.readlen = hp->readlen;
    if (max_l_seq > 0 && (int64_t)max_l_seq + 32 <= RS_STAGE) {
        const int tile_reads = scan_tile_reads(max_l_seq);
        const size_t smem2 = 2 * RS_STAGE + RS_SPOS * sizeof(int32_t);
        static bool attr2 = false;
        if (!attr2) {
            UNFZ_CHECK(ctx, cudaFuncSetAttribute(read_scan_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
            attr2 = true;
        }
        const int64_t tiles = (reads->n_reads + tile_reads - 1) / tile_reads;
        int64_t g = (int64_t)ctx->sm_count * 3;     // 3 CTAs x 72 KB of staging per SM
        if (g > tiles) g = tiles;
        read_scan_pipe_kernel<<<(unsigned)g, RS_THREADS, smem2, (cudaStream_t)stream>>>(*reads, *sites, mark_prefix, P, tile_reads,
                                                                                     out, row_lb, blk_maxspan, tile_tot);
        UNFZ_LAUNCH_CHECK(ctx);
        return 0;
    }
    const size_t smem = RS_QBUF + 16 + RS_SPOS * sizeof(int32_t);
    static bool attr_set = false;
    if (!attr_set) {
        UNFZ_CHECK(ctx, cudaFuncSetAttribute(read_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int64_t n_tiles = (reads->n_reads + RS_THREADS - 1) / RS_THREADS;
    int64_t grid = (int64_t)ctx->sm_count * 4;     // 4 x ~49 KB of staging per SM
    if (grid > n_tiles) grid = n_tiles;
    read_scan_kernel<<<(unsigned)grid, RS_THREADS, smem, (cudaStream_t)stream>>>(*reads, *sites, mark_prefix, P, out, row_lb, blk_maxspan, tile_tot);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int unfz_read_site_alleles(UnfzCtx* ctx, const UnfzReadCols* reads, const UnfzSiteCols* sites,
                                      const uint8_t* row_mark, const int32_t* mark_prefix,
                                      const UnfzReadSum* rsum, const int32_t* row_lb, const uint32_t* tile_base,
                                      int32_t tile_reads, uint32_t* hits, void* stream) {
    if (reads->n_reads <= 0) return 0;
    const int64_t blocks = (reads->n_reads + 255) / 256;
    read_site_alleles_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*reads, *sites, row_mark, mark_prefix, rsum, row_lb,
                                                                                 tile_base, tile_reads, hits);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}
