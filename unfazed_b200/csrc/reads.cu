// Read path, part 1: the streaming read scan (goodread + pair-independent filters + overlap counts)
// and the read-by-site allele lookup.
//
// read_scan is the bandwidth kernel of the read path: per read it must see the 32 B header, the
// CIGAR words and one low-quality bit per base.  Reads are stored in file order, so the quality bits
// (and the CIGAR words) of consecutive reads are one contiguous span that is staged in shared memory
// by asynchronous copies while earlier reads are being processed; every thread then counts the
// low-quality bases of its own read out of shared memory (five or six POPCs for 150 bases).
//   read_scan_warp_kernel  one cp.async pipeline per warp, 32 reads per tile; reads too long for a
//                          warp's slice are counted straight out of global memory
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ------------------------------------------------------------------------------------------------
// K2: read scan
// ------------------------------------------------------------------------------------------------
struct ScanParams {
    int32_t min_mapq;
    int32_t readlen;
};

// number of set bits in [b0, b1) of a little-endian bit array (bit i = bit i&31 of word i>>5): the low-quality
// bases of one read.  A 150-base read is five or six words; the first and the last one are masked.
template <typename Ptr>
__device__ __forceinline__ int count_bits(Ptr w, int b0, int b1) {
    if (b1 <= b0) return 0;
    const int w0 = b0 >> 5, w1 = (b1 - 1) >> 5;
    const uint32_t head = 0xffffffffu << (b0 & 31);
    const uint32_t tail = 0xffffffffu >> (31 - ((b1 - 1) & 31));
    if (w0 == w1) return __popc(w[w0] & head & tail);
    int acc = __popc(w[w0] & head);
#pragma unroll 4
    for (int i = w0 + 1; i < w1; ++i) acc += __popc(w[i]);
    return acc + __popc(w[w1] & tail);
}

// ------------------------------------------------------------------------------------------------
// K2, streaming kernel: one independent
// software pipeline PER WARP, no CTA-wide synchronisation at all.
//
// Measured on B200 (profiles/README.md, round 1d): a single asynchronous copy stream -- one TMA bulk
// copy or the LDGSTS copies of one warp -- moves about 7 KB/us, so a CTA-wide double buffer filled by
// one producer tops out near 60 % of HBM bandwidth however the tiles are sized (the consumers wait
// for the fill, the producer waits for the consumers).  What saturates HBM is MANY concurrent
// streams, so here every warp is its own stream:
//  * a warp owns a contiguous range of 32-read tiles ("hit tiles") and two private slices of shared
//    memory; while it works on tile T out of one slice, the quality bits and CIGAR words of tile
//    T+1 -- each one contiguous span, because reads are stored in file order -- are in flight into
//    the other slice (cp.async 16-byte chunks, one commit group per tile), and the headers of tile
//    T+2 are in flight into registers;
//  * the site rows a tile can overlap are a window of 64 positions carried from tile to tile (reads
//    are sorted by start): it advances 32 rows at a time out of a register prefetched one step
//    ahead, so the steady state issues no dependent global load for it.  When a tile runs into the
//    next read block that block's window is opened next to it and promoted when the tiles get there;
//  * per read: CIGAR walk out of shared memory, forward probe of the window (row positions and their mark
//    prefixes side by side: a read's hit count is a difference of two window entries), low-quality count by
//    POPC over the staged bit span;
//  * hit slots are numbered per tile: one warp scan, one total per tile, no cross-warp scan.
// ------------------------------------------------------------------------------------------------
// 6 warps x 4 CTAs per SM (80 registers, no spills).  With quality BYTES staged the kernel was bandwidth-bound and
// 5 x 4 (96 registers) was the best point; with one BIT per base it is bound by instruction issue (about 800 issue
// slots per 32-read tile) and wants the extra warps: measured 0.522 ms (5 x 4), 0.493 ms (6 x 4), 0.613 ms (8 x 4: 64
// registers, spills) for 22.7 M reads.
#ifndef WP_MINB
#define WP_MINB 4
#endif
#ifndef WP_WARPS_N
#define WP_WARPS_N 6
#endif
constexpr int WP_WARPS = WP_WARPS_N;
constexpr int WP_THREADS = WP_WARPS * 32;
constexpr int WP_CIGW = 64;                               // staged CIGAR words per tile
constexpr int WP_SPOS = 64;                               // site positions per window
constexpr int WP_FIXED = 2 * WP_CIGW * 4 + 4 * WP_SPOS * 4;   // per warp, next to the two quality slices: CIGAR words, row positions + mark prefixes

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
// predicated 16-byte copy at a compile-time offset from both bases (no address arithmetic per chunk)
template <int OFF>
__device__ __forceinline__ void cp_async16_at(uint32_t dst_smem, const void* src, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\t"
                 "@p cp.async.cg.shared.global [%0 + %3], [%1 + %3], 16;\n\t}"
                 ::"r"(dst_smem), "l"(src), "r"((int)pred), "n"(OFF) : "memory");
}
template <int I, int N>
struct CpChunks {
    static __device__ __forceinline__ void run(uint32_t dst, const uint8_t* src, uint32_t o, uint32_t bytes) {
        cp_async16_at<I * 512>(dst, src, o + I * 512u < bytes);
        CpChunks<I + 1, N>::run(dst, src, o, bytes);
    }
};
template <int N>
struct CpChunks<N, N> {
    static __device__ __forceinline__ void run(uint32_t, const uint8_t*, uint32_t, uint32_t) {}
};
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// reference_end, query bases without a reference position and non-M/= operations of one CIGAR
template <typename Ptr>
__device__ __forceinline__ void cigar_summary(Ptr cg, int n_cigar, int32_t& end, int& none_cnt, int& non_m) {
    for (int k = 0; k < n_cigar; ++k) {
        const uint32_t w = cg[k];
        const uint32_t op = w & 15u, ln = w >> 4;
        if ((0x18du >> op) & 1u) end += (int32_t)ln;          // M D N = X
        if ((0x012u >> op) & 1u) none_cnt += (int)ln;         // I S
        non_m += (int)((0xff7eu >> op) & 1u);                 // everything but M and =
    }
}

// first index in [lo, hi) with a[i] >= v, searched by a whole warp: 32 probes per round trip
__device__ __forceinline__ int64_t warp_lower_bound(const int32_t* __restrict__ a, int64_t lo, int64_t hi, int32_t v, int lane) {
    while (hi - lo > 32) {
        const int64_t step = (hi - lo + 32) / 33;
        const int64_t i = lo + (int64_t)(lane + 1) * step - 1;
        const bool lt = i < hi && __ldg(a + i) < v;
        const int k = __popc(__ballot_sync(0xffffffffu, lt));      // probes are monotone
        const int64_t nhi = k < 32 ? min(hi, lo + (int64_t)(k + 1) * step - 1) : hi;
        lo += (int64_t)k * step;
        hi = nhi;
    }
    const bool lt = lo + lane < hi && __ldg(a + lo + lane) < v;
    return lo + __popc(__ballot_sync(0xffffffffu, lt));
}

struct WpSpan {               // where the staged spans of a tile start, and whether they were staged
    uint32_t qa_lo, ca;
    bool q_ok, cig_ok;
};

__global__ void __launch_bounds__(WP_THREADS, WP_MINB)
read_scan_warp_kernel(UnfzReadCols reads, UnfzSiteCols sites, const int32_t* __restrict__ mark_prefix, ScanParams P,
                      int qslice, UnfzReadSum* __restrict__ out,
                      int32_t* __restrict__ blk_maxspan, uint32_t* __restrict__ tile_tot, uint2* __restrict__ tile_info,
                      const int32_t* __restrict__ guard) {
    UNFZ_GUARD(guard);
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* my = smem + warp * (2 * qslice + WP_FIXED);
    uint32_t* cigbuf = reinterpret_cast<uint32_t*>(my + 2 * qslice);
    int32_t* spos = reinterpret_cast<int32_t*>(my + 2 * qslice + 2 * WP_CIGW * 4);
    int32_t* smp = spos + 2 * WP_SPOS;                       // mark_prefix of the same rows: a read's hit count is a difference of two

    // read and row indices fit 32 bits (mate / row_lb are int32 in the ABI)
    const int n = (int)reads.n_reads;
    const int n_tiles = (n + 31) >> 5;
    const int n_warps = (int)gridDim.x * WP_WARPS;
    const int tpw = (n_tiles + n_warps - 1) / n_warps;
    const int t0 = ((int)blockIdx.x * WP_WARPS + warp) * tpw;
    const int t1 = min(t0 + tpw, n_tiles);
    if (t0 >= t1) return;
    const uint32_t FULL = 0xffffffffu;

    auto load_hdr = [&](int T, uint4& a, uint4& b) {
        a = make_uint4(0, 0, 0, 0);
        b = a;
        const int r = (T << 5) + lane;
        if (T < t1 && r < n) {
            const uint4* q = reinterpret_cast<const uint4*>(reads.hdr + r);
            a = __ldg(q);
            b = __ldg(q + 1);
        }
    };
    // start the copies of tile T (headers a, b) into slice k; exactly one commit group per call
    auto issue_fill = [&](int T, const uint4& a, const uint4& b, int k) -> WpSpan {
        WpSpan sp;
        sp.qa_lo = 0; sp.ca = 0; sp.q_ok = false; sp.cig_ok = false;
        if (T < t1) {
            const int nl = min(32, n - (T << 5));
            const long long q_self = (long long)b.x | ((long long)((b.w >> 16) & 0xffu) << 32);
            const long long f_q = __shfl_sync(FULL, q_self, 0);
            const long long l_q = __shfl_sync(FULL, q_self + (long long)(int32_t)b.y, nl - 1);
            const uint32_t f_c = __shfl_sync(FULL, a.w, 0);
            const uint32_t l_c = __shfl_sync(FULL, a.w + (b.z >> 16), nl - 1);
            // bytes of the bit plane that hold bases [f_q, l_q), widened to 16-byte chunks
            const long long qa = (f_q >> 3) & ~15ll;
            const long long qspan = l_q > f_q ? ((((l_q + 7) >> 3) - qa + 15) & ~15ll) : 0;
            const uint32_t ca = f_c & ~3u;
            const long long cspan = l_c > ca ? (((long long)l_c - ca + 3) & ~3ll) : 0;
            sp.q_ok = qspan <= qslice;
            sp.cig_ok = cspan <= WP_CIGW && (long long)ca + cspan <= reads.n_cigar;
            sp.qa_lo = (uint32_t)(qa << 3);                  // first staged BIT (low 32 bits; differences stay small)
            sp.ca = ca;
            const uint32_t o = lane * 16u;
            if (sp.q_ok) {
                // lane-strided 16-byte chunks: the first two (1 KB: 32 reads of up to 250 bases) are unrolled
                // with immediate offsets, longer slices finish in a loop
                const uint32_t dst = smem_u32(my + k * qslice) + o;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(reads.lowq) + qa + o;
                CpChunks<0, 2>::run(dst, src, o, (uint32_t)qspan);
                for (uint32_t x = o + 1024u; x < (uint32_t)qspan; x += 512u)
                    cp_async16(my + k * qslice + x, reinterpret_cast<const uint8_t*>(reads.lowq) + qa + x);
            }
            if (sp.cig_ok)                                      // <= 256 bytes: one chunk per lane
                cp_async16_at<0>(smem_u32(cigbuf + k * WP_CIGW) + o, reinterpret_cast<const uint8_t*>(reads.cigar + ca) + o,
                                 o < (uint32_t)cspan * 4u);
        }
        cp_async_commit();
        return sp;
    };

    // ---- carried state: the read block of the current tile and its window of site rows -------------
    int rb0 = 0, sblk = -1;
    int blk_end = -1, row_base = 0, row_end = 0;
    long long cul = -1;                                           // floor(concordant_upper_len): |insert| is an integer
    // The window lives in shared memory only: registers written by global loads would make every hot-path
    // instruction that reads them wait on the (statically assigned) scoreboard of whatever load is in flight.
    int32_t w2 = 0x7fffffff, w2m = 0;                            // rows row_base + 64 + lane, read only when the window advances
    int cur = 0;                                                 // first window entry >= the tile's first start (uniform)
    int bmax = 0;                                                // longest span seen in the current block
    int rb1 = -1, sblk1 = -1;                                    // the next block, opened when a tile runs into it
    int blk_end1 = 0, row_base1 = 0, row_end1 = 0;
    long long cul1 = -1;
    // (double)ins <= c for an integer ins >= 0  <=>  ins <= floor(c); NaN and negative bounds admit nothing
    auto cul_floor = [](double c) -> long long { return c >= 0.0 ? (c < 9.0e18 ? (long long)floor(c) : 0x7fffffffffffffffll) : -1; };
    auto row_at = [&](int i, int rend) -> int32_t { return i < rend ? __ldg(sites.pos + i) : 0x7fffffff; };
    auto mp_at = [&](int i, int rend) -> int32_t { return __ldg(mark_prefix + min(i, rend)); };   // rows past the block: its total
    auto open_block = [&](int rb, int32_t first_start, int& sb, int& bend, long long& c, int& rbase, int& rend) {
        bend = (int)reads.blk_off[rb + 1];
        sb = reads.blk_sblk[rb];
        c = cul_floor(reads.blk_cul[rb]);
        rbase = rend = 0;
        if (sb >= 0) {
            rend = (int)sites.blk_off[sb + 1];
            rbase = (int)warp_lower_bound(sites.pos, sites.blk_off[sb], rend, first_start, lane);
        }
    };

    uint4 hAa, hAb, hBa, hBb, hCa, hCb;
    load_hdr(t0, hAa, hAb);
    load_hdr(t0 + 1, hBa, hBb);
    WpSpan spA = issue_fill(t0, hAa, hAb, 0), spB = spA;

    for (int T = t0; T < t1; ++T) {
        const int k = (T - t0) & 1;
        spB = issue_fill(T + 1, hBa, hBb, k ^ 1);              // in flight while this tile is processed
        load_hdr(T + 2, hCa, hCb);                             // consumed by the next iteration's fill

        const int r0 = T << 5;
        const int nl = min(32, n - r0);
        const int last = r0 + nl - 1;
        const int32_t f_start = (int32_t)__shfl_sync(FULL, hAa.x, 0);
        if (r0 >= blk_end) {                                   // the tile starts in a new read block
            if (blk_end < 0) rb0 = (int)(upper_bound_dev(reads.blk_off, 0, (int64_t)reads.n_blocks + 1, r0) - 1);
            else do { ++rb0; } while (r0 >= reads.blk_off[rb0 + 1]);
            if (rb0 == rb1) {                                  // opened while the previous tile ran into it
                sblk = sblk1; blk_end = blk_end1; cul = cul1; row_base = row_base1; row_end = row_end1;
                const int32_t a = spos[WP_SPOS + lane], b = spos[WP_SPOS + 32 + lane];
                const int32_t am = smp[WP_SPOS + lane], bm = smp[WP_SPOS + 32 + lane];
                spos[lane] = a;
                spos[32 + lane] = b;
                smp[lane] = am;
                smp[32 + lane] = bm;
            } else {
                open_block(rb0, f_start, sblk, blk_end, cul, row_base, row_end);
                spos[lane] = row_at(row_base + lane, row_end);
                spos[32 + lane] = row_at(row_base + 32 + lane, row_end);
                smp[lane] = mp_at(row_base + lane, row_end);
                smp[32 + lane] = mp_at(row_base + 32 + lane, row_end);
            }
            w2 = row_at(row_base + 64 + lane, row_end);
            w2m = mp_at(row_base + 64 + lane, row_end);
            cur = 0;
            bmax = 0;
            __syncwarp();
        }
        // reads are sorted by start: rows before the tile's first start are behind every read from here on
        while (cur < WP_SPOS && spos[cur] < f_start) ++cur;
        if (cur >= 32) {                                       // advance the window by 32 rows (or re-seat it after a gap)
            if (cur == WP_SPOS) {
                row_base = (int)warp_lower_bound(sites.pos, row_base + WP_SPOS, row_end, f_start, lane);
                spos[lane] = row_at(row_base + lane, row_end);
                spos[32 + lane] = row_at(row_base + 32 + lane, row_end);
                smp[lane] = mp_at(row_base + lane, row_end);
                smp[32 + lane] = mp_at(row_base + 32 + lane, row_end);
                cur = 0;
            } else {
                const int32_t b = spos[32 + lane], bm = smp[32 + lane];
                spos[lane] = b;
                spos[32 + lane] = w2;
                smp[lane] = bm;
                smp[32 + lane] = w2m;
                row_base += 32;
                cur -= 32;
            }
            w2 = row_at(row_base + 64 + lane, row_end);        // consumed at the next advance
            w2m = mp_at(row_base + 64 + lane, row_end);
            __syncwarp();
        }
        const bool two = last >= blk_end;                      // the tile runs into the next read block
        if (two && rb1 != rb0 + 1) {
            rb1 = rb0 + 1;
            const int32_t first1 = __ldg(&reads.hdr[blk_end].start);
            open_block(rb1, first1, sblk1, blk_end1, cul1, row_base1, row_end1);
            spos[WP_SPOS + lane] = row_at(row_base1 + lane, row_end1);
            spos[WP_SPOS + 32 + lane] = row_at(row_base1 + 32 + lane, row_end1);
            smp[WP_SPOS + lane] = mp_at(row_base1 + lane, row_end1);
            smp[WP_SPOS + 32 + lane] = mp_at(row_base1 + 32 + lane, row_end1);
        }

        cp_async_wait<1>();                                    // this tile's spans have landed (all but the newest group)
        __syncwarp();

        // ---- one read per lane ---------------------------------------------------------------------
        const bool live = lane < nl;
        const int r = r0 + lane;
        const int32_t start = (int32_t)hAa.x, tlen = (int32_t)hAa.y, mate = (int32_t)hAa.z;
        const int l_seq = (int)hAb.y, n_cigar = (int)(hAb.z >> 16);
        const uint32_t f = hAb.z & 0xffffu, mapq = hAb.w & 0xffu, aux = (hAb.w >> 8) & 0xffu;

        // which staged block the read belongs to: 0 the tile's, 1 the next one, 2 neither (tiny blocks: rare)
        int which = 0;
        if (two && live && r >= blk_end) which = r < blk_end1 ? 1 : 2;
        int rb = rb0 + which, sb = which ? sblk1 : sblk;
        long long my_cul = which ? cul1 : cul;
        const int rbase = which ? row_base1 : row_base, rend = which ? row_end1 : row_end;
        if (which == 2) {
            rb = (int)(upper_bound_dev(reads.blk_off, (int64_t)rb0, (int64_t)reads.n_blocks + 1, r) - 1);
            sb = reads.blk_sblk[rb];
            my_cul = cul_floor(reads.blk_cul[rb]);
        }

        int32_t end = start;
        uint32_t flags = 0;
        if (live) {
            int none_cnt = 0, non_m = 0;
            if (spA.cig_ok) cigar_summary(cigbuf + k * WP_CIGW + (hAa.w - spA.ca), n_cigar, end, none_cnt, non_m);
            else cigar_summary(reads.cigar + hAa.w, n_cigar, end, none_cnt, non_m);
            const bool base_ok = !(f & (0x200u | 0x4u | 0x400u | 0x100u | 0x800u | 0x8u)) &&
                                 (int)mapq >= P.min_mapq && (aux & 1u);
            if (base_ok) flags |= UNFZ_RS_GOOD_DISC;
            if (none_cnt <= 5) flags |= UNFZ_RS_NONE_OK;
            if (non_m <= 5) flags |= UNFZ_RS_EXT_OK;
            long long ins = (long long)tlen - 2ll * P.readlen;
            if (ins < 0) ins = -ins;
            if (ins <= my_cul) flags |= UNFZ_RS_INS_OK;
            if ((f & 1u) && !(f & 8u) && mate >= 0) flags |= UNFZ_RS_HAS_MATE;
        }
        // longest reference span per read block (bounds the fetch windows of the chaining kernels)
        {
            const int span = live ? end - start : 0;
            const int m0 = __reduce_max_sync(FULL, which == 0 ? span : 0);
            if (m0 > bmax) {
                bmax = m0;
                if (lane == 0) atomicMax(blk_maxspan + rb0, m0);
            }
            if (two) {
                const int m1 = __reduce_max_sync(FULL, which == 1 ? span : 0);
                if (lane == 0 && m1 > 0) atomicMax(blk_maxspan + rb0 + 1, m1);
                if (which == 2 && span > 0) atomicMax(blk_maxspan + rb, span);
            }
        }

        // ---- marked-site overlap: rows with start <= pos < end ----------------------------------------
        int32_t fmark = 0, emark = 0;
        int lbs = 0, lbe = 0;
        if (live && sb >= 0) {
            if (which < 2) {
                // the tile's reads start within a row or two of the cursor and span a row or two: probe forward
                const int32_t* win = spos + which * WP_SPOS;
                int lo = which ? 0 : cur;
                while (lo < WP_SPOS && win[lo] < start) ++lo;
                int lo2 = lo;
                while (lo2 < WP_SPOS && win[lo2] < end) ++lo2;
                lbs = rbase + lo;
                lbe = rbase + lo2;
                if (lo2 == WP_SPOS) {                          // ran off the staged window: finish in global memory
                    if (lo == WP_SPOS) lbs = (int)lower_bound_dev(sites.pos, rbase + WP_SPOS - 1, rend, start);
                    lbe = (int)lower_bound_dev(sites.pos, max(lbs, rbase + WP_SPOS - 1), rend, end);
                    fmark = __ldg(mark_prefix + lbs);
                    emark = __ldg(mark_prefix + lbe);
                } else {                                       // the mark prefixes of the window's rows sit next to their positions
                    const int32_t* wmp = smp + which * WP_SPOS;
                    fmark = wmp[lo];
                    emark = wmp[lo2];
                }
            } else {
                const int64_t a = sites.blk_off[sb], b = sites.blk_off[sb + 1];
                lbs = (int)lower_bound_dev(sites.pos, a, b, start);
                lbe = (int)lower_bound_dev(sites.pos, lbs, b, end);
                fmark = __ldg(mark_prefix + lbs);
                emark = __ldg(mark_prefix + lbe);
            }
        }

        // ---- goodread: low-quality bases out of the staged span ---------------------------------------
        int low = 0;
        if (live && l_seq > 0) {
            if (spA.q_ok) {
                const int off = (int)(hAb.x - spA.qa_lo);
                low = count_bits(reinterpret_cast<const uint32_t*>(my + k * qslice), off, off + l_seq);
            } else {
                const int64_t q0 = (int64_t)hAb.x | ((int64_t)((hAb.w >> 16) & 0xffu) << 32);
                const int off = (int)(q0 & 31);
                low = count_bits(reads.lowq + (q0 >> 5), off, off + l_seq);
            }
        }
        if (live && (flags & UNFZ_RS_GOOD_DISC) && low <= 10 && n_cigar <= 10) flags |= UNFZ_RS_GOOD_CONC;

        const int32_t cnt = min(emark - fmark, 0xffff);
        // ---- hit slots: offset inside the tile + the tile's total -----------------------------------------
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += y; }
        if (lane == 31) tile_tot[T] = (uint32_t)incl;
        {
            const unsigned has = __ballot_sync(FULL, cnt > 0);    // who has hits + the tile's read block, for the lookup kernel
            if (tile_info && lane == 0) tile_info[T] = make_uint2(has, (uint32_t)rb0);
        }
        if (live) {
            int4* o = reinterpret_cast<int4*>(out + r);        // the 32-byte summary as two 16-byte stores
            o[0] = make_int4(end, fmark, (int)((uint32_t)(flags & 0xffffu) | ((uint32_t)cnt << 16)), incl - cnt);
            o[1] = make_int4(start, mate, lbs, 0);
        }
        __syncwarp();                                          // the slice may be refilled by the next iteration
        hAa = hBa; hAb = hBb; hBa = hCa; hBb = hCb;
        spA = spB;
    }
    cp_async_wait<0>();
}

// query index of reference position p, or -1 (pysam get_reference_positions(full_length=True).index)
__device__ __forceinline__ int cigar_qpos(const uint32_t* __restrict__ cg, int n_cigar, int32_t start, int32_t p, uint32_t w0) {
    int32_t cur = start;
    int q = 0;
    for (int k = 0; k < n_cigar; ++k) {
        const uint32_t w = k == 0 ? w0 : __ldg(cg + k);     // the first word was requested together with the row data
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




// ------------------------------------------------------------------------------------------------
// K3: read x marked-site allele lookup
// ------------------------------------------------------------------------------------------------
// the lookup of ONE read with hits: walk the site rows it spans, CIGAR walk per marked row, three bit / base gathers.
// The kernel lives on the latency of dependent gathers, so everything whose address is known is requested at once:
// summary and header together; then the first CIGAR word, the first row's position and its two mark prefixes; inside
// the walk the next row is requested before the current one is looked up.  Three round trips for the common read
// (one M operation, one marked site) instead of seven.
__device__ __forceinline__ void read_alleles_one(const UnfzReadCols& reads, const UnfzSiteCols& sites,
                                                 const int32_t* __restrict__ mark_prefix, const UnfzReadSum* __restrict__ rsum,
                                                 const uint32_t* __restrict__ tile_base, int32_t tile_reads,
                                                 uint32_t* __restrict__ hits, int64_t r) {
    const UnfzReadSum s = load_rsum(rsum + r);
    const UnfzRead h = load_read(reads.hdr + r);           // independent of the summary: both sectors fly together
    if (s.cnt == 0) return;
    // nine reads in ten are one M operation covering every base (aux bit6, set at upload by unfz_read_starts): their
    // only CIGAR word is known without fetching it -- one gathered sector less per read
    const bool simple = (h.aux & 0x40u) != 0;
    // The scan counted s.cnt marked rows with start <= pos < end from row_lb on, all inside the read's site block:
    // the walk ends when they are written, so the block's end is never consulted (n_rows only bounds the prefetch)
    const int64_t n_rows = sites.n_rows;
    int64_t row = s.row_lb;                                // first site row with pos >= start (from read_scan)
    const uint32_t* cg = reads.cigar + h.cigar_off;
    const uint32_t cg0 = simple ? ((uint32_t)h.l_seq << 4) : (h.n_cigar > 0 ? __ldg(cg) : 0u);   // a simple read IS "l_seq M"
    int32_t p = row < n_rows ? __ldg(sites.pos + row) : 0x7fffffff;
    int mp0 = __ldg(mark_prefix + min(row, n_rows));
    int mp1 = __ldg(mark_prefix + min(row + 1, n_rows));
    const int64_t q0 = read_qoff(h);
    const int64_t hbase = (int64_t)__ldg(tile_base + (uint32_t)r / (uint32_t)tile_reads) + s.hoff;
    int written = 0;
    while (written < s.cnt && p < s.end) {
        const int64_t nrow = row + 1;                      // requested now, used by the next iteration
        const int32_t pn = nrow < n_rows ? __ldg(sites.pos + nrow) : 0x7fffffff;
        const int mp2 = __ldg(mark_prefix + min(nrow + 1, n_rows));
        if (mp1 != mp0) {                                  // a marked row
            const int k = mp0 - s.fmark;
            uint32_t word = 0;
            const int q = cigar_qpos(cg, h.n_cigar, h.start, p, cg0);
            if (q >= 0 && q < 0xffff) {
                const int64_t g = q0 + q;
                const uint32_t lq = (__ldg(reads.lowq + (g >> 5)) >> (g & 31)) & 1u;
                const uint32_t nb = (h.aux & 0x80u) ? ((__ldg(reads.nmask + (g >> 5)) >> (g & 31)) & 1u) : 0u;
                const uint32_t code = (__ldg(reads.seq2 + (g >> 2)) >> ((g & 3) << 1)) & 3u;
                word = (uint32_t)(q + 1) | (lq << 16) | (nb << 23) | (code << 24) | ((q + 1 < h.l_seq) ? (1u << 26) : 0u);
            }
            if (k >= 0 && k < s.cnt) hits[hbase + k] = word;
            ++written;
        }
        row = nrow; p = pn; mp0 = mp1; mp1 = mp2;
    }
}

// One warp per LK_TILES consecutive 32-read tiles.  Only one read in six overlaps a marked site: a thread per read
// leaves five lanes of six idle through a chain of dependent gathers.  Here the warp first reads the tiles' "has hits"
// masks (8 bytes per tile, written by the scan), then deals the reads that do have hits to its lanes densely -- the
// j-th such read of the group goes to lane j mod 32 -- so every lane of an iteration carries a gather chain.
#ifndef LK_TILES_N
#define LK_TILES_N 16
#endif
constexpr int LK_TILES = LK_TILES_N;
// the lookup is a chain of dependent gathers per read: it lives on resident warps.  8 CTAs x 256 threads per SM (32
// registers, a few spilled words) beat 6 x 40 registers without spills (measured with the single-M shortcut build: 0.153 vs 0.173 ms)
#ifndef LK_MINB
#define LK_MINB 8
#endif

__global__ void __launch_bounds__(256, LK_MINB)
read_site_alleles_kernel(UnfzReadCols reads, UnfzSiteCols sites, const uint8_t* __restrict__ row_mark,
                         const int32_t* __restrict__ mark_prefix, const UnfzReadSum* __restrict__ rsum,
                         const uint32_t* __restrict__ tile_base, int32_t tile_reads,
                         uint32_t* __restrict__ hits, const uint2* __restrict__ tile_info, const int32_t* __restrict__ guard) {
    UNFZ_GUARD(guard);
    (void)row_mark;
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_tiles = (reads.n_reads + 31) >> 5;
    const int64_t t0 = w * LK_TILES;
    if (t0 >= n_tiles) return;
    uint32_t mask = 0;
    if (lane < LK_TILES && t0 + lane < n_tiles) {
        const uint2 ti = __ldg(tile_info + t0 + lane);
        mask = ti.x;
        if (((t0 + lane) << 5) + 32 > reads.n_reads) mask &= (1u << (reads.n_reads - ((t0 + lane) << 5))) - 1u;   // ragged last tile
    }
    // exclusive prefix of the populations over the group's tiles (lanes 0..LK_TILES-1)
    const int c = __popc(mask);
    int incl = c;
#pragma unroll
    for (int o = 1; o < LK_TILES; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
    const int total = __shfl_sync(0xffffffffu, incl, LK_TILES - 1);
    for (int j0 = 0; j0 < total; j0 += 32) {
        const int j = j0 + lane;
        // tile of the j-th read with hits: the first tile whose inclusive prefix exceeds j
        int t = 0;
#pragma unroll
        for (int q = 0; q < LK_TILES; ++q) t += (__shfl_sync(0xffffffffu, incl, q) <= j) ? 1 : 0;
        const int tt = min(t, LK_TILES - 1);
        const uint32_t m = __shfl_sync(0xffffffffu, mask, tt);
        const int before = __shfl_sync(0xffffffffu, incl - c, tt);
        if (j < total) {
            const int bit = __fns(m, 0, j - before + 1);           // the (j - before)-th set bit of the tile's mask
            read_alleles_one(reads, sites, mark_prefix, rsum, tile_base, tile_reads, hits, ((t0 + tt) << 5) + bit);
        }
    }
}

}  // namespace

// bytes of one quality-bit slice of the streaming kernel: 32 reads + 16-byte alignment slop on either side
static int wp_slice_bytes(int32_t max_l_seq) { return (((32 * max_l_seq + 7) / 8 + 15) & ~15) + 32; }

// hit slots are numbered per tile of this many consecutive reads (tile_tot / tile_base granularity)
extern "C" int32_t unfz_read_scan_tile_reads(int32_t max_l_seq) { (void)max_l_seq; return 32; }

extern "C" int unfz_read_scan(UnfzCtx* ctx, const UnfzReadCols* reads, const UnfzSiteCols* sites,
                              const int32_t* mark_prefix, const UnfzParams* hp, int32_t max_l_seq, UnfzReadSum* out,
                              int32_t* blk_maxspan, uint32_t* tile_tot, uint32_t* tile_info, void* stream) {
    if (reads->n_reads <= 0) return 0;
    ScanParams P;
    P.min_mapq = hp->min_map_qual;
    P.readlen = hp->readlen;
    // reads too long for a 16 KB slice (> 4000 bases) are counted out of global memory (qslice 0: nothing staged)
    int qslice = max_l_seq > 0 ? wp_slice_bytes(max_l_seq) : 0;
    if (qslice > 16 * 1024) qslice = 0;
    const size_t smem2 = (size_t)WP_WARPS * (2 * (size_t)qslice + WP_FIXED);
    // the opt-in shared-memory size is an attribute of the function ON ONE DEVICE: cached per device
    if (ctx->scan_smem_attr < smem2) {
        UNFZ_CHECK(ctx, cudaFuncSetAttribute(read_scan_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        ctx->scan_smem_attr = smem2;
    }
    int per_sm = (int)((size_t)(228 * 1024) / (smem2 + 1024));
    if (per_sm > WP_MINB) per_sm = WP_MINB;
    if (per_sm < 1) per_sm = 1;
    const int64_t tiles = (reads->n_reads + 31) / 32;
    int64_t g = (int64_t)ctx->sm_count * per_sm;
    if (g * WP_WARPS > tiles) g = (tiles + WP_WARPS - 1) / WP_WARPS;
    read_scan_warp_kernel<<<(unsigned)g, WP_THREADS, smem2, (cudaStream_t)stream>>>(*reads, *sites, mark_prefix, P, qslice,
                                                                                 out, blk_maxspan, tile_tot,
                                                                                 reinterpret_cast<uint2*>(tile_info), ctx->guard);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// upload completion: sparse list of non-ACGT bases -> bit plane + per-read flag
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void expand_nlist_kernel(UnfzReadCols reads, UnfzRead* __restrict__ hdr, uint32_t* __restrict__ nmask,
                                    const int64_t* __restrict__ nidx, int64_t n_idx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_idx) return;
    const int64_t g = nidx[i];
    if (g < 0 || g >= reads.n_qual) return;
    atomicOr(nmask + (g >> 5), 1u << (g & 31));
    // owner: the last non-empty read whose first base is at or before g (offsets never decrease with the read index)
    int64_t lo = 0, hi = reads.n_reads;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (read_qoff(hdr[mid]) <= g) lo = mid + 1; else hi = mid;
    }
    int64_t r = lo - 1;
    while (r >= 0 && hdr[r].l_seq == 0) --r;
    if (r < 0 || read_qoff(hdr[r]) + hdr[r].l_seq <= g) return;     // a base index between two reads: nobody owns it
    // aux is byte 29 of the 32-byte header: set bit7 with a word-wide atomic (the neighbours are mapq, qoff_hi, pad)
    unsigned* w = reinterpret_cast<unsigned*>(hdr + r) + 7;
    atomicOr(w, 0x80u << 8);
}
}  // namespace

namespace {
__global__ void read_starts_kernel(UnfzRead* __restrict__ hdr, const uint32_t* __restrict__ cigar, int64_t n_cigar_words,
                                   int64_t n, int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const UnfzRead h = hdr[i];
    out[i] = h.start;
    // one M or = operation covering every base: mark the read (aux bit6) -- query index q is at start + q.  The two
    // device-only bits are (re)written here, whatever the host table carried; unfz_expand_nlist runs after this kernel
    bool simple = false;
    if (h.n_cigar == 1 && h.l_seq > 0 && (int64_t)h.cigar_off < n_cigar_words) {
        const uint32_t w = cigar[h.cigar_off];
        const uint32_t op = w & 15u;
        simple = (op == 0u || op == 7u) && (w >> 4) == (uint32_t)h.l_seq;
    }
    unsigned* wp = reinterpret_cast<unsigned*>(hdr + i) + 7;        // mapq | aux << 8 | qoff_hi << 16 | pad << 24
    *wp = (*wp & ~(0xc0u << 8)) | (simple ? (0x40u << 8) : 0u);
}
}  // namespace

extern "C" int unfz_read_starts(UnfzCtx* ctx, const UnfzReadCols* reads, int32_t* start_rw, void* stream) {
    if (reads->n_reads <= 0) return 0;
    read_starts_kernel<<<(unsigned)((reads->n_reads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        const_cast<UnfzRead*>(reads->hdr), reads->cigar, reads->n_cigar, reads->n_reads, start_rw);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int unfz_expand_nlist(UnfzCtx* ctx, const UnfzReadCols* reads, UnfzRead* hdr_rw, uint32_t* nmask_rw,
                                 const int64_t* nidx, int64_t n_idx, void* stream) {
    if (n_idx <= 0 || reads->n_reads <= 0) return 0;
    expand_nlist_kernel<<<(unsigned)((n_idx + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*reads, hdr_rw, nmask_rw, nidx, n_idx);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int unfz_read_site_alleles(UnfzCtx* ctx, const UnfzReadCols* reads, const UnfzSiteCols* sites,
                                      const uint8_t* row_mark, const int32_t* mark_prefix,
                                      const UnfzReadSum* rsum, const uint32_t* tile_base,
                                      int32_t tile_reads, uint32_t* hits, const uint32_t* tile_info, void* stream) {
    if (reads->n_reads <= 0) return 0;
    if (tile_reads != 32 || tile_info == nullptr) return unfz_fail(ctx, -21, "unfz_read_site_alleles needs the tile_info of unfz_read_scan (32-read tiles)");
    const int64_t tiles = (reads->n_reads + 31) / 32;
    const int64_t warps = (tiles + LK_TILES - 1) / LK_TILES;
    const int64_t blocks = (warps + 7) / 8;                 // 8 warps per CTA
    read_site_alleles_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*reads, *sites, row_mark, mark_prefix, rsum,
                                                                                 tile_base, tile_reads, hits,
                                                                                 reinterpret_cast<const uint2*>(tile_info), ctx->guard);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}
