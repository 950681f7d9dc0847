// estimate_concordant_insert_len (read_collector.py:11-25) on the device.
//
// The reference takes np.percentile(|tlen - 2*readlen|, 99.5) over the first reads of the BAM; numpy
// interpolates linearly between two ORDER STATISTICS.  Those two integers are selected exactly here
// with a two-level radix histogram (high 16 bits, then low 16 bits inside the selected bins) -- no
// sort -- and the host applies numpy's own interpolation formula to them, so the result is
// bit-identical to the reference's.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t insert_value(const UnfzRead* __restrict__ hdr, int64_t r, int32_t readlen) {
    long long v = (long long)__ldg(&hdr[r].tlen) - 2ll * readlen;
    if (v < 0) v = -v;
    return (uint32_t)v;                                   // |int32 - small| always fits 32 bits
}

__global__ void isz_hist_hi(const UnfzRead* __restrict__ hdr, int64_t first, int64_t n, int32_t readlen, uint32_t* __restrict__ hist) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[insert_value(hdr, first + i, readlen) >> 16], 1u);
}

constexpr int ISZ_Q = 4;        // order statistics selected per call

// single block: for each requested rank, the bin that holds it and the rank inside that bin.
// pass 0 reads the shared high-bits histogram, pass 1 the per-rank low-bits histograms.
__global__ void __launch_bounds__(1024)
isz_pick(const uint32_t* hist, const unsigned long long* ranks_in, int which_pass, uint32_t* pick) {
    __shared__ unsigned long long part[1024];
    for (int q = 0; q < ISZ_Q; ++q) {
        const uint32_t* h = hist + (size_t)q * 65536 * (which_pass ? 1 : 0);
        const unsigned long long rank = which_pass ? (unsigned long long)pick[2 * q + 1] : ranks_in[q];
        unsigned long long s = 0;
        const int b0 = threadIdx.x * 64;
        for (int b = 0; b < 64; ++b) s += h[b0 + b];
        part[threadIdx.x] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long run = 0;
            int t = 0;
            for (; t < 1023; ++t) { if (run + part[t] > rank) break; run += part[t]; }
            int b = t * 64;
            for (; b < t * 64 + 63; ++b) { if (run + h[b] > rank) break; run += h[b]; }
            pick[2 * ISZ_Q * which_pass + 2 * q] = (uint32_t)b;                    // bin
            pick[2 * ISZ_Q * which_pass + 2 * q + 1] = (uint32_t)(rank - run);     // rank inside the bin
        }
        __syncthreads();
    }
}

__global__ void isz_hist_lo(const UnfzRead* __restrict__ hdr, int64_t first, int64_t n, int32_t readlen,
                            const uint32_t* __restrict__ pick, uint32_t* __restrict__ hist_lo) {
    uint32_t bin[ISZ_Q];
#pragma unroll
    for (int q = 0; q < ISZ_Q; ++q) bin[q] = pick[2 * q];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = insert_value(hdr, first + i, readlen);
#pragma unroll
        for (int q = 0; q < ISZ_Q; ++q)
            if ((v >> 16) == bin[q]) atomicAdd(&hist_lo[(size_t)q * 65536 + (v & 0xffffu)], 1u);
    }
}

__global__ void isz_finish(const uint32_t* __restrict__ pick, uint32_t* __restrict__ out) {
    for (int q = 0; q < ISZ_Q; ++q) out[q] = (pick[2 * q] << 16) | pick[2 * ISZ_Q + 2 * q];
}

}  // namespace

extern "C" int64_t unfz_insert_size_work_bytes(void) { return (int64_t)(1 + ISZ_Q) * 65536 * 4 + 4 * ISZ_Q * 4 + ISZ_Q * 8 + 64; }

// Four order statistics (0-based ranks, HOST array h_ranks[4], clamped by the caller) of
// |tlen - 2*readlen| over the reads of `n_ranges` index ranges [first, first+count) (host arrays).
// work: unfz_insert_size_work_bytes() bytes, zeroed by the caller.  out[0..3] (device uint32).
extern "C" int unfz_insert_size_order_stats(UnfzCtx* ctx, const UnfzReadCols* reads, const int64_t* h_first,
                                            const int64_t* h_count, int32_t n_ranges, int32_t readlen,
                                            const uint64_t* h_ranks, void* work, uint32_t* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t* hist_hi = reinterpret_cast<uint32_t*>(work);
    uint32_t* hist_lo = hist_hi + 65536;
    uint32_t* pick = hist_lo + (size_t)ISZ_Q * 65536;
    unsigned long long* ranks = reinterpret_cast<unsigned long long*>(pick + 4 * ISZ_Q);
    UNFZ_CHECK(ctx, cudaMemcpyAsync(ranks, h_ranks, sizeof(unsigned long long) * ISZ_Q, cudaMemcpyHostToDevice, st));
    for (int i = 0; i < n_ranges; ++i) {
        if (h_count[i] <= 0) continue;
        const int grid = (int)((h_count[i] + 255) / 256 > 1184 ? 1184 : (h_count[i] + 255) / 256);
        isz_hist_hi<<<grid, 256, 0, st>>>(reads->hdr, h_first[i], h_count[i], readlen, hist_hi);
    }
    isz_pick<<<1, 1024, 0, st>>>(hist_hi, ranks, 0, pick);
    for (int i = 0; i < n_ranges; ++i) {
        if (h_count[i] <= 0) continue;
        const int grid = (int)((h_count[i] + 255) / 256 > 1184 ? 1184 : (h_count[i] + 255) / 256);
        isz_hist_lo<<<grid, 256, 0, st>>>(reads->hdr, h_first[i], h_count[i], readlen, pick, hist_lo);
    }
    isz_pick<<<1, 1024, 0, st>>>(hist_lo, ranks, 1, pick);
    isz_finish<<<1, 1, 0, st>>>(pick, out);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}
