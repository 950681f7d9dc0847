// Device-wide exclusive prefix sums used to turn per-item counts into CSR offsets.
// Three-phase scan (tile reduce -> recursive scan of tile sums -> tile scan), hand written; the
// inputs here are tiny next to the streaming kernels so a decoupled look-back is not needed.
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename TIn>
__device__ __forceinline__ int64_t load_strided(const TIn* in, int64_t stride_bytes, int64_t i) {
    return (int64_t)*reinterpret_cast<const TIn*>(reinterpret_cast<const char*>(in) + i * stride_bytes);
}

__device__ __forceinline__ int64_t block_exclusive_scan(int64_t v, int64_t* total) {
    __shared__ int64_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        int64_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            int64_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    int64_t base = warp ? warp_sums[warp - 1] : 0;
    *total = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return base + x - v;
}

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS)
tile_sums_kernel(const TIn* __restrict__ in, int64_t stride, int64_t n, int64_t* __restrict__ sums) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += load_strided(in, stride, base + k);
    int64_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS)
tile_scan_kernel(const TIn* __restrict__ in, int64_t stride, int64_t n, const int64_t* __restrict__ tile_off,
                 TOut* __restrict__ out, int64_t out_stride, int write_total, int64_t* __restrict__ total_out) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? load_strided(in, stride, base + k) : 0;
        s += v[k];
    }
    int64_t total;
    int64_t run = block_exclusive_scan(s, &total) + (tile_off ? tile_off[blockIdx.x] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n)
            *reinterpret_cast<TOut*>(reinterpret_cast<char*>(out) + (base + k) * out_stride) = (TOut)run;
        run += v[k];
    }
    // out[n] = grand total (CSR convention) -- written by the thread that owns item n-1
    if (n > 0 && base <= n - 1 && n - 1 < base + SCAN_ITEMS) {
        if (write_total) *reinterpret_cast<TOut*>(reinterpret_cast<char*>(out) + n * out_stride) = (TOut)run;
        if (total_out) *total_out = run;
    }
}

__global__ void zero_total_kernel(void* out, int bytes, int64_t* total_out) {
    if (out && threadIdx.x < bytes) reinterpret_cast<char*>(out)[threadIdx.x] = 0;
    if (total_out && threadIdx.x == 0) *total_out = 0;
}

template <typename TIn, typename TOut>
int scan_impl(UnfzCtx* ctx, const TIn* in, int64_t in_stride, TOut* out, int64_t out_stride, int64_t n,
              void* work, cudaStream_t st, int write_total, int64_t* total_out) {
    if (n <= 0) {
        if (write_total || total_out) {
            zero_total_kernel<<<1, 32, 0, st>>>(write_total ? (void*)out : nullptr, (int)sizeof(TOut), total_out);
            UNFZ_LAUNCH_CHECK(ctx);
        }
        return 0;
    }
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t* sums = reinterpret_cast<int64_t*>(work);
    if (tiles == 1) {
        tile_scan_kernel<TIn, TOut><<<1, SCAN_THREADS, 0, st>>>(in, in_stride, n, nullptr, out, out_stride, write_total, total_out);
        UNFZ_LAUNCH_CHECK(ctx);
        return 0;
    }
    tile_sums_kernel<TIn><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, in_stride, n, sums);
    UNFZ_LAUNCH_CHECK(ctx);
    // recursive exclusive scan of the tile sums, in place after the sums array
    int64_t* sums_scanned = sums + tiles;
    int rc = scan_impl<int64_t, int64_t>(ctx, sums, 8, sums_scanned, 8, tiles, sums_scanned + tiles + 1, st, 0, nullptr);
    if (rc) return rc;
    tile_scan_kernel<TIn, TOut><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, in_stride, n, sums_scanned, out, out_stride, write_total, total_out);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

// several independent rows in ONE launch: one CTA per row, sequential tiles with a running carry
__global__ void __launch_bounds__(SCAN_THREADS)
rows_scan_kernel(const int64_t* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
    const int64_t* src = in + (int64_t)blockIdx.x * n;
    int64_t* dst = out + (int64_t)blockIdx.x * (n + 1);
    int64_t carry = 0;
    for (int64_t t0 = 0; t0 < n; t0 += SCAN_TILE) {
        const int64_t base = t0 + (int64_t)threadIdx.x * SCAN_ITEMS;
        int64_t v[SCAN_ITEMS];
        int64_t s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            v[k] = (base + k < n) ? src[base + k] : 0;
            s += v[k];
        }
        int64_t total;
        int64_t run = carry + block_exclusive_scan(s, &total);
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (base + k < n) dst[base + k] = run;
            run += v[k];
        }
        carry += total;
    }
    if (threadIdx.x == 0) dst[n] = carry;
}

}  // namespace

extern "C" int unfz_exclusive_scan_rows_i64(UnfzCtx* ctx, const int64_t* in, int64_t* out, int32_t n_rows, int64_t n, void* stream) {
    if (n_rows <= 0) return 0;
    rows_scan_kernel<<<n_rows, SCAN_THREADS, 0, (cudaStream_t)stream>>>(in, out, n);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" int64_t unfz_scan_work_bytes(int64_t n) {
    // tile sums + scanned tile sums at every recursion level (geometric), generous bound
    int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t total = 0;
    while (tiles > 1) {
        total += 2 * tiles + 2;
        tiles = (tiles + SCAN_TILE - 1) / SCAN_TILE;
    }
    return (total + 16) * 8;
}

extern "C" int unfz_exclusive_scan_i64(UnfzCtx* ctx, const int64_t* in, int64_t* out, int64_t n, void* work, void* stream) {
    return scan_impl<int64_t, int64_t>(ctx, in, 8, out, 8, n, work, (cudaStream_t)stream, 1, nullptr);
}

extern "C" int unfz_exclusive_scan_u32(UnfzCtx* ctx, const uint32_t* in, uint32_t* out, int64_t n, int64_t* total_out,
                                       void* work, void* stream) {
    return scan_impl<uint32_t, uint32_t>(ctx, in, 4, out, 4, n, work, (cudaStream_t)stream, 1, total_out);
}

extern "C" int unfz_exclusive_scan_u8_i32(UnfzCtx* ctx, const uint8_t* in, int32_t* out, int64_t n, void* work, void* stream) {
    return scan_impl<uint8_t, int32_t>(ctx, in, 1, out, 4, n, work, (cudaStream_t)stream, 1, nullptr);
}
