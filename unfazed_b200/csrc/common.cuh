// Shared helpers for the sm_100a kernels of libunfazed_sm100.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/unfazed_sm100.h"

struct UnfzGraphSlot {
    uint64_t key;             // hash of everything the captured launches depend on; 0 = empty
    cudaGraphExec_t exec;
    uint64_t tick;            // last use, for eviction
};

struct UnfzCtx {
    int device;
    int sm_count;
    const int32_t* guard;     // device flag of the speculative-sizing mode (unfz_ctx_set_guard), or null
    bool chain_carveout_set;  // shared-memory carve-out preference of the chaining kernel set on THIS device
    size_t scan_smem_attr;    // opt-in dynamic shared memory already granted to the read scan on THIS device
    void* cls_params;         // classifier thresholds + allele-balance interval table of cls_key (sites.cu), malloc'ed
    uint32_t* cls_tab_dev;    // the same table on the device (cudaMalloc'ed with the first classification)
    double cls_key[8];        // ab_homref, ab_het, ab_homalt, min_gt_qual, min_depth the table was made for
    UnfzGraphSlot graphs[8];  // instantiated batch graphs (unfz_run_batch_graph)
    uint64_t graph_tick;
    char err[512];
};

// every kernel that writes into speculatively sized buffers starts with this
#define UNFZ_GUARD(g) do { if ((g) != nullptr && *(g) != 0) return; } while (0)

#define UNFZ_CHECK(ctx, expr)                                                                   \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #expr, \
                     cudaGetErrorString(_e));                                                   \
            return -(int)_e - 1000;                                                             \
        }                                                                                       \
    } while (0)

#define UNFZ_LAUNCH_CHECK(ctx) UNFZ_CHECK(ctx, cudaGetLastError())

// internal: makes sure the classifier's interval table for these thresholds is on the device (synchronous copy when the
// thresholds change: must run outside stream capture -- unfz_run_batch_graph calls it before it starts capturing)
extern "C" int unfz_classify_prepare(UnfzCtx* ctx, const UnfzParams* hp);

static inline int unfz_fail(UnfzCtx* ctx, int code, const char* msg) {
    snprintf(ctx->err, sizeof(ctx->err), "%s", msg);
    return code;
}

// first index i in [lo, hi) with a[i] >= v
template <typename T, typename V>
__device__ __forceinline__ int64_t lower_bound_dev(const T* __restrict__ a, int64_t lo, int64_t hi, V v) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first index i in [lo, hi) with a[i] > v
template <typename T, typename V>
__device__ __forceinline__ int64_t upper_bound_dev(const T* __restrict__ a, int64_t lo, int64_t hi, V v) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ int64_t read_qoff(const UnfzRead& h) {
    return (int64_t)h.qoff_lo | ((int64_t)h.qoff_hi << 32);
}

// 16-byte header halves through the read-only path
__device__ __forceinline__ UnfzRead load_read(const UnfzRead* __restrict__ p) {
    union { UnfzRead r; int4 v[2]; } u;
    const int4* q = reinterpret_cast<const int4*>(p);
    u.v[0] = __ldg(q);
    u.v[1] = __ldg(q + 1);
    return u.r;
}

__device__ __forceinline__ UnfzReadSum load_rsum(const UnfzReadSum* __restrict__ p) {
    union { UnfzReadSum r; int4 v[2]; } u;
    const int4* q = reinterpret_cast<const int4*>(p);
    u.v[0] = __ldg(q);
    u.v[1] = __ldg(q + 1);
    return u.r;
}
