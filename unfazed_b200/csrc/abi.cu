// Context management of libunfazed_sm100.so.
#include "common.cuh"

extern "C" int unfz_abi_version(void) { return UNFZ_ABI_VERSION; }

extern "C" int unfz_ctx_create(int device, UnfzCtx** out) {
    if (!out) return -1;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return -2;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -3;
    if (prop.major < 10) return -4;   // sm_100a code only
    if (cudaSetDevice(device) != cudaSuccess) return -5;
    UnfzCtx* c = new UnfzCtx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->guard = nullptr;
    c->err[0] = 0;
    *out = c;
    return 0;
}

extern "C" void unfz_ctx_destroy(UnfzCtx* ctx) { delete ctx; }

extern "C" const char* unfz_last_error(UnfzCtx* ctx) { return ctx ? ctx->err : "null context"; }

// ------------------------------------------------------------------------------------------------
// Speculative sizing.  The sizes of the variable outputs (pairs; hits; chaining scratch) are only
// known on the device.  A caller that has run a similar batch before can allocate from those
// capacities, let this kernel compare the real totals with them and launch everything without a
// host round trip: if a capacity is exceeded the flag is raised, every guarded kernel of the context
// returns at once, and the caller re-runs the batch with exact sizes.
// ------------------------------------------------------------------------------------------------
namespace {
struct CapArgs {
    const int64_t* totals[8];
    int64_t caps[8];
};
__global__ void check_caps_kernel(CapArgs a, int k, int32_t* flag, int64_t* actual) {
    const int i = threadIdx.x;
    if (i < k) {
        const int64_t v = *a.totals[i];
        actual[i] = v;
        if (v > a.caps[i]) atomicOr(flag, 1);
    }
}
}  // namespace

extern "C" int unfz_ctx_set_guard(UnfzCtx* ctx, const int32_t* flag) {
    if (!ctx) return -1;
    ctx->guard = flag;
    return 0;
}

extern "C" int unfz_check_caps(UnfzCtx* ctx, int32_t k, const int64_t* const* totals, const int64_t* caps,
                               int32_t* flag, int64_t* actual, void* stream) {
    if (k < 0 || k > 8) return unfz_fail(ctx, -30, "unfz_check_caps: at most 8 totals");
    if (k == 0) return 0;
    CapArgs a;
    for (int i = 0; i < 8; ++i) { a.totals[i] = i < k ? totals[i] : nullptr; a.caps[i] = i < k ? caps[i] : 0; }
    check_caps_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a, k, flag, actual);
    UNFZ_LAUNCH_CHECK(ctx);
    return 0;
}
