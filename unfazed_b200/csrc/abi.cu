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
    c->err[0] = 0;
    *out = c;
    return 0;
}

extern "C" void unfz_ctx_destroy(UnfzCtx* ctx) { delete ctx; }

extern "C" const char* unfz_last_error(UnfzCtx* ctx) { return ctx ? ctx->err : "null context"; }
