// Context management of libunfazed_sm100.so.
#include <stdlib.h>

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
    c->scan_smem_attr = 0;
    c->chain_carveout_set = false;
    c->cls_params = nullptr;
    c->cls_tab_dev = nullptr;
    memset(c->graphs, 0, sizeof(c->graphs));
    c->graph_tick = 0;
    c->err[0] = 0;
    *out = c;
    return 0;
}

extern "C" void unfz_ctx_destroy(UnfzCtx* ctx) {
    if (!ctx) return;
    for (auto& g : ctx->graphs)
        if (g.key) cudaGraphExecDestroy(g.exec);
    free(ctx->cls_params);
    if (ctx->cls_tab_dev) cudaFree(ctx->cls_tab_dev);
    delete ctx;
}

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

// ------------------------------------------------------------------------------------------------
// One batch in one call (speculatively sized buffers): the body of Engine.run without the per-call
// cost of the host language.
// ------------------------------------------------------------------------------------------------
#define UNFZ_RC(expr) do { const int rc_ = (expr); if (rc_ != 0) { ctx->guard = nullptr; return rc_; } } while (0)

extern "C" int unfz_batch_struct_bytes(void) { return (int)sizeof(UnfzBatch); }

extern "C" int unfz_run_batch(UnfzCtx* ctx, const UnfzBatch* b, void* s) {
    if (!ctx || !b) return -1;
    const int32_t n = b->n_dnms, S = b->n_segs;
    const int64_t V = b->sites->n_rows;
    ctx->guard = b->guard;
    UNFZ_RC(unfz_window_search(ctx, b->sites, b->segs, S, b->seg_row_lo, b->seg_count, s));
    UNFZ_RC(unfz_exclusive_scan_i64(ctx, b->seg_count, b->seg_pair_off, S, b->scan_work, s));
    {
        const int64_t* t[1] = {b->seg_pair_off + S};
        const int64_t c[1] = {b->cap_pairs};
        UNFZ_RC(unfz_check_caps(ctx, 1, t, c, b->guard, b->actual, s));
    }
    UNFZ_RC(unfz_classify_sites(ctx, b->sites, b->segs, b->seg_row_lo, b->seg_pair_off, S, b->cap_pairs, b->h_params, b->cls, s));
    UNFZ_RC(unfz_compact_sites(ctx, b->dnms, n, b->segs, b->seg_row_lo, b->seg_pair_off, b->cls, b->het_list, b->n_het,
                               b->cand_list, b->n_cand, b->cnv_dad, b->cnv_mom, b->row_mark, s));
    UNFZ_RC(unfz_exclusive_scan_u8_i32(ctx, b->row_mark, b->mark_prefix, V, b->scan_work, s));
    if (b->reads != nullptr) {
        int64_t* total_hits = b->off + 6 * ((int64_t)n + 1);
        UNFZ_RC(unfz_read_scan(ctx, b->reads, b->sites, b->mark_prefix, b->h_params, b->max_l_seq, b->rsum,
                               b->blk_maxspan, b->tile_tot, b->tile_info, s));
        UNFZ_RC(unfz_exclusive_scan_u32(ctx, b->tile_tot, b->tile_base, b->n_tiles, total_hits, b->scan_work, s));
        UNFZ_RC(unfz_chain_size(ctx, b->dnms, n, b->segs, b->seg_pair_off, b->sites, b->reads, b->rsum, b->blk_maxspan,
                                b->het_list, b->n_het, b->cand_list, b->n_cand, b->win, b->need, b->site_lo, b->site_n,
                                b->seed_win, s));
        UNFZ_RC(unfz_exclusive_scan_rows_i64(ctx, b->need, b->off, 6, n, s));
        {
            const int64_t* t[7];
            int64_t c[7];
            for (int i = 0; i < 6; ++i) { t[i] = b->off + (int64_t)(i + 1) * (n + 1) - 1; c[i] = b->cap_chain[i]; }
            t[6] = total_hits;
            c[6] = b->cap_hits;
            UNFZ_RC(unfz_check_caps(ctx, 7, t, c, b->guard, b->actual + 1, s));
        }
        UNFZ_RC(unfz_read_site_alleles(ctx, b->reads, b->sites, b->row_mark, b->mark_prefix, b->rsum, b->tile_base,
                                       b->tile_reads, b->hits, b->tile_info, s));
        UNFZ_RC(unfz_chain_tally(ctx, b->dnms, n, b->segs, b->seg_pair_off, b->sites, b->reads, b->rsum, b->blk_maxspan,
                                 b->hits, b->tile_base, b->tile_reads, b->mark_prefix, b->het_list, b->n_het, b->cand_list,
                                 b->n_cand, b->alleles, b->win, b->site_lo, b->site_n, b->seed_win, b->off, b->cap_chain,
                                 b->h_params, b->scratch, b->scratch_bytes, b->slot_label, b->slot_evid, b->cand_evid,
                                 b->tally, b->ev_need, s));
        if (b->ev_need != nullptr) {
            UNFZ_RC(unfz_exclusive_scan_rows_i64(ctx, b->ev_need, b->ev_off, 4, n, s));
            UNFZ_RC(unfz_evidence_lists(ctx, b->dnms, n, b->seg_pair_off, b->sites, b->cand_list, b->n_cand, b->cand_evid,
                                        b->win, b->off, b->slot_evid, b->ev_off, b->ev_read_dad, b->ev_read_mom, b->ev_pos_dad,
                                        b->ev_pos_mom, s));
        }
    }
    UNFZ_RC(unfz_summarize(ctx, b->dnms, n, b->tally, b->cnv_dad, b->cnv_mom, b->n_cand, b->h_params, b->calls_strict,
                           b->calls_ambiguous, s));
    ctx->guard = nullptr;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// The same batch as ONE CUDA graph: the zero-fills of the caller's arenas, the ~25 launches of
// unfz_run_batch and the device-to-host copy of the result block are captured once per buffer layout
// and replayed with a single cudaGraphLaunch -- the per-launch gaps of the ten small kernels in front
// of the read scan disappear, and nothing but this library's kernels runs on the stream.
// The graph is keyed by a hash of every value the captured launches depend on (the batch struct, the
// column structs and parameters behind its host pointers, the spans to clear, the copy); a caller that
// recycles its buffers -- Engine.run does -- hits the cache from the third batch on.
// `stream` must not be the legacy default stream (CUDA cannot capture it).
// ------------------------------------------------------------------------------------------------
static uint64_t fnv1a(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

extern "C" int unfz_run_batch_graph(UnfzCtx* ctx, const UnfzBatch* b, const UnfzSpan* zero, int32_t n_zero,
                                    void* h_dst, const void* d_src, int64_t d2h_bytes, void* stream) {
    if (!ctx || !b) return -1;
    cudaStream_t s = (cudaStream_t)stream;
    uint64_t key = 1469598103934665603ull;
    key = fnv1a(key, b, sizeof(*b));
    key = fnv1a(key, b->sites, sizeof(*b->sites));
    if (b->reads) key = fnv1a(key, b->reads, sizeof(*b->reads));
    key = fnv1a(key, b->h_params, sizeof(*b->h_params));
    key = fnv1a(key, zero, sizeof(UnfzSpan) * (size_t)(n_zero > 0 ? n_zero : 0));
    key = fnv1a(key, &h_dst, sizeof(h_dst));
    key = fnv1a(key, &d_src, sizeof(d_src));
    key = fnv1a(key, &d2h_bytes, sizeof(d2h_bytes));
    if (key == 0) key = 1;
    UnfzGraphSlot* slot = nullptr;
    UnfzGraphSlot* victim = nullptr;                  // an empty slot if there is one, else the least recently used
    for (auto& g : ctx->graphs) {
        if (g.key == key) { slot = &g; break; }
        if (!victim || (victim->key != 0 && (g.key == 0 || g.tick < victim->tick))) victim = &g;
    }
    if (!slot) {
        cudaGraph_t graph = nullptr;
        { const int rc0 = unfz_classify_prepare(ctx, b->h_params); if (rc0) return rc0; }   // a synchronous copy: not under capture
        UNFZ_CHECK(ctx, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        int rc = 0;
        for (int i = 0; i < n_zero && rc == 0; ++i)
            if (zero[i].bytes > 0 && cudaMemsetAsync(zero[i].ptr, 0, (size_t)zero[i].bytes, s) != cudaSuccess) rc = -40;
        if (rc == 0) rc = unfz_run_batch(ctx, b, s);
        if (rc == 0 && d2h_bytes > 0 && cudaMemcpyAsync(h_dst, d_src, (size_t)d2h_bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess) rc = -41;
        const cudaError_t e = cudaStreamEndCapture(s, &graph);
        if (rc != 0 || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return rc != 0 ? rc : unfz_fail(ctx, -42, "unfz_run_batch_graph: stream capture failed (legacy default stream?)");
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) return unfz_fail(ctx, -43, cudaGetErrorString(ei));
        if (victim->key) cudaGraphExecDestroy(victim->exec);
        victim->key = key;
        victim->exec = exec;
        slot = victim;
    }
    slot->tick = ++ctx->graph_tick;
    UNFZ_CHECK(ctx, cudaGraphLaunch(slot->exec, s));
    return 0;
}
