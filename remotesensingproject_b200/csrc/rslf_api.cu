/*
 * rslf_api.cu — C ABI (include/rslf_b200.h) and host orchestration of the
 * B200-native EPI depth path: streams, device buffers, the s_hat pass loop and
 * the fine-to-coarse level loop.  No CPU fallback: every entry point needs a
 * CUDA device.
 */
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <thread>

#include "rslf_common.cuh"
#include "k_edge.cuh"
#include "k_depth.cuh"
#ifndef RSLF_DEPTH_TMEM_DEFAULT
#define RSLF_DEPTH_TMEM_DEFAULT 1
#endif
#include "k_depth_tm.cuh"
#include "k_median.cuh"
#include "k_propagate.cuh"
#include "k_pyramid.cuh"
#include "k_colour.cuh"
#include "rslf_comm.cuh"
#include "k_peak.cuh"

#define RSLF_ABI_VERSION 4
static const size_t RSLF_RING_BYTES = (size_t)64 << 20;   /* one buffer of the pinned transfer ring */
static const size_t RSLF_COPY_THREADS = 6;                /* host threads that fill / empty a ring buffer */
#define RSLF_COUNT_SLOTS 65536

/* ------------------------------------------------------------------ helpers */
template <typename T>
static int dev_alloc(rslf_ctx* ctx, T** p, size_t n)
{
    if (*p) { cudaFree(*p); *p = nullptr; }
    if (n == 0) return RSLF_OK;
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e));
        *p = nullptr;
        return e == cudaErrorMemoryAllocation ? RSLF_ERR_NOMEM : RSLF_ERR_CUDA;
    }
    return RSLF_OK;
}
template <typename T> static void dev_free(T** p) { if (*p) { cudaFree(*p); *p = nullptr; } }

static int stream_grid(const rslf_ctx* ctx, size_t n, int threads = 256)
{
    size_t b = (n + threads - 1) / threads;
    size_t cap = (size_t)ctx->num_sm * 16;
    return (int)std::max<size_t>(1, std::min(b, cap));
}

/* stage clock: CUDA events on the ctx stream, resolved after the final sync */
static void clk_reset(rslf_ctx* ctx) { ctx->clk.used = 0; ctx->clk.spans.clear(); }
static size_t clk_mark(rslf_ctx* ctx)
{
    rslf_stage_clock& c = ctx->clk;
    if (c.used == c.pool.size()) {
        cudaEvent_t e; cudaEventCreate(&e); c.pool.push_back(e);
    }
    cudaEventRecord(c.pool[c.used], ctx->stream);
    return c.used++;
}
struct stage_scope {
    rslf_ctx* ctx; int stage; size_t a; bool on;
    /* stage_timing 0: no events inside a run; 1 (default): only the dominant (depth) kernel is bracketed, two event
     * records per pass; 2: every stage (diagnostics: ~12 records per pass, which cost ~2 % of a one-GPU step and ~10 % of
     * an 8-GPU step, where a pass is a few hundred microseconds) */
    stage_scope(rslf_ctx* c, int st) : ctx(c), stage(st), a(0), on(c->stage_timing >= 2 || (c->stage_timing == 1 && st == ST_DEPTH)) { if (on) a = clk_mark(ctx); }
    ~stage_scope() { if (on) { size_t b = clk_mark(ctx); ctx->clk.spans.push_back({stage, a, b}); } }
};
static void clk_resolve(rslf_ctx* ctx)
{
    float acc[ST_COUNT] = {0};
    ctx->clk.span_ms.clear();
    for (auto& sp : ctx->clk.spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->clk.pool[sp.a], ctx->clk.pool[sp.b]) == cudaSuccess) acc[sp.stage] += ms;
        ctx->clk.span_ms.push_back(ms);
    }
    ctx->timing.ms_edge = acc[ST_EDGE]; ctx->timing.ms_depth = acc[ST_DEPTH]; ctx->timing.ms_reduce = acc[ST_REDUCE];
    ctx->timing.ms_median = acc[ST_MEDIAN]; ctx->timing.ms_propagate = acc[ST_PROP]; ctx->timing.ms_pyramid = acc[ST_PYR];
}

static void free_level(rslf_level& L, bool keep_raw_borrowed_ptr = false)
{
    (void)keep_raw_borrowed_ptr;
    dev_free(&L.raw); dev_free(&L.epi_full); L.epi = nullptr; dev_free(&L.ce); dev_free(&L.cd); dev_free(&L.depth); dev_free(&L.rbar);
    dev_free(&L.dmin); dev_free(&L.dmax); dev_free(&L.emask); dev_free(&L.remaining); dev_free(&L.valid); dev_free(&L.rowdark); dev_free(&L.dark_lo); dev_free(&L.dark_hi);
    L.cap_px = 0; L.cap_stack = 0; L.V = L.U = 0; L.C = 0; L.have_bounds = false;
}

static void free_scratch(rslf_ctx* ctx)
{
    dev_free(&ctx->items); dev_free(&ctx->filtered); dev_free(&ctx->winner); dev_free(&ctx->arrive);
    dev_free(&ctx->rec_own); ctx->rec = nullptr;
    if (ctx->partials) { cudaFree(ctx->partials); ctx->partials = nullptr; }
    ctx->partials_cap = 0;
    dev_free(&ctx->nearest_l); dev_free(&ctx->nearest_r);
    dev_free(&ctx->fuse_a); dev_free(&ctx->fuse_b); dev_free(&ctx->fuse_ma); dev_free(&ctx->fuse_mb);
    dev_free(&ctx->out_map); dev_free(&ctx->out_valid); dev_free(&ctx->pile_depth_raw);
    dev_free(&ctx->g_depth); dev_free(&ctx->g_colour); dev_free(&ctx->g_mask); ctx->g_plane_cap = 0;
    dev_free(&ctx->g_raw); ctx->g_raw_cap = 0;
    dev_free(&ctx->g_map); dev_free(&ctx->g_map2); dev_free(&ctx->g_mk); dev_free(&ctx->g_mk2); ctx->g_map_cap = 0;
    ctx->scratch_px = 0;
}

/* scratch shared by all levels, sized for level 0 */
static int ensure_scratch(rslf_ctx* ctx, bool need_2d, bool need_ftc, size_t plane_override = 0)
{
    /* the largest V x U plane any level of the run holds on this rank (a replicated coarse level may exceed
     * the rank's share of level 0) */
    const size_t plane = plane_override ? plane_override : (size_t)ctx->V * ctx->U;
    const size_t px = plane * ctx->S;
    if (ctx->scratch_px != px || ctx->scratch_plane != plane || ctx->scratch_world != ctx->world) {
        ctx->scratch_world = ctx->world;
        free_scratch(ctx);
        RSLF_TRY(dev_alloc(ctx, &ctx->items, plane));
        RSLF_TRY(dev_alloc(ctx, &ctx->rec_own, plane));
        RSLF_TRY(dev_alloc(ctx, &ctx->filtered, plane));
        /* a rank of a pass-balanced run evaluates 1/world of the pass's pixels of ALL ranks: its share can
         * exceed its own plane when the row blocks are unequal */
        ctx->share_cap = plane;
        if (ctx->world > 1) ctx->share_cap = std::max(plane, ((size_t)ctx->V_total * ctx->U + ctx->world - 1) / ctx->world + 64);
        RSLF_TRY(dev_alloc(ctx, &ctx->arrive, ctx->share_cap));
        RSLF_TRY(dev_alloc(ctx, &ctx->pile_depth_raw, plane));
        RSLF_CUDA_TRY(ctx, cudaMemsetAsync(ctx->arrive, 0, ctx->share_cap * sizeof(int), ctx->stream));
        ctx->scratch_px = px; ctx->scratch_plane = plane;
    }
    if (need_2d && !ctx->winner) {
        RSLF_TRY(dev_alloc(ctx, &ctx->winner, px));
        RSLF_CUDA_TRY(ctx, cudaMemsetAsync(ctx->winner, 0x7f, px * sizeof(int), ctx->stream));  /* >= any u */
    }
    if (need_ftc && !ctx->out_map) {
        RSLF_TRY(dev_alloc(ctx, &ctx->nearest_l, px));
        RSLF_TRY(dev_alloc(ctx, &ctx->nearest_r, px));
        RSLF_TRY(dev_alloc(ctx, &ctx->fuse_a, px));
        RSLF_TRY(dev_alloc(ctx, &ctx->fuse_b, px));
        RSLF_TRY(dev_alloc(ctx, &ctx->fuse_ma, px));
        RSLF_TRY(dev_alloc(ctx, &ctx->fuse_mb, px));
        RSLF_TRY(dev_alloc(ctx, &ctx->out_map, px));
        RSLF_TRY(dev_alloc(ctx, &ctx->out_valid, px));
    }
    return RSLF_OK;
}

static int ensure_partials(rslf_ctx* ctx, size_t n)
{
    if (ctx->partials_cap >= n) return RSLF_OK;
    if (ctx->partials) { cudaFree(ctx->partials); ctx->partials = nullptr; }
    cudaError_t e = cudaMalloc(&ctx->partials, n * sizeof(rslf_partial));
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(partials %zu): %s", n, cudaGetErrorString(e));
        return RSLF_ERR_NOMEM;
    }
    ctx->partials_cap = n;
    return RSLF_OK;
}

/* per-level maps; `full` = all S lines (2D), else one plane (pile).  V rows of maps from global row v0 on; the
 * EPI stack of the level holds all Vtot rows on every rank (Vtot < 0: V, a single-rank level). */
static int ensure_level(rslf_ctx* ctx, int p, int V, int U, bool full, bool with_bounds, int v0 = 0, int Vtot = -1)
{
    rslf_level& L = ctx->lv[p];
    if (Vtot < 0) { Vtot = V; v0 = 0; }
    const size_t planes = full ? (size_t)ctx->S : 1;
    const size_t px = planes * V * U;
    const size_t stack = (size_t)Vtot * ctx->S * U * ctx->C + 64;     /* + slack: TMA segments end on 16-byte boundaries */
    if (L.V != V || L.U != U || L.cap_px != px || L.C != ctx->C) {
        float* raw = L.raw; float* epi = L.epi_full; size_t cap_stack = L.cap_stack;   /* stacks survive a map re-allocation */
        L.raw = nullptr; L.epi_full = nullptr;
        free_level(L);
        L.raw = raw; L.epi_full = epi; L.cap_stack = cap_stack;
        L.V = V; L.U = U; L.C = ctx->C;
        RSLF_TRY(dev_alloc(ctx, &L.ce, px)); RSLF_TRY(dev_alloc(ctx, &L.cd, px));
        RSLF_TRY(dev_alloc(ctx, &L.depth, px)); RSLF_TRY(dev_alloc(ctx, &L.rbar, px * ctx->C));
        RSLF_TRY(dev_alloc(ctx, &L.emask, px)); RSLF_TRY(dev_alloc(ctx, &L.remaining, px));
        RSLF_TRY(dev_alloc(ctx, &L.valid, px));
        RSLF_TRY(dev_alloc(ctx, &L.rowdark, (planes + 1) * (size_t)V));
        RSLF_TRY(dev_alloc(ctx, &L.dark_lo, planes * (size_t)V)); RSLF_TRY(dev_alloc(ctx, &L.dark_hi, planes * (size_t)V));
        L.cap_px = px;
    }
    if (L.cap_stack < stack) {
        RSLF_TRY(dev_alloc(ctx, &L.epi_full, stack));
        if (p > 0) RSLF_TRY(dev_alloc(ctx, &L.raw, stack));
        L.cap_stack = stack;
    }
    L.v0 = v0; L.Vtot = Vtot;
    L.epi = L.epi_full + (size_t)v0 * ctx->S * U * ctx->C;
    if (with_bounds && !L.dmin) {
        RSLF_TRY(dev_alloc(ctx, &L.dmin, px)); RSLF_TRY(dev_alloc(ctx, &L.dmax, px));
    }
    return RSLF_OK;
}

static int check_params(rslf_ctx* ctx, const rslf_params* P, int dim_d)
{
    if (!ctx || !P) return RSLF_ERR_ARG;
    if (!ctx->have_input) { snprintf(ctx->err, sizeof(ctx->err), "no EPIs uploaded"); return RSLF_ERR_STATE; }
    if (dim_d < 2) { snprintf(ctx->err, sizeof(ctx->err), "dim_d must be >= 2"); return RSLF_ERR_ARG; }
    if (ctx->C != 1 && ctx->C != 3) { snprintf(ctx->err, sizeof(ctx->err), "C must be 1 or 3"); return RSLF_ERR_ARG; }
    return RSLF_OK;
}

/* kern.hpp:43: inv_m_h_sq = 1.0 / (h*h); the 1-channel kernel uses 3 * inv (kern.cpp:21) */
static float kernel_inv(const rslf_params& P, int C)
{
    float hh = P.kernel_h * P.kernel_h;
    float inv = (float)(1.0 / (double)hh);
    if (C == 1) inv = (float)(3 * inv);
    return inv;
}

/* bytes of one value of a raw stack */
static inline size_t depth_esz(int cv_depth) { return cv_depth == RSLF_DEPTH_8U ? 1 : cv_depth == RSLF_DEPTH_16U ? 2 : 4; }

/* Normalises the raw stack of level p (all Vtot rows: every rank holds the whole stack, so the stack maximum
 * needs no collective) into L.epi_full (ctor scaling, dc.hpp:442-477). */
static int normalise_level(rslf_ctx* ctx, int p, const void* raw, int cv_depth)
{
    rslf_level& L = ctx->lv[p];
    stage_scope sc(ctx, ST_PYR);
    const size_t n = (size_t)L.Vtot * ctx->S * L.U * ctx->C;
    if (p == 0 && ctx->pre_norm_epoch == ctx->input_epoch && ctx->world <= 1) {
        L.nonneg = ctx->pre_nonneg;                          /* normalised while it was uploaded (rslf_cuda_upload_epis_pipelined) */
        return RSLF_OK;
    }
    if (cv_depth == RSLF_DEPTH_8U) {
        normalise_u8_kernel<<<stream_grid(ctx, n), 256, 0, ctx->stream>>>((const uint8_t*)raw, n, L.epi_full);
        L.nonneg = 1;
        ctx->timing.kernel_launches += 1;
    } else if (cv_depth == RSLF_DEPTH_16U) {
        float sf = ctx->scale_factor;
        if (sf < 0.f) {
            /* the stack maximum (start value = the scale factor, dc.hpp:445) */
            float init[2] = {sf, 0.f}, mx = sf;
            RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->minmax, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
            stack_max_u16_kernel<<<stream_grid(ctx, n), 256, 0, ctx->stream>>>((const uint16_t*)raw, n, ctx->minmax);
            RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(&mx, ctx->minmax, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
            RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            sf = mx;
            ctx->timing.kernel_launches += 1;
        }
        L.nonneg = (sf > 0.f) ? 1 : 0;
        normalise_u16_kernel<<<stream_grid(ctx, n), 256, 0, ctx->stream>>>((const uint16_t*)raw, n, sf, L.epi_full);
        ctx->timing.kernel_launches += 1;
    } else {
        /* max (start value = the scale factor, dc.hpp:445) and min of the stack; those of the uploaded stack (level 0)
         * are a property of the input: computed once per upload, not once per run */
        float mm[2];
        if (p == 0 && ctx->mm0_epoch == ctx->input_epoch) { mm[0] = ctx->mm0[0]; mm[1] = ctx->mm0[1]; }
        else {
            float init[2] = {ctx->scale_factor, std::numeric_limits<float>::infinity()};
            RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->minmax, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
            stack_minmax_kernel<<<stream_grid(ctx, n), 256, 0, ctx->stream>>>((const float*)raw, n, ctx->minmax);
            RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(mm, ctx->minmax, sizeof(mm), cudaMemcpyDeviceToHost, ctx->stream));
            RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            ctx->timing.kernel_launches += 1;
            if (p == 0) { ctx->mm0[0] = mm[0]; ctx->mm0[1] = mm[1]; ctx->mm0_epoch = ctx->input_epoch; }
        }
        float sf = ctx->scale_factor;
        if (sf < 0.f) sf = mm[0];
        /* sign of a normalised value: sign(x) * sign(1/sf) */
        L.nonneg = ((mm[1] >= 0.f && sf > 0.f)) ? 1 : 0;
        normalise_f32_kernel<<<stream_grid(ctx, n), 256, 0, ctx->stream>>>((const float*)raw, n, nullptr, sf, L.epi_full);
        ctx->timing.kernel_launches += 1;
    }
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    return RSLF_OK;
}

/* The raw level-0 stack with ALL rows: the uploaded stack itself on one GPU; in a row-sharded run every rank
 * uploads its block and the blocks are all-gathered over NVLink once per input (cached until the next upload). */
static int full_input(rslf_ctx* ctx, const void** out)
{
    if (ctx->world <= 1 || !ctx->have_shards) { *out = ctx->raw_in; return RSLF_OK; }
    if (ctx->raw_full_epoch != ctx->input_epoch) {
        const size_t row_bytes = (size_t)ctx->S * ctx->U * ctx->C * depth_esz(ctx->cv_depth);
        const size_t bytes = row_bytes * ctx->V_total + 256;
        if (ctx->raw_full_cap < bytes) {
            if (ctx->raw_full) cudaFree(ctx->raw_full);
    if (ctx->img_staging) cudaFree(ctx->img_staging);
    if (ctx->open_ce) cudaFree(ctx->open_ce);
    if (ctx->open_mask) cudaFree(ctx->open_mask);
    if (ctx->open_tmp) cudaFree(ctx->open_tmp);
    for (int i = 0; i < 2; ++i) { if (ctx->ring[i]) cudaFreeHost(ctx->ring[i]); if (ctx->ring_ev[i]) cudaEventDestroy(ctx->ring_ev[i]); }
    if (ctx->ev_img) cudaEventDestroy(ctx->ev_img);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
            ctx->raw_full = nullptr; ctx->raw_full_cap = 0;
            cudaError_t e = cudaMalloc(&ctx->raw_full, bytes);
            if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(full raw stack %zu): %s", bytes, cudaGetErrorString(e)); return RSLF_ERR_NOMEM; }
            ctx->raw_full_cap = bytes;
        }
        shard_tab t; t.n = ctx->world;
        for (int r = 0; r <= ctx->world; ++r) t.b[r] = ctx->row_starts[r];
        RSLF_TRY(comm_gather_rows(ctx, ctx->raw_in, row_bytes, t, ctx->raw_full));
        ctx->raw_full_epoch = ctx->input_epoch;
    }
    *out = ctx->raw_full;
    return RSLF_OK;
}

/* ------------------------------------------------------------------ lifetime */
extern "C" int rslf_cuda_abi_version(void) { return RSLF_ABI_VERSION; }

extern "C" void rslf_params_default(rslf_params* p)
{
    if (!p) return;
    p->edge_score_threshold = 0.02f; p->line_score_threshold = 0.02f;
    p->disp_score_threshold = 0.01f; p->raw_score_threshold = 0.f;
    p->mean_shift_max_iter = 10; p->edge_confidence_filter_size = 9;
    p->edge_confidence_opening_type = 2; p->edge_confidence_opening_size = 1;
    p->median_filter_size = 5; p->median_filter_epsilon = 0.1f; p->propagation_epsilon = 0.1f;
    p->slope_factor = 1.0f; p->cut_shadows = 1; p->shadow_level = (float)(0.05 * 1.73205080757);
    p->kernel_h = 0.2f;
}

extern "C" const char* rslf_cuda_strerror(int code)
{
    switch (code) {
        case RSLF_OK: return "ok";
        case RSLF_ERR_CUDA: return "CUDA runtime error";
        case RSLF_ERR_ARG: return "invalid argument";
        case RSLF_ERR_STATE: return "call order violated";
        case RSLF_ERR_UNSUPPORTED: return "not implemented in the CUDA path";
        case RSLF_ERR_NCCL: return "NCCL error";
        case RSLF_ERR_NOMEM: return "out of device memory";
        case RSLF_ERR_PEER: return "a peer GPU did not answer in time";
        default: return "unknown error";
    }
}

extern "C" int rslf_cuda_create(int device, rslf_ctx** out)
{
    if (!out) return RSLF_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return RSLF_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return RSLF_ERR_CUDA;
    rslf_ctx* ctx = new rslf_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return RSLF_ERR_CUDA; }
    ctx->num_sm = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return RSLF_ERR_CUDA; }
    cudaEventCreate(&ctx->ev_a); cudaEventCreate(&ctx->ev_b);
    if (cudaMalloc((void**)&ctx->count, RSLF_COUNT_SLOTS * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->queue, RSLF_COUNT_SLOTS * sizeof(int)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->dev_err, sizeof(int)) != cudaSuccess ||
        cudaMemset(ctx->dev_err, 0, sizeof(int)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->total_px, 2 * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc((void**)&ctx->minmax, 4 * sizeof(float)) != cudaSuccess) {
        rslf_cuda_destroy(ctx); return RSLF_ERR_CUDA;
    }
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    const char* st = getenv("RSLF_STAGE_TIMING");
    if (st) ctx->stage_timing = atoi(st);
    const char* fm = getenv("RSLF_FAST_MATH");
    if (fm) ctx->fast_math = atoi(fm) != 0;
    const char* bl = getenv("RSLF_BALANCE");
    if (bl) ctx->balance = atoi(bl) != 0;
    *out = ctx;
    return RSLF_OK;
}

extern "C" void rslf_cuda_destroy(rslf_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    comm_destroy(ctx);
    for (int p = 0; p < RSLF_MAX_LEVELS; ++p) free_level(ctx->lv[p]);
    free_scratch(ctx);
    if (ctx->raw_in && !ctx->raw_borrowed) cudaFree(ctx->raw_in);
    if (ctx->raw_full) cudaFree(ctx->raw_full);
    if (ctx->img_staging) cudaFree(ctx->img_staging);
    if (ctx->open_ce) cudaFree(ctx->open_ce);
    if (ctx->open_mask) cudaFree(ctx->open_mask);
    if (ctx->open_tmp) cudaFree(ctx->open_tmp);
    if (ctx->ev_img) cudaEventDestroy(ctx->ev_img);
    if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
    dev_free(&ctx->queue); dev_free(&ctx->dev_err); dev_free(&ctx->dlog); dev_free(&ctx->dlog_count);
    dev_free(&ctx->count); dev_free(&ctx->total_px); dev_free(&ctx->minmax); dev_free(&ctx->rowwork);
    dev_free(&ctx->colour_hist); dev_free(&ctx->colour_lut);
    if (ctx->l2_flush) cudaFree(ctx->l2_flush);
    for (auto e : ctx->clk.pool) cudaEventDestroy(e);
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* rslf_cuda_last_error_text(const rslf_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

extern "C" int rslf_cuda_last_timing(const rslf_ctx* ctx, rslf_timing* out)
{
    if (!ctx || !out) return RSLF_ERR_ARG;
    *out = ctx->timing;
    return RSLF_OK;
}

/* Diagnostics: the individual event spans of the last run in issue order (stage id 0 edge, 1 depth, 2 reduce, 3 median,
 * 4 propagate, 5 pyramid; milliseconds); with rslf_cuda_set_stage_timing(2) that is one entry per stage and pass. */
extern "C" int rslf_cuda_last_spans(const rslf_ctx* ctx, int* stage, float* ms, size_t max_spans, size_t* count)
{
    if (!ctx || !count) return RSLF_ERR_ARG;
    const size_t n = std::min(ctx->clk.spans.size(), ctx->clk.span_ms.size());
    *count = n;
    for (size_t i = 0; i < n && i < max_spans; ++i) { if (stage) stage[i] = ctx->clk.spans[i].stage; if (ms) ms[i] = ctx->clk.span_ms[i]; }
    return RSLF_OK;
}

extern "C" int rslf_cuda_set_stage_timing(rslf_ctx* ctx, int on)
{
    if (!ctx) return RSLF_ERR_ARG;
    ctx->stage_timing = on;
    return RSLF_OK;
}

extern "C" int rslf_cuda_set_fast_math(rslf_ctx* ctx, int on)
{
    if (!ctx) return RSLF_ERR_ARG;
    ctx->fast_math = on != 0;
    return RSLF_OK;
}

extern "C" int rslf_cuda_set_decision_log(rslf_ctx* ctx, size_t capacity)
{
    if (!ctx) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->dlog) { cudaFree(ctx->dlog); ctx->dlog = nullptr; }
    ctx->dlog_cap = 0;
    if (capacity == 0) return RSLF_OK;
    if (!ctx->dlog_count) RSLF_TRY(dev_alloc(ctx, &ctx->dlog_count, 1));
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(ctx->dlog_count, 0, sizeof(int), ctx->stream));
    RSLF_TRY(dev_alloc(ctx, &ctx->dlog, capacity));
    ctx->dlog_cap = capacity;
    return RSLF_OK;
}

extern "C" int rslf_cuda_get_decision_log(rslf_ctx* ctx, rslf_decision* out, size_t max_records, size_t* count)
{
    if (!ctx || !count) return RSLF_ERR_ARG;
    *count = 0;
    if (!ctx->dlog) { snprintf(ctx->err, sizeof(ctx->err), "no decision log (rslf_cuda_set_decision_log)"); return RSLF_ERR_STATE; }
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int n = 0;
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(&n, ctx->dlog_count, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    *count = (size_t)n;
    const size_t m = std::min<size_t>(std::min<size_t>((size_t)n, ctx->dlog_cap), max_records);
    if (out && m) {
        RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->dlog, m * sizeof(rslf_decision), cudaMemcpyDeviceToHost, ctx->stream));
        RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return RSLF_OK;
}

/* 0: edge confidence (the reference's default build); 1: disparity confidence, the reference as intended with
 * -D_USE_DISP_CONFIDENCE_SCORE (`#elseif` read as `#elif`): core.hpp:1097-1098, dc.hpp:901-902 */
extern "C" int rslf_cuda_set_confidence_criterion(rslf_ctx* ctx, int criterion)
{
    if (!ctx) return RSLF_ERR_ARG;
    if (criterion == 2) { snprintf(ctx->err, sizeof(ctx->err), "the line-confidence criterion (core.hpp:1032-1081) is not implemented"); return RSLF_ERR_UNSUPPORTED; }
    if (criterion != 0 && criterion != 1) return RSLF_ERR_ARG;
    ctx->criterion = criterion;
    return RSLF_OK;
}

/* 1 (default): pass-balanced depth kernel in row-sharded runs; 0: lock-step row blocks (k_balance.cuh) */
extern "C" int rslf_cuda_set_balance(rslf_ctx* ctx, int on)
{
    if (!ctx) return RSLF_ERR_ARG;
    ctx->balance = on != 0;
    return RSLF_OK;
}

extern "C" int rslf_cuda_sync(rslf_ctx* ctx)
{
    if (!ctx) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RSLF_OK;
}

extern "C" int rslf_cuda_flush_l2(rslf_ctx* ctx)
{
    if (!ctx) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t bytes = 256u << 20;           /* > 126 MB L2 */
    if (!ctx->l2_flush) RSLF_CUDA_TRY(ctx, cudaMalloc(&ctx->l2_flush, bytes));
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(ctx->l2_flush, 0x5a, bytes, ctx->stream));
    return RSLF_OK;
}

/* ------------------------------------------------------------------ input */
static int set_dims(rslf_ctx* ctx, int V, int S, int U, int C, int cv_depth, float scale)
{
    if (V < 1 || S < 1 || U < 1 || (C != 1 && C != 3)) { snprintf(ctx->err, sizeof(ctx->err), "bad dimensions"); return RSLF_ERR_ARG; }
    if (cv_depth != RSLF_DEPTH_8U && cv_depth != RSLF_DEPTH_16U && cv_depth != RSLF_DEPTH_32F) {
        snprintf(ctx->err, sizeof(ctx->err), "only CV_8U, CV_16U and CV_32F inputs are implemented"); return RSLF_ERR_UNSUPPORTED;
    }
    if ((size_t)S * 4 > RSLF_COUNT_SLOTS / RSLF_MAX_LEVELS) { snprintf(ctx->err, sizeof(ctx->err), "S too large"); return RSLF_ERR_UNSUPPORTED; }
    ctx->V = V; ctx->S = S; ctx->U = U; ctx->C = C; ctx->cv_depth = cv_depth; ctx->scale_factor = scale;
    ++ctx->input_epoch;                                      /* the gathered copy of the stack (full_input) is stale */
    if (ctx->V_total == 0 || ctx->world == 1) { ctx->v0 = 0; ctx->V_total = V; }
    return RSLF_OK;
}

static int own_raw(rslf_ctx* ctx, size_t bytes)
{
    if (ctx->raw_borrowed) { ctx->raw_in = nullptr; ctx->raw_borrowed = false; ctx->raw_cap = 0; }
    if (ctx->raw_cap < bytes) {
        if (ctx->raw_in) cudaFree(ctx->raw_in);
        ctx->raw_in = nullptr; ctx->raw_cap = 0;
        cudaError_t e = cudaMalloc(&ctx->raw_in, bytes);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(raw %zu): %s", bytes, cudaGetErrorString(e)); return RSLF_ERR_NOMEM; }
        ctx->raw_cap = bytes;
    }
    return RSLF_OK;
}

static int images_to_device(rslf_ctx* ctx, void* dev, int n_imgs, int rows, size_t row_bytes, const void* const* host, size_t step);
static int planes_to_host(rslf_ctx* ctx, const void* dev, int planes, int rows, size_t row_bytes, void* const* host, size_t step);

extern "C" int rslf_cuda_upload_epis(rslf_ctx* ctx, const void* const* epi_ptrs, int V, int S, int U, int C,
                                     int cv_depth, size_t row_step_bytes, float epi_scale_factor)
{
    if (!ctx || !epi_ptrs) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_TRY(set_dims(ctx, V, S, U, C, cv_depth, epi_scale_factor));
    const size_t esz = depth_esz(cv_depth);
    const size_t row = (size_t)U * C * esz;
    if (row_step_bytes < row) { snprintf(ctx->err, sizeof(ctx->err), "row step smaller than a row"); return RSLF_ERR_ARG; }
    const size_t epi_bytes = row * S;
    RSLF_TRY(own_raw(ctx, epi_bytes * V));
    cudaEventRecord(ctx->ev_a, ctx->stream);
    /* V images of S rows: pinned / registered memory by DMA, pageable cv::Mat storage through the pinned ring */
    RSLF_TRY(images_to_device(ctx, ctx->raw_in, V, S, row, epi_ptrs, row_step_bytes));
    cudaEventRecord(ctx->ev_b, ctx->stream);
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.ms_h2d, ctx->ev_a, ctx->ev_b);
    ctx->have_input = true;
    return RSLF_OK;
}

/* min over a range, for the sign flag of the normalised stack (stack_minmax_kernel needs both slots initialised) */
static bool edge_params_equal(const rslf_params& a, const rslf_params& b)
{
    return a.edge_score_threshold == b.edge_score_threshold && a.edge_confidence_filter_size == b.edge_confidence_filter_size &&
           a.edge_confidence_opening_size == b.edge_confidence_opening_size && a.edge_confidence_opening_type == b.edge_confidence_opening_type &&
           a.cut_shadows == b.cut_shadows && a.shadow_level == b.shadow_level && a.propagation_epsilon == b.propagation_epsilon;
}

/* Ingest fused with the computers' input handling: the V EPIs are uploaded in chunks and, while the
 * next chunk is still crossing PCIe, the chunk that has arrived is normalised to float32 (dc.hpp:463-477) and —
 * when `params` is given — its edge confidence C_e and mask are computed for all S lines (core.hpp:901-931) on a
 * second stream.  The following Depth2DComputer / FineToCoarse run on the same input (and the same edge parameters)
 * starts from those level-0 maps instead of recomputing them.  Needs the scale to be known up front (8-bit input, or
 * epi_scale_factor >= 0: with scale < 0 the stack maximum is only known after the last byte) and a single-rank
 * context without the opening; otherwise this is rslf_cuda_upload_epis. */
extern "C" int rslf_cuda_upload_epis_pipelined(rslf_ctx* ctx, const void* const* epi_ptrs, int V, int S, int U, int C,
                                               int cv_depth, size_t row_step_bytes, float epi_scale_factor, const rslf_params* params)
{
    if (!ctx || !epi_ptrs) return RSLF_ERR_ARG;
    const bool scale_known = (cv_depth == RSLF_DEPTH_8U) || (epi_scale_factor >= 0.f);
    if (!scale_known || ctx->world > 1 || (params && params->edge_confidence_opening_size > 1) || (epi_scale_factor == 0.f && cv_depth != RSLF_DEPTH_8U))
        return rslf_cuda_upload_epis(ctx, epi_ptrs, V, S, U, C, cv_depth, row_step_bytes, epi_scale_factor);
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_TRY(set_dims(ctx, V, S, U, C, cv_depth, epi_scale_factor));
    const size_t esz = depth_esz(cv_depth);
    const size_t row = (size_t)U * C * esz;
    if (row_step_bytes < row) { snprintf(ctx->err, sizeof(ctx->err), "row step smaller than a row"); return RSLF_ERR_ARG; }
    const size_t epi_bytes = row * S, epi_vals = (size_t)S * U * C;
    RSLF_TRY(own_raw(ctx, epi_bytes * V));
    ctx->row_starts[0] = 0; ctx->row_starts[1] = V; ctx->v0 = 0; ctx->V_total = V;
    RSLF_TRY(ensure_scratch(ctx, true, false));
    RSLF_TRY(ensure_level(ctx, 0, V, U, true, false));
    rslf_level& L = ctx->lv[0];
    if (!ctx->stream2) RSLF_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    if (!ctx->ev_img) RSLF_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_img, cudaEventDisableTiming));
    cudaEventRecord(ctx->ev_a, ctx->stream);
    float init[2] = {0.f, std::numeric_limits<float>::infinity()};
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->minmax, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    if (params) {
        RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.rowdark, 0, ((size_t)S + 1) * V * sizeof(int), ctx->stream));
        RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.dark_lo, 0x7f, (size_t)S * V * sizeof(int), ctx->stream));
        RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.dark_hi, 0xff, (size_t)S * V * sizeof(int), ctx->stream));
    }
    const float sf = epi_scale_factor;
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)V, RSLF_RING_BYTES / epi_bytes));
    for (int v0 = 0; v0 < V; v0 += chunk) {
        const int n = std::min(chunk, V - v0);
        RSLF_TRY(images_to_device(ctx, (char*)ctx->raw_in + (size_t)v0 * epi_bytes, n, S, row, epi_ptrs + v0, row_step_bytes));
        RSLF_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_img, ctx->stream));
        RSLF_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_img, 0));
        const size_t nv = (size_t)n * epi_vals;
        float* out = L.epi_full + (size_t)v0 * epi_vals;
        if (cv_depth == RSLF_DEPTH_8U)
            normalise_u8_kernel<<<stream_grid(ctx, nv), 256, 0, ctx->stream2>>>((const uint8_t*)ctx->raw_in + (size_t)v0 * epi_vals, nv, out);
        else if (cv_depth == RSLF_DEPTH_16U)
            normalise_u16_kernel<<<stream_grid(ctx, nv), 256, 0, ctx->stream2>>>((const uint16_t*)ctx->raw_in + (size_t)v0 * epi_vals, nv, sf, out);
        else {
            stack_minmax_kernel<<<stream_grid(ctx, nv), 256, 0, ctx->stream2>>>((const float*)ctx->raw_in + (size_t)v0 * epi_vals, nv, ctx->minmax);
            normalise_f32_kernel<<<stream_grid(ctx, nv), 256, 0, ctx->stream2>>>((const float*)ctx->raw_in + (size_t)v0 * epi_vals, nv, nullptr, sf, out);
        }
        ctx->timing.kernel_launches += 1;
        if (params) RSLF_TRY(launch_edge_confidence_rows(ctx, ctx->stream2, L.epi_full, V, S, U, C, v0, n, *params, L.ce, L.emask, L.rowdark, L.dark_lo, L.dark_hi));
    }
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    RSLF_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_img, ctx->stream2));
    RSLF_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_img, 0));
    float mm[2] = {0.f, 0.f};
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(mm, ctx->minmax, sizeof(mm), cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(ctx->ev_b, ctx->stream);
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.ms_h2d, ctx->ev_a, ctx->ev_b);
    ctx->have_input = true;
    ctx->pre_norm_epoch = ctx->input_epoch;
    ctx->pre_nonneg = (cv_depth == RSLF_DEPTH_8U) ? 1 : (cv_depth == RSLF_DEPTH_16U) ? (sf > 0.f ? 1 : 0) : ((mm[1] >= 0.f && sf > 0.f) ? 1 : 0);
    if (params) { ctx->pre_edge_epoch = ctx->input_epoch; ctx->pre_params = *params; }
    return RSLF_OK;
}

extern "C" int rslf_cuda_set_epis_device(rslf_ctx* ctx, const void* d_epis, int V, int S, int U, int C,
                                         int cv_depth, float epi_scale_factor)
{
    if (!ctx || !d_epis) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_TRY(set_dims(ctx, V, S, U, C, cv_depth, epi_scale_factor));
    if (ctx->raw_in && !ctx->raw_borrowed) cudaFree(ctx->raw_in);
    ctx->raw_in = const_cast<void*>(d_epis);
    ctx->raw_borrowed = true; ctx->raw_cap = 0;
    ctx->have_input = true;
    ctx->timing.ms_h2d = 0.f;
    return RSLF_OK;
}

/* [S][V][U][C] -> [V][S][U][C] (rslf::build_epis_from_imgs, src/rslf_io.cpp:194-227) */
template <typename T>
__global__ void build_epis_kernel(const T* __restrict__ imgs, int S, int V, size_t rowlen, T* __restrict__ epis, int s_first)
{
    const int s = s_first + blockIdx.y, v = blockIdx.z;
    const T* src = imgs + ((size_t)s * V + v) * rowlen;
    T* dst = epis + ((size_t)v * S + s) * rowlen;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < rowlen; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

extern "C" int rslf_cuda_upload_images(rslf_ctx* ctx, const void* const* img_ptrs, int S, int V, int U, int C,
                                       int cv_depth, size_t row_step_bytes, float epi_scale_factor)
{
    if (!ctx || !img_ptrs) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_TRY(set_dims(ctx, V, S, U, C, cv_depth, epi_scale_factor));
    const size_t esz = depth_esz(cv_depth);
    const size_t row = (size_t)U * C * esz;
    if (row_step_bytes < row) return RSLF_ERR_ARG;
    const size_t img_bytes = row * V;
    RSLF_TRY(own_raw(ctx, img_bytes * S));
    /* staging area for the image stack, kept for the next upload (no allocation per call) */
    if (ctx->img_staging_cap < img_bytes * S) {
        if (ctx->img_staging) cudaFree(ctx->img_staging);
    if (ctx->open_ce) cudaFree(ctx->open_ce);
    if (ctx->open_mask) cudaFree(ctx->open_mask);
    if (ctx->open_tmp) cudaFree(ctx->open_tmp);
        ctx->img_staging = nullptr; ctx->img_staging_cap = 0;
        cudaError_t e = cudaMalloc(&ctx->img_staging, img_bytes * S);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(image staging %zu): %s", img_bytes * S, cudaGetErrorString(e)); return RSLF_ERR_NOMEM; }
        ctx->img_staging_cap = img_bytes * S;
    }
    cudaEventRecord(ctx->ev_a, ctx->stream);
    /* image by image: the transfer of image s + 1 (copy engine) overlaps the transposition of image s (SMs, second
     * stream), which reads each staged image once and writes its V scanlines into the EPIs (rslf_io.cpp:203-219) */
    if (!ctx->stream2) RSLF_CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
    if (!ctx->ev_img) RSLF_CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_img, cudaEventDisableTiming));
    for (int s = 0; s < S; ++s) {
        const void* one = img_ptrs[s];
        RSLF_TRY(images_to_device(ctx, (char*)ctx->img_staging + (size_t)s * img_bytes, 1, V, row, &one, row_step_bytes));
        RSLF_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_img, ctx->stream));
        RSLF_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_img, 0));
        dim3 grid(std::max(1, std::min(8, rslf_div_up((long long)U * C, 256))), 1, V);
        if (esz == 1) build_epis_kernel<uint8_t><<<grid, 256, 0, ctx->stream2>>>((const uint8_t*)ctx->img_staging, S, V, (size_t)U * C, (uint8_t*)ctx->raw_in, s);
        else if (esz == 2) build_epis_kernel<uint16_t><<<grid, 256, 0, ctx->stream2>>>((const uint16_t*)ctx->img_staging, S, V, (size_t)U * C, (uint16_t*)ctx->raw_in, s);
        else build_epis_kernel<float><<<grid, 256, 0, ctx->stream2>>>((const float*)ctx->img_staging, S, V, (size_t)U * C, (float*)ctx->raw_in, s);
    }
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    RSLF_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_img, ctx->stream2));
    RSLF_CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_img, 0));
    cudaEventRecord(ctx->ev_b, ctx->stream);
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.ms_h2d, ctx->ev_a, ctx->ev_b);
    ctx->have_input = true;
    return RSLF_OK;
}

/* gathered planes of one line (row-sharded runs) */
static int ensure_gather_planes(rslf_ctx* ctx, size_t plane_tot)
{
    if (ctx->g_plane_cap >= plane_tot) return RSLF_OK;
    RSLF_TRY(dev_alloc(ctx, &ctx->g_depth, plane_tot));
    RSLF_TRY(dev_alloc(ctx, &ctx->g_colour, plane_tot * ctx->C));
    RSLF_TRY(dev_alloc(ctx, &ctx->g_mask, plane_tot));
    ctx->g_plane_cap = plane_tot;
    return RSLF_OK;
}

/* Global view of a row-sharded [S][Vloc][U] map: the map itself on one GPU, else gathered into slot 0 / 1. */
static int global_svu_f32(rslf_ctx* ctx, const float* local, int S, int U, const shard_tab& t, int slot, const float** out)
{
    if (ctx->world <= 1) { *out = local; return RSLF_OK; }
    float* dst = slot ? ctx->g_map2 : ctx->g_map;
    RSLF_TRY(comm_gather_svu(ctx, local, 4, S, U, t, dst));
    *out = dst;
    return RSLF_OK;
}
static int global_svu_u8(rslf_ctx* ctx, const uint8_t* local, int S, int U, const shard_tab& t, int slot, const uint8_t** out)
{
    if (ctx->world <= 1) { *out = local; return RSLF_OK; }
    uint8_t* dst = slot ? ctx->g_mk2 : ctx->g_mk;
    RSLF_TRY(comm_gather_svu(ctx, local, 1, S, U, t, dst));
    *out = dst;
    return RSLF_OK;
}

__global__ void row_sum_kernel(const int* __restrict__ rowdark, int S, int V, int* __restrict__ out)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    int t = 0;
    for (int s = 0; s < S; ++s) t += rowdark[(size_t)v * S + s];        /* [V][S]: the views of a row lie side by side */
    out[v] = t;
}

/* ------------------------------------------------------------------ one s_hat pass */
struct pass_io {
    int level; int s_hat; int D; float dmin, dmax; bool use_bound_maps;
    bool bounds_in_range;     /* every per-pixel bound lies inside [dmin, dmax] (always true for the pyramid's own bounds) */
    bool pile;                /* 1D pile: single-plane maps, no remaining mask */
    int count_slot;
};

/* Peer memory for the row-sharded run: the arena of every rank (rslf_comm.cuh), sized for the largest level-0
 * block.  Collective; RSLF_ERR_UNSUPPORTED (on every rank) when peer memory is unavailable. */
static int ensure_arena(rslf_ctx* ctx)
{
    if (ctx->world <= 1) return RSLF_ERR_UNSUPPORTED;
    int max_rows = 1;
    for (int q = 0; q < ctx->world; ++q) max_rows = std::max(max_rows, ctx->row_starts[q + 1] - ctx->row_starts[q]);
    return comm_arena_setup(ctx, ctx->U, ctx->C, (size_t)max_rows * ctx->U);
}

/* compute_1D_depth_epi_pile (core.hpp:772-893) on line s_hat of level L: compaction,
 * depth kernel, selective median into ctx->filtered.
 *
 * Row-sharded level, three ways to cross the block borders:
 *   balanced  (default): pass-balanced depth kernel + median halo through peer memory (k_balance.cuh)
 *   lock-step (RSLF_BALANCE=0, or the 1D pile): every rank evaluates its own rows, median halo through peer memory
 *                        or, without peer memory, one small all-gather
 *   gather               thin blocks / windows wider than 5x5: all-gather of the planes of line s_hat */
static int run_depth_pass(rslf_ctx* ctx, const rslf_params& P, const pass_io& io)
{
    rslf_level& L = ctx->lv[io.level];
    const int V = L.V, U = L.U, S = ctx->S, C = ctx->C;
    const size_t plane = (size_t)V * U;
    const size_t po = io.pile ? 0 : (size_t)io.s_hat * plane;
    int* count = ctx->count + io.count_slot;
    int* queue = ctx->queue + io.count_slot;
    const bool sharded = (ctx->world > 1 && !L.replicated);
    const float* colour0 = L.epi + (size_t)io.s_hat * U * C;      /* colours of line s_hat: row v at + v * S*U*C */
    const size_t colour_stride = (size_t)S * U * C;
    const uint8_t* fresh = io.pile ? nullptr : L.remaining + po;
    const int* rdv = io.pile ? nullptr : L.rowdark + (size_t)S * V;
    median_gate gate; memset(&gate, 0, sizeof(gate));
    gate.rowdark = L.rowdark; gate.lo = L.dark_lo; gate.hi = L.dark_hi; gate.S = S; gate.s_hat = io.s_hat; gate.slope = P.slope_factor;
    gate.dlo = std::min(io.dmin, io.dmax); gate.dhi = std::max(io.dmin, io.dmax);
    gate.on = (!io.pile && io.bounds_in_range) ? 1 : 0;
    { const char* eg = getenv("RSLF_MEDIAN_GATE"); if (eg && eg[0] == '0') gate.on = 0; }      /* test hook */
    /* row-sharded level: how the rows next to the block reach the neighbours */
    shard_tab t; bool halo_path = false, p2p = false, balanced = false;
    if (sharded) {
        t = level_shards(ctx, io.level, L.Vtot);
        int min_rows = L.Vtot;
        for (int q = 0; q < t.n; ++q) min_rows = std::min(min_rows, t.b[q + 1] - t.b[q]);
        const char* force_full = getenv("RSLF_MEDIAN_GATHER");         /* "full": test hook for the fallback */
        halo_path = (P.median_filter_size - 1) / 2 <= 2 && min_rows >= 2 && !(force_full && force_full[0] == 'f');
        p2p = (ensure_arena(ctx) == RSLF_OK);
        balanced = p2p && halo_path && ctx->balance && !io.pile && ctx->world <= RSLF_MAX_PEERS;
    }
    const arena_layout al = p2p ? arena_current(ctx) : arena_layout();
    int4* rec = balanced ? reinterpret_cast<int4*>(ctx->arena + al.off_rec) : ctx->rec_own;
    const unsigned seq = balanced ? ++ctx->bal_seq : 0u;
    {
        stage_scope sc(ctx, ST_REDUCE);
        compact_args c; memset(&c, 0, sizeof(c));
        c.emask = L.emask + po; c.remaining = io.pile ? nullptr : L.remaining + po; c.n = (int)plane;
        c.items = ctx->items; c.rec = rec; c.count = count; c.total = ctx->total_px + (L.replicated ? 1 : 0);
        c.rowwork = L.replicated ? nullptr : ctx->rowwork; c.U = U; c.v0 = L.v0; c.shift = io.level; c.v0_base = ctx->v0; c.rows_base = ctx->V;
        c.V = V; c.select = 0;
        c.dmin_map = io.use_bound_maps ? L.dmin + po : nullptr; c.dmax_map = io.use_bound_maps ? L.dmax + po : nullptr;
        c.dmin_c = io.dmin; c.dmax_c = io.dmax; c.ce = L.ce + po;
        c.pix_off = L.v0 * U;                                    /* records address the level's full stack */
        if (balanced) {
            c.bal_n = ctx->world; c.seq = seq; c.blocks_done = ctx->p2p_done + 1;
            for (int q = 0; q < ctx->world; ++q)
                c.counts_peer[q] = reinterpret_cast<unsigned long long*>(ctx->arena_peer[q] + al.off_counts) + ctx->rank;
        }
        compact_kernel<<<stream_grid(ctx, plane), 256, 0, ctx->stream>>>(c);
        RSLF_CUDA_TRY(ctx, cudaGetLastError());
        ctx->timing.kernel_launches += 1;
    }
    depth_plan plan = plan_depth(ctx, S, C, io.D, io.s_hat, io.dmin, io.dmax, P.slope_factor);
    /* tensor-memory variant (k_depth_tm.cuh): RSLF_DEPTH_TMEM = 0 off, 1 RGB stacks (the planner's H = 1 case), 2 every stack */
    depth_tm_layout tml; bool use_tm = false;
    {
        const char* et = getenv("RSLF_DEPTH_TMEM");
        const int mode = et ? atoi(et) : RSLF_DEPTH_TMEM_DEFAULT;
        const bool forced = getenv("RSLF_DEPTH_H") || getenv("RSLF_DEPTH_RV");
        if (mode > 0 && !forced && (mode >= 2 || C == 3)) {
            const int wpv1 = depth_wpv_q16(io.D, 1, io.dmin, io.dmax, P.slope_factor);
            if (depth_tm_build(tml, S, C, io.s_hat, wpv1, ctx->smem_optin)) {
                use_tm = true;
                plan.H = 1; plan.RV = DEPTH_TM_RV; plan.reg_last = tml.reg_last; plan.wpv_q16 = wpv1;
                plan.chunks = rslf_div_up(io.D, 32);
            }
        }
    }
    if (ctx->fast_math && !use_tm) {
        snprintf(ctx->err, sizeof(ctx->err), "fast_math: the contracted mean shift exists only in the tensor-memory depth kernel, "
                 "which does not take this pass (C=%d S=%d D=%d)", C, S, io.D);
        return RSLF_ERR_UNSUPPORTED;
    }
    if (plan.chunks > 1) RSLF_TRY(ensure_partials(ctx, ctx->share_cap * plan.chunks));
    depth_args a; memset(&a, 0, sizeof(a));
    a.epi = L.epi_full; a.V = L.Vtot; a.S = S; a.U = U; a.D = io.D; a.s_hat = io.s_hat; a.slope = P.slope_factor;
    a.inv = kernel_inv(P, C); a.iters = P.mean_shift_max_iter; a.negzero = -0.0f;
    a.rec = rec; a.count = count; a.queue = queue; a.pix_off = L.v0 * U;
    a.ce = L.ce + po; a.emask = L.emask + po; a.cd = L.cd + po;
    a.depth = io.pile ? ctx->pile_depth_raw : L.depth + po;
    a.rbar = L.rbar + po * C;
    a.raw_thr = P.raw_score_threshold;
    a.wpv_q16 = plan.wpv_q16; a.reg_last = plan.reg_last;
    a.chunks = plan.chunks; a.partials = (rslf_partial*)ctx->partials; a.arrive = ctx->arrive;
    a.log = ctx->dlog; a.log_count = ctx->dlog_count; a.log_cap = (int)std::min<size_t>(ctx->dlog_cap, 0x7fffffff); a.level = io.level;
    if (balanced) {
        a.bal.n = ctx->world; a.bal.rank = ctx->rank; a.bal.seq = seq;
        for (int q = 0; q < ctx->world; ++q) {
            a.bal.rec[q] = reinterpret_cast<const int4*>(ctx->arena_peer[q] + al.off_rec);
            a.bal.res[q] = reinterpret_cast<float4*>(ctx->arena_peer[q] + al.off_res);
            a.bal.done[q] = reinterpret_cast<unsigned long long*>(ctx->arena_peer[q] + al.off_done) + ctx->rank;
        }
        a.bal.counts = reinterpret_cast<const volatile unsigned long long*>(ctx->arena + al.off_counts);
        a.bal.blocks_done = ctx->p2p_done + 2; a.bal.err = ctx->dev_err;
    }
    {
        stage_scope sc(ctx, ST_DEPTH);
        if (use_tm) RSLF_TRY(launch_depth_tm(ctx, C, L.nonneg != 0, a, tml, ctx->fast_math != 0));
        else RSLF_TRY(launch_depth(ctx, C, L.nonneg != 0, a, plan));
    }
    {
        stage_scope sc(ctx, ST_MEDIAN);
        median_halo halo; memset(&halo, 0, sizeof(halo));
        if (balanced) {
            /* the owner applies the result records of its rows, then its first / last rows go to the neighbours;
             * colours are read in place: every rank holds the whole stack */
            bal_apply_args g; memset(&g, 0, sizeof(g));
            g.items = ctx->items; g.count = count; g.res = reinterpret_cast<const float4*>(ctx->arena + al.off_res); g.C = C;
            g.depth = a.depth; g.cd = a.cd; g.rbar = a.rbar; g.ce = a.ce; g.emask = a.emask;
            g.done = reinterpret_cast<const volatile unsigned long long*>(ctx->arena + al.off_done);
            g.n = ctx->world; g.seq = seq; g.err = ctx->dev_err; g.blocks_done = ctx->p2p_done + 3; g.do_push = 1;
            p2p_prepare_halo(ctx, a.depth, L.emask + po, nullptr, colour_stride, U, C, t, &g.push, &halo);
            bal_apply_kernel<<<ctx->num_sm, 256, 0, ctx->stream>>>(g);
            RSLF_CUDA_TRY(ctx, cudaGetLastError());
            ctx->timing.kernel_launches += 1;
        } else if (halo_path) {
            /* the rows next to the block go to the neighbours: peer-to-peer stores over NVLink when CUDA IPC between
             * the ranks works, else one small all-gather */
            if (p2p) RSLF_TRY(comm_p2p_exchange_median_halo(ctx, a.depth, L.emask + po, nullptr, colour_stride, U, C, t, &halo));
            else RSLF_TRY(comm_exchange_median_halo(ctx, a.depth, L.emask + po, colour0, colour_stride, U, C, t, &halo));
        }
        if (halo_path && p2p) {
            /* colour rows of the neighbours: in place in this rank's copy of the stack */
            halo.colour_halo_stride = colour_stride;
            if (halo.top_depth) halo.top_colour = colour0 - 2 * colour_stride;
            if (halo.bot_depth) halo.bot_colour = colour0 + (size_t)V * colour_stride;
        }
        if (halo_path) {
            RSLF_TRY(launch_selective_median(ctx, a.depth, L.emask + po, colour0, colour_stride, V, U, C,
                                             P.median_filter_size, P.median_filter_epsilon, ctx->filtered, 0, -1, fresh, rdv, &halo, &gate));
        } else if (sharded) {
            /* thin blocks or wide windows: gather the depth / mask planes of line s_hat (colours: the whole stack is here) */
            RSLF_TRY(ensure_gather_planes(ctx, (size_t)L.Vtot * U));
            RSLF_TRY(comm_gather_median_planes(ctx, a.depth, L.emask + po, colour0, colour_stride, U, C, t));
            RSLF_TRY(launch_selective_median(ctx, ctx->g_depth, ctx->g_mask, ctx->g_colour, (size_t)U * C, L.Vtot, U, C,
                                             P.median_filter_size, P.median_filter_epsilon, ctx->filtered, L.v0, V, fresh, rdv));
        } else {
            RSLF_TRY(launch_selective_median(ctx, a.depth, L.emask + po, colour0, colour_stride,
                                             V, U, C, P.median_filter_size, P.median_filter_epsilon, ctx->filtered, 0, -1, fresh, rdv, nullptr, &gate));
        }
    }
    return RSLF_OK;
}

static std::vector<int> visiting_order(int S)
{
    /* core.hpp:953, 981-990 */
    int s_hat = (int)std::floor(S / 2.0);
    std::vector<int> order;
    order.push_back(s_hat);
    for (int off = 1; off < S - s_hat; ++off) {
        order.push_back(s_hat + off);
        if (s_hat - off > -1) order.push_back(s_hat - off);
    }
    return order;
}

/* Depth2DComputer::run (dc.hpp:748-805) on level p (stack already normalised in L.epi). */
static int run_depth2d_level(rslf_ctx* ctx, int p, const rslf_params& P, float dmin, float dmax, int D, bool use_bound_maps,
                             bool bounds_in_range = true)
{
    rslf_level& L = ctx->lv[p];
    const int V = L.V, U = L.U, S = ctx->S, C = ctx->C;
    const size_t plane = (size_t)V * U, px = plane * S;
    /* result maps start from zero (dc.hpp:741-744; C_e / C_d: fresh zero pages) */
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.depth, 0, px * sizeof(float), ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.cd, 0, px * sizeof(float), ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.rbar, 0, px * C * sizeof(float), ctx->stream));
    /* level 0 of an input whose edge confidence was computed while it was uploaded, with the same parameters? */
    const bool pre_edge = (p == 0 && ctx->world <= 1 && ctx->pre_edge_epoch == ctx->input_epoch && edge_params_equal(ctx->pre_params, P));
    if (!pre_edge) {
        RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.rowdark, 0, ((size_t)S + 1) * V * sizeof(int), ctx->stream));
        RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.dark_lo, 0x7f, (size_t)S * V * sizeof(int), ctx->stream));
        RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.dark_hi, 0xff, (size_t)S * V * sizeof(int), ctx->stream));
    } else RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.rowdark + (size_t)S * V, 0, (size_t)V * sizeof(int), ctx->stream));
    if (pre_edge) {
        stage_scope sc(ctx, ST_EDGE);
        row_sum_kernel<<<rslf_div_up(V, 256), 256, 0, ctx->stream>>>(L.rowdark, S, V, L.rowdark + (size_t)S * V);
        RSLF_CUDA_TRY(ctx, cudaGetLastError());
        ctx->timing.kernel_launches += 1;
        ctx->pre_edge_epoch = 0;                             /* the run consumes the maps (the depth kernel clears C_e of rejected pixels) */
    } else {
        stage_scope sc(ctx, ST_EDGE);
        const bool sh = ctx->world > 1 && !L.replicated;
        RSLF_TRY(launch_edge_confidence(ctx, L.epi, V, S, U, C, 0, S, P, L.ce, L.emask, L.rowdark, L.remaining,   /* L.remaining is free until core.hpp:958-963 below */
                                        sh ? L.v0 : 0, sh ? L.Vtot - L.v0 - V : 0, L.dark_lo, L.dark_hi));
        /* rows that hold a dark target in any view: rowdark[S][v] = sum over s (an upper bound for the whole level) */
        row_sum_kernel<<<rslf_div_up(V, 256), 256, 0, ctx->stream>>>(L.rowdark, S, V, L.rowdark + (size_t)S * V);
        RSLF_CUDA_TRY(ctx, cudaGetLastError());
        ctx->timing.kernel_launches += 1;
    }
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(L.remaining, L.emask, px, cudaMemcpyDeviceToDevice, ctx->stream));   /* core.hpp:958-963 */
    std::vector<int> order = visiting_order(S);
    int pass = 0;
    for (int s_hat : order) {
        pass_io io;
        io.level = p; io.s_hat = s_hat; io.D = D; io.dmin = dmin; io.dmax = dmax; io.use_bound_maps = use_bound_maps;
        io.pile = false; io.count_slot = p * 2 * S + pass; io.bounds_in_range = bounds_in_range;
        RSLF_TRY(run_depth_pass(ctx, P, io));
        {
            stage_scope sc(ctx, ST_PROP);
            prop_args a;
            a.epi = L.epi; a.V = V; a.S = S; a.U = U; a.s_hat = s_hat; a.slope = P.slope_factor;
            a.eps = P.propagation_epsilon; a.eps_T = rslf_sq_threshold(P.propagation_epsilon);
            a.emask_p = L.emask + (size_t)s_hat * plane; a.filtered = ctx->filtered;
            a.rbar_p = L.rbar + (size_t)s_hat * plane * C; a.cd_p = L.cd + (size_t)s_hat * plane;
            a.depth = L.depth; a.cd = L.cd; a.remaining = L.remaining; a.winner = ctx->winner;
            a.items = ctx->items; a.count = ctx->count + io.count_slot; a.rowdark = L.rowdark;
            a.dark_lo = L.dark_lo; a.dark_hi = L.dark_hi;
            a.items2 = nullptr; a.count2 = nullptr;
            a.criterion = ctx->criterion; a.disp_thr = P.disp_score_threshold;
            RSLF_TRY(launch_propagate(ctx, C, a));
        }
        ++pass;
        ctx->timing.passes += 1;
    }
    return RSLF_OK;
}

static int begin_run(rslf_ctx* ctx)
{
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    float h2d = ctx->timing.ms_h2d;
    memset(&ctx->timing, 0, sizeof(ctx->timing));
    ctx->timing.ms_h2d = h2d;
    clk_reset(ctx);
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(ctx->count, 0, RSLF_COUNT_SLOTS * sizeof(int), ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(ctx->queue, 0, RSLF_COUNT_SLOTS * sizeof(int), ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(ctx->total_px, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (ctx->dlog_count) RSLF_CUDA_TRY(ctx, cudaMemsetAsync(ctx->dlog_count, 0, sizeof(int), ctx->stream));
    if (ctx->rowwork_cap < (size_t)ctx->V) {
        RSLF_TRY(dev_alloc(ctx, &ctx->rowwork, (size_t)ctx->V));
        ctx->rowwork_cap = (size_t)ctx->V;
    }
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(ctx->rowwork, 0, (size_t)ctx->V * sizeof(unsigned), ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, ctx->stream));
    return RSLF_OK;
}

static int end_run(rslf_ctx* ctx, int D)
{
    RSLF_CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, ctx->stream));
    unsigned long long both[2] = {0, 0};
    int peer_err = 0;
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(both, ctx->total_px, sizeof(both), cudaMemcpyDeviceToHost, ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(&peer_err, ctx->dev_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (peer_err) {
        cudaMemsetAsync(ctx->dev_err, 0, sizeof(int), ctx->stream);
        snprintf(ctx->err, sizeof(ctx->err), "rank %d: a wait for a peer GPU timed out (results of this run are invalid)", ctx->rank);
        return RSLF_ERR_PEER;
    }
    RSLF_CUDA_TRY(ctx, cudaEventElapsedTime(&ctx->timing.ms_total, ctx->ev_a, ctx->ev_b));
    if (ctx->stage_timing) clk_resolve(ctx);
    /* levels that every rank computes whole are counted once (by rank 0), so that the sum over ranks is the job's work */
    const unsigned long long total = both[0] + ((ctx->world <= 1 || ctx->rank == 0) ? both[1] : 0ULL);
    ctx->timing.computed_pixels = (double)total;
    ctx->timing.samples = (double)total * D * ctx->S;
    return RSLF_OK;
}

static int copy_out(rslf_ctx* ctx, void* host, const void* dev, size_t bytes)
{
    if (!host) return RSLF_OK;
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return RSLF_OK;
}

/* ---- host images (cv::Mat storage) <-> dense device arrays ------------------------------------------------
 * A result map is [planes][rows][row_bytes] on the device; the caller's images are `planes` separate host
 * buffers with `step` bytes between rows (cv::Mat::data / step).  Pinned (page-locked) or registered host
 * memory is reached by one strided cudaMemcpy2DAsync per plane at full PCIe rate.  Pageable memory (an ordinary
 * cv::Mat) cannot be the target of an asynchronous DMA: the rows go through a ring of two pinned buffers, the
 * copy engine filling one while the host thread empties the other into the image rows — one CPU pass over the
 * data, overlapped with the transfer (the driver's own pageable path would add a second one in the caller). */
static bool host_is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
/* The ring is per DEVICE and lives as long as the process: page-locking 2 x 64 MB costs tens of milliseconds, and a
 * computer object (hence a context) is typically made per light field.  It is deliberately never released: on B200 /
 * driver 580 cudaFreeHost of it faulted in a process that had had peer (CUDA IPC) mappings to another GPU. */
struct rslf_ring { void* buf[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr}; unsigned next = 0; };
static rslf_ring g_ring[64];
/* Set once a context of this process joins a multi-GPU communicator (peer mappings through CUDA IPC / NCCL): on B200 /
 * driver 580 the ring's page-locked buffers and events then misbehave (cudaFreeHost faulted, an event recorded before
 * the mappings were closed became "invalid device context").  From then on pageable images take the driver's own
 * staged copies; page-locked images (the bench's buffers, rslf_host_alloc) are unaffected: they never use the ring. */
static bool g_no_ring = false;
static void rslf_disable_ring() { g_no_ring = true; }
static int ensure_ring(rslf_ctx* ctx)
{
    rslf_ring& g = g_ring[ctx->device & 63];
    for (int i = 0; i < 2; ++i) {
        if (!g.buf[i] && cudaHostAlloc(&g.buf[i], RSLF_RING_BYTES, cudaHostAllocDefault) != cudaSuccess) {
            g.buf[i] = nullptr;
            snprintf(ctx->err, sizeof(ctx->err), "cudaHostAlloc(%zu) for the transfer ring failed", RSLF_RING_BYTES);
            return RSLF_ERR_NOMEM;
        }
        if (!g.ev[i]) RSLF_CUDA_TRY(ctx, cudaEventCreateWithFlags(&g.ev[i], cudaEventDisableTiming));
        ctx->ring[i] = g.buf[i]; ctx->ring_ev[i] = g.ev[i];
    }
    return RSLF_OK;
}

/* A ring job: consecutive pieces of host images that map to ONE contiguous device range of at most a ring buffer. */
struct ring_piece { char* host; int rows; };
struct ring_job { size_t dev_off, bytes; std::vector<ring_piece> pieces; };

static std::vector<ring_job> ring_jobs(int n_imgs, int rows, size_t row_bytes, void* const* host, size_t step)
{
    std::vector<ring_job> jobs;
    const int max_rows = (int)std::max<size_t>(1, RSLF_RING_BYTES / row_bytes);
    ring_job cur; cur.dev_off = 0; cur.bytes = 0;
    auto flush = [&]() { if (!cur.pieces.empty()) jobs.push_back(cur); cur.pieces.clear(); cur.bytes = 0; };
    for (int p = 0; p < n_imgs; ++p) {
        if (!host[p]) { flush(); continue; }
        for (int r0 = 0; r0 < rows; ) {
            const int room = max_rows - (int)(cur.bytes / row_bytes);
            if (room <= 0) { flush(); continue; }
            const int n = std::min(room, rows - r0);
            if (cur.pieces.empty()) cur.dev_off = ((size_t)p * rows + r0) * row_bytes;
            cur.pieces.push_back({(char*)host[p] + (size_t)r0 * step, n});
            cur.bytes += (size_t)n * row_bytes;
            r0 += n;
        }
    }
    flush();
    return jobs;
}

/* moves the pieces of a job between the images and a ring buffer with a few host threads (pieces split by bytes) */
static void ring_copy(const ring_job& j, char* ring, size_t row_bytes, size_t step, bool to_ring)
{
    int nt = (int)std::min<size_t>(RSLF_COPY_THREADS, std::max<size_t>(1, j.bytes >> 22));
    unsigned hw = std::thread::hardware_concurrency();
    if (hw && (int)hw < nt) nt = (int)hw;
    const long long total_rows = (long long)(j.bytes / row_bytes);
    auto part = [&j, ring, row_bytes, step, to_ring, total_rows](int t, int n) {
        const long long a = total_rows * t / n, b = total_rows * (t + 1) / n;
        long long at = 0;
        for (const ring_piece& pc : j.pieces) {
            const long long lo = std::max(a, at), hi = std::min(b, at + pc.rows);
            for (long long r = lo; r < hi; ++r) {
                char* h = pc.host + (size_t)(r - at) * step;
                char* g = ring + (size_t)r * row_bytes;
                if (to_ring) memcpy(g, h, row_bytes); else memcpy(h, g, row_bytes);
            }
            at += pc.rows;
            if (at >= b) break;
        }
    };
    if (nt <= 1) { part(0, 1); return; }
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(part, t, nt);
    part(0, nt);
    for (auto& x : th) x.join();
}

/* device [planes][rows][row_bytes] -> host planes (nullptr planes are skipped) */
static int planes_to_host(rslf_ctx* ctx, const void* dev, int planes, int rows, size_t row_bytes, void* const* host, size_t step)
{
    if (!host) return RSLF_OK;
    if (step < row_bytes) { snprintf(ctx->err, sizeof(ctx->err), "row step smaller than a row"); return RSLF_ERR_ARG; }
    bool pinned = true;
    for (int p = 0; p < planes && pinned; ++p) if (host[p]) pinned = host_is_pinned(host[p]);
    if (pinned) {
        for (int p = 0; p < planes; ++p)
            if (host[p]) RSLF_CUDA_TRY(ctx, cudaMemcpy2DAsync(host[p], step, (const char*)dev + (size_t)p * rows * row_bytes, row_bytes,
                                                             row_bytes, rows, cudaMemcpyDeviceToHost, ctx->stream));
        return RSLF_OK;
    }
    if (g_no_ring) {
        for (int p = 0; p < planes; ++p)
            if (host[p]) RSLF_CUDA_TRY(ctx, cudaMemcpy2DAsync(host[p], step, (const char*)dev + (size_t)p * rows * row_bytes, row_bytes,
                                                             row_bytes, rows, cudaMemcpyDeviceToHost, ctx->stream));
        return RSLF_OK;
    }
    RSLF_TRY(ensure_ring(ctx));
    const std::vector<ring_job> jobs = ring_jobs(planes, rows, row_bytes, host, step);
    for (size_t i = 0; i < jobs.size() + 2; ++i) {
        const int slot = (int)(i & 1);
        if (i >= 2) {                                                       /* empty the buffer filled two jobs ago */
            RSLF_CUDA_TRY(ctx, cudaEventSynchronize(ctx->ring_ev[slot]));
            ring_copy(jobs[i - 2], (char*)ctx->ring[slot], row_bytes, step, false);
        }
        if (i < jobs.size()) {
            if (i < 2) RSLF_CUDA_TRY(ctx, cudaEventSynchronize(ctx->ring_ev[slot]));   /* an earlier call's transfer may still use the buffer */
            RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ring[slot], (const char*)dev + jobs[i].dev_off, jobs[i].bytes, cudaMemcpyDeviceToHost, ctx->stream));
            RSLF_CUDA_TRY(ctx, cudaEventRecord(ctx->ring_ev[slot], ctx->stream));
        }
    }
    return RSLF_OK;
}

/* host images -> device [n_imgs][rows][row_bytes] */
static int images_to_device(rslf_ctx* ctx, void* dev, int n_imgs, int rows, size_t row_bytes, const void* const* host, size_t step)
{
    bool pinned = g_no_ring;                          /* no ring: the driver stages pageable memory itself */
    if (!pinned) {
        pinned = true;
        for (int p = 0; p < n_imgs && pinned; ++p) pinned = host_is_pinned(host[p]);
    }
    if (pinned) {
        /* one dense copy when the images are continuous and back to back, else one strided copy per image */
        bool dense = (step == row_bytes);
        for (int p = 1; dense && p < n_imgs; ++p) dense = ((const char*)host[p] == (const char*)host[0] + (size_t)p * rows * row_bytes);
        if (dense) RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(dev, host[0], (size_t)n_imgs * rows * row_bytes, cudaMemcpyHostToDevice, ctx->stream));
        else for (int p = 0; p < n_imgs; ++p)
            RSLF_CUDA_TRY(ctx, cudaMemcpy2DAsync((char*)dev + (size_t)p * rows * row_bytes, row_bytes, host[p], step, row_bytes, rows,
                                                 cudaMemcpyHostToDevice, ctx->stream));
        return RSLF_OK;
    }
    RSLF_TRY(ensure_ring(ctx));
    const std::vector<ring_job> jobs = ring_jobs(n_imgs, rows, row_bytes, (void* const*)host, step);
    for (size_t i = 0; i < jobs.size(); ++i) {
        const int slot = (int)(g_ring[ctx->device & 63].next++ & 1u);       /* alternates across calls too (chunked uploads) */
        RSLF_CUDA_TRY(ctx, cudaEventSynchronize(ctx->ring_ev[slot]));       /* its previous transfer (this call's or an earlier one's) has left the buffer */
        ring_copy(jobs[i], (char*)ctx->ring[slot], row_bytes, step, true);
        RSLF_CUDA_TRY(ctx, cudaMemcpyAsync((char*)dev + jobs[i].dev_off, ctx->ring[slot], jobs[i].bytes, cudaMemcpyHostToDevice, ctx->stream));
        RSLF_CUDA_TRY(ctx, cudaEventRecord(ctx->ring_ev[slot], ctx->stream));
    }
    return RSLF_OK;
}

extern "C" void* rslf_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void rslf_host_free(void* p) { if (p) cudaFreeHost(p); }

static int prepare_shards(rslf_ctx* ctx, int levels);

/* ------------------------------------------------------------------ Depth1DComputer_pile */
extern "C" int rslf_cuda_depth1d_pile_run(rslf_ctx* ctx, float dmin, float dmax, int dim_d, int s_hat,
                                          const rslf_params* params)
{
    RSLF_TRY(check_params(ctx, params, dim_d));
    const rslf_params& P = *params;
    RSLF_TRY(begin_run(ctx));
    const int V = ctx->V, U = ctx->U, S = ctx->S, C = ctx->C;
    if (s_hat < 0 || s_hat > S - 1) s_hat = (int)std::floor((0.0 + S) / 2);      /* dc.hpp:489-498 */
    RSLF_TRY(prepare_shards(ctx, 1));
    RSLF_TRY(ensure_scratch(ctx, false, false));
    RSLF_TRY(ensure_level(ctx, 0, V, U, false, false, ctx->v0, ctx->V_total));
    ctx->n_levels = 1;
    rslf_level& L = ctx->lv[0];
    L.replicated = false;
    L.slope = P.slope_factor;
    const void* raw0 = nullptr;
    RSLF_TRY(full_input(ctx, &raw0));
    RSLF_TRY(normalise_level(ctx, 0, raw0, ctx->cv_depth));
    const size_t plane = (size_t)V * U;
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(ctx->pile_depth_raw, 0, plane * sizeof(float), ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.cd, 0, plane * sizeof(float), ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(L.rbar, 0, plane * C * sizeof(float), ctx->stream));
    {
        stage_scope sc(ctx, ST_EDGE);
        RSLF_TRY(launch_edge_confidence(ctx, L.epi, V, S, U, C, s_hat, 1, P, L.ce, L.emask, nullptr, L.remaining,
                                        ctx->world > 1 ? L.v0 : 0, ctx->world > 1 ? L.Vtot - L.v0 - V : 0));
    }
    pass_io io;
    io.level = 0; io.s_hat = s_hat; io.D = dim_d; io.dmin = dmin; io.dmax = dmax; io.use_bound_maps = false;
    io.pile = true; io.count_slot = 0; io.bounds_in_range = false;
    RSLF_TRY(run_depth_pass(ctx, P, io));
    ctx->timing.passes = 1; ctx->timing.levels = 1;
    ctx->last_kind = 1; ctx->pile_s_hat = s_hat;
    return end_run(ctx, dim_d);
}

extern "C" int rslf_cuda_depth1d_pile_get(rslf_ctx* ctx, float* best_depth_vu, float* edge_conf_vu,
                                          uint8_t* edge_mask_vu, float* disp_conf_vu, float* rbar_vuc)
{
    if (!ctx) return RSLF_ERR_ARG;
    if (ctx->last_kind != 1) { snprintf(ctx->err, sizeof(ctx->err), "no pile result"); return RSLF_ERR_STATE; }
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    rslf_level& L = ctx->lv[0];
    const size_t plane = (size_t)L.V * L.U;
    cudaEventRecord(ctx->ev_a, ctx->stream);
    RSLF_TRY(copy_out(ctx, best_depth_vu, ctx->filtered, plane * 4));          /* the filtered map (core.hpp:892) */
    RSLF_TRY(copy_out(ctx, edge_conf_vu, L.ce, plane * 4));
    RSLF_TRY(copy_out(ctx, edge_mask_vu, L.emask, plane));
    RSLF_TRY(copy_out(ctx, disp_conf_vu, L.cd, plane * 4));
    RSLF_TRY(copy_out(ctx, rbar_vuc, L.rbar, plane * 4 * ctx->C));
    cudaEventRecord(ctx->ev_b, ctx->stream);
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.ms_d2h, ctx->ev_a, ctx->ev_b);
    return RSLF_OK;
}

extern "C" int rslf_cuda_depth1d_pile(rslf_ctx* ctx, float dmin, float dmax, int dim_d, int s_hat,
                                      const rslf_params* params, float* best_depth_vu, float* edge_conf_vu,
                                      uint8_t* edge_mask_vu, float* disp_conf_vu, float* rbar_vuc)
{
    RSLF_TRY(rslf_cuda_depth1d_pile_run(ctx, dmin, dmax, dim_d, s_hat, params));
    return rslf_cuda_depth1d_pile_get(ctx, best_depth_vu, edge_conf_vu, edge_mask_vu, disp_conf_vu, rbar_vuc);
}

/* unfiltered argmax depths of the last pile run (diagnostic; m_best_depth_v_u before core.hpp:881-892) */
extern "C" int rslf_cuda_depth1d_pile_get_raw_depth(rslf_ctx* ctx, float* raw_depth_vu)
{
    if (!ctx || !raw_depth_vu) return RSLF_ERR_ARG;
    if (ctx->last_kind != 1) return RSLF_ERR_STATE;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_TRY(copy_out(ctx, raw_depth_vu, ctx->pile_depth_raw, (size_t)ctx->lv[0].V * ctx->lv[0].U * 4));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RSLF_OK;
}

/* ------------------------------------------------------------------ Depth2DComputer */
extern "C" int rslf_cuda_depth2d_run(rslf_ctx* ctx, float dmin, float dmax, int dim_d, const rslf_params* params,
                                     const float* dmin_svu, const float* dmax_svu)
{
    RSLF_TRY(check_params(ctx, params, dim_d));
    if ((dmin_svu == nullptr) != (dmax_svu == nullptr)) return RSLF_ERR_ARG;
    const rslf_params& P = *params;
    const int V = ctx->V, U = ctx->U, S = ctx->S;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_TRY(prepare_shards(ctx, 1));
    RSLF_TRY(ensure_scratch(ctx, true, false));
    RSLF_TRY(ensure_level(ctx, 0, V, U, true, dmin_svu != nullptr, ctx->v0, ctx->V_total));
    ctx->n_levels = 1;
    rslf_level& L = ctx->lv[0];
    L.replicated = false;
    const size_t px = (size_t)S * V * U;
    L.bounds_const = false;
    if (dmin_svu) {
        RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(L.dmin, dmin_svu, px * 4, cudaMemcpyHostToDevice, ctx->stream));
        RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(L.dmax, dmax_svu, px * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    const void* raw0 = nullptr;
    RSLF_TRY(full_input(ctx, &raw0));
    RSLF_TRY(begin_run(ctx));
    RSLF_TRY(normalise_level(ctx, 0, raw0, ctx->cv_depth));
    RSLF_TRY(run_depth2d_level(ctx, 0, P, dmin, dmax, dim_d, dmin_svu != nullptr, dmin_svu == nullptr));   /* user-edited bounds may leave [dmin, dmax] */
    ctx->timing.levels = 1;
    ctx->last_kind = 2;
    return end_run(ctx, dim_d);
}

static int get_level_maps(rslf_ctx* ctx, int p, float* best_depth, float* edge_conf, uint8_t* edge_mask,
                          float* disp_conf, float* rbar, float* dmin_svu, float* dmax_svu)
{
    rslf_level& L = ctx->lv[p];
    const size_t px = (size_t)ctx->S * L.V * L.U;
    cudaEventRecord(ctx->ev_a, ctx->stream);
    RSLF_TRY(copy_out(ctx, best_depth, L.depth, px * 4));
    RSLF_TRY(copy_out(ctx, edge_conf, L.ce, px * 4));
    RSLF_TRY(copy_out(ctx, edge_mask, L.emask, px));
    RSLF_TRY(copy_out(ctx, disp_conf, L.cd, px * 4));
    RSLF_TRY(copy_out(ctx, rbar, L.rbar, px * 4 * ctx->C));
    if (L.dmin) {
        if (L.bounds_const && (dmin_svu || dmax_svu)) {
            fill_f32_kernel<<<stream_grid(ctx, px), 256, 0, ctx->stream>>>(L.dmin, px, L.bounds_lo);
            fill_f32_kernel<<<stream_grid(ctx, px), 256, 0, ctx->stream>>>(L.dmax, px, L.bounds_hi);
            L.bounds_const = false;
        }
        RSLF_TRY(copy_out(ctx, dmin_svu, L.dmin, px * 4)); RSLF_TRY(copy_out(ctx, dmax_svu, L.dmax, px * 4));
    }
    cudaEventRecord(ctx->ev_b, ctx->stream);
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.ms_d2h, ctx->ev_a, ctx->ev_b);
    return RSLF_OK;
}

extern "C" int rslf_cuda_depth2d_get(rslf_ctx* ctx, float* best_depth_svu, float* edge_conf_svu,
                                     uint8_t* edge_mask_svu, float* disp_conf_svu, float* rbar_svuc)
{
    if (!ctx) return RSLF_ERR_ARG;
    if (ctx->last_kind != 2) { snprintf(ctx->err, sizeof(ctx->err), "no depth2d result"); return RSLF_ERR_STATE; }
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return get_level_maps(ctx, 0, best_depth_svu, edge_conf_svu, edge_mask_svu, disp_conf_svu, rbar_svuc, nullptr, nullptr);
}

/* Results of the last depth2d / fine-to-coarse-level run straight into S host images per map (cv::Mat::data, step) */
static int level_maps_to_mats(rslf_ctx* ctx, int p, float* const* best_depth, float* const* edge_conf, uint8_t* const* edge_mask,
                              float* const* disp_conf, float* const* rbar, size_t step_f32, size_t step_u8, size_t step_rbar)
{
    rslf_level& L = ctx->lv[p];
    const int S = ctx->S, V = L.V, U = L.U, C = ctx->C;
    cudaEventRecord(ctx->ev_a, ctx->stream);
    RSLF_TRY(planes_to_host(ctx, L.depth, S, V, (size_t)U * 4, (void* const*)best_depth, step_f32));
    RSLF_TRY(planes_to_host(ctx, L.ce, S, V, (size_t)U * 4, (void* const*)edge_conf, step_f32));
    RSLF_TRY(planes_to_host(ctx, L.emask, S, V, (size_t)U, (void* const*)edge_mask, step_u8));
    RSLF_TRY(planes_to_host(ctx, L.cd, S, V, (size_t)U * 4, (void* const*)disp_conf, step_f32));
    RSLF_TRY(planes_to_host(ctx, L.rbar, S, V, (size_t)U * 4 * C, (void* const*)rbar, step_rbar));
    cudaEventRecord(ctx->ev_b, ctx->stream);
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.ms_d2h, ctx->ev_a, ctx->ev_b);
    return RSLF_OK;
}

extern "C" int rslf_cuda_depth2d_get_mats(rslf_ctx* ctx, float* const* best_depth_s, float* const* edge_conf_s,
                                          uint8_t* const* edge_mask_s, float* const* disp_conf_s, float* const* rbar_s,
                                          size_t step_f32, size_t step_u8, size_t step_rbar)
{
    if (!ctx) return RSLF_ERR_ARG;
    if (ctx->last_kind != 2) { snprintf(ctx->err, sizeof(ctx->err), "no depth2d result"); return RSLF_ERR_STATE; }
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return level_maps_to_mats(ctx, 0, best_depth_s, edge_conf_s, edge_mask_s, disp_conf_s, rbar_s, step_f32, step_u8, step_rbar);
}

extern "C" int rslf_cuda_depth2d(rslf_ctx* ctx, float dmin, float dmax, int dim_d, const rslf_params* params,
                                 const float* dmin_svu, const float* dmax_svu, float* best_depth_svu,
                                 float* edge_conf_svu, uint8_t* edge_mask_svu, float* disp_conf_svu, float* rbar_svuc)
{
    RSLF_TRY(rslf_cuda_depth2d_run(ctx, dmin, dmax, dim_d, params, dmin_svu, dmax_svu));
    return rslf_cuda_depth2d_get(ctx, best_depth_svu, edge_conf_svu, edge_mask_svu, disp_conf_svu, rbar_svuc);
}

extern "C" int rslf_cuda_depth2d_get_valid_mask(rslf_ctx* ctx, int accept_all, const rslf_params* params,
                                                uint8_t* valid_svu)
{
    if (!ctx || !params || !valid_svu) return RSLF_ERR_ARG;
    if (ctx->last_kind != 2) return RSLF_ERR_STATE;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    rslf_level& L = ctx->lv[0];
    const size_t px = (size_t)ctx->S * L.V * L.U;
    valid_mask_kernel<<<stream_grid(ctx, px), 256, 0, ctx->stream>>>(L.ce, ctx->criterion == 1 ? L.cd : L.ce, px,
        ctx->criterion == 1 ? params->disp_score_threshold : params->edge_score_threshold, accept_all, L.valid);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    RSLF_TRY(copy_out(ctx, valid_svu, L.valid, px));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RSLF_OK;
}

/* ------------------------------------------------------------------ FineToCoarse */
static inline int cv_round_half(int n) { return (int)std::nearbyint(n * 0.5); }   /* cvRound: half to even */

static int launch_downsample(rslf_ctx* ctx, const float* in, int V, int S, int U, int C, float* out, int V2, int U2,
                             int ov_begin = 0, int ov_count = -1)
{
    if (ov_count < 0) ov_count = V2;
    dim3 grid(rslf_div_up(U2, DS_TU), rslf_div_up(ov_count, DS_TV), S);
    if (C == 1) downsample_kernel<1><<<grid, DS_THREADS, 0, ctx->stream>>>(in, V, S, U, out, V2, U2, ov_begin, ov_count);
    else downsample_kernel<3><<<grid, DS_THREADS, 0, ctx->stream>>>(in, V, S, U, out, V2, U2, ov_begin, ov_count);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    return RSLF_OK;
}

template <typename T>
static int launch_downsample_int(rslf_ctx* ctx, const T* in, int V, int S, int U, int C, T* out, int V2, int U2,
                                 int ov_begin = 0, int ov_count = -1)
{
    if (ov_count < 0) ov_count = V2;
    dim3 grid(rslf_div_up(U2, DS_TU), rslf_div_up(ov_count, DS_TV), S);
    if (C == 1) downsample_int_kernel<T, 1><<<grid, DS_THREADS, 0, ctx->stream>>>(in, V, S, U, out, V2, U2, ov_begin, ov_count);
    else downsample_int_kernel<T, 3><<<grid, DS_THREADS, 0, ctx->stream>>>(in, V, S, U, out, V2, U2, ov_begin, ov_count);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    return RSLF_OK;
}

/* Rows of level p (Vs rows) that rank r's rows [y0, y1) of the finer level (Vd rows) read when upsampling: all inside
 * its own block of level p plus FUSE_HALO rows either side? (same arithmetic as the kernel: fuse_src_rows) */
static bool fuse_halo_suffices(const shard_tab& coarse, const shard_tab& fine, int Vs, int Vd)
{
    for (int r = 0; r < fine.n; ++r) {
        const int lo = coarse.b[r] - FUSE_HALO, hi = coarse.b[r + 1] + FUSE_HALO;       /* rows available: [lo, hi) */
        if (coarse.b[r + 1] - coarse.b[r] < FUSE_HALO) return false;                     /* the neighbour sends what it owns */
        for (int y = fine.b[r]; y < fine.b[r + 1]; ++y) {
            int iy0, iy1, ny; float fy;
            fuse_src_rows(y, Vs, Vd, &iy0, &iy1, &fy, &ny);
            if (iy0 < lo || iy1 >= hi || ny < lo || ny >= hi) return false;
        }
    }
    return true;
}

/* fuse_disp_maps (ftc_core.cpp:69-135).  Vp: GLOBAL rows per level; disp / valid / outputs of a sharded level hold
 * the rank's rows (tabs[p]); a replicated level (rep[p]) is whole on every rank.  Across the block borders the
 * upsampling and the final 3x3 median read FUSE_HALO rows of the neighbours, sent through peer memory
 * (comm_svu_halo); without peer memory, or when a block is thinner than the halo, the maps are all-gathered. */
static int launch_fuse(rslf_ctx* ctx, int levels, const int* Vp, const int* Up, const shard_tab* tabs, const bool* rep,
                       const float* const* disp, const uint8_t* const* valid, float* out_map, uint8_t* out_valid)
{
    const int S = ctx->S, r = ctx->rank;
    auto whole = [&](int p) { return ctx->world <= 1 || rep[p]; };
    const bool p2p = ctx->world > 1 && ctx->p2p_state > 0 && levels <= FUSE_SLOTS + 64 && getenv("RSLF_FUSE_GATHER") == nullptr;
    if (p2p) ++ctx->fuse_seq;
    /* the view of a level's (fused) map for the kernels that read across block borders */
    auto view_of = [&](int p, const float* m, const uint8_t* k, bool need_halo_ok, svu_view<float>* vm, svu_view<uint8_t>* vk, int slot) -> int {
        if (whole(p)) { *vm = svu_whole(m, Vp[p]); if (vk) *vk = svu_whole(k, Vp[p]); return RSLF_OK; }
        if (p2p && need_halo_ok && p < FUSE_SLOTS) return comm_svu_halo(ctx, p, m, k, S, Up[p], tabs[p], vm, vk);
        const float* gm = nullptr; const uint8_t* gk = nullptr;
        RSLF_TRY(global_svu_f32(ctx, m, S, Up[p], tabs[p], slot, &gm));
        *vm = svu_whole(gm, Vp[p]);
        if (vk) { RSLF_TRY(global_svu_u8(ctx, k, S, Up[p], tabs[p], slot, &gk)); *vk = svu_whole(gk, Vp[p]); }
        return RSLF_OK;
    };
    const float* map_cur = disp[levels - 1]; const uint8_t* mask_cur = valid[levels - 1];
    float* fa = ctx->fuse_a; float* fb = ctx->fuse_b; uint8_t* ma = ctx->fuse_ma; uint8_t* mb = ctx->fuse_mb;
    for (int p = levels - 1; p > 0; --p) {
        const int Vn = Vp[p - 1], Un = Up[p - 1];
        const int y0 = whole(p - 1) ? 0 : tabs[p - 1].b[r], Vn_loc = whole(p - 1) ? Vn : tabs[p - 1].b[r + 1] - y0;
        svu_view<float> vm; svu_view<uint8_t> vk;
        const bool halo_ok = !whole(p) && !whole(p - 1) && fuse_halo_suffices(tabs[p], tabs[p - 1], Vp[p], Vn);
        RSLF_TRY(view_of(p, map_cur, mask_cur, halo_ok, &vm, &vk, p & 1));
        dim3 grid(rslf_div_up(Un, 128), Vn_loc, S);
        uint8_t* mout = (p == 1) ? out_valid : ma;
        fuse_level_kernel<<<grid, 128, 0, ctx->stream>>>(vm, vk, Vp[p], Up[p], disp[p - 1], valid[p - 1], Vn, Un, fa, mout, y0, Vn_loc);
        RSLF_CUDA_TRY(ctx, cudaGetLastError());
        ctx->timing.kernel_launches += 1;
        map_cur = fa; mask_cur = mout;
        std::swap(fa, fb); std::swap(ma, mb);
    }
    const int v0 = whole(0) ? 0 : tabs[0].b[r], V_loc = whole(0) ? Vp[0] : tabs[0].b[r + 1] - v0;
    const size_t px = (size_t)S * V_loc * Up[0];
    if (levels == 1) RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(out_valid, valid[0], px, cudaMemcpyDeviceToDevice, ctx->stream));
    /* the final 3x3 median reads +-1 row */
    svu_view<float> vm;
    bool thick = true;
    if (!whole(0)) for (int q = 0; q < tabs[0].n; ++q) thick = thick && (tabs[0].b[q + 1] - tabs[0].b[q] >= FUSE_HALO);
    RSLF_TRY(view_of(0, map_cur, nullptr, thick, &vm, nullptr, 0));
    dim3 grid(rslf_div_up(Up[0], 128), V_loc, S);
    median3x3_kernel<<<grid, 128, 0, ctx->stream>>>(vm, Vp[0], Up[0], out_map, v0, V_loc);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    return RSLF_OK;
}

static shard_tab single_tab(int V) { shard_tab t; t.n = 1; t.b[0] = 0; t.b[1] = V; return t; }

/* Row-shard bookkeeping of a run: sets the level-0 table for a single rank, checks a multi-rank one. */
static int prepare_shards(rslf_ctx* ctx, int levels)
{
    if (ctx->world <= 1) {
        ctx->row_starts[0] = 0; ctx->row_starts[1] = ctx->V; ctx->v0 = 0; ctx->V_total = ctx->V; ctx->rank = 0;
        return RSLF_OK;
    }
    if (!ctx->have_shards) { snprintf(ctx->err, sizeof(ctx->err), "multi-rank run without rslf_cuda_set_row_shards"); return RSLF_ERR_STATE; }
    if (ctx->row_starts[ctx->rank + 1] - ctx->row_starts[ctx->rank] != ctx->V) {
        snprintf(ctx->err, sizeof(ctx->err), "uploaded %d rows but the shard table gives this rank %d", ctx->V,
                 ctx->row_starts[ctx->rank + 1] - ctx->row_starts[ctx->rank]);
        return RSLF_ERR_ARG;
    }
    (void)levels;
    return RSLF_OK;
}

extern "C" int rslf_cuda_fine_to_coarse_run(rslf_ctx* ctx, float dmin, float dmax, int dim_d,
                                            const rslf_params* params, int max_pyr_depth, int accept_all_last_scale)
{
    RSLF_TRY(check_params(ctx, params, dim_d));
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int S = ctx->S, C = ctx->C, r = ctx->world > 1 ? ctx->rank : 0;
    /* pyramid sizes from the GLOBAL field: while (V > 10 && U > 10 && depth < max) (ftc.hpp:130), dsize = cvRound(n * 0.5) */
    int Vp[RSLF_MAX_LEVELS], Up[RSLF_MAX_LEVELS];
    int levels = 0;
    {
        int V = (ctx->world > 1 && ctx->have_shards) ? ctx->V_total : ctx->V, U = ctx->U;
        int maxd = (max_pyr_depth < 1) ? RSLF_MAX_LEVELS : std::min(max_pyr_depth, RSLF_MAX_LEVELS);
        while (V > 10 && U > 10 && levels < maxd) {
            Vp[levels] = V; Up[levels] = U; ++levels;
            V = cv_round_half(V); U = cv_round_half(U);
        }
    }
    if (levels == 0) { snprintf(ctx->err, sizeof(ctx->err), "light field smaller than the minimum pyramid size"); return RSLF_ERR_ARG; }
    RSLF_TRY(prepare_shards(ctx, levels));
    /* Multi-rank runs shard the fine levels by rows and REPLICATE the coarse ones: level p can be sharded while
     * every boundary is a multiple of 2^p (so that levels split at the same image position and the bound
     * propagation rows 2v, 2v+1 -> v stay local) and every rank keeps >= 2 rows; the remaining levels are a
     * few thousand pixels and are computed whole by every rank (identical results), which avoids both thin
     * blocks and coarse alignment of the level-0 shards (load balance). */
    shard_tab tabs[RSLF_MAX_LEVELS];
    bool rep[RSLF_MAX_LEVELS];
    int Vl[RSLF_MAX_LEVELS];                                  /* rows of the level held by this rank */
    int sharded_levels = levels;
    if (ctx->world > 1) {
        sharded_levels = 1;
        while (sharded_levels < levels) {
            bool ok = true;
            for (int q = 1; q < ctx->world; ++q) ok = ok && (ctx->row_starts[q] % (1 << sharded_levels) == 0);
            const shard_tab t = level_shards(ctx, sharded_levels, Vp[sharded_levels]);
            for (int q = 0; q < ctx->world; ++q) ok = ok && (t.b[q + 1] - t.b[q] >= 2);
            if (!ok) break;
            ++sharded_levels;
        }
    }
    size_t max_plane = 0;
    for (int p = 0; p < levels; ++p) {
        rep[p] = (ctx->world > 1 && p >= sharded_levels);
        tabs[p] = rep[p] ? single_tab(Vp[p]) : level_shards(ctx, p, Vp[p]);
        Vl[p] = rep[p] ? Vp[p] : tabs[p].b[r + 1] - tabs[p].b[r];
        if (Vl[p] < 1) { snprintf(ctx->err, sizeof(ctx->err), "rank %d would hold no row of pyramid level %d; use fewer ranks", r, p); return RSLF_ERR_ARG; }
        max_plane = std::max(max_plane, (size_t)Vl[p] * Up[p]);
    }
    /* the first replicated level derives its bounds from the gathered maps of the last sharded level */
    if (ctx->world > 1 && sharded_levels < levels)
        max_plane = std::max(max_plane, (size_t)Vp[sharded_levels - 1] * Up[sharded_levels - 1]);
    RSLF_TRY(ensure_scratch(ctx, true, true, max_plane));
    for (int p = 0; p < levels; ++p) {
        RSLF_TRY(ensure_level(ctx, p, Vl[p], Up[p], true, true, rep[p] ? 0 : tabs[p].b[r], Vp[p]));
        ctx->lv[p].replicated = rep[p];
    }
    if (ctx->world > 1) {
        const size_t gpx = (size_t)S * Vp[0] * Up[0];
        if (ctx->g_map_cap < gpx) {
            RSLF_TRY(dev_alloc(ctx, &ctx->g_map, gpx)); RSLF_TRY(dev_alloc(ctx, &ctx->g_map2, gpx));
            RSLF_TRY(dev_alloc(ctx, &ctx->g_mk, gpx)); RSLF_TRY(dev_alloc(ctx, &ctx->g_mk2, gpx));
            ctx->g_map_cap = gpx;
        }
    }
    ctx->n_levels = levels;
    const void* raw0 = nullptr;
    RSLF_TRY(full_input(ctx, &raw0));
    RSLF_TRY(begin_run(ctx));
    const rslf_params& P0 = *params;
    for (int p = 0; p < levels; ++p) {
        rslf_level& L = ctx->lv[p];
        const size_t px = (size_t)S * Vl[p] * Up[p];
        const void* raw = (p == 0) ? raw0 : (const void*)L.raw;
        /* 8-bit / 16-bit stacks keep their depth between levels (OpenCV's integer blur / resize), float stacks stay float */
        RSLF_TRY(normalise_level(ctx, p, raw, ctx->cv_depth));
        rslf_params P = P0;
        P.slope_factor = (float)((0.0 + Up[p]) / Up[0]);                        /* ftc.hpp:139 */
        L.slope = P.slope_factor;
        /* level 0 has the global bounds everywhere (ftc.hpp:160-168): they travel as constants in the work records; the
         * maps are only filled if somebody asks for them (rslf_cuda_fine_to_coarse_get_level) */
        if (p == 0) L.bounds_const = true, L.bounds_lo = dmin, L.bounds_hi = dmax;
        else L.bounds_const = false;
        RSLF_TRY(run_depth2d_level(ctx, p, P, dmin, dmax, dim_d, p > 0));
        {
            stage_scope sc(ctx, ST_PYR);
            const int accept_all = (accept_all_last_scale && p == levels - 1) ? 1 : 0;   /* ftc.hpp:157-158 */
            valid_mask_kernel<<<stream_grid(ctx, px), 256, 0, ctx->stream>>>(L.ce, ctx->criterion == 1 ? L.cd : L.ce, px,
                ctx->criterion == 1 ? P.disp_score_threshold : P.edge_score_threshold, accept_all, L.valid);
            ctx->timing.kernel_launches += 1;
            if (p + 1 < levels) {
                rslf_level& N = ctx->lv[p + 1];
                const bool sh = (ctx->world > 1 && !rep[p]);             /* level p is sharded */
                /* next level's raw stack (ftc.hpp:146): every rank holds all rows of every level's stack, so the 7x7 blur
                 * and the half-scale resize run on the whole level without any exchange */
                const int ov0 = rep[p + 1] ? 0 : tabs[p + 1].b[r];
                if (ctx->cv_depth == RSLF_DEPTH_8U)
                    RSLF_TRY(launch_downsample_int<uint8_t>(ctx, (const uint8_t*)raw, Vp[p], S, Up[p], C, (uint8_t*)N.raw, Vp[p + 1], Up[p + 1]));
                else if (ctx->cv_depth == RSLF_DEPTH_16U)
                    RSLF_TRY(launch_downsample_int<uint16_t>(ctx, (const uint16_t*)raw, Vp[p], S, Up[p], C, (uint16_t*)N.raw, Vp[p + 1], Up[p + 1]));
                else
                    RSLF_TRY(launch_downsample(ctx, (const float*)raw, Vp[p], S, Up[p], C, N.raw, Vp[p + 1], Up[p + 1]));
                /* bounds of level p+1 from level p (ftc.hpp:201-294): rows 2v, 2v+1 are local when both levels are
                 * sharded (aligned blocks); the first replicated level reads the gathered maps of level p */
                const float* b_depth = L.depth; const uint8_t* b_valid = L.valid;
                int b_rows = Vl[p], b_v0 = sh ? tabs[p].b[r] : 0;
                if (sh && rep[p + 1]) {
                    RSLF_TRY(global_svu_f32(ctx, L.depth, S, Up[p], tabs[p], 0, &b_depth));
                    RSLF_TRY(global_svu_u8(ctx, L.valid, S, Up[p], tabs[p], 0, &b_valid));
                    b_rows = Vp[p]; b_v0 = 0;
                }
                const int rows = S * b_rows;
                nearest_valid_kernel<<<rslf_div_up((long long)rows * 32, 256), 256, 0, ctx->stream>>>(b_valid, rows, Up[p], ctx->nearest_l, ctx->nearest_r);
                dim3 grid(rslf_div_up(Up[p + 1], 128), Vl[p + 1], S);
                set_bounds_kernel<<<grid, 128, 0, ctx->stream>>>(b_depth, ctx->nearest_l, ctx->nearest_r, S, b_rows, Up[p],
                                                                 Vl[p + 1], Up[p + 1], N.dmin, N.dmax, Vp[p], b_v0, ov0, 1, dmin, dmax);
                ctx->timing.kernel_launches += 2;
            }
            RSLF_CUDA_TRY(ctx, cudaGetLastError());
        }
        ctx->timing.levels += 1;
    }
    {
        stage_scope sc(ctx, ST_PYR);
        const float* dp[RSLF_MAX_LEVELS]; const uint8_t* vp[RSLF_MAX_LEVELS];
        for (int p = 0; p < levels; ++p) { dp[p] = ctx->lv[p].depth; vp[p] = ctx->lv[p].valid; }
        RSLF_TRY(launch_fuse(ctx, levels, Vp, Up, tabs, rep, dp, vp, ctx->out_map, ctx->out_valid));
    }
    ctx->last_kind = 3;
    return end_run(ctx, dim_d);
}

extern "C" int rslf_cuda_fine_to_coarse_get(rslf_ctx* ctx, float* out_map_svu, uint8_t* out_valid_svu)
{
    if (!ctx) return RSLF_ERR_ARG;
    if (ctx->last_kind != 3) { snprintf(ctx->err, sizeof(ctx->err), "no fine-to-coarse result"); return RSLF_ERR_STATE; }
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t px = (size_t)ctx->S * ctx->V * ctx->U;
    cudaEventRecord(ctx->ev_a, ctx->stream);
    RSLF_TRY(copy_out(ctx, out_map_svu, ctx->out_map, px * 4));
    RSLF_TRY(copy_out(ctx, out_valid_svu, ctx->out_valid, px));
    cudaEventRecord(ctx->ev_b, ctx->stream);
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.ms_d2h, ctx->ev_a, ctx->ev_b);
    return RSLF_OK;
}

extern "C" int rslf_cuda_fine_to_coarse_get_mats(rslf_ctx* ctx, float* const* out_map_s, size_t map_step,
                                                 uint8_t* const* out_valid_s, size_t valid_step)
{
    if (!ctx) return RSLF_ERR_ARG;
    if (ctx->last_kind != 3) { snprintf(ctx->err, sizeof(ctx->err), "no fine-to-coarse result"); return RSLF_ERR_STATE; }
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaEventRecord(ctx->ev_a, ctx->stream);
    RSLF_TRY(planes_to_host(ctx, ctx->out_map, ctx->S, ctx->V, (size_t)ctx->U * 4, (void* const*)out_map_s, map_step));
    RSLF_TRY(planes_to_host(ctx, ctx->out_valid, ctx->S, ctx->V, (size_t)ctx->U, (void* const*)out_valid_s, valid_step));
    cudaEventRecord(ctx->ev_b, ctx->stream);
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaEventElapsedTime(&ctx->timing.ms_d2h, ctx->ev_a, ctx->ev_b);
    return RSLF_OK;
}

/* The normalised EPIs the run worked on (Depth2DComputer::get_epis, dc.hpp:207; level 0 of a pyramid): V host
 * images of S rows x U x C float32 */
extern "C" int rslf_cuda_get_epis(rslf_ctx* ctx, float* const* epi_v, size_t step)
{
    if (!ctx || !epi_v) return RSLF_ERR_ARG;
    if (ctx->last_kind == 0 || !ctx->lv[0].epi) { snprintf(ctx->err, sizeof(ctx->err), "no run yet"); return RSLF_ERR_STATE; }
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_TRY(planes_to_host(ctx, ctx->lv[0].epi, ctx->lv[0].V, ctx->S, (size_t)ctx->lv[0].U * ctx->C * 4, (void* const*)epi_v, step));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RSLF_OK;
}

extern "C" int rslf_cuda_fine_to_coarse(rslf_ctx* ctx, float dmin, float dmax, int dim_d, const rslf_params* params,
                                        int max_pyr_depth, int accept_all_last_scale, float* out_map_svu,
                                        uint8_t* out_valid_svu)
{
    RSLF_TRY(rslf_cuda_fine_to_coarse_run(ctx, dmin, dmax, dim_d, params, max_pyr_depth, accept_all_last_scale));
    return rslf_cuda_fine_to_coarse_get(ctx, out_map_svu, out_valid_svu);
}

/* FineToCoarse::get_coloured_depth_maps (ftc.hpp:324-377) of the last fine-to-coarse run; see k_colour.cuh */
extern "C" int rslf_cuda_fine_to_coarse_get_coloured(rslf_ctx* ctx, const uint8_t* lut_bgr_256x3, int saturate,
                                                     const rslf_params* params, uint8_t* out_bgr_svu3, double* fit_min_max)
{
    if (!ctx || !lut_bgr_256x3 || !params || !out_bgr_svu3) return RSLF_ERR_ARG;
    if (ctx->last_kind != 3) { snprintf(ctx->err, sizeof(ctx->err), "no fine-to-coarse result"); return RSLF_ERR_STATE; }
    if (!saturate) {
        snprintf(ctx->err, sizeof(ctx->err), "ImageConverter_uchar::fit without saturation (mean + 12 std) is not implemented");
        return RSLF_ERR_UNSUPPORTED;
    }
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int S = ctx->S, V = ctx->V, U = ctx->U, C = ctx->C;
    const size_t plane = (size_t)V * U, px = plane * S;
    /* row-sharded run: every rank colours its rows; the quantile fit is over the WHOLE plane of view s_fit: the
     * key histograms of the radix select are summed over the ranks (NCCL all-reduce), the ranks of the quantiles
     * count in the whole plane */
    const bool sharded = ctx->world > 1 && ctx->have_shards;
    const size_t plane_tot = sharded ? (size_t)ctx->V_total * U : plane;
    const int s_fit = (int)std::round(S / 2.0);                       /* ftc.hpp:345 */
    if (s_fit >= S) { snprintf(ctx->err, sizeof(ctx->err), "S = %d: the reference fits on view round(S / 2) = %d, which does not exist", S, s_fit); return RSLF_ERR_ARG; }
    if (V > 65535 || S > 65535) return RSLF_ERR_UNSUPPORTED;
    if (!ctx->colour_hist) RSLF_TRY(dev_alloc(ctx, &ctx->colour_hist, (size_t)3 * COLOUR_BINS));
    if (!ctx->colour_lut) RSLF_TRY(dev_alloc(ctx, &ctx->colour_lut, (size_t)768));
    unsigned* h_hi = ctx->colour_hist; unsigned* h_a = h_hi + COLOUR_BINS; unsigned* h_b = h_a + COLOUR_BINS;
    const float* fit_plane = ctx->out_map + (size_t)s_fit * plane;
    std::vector<unsigned> host((size_t)2 * COLOUR_BINS);
    /* rslf_plot.cpp:73-79: ranks floor(0.02 N) and floor(0.98 N) of the sorted plane */
    const size_t k_min = (size_t)std::floor(0.02 * (double)plane_tot), k_max = (size_t)std::floor(0.98 * (double)plane_tot);
    RSLF_CUDA_TRY(ctx, cudaMemsetAsync(h_hi, 0, (size_t)3 * COLOUR_BINS * sizeof(unsigned), ctx->stream));
    colour_hist_hi_kernel<<<stream_grid(ctx, plane), 256, 0, ctx->stream>>>(fit_plane, plane, h_hi);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    if (sharded) RSLF_TRY(comm_allreduce_sum_u32(ctx, h_hi, COLOUR_BINS));
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(host.data(), h_hi, COLOUR_BINS * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned hi_a = 0, hi_b = 0, lo_a = 0, lo_b = 0; size_t r_a = 0, r_b = 0, dummy = 0;
    if (!colour_find_bin(host.data(), k_min, &hi_a, &r_a) || !colour_find_bin(host.data(), k_max, &hi_b, &r_b)) {
        snprintf(ctx->err, sizeof(ctx->err), "coloured maps: histogram does not cover the plane"); return RSLF_ERR_CUDA;
    }
    colour_hist_lo_kernel<<<stream_grid(ctx, plane), 256, 0, ctx->stream>>>(fit_plane, plane, hi_a, hi_b, h_a, h_b);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    if (sharded) RSLF_TRY(comm_allreduce_sum_u32(ctx, h_a, (size_t)2 * COLOUR_BINS));
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(host.data(), h_a, (size_t)2 * COLOUR_BINS * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (!colour_find_bin(host.data(), r_a, &lo_a, &dummy) || !colour_find_bin(host.data() + COLOUR_BINS, r_b, &lo_b, &dummy)) {
        snprintf(ctx->err, sizeof(ctx->err), "coloured maps: second-level histogram does not cover the bin"); return RSLF_ERR_CUDA;
    }
    const double mn = colour_key_to_float((hi_a << 16) | lo_a), mx = colour_key_to_float((hi_b << 16) | lo_b);
    if (fit_min_max) { fit_min_max[0] = mn; fit_min_max[1] = mx; }
    /* rslf_plot.cpp:105-106: float alpha = 255.0 / (max - min); convertTo(dst, CV_8U, alpha, -alpha * min) */
    const float alpha = (float)(255.0 / (mx - mn));
    const double beta = -alpha * mn;
    const float a = alpha, b = (float)beta;
    const float shadow = (float)(0.05 * 1.73205080757);               /* _SHADOW_NORMALIZED_LEVEL (core.hpp:30), compared in float */
    const double shadow_T = rslf_sq_threshold(shadow);
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->colour_lut, lut_bgr_256x3, 768, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* d_out = reinterpret_cast<uint8_t*>(ctx->fuse_a);          /* scratch of the fusion: 4 B per pixel >= 3 */
    dim3 grid(rslf_div_up(U, 128), V, S);
    if (C == 1)
        colour_maps_kernel<1><<<grid, 128, 0, ctx->stream>>>(ctx->out_map, ctx->out_valid, ctx->lv[0].epi, V, S, U, a, b,
                                                              params->cut_shadows, shadow, shadow_T, ctx->colour_lut, d_out);
    else
        colour_maps_kernel<3><<<grid, 128, 0, ctx->stream>>>(ctx->out_map, ctx->out_valid, ctx->lv[0].epi, V, S, U, a, b,
                                                              params->cut_shadows, shadow, shadow_T, ctx->colour_lut, d_out);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 3;
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(out_bgr_svu3, d_out, px * 3, cudaMemcpyDeviceToHost, ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RSLF_OK;
}

extern "C" int rslf_cuda_fine_to_coarse_level_dims(const rslf_ctx* ctx, int level, int* V, int* U)
{
    if (!ctx) return RSLF_ERR_ARG;
    if (ctx->last_kind != 3) return RSLF_ERR_STATE;
    if (level >= 0 && level < ctx->n_levels) {
        if (V) *V = ctx->lv[level].V;
        if (U) *U = ctx->lv[level].U;
    }
    return ctx->n_levels;
}

extern "C" int rslf_cuda_fine_to_coarse_get_level(rslf_ctx* ctx, int level, float* best_depth_svu,
                                                  float* edge_conf_svu, uint8_t* edge_mask_svu, float* disp_conf_svu,
                                                  float* dmin_svu, float* dmax_svu)
{
    if (!ctx) return RSLF_ERR_ARG;
    if (ctx->last_kind != 3 || level < 0 || level >= ctx->n_levels) return RSLF_ERR_STATE;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return get_level_maps(ctx, level, best_depth_svu, edge_conf_svu, edge_mask_svu, disp_conf_svu, nullptr, dmin_svu, dmax_svu);
}

/* ------------------------------------------------------------------ free functions */
extern "C" int rslf_cuda_edge_confidence(rslf_ctx* ctx, int s, const rslf_params* params, float* edge_conf_vu,
                                         uint8_t* edge_mask_vu)
{
    if (!ctx || !params) return RSLF_ERR_ARG;
    if (!ctx->have_input) return RSLF_ERR_STATE;
    if (s < 0 || s >= ctx->S) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_TRY(ensure_scratch(ctx, false, false));
    RSLF_TRY(ensure_level(ctx, 0, ctx->V, ctx->U, false, false));
    ctx->last_kind = 0;
    RSLF_TRY(normalise_level(ctx, 0, ctx->raw_in, ctx->cv_depth));
    rslf_level& L = ctx->lv[0];
    RSLF_TRY(launch_edge_confidence(ctx, L.epi, ctx->V, ctx->S, ctx->U, ctx->C, s, 1, *params, L.ce, L.emask, nullptr, L.remaining));
    const size_t plane = (size_t)ctx->V * ctx->U;
    RSLF_TRY(copy_out(ctx, edge_conf_vu, L.ce, plane * 4));
    RSLF_TRY(copy_out(ctx, edge_mask_vu, L.emask, plane));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RSLF_OK;
}

extern "C" int rslf_cuda_selective_median(rslf_ctx* ctx, const float* src_vu, const uint8_t* mask_vu, int s_hat,
                                          int size, float epsilon, float* dst_vu)
{
    if (!ctx || !src_vu || !mask_vu || !dst_vu) return RSLF_ERR_ARG;
    if (!ctx->have_input) return RSLF_ERR_STATE;
    if (s_hat < 0 || s_hat >= ctx->S) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_TRY(ensure_scratch(ctx, false, false));
    RSLF_TRY(ensure_level(ctx, 0, ctx->V, ctx->U, false, false));
    ctx->last_kind = 0;
    RSLF_TRY(normalise_level(ctx, 0, ctx->raw_in, ctx->cv_depth));
    rslf_level& L = ctx->lv[0];
    const size_t plane = (size_t)ctx->V * ctx->U;
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pile_depth_raw, src_vu, plane * 4, cudaMemcpyHostToDevice, ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(L.emask, mask_vu, plane, cudaMemcpyHostToDevice, ctx->stream));
    RSLF_TRY(launch_selective_median(ctx, ctx->pile_depth_raw, L.emask, L.epi + (size_t)s_hat * ctx->U * ctx->C,
                                     (size_t)ctx->S * ctx->U * ctx->C, ctx->V, ctx->U, ctx->C, size, epsilon, ctx->filtered));
    RSLF_TRY(copy_out(ctx, dst_vu, ctx->filtered, plane * 4));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RSLF_OK;
}

extern "C" int rslf_cuda_downsample_epis(rslf_ctx* ctx, const float* in_epis, int V, int S, int U, int C,
                                         float* out_epis, int* V2o, int* U2o)
{
    if (!ctx || !in_epis || !out_epis || (C != 1 && C != 3) || V < 1 || S < 1 || U < 1) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int V2 = cv_round_half(V), U2 = cv_round_half(U);
    if (V2o) *V2o = V2;
    if (U2o) *U2o = U2;
    if (V2 < 1 || U2 < 1) return RSLF_ERR_ARG;
    float *din = nullptr, *dout = nullptr;
    const size_t nin = (size_t)V * S * U * C, nout = (size_t)V2 * S * U2 * C;
    RSLF_TRY(dev_alloc(ctx, &din, nin));
    int rc = dev_alloc(ctx, &dout, nout);
    if (rc == RSLF_OK) {
        cudaMemcpyAsync(din, in_epis, nin * 4, cudaMemcpyHostToDevice, ctx->stream);
        rc = launch_downsample(ctx, din, V, S, U, C, dout, V2, U2);
        if (rc == RSLF_OK) {
            cudaMemcpyAsync(out_epis, dout, nout * 4, cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = RSLF_ERR_CUDA;
        }
    }
    dev_free(&din); dev_free(&dout);
    return rc;
}

template <typename T>
static int downsample_epis_int(rslf_ctx* ctx, const T* in_epis, int V, int S, int U, int C, T* out_epis, int* V2o, int* U2o)
{
    if (!ctx || !in_epis || !out_epis || (C != 1 && C != 3) || V < 1 || S < 1 || U < 1) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const int V2 = cv_round_half(V), U2 = cv_round_half(U);
    if (V2o) *V2o = V2;
    if (U2o) *U2o = U2;
    if (V2 < 1 || U2 < 1) return RSLF_ERR_ARG;
    T *din = nullptr, *dout = nullptr;
    const size_t nin = (size_t)V * S * U * C, nout = (size_t)V2 * S * U2 * C;
    RSLF_TRY(dev_alloc(ctx, &din, nin));
    int rc = dev_alloc(ctx, &dout, nout);
    if (rc == RSLF_OK) {
        cudaMemcpyAsync(din, in_epis, nin * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
        rc = launch_downsample_int<T>(ctx, din, V, S, U, C, dout, V2, U2);
        if (rc == RSLF_OK) {
            cudaMemcpyAsync(out_epis, dout, nout * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = RSLF_ERR_CUDA;
        }
    }
    dev_free(&din); dev_free(&dout);
    return rc;
}

extern "C" int rslf_cuda_downsample_epis_u8(rslf_ctx* ctx, const uint8_t* in_epis, int V, int S, int U, int C,
                                            uint8_t* out_epis, int* V2o, int* U2o)
{
    return downsample_epis_int<uint8_t>(ctx, in_epis, V, S, U, C, out_epis, V2o, U2o);
}

extern "C" int rslf_cuda_downsample_epis_u16(rslf_ctx* ctx, const uint16_t* in_epis, int V, int S, int U, int C,
                                             uint16_t* out_epis, int* V2o, int* U2o)
{
    return downsample_epis_int<uint16_t>(ctx, in_epis, V, S, U, C, out_epis, V2o, U2o);
}

extern "C" int rslf_cuda_set_bounds(rslf_ctx* ctx, const float* depth_up, const uint8_t* valid_up, int S, int Vu, int Uu,
                                    int Vd, int Ud, float* dmin_map, float* dmax_map)
{
    if (!ctx || !depth_up || !valid_up || !dmin_map || !dmax_map) return RSLF_ERR_ARG;
    if (S < 1 || Vu < 1 || Uu < 1 || Vd < 1 || Ud < 1) { snprintf(ctx->err, sizeof(ctx->err), "rslf_cuda_set_bounds: dimensions must be positive"); return RSLF_ERR_ARG; }
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const size_t nu = (size_t)S * Vu * Uu, nd = (size_t)S * Vd * Ud;
    float *d_dep = nullptr, *d_mn = nullptr, *d_mx = nullptr; uint8_t* d_val = nullptr; int *d_l = nullptr, *d_r = nullptr;
    int rc = dev_alloc(ctx, &d_dep, nu);
    if (rc == RSLF_OK) rc = dev_alloc(ctx, &d_val, nu);
    if (rc == RSLF_OK) rc = dev_alloc(ctx, &d_l, nu);
    if (rc == RSLF_OK) rc = dev_alloc(ctx, &d_r, nu);
    if (rc == RSLF_OK) rc = dev_alloc(ctx, &d_mn, nd);
    if (rc == RSLF_OK) rc = dev_alloc(ctx, &d_mx, nd);
    if (rc == RSLF_OK) {
        cudaMemcpyAsync(d_dep, depth_up, nu * 4, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_val, valid_up, nu, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_mn, dmin_map, nd * 4, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(d_mx, dmax_map, nd * 4, cudaMemcpyHostToDevice, ctx->stream);
        const int rows = S * Vu;
        nearest_valid_kernel<<<rslf_div_up((long long)rows * 32, 256), 256, 0, ctx->stream>>>(d_val, rows, Uu, d_l, d_r);
        dim3 grid(rslf_div_up(Ud, 128), Vd, S);
        set_bounds_kernel<<<grid, 128, 0, ctx->stream>>>(d_dep, d_l, d_r, S, Vu, Uu, Vd, Ud, d_mn, d_mx, Vu, 0, 0, 0, 0.f, 0.f);
        cudaMemcpyAsync(dmin_map, d_mn, nd * 4, cudaMemcpyDeviceToHost, ctx->stream);
        cudaMemcpyAsync(dmax_map, d_mx, nd * 4, cudaMemcpyDeviceToHost, ctx->stream);
        if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) rc = RSLF_ERR_CUDA;
    }
    dev_free(&d_dep); dev_free(&d_val); dev_free(&d_l); dev_free(&d_r); dev_free(&d_mn); dev_free(&d_mx);
    return rc;
}

extern "C" int rslf_cuda_fuse_disp_maps(rslf_ctx* ctx, int levels, int S, const int* Vp, const int* Up,
                                        const float* const* disp_p, const uint8_t* const* valid_p,
                                        float* out_map_svu, uint8_t* out_valid_svu)
{
    if (!ctx || levels < 1 || levels > RSLF_MAX_LEVELS || !Vp || !Up || !disp_p || !valid_p || S < 1) return RSLF_ERR_ARG;
    for (int p = 0; p < levels; ++p)
        if (Vp[p] < 1 || Up[p] < 1 || !disp_p[p] || !valid_p[p]) { snprintf(ctx->err, sizeof(ctx->err), "rslf_cuda_fuse_disp_maps: level %d is empty", p); return RSLF_ERR_ARG; }
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<float*> dd(levels, nullptr); std::vector<uint8_t*> dv(levels, nullptr);
    const size_t px0 = (size_t)S * Vp[0] * Up[0];
    float *fa = nullptr, *fb = nullptr, *om = nullptr; uint8_t *ma = nullptr, *mb = nullptr, *ov = nullptr;
    int rc = RSLF_OK;
    for (int p = 0; p < levels && rc == RSLF_OK; ++p) {
        const size_t px = (size_t)S * Vp[p] * Up[p];
        rc = dev_alloc(ctx, &dd[p], px);
        if (rc == RSLF_OK) rc = dev_alloc(ctx, &dv[p], px);
        if (rc == RSLF_OK) {
            cudaMemcpyAsync(dd[p], disp_p[p], px * 4, cudaMemcpyHostToDevice, ctx->stream);
            cudaMemcpyAsync(dv[p], valid_p[p], px, cudaMemcpyHostToDevice, ctx->stream);
        }
    }
    if (rc == RSLF_OK) rc = dev_alloc(ctx, &fa, px0);
    if (rc == RSLF_OK) rc = dev_alloc(ctx, &fb, px0);
    if (rc == RSLF_OK) rc = dev_alloc(ctx, &ma, px0);
    if (rc == RSLF_OK) rc = dev_alloc(ctx, &mb, px0);
    if (rc == RSLF_OK) rc = dev_alloc(ctx, &om, px0);
    if (rc == RSLF_OK) rc = dev_alloc(ctx, &ov, px0);
    if (rc == RSLF_OK) {
        /* temporarily borrow the ctx fuse buffers' slots */
        float* sa = ctx->fuse_a; float* sb = ctx->fuse_b; uint8_t* sma = ctx->fuse_ma; uint8_t* smb = ctx->fuse_mb; int sS = ctx->S;
        ctx->fuse_a = fa; ctx->fuse_b = fb; ctx->fuse_ma = ma; ctx->fuse_mb = mb; ctx->S = S;
        {
            /* host maps are whole: a single-rank table, whatever the ctx's communicator */
            std::vector<shard_tab> tabs(levels);
            for (int p = 0; p < levels; ++p) tabs[p] = single_tab(Vp[p]);
            int sw = ctx->world, sr = ctx->rank; ctx->world = 1; ctx->rank = 0;
            bool norep[RSLF_MAX_LEVELS] = {false};
            rc = launch_fuse(ctx, levels, Vp, Up, tabs.data(), norep, dd.data(), dv.data(), om, ov);
            ctx->world = sw; ctx->rank = sr;
        }
        ctx->fuse_a = sa; ctx->fuse_b = sb; ctx->fuse_ma = sma; ctx->fuse_mb = smb; ctx->S = sS;
        if (rc == RSLF_OK) {
            if (out_map_svu) cudaMemcpyAsync(out_map_svu, om, px0 * 4, cudaMemcpyDeviceToHost, ctx->stream);
            if (out_valid_svu) cudaMemcpyAsync(out_valid_svu, ov, px0, cudaMemcpyDeviceToHost, ctx->stream);
            if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = RSLF_ERR_CUDA;
        }
    }
    for (int p = 0; p < levels; ++p) { dev_free(&dd[p]); dev_free(&dv[p]); }
    dev_free(&fa); dev_free(&fb); dev_free(&ma); dev_free(&mb); dev_free(&om); dev_free(&ov);
    return rc;
}

/* ------------------------------------------------------------------ measurement */
extern "C" int rslf_plan_depth_tm(int S, int C, int D, int s_hat, float dmin, float dmax, float slope, size_t smem_limit,
                                  int* out8, int* rounds, int* blocks)
{
    if (!out8 || S < 1 || (C != 1 && C != 3) || D < 2) return RSLF_ERR_ARG;
    depth_tm_layout L; memset(&L, 0, sizeof(L));
    const int wpv = depth_wpv_q16(D, 1, dmin, dmax, slope);
    const bool ok = depth_tm_build(L, S, C, s_hat, wpv, smem_limit);
    out8[0] = ok ? 1 : 0; out8[1] = DEPTH_TM_RV; out8[2] = L.TV; out8[3] = L.SV; out8[4] = L.reg_last;
    out8[5] = L.nrounds; out8[6] = L.warp_bytes; out8[7] = DEPTH_TM_WARPS;
    if (!ok) return RSLF_OK;
    if (rounds)
        for (int r = 0; r < L.nrounds; ++r) { rounds[3 * r] = L.round_b0[r]; rounds[3 * r + 1] = L.round_nb[r]; rounds[3 * r + 2] = L.round_kind[r]; }
    if (blocks)
        for (int b = 0; b < depth_padded_views(S) / DEPTH_UNR; ++b) { blocks[2 * b] = 4 * (int)L.blk_off4[b]; blocks[2 * b + 1] = L.blk_pitch[b]; }
    return RSLF_OK;
}

/* Host-only: the share of a pass that rank `rank` of `world` evaluates when the ranks' work lists hold counts[0..world)
 * pixels (pass-balanced multi-GPU mode, k_balance.cuh): first pixel and number of pixels in the concatenated lists,
 * and for every owner how many of those pixels come from its list (from_owner[world], optional).  Same arithmetic as
 * the device side (bal_share). */
extern "C" int rslf_balance_share(const int* counts, int world, int rank, long long* first, int* count, int* from_owner)
{
    if (!counts || world < 1 || world > RSLF_MAX_PEERS || rank < 0 || rank >= world || !first || !count) return RSLF_ERR_ARG;
    long long T = 0;
    for (int r = 0; r < world; ++r) { if (counts[r] < 0) return RSLF_ERR_ARG; T += counts[r]; }
    int lo = 0, n = 0;
    bal_share(T, rank, world, &lo, &n);
    *first = lo; *count = n;
    if (from_owner) {
        long long pre = 0;
        for (int r = 0; r < world; ++r) {
            const long long a = std::max<long long>(pre, lo), b = std::min<long long>(pre + counts[r], (long long)lo + n);
            from_owner[r] = (int)std::max<long long>(0, b - a);
            pre += counts[r];
        }
    }
    return RSLF_OK;
}

extern "C" int rslf_cuda_measure_fp32_peak(rslf_ctx* ctx, double* gops_nofma, double* gflops_fma)
{
    if (!ctx) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return measure_fp32_peak(ctx, gops_nofma, gflops_fma);
}

/* Gop/s of separately rounded FP32 multiplies / adds issued as packed f32x2 instructions (FFMA2 / FADD2). */
extern "C" int rslf_cuda_measure_fp32x2_peak(rslf_ctx* ctx, double* gops_nofma_packed)
{
    if (!ctx) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    return measure_fp32x2_peak(ctx, gops_nofma_packed);
}

/* Pixels evaluated per image row of this rank's block in the last run (all passes; pyramid levels mapped back to
 * level-0 rows; replicated coarse levels excluded).  The cost of a pixel is D x S samples whatever its level, so
 * this is the work profile a caller needs to balance the row blocks of a multi-GPU run. */
extern "C" int rslf_cuda_get_row_work(rslf_ctx* ctx, unsigned* rows_out)
{
    if (!ctx || !rows_out) return RSLF_ERR_ARG;
    if (!ctx->rowwork || ctx->last_kind == 0) return RSLF_ERR_STATE;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(rows_out, ctx->rowwork, (size_t)ctx->V * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return RSLF_OK;
}
