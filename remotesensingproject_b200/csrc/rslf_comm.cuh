/*
 * rslf_comm.cuh — NCCL plumbing for the row-sharded run (one process per GPU).
 * NCCL is resolved with dlopen at communicator creation, so the library loads
 * (and single-GPU runs work) on hosts without libnccl.  NCCL carries the one-off
 * all-gather of the raw stack (every rank keeps the whole EPI stack, the result
 * maps are sharded by rows), the gathers around the first replicated pyramid level
 * and the fallbacks; the per-pass traffic — work / result records of the balanced
 * depth kernel, the median halo rows (rslf_depth_computation_core.hpp:698-709) —
 * goes through peer memory over NVLink (second half of this file, k_balance.cuh).
 */
#pragma once
#include <dlfcn.h>
#include <algorithm>
#include <vector>
#include "rslf_common.cuh"

typedef struct { char internal[128]; } rslf_nccl_uid;
typedef int (*fn_ncclGetUniqueId)(rslf_nccl_uid*);
typedef int (*fn_ncclCommInitRank)(void**, int, rslf_nccl_uid, int);
typedef int (*fn_ncclCommDestroy)(void*);
typedef int (*fn_ncclAllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_ncclAllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_ncclGetErrorString)(int);

struct rslf_nccl_api {
    void* lib = nullptr;
    fn_ncclGetUniqueId GetUniqueId = nullptr;
    fn_ncclCommInitRank CommInitRank = nullptr;
    fn_ncclCommDestroy CommDestroy = nullptr;
    fn_ncclAllGather AllGather = nullptr;
    fn_ncclAllReduce AllReduce = nullptr;
    fn_ncclGetErrorString GetErrorString = nullptr;
};
static rslf_nccl_api g_nccl;

/* ncclDataType_t / ncclRedOp_t values (stable across NCCL 2.x) */
enum { RSLF_NCCL_UINT8 = 1, RSLF_NCCL_UINT32 = 3, RSLF_NCCL_FLOAT = 7, RSLF_NCCL_SUM = 0, RSLF_NCCL_MAX = 2 };

static int nccl_load(char* err, size_t errlen)
{
    if (g_nccl.lib) return RSLF_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !g_nccl.lib; ++i) g_nccl.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!g_nccl.lib) { snprintf(err, errlen, "dlopen(libnccl.so.2): %s", dlerror()); return RSLF_ERR_NCCL; }
    g_nccl.GetUniqueId = (fn_ncclGetUniqueId)dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (fn_ncclCommInitRank)dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (fn_ncclCommDestroy)dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.AllGather = (fn_ncclAllGather)dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.AllReduce = (fn_ncclAllReduce)dlsym(g_nccl.lib, "ncclAllReduce");
    g_nccl.GetErrorString = (fn_ncclGetErrorString)dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.AllReduce) {
        snprintf(err, errlen, "libnccl lacks a required symbol");
        return RSLF_ERR_NCCL;
    }
    return RSLF_OK;
}

#define RSLF_NCCL_TRY(ctx, call)                                                               \
    do {                                                                                       \
        int _r = (call);                                                                       \
        if (_r != 0) {                                                                         \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #call, \
                     g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error");         \
            return RSLF_ERR_NCCL;                                                              \
        }                                                                                      \
    } while (0)

static void arena_close(rslf_ctx* ctx);
static void rslf_disable_ring();

static void comm_destroy(rslf_ctx* ctx)
{
    arena_close(ctx);
    if (ctx->p2p_done) cudaFree(ctx->p2p_done);
    ctx->p2p_done = nullptr; ctx->p2p_state = 0;
    if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
}

/* in-place max over ranks of n device floats; optionally read element 0 back */
static int comm_allreduce_max(rslf_ctx* ctx, float* dev, int n, float* host0)
{
    if (ctx->world <= 1 || !ctx->nccl_comm) return RSLF_OK;
    RSLF_NCCL_TRY(ctx, g_nccl.AllReduce(dev, dev, (size_t)n, RSLF_NCCL_FLOAT, RSLF_NCCL_MAX, ctx->nccl_comm, ctx->stream));
    if (host0) {
        RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(host0, dev, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        RSLF_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return RSLF_OK;
}

/* in-place sum over ranks of n device uint32 counters (histograms of the coloured maps' quantile fit) */
static int comm_allreduce_sum_u32(rslf_ctx* ctx, unsigned* dev, size_t n)
{
    if (ctx->world <= 1 || !ctx->nccl_comm) return RSLF_OK;
    RSLF_NCCL_TRY(ctx, g_nccl.AllReduce(dev, dev, n, RSLF_NCCL_UINT32, RSLF_NCCL_SUM, ctx->nccl_comm, ctx->stream));
    return RSLF_OK;
}

static int comm_allreduce_max_host(rslf_ctx* ctx, float v, float* out)
{
    *out = v;
    if (ctx->world <= 1 || !ctx->nccl_comm) return RSLF_OK;
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->minmax + 2, &v, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    return comm_allreduce_max(ctx, ctx->minmax + 2, 1, out);
}

/* every rank contributes `bytes` from send; recv holds world * bytes in rank order */
static int comm_allgather_bytes(rslf_ctx* ctx, const void* send, void* recv, size_t bytes)
{
    if (ctx->world <= 1 || !ctx->nccl_comm) return RSLF_OK;
    RSLF_NCCL_TRY(ctx, g_nccl.AllGather(send, recv, bytes, RSLF_NCCL_UINT8, ctx->nccl_comm, ctx->stream));
    return RSLF_OK;
}

/* ---- row shards -------------------------------------------------------------
 * Rank r holds the global rows [b[r], b[r+1]) of a level.  Level-0 boundaries are multiples of
 * 2^(levels-1), so level p splits at b0[r] >> p and bound propagation (rows 2v, 2v+1 -> v) is local.
 */
struct shard_tab { int n; int b[65]; };

static shard_tab level_shards(const rslf_ctx* ctx, int p, int Vtot_p)
{
    shard_tab t; t.n = ctx->world;
    for (int r = 0; r <= ctx->world; ++r) {
        int v = ctx->row_starts[r] >> p;
        t.b[r] = v < Vtot_p ? v : Vtot_p;
    }
    t.b[0] = 0; t.b[ctx->world] = Vtot_p;
    return t;
}
static int shard_max_rows(const shard_tab& t) { int m = 0; for (int r = 0; r < t.n; ++r) m = std::max(m, t.b[r + 1] - t.b[r]); return m; }
static bool shard_equal(const shard_tab& t) { for (int r = 0; r < t.n; ++r) if (t.b[r + 1] - t.b[r] != t.b[1] - t.b[0]) return false; return true; }

static int comm_ensure_stage(rslf_ctx* ctx, size_t blockbytes)
{
    if (ctx->g_stage_cap >= blockbytes) return RSLF_OK;
    if (ctx->g_send) cudaFree(ctx->g_send);
    if (ctx->g_recv) cudaFree(ctx->g_recv);
    ctx->g_send = ctx->g_recv = nullptr; ctx->g_stage_cap = 0;
    if (cudaMalloc(&ctx->g_send, blockbytes) != cudaSuccess || cudaMalloc(&ctx->g_recv, blockbytes * ctx->world) != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(gather staging %zu x %d)", blockbytes, ctx->world);
        return RSLF_ERR_NOMEM;
    }
    ctx->g_stage_cap = blockbytes;
    return RSLF_OK;
}

/* all ranks' rows of a row-major [V][row_bytes] array -> global [Vtot][row_bytes] on every rank */
static int comm_gather_rows(rslf_ctx* ctx, const void* local, size_t row_bytes, const shard_tab& t, void* global)
{
    const int r0 = ctx->rank;
    const size_t mine = (size_t)(t.b[r0 + 1] - t.b[r0]) * row_bytes;
    if (shard_equal(t)) {
        RSLF_NCCL_TRY(ctx, g_nccl.AllGather(local, global, mine, RSLF_NCCL_UINT8, ctx->nccl_comm, ctx->stream));
        return RSLF_OK;
    }
    const size_t blockbytes = (size_t)shard_max_rows(t) * row_bytes;
    RSLF_TRY(comm_ensure_stage(ctx, blockbytes));
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(ctx->g_send, local, mine, cudaMemcpyDeviceToDevice, ctx->stream));
    RSLF_NCCL_TRY(ctx, g_nccl.AllGather(ctx->g_send, ctx->g_recv, blockbytes, RSLF_NCCL_UINT8, ctx->nccl_comm, ctx->stream));
    for (int r = 0; r < t.n; ++r) {
        const size_t n = (size_t)(t.b[r + 1] - t.b[r]) * row_bytes;
        if (n) RSLF_CUDA_TRY(ctx, cudaMemcpyAsync((char*)global + (size_t)t.b[r] * row_bytes, (char*)ctx->g_recv + r * blockbytes, n,
                                                  cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return RSLF_OK;
}

/* all ranks' rows of a [S][Vloc][U] map -> global [S][Vtot][U] on every rank */
static int comm_gather_svu(rslf_ctx* ctx, const void* local, size_t elem, int S, int U, const shard_tab& t, void* global)
{
    const int r0 = ctx->rank, Vtot = t.b[t.n];
    const size_t maxrows = shard_max_rows(t), blockbytes = (size_t)S * maxrows * U * elem;
    RSLF_TRY(comm_ensure_stage(ctx, blockbytes));
    const size_t wloc = (size_t)(t.b[r0 + 1] - t.b[r0]) * U * elem;
    if (wloc) RSLF_CUDA_TRY(ctx, cudaMemcpy2DAsync(ctx->g_send, maxrows * U * elem, local, wloc, wloc, S, cudaMemcpyDeviceToDevice, ctx->stream));
    RSLF_NCCL_TRY(ctx, g_nccl.AllGather(ctx->g_send, ctx->g_recv, blockbytes, RSLF_NCCL_UINT8, ctx->nccl_comm, ctx->stream));
    for (int r = 0; r < t.n; ++r) {
        const size_t wr = (size_t)(t.b[r + 1] - t.b[r]) * U * elem;
        if (wr) RSLF_CUDA_TRY(ctx, cudaMemcpy2DAsync((char*)global + (size_t)t.b[r] * U * elem, (size_t)Vtot * U * elem,
                                                     (char*)ctx->g_recv + r * blockbytes, maxrows * U * elem, wr, S,
                                                     cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return RSLF_OK;
}

/* per-pass exchange for the selective median: depth, colour and mask rows of line s_hat of every rank */
__global__ void unpack_median_kernel(const char* __restrict__ recv, size_t blockbytes, size_t off_colour, size_t off_mask,
                                     shard_tab t, int U, int C, float* __restrict__ g_depth, float* __restrict__ g_colour,
                                     uint8_t* __restrict__ g_mask)
{
    const int k = blockIdx.y;                   /* global row */
    int r = 0;
    while (r + 1 < t.n && k >= t.b[r + 1]) ++r;
    const int lv = k - t.b[r];
    const char* blk = recv + (size_t)r * blockbytes;
    const float* d = reinterpret_cast<const float*>(blk) + (size_t)lv * U;
    const float* c = reinterpret_cast<const float*>(blk + off_colour) + (size_t)lv * U * C;
    const uint8_t* m = reinterpret_cast<const uint8_t*>(blk + off_mask) + (size_t)lv * U;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < U; u += gridDim.x * blockDim.x) {
        g_depth[(size_t)k * U + u] = d[u];
        g_mask[(size_t)k * U + u] = m[u];
        for (int cc = 0; cc < C; ++cc) g_colour[((size_t)k * U + u) * C + cc] = c[(size_t)u * C + cc];
    }
}

static int comm_gather_median_planes(rslf_ctx* ctx, const float* depth_plane, const uint8_t* mask_plane, const float* colour0,
                                     size_t colour_row_stride_floats, int U, int C, const shard_tab& t)
{
    const int r0 = ctx->rank, Vloc = t.b[r0 + 1] - t.b[r0], Vtot = t.b[t.n];
    const size_t maxrows = shard_max_rows(t);
    const size_t off_colour = maxrows * U * 4, off_mask = off_colour + maxrows * U * C * 4;
    const size_t blockbytes = (off_mask + maxrows * U + 15) & ~(size_t)15;
    RSLF_TRY(comm_ensure_stage(ctx, blockbytes));
    char* send = (char*)ctx->g_send;
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(send, depth_plane, (size_t)Vloc * U * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaMemcpy2DAsync(send + off_colour, (size_t)U * C * 4, colour0, colour_row_stride_floats * 4, (size_t)U * C * 4, Vloc,
                                         cudaMemcpyDeviceToDevice, ctx->stream));
    RSLF_CUDA_TRY(ctx, cudaMemcpyAsync(send + off_mask, mask_plane, (size_t)Vloc * U, cudaMemcpyDeviceToDevice, ctx->stream));
    RSLF_NCCL_TRY(ctx, g_nccl.AllGather(ctx->g_send, ctx->g_recv, blockbytes, RSLF_NCCL_UINT8, ctx->nccl_comm, ctx->stream));
    dim3 grid(std::max(1, std::min(8, (U + 127) / 128)), Vtot);
    unpack_median_kernel<<<grid, 128, 0, ctx->stream>>>((const char*)ctx->g_recv, blockbytes, off_colour, off_mask, t, U, C,
                                                        ctx->g_depth, ctx->g_colour, ctx->g_mask);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    return RSLF_OK;
}

/* per-pass halo exchange for the selective median when every rank holds >= 2 rows: each rank contributes
 * its first two and last two rows of the depth / colour / mask planes of line s_hat (4 rows), one small
 * all-gather, and the median kernel reads its neighbours' rows straight out of the receive buffer */
__global__ void pack_median_halo_kernel(const float* __restrict__ depth, const uint8_t* __restrict__ mask,
                                        const float* __restrict__ colour0, size_t colour_row_stride, int Vloc, int U, int C,
                                        char* __restrict__ send, size_t off_colour, size_t off_mask)
{
    const int j = blockIdx.y;                                  /* 0,1: first rows; 2,3: last rows */
    const int v = (j < 2) ? j : Vloc - 4 + j;
    float* d = reinterpret_cast<float*>(send) + (size_t)j * U;
    float* c = reinterpret_cast<float*>(send + off_colour) + (size_t)j * U * C;
    uint8_t* m = reinterpret_cast<uint8_t*>(send + off_mask) + (size_t)j * U;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < U; u += gridDim.x * blockDim.x) {
        d[u] = depth[(size_t)v * U + u];
        m[u] = mask[(size_t)v * U + u];
        for (int cc = 0; cc < C; ++cc) c[(size_t)u * C + cc] = colour0[(size_t)v * colour_row_stride + (size_t)u * C + cc];
    }
}

static int comm_exchange_median_halo(rslf_ctx* ctx, const float* depth_plane, const uint8_t* mask_plane, const float* colour0,
                                     size_t colour_row_stride_floats, int U, int C, const shard_tab& t, median_halo* out)
{
    const int r0 = ctx->rank, Vloc = t.b[r0 + 1] - t.b[r0];
    const size_t off_colour = (size_t)4 * U * 4, off_mask = off_colour + (size_t)4 * U * C * 4;
    const size_t blockbytes = (off_mask + (size_t)4 * U + 15) & ~(size_t)15;
    RSLF_TRY(comm_ensure_stage(ctx, blockbytes));
    dim3 grid(std::max(1, std::min(8, (U + 127) / 128)), 4);
    pack_median_halo_kernel<<<grid, 128, 0, ctx->stream>>>(depth_plane, mask_plane, colour0, colour_row_stride_floats, Vloc, U, C,
                                                           (char*)ctx->g_send, off_colour, off_mask);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    RSLF_NCCL_TRY(ctx, g_nccl.AllGather(ctx->g_send, ctx->g_recv, blockbytes, RSLF_NCCL_UINT8, ctx->nccl_comm, ctx->stream));
    memset(out, 0, sizeof(*out));
    out->colour_halo_stride = (size_t)U * C;
    const char* recv = (const char*)ctx->g_recv;
    if (r0 > 0) {                                              /* rows 2, 3 of the rank above = its last two rows */
        const char* b = recv + (size_t)(r0 - 1) * blockbytes;
        out->top_depth = reinterpret_cast<const float*>(b) + (size_t)2 * U;
        out->top_colour = reinterpret_cast<const float*>(b + off_colour) + (size_t)2 * U * C;
        out->top_mask = reinterpret_cast<const uint8_t*>(b + off_mask) + (size_t)2 * U;
    }
    if (r0 + 1 < t.n) {                                        /* rows 0, 1 of the rank below = its first two rows */
        const char* b = recv + (size_t)(r0 + 1) * blockbytes;
        out->bot_depth = reinterpret_cast<const float*>(b);
        out->bot_colour = reinterpret_cast<const float*>(b + off_colour);
        out->bot_mask = reinterpret_cast<const uint8_t*>(b + off_mask);
    }
    return RSLF_OK;
}

/* ---- peer memory over NVLink (CUDA IPC) -------------------------------------------------------------------
 * Every rank allocates one ARENA with the same layout, exports it with cudaIpcGetMemHandle and maps the
 * arenas of all peers.  It holds everything another GPU reads or writes during a run:
 *
 *   [halo areas: 2 slots x {from above, from below}]   rows next to the block for the 5x5 selective median
 *   [halo flags]  [counts[r]]  [done[r]]                 sequence words raised by the peers (k_balance.cuh)
 *   [records: cap x 16 B]  [results: cap x 32 B]         work / result records of the pass-balanced depth kernel
 *
 * Halo exchange: instead of a collective, every rank stores its two first / two last rows of line s_hat straight
 * into its neighbours' areas with ordinary global stores over NVLink and raises a sequence flag there; the
 * neighbour's median kernel waits for the flag.  No rank ever waits inside a producer, so there is no circular
 * wait; two slots alternate because a rank can be at most one pass ahead of its neighbour.
 */
static size_t p2p_area_bytes(int U, int C) { return (((size_t)2 * U * 4 + (size_t)2 * U * C * 4 + (size_t)2 * U) + 255) & ~(size_t)255; }

#define FUSE_HALO 2               /* rows of a fused map a neighbour may need (bilinear upsampling, 3x3 median) */
#define FUSE_SLOTS 8              /* sharded pyramid levels whose fused maps travel through the arena */

struct arena_layout { size_t area, off_flags, off_counts, off_done, off_rec, off_res, off_fuse, fuse_area, total; };
static arena_layout arena_make_layout(int U, int C, size_t cap_rec, int S)
{
    arena_layout l;
    l.area = p2p_area_bytes(U, C);
    l.off_flags = 4 * l.area;                     /* 4 median-halo flags, then at + 64: FUSE_SLOTS x 2 fuse-halo flags */
    l.off_counts = l.off_flags + 256;
    l.off_done = l.off_counts + 8 * RSLF_MAX_PEERS;
    l.off_rec = (l.off_done + 8 * RSLF_MAX_PEERS + 255) & ~(size_t)255;
    l.off_res = l.off_rec + ((cap_rec * 16 + 255) & ~(size_t)255);
    l.off_fuse = (l.off_res + cap_rec * 32 + 255) & ~(size_t)255;
    l.fuse_area = ((size_t)S * FUSE_HALO * U * 5 + 255) & ~(size_t)255;     /* [S][H][U] float map, then [S][H][U] mask */
    l.total = l.off_fuse + (size_t)FUSE_SLOTS * 2 * l.fuse_area + 256;
    return l;
}

/* Unmaps the peers' arenas and frees this rank's.  An exporter must not free an allocation that an importer still has
 * mapped (the importer's cudaIpcCloseMemHandle then faults), so between the two steps the ranks meet in a tiny
 * all-reduce: every rank has unmapped before any rank frees.  Collective whenever an arena is live (all ranks tear
 * their contexts down, or rebuild the arena, at the same point of the program). */
static int comm_allreduce_max_host(rslf_ctx* ctx, float v, float* out);
static void arena_close(rslf_ctx* ctx)
{
    bool mapped = false;
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < RSLF_MAX_RANKS; ++r) {
        if (ctx->arena_peer[r] && ctx->arena_peer[r] != ctx->arena) { cudaIpcCloseMemHandle(ctx->arena_peer[r]); mapped = true; }
        ctx->arena_peer[r] = nullptr;
    }
    if (mapped && ctx->arena && ctx->nccl_comm && ctx->world > 1) { float out = 0.f; comm_allreduce_max_host(ctx, 0.f, &out); }
    cudaGetLastError();
    if (ctx->arena) cudaFree(ctx->arena);
    ctx->arena = nullptr; ctx->arena_bytes = 0; ctx->arena_cap_rec = 0;
    ctx->p2p_up = ctx->p2p_dn = nullptr;
}

/* Collective over the ranks: (re)creates the arenas for rows of U x C values and cap_rec records per rank.
 * Returns RSLF_ERR_UNSUPPORTED on EVERY rank when any rank cannot provide or map peer memory (the caller then
 * uses the NCCL exchanges).  Local failures never leave the other ranks alone in a collective. */
static int comm_arena_setup(rslf_ctx* ctx, int U, int C, size_t cap_rec)
{
    const arena_layout want = arena_make_layout(U, C, cap_rec, ctx->S);
    if (ctx->p2p_state > 0 && ctx->arena_U == U && ctx->arena_C == C && ctx->arena_S == ctx->S && ctx->arena_cap_rec >= cap_rec) return RSLF_OK;
    if (ctx->p2p_state > 0) ctx->p2p_state = 0;             /* other geometry: rebuild (collective: every rank sees the same change) */
    if (ctx->p2p_state < 0) return RSLF_ERR_UNSUPPORTED;
    const char* mode = getenv("RSLF_HALO");
    if (mode && mode[0] == 'n') { ctx->p2p_state = -1; return RSLF_ERR_UNSUPPORTED; }
    /* every rank that gets here takes part in the two collectives below, whatever fails locally */
    float ok = (ctx->world <= RSLF_MAX_RANKS) ? 1.f : 0.f;
    cudaStreamSynchronize(ctx->stream);
    arena_close(ctx);
    ctx->p2p_state = -1;
    const arena_layout l = want;
    if (ok > 0.f && cudaMalloc((void**)&ctx->arena, l.total) != cudaSuccess) { ctx->arena = nullptr; ok = 0.f; }
    if (!ctx->p2p_done && cudaMalloc((void**)&ctx->p2p_done, 8 * sizeof(int)) != cudaSuccess) { ctx->p2p_done = nullptr; ok = 0.f; }
    cudaIpcMemHandle_t mine; memset(&mine, 0, sizeof(mine));
    if (ok > 0.f) {
        cudaMemsetAsync(ctx->arena, 0, l.total, ctx->stream);
        cudaMemsetAsync(ctx->p2p_done, 0, 8 * sizeof(int), ctx->stream);
        if (cudaIpcGetMemHandle(&mine, ctx->arena) != cudaSuccess) ok = 0.f;
    }
    cudaGetLastError();
    /* exchange the handles with one all-gather (through device staging) */
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    std::vector<cudaIpcMemHandle_t> all(ctx->world);
    bool coll_ok = true;
    if (comm_ensure_stage(ctx, 256) != RSLF_OK) coll_ok = false;
    if (coll_ok) {
        coll_ok = cudaMemcpyAsync(ctx->g_send, &mine, hs, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
                  g_nccl.AllGather(ctx->g_send, ctx->g_recv, hs, RSLF_NCCL_UINT8, ctx->nccl_comm, ctx->stream) == 0 &&
                  cudaMemcpyAsync(all.data(), ctx->g_recv, hs * ctx->world, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
                  cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    }
    if (!coll_ok) { snprintf(ctx->err, sizeof(ctx->err), "peer-memory setup: handle exchange failed"); cudaGetLastError(); return RSLF_ERR_NCCL; }
    if (ok > 0.f) {
        for (int r = 0; r < ctx->world && ok > 0.f; ++r) {
            if (r == ctx->rank) { ctx->arena_peer[r] = ctx->arena; continue; }
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0.f; break; }
            ctx->arena_peer[r] = (char*)ptr;
        }
    }
    cudaGetLastError();
    float bad = (ok > 0.f) ? 0.f : 1.f, anybad = bad;
    RSLF_TRY(comm_allreduce_max_host(ctx, bad, &anybad));
    if (anybad > 0.f) {
        arena_close(ctx);
        return RSLF_ERR_UNSUPPORTED;                        /* every rank falls back to the NCCL exchanges */
    }
    ctx->p2p_up = ctx->rank > 0 ? ctx->arena_peer[ctx->rank - 1] : nullptr;
    ctx->p2p_dn = ctx->rank + 1 < ctx->world ? ctx->arena_peer[ctx->rank + 1] : nullptr;
    ctx->p2p_area = l.area; ctx->arena_bytes = l.total; ctx->arena_cap_rec = cap_rec; ctx->arena_U = U; ctx->arena_C = C; ctx->arena_S = ctx->S;
    ctx->p2p_seq = 0; ctx->bal_seq = 0; ctx->fuse_seq = 0; ctx->p2p_state = 1;
    return RSLF_OK;
}
static arena_layout arena_current(const rslf_ctx* ctx) { return arena_make_layout(ctx->arena_U, ctx->arena_C, ctx->arena_cap_rec, ctx->arena_S); }

struct p2p_push_args {
    const float* depth; const uint8_t* mask; const float* colour0; size_t colour_row_stride; int Vloc, U, C;
    char* up_area; char* dn_area;                   /* where my first rows go (neighbour above: its "from below" area) / my last rows */
    unsigned* up_flag; unsigned* dn_flag; unsigned seq; int* done; int nblocks;
};

/* rows j = 0, 1 (first rows -> rank above) and 2, 3 (last rows -> rank below), columns u0, u0 + ustep, ...;
 * colour0 == nullptr: the receiver reads the colours from its own copy of the stack */
__device__ __forceinline__ void p2p_push_rows(const p2p_push_args& a, int j, int u0, int ustep)
{
    char* dst = (j < 2) ? a.up_area : a.dn_area;
    if (!dst) return;
    const int v = (j < 2) ? j : a.Vloc - 4 + j, jj = j & 1;
    float* d = reinterpret_cast<float*>(dst) + (size_t)jj * a.U;
    float* c = reinterpret_cast<float*>(dst + (size_t)2 * a.U * 4) + (size_t)jj * a.U * a.C;
    uint8_t* m = reinterpret_cast<uint8_t*>(dst + (size_t)2 * a.U * 4 + (size_t)2 * a.U * a.C * 4) + (size_t)jj * a.U;
    for (int u = u0; u < a.U; u += ustep) {
        d[u] = a.depth[(size_t)v * a.U + u];
        m[u] = a.mask[(size_t)v * a.U + u];
        if (a.colour0)
            for (int cc = 0; cc < a.C; ++cc) c[(size_t)u * a.C + cc] = a.colour0[(size_t)v * a.colour_row_stride + (size_t)u * a.C + cc];
    }
}

__global__ void p2p_push_halo_kernel(const p2p_push_args a)
{
    p2p_push_rows(a, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
    /* the last block to finish raises the neighbours' flags, after every store of the grid is visible system-wide */
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int old = atomicAdd(a.done, 1);
        if (old == a.nblocks - 1) {
            *a.done = 0;
            __threadfence_system();
            if (a.up_flag) *reinterpret_cast<volatile unsigned*>(a.up_flag) = a.seq;
            if (a.dn_flag) *reinterpret_cast<volatile unsigned*>(a.dn_flag) = a.seq;
        }
    }
}

/* Balanced mode, owner side: waits for every rank's done flag of the pass, scatters the result records of this
 * rank's pixels into its planes of line s_hat (core.hpp:630-657), and lets the last block push the halo rows. */
struct bal_apply_args {
    const int* items; const int* count; const float4* res; int C;
    float* depth; float* cd; float* rbar; float* ce; uint8_t* emask;
    const volatile unsigned long long* done; int n; unsigned seq; int* err; int* blocks_done;
    int do_push; p2p_push_args push;
};

__global__ void __launch_bounds__(256) bal_apply_kernel(const bal_apply_args a)
{
    __shared__ int s_last;
    if ((int)threadIdx.x < a.n) peer_wait64(a.done + threadIdx.x, a.seq, a.err);
    __threadfence_system();
    __syncthreads();
    const int n = *a.count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 r0 = ld_cv_float4(a.res + 2 * (size_t)i), r1 = ld_cv_float4(a.res + 2 * (size_t)i + 1);
        const int pix = a.items[i];
        if (__float_as_int(r1.y)) {
            a.depth[pix] = r0.x; a.cd[pix] = r0.y;
            a.rbar[(size_t)pix * a.C] = r0.z;
            if (a.C == 3) { a.rbar[(size_t)pix * 3 + 1] = r0.w; a.rbar[(size_t)pix * 3 + 2] = r1.x; }
        } else {
            a.ce[pix] = 0.f; a.emask[pix] = 0;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int old = atomicAdd(a.blocks_done, 1);
        s_last = (old == (int)gridDim.x - 1);
        if (s_last) *a.blocks_done = 0;
    }
    __syncthreads();
    if (!s_last || !a.do_push) return;
    __threadfence();
    for (int j = 0; j < 4; ++j) p2p_push_rows(a.push, j, threadIdx.x, blockDim.x);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (a.push.up_flag) *reinterpret_cast<volatile unsigned*>(a.push.up_flag) = a.push.seq;
        if (a.push.dn_flag) *reinterpret_cast<volatile unsigned*>(a.push.dn_flag) = a.push.seq;
    }
}

/* layout helpers of the halo part: area(slot, side) with side 0 = from above, 1 = from below; flags after the areas */
static inline size_t p2p_area_off(const rslf_ctx* ctx, int slot, int side) { return ((size_t)slot * 2 + side) * ctx->p2p_area; }
static inline size_t p2p_flag_off(const rslf_ctx* ctx, int slot, int side) { return 4 * ctx->p2p_area + ((size_t)slot * 2 + side) * sizeof(unsigned); }

/* Fills the push arguments and the receiving side (median_halo) of one pass.  colour0 == nullptr: colours are read
 * in place from the replicated stack by the receiver. */
static void p2p_prepare_halo(rslf_ctx* ctx, const float* depth_plane, const uint8_t* mask_plane, const float* colour0,
                             size_t colour_row_stride_floats, int U, int C, const shard_tab& t, p2p_push_args* push, median_halo* out)
{
    const int r0 = ctx->rank, Vloc = t.b[r0 + 1] - t.b[r0];
    const unsigned seq = ++ctx->p2p_seq;
    const int slot = (int)(seq & 1u);
    p2p_push_args& a = *push;
    a.depth = depth_plane; a.mask = mask_plane; a.colour0 = colour0; a.colour_row_stride = colour_row_stride_floats;
    a.Vloc = Vloc; a.U = U; a.C = C; a.seq = seq; a.done = ctx->p2p_done; a.nblocks = 0;
    /* my first rows are the rows BELOW the block of the rank above: its side 1; my last rows: side 0 of the rank below */
    a.up_area = ctx->p2p_up ? ctx->p2p_up + p2p_area_off(ctx, slot, 1) : nullptr;
    a.dn_area = ctx->p2p_dn ? ctx->p2p_dn + p2p_area_off(ctx, slot, 0) : nullptr;
    a.up_flag = ctx->p2p_up ? reinterpret_cast<unsigned*>(ctx->p2p_up + p2p_flag_off(ctx, slot, 1)) : nullptr;
    a.dn_flag = ctx->p2p_dn ? reinterpret_cast<unsigned*>(ctx->p2p_dn + p2p_flag_off(ctx, slot, 0)) : nullptr;
    memset(out, 0, sizeof(*out));
    out->seq = seq; out->err = ctx->dev_err; out->colour_halo_stride = (size_t)U * C;
    if (r0 > 0) {
        const char* b = ctx->arena + p2p_area_off(ctx, slot, 0);
        out->top_depth = reinterpret_cast<const float*>(b);
        out->top_colour = reinterpret_cast<const float*>(b + (size_t)2 * U * 4);
        out->top_mask = reinterpret_cast<const uint8_t*>(b + (size_t)2 * U * 4 + (size_t)2 * U * C * 4);
        out->flag_top = reinterpret_cast<const volatile unsigned*>(ctx->arena + p2p_flag_off(ctx, slot, 0));
    }
    if (r0 + 1 < t.n) {
        const char* b = ctx->arena + p2p_area_off(ctx, slot, 1);
        out->bot_depth = reinterpret_cast<const float*>(b);
        out->bot_colour = reinterpret_cast<const float*>(b + (size_t)2 * U * 4);
        out->bot_mask = reinterpret_cast<const uint8_t*>(b + (size_t)2 * U * 4 + (size_t)2 * U * C * 4);
        out->flag_bot = reinterpret_cast<const volatile unsigned*>(ctx->arena + p2p_flag_off(ctx, slot, 1));
    }
}

static int comm_p2p_exchange_median_halo(rslf_ctx* ctx, const float* depth_plane, const uint8_t* mask_plane, const float* colour0,
                                         size_t colour_row_stride_floats, int U, int C, const shard_tab& t, median_halo* out)
{
    p2p_push_args a;
    p2p_prepare_halo(ctx, depth_plane, mask_plane, colour0, colour_row_stride_floats, U, C, t, &a, out);
    dim3 grid(std::max(1, std::min(8, (U + 127) / 128)), 4);
    a.nblocks = (int)(grid.x * grid.y);
    p2p_push_halo_kernel<<<grid, 128, 0, ctx->stream>>>(a);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    return RSLF_OK;
}

/* ---- halo rows of a row-sharded [S][Vloc][U] map + mask (fuse_disp_maps: bilinear upsampling across the block border,
 * final 3x3 median): the first / last FUSE_HALO rows of every plane go into the neighbours' per-level areas ---- */
struct svu_push_args {
    const float* map; const uint8_t* mask; int S, Vloc, U, H;
    char* up_area; char* dn_area; unsigned* up_flag; unsigned* dn_flag; unsigned seq; int* done; int nblocks;
};
__global__ void svu_push_halo_kernel(const svu_push_args a)
{
    const int j = blockIdx.y, s = blockIdx.z;                      /* j < H: my row j -> rank above; else my row Vloc - 2H + j -> rank below */
    char* dst = (j < a.H) ? a.up_area : a.dn_area;
    const int jj = (j < a.H) ? j : j - a.H;
    const int v = (j < a.H) ? j : a.Vloc - 2 * a.H + j;
    if (dst && v >= 0 && v < a.Vloc) {
        float* dm = reinterpret_cast<float*>(dst) + ((size_t)s * a.H + jj) * a.U;
        uint8_t* dk = reinterpret_cast<uint8_t*>(dst + (size_t)a.S * a.H * a.U * 4) + ((size_t)s * a.H + jj) * a.U;
        const float* sm = a.map + ((size_t)s * a.Vloc + v) * a.U;
        const uint8_t* sk = a.mask ? a.mask + ((size_t)s * a.Vloc + v) * a.U : nullptr;
        for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < a.U; u += gridDim.x * blockDim.x) {
            dm[u] = sm[u];
            if (sk) dk[u] = sk[u];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int old = atomicAdd(a.done, 1);
        if (old == a.nblocks - 1) {
            *a.done = 0;
            __threadfence_system();
            if (a.up_flag) *reinterpret_cast<volatile unsigned*>(a.up_flag) = a.seq;
            if (a.dn_flag) *reinterpret_cast<volatile unsigned*>(a.dn_flag) = a.seq;
        }
    }
}
__global__ void peer_wait_kernel(const volatile unsigned* f0, const volatile unsigned* f1, unsigned seq, int* err)
{
    if (threadIdx.x == 0 && f0) peer_wait32(f0, seq, err);
    if (threadIdx.x == 1 && f1) peer_wait32(f1, seq, err);
    __threadfence_system();
}

static inline size_t fuse_area_off(const arena_layout& l, int slot, int side) { return l.off_fuse + ((size_t)slot * 2 + side) * l.fuse_area; }
static inline size_t fuse_flag_off(const arena_layout& l, int slot, int side) { return l.off_flags + 64 + ((size_t)slot * 2 + side) * sizeof(unsigned); }

/* Sends this rank's border rows of `map` / `mask` ([S][Vloc][U], level slot `slot`) to the neighbours, waits for
 * theirs, and returns the windows (own rows + halo rows in the arena) for the kernels that read across the border. */
static int comm_svu_halo(rslf_ctx* ctx, int slot, const float* map, const uint8_t* mask, int S, int U, const shard_tab& t,
                         svu_view<float>* vm, svu_view<uint8_t>* vk)
{
    const arena_layout l = arena_current(ctx);
    const int r0 = ctx->rank, Vloc = t.b[r0 + 1] - t.b[r0], H = FUSE_HALO;
    const unsigned seq = ctx->fuse_seq;
    svu_push_args a;
    a.map = map; a.mask = mask; a.S = S; a.Vloc = Vloc; a.U = U; a.H = H; a.seq = seq; a.done = ctx->p2p_done + 4;
    a.up_area = ctx->p2p_up ? ctx->p2p_up + fuse_area_off(l, slot, 1) : nullptr;
    a.dn_area = ctx->p2p_dn ? ctx->p2p_dn + fuse_area_off(l, slot, 0) : nullptr;
    a.up_flag = ctx->p2p_up ? reinterpret_cast<unsigned*>(ctx->p2p_up + fuse_flag_off(l, slot, 1)) : nullptr;
    a.dn_flag = ctx->p2p_dn ? reinterpret_cast<unsigned*>(ctx->p2p_dn + fuse_flag_off(l, slot, 0)) : nullptr;
    dim3 grid(std::max(1, std::min(4, (U + 255) / 256)), 2 * H, S);
    a.nblocks = (int)(grid.x * grid.y * grid.z);
    svu_push_halo_kernel<<<grid, 256, 0, ctx->stream>>>(a);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    const volatile unsigned* f_top = r0 > 0 ? reinterpret_cast<const volatile unsigned*>(ctx->arena + fuse_flag_off(l, slot, 0)) : nullptr;
    const volatile unsigned* f_bot = r0 + 1 < t.n ? reinterpret_cast<const volatile unsigned*>(ctx->arena + fuse_flag_off(l, slot, 1)) : nullptr;
    peer_wait_kernel<<<1, 32, 0, ctx->stream>>>(f_top, f_bot, seq, ctx->dev_err);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 2;
    const char* top = ctx->arena + fuse_area_off(l, slot, 0);
    const char* bot = ctx->arena + fuse_area_off(l, slot, 1);
    vm->mid = map; vm->top = reinterpret_cast<const float*>(top); vm->bot = reinterpret_cast<const float*>(bot);
    vm->r0 = t.b[r0]; vm->rows = Vloc; vm->H = H;
    if (vk) {
        vk->mid = mask; vk->top = reinterpret_cast<const uint8_t*>(top + (size_t)S * H * U * 4);
        vk->bot = reinterpret_cast<const uint8_t*>(bot + (size_t)S * H * U * 4);
        vk->r0 = t.b[r0]; vk->rows = Vloc; vk->H = H;
    }
    return RSLF_OK;
}

extern "C" int rslf_cuda_nccl_unique_id(void* id128)
{
    if (!id128) return RSLF_ERR_ARG;
    char err[256];
    if (nccl_load(err, sizeof(err)) != RSLF_OK) return RSLF_ERR_NCCL;
    rslf_nccl_uid uid;
    if (g_nccl.GetUniqueId(&uid) != 0) return RSLF_ERR_NCCL;
    memcpy(id128, &uid, sizeof(uid));
    return RSLF_OK;
}

extern "C" int rslf_cuda_comm_init(rslf_ctx* ctx, const void* id128, int rank, int world)
{
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return RSLF_ERR_ARG;
    RSLF_CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    RSLF_TRY(nccl_load(ctx->err, sizeof(ctx->err)));
    comm_destroy(ctx);
    rslf_nccl_uid uid; memcpy(&uid, id128, sizeof(uid));
    rslf_disable_ring();                                 /* see g_no_ring (rslf_api.cu) */
    RSLF_NCCL_TRY(ctx, g_nccl.CommInitRank(&ctx->nccl_comm, world, uid, rank));
    ctx->rank = rank; ctx->world = world;
    return RSLF_OK;
}

extern "C" int rslf_cuda_set_row_shards(rslf_ctx* ctx, const int* row_starts, int n_ranks)
{
    if (!ctx || !row_starts || n_ranks < 1 || n_ranks > 64) return RSLF_ERR_ARG;
    if (n_ranks != ctx->world) { snprintf(ctx->err, sizeof(ctx->err), "shard table has %d ranks, communicator %d", n_ranks, ctx->world); return RSLF_ERR_ARG; }
    for (int r = 0; r < n_ranks; ++r)
        if (row_starts[r] < 0 || row_starts[r + 1] <= row_starts[r]) { snprintf(ctx->err, sizeof(ctx->err), "shard table must be strictly increasing"); return RSLF_ERR_ARG; }
    if (row_starts[0] != 0) return RSLF_ERR_ARG;
    for (int r = 0; r <= n_ranks; ++r) ctx->row_starts[r] = row_starts[r];
    ctx->v0 = row_starts[ctx->rank]; ctx->V_total = row_starts[n_ranks];
    ctx->have_shards = true;
    ++ctx->input_epoch;                                  /* the gathered copy of the stack follows the shard table */
    return RSLF_OK;
}
