/*
 * rslf_comm.cuh — NCCL plumbing for the row-sharded run (one process per GPU).
 * NCCL is resolved with dlopen at communicator creation, so the library loads
 * (and single-GPU runs work) on hosts without libnccl.  The path has two
 * exchange steps only: the per-pass all-gather of the depth / mask / colour rows
 * of line s_hat that the cross-row selective median reads
 * (rslf_depth_computation_core.hpp:698-709), and the scalar max of the input
 * normalisation (rslf_depth_computation.hpp:442-460).
 */
#pragma once
#include <dlfcn.h>
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
enum { RSLF_NCCL_UINT8 = 1, RSLF_NCCL_FLOAT = 7, RSLF_NCCL_MAX = 2 };

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

static void comm_destroy(rslf_ctx* ctx)
{
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
    RSLF_NCCL_TRY(ctx, g_nccl.CommInitRank(&ctx->nccl_comm, world, uid, rank));
    ctx->rank = rank; ctx->world = world;
    return RSLF_OK;
}

extern "C" int rslf_cuda_set_row_shard(rslf_ctx* ctx, int v0, int V_total)
{
    if (!ctx || v0 < 0 || V_total < 1) return RSLF_ERR_ARG;
    ctx->v0 = v0; ctx->V_total = V_total;
    return RSLF_OK;
}
