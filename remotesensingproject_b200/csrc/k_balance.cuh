/*
 * k_balance.cuh — work records of a pass and the pass-balanced multi-GPU protocol.
 *
 * Work records.  compact_kernel (k_depth.cuh) lists the pixels of line s_hat that are still to be computed
 * (core.hpp:510-516) and writes, next to the list, one 16-byte RECORD per pixel: {pixel index in the level's
 * full EPI stack, dmin, dmax, C_e}.  The depth kernels read nothing else about a pixel, so a pixel can be
 * evaluated by ANY GPU that holds the EPI stack.
 *
 * Pass-balanced row sharding (multi-GPU).  The result maps stay sharded by image rows (the reference's OpenMP
 * axis, core.hpp:743, 799, 1088) but the EPI stack of every level is replicated, and the pixels of a pass are
 * split EVENLY over the ranks whatever rows they lie in: the ranks advance in lock-step through the per-pass
 * median halo, so only an even split of every single pass keeps all GPUs busy (a contiguous row cut can balance
 * the sum over the passes, never each pass).  Per pass, all over NVLink peer memory (CUDA IPC), no collective:
 *
 *   owner r   compact_kernel:   list + records of its rows; the last block publishes the count n_r to every
 *                               peer: counts[r] <- (seq << 32 | n_r)                                   (P2P store)
 *   worker q  depth kernel:     waits for the N counts, takes pixels [T q / N, T (q+1) / N) of the concatenated
 *                               lists (T = sum n_r), loads their records from the owners (P2P load, .cv),
 *                               stores a 32-byte RESULT record per pixel into the owner's result list (P2P
 *                               store); its last block raises done[q] <- seq on every peer
 *   owner r   bal_apply_kernel: waits for the N done flags, scatters the result records into its maps, and its
 *                               last block pushes the two first / last rows of the depth and mask planes of line
 *                               s_hat into the neighbours' halo areas (the 5x5 selective median, core.hpp:698-709)
 *   owner r   median, propagation: row-local (k_median.cuh waits for the neighbours' halo flags)
 *
 * Every buffer is single-buffered and race-free: an owner rewrites its records / counts only after its apply
 * kernel saw all done flags of the pass, i.e. after every worker finished reading them; a worker writes results
 * of pass k+1 only after the owner published the count of pass k+1, i.e. after the owner applied pass k.
 * Waits are bounded by the global timer; a peer that never answers raises an error word that the host reports
 * (RSLF_ERR_PEER) instead of trapping the context.
 */
#pragma once
#include "rslf_common.cuh"

#define RSLF_MAX_PEERS 16
#define RSLF_PEER_TIMEOUT_NS 60000000000ULL          /* 60 s */

struct depth_bal {
    int n, rank;                                     /* ranks sharing the pass; n = 0: direct (single-rank) mode */
    unsigned seq;                                    /* sequence number of the pass */
    const int4* rec[RSLF_MAX_PEERS];                 /* record list of every rank (own entry: local pointer) */
    float4* res[RSLF_MAX_PEERS];                     /* result list of every rank, two float4 per pixel */
    const volatile unsigned long long* counts;       /* local: counts[r] = (seq << 32 | n_r), written by rank r */
    unsigned long long* done[RSLF_MAX_PEERS];        /* rank r's done[this rank] */
    int* blocks_done;                                /* local block counter (last-block detection) */
    int* err;                                        /* local error word */
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long rslf_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

/* waits until the sequence half of *p reaches seq; raises *err after RSLF_PEER_TIMEOUT_NS */
__device__ __forceinline__ unsigned long long peer_wait64(const volatile unsigned long long* p, unsigned seq, int* err)
{
    unsigned long long v = *p;
    if ((int)((unsigned)(v >> 32) - seq) >= 0) return v;
    const unsigned long long t0 = rslf_globaltimer();
    for (;;) {
        v = *p;
        if ((int)((unsigned)(v >> 32) - seq) >= 0) return v;
        __nanosleep(100);
        if (rslf_globaltimer() - t0 > RSLF_PEER_TIMEOUT_NS) { atomicExch(err, 1); return v; }
    }
}
__device__ __forceinline__ void peer_wait32(const volatile unsigned* p, unsigned seq, int* err)
{
    if ((int)(*p - seq) >= 0) return;
    const unsigned long long t0 = rslf_globaltimer();
    for (;;) {
        if ((int)(*p - seq) >= 0) return;
        __nanosleep(100);
        if (rslf_globaltimer() - t0 > RSLF_PEER_TIMEOUT_NS) { if (err) atomicExch(err, 1); return; }
    }
}

/* this rank's share of the pass: pixels [lo, lo + n) of the concatenated lists; lane r < ranks keeps the
 * exclusive / inclusive prefix of rank r's count */
struct bal_slice { int lo, n, pre_excl, pre_incl; };

/* Host/device: the share of rank q of T pixels over n ranks */
static __host__ __device__ inline void bal_share(long long T, int q, int n, int* lo, int* cnt)
{
    const long long a = T * q / n, b = T * (q + 1) / n;
    *lo = (int)a; *cnt = (int)(b - a);
}

__device__ __forceinline__ bal_slice bal_begin(const depth_bal& b, int lane)
{
    int c = 0;
    if (lane < b.n) c = (int)(unsigned)(peer_wait64(b.counts + lane, b.seq, b.err) & 0xffffffffULL);
    __threadfence_system();
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    const int T = __shfl_sync(0xffffffffu, incl, 31);
    bal_slice s;
    s.pre_incl = incl; s.pre_excl = incl - c;
    bal_share(T, b.rank, b.n, &s.lo, &s.n);
    return s;
}

/* owner and index inside the owner's list of global pixel g (warp-uniform) */
__device__ __forceinline__ void bal_locate(const bal_slice& s, int g, int ranks, int lane, int& owner, int& idx)
{
    const unsigned m = __ballot_sync(0xffffffffu, lane < ranks && s.pre_incl <= g);
    owner = __popc(m);
    owner = owner < ranks ? owner : ranks - 1;
    idx = g - __shfl_sync(0xffffffffu, s.pre_excl, owner);
}

__device__ __forceinline__ int4 ld_cv_int4(const int4* p)
{
    int4 v;
    asm volatile("ld.global.cv.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_cv_float4(const float4* p)
{
    float4 v;
    asm volatile("ld.global.cv.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

/* last block of a grid raises flag[rank] <- (seq << 32 | value) on every peer; every thread that wrote peer
 * memory must have executed __threadfence_system() before the block calls this (all threads of the block) */
__device__ __forceinline__ void bal_last_block_signal(int* blocks_done, unsigned long long* const* peer_slots, int ranks,
                                                      unsigned seq, unsigned value)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int old = atomicAdd(blocks_done, 1);
        if (old == (int)(gridDim.x * gridDim.y * gridDim.z) - 1) {
            *blocks_done = 0;
            __threadfence_system();
            const unsigned long long v = ((unsigned long long)seq << 32) | (unsigned long long)value;
            for (int r = 0; r < ranks; ++r) *reinterpret_cast<volatile unsigned long long*>(peer_slots[r]) = v;
        }
    }
}
#endif
