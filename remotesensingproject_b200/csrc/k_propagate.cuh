/*
 * k_propagate.cuh — propagation of the depths of line s_hat along their EPI
 * lines into all other views, with the "still to compute" bookkeeping.
 *
 * Replaces the propagation loop of compute_2D_depth_epi
 * (rslf_depth_computation_core.hpp:1086-1129).  For every pixel (v,u) of line
 * s_hat whose edge mask is set (the `#else` branch, :1097-1103) and every view s:
 *     u' = u + (int)round(depth(u) * (s_hat - s) * slope)
 *     if 0 <= u' < U  and  remaining[s](v,u')  and  norm(E_v(s,u') - rbar(u)) < eps:
 *         depth[s](v,u') = depth(u); C_d[s](v,u') = C_d(u); remaining[s](v,u') = 0
 * The reference walks u in ascending order inside a row and the first writer
 * clears the mask, so the winner of a target is the LOWEST source u that passes
 * the (static) tests.  Here: pass 1 takes atomicMin(u) per target in an
 * arbitration buffer, pass 2 lets exactly the winner commit and re-arms the
 * buffer.  Rows are independent (the reference's OpenMP axis, :1088).
 *
 * Sources come in two kinds.  (1) Pixels computed in this pass: they are the compacted work list of the
 * pass (sparse after the first pass) and carry their mean-shift radiance r_bar; one warp per list entry,
 * lanes over the views.  (2) Pixels of the line that were painted by an earlier pass: their r_bar is the zero it was
 * initialised with (dc.hpp:744), so they can only paint targets whose own colour norm is below eps, i.e.
 * confident-but-dark pixels; per (s, v) row the edge-confidence kernel counted those (rowdark), and a block
 * of the dense kernel returns at once when its rows hold none.  A pixel computed in this pass whose r_bar
 * happens to be exactly zero is handled by both kinds; the commit is exclusive (compare-and-swap on the arbitration entry).
 */
#pragma once
#include "rslf_common.cuh"

#define PROP_THREADS 128
#define PROP_SG 32              /* at most this many views per block of the dense kind (one bit each in the block's `live` word) */
#define PROP_DENSE_MIN 8192     /* list entries from which the list kind walks sources by lane instead of views by lane */
#define PROP_VS 4               /* view slices per group of 32 sources in that mode */

struct prop_args {
    const float* epi; int V, S, U; int s_hat; float slope; float eps; double eps_T;
    const uint8_t* emask_p;      /* plane s_hat */
    const float* filtered;       /* median-filtered depth plane of s_hat */
    const float* rbar_p;         /* plane s_hat, [V][U][C] */
    const float* cd_p;           /* plane s_hat */
    float* depth; float* cd; uint8_t* remaining; int* winner;   /* [S][V][U] */
    const int* items; const int* count;                         /* work list of the pass */
    const int* items2; const int* count2;                       /* second part of the list (row-sharded runs), or nullptr */
    int* rowdark;                                                /* [S][V] confident-and-dark pixels still unpainted (upper bound) */
    const int* dark_lo; const int* dark_hi;                      /* [S][V] column range that holds them (a bound; nullptr: no range test) */
    int criterion; float disp_thr;                               /* 1: sources are the pixels with C_d > disp_thr (core.hpp:1097-1098 as intended) */
};

/* "Only paint if the confidence threshold was high enough" (core.hpp:1096-1103) */
__device__ __forceinline__ bool prop_source(const prop_args& a, size_t pix)
{
    return a.criterion == 1 ? (a.cd_p[pix] > a.disp_thr) : (a.emask_p[pix] != 0);
}

/* one (source, view) pair; PHASE 0: arbitration, PHASE 1: commit */
template <int C, int PHASE>
__device__ __forceinline__ void propagate_one(const prop_args& a, int v, int u, int s, float cur, const float (&rb)[C], float cdv)
{
    float t = cur * (float)(a.s_hat - s);
    t = t * a.slope;
    const int q = u + (int)roundf(t);
    if (q < 0 || q >= a.U) return;
    const size_t tgt = ((size_t)s * a.V + v) * (size_t)a.U + q;
    if (!a.remaining[tgt]) return;
    if (PHASE == 0) {
        const float* e = a.epi + (((size_t)v * a.S + s) * (size_t)a.U + q) * C;
        float ec[C];
#pragma unroll
        for (int c = 0; c < C; ++c) ec[c] = __ldg(e + c);
        if (rslf_norm_diff_lt<C>(ec, rb, a.eps, a.eps_T)) atomicMin(a.winner + tgt, u);
    } else if (a.winner[tgt] == u && atomicCAS(a.winner + tgt, u, 0x7fffffff) == u) {
        /* the arbitration entry equals u only if this source passed every test in phase 0 and is the lowest; the
         * compare-and-swap re-arms the entry and makes the commit exclusive: a source computed in this pass whose
         * r_bar is exactly zero is seen by both kinds of blocks, and only one of them may count the dark target down */
        a.depth[tgt] = cur;
        a.cd[tgt] = cdv;
        a.remaining[tgt] = 0;
        if (a.rowdark) {
            const float* e = a.epi + (((size_t)v * a.S + s) * (size_t)a.U + q) * C;
            float ec[C], z[C];
#pragma unroll
            for (int c = 0; c < C; ++c) { ec[c] = __ldg(e + c); z[c] = 0.f; }
            if (rslf_norm_diff_lt<C>(ec, z, a.eps, a.eps_T)) atomicSub(a.rowdark + (size_t)v * a.S + s, 1);
        }
    }
}

/* One launch per phase: the first list_blocks blocks walk the pass's work list (kind 1: one warp per list
 * entry, whose source data is loaded once, lanes over views), the remaining blocks are (row v, group of PROP_SG
 * views) pairs for the pixels painted earlier (kind 2: r_bar == 0), threads over u. */
template <int C, int PHASE>
__global__ void __launch_bounds__(PROP_THREADS)
propagate_kernel(const prop_args a, int list_blocks, int sg)
{
    /* sg: views per block of the second kind (<= PROP_SG) */
    if ((int)blockIdx.x < list_blocks) {
        const int n1 = *a.count, n = n1 + (a.count2 ? *a.count2 : 0);
        const int lane = threadIdx.x & 31;
        const int warps = (list_blocks * PROP_THREADS) >> 5;
        if (n >= PROP_DENSE_MIN) {
            /* Dense list (the first passes of a level: up to a million sources): lanes over 32 CONSECUTIVE list entries —
             * the compaction keeps the pixels of a 32-pixel stretch in order — and a loop over a slice of the views.  In
             * one view the 32 lanes then touch neighbouring targets (sources that lie side by side mostly carry similar
             * depths): the mask bytes and the colours of a warp's access fall into a few sectors instead of 32 planes. */
            const int groups = (n + 31) >> 5;
            for (int item = (blockIdx.x * PROP_THREADS + threadIdx.x) >> 5; item < groups * PROP_VS; item += warps) {
                const int g = item / PROP_VS, vs = item - g * PROP_VS;
                const int e = (g << 5) + lane;
                int pix = 0; bool ok = false;
                if (e < n) { pix = (e < n1) ? a.items[e] : a.items2[e - n1]; ok = prop_source(a, pix); }
                const int v = pix / a.U, u = pix - v * a.U;
                float rb[C];
#pragma unroll
                for (int c = 0; c < C; ++c) rb[c] = ok ? a.rbar_p[(size_t)pix * C + c] : 0.f;
                const float cur = ok ? a.filtered[pix] : 0.f;
                const float cdv = (PHASE && ok) ? a.cd_p[pix] : 0.f;
                const int s0 = a.S * vs / PROP_VS, s1 = a.S * (vs + 1) / PROP_VS;
                if (ok)
                    for (int s = s0; s < s1; ++s) propagate_one<C, PHASE>(a, v, u, s, cur, rb, cdv);
            }
            return;
        }
        for (int it = (blockIdx.x * PROP_THREADS + threadIdx.x) >> 5; it < n; it += warps) {
            const int pix = (it < n1) ? a.items[it] : a.items2[it - n1];
            if (!prop_source(a, pix)) continue;             /* e.g. score <= threshold: dropped by the depth kernel */
            const int v = pix / a.U, u = pix - v * a.U;
            float rb[C];
#pragma unroll
            for (int c = 0; c < C; ++c) rb[c] = a.rbar_p[(size_t)pix * C + c];
            const float cur = a.filtered[pix];
            const float cdv = PHASE ? a.cd_p[pix] : 0.f;
            for (int s = lane; s < a.S; s += 32) propagate_one<C, PHASE>(a, v, u, s, cur, rb, cdv);
        }
        return;
    }
    const int idx = (int)blockIdx.x - list_blocks;
    const int v = idx % a.V;
    const int s_begin = (idx / a.V) * sg, s_end = min(a.S, s_begin + sg);
    /* views of the group that still hold a dark target in this row, and the column range those targets lie in: a source
     * without r_bar can only paint such targets, so a (source, view) pair whose target column falls outside the range
     * is rejected before any memory is touched */
    __shared__ int s_lo[PROP_SG], s_hi[PROP_SG];
    __shared__ unsigned s_live;
    if (threadIdx.x == 0) s_live = 0;
    __syncthreads();
    if ((int)threadIdx.x < s_end - s_begin) {
        const int s = s_begin + threadIdx.x;
        const bool lv = a.rowdark[(size_t)v * a.S + s] > 0;
        s_lo[threadIdx.x] = a.dark_lo ? a.dark_lo[(size_t)v * a.S + s] : 0;
        s_hi[threadIdx.x] = a.dark_hi ? a.dark_hi[(size_t)v * a.S + s] : a.U - 1;
        if (lv) atomicOr(&s_live, 1u << threadIdx.x);
    }
    __syncthreads();
    const unsigned live = s_live;
    if (!live) return;
    for (int u = threadIdx.x; u < a.U; u += PROP_THREADS) {
        const size_t o = (size_t)v * a.U + u;
        if (!prop_source(a, o)) continue;
        float rb[C]; bool zero = true;
#pragma unroll
        for (int c = 0; c < C; ++c) { rb[c] = a.rbar_p[o * C + c]; zero = zero && (rb[c] == 0.f); }
        if (!zero) continue;
        const float cur = a.filtered[o];
        const float cdv = PHASE ? a.cd_p[o] : 0.f;
        for (int s = s_begin; s < s_end; ++s) {
            if (!(live & (1u << (s - s_begin)))) continue;
            float t = cur * (float)(a.s_hat - s);                       /* the target column, as in propagate_one */
            t = t * a.slope;
            const int q = u + (int)roundf(t);
            if (q < s_lo[s - s_begin] || q > s_hi[s - s_begin]) continue;
            propagate_one<C, PHASE>(a, v, u, s, cur, rb, cdv);
        }
    }
}

static int launch_propagate(rslf_ctx* ctx, int C, const prop_args& a)
{
    const int lb = ctx->num_sm * 16;
    /* second kind: (row, group of sg views) blocks.  Many rows: 32 views per block keep the launch small (most blocks
     * return at once); few rows (a rank's block of a row-sharded level, a coarse level): 8 views per block, so that the
     * rows that do hold dark targets are spread over enough blocks */
    const int sg = ((long long)a.V * rslf_div_up(a.S, PROP_SG) >= 2048) ? PROP_SG : 8;
    const int grid = lb + a.V * rslf_div_up(a.S, sg);
    if (C == 1) {
        propagate_kernel<1, 0><<<grid, PROP_THREADS, 0, ctx->stream>>>(a, lb, sg);
        propagate_kernel<1, 1><<<grid, PROP_THREADS, 0, ctx->stream>>>(a, lb, sg);
    } else {
        propagate_kernel<3, 0><<<grid, PROP_THREADS, 0, ctx->stream>>>(a, lb, sg);
        propagate_kernel<3, 1><<<grid, PROP_THREADS, 0, ctx->stream>>>(a, lb, sg);
    }
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 2;
    return RSLF_OK;
}
