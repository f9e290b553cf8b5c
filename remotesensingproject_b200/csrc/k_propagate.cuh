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
 * One thread per (source pixel, group of PROP_SG views).
 */
#pragma once
#include "rslf_common.cuh"

#define PROP_THREADS 128
#define PROP_SG 8

struct prop_args {
    const float* epi; int V, S, U; int s_hat; float slope; float eps; double eps_T;
    const uint8_t* emask_p;      /* plane s_hat */
    const float* filtered;       /* median-filtered depth plane of s_hat */
    const float* rbar_p;         /* plane s_hat, [V][U][C] */
    const float* cd_p;           /* plane s_hat */
    float* depth; float* cd; uint8_t* remaining; int* winner;   /* [S][V][U] */
};

template <int C, int PHASE>
__global__ void __launch_bounds__(PROP_THREADS)
propagate_kernel(const prop_args a)
{
    const int u = blockIdx.x * PROP_THREADS + threadIdx.x;
    const int v = blockIdx.y;
    if (u >= a.U) return;
    const size_t o = (size_t)v * a.U + u;
    if (!a.emask_p[o]) return;
    const float cur = a.filtered[o];
    float rb[C];
#pragma unroll
    for (int c = 0; c < C; ++c) rb[c] = a.rbar_p[o * C + c];
    const float cdv = (PHASE == 1) ? a.cd_p[o] : 0.f;
    const size_t plane = (size_t)a.V * a.U;
    const int s_begin = blockIdx.z * PROP_SG;
    const int s_end = min(a.S, s_begin + PROP_SG);
    for (int s = s_begin; s < s_end; ++s) {
        float t = cur * (float)(a.s_hat - s);
        t = t * a.slope;
        const int q = u + (int)roundf(t);
        if (q < 0 || q >= a.U) continue;
        const size_t tgt = (size_t)s * plane + (size_t)v * a.U + q;
        if (!a.remaining[tgt]) continue;
        const float* e = a.epi + (((size_t)v * a.S + s) * (size_t)a.U + q) * C;
        float ec[C];
#pragma unroll
        for (int c = 0; c < C; ++c) ec[c] = __ldg(e + c);
        if (!rslf_norm_diff_lt<C>(ec, rb, a.eps, a.eps_T)) continue;
        if (PHASE == 0) {
            atomicMin(a.winner + tgt, u);
        } else if (a.winner[tgt] == u) {
            a.depth[tgt] = cur;
            a.cd[tgt] = cdv;
            a.remaining[tgt] = 0;
            a.winner[tgt] = 0x7fffffff;
        }
    }
}

static int launch_propagate(rslf_ctx* ctx, int C, const prop_args& a)
{
    dim3 grid(rslf_div_up(a.U, PROP_THREADS), a.V, rslf_div_up(a.S, PROP_SG));
    if (C == 1) {
        propagate_kernel<1, 0><<<grid, PROP_THREADS, 0, ctx->stream>>>(a);
        propagate_kernel<1, 1><<<grid, PROP_THREADS, 0, ctx->stream>>>(a);
    } else {
        propagate_kernel<3, 0><<<grid, PROP_THREADS, 0, ctx->stream>>>(a);
        propagate_kernel<3, 1><<<grid, PROP_THREADS, 0, ctx->stream>>>(a);
    }
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 2;
    return RSLF_OK;
}
