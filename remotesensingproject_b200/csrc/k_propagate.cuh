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
    const size_t plane = (size_t)a.V * a.U;
    const int s_begin = blockIdx.z * PROP_SG;
    /* targets of the PROP_SG views of this thread; the mask bytes are fetched together */
    size_t tgt[PROP_SG];
    uint8_t rem[PROP_SG];
#pragma unroll
    for (int j = 0; j < PROP_SG; ++j) {
        const int s = s_begin + j;
        float t = cur * (float)(a.s_hat - s);
        t = t * a.slope;
        const int q = u + (int)roundf(t);
        const bool in = (s < a.S) && (q >= 0) && (q < a.U);
        tgt[j] = (size_t)min(s, a.S - 1) * plane + (size_t)v * a.U + min(max(q, 0), a.U - 1);
        rem[j] = in ? a.remaining[tgt[j]] : (uint8_t)0;
    }
    bool any = false;
#pragma unroll
    for (int j = 0; j < PROP_SG; ++j) any |= (rem[j] != 0);
    if (!any) return;
    if (PHASE == 0) {
        float rb[C];
#pragma unroll
        for (int c = 0; c < C; ++c) rb[c] = a.rbar_p[o * C + c];
#pragma unroll
        for (int j = 0; j < PROP_SG; ++j) {
            if (!rem[j]) continue;
            /* colour of the target pixel: E_v(s, q); tgt = (s*V + v)*U + q  ->  epi index ((v*S + s)*U + q)*C */
            const int s = s_begin + j;
            const size_t q = tgt[j] - ((size_t)s * plane + (size_t)v * a.U);
            const float* e = a.epi + (((size_t)v * a.S + s) * (size_t)a.U + q) * C;
            float ec[C];
#pragma unroll
            for (int c = 0; c < C; ++c) ec[c] = __ldg(e + c);
            if (rslf_norm_diff_lt<C>(ec, rb, a.eps, a.eps_T)) atomicMin(a.winner + tgt[j], u);
        }
    } else {
        /* the arbitration entry equals u only if this source passed every test in phase 0 and is the lowest */
        const float cdv = a.cd_p[o];
#pragma unroll
        for (int j = 0; j < PROP_SG; ++j) {
            if (!rem[j]) continue;
            if (a.winner[tgt[j]] == u) {
                a.depth[tgt[j]] = cur;
                a.cd[tgt[j]] = cdv;
                a.remaining[tgt[j]] = 0;
                a.winner[tgt[j]] = 0x7fffffff;
            }
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
