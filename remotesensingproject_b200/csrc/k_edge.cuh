/*
 * k_edge.cuh — edge confidence C_e and its mask.
 *
 * Replaces compute_1D_edge_confidence (rslf_depth_computation_core.hpp:426-478,
 * squares summed by _square_sum_channels_into, src/rslf_depth_computation_core.cpp:6-23)
 * and its wrappers _pile (:728-770) and compute_2D_edge_confidence (:901-931).
 *
 *   C_e(u) = sum_{j != centre} sum_c (E_c(u) - E_c(reflect101(u + j - centre)))^2
 * accumulated in the order j ascending, c ascending, one float rounding per
 * subtract / square / add; then C_e = 0 where norm(E(u)) < shadow level, and
 * mask = C_e > threshold.
 *
 * HBM-bound stream: 4C bytes read + 5 bytes written per pixel.  One block owns
 * a segment of one EPI scanline (v, s); the segment plus its halo is staged in
 * shared memory once (coalesced, border rule applied while staging), so HBM
 * sees every value once; the fs taps are then read from shared memory
 * (stride-C word addresses: conflict-free for C = 1 and 3).
 */
#pragma once
#include "rslf_common.cuh"

#define EDGE_THREADS 128
#define EDGE_PPT 4
#define EDGE_TILE (EDGE_THREADS * EDGE_PPT)
#define EDGE_MAX_FS 17

template <int C>
__global__ void __launch_bounds__(EDGE_THREADS)
edge_confidence_kernel(const float* __restrict__ epi, int V, int S, int U, int s_first, int s_count,
                       int fs, int cut_shadows, float shadow_level, double shadow_T, float thr,
                       float* __restrict__ ce_out, uint8_t* __restrict__ mask_out,
                       float dark_eps, double dark_T, int* __restrict__ rowdark)
{
    extern __shared__ float seg[];                 /* (EDGE_TILE + fs - 1) * C floats */
    const int centre = (fs - 1) / 2;
    const int row = blockIdx.x;                    /* row = v * s_count + si */
    const int v = row / s_count, si = row % s_count;
    const int s = s_first + si;
    const int u0 = blockIdx.y * EDGE_TILE;
    const float* src = epi + ((size_t)v * S + s) * (size_t)U * C;
    const int span = min(EDGE_TILE, U - u0) + fs - 1;
    /* stage [u0 - centre, u0 + tile + fs - 1 - centre) with the border rule applied */
    for (int i = threadIdx.x; i < span * C; i += EDGE_THREADS) {
        int p = i / C, c = i - p * C;
        int q = rslf_reflect101(u0 + p - centre, U);
        seg[i] = __ldg(src + (size_t)q * C + c);
    }
    __syncthreads();
    const size_t out_row = ((size_t)si * V + v) * (size_t)U;
    int ndark = 0;
#pragma unroll
    for (int k = 0; k < EDGE_PPT; ++k) {
        /* pixel index inside the tile: consecutive threads take consecutive pixels (coalesced stores) */
        const int p = k * EDGE_THREADS + threadIdx.x;
        const int u = u0 + p;
        if (u >= U) continue;
        float x[C];
#pragma unroll
        for (int c = 0; c < C; ++c) x[c] = seg[(p + centre) * C + c];
        float acc = 0.f;
        for (int j = 0; j < fs; ++j) {
            if (j == centre) continue;
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float diff = x[c] - seg[(p + j) * C + c];
                float sq = diff * diff;
                acc = acc + sq;
            }
        }
        if (cut_shadows) {
            bool dark;
            if (C == 1) dark = rslf_norm1_lt(x[0], shadow_level);
            else dark = rslf_norm3_lt(x[0], x[C > 1 ? 1 : 0], x[C > 2 ? 2 : 0], shadow_T);
            if (dark) acc = 0.f;
        }
        ce_out[out_row + u] = acc;
        mask_out[out_row + u] = (acc > thr) ? 255 : 0;
        if (rowdark && acc > thr) {
            /* confident pixels darker than the propagation epsilon: the only ones a source without r_bar
             * (a pixel that was itself propagated) can still paint (see k_propagate.cuh) */
            bool dk;
            if (C == 1) dk = rslf_norm1_lt(x[0], dark_eps);
            else dk = rslf_norm3_lt(x[0], x[C > 1 ? 1 : 0], x[C > 2 ? 2 : 0], dark_T);
            ndark += dk ? 1 : 0;
        }
    }
    if (rowdark) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ndark += __shfl_xor_sync(0xffffffffu, ndark, off);
        if ((threadIdx.x & 31) == 0 && ndark) atomicAdd(rowdark + (size_t)si * V + v, ndark);
    }
}

/* Launch for lines [s_first, s_first + s_count) of every row; output planes are
 * [s_count][V][U] starting at ce_out / mask_out. */
static int launch_edge_confidence(rslf_ctx* ctx, const float* epi, int V, int S, int U, int C, int s_first,
                                  int s_count, const rslf_params& P, float* ce_out, uint8_t* mask_out,
                                  int* rowdark = nullptr)
{
    int fs = P.edge_confidence_filter_size;
    if (fs < 1 || fs > EDGE_MAX_FS) {
        snprintf(ctx->err, sizeof(ctx->err), "edge_confidence_filter_size %d outside [1,%d]", fs, EDGE_MAX_FS);
        return RSLF_ERR_UNSUPPORTED;
    }
    if (P.edge_confidence_opening_size > 1) {
        snprintf(ctx->err, sizeof(ctx->err), "morphological opening of the edge mask is not implemented");
        return RSLF_ERR_UNSUPPORTED;
    }
    dim3 grid((unsigned)((size_t)V * s_count), rslf_div_up(U, EDGE_TILE));
    size_t smem = (size_t)(EDGE_TILE + fs - 1) * C * sizeof(float);
    double T = rslf_sq_threshold(P.shadow_level);
    const double dT = rslf_sq_threshold(P.propagation_epsilon);
    if (C == 1)
        edge_confidence_kernel<1><<<grid, EDGE_THREADS, smem, ctx->stream>>>(
            epi, V, S, U, s_first, s_count, fs, P.cut_shadows, P.shadow_level, T, P.edge_score_threshold, ce_out, mask_out,
            P.propagation_epsilon, dT, rowdark);
    else
        edge_confidence_kernel<3><<<grid, EDGE_THREADS, smem, ctx->stream>>>(
            epi, V, S, U, s_first, s_count, fs, P.cut_shadows, P.shadow_level, T, P.edge_score_threshold, ce_out, mask_out,
            P.propagation_epsilon, dT, rowdark);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    return RSLF_OK;
}
