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
                       float dark_eps, double dark_T, int* __restrict__ rowdark, int v_first,
                       int* __restrict__ dark_lo, int* __restrict__ dark_hi)
{
    /* the grid covers rows v_first .. of the V-row planes (a chunk of EPIs during the pipelined ingest, else all) */
    extern __shared__ float seg[];                 /* (EDGE_TILE + fs - 1) * C floats */
    const int centre = (fs - 1) / 2;
    const int row = blockIdx.x;                    /* row = (v - v_first) * s_count + si */
    const int v = v_first + row / s_count, si = row % s_count;
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
    int ndark = 0, dlo = 0x7fffffff, dhi = -1;         /* dark confident pixels of the row: count and column range */
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
            if (dk) { dlo = min(dlo, u); dhi = max(dhi, u); }
        }
    }
    if (rowdark) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            ndark += __shfl_xor_sync(0xffffffffu, ndark, off);
            dlo = min(dlo, __shfl_xor_sync(0xffffffffu, dlo, off)); dhi = max(dhi, __shfl_xor_sync(0xffffffffu, dhi, off));
        }
        if ((threadIdx.x & 31) == 0 && ndark) {
            atomicAdd(rowdark + (size_t)v * s_count + si, ndark);
            if (dark_lo) { atomicMin(dark_lo + (size_t)v * s_count + si, dlo); atomicMax(dark_hi + (size_t)v * s_count + si, dhi); }
        }
    }
}

/*
 * Optional morphological opening of every V x U edge mask (core.hpp:759-769):
 * cv::morphologyEx(mask, mask, MORPH_OPEN, getStructuringElement(type, Size(k, k))), i.e. an erosion followed by a
 * dilation with the same element, anchor (k/2, k/2), BORDER_CONSTANT with morphologyDefaultBorderValue() (positions
 * outside the image take no part).  Every row of a RECT / CROSS / ELLIPSE element is one contiguous span
 * [j1, j2), so the element travels in the kernel arguments as k spans.  HBM-bound byte stream: the k x k
 * neighbourhood of a pixel comes out of L1/L2, 1 B read + 1 B written per pixel and pass from HBM.
 */
#define MORPH_MAX_K 31
#define MORPH_THREADS 128
struct morph_elem { int k; signed char j1[MORPH_MAX_K]; signed char j2[MORPH_MAX_K]; };

/* cv::getStructuringElement(shape, Size(k, k)), default anchor: 0 = MORPH_RECT, 1 = MORPH_CROSS, 2 = MORPH_ELLIPSE */
static bool morph_build(morph_elem& e, int shape, int k)
{
    if (k < 1 || k > MORPH_MAX_K || shape < 0 || shape > 2) return false;
    e.k = k;
    const int r = k / 2, c = k / 2;
    const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < MORPH_MAX_K; ++i) { e.j1[i] = 0; e.j2[i] = 0; }
    for (int i = 0; i < k; ++i) {
        int j1 = 0, j2 = 0;
        if (shape == 0 || (shape == 1 && i == r)) j2 = k;
        else if (shape == 1) { j1 = c; j2 = c + 1; }
        else {
            const int dy = i - r;
            if (abs(dy) <= r) {
                const int dx = (int)nearbyint(c * sqrt((r * r - dy * dy) * inv_r2));     /* cvRound: half to even */
                j1 = c - dx > 0 ? c - dx : 0;
                j2 = c + dx + 1 < k ? c + dx + 1 : k;
            }
        }
        e.j1[i] = (signed char)j1; e.j2[i] = (signed char)j2;
    }
    return true;
}

/* grid (U tiles, V, planes); ERODE: min over the element, else max */
template <int ERODE>
__global__ void __launch_bounds__(MORPH_THREADS)
morph_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int V, int U, const morph_elem e)
{
    const int u = blockIdx.x * MORPH_THREADS + threadIdx.x;
    if (u >= U) return;
    const int v = blockIdx.y;
    const size_t plane_off = (size_t)blockIdx.z * (size_t)V * U;
    const int a = e.k / 2;
    int acc = ERODE ? 255 : 0;
    for (int ky = 0; ky < e.k; ++ky) {
        const int yy = v + ky - a;
        if (yy < 0 || yy >= V) continue;
        const uint8_t* row = in + plane_off + (size_t)yy * U;
        const int x0 = max(u + (int)e.j1[ky] - a, 0), x1 = min(u + (int)e.j2[ky] - a, U);
        for (int x = x0; x < x1; ++x) {
            const int val = row[x];
            acc = ERODE ? min(acc, val) : max(acc, val);
        }
    }
    out[plane_off + (size_t)v * U + u] = (uint8_t)acc;
}

/* confident-and-dark pixels per (line, row) of a mask that was opened after the edge kernel (same count as the edge
 * kernel's `rowdark`); one block per (v, si) row like the edge kernel */
template <int C>
__global__ void __launch_bounds__(EDGE_THREADS)
rowdark_count_kernel(const float* __restrict__ epi, const uint8_t* __restrict__ mask, int V, int S, int U, int s_first,
                     int s_count, float dark_eps, double dark_T, int* __restrict__ rowdark,
                     int* __restrict__ dark_lo, int* __restrict__ dark_hi)
{
    const int row = blockIdx.x;
    const int v = row / s_count, si = row % s_count;
    const float* src = epi + ((size_t)v * S + (s_first + si)) * (size_t)U * C;
    const uint8_t* m = mask + ((size_t)si * V + v) * (size_t)U;
    int ndark = 0, dlo = 0x7fffffff, dhi = -1;
    for (int u = threadIdx.x; u < U; u += EDGE_THREADS) {
        if (!m[u]) continue;
        float x[C];
#pragma unroll
        for (int c = 0; c < C; ++c) x[c] = __ldg(src + (size_t)u * C + c);
        bool dk;
        if (C == 1) dk = rslf_norm1_lt(x[0], dark_eps);
        else dk = rslf_norm3_lt(x[0], x[C > 1 ? 1 : 0], x[C > 2 ? 2 : 0], dark_T);
        ndark += dk ? 1 : 0;
        if (dk) { dlo = min(dlo, u); dhi = max(dhi, u); }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        ndark += __shfl_xor_sync(0xffffffffu, ndark, off);
        dlo = min(dlo, __shfl_xor_sync(0xffffffffu, dlo, off)); dhi = max(dhi, __shfl_xor_sync(0xffffffffu, dhi, off));
    }
    if ((threadIdx.x & 31) == 0 && ndark) {
        atomicAdd(rowdark + (size_t)v * s_count + si, ndark);
        if (dark_lo) { atomicMin(dark_lo + (size_t)v * s_count + si, dlo); atomicMax(dark_hi + (size_t)v * s_count + si, dhi); }
    }
}

/* Launch for lines [s_first, s_first + s_count) of every row; output planes are
 * [s_count][V][U] starting at ce_out / mask_out.  mask_tmp: scratch of the same size as mask_out, needed only
 * when the opening is enabled (edge_confidence_opening_size > 1). */
static int launch_edge_confidence(rslf_ctx* ctx, const float* epi, int V, int S, int U, int C, int s_first,
                                  int s_count, const rslf_params& P, float* ce_out, uint8_t* mask_out,
                                  int* rowdark = nullptr, uint8_t* mask_tmp = nullptr, int rows_above = 0, int rows_below = 0,
                                  int* dark_lo = nullptr, int* dark_hi = nullptr)
{
    /* Row-sharded level with the opening switched on: the element reaches k/2 rows beyond the block in the erosion and
     * again in the dilation.  Every rank holds the whole EPI stack, so it computes C_e and the mask on its rows PLUS
     * 2 (k/2) rows either side (as far as the image goes: rows_above / rows_below say how many exist), opens that
     * extended mask — whose outer 2 (k/2) rows are then wrong, and whose inner rows are exact — and keeps its rows.
     * No exchange. */
    if (P.edge_confidence_opening_size > 1 && (rows_above > 0 || rows_below > 0)) {
        const int h = 2 * (P.edge_confidence_opening_size / 2);
        const int ha = std::min(h, rows_above), hb = std::min(h, rows_below), Vx = V + ha + hb;
        const size_t pxx = (size_t)s_count * Vx * U;
        if (ctx->open_cap < pxx) {
            if (ctx->open_ce) cudaFree(ctx->open_ce);
            if (ctx->open_mask) cudaFree(ctx->open_mask);
            if (ctx->open_tmp) cudaFree(ctx->open_tmp);
            ctx->open_ce = nullptr; ctx->open_mask = ctx->open_tmp = nullptr; ctx->open_cap = 0;
            if (cudaMalloc((void**)&ctx->open_ce, pxx * 4) != cudaSuccess || cudaMalloc((void**)&ctx->open_mask, pxx) != cudaSuccess ||
                cudaMalloc((void**)&ctx->open_tmp, pxx) != cudaSuccess) {
                snprintf(ctx->err, sizeof(ctx->err), "cudaMalloc(extended edge planes, %zu pixels)", pxx);
                return RSLF_ERR_NOMEM;
            }
            ctx->open_cap = pxx;
        }
        const float* epi_x = epi - (size_t)ha * S * U * C;
        RSLF_TRY(launch_edge_confidence(ctx, epi_x, Vx, S, U, C, s_first, s_count, P, ctx->open_ce, ctx->open_mask, nullptr, ctx->open_tmp, 0, 0));
        /* keep rows [ha, ha + V) of every plane */
        RSLF_CUDA_TRY(ctx, cudaMemcpy2DAsync(ce_out, (size_t)V * U * 4, ctx->open_ce + (size_t)ha * U, (size_t)Vx * U * 4, (size_t)V * U * 4, s_count,
                                             cudaMemcpyDeviceToDevice, ctx->stream));
        RSLF_CUDA_TRY(ctx, cudaMemcpy2DAsync(mask_out, (size_t)V * U, ctx->open_mask + (size_t)ha * U, (size_t)Vx * U, (size_t)V * U, s_count,
                                             cudaMemcpyDeviceToDevice, ctx->stream));
        if (rowdark) {
            const double dT = rslf_sq_threshold(P.propagation_epsilon);
            const unsigned rows = (unsigned)((size_t)V * s_count);
            if (C == 1) rowdark_count_kernel<1><<<rows, EDGE_THREADS, 0, ctx->stream>>>(epi, mask_out, V, S, U, s_first, s_count, P.propagation_epsilon, dT, rowdark, dark_lo, dark_hi);
            else rowdark_count_kernel<3><<<rows, EDGE_THREADS, 0, ctx->stream>>>(epi, mask_out, V, S, U, s_first, s_count, P.propagation_epsilon, dT, rowdark, dark_lo, dark_hi);
            RSLF_CUDA_TRY(ctx, cudaGetLastError());
            ctx->timing.kernel_launches += 1;
        }
        return RSLF_OK;
    }
    int fs = P.edge_confidence_filter_size;
    if (fs < 1 || fs > EDGE_MAX_FS) {
        snprintf(ctx->err, sizeof(ctx->err), "edge_confidence_filter_size %d outside [1,%d]", fs, EDGE_MAX_FS);
        return RSLF_ERR_UNSUPPORTED;
    }
    const bool opening = P.edge_confidence_opening_size > 1;
    morph_elem elem;
    if (opening) {
        if (!morph_build(elem, P.edge_confidence_opening_type, P.edge_confidence_opening_size)) {
            snprintf(ctx->err, sizeof(ctx->err), "edge_confidence_opening: type %d (0..2), size %d (<= %d)",
                     P.edge_confidence_opening_type, P.edge_confidence_opening_size, MORPH_MAX_K);
            return RSLF_ERR_UNSUPPORTED;
        }
        if (!mask_tmp) { snprintf(ctx->err, sizeof(ctx->err), "edge_confidence_opening: no scratch mask"); return RSLF_ERR_STATE; }
        if (V > 65535 || s_count > 65535) { snprintf(ctx->err, sizeof(ctx->err), "edge_confidence_opening: more than 65535 rows or views"); return RSLF_ERR_UNSUPPORTED; }
    }
    int* rowdark_late = opening ? rowdark : nullptr;      /* counted on the opened mask instead */
    if (opening) rowdark = nullptr;
    dim3 grid((unsigned)((size_t)V * s_count), rslf_div_up(U, EDGE_TILE));
    size_t smem = (size_t)(EDGE_TILE + fs - 1) * C * sizeof(float);
    double T = rslf_sq_threshold(P.shadow_level);
    const double dT = rslf_sq_threshold(P.propagation_epsilon);
    if (C == 1)
        edge_confidence_kernel<1><<<grid, EDGE_THREADS, smem, ctx->stream>>>(
            epi, V, S, U, s_first, s_count, fs, P.cut_shadows, P.shadow_level, T, P.edge_score_threshold, ce_out, mask_out,
            P.propagation_epsilon, dT, rowdark, 0, dark_lo, dark_hi);
    else
        edge_confidence_kernel<3><<<grid, EDGE_THREADS, smem, ctx->stream>>>(
            epi, V, S, U, s_first, s_count, fs, P.cut_shadows, P.shadow_level, T, P.edge_score_threshold, ce_out, mask_out,
            P.propagation_epsilon, dT, rowdark, 0, dark_lo, dark_hi);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    if (opening) {
        dim3 mg(rslf_div_up(U, MORPH_THREADS), V, s_count);
        morph_kernel<1><<<mg, MORPH_THREADS, 0, ctx->stream>>>(mask_out, mask_tmp, V, U, elem);
        morph_kernel<0><<<mg, MORPH_THREADS, 0, ctx->stream>>>(mask_tmp, mask_out, V, U, elem);
        ctx->timing.kernel_launches += 2;
        if (rowdark_late) {
            const unsigned rows = (unsigned)((size_t)V * s_count);
            if (C == 1) rowdark_count_kernel<1><<<rows, EDGE_THREADS, 0, ctx->stream>>>(epi, mask_out, V, S, U, s_first, s_count, P.propagation_epsilon, dT, rowdark_late, dark_lo, dark_hi);
            else rowdark_count_kernel<3><<<rows, EDGE_THREADS, 0, ctx->stream>>>(epi, mask_out, V, S, U, s_first, s_count, P.propagation_epsilon, dT, rowdark_late, dark_lo, dark_hi);
            ctx->timing.kernel_launches += 1;
        }
        RSLF_CUDA_TRY(ctx, cudaGetLastError());
    }
    return RSLF_OK;
}

/* Edge confidence of all S lines of the EPIs [v_first, v_first + v_count) on `stream` (pipelined ingest: a chunk of
 * EPIs that has just been uploaded and normalised); planes are [S][V][U].  No opening (that needs the whole mask). */
static int launch_edge_confidence_rows(rslf_ctx* ctx, cudaStream_t stream, const float* epi, int V, int S, int U, int C, int v_first,
                                       int v_count, const rslf_params& P, float* ce_out, uint8_t* mask_out, int* rowdark,
                                       int* dark_lo, int* dark_hi)
{
    const int fs = P.edge_confidence_filter_size;
    if (fs < 1 || fs > EDGE_MAX_FS) return RSLF_ERR_UNSUPPORTED;
    dim3 grid((unsigned)((size_t)v_count * S), rslf_div_up(U, EDGE_TILE));
    const size_t smem = (size_t)(EDGE_TILE + fs - 1) * C * sizeof(float);
    const double T = rslf_sq_threshold(P.shadow_level), dT = rslf_sq_threshold(P.propagation_epsilon);
    if (C == 1)
        edge_confidence_kernel<1><<<grid, EDGE_THREADS, smem, stream>>>(epi, V, S, U, 0, S, fs, P.cut_shadows, P.shadow_level, T, P.edge_score_threshold,
                                                                        ce_out, mask_out, P.propagation_epsilon, dT, rowdark, v_first, dark_lo, dark_hi);
    else
        edge_confidence_kernel<3><<<grid, EDGE_THREADS, smem, stream>>>(epi, V, S, U, 0, S, fs, P.cut_shadows, P.shadow_level, T, P.edge_score_threshold,
                                                                        ce_out, mask_out, P.propagation_epsilon, dT, rowdark, v_first, dark_lo, dark_hi);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    return RSLF_OK;
}
