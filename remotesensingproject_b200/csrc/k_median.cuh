/*
 * k_median.cuh — selective median filter of the depth plane of line s_hat.
 *
 * Replaces selective_median_filter (rslf_depth_computation_core.hpp:663-718):
 * for every masked pixel, the element of rank n/2 (std::nth_element, i.e. the
 * upper median) of src(k,l) over the (2w+1)^2 window clipped at the image
 * borders, restricted to masked neighbours whose colour on line s_hat is within
 * epsilon (norm) of the centre pixel's; unmasked outputs are 0.
 *
 * One thread per pixel; the window values sit in registers (fully unrolled) and
 * are sorted by a fixed compare-exchange network (Batcher), so no data-dependent
 * indexing is needed.  Colours are addressed through (colour, row_stride), and the rows
 * to filter through (v_begin, v_count), so that a row-sharded run filters its own rows
 * out of the gathered planes of line s_hat.
 */
#pragma once
#include "rslf_common.cuh"
#include "k_balance.cuh"
#include <math_constants.h>

/* compare-exchange on registers */
__device__ __forceinline__ void med_cswap(float& a, float& b) { const float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }

/* Batcher's odd-even merge sort as a fixed compare-exchange network for any N (Knuth 5.2.2 M);
 * every index is a compile-time constant after unrolling, so the array stays in registers. */
template <int N>
__device__ __forceinline__ void sort_network(float (&a)[N])
{
#pragma unroll
    for (int p = 1; p < N; p <<= 1)
#pragma unroll
        for (int k = p; k >= 1; k >>= 1)
#pragma unroll
            for (int j = k % p; j + k < N; j += 2 * k)
#pragma unroll
                for (int i = 0; i < k; ++i)
                    if (i + j + k < N && (i + j) / (2 * p) == (i + j + k) / (2 * p)) med_cswap(a[i + j], a[i + j + k]);
}

/* rows of the neighbouring ranks (row-sharded runs): the two rows above / below this rank's block of the
 * depth, mask and colour planes of line s_hat; nullptr where the block touches the image border */
struct median_halo {
    const float* top_depth; const float* top_colour; const uint8_t* top_mask;     /* rows v = -2, -1 */
    const float* bot_depth; const float* bot_colour; const uint8_t* bot_mask;     /* rows v = V, V + 1 */
    size_t colour_halo_stride;     /* floats between the two colour rows of a halo (U * C in a receive area, the stack's row stride in place) */
    /* peer-to-peer exchange: the neighbours store the rows into this rank's arena and then raise these
     * flags to `seq`; the kernel waits for them before it touches a halo row (nullptr: NCCL path, no wait).
     * The wait is bounded by the global timer and raises *err instead of trapping (k_balance.cuh). */
    const volatile unsigned* flag_top; const volatile unsigned* flag_bot; unsigned seq; int* err;
};

/* Which painted pixels of a row that still holds dark targets can reach one (k_propagate.cuh, second kind of source)?
 * A source at column u paints column u + round(d k_s slope) in view s, with its filtered depth d somewhere in the level's
 * disparity range [dlo, dhi]; the dark targets of (row, view s) lie in columns [lo_s, hi_s].  So only sources with
 * u in the hull over the live views of [lo_s - max offset, hi_s - min offset] need a filtered value at all.
 * on == 0 (user-edited bounds may leave the range): every painted pixel of such a row is filtered. */
struct median_gate { const int* rowdark; const int* lo; const int* hi; int S, s_hat; float slope, dlo, dhi; int on; };

template <int C, int WIDTH, bool HALO>
__global__ void __launch_bounds__(128)
selective_median_kernel(const float* __restrict__ src, const uint8_t* __restrict__ mask,
                        const float* __restrict__ colour, size_t colour_row_stride,
                        int V, int U, int v_begin, float eps, double eps_T, float* __restrict__ dst,
                        const uint8_t* __restrict__ fresh, const int* __restrict__ rowdark_v, const median_halo halo,
                        const median_gate gate)
{
    constexpr int N = (2 * WIDTH + 1) * (2 * WIDTH + 1);
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y + v_begin;            /* row in the (possibly gathered) planes */
    if (HALO && (halo.flag_top || halo.flag_bot)) {
        /* blocks whose window reaches a neighbour's rows wait until those rows have landed (bounded spin) */
        const bool need_top = halo.flag_top && v < WIDTH, need_bot = halo.flag_bot && v >= V - WIDTH;
        if ((need_top || need_bot) && threadIdx.x == 0) {
            if (need_top) peer_wait32(halo.flag_top, halo.seq, halo.err);
            if (need_bot) peer_wait32(halo.flag_bot, halo.seq, halo.err);
            __threadfence_system();
        }
        __syncthreads();
    }
    __shared__ int s_ulo, s_uhi;
    bool dark_row = false;
    if (fresh) {
        dark_row = rowdark_v[blockIdx.y] > 0;
        if (dark_row && gate.on) {
            if (threadIdx.x == 0) { s_ulo = 0x7fffffff; s_uhi = -1; }
            __syncthreads();
            for (int s = threadIdx.x; s < gate.S; s += blockDim.x) {
                if (gate.rowdark[(size_t)blockIdx.y * gate.S + s] <= 0) continue;
                const float k = (float)(gate.s_hat - s);
                float t1 = gate.dlo * k; t1 = t1 * gate.slope;
                float t2 = gate.dhi * k; t2 = t2 * gate.slope;
                const int o1 = (int)roundf(fminf(t1, t2)) - 1, o2 = (int)roundf(fmaxf(t1, t2)) + 1;
                atomicMin(&s_ulo, gate.lo[(size_t)blockIdx.y * gate.S + s] - o2);
                atomicMax(&s_uhi, gate.hi[(size_t)blockIdx.y * gate.S + s] - o1);
            }
            __syncthreads();
        }
    }
    if (u >= U) return;
    const size_t o = (size_t)v * U + u;
    const size_t od = (size_t)blockIdx.y * U + u;     /* dst holds this rank's rows only */
    if (!mask[o]) { dst[od] = 0.f; return; }
    /* the filtered value is only ever read by the propagation: needed for the pixels computed in this pass
     * (fresh: still flagged in the line's remaining mask) and, in rows that still hold dark targets, for the
     * pixels painted earlier (see k_propagate.cuh).  fresh == nullptr: filter every masked pixel. */
    if (fresh && !fresh[od] && (!dark_row || (gate.on && (u < s_ulo || u > s_uhi)))) { dst[od] = 0.f; return; }
    float pc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) pc[c] = __ldg(colour + (size_t)v * colour_row_stride + (size_t)u * C + c);
    float val[N];
    int n = 0;
#pragma unroll
    for (int dk = -WIDTH; dk <= WIDTH; ++dk) {
        const int k = v + dk;
        /* the window row: inside the block, in a neighbour's halo rows, or outside the image */
        const float* srow = nullptr; const uint8_t* mrow = nullptr; const float* crow = nullptr;
        if (k >= 0 && k < V) {
            srow = src + (size_t)k * U; mrow = mask + (size_t)k * U; crow = colour + (size_t)k * colour_row_stride;
        } else if (HALO && k < 0 && k >= -2 && halo.top_depth) {
            srow = halo.top_depth + (size_t)(k + 2) * U; mrow = halo.top_mask + (size_t)(k + 2) * U;
            crow = halo.top_colour + (size_t)(k + 2) * halo.colour_halo_stride;
        } else if (HALO && k >= V && k < V + 2 && halo.bot_depth) {
            srow = halo.bot_depth + (size_t)(k - V) * U; mrow = halo.bot_mask + (size_t)(k - V) * U;
            crow = halo.bot_colour + (size_t)(k - V) * halo.colour_halo_stride;
        }
#pragma unroll
        for (int dl = -WIDTH; dl <= WIDTH; ++dl) {
            const int l = u + dl;
            float x = CUDART_INF_F;
            if (srow && l >= 0 && l < U &&
                ((HALO && (k < 0 || k >= V)) ? *reinterpret_cast<const volatile uint8_t*>(mrow + l) : mrow[l])) {
                float q[C];
#pragma unroll
                for (int c = 0; c < C; ++c) q[c] = (HALO && (k < 0 || k >= V)) ? crow[(size_t)l * C + c] : __ldg(crow + (size_t)l * C + c);
                if (rslf_norm_diff_lt<C>(pc, q, eps, eps_T)) {
                    x = (HALO && (k < 0 || k >= V)) ? *reinterpret_cast<const volatile float*>(srow + l) : srow[l];
                    ++n;
                }
            }
            val[(dk + WIDTH) * (2 * WIDTH + 1) + (dl + WIDTH)] = x;
        }
    }
    /* element of rank n/2 among the n selected values (std::nth_element at n/2, core.hpp:712-713): sort the
     * window, the rejected slots are +inf and end up last.  Depths are finite, so +inf is unambiguous. */
    sort_network<N>(val);
    const int rank = n / 2;
    float out = val[0];
#pragma unroll
    for (int i = 1; i <= N / 2; ++i) out = (rank == i) ? val[i] : out;
    dst[od] = out;
}

static int launch_selective_median(rslf_ctx* ctx, const float* src, const uint8_t* mask, const float* colour,
                                   size_t colour_row_stride, int V, int U, int C, int size, float eps, float* dst,
                                   int v_begin = 0, int v_count = -1, const uint8_t* fresh = nullptr,
                                   const int* rowdark_v = nullptr, const median_halo* halo = nullptr, const median_gate* gate = nullptr)
{
    const int width = (size - 1) / 2;
    if (v_count < 0) v_count = V;
    dim3 grid(rslf_div_up(U, 128), v_count);
    const double T = rslf_sq_threshold(eps);
    median_halo h; memset(&h, 0, sizeof(h));
    if (halo) h = *halo;
    median_gate g; memset(&g, 0, sizeof(g));
    if (gate) g = *gate;
#define RSLF_MED_CASE(CC, WW)                                                                              \
    if (C == CC && width == WW) {                                                                          \
        if (halo) selective_median_kernel<CC, WW, true><<<grid, 128, 0, ctx->stream>>>(                    \
            src, mask, colour, colour_row_stride, V, U, v_begin, eps, T, dst, fresh, rowdark_v, h, g);     \
        else selective_median_kernel<CC, WW, false><<<grid, 128, 0, ctx->stream>>>(                        \
            src, mask, colour, colour_row_stride, V, U, v_begin, eps, T, dst, fresh, rowdark_v, h, g);     \
        RSLF_CUDA_TRY(ctx, cudaGetLastError());                                                            \
        ctx->timing.kernel_launches += 1;                                                                  \
        return RSLF_OK;                                                                                    \
    }
    RSLF_MED_CASE(1, 0) RSLF_MED_CASE(1, 1) RSLF_MED_CASE(1, 2) RSLF_MED_CASE(1, 3)
    RSLF_MED_CASE(3, 0) RSLF_MED_CASE(3, 1) RSLF_MED_CASE(3, 2) RSLF_MED_CASE(3, 3)
#undef RSLF_MED_CASE
    snprintf(ctx->err, sizeof(ctx->err), "median_filter_size %d (window > 7x7) or C=%d not implemented", size, C);
    return RSLF_ERR_UNSUPPORTED;
}
