/*
 * k_colour.cuh — coloured disparity maps of a fine-to-coarse result on the device.
 *
 * Replaces FineToCoarse::get_coloured_depth_maps (rslf_fine_to_coarse.hpp:324-377) with
 * ImageConverter_uchar (src/rslf_plot.cpp:66-107):
 *   fit            the values of rank floor(0.02 N) and floor(0.98 N) of the sorted fused map of view
 *                  (int)round(S / 2.0) (cv::sort of the reshaped plane, :68-79);
 *   copy_and_scale float alpha = 255.0 / (max - min); convertTo(CV_8U, alpha, -alpha * min):
 *                  cvRound(x * float(alpha) + float(beta)), saturated to 0..255 (:100-107);
 *   applyColorMap  256-entry BGR table, supplied by the caller (OpenCV's tables are data of OpenCV);
 *   black          where the validity mask is 0 and, with par_cut_shadows, where
 *                  norm(E(s, u)) < _SHADOW_NORMALIZED_LEVEL (ftc.hpp:352-369).
 *
 * The two order statistics come from a two-level radix select instead of a sort: a 65536-bin histogram of the
 * high halves of the order-preserving integer keys of the plane, then of the low halves inside the two bins
 * that hold the wanted ranks.  Both passes and the colouring are HBM-bound streams: 4 B read per pixel of the
 * plane and pass; 4 + 1 + 4C B read and 3 B written per pixel of the result.
 */
#pragma once
#include "rslf_common.cuh"

#define COLOUR_BINS 65536

/* ascending float order == ascending unsigned order of the keys (the fused maps hold no NaN) */
__device__ __forceinline__ unsigned colour_key(float x)
{
    const unsigned u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
static inline float colour_key_to_float(unsigned key)
{
    const unsigned u = (key & 0x80000000u) ? (key ^ 0x80000000u) : ~key;
    float f; memcpy(&f, &u, 4);
    return f;
}

/* one count per lane into hist[bin]; lanes of a warp that hit the same bin add once (disparity maps hold few
 * distinct values, so plain atomics would serialise on a handful of addresses) */
__device__ __forceinline__ void colour_hist_add(unsigned* __restrict__ hist, unsigned bin)
{
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, bin);
    if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(hist + bin, (unsigned)__popc(peers));
}

__global__ void __launch_bounds__(256)
colour_hist_hi_kernel(const float* __restrict__ x, size_t n, unsigned* __restrict__ hist)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        colour_hist_add(hist, colour_key(x[i]) >> 16);
}

/* low halves of the keys whose high half is hi_a (-> hist_a) / hi_b (-> hist_b); hi_a may equal hi_b */
__global__ void __launch_bounds__(256)
colour_hist_lo_kernel(const float* __restrict__ x, size_t n, unsigned hi_a, unsigned hi_b,
                      unsigned* __restrict__ hist_a, unsigned* __restrict__ hist_b)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned key = colour_key(x[i]);
        const unsigned hi = key >> 16, lo = key & 0xffffu;
        if (hi == hi_a) atomicAdd(hist_a + lo, 1u);
        if (hi == hi_b) atomicAdd(hist_b + lo, 1u);
    }
}

/* grid (U tiles, V, S): out[s][v][u][0..2] */
template <int C>
__global__ void __launch_bounds__(128)
colour_maps_kernel(const float* __restrict__ map, const uint8_t* __restrict__ valid, const float* __restrict__ epi0,
                   int V, int S, int U, float a, float b, int cut_shadows, float shadow_level, double shadow_T,
                   const uint8_t* __restrict__ lut, uint8_t* __restrict__ out)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= U) return;
    const int v = blockIdx.y, s = blockIdx.z;
    const size_t o = ((size_t)s * V + v) * (size_t)U + u;
    float x = map[o] * a;                          /* two roundings: the library is compiled with -fmad=false */
    x = x + b;
    int r = __float2int_rn(x);                     /* cvRound: half to even */
    r = min(max(r, 0), 255);
    bool black = valid[o] == 0;
    if (cut_shadows) {
        const float* e = epi0 + (((size_t)v * S + s) * (size_t)U + u) * C;
        if (C == 1) black = black || rslf_norm1_lt(__ldg(e), shadow_level);
        else black = black || rslf_norm3_lt(__ldg(e), __ldg(e + (C > 1 ? 1 : 0)), __ldg(e + (C > 2 ? 2 : 0)), shadow_T);
    }
    uint8_t* q = out + o * 3;
    const uint8_t* t = lut + 3 * r;
    q[0] = black ? 0 : t[0]; q[1] = black ? 0 : t[1]; q[2] = black ? 0 : t[2];
}

/* the bin of a histogram that holds the element of rank k (0-based), and k's rank inside that bin */
static bool colour_find_bin(const unsigned* hist, size_t k, unsigned* bin, size_t* rank_in_bin)
{
    size_t cum = 0;
    for (unsigned b = 0; b < COLOUR_BINS; ++b) {
        if (k < cum + hist[b]) { *bin = b; *rank_in_bin = k - cum; return true; }
        cum += hist[b];
    }
    return false;
}
