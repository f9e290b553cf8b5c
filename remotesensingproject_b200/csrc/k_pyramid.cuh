/*
 * k_pyramid.cuh — input normalisation and the fine-to-coarse pyramid kernels.
 *
 *  - stack_minmax / normalise: computers' ctor input scaling
 *    (rslf_depth_computation.hpp:442-477, 668-704)
 *  - downsample: rslf::downsample_EPIs (src/rslf_fine_to_coarse_core.cpp:14-60):
 *    GaussianBlur 7x7 sigma 0 BORDER_REFLECT + resize fx=fy=0.5 (2x2 area mean)
 *  - nearest_valid / set_bounds: FineToCoarse::run's bound propagation
 *    (rslf_fine_to_coarse.hpp:201-294)
 *  - valid_mask: Depth2DComputer::get_valid_depths_mask_s_v_u (dc.hpp:893-915)
 *  - fuse_level / median3x3: rslf::fuse_disp_maps (ftc_core.cpp:69-135)
 * All are HBM-bound streams; arithmetic order follows the float32 paths of the
 * OpenCV calls (checked against cv2 through the oracle's golden vectors).
 */
#pragma once
#include "rslf_common.cuh"
#include <math_constants.h>

/* ---- min / max of a float stack (values are compared, order-independent) ---- */
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    /* ordered-int trick, valid for any mix of signs */
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

/* mm[0] = max (initialised by the caller to the reference's start value), mm[1] = min */
__global__ void stack_minmax_kernel(const float* __restrict__ x, size_t n, float* __restrict__ mm)
{
    float mx = -CUDART_INF_F, mn = CUDART_INF_F;
    /* 16-byte loads over the aligned body, scalar head / tail */
    const size_t head = min(n, (size_t)((16 - ((uintptr_t)x & 15)) & 15) / 4), n4 = (n - head) / 4;
    const float4* x4 = reinterpret_cast<const float4*>(x + head);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = x4[i];
        mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w))); mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n - 4 * n4; i += (size_t)gridDim.x * blockDim.x) {
        const float v = (i < head) ? x[i] : x[4 * n4 + i];
        mx = fmaxf(mx, v); mn = fminf(mn, v);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, off));
    }
    if ((threadIdx.x & 31) == 0) { atomic_max_float(mm, mx); atomic_min_float(mm + 1, mn); }
}

/* out = in * a with a = float(1.0 / scale); scale = *scale_dev if given (the stack maximum) */
__global__ void normalise_f32_kernel(const float* __restrict__ in, size_t n, const float* __restrict__ scale_dev,
                                     float scale_host, float* __restrict__ out)
{
    const float sf = scale_dev ? *scale_dev : scale_host;
    const float a = (float)(1.0 / (double)sf);
    if ((((uintptr_t)in | (uintptr_t)out) & 15) == 0) {
        /* HBM-bound stream: 16-byte loads and stores */
        const size_t n4 = n / 4;
        const float4* in4 = reinterpret_cast<const float4*>(in);
        float4* out4 = reinterpret_cast<float4*>(out);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const float4 v = in4[i];
            out4[i] = make_float4(v.x * a, v.y * a, v.z * a, v.w * a);
        }
        for (size_t i = 4 * n4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i] * a;
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = in[i] * a;
}
__global__ void normalise_u8_kernel(const uint8_t* __restrict__ in, size_t n, float* __restrict__ out)
{
    const float a = (float)(1.0 / 255.0);
    if ((((uintptr_t)in & 3) | ((uintptr_t)out & 15)) == 0) {
        const size_t n4 = n / 4;
        const uchar4* in4 = reinterpret_cast<const uchar4*>(in);
        float4* out4 = reinterpret_cast<float4*>(out);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
            const uchar4 v = in4[i];
            out4[i] = make_float4((float)v.x * a, (float)v.y * a, (float)v.z * a, (float)v.w * a);
        }
        for (size_t i = 4 * n4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (float)in[i] * a;
        return;
    }
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i] * a;
}
/* CV_16U stacks: maximum for scale < 0 (minMaxLoc -> (float)max, dc.hpp:453-456; mm[0] holds the start value) and
 * convertTo(CV_32F, 1.0 / scale) = float(x) * float(1.0 / scale) (dc.hpp:472-474) */
__global__ void stack_max_u16_kernel(const uint16_t* __restrict__ x, size_t n, float* __restrict__ mm)
{
    unsigned mx = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        mx = max(mx, (unsigned)x[i]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) atomic_max_float(mm, (float)mx);
}
__global__ void normalise_u16_kernel(const uint16_t* __restrict__ in, size_t n, float scale, float* __restrict__ out)
{
    const float a = (float)(1.0 / (double)scale);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (float)in[i] * a;
}

/* ---- downsample_EPIs, float32 ------------------------------------------------
 * One block = one view s, one tile of DS_TV x DS_TU output pixels.  The input tile
 * (with the 3-pixel blur halo, border rule applied while staging) goes to shared
 * memory, the horizontal pass writes a second shared buffer, the vertical pass and
 * the 2x2 mean produce the output.  Kernel [1,3.5,7,9,7,3.5,1]/32, symmetric form
 * k0*x0 + k1*(x-1 + x+1) + k2*(x-2 + x+2) + k3*(x-3 + x+3), rows then columns.
 */
#define DS_TU 32
#define DS_TV 8
#define DS_THREADS 256

template <int C>
__global__ void __launch_bounds__(DS_THREADS)
downsample_kernel(const float* __restrict__ in, int V, int S, int U, float* __restrict__ out, int V2, int U2,
                  int ov_begin, int ov_count)
{
    constexpr int IW = 2 * DS_TU + 6, IH = 2 * DS_TV + 6, BW = 2 * DS_TU;
    __shared__ float tin[IH][IW * C];
    __shared__ float hb[IH][BW * C];
    const float k0 = 9.f / 32.f, k1 = 7.f / 32.f, k2 = 3.5f / 32.f, k3 = 1.f / 32.f;
    const int s = blockIdx.z;
    /* `in` holds all V rows of the level; this launch writes the ov_count output rows from ov_begin on
     * into `out` (local row 0 = output row ov_begin), so a row-sharded rank produces only its rows */
    const int ou0 = blockIdx.x * DS_TU, ov0 = ov_begin + blockIdx.y * DS_TV;
    const int iu0 = 2 * ou0 - 3, iv0 = 2 * ov0 - 3;
    for (int i = threadIdx.x; i < IH * IW * C; i += DS_THREADS) {
        int r = i / (IW * C), rem = i - r * (IW * C);
        int p = rem / C, c = rem - p * C;
        /* blurred rows / columns beyond the image are only reached through the clamped
         * 2x2 source index, i.e. never: clamp first, then apply BORDER_REFLECT */
        int vv = rslf_reflect(min(iv0 + r, V - 1 + 3), V);
        int uu = rslf_reflect(min(iu0 + p, U - 1 + 3), U);
        tin[r][rem] = __ldg(in + (((size_t)vv * S + s) * (size_t)U + uu) * C + c);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IH * BW * C; i += DS_THREADS) {
        int r = i / (BW * C), rem = i - r * (BW * C);
        int p = rem / C, c = rem - p * C;
        const float* x = &tin[r][(p + 3) * C + c];
        float acc = k0 * x[0];
        float t1 = x[-1 * C] + x[1 * C]; t1 = k1 * t1; acc = acc + t1;
        float t2 = x[-2 * C] + x[2 * C]; t2 = k2 * t2; acc = acc + t2;
        float t3 = x[-3 * C] + x[3 * C]; t3 = k3 * t3; acc = acc + t3;
        hb[r][rem] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DS_TV * DS_TU * C; i += DS_THREADS) {
        int r = i / (DS_TU * C), rem = i - r * (DS_TU * C);
        int p = rem / C, c = rem - p * C;
        const int ov = ov0 + r, ou = ou0 + p;
        if (ov >= V2 || ov >= ov_begin + ov_count || ou >= U2) continue;
        /* source rows / columns of the 2x2 mean, clamped at odd edges */
        const int r0 = min(2 * ov, V - 1) - iv0, r1 = min(2 * ov + 1, V - 1) - iv0;
        const int c0 = min(2 * ou, U - 1) - 2 * ou0, c1 = min(2 * ou + 1, U - 1) - 2 * ou0;
        float bl[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int rr = a ? r1 : r0;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int cc = (b ? c1 : c0) * C + c;
                float acc = k0 * hb[rr][cc];
                float t1 = hb[rr - 1][cc] + hb[rr + 1][cc]; t1 = k1 * t1; acc = acc + t1;
                float t2 = hb[rr - 2][cc] + hb[rr + 2][cc]; t2 = k2 * t2; acc = acc + t2;
                float t3 = hb[rr - 3][cc] + hb[rr + 3][cc]; t3 = k3 * t3; acc = acc + t3;
                bl[a][b] = acc;
            }
        }
        const float sum = ((bl[0][0] + bl[0][1]) + bl[1][0]) + bl[1][1];
        out[(((size_t)(ov - ov_begin) * S + s) * (size_t)U2 + ou) * C + c] = sum * 0.25f;
    }
}

/* ---- downsample_EPIs, 8-bit and 16-bit stacks (they keep their depth between levels, ftc.hpp:142-147) ----
 * OpenCV's integer paths, bit-exact against cv2 (tests/golden/down_u8_*.npz, tests/test_oracle_vs_cv2.py): integer
 * kernel [8,28,56,72,56,28,8] (sum 256) along rows then columns, (acc + 2^15) >> 16, BORDER_REFLECT; half-size
 * resize = (a+b+c+d+2) >> 2, or the half-to-even rounded mean where an odd dimension leaves a single
 * source row / column.  Same tiling as the float kernel.  T = uint8_t / uint16_t; unsigned 32-bit
 * accumulators (256 * 256 * 65535 + 2^15 < 2^32).
 */
template <typename T, int C>
__global__ void __launch_bounds__(DS_THREADS)
downsample_int_kernel(const T* __restrict__ in, int V, int S, int U, T* __restrict__ out, int V2, int U2,
                      int ov_begin, int ov_count)
{
    constexpr int IW = 2 * DS_TU + 6, IH = 2 * DS_TV + 6, BW = 2 * DS_TU;
    __shared__ T tin[IH][IW * C];
    __shared__ unsigned hb[IH][BW * C];
    const unsigned K[7] = {8u, 28u, 56u, 72u, 56u, 28u, 8u};
    const int s = blockIdx.z;
    const int ou0 = blockIdx.x * DS_TU, ov0 = ov_begin + blockIdx.y * DS_TV;
    const int iu0 = 2 * ou0 - 3, iv0 = 2 * ov0 - 3;
    for (int i = threadIdx.x; i < IH * IW * C; i += DS_THREADS) {
        int r = i / (IW * C), rem = i - r * (IW * C);
        int p = rem / C, c = rem - p * C;
        int vv = rslf_reflect(min(iv0 + r, V - 1 + 3), V);
        int uu = rslf_reflect(min(iu0 + p, U - 1 + 3), U);
        tin[r][rem] = in[(((size_t)vv * S + s) * (size_t)U + uu) * C + c];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < IH * BW * C; i += DS_THREADS) {
        int r = i / (BW * C), rem = i - r * (BW * C);
        int p = rem / C, c = rem - p * C;
        unsigned acc = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j) acc += K[j] * (unsigned)tin[r][(p + j) * C + c];
        hb[r][rem] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DS_TV * DS_TU * C; i += DS_THREADS) {
        int r = i / (DS_TU * C), rem = i - r * (DS_TU * C);
        int p = rem / C, c = rem - p * C;
        const int ov = ov0 + r, ou = ou0 + p;
        if (ov >= V2 || ov >= ov_begin + ov_count || ou >= U2) continue;
        const int nr = (2 * ov + 1 < V) ? 2 : 1, nc = (2 * ou + 1 < U) ? 2 : 1;
        unsigned sum = 0;
        for (int a = 0; a < nr; ++a) {
            const int rr = 2 * ov + a - iv0;
            for (int b = 0; b < nc; ++b) {
                const int cc = (2 * ou + b - 2 * ou0) * C + c;
                unsigned acc = 0;
#pragma unroll
                for (int j = 0; j < 7; ++j) acc += K[j] * hb[rr + j - 3][cc];
                sum += (acc + 32768u) >> 16;
            }
        }
        const int n = nr * nc;
        unsigned res = sum;
        if (n == 4) res = (sum + 2u) >> 2;
        else if (n == 2) res = (sum + ((sum >> 1) & 1u)) >> 1;          /* half to even */
        out[(((size_t)(ov - ov_begin) * S + s) * (size_t)U2 + ou) * C + c] = (T)res;
    }
}

/* ---- validity mask (Depth2DComputer::get_valid_depths_mask_s_v_u, dc.hpp:893-915): conf > thr with conf = C_e (default
 * build, :905-906) or C_d (disparity-confidence criterion, :901-902); an accept-all level takes C_e > -1 ---- */
__global__ void valid_mask_kernel(const float* __restrict__ ce, const float* __restrict__ conf, size_t n, float thr, int accept_all,
                                  uint8_t* __restrict__ valid)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        valid[i] = accept_all ? ((ce[i] > -1.f) ? 255 : 0) : ((conf[i] > thr) ? 255 : 0);
}

__global__ void fill_f32_kernel(float* __restrict__ x, size_t n, float v)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = v;
}

/* ---- nearest valid neighbours of every pixel of a row (one warp per row) ----
 * left[u]  = largest index j <= u with valid[j] and j >= 1   (-1 if none)
 * right[u] = smallest index j >= u with valid[j]              (U if none)
 * The reference scans u_up-1 .. 1 to the left (index 0 is never tested,
 * ftc.hpp:222-236) and u_up+1 .. U-1 to the right (:238-252).
 */
__global__ void nearest_valid_kernel(const uint8_t* __restrict__ valid, int rows, int U,
                                     int* __restrict__ left, int* __restrict__ right)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const uint8_t* vr = valid + (size_t)warp * U;
    int* lr = left + (size_t)warp * U;
    int* rr = right + (size_t)warp * U;
    int carry = -1;
    for (int base = 0; base < U; base += 32) {
        const int u = base + lane;
        const bool ok = (u < U) && (u >= 1) && (vr[u] > 0);
        const unsigned bits = __ballot_sync(0xffffffffu, ok);
        const unsigned upto = bits & (0xffffffffu >> (31 - lane));
        if (u < U) lr[u] = upto ? (base + 31 - __clz(upto)) : carry;
        if (bits) carry = base + 31 - __clz(bits);
    }
    carry = U;
    for (int base = ((U - 1) / 32) * 32; base >= 0; base -= 32) {
        const int u = base + lane;
        const bool ok = (u < U) && (vr[u] > 0);
        const unsigned bits = __ballot_sync(0xffffffffu, ok);
        const unsigned from = bits & (0xffffffffu << lane);
        if (u < U) rr[u] = from ? (base + __ffs(from) - 1) : carry;
        if (bits) carry = base + __ffs(bits) - 1;
    }
}

/* bounds of level p+1 (down) from depth / nearest-valid tables of level p (up) */
__global__ void set_bounds_kernel(const float* __restrict__ depth_up, const int* __restrict__ left,
                                  const int* __restrict__ right, int S, int Vu, int Uu, int Vd, int Ud,
                                  float* __restrict__ dmin_map, float* __restrict__ dmax_map,
                                  int Vu_tot, int vu0, int vd0, int write_default, float dmin_default, float dmax_default)
{
    /* write_default: pixels without a bound from the level above take the global bounds here (ftc.hpp:160-168 fills the
     * maps first); otherwise the maps hold them already (rslf_cuda_set_bounds) */
    /* Vu / Vd: rows held locally (first global rows vu0 / vd0); Vu_tot: global rows of the upper level */
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y, s = blockIdx.z;
    if (u >= Ud) return;
    const int v_up_g = min(2 * (v + vd0), Vu_tot - 1);
    const int v_up = v_up_g - vu0;
    const int u_up = min(2 * u, Uu - 1);
    float mn = 0.f, mx = 0.f; int n = 0;
    for (int line = 0; line < 2; ++line) {
        const int vv = v_up + line;
        if (line == 1 && !(v_up_g + 1 < Vu_tot)) break;
        if (vv < 0 || vv >= Vu) continue;          /* cannot happen with aligned shards */
        const size_t ro = ((size_t)s * Vu + vv) * (size_t)Uu;
        const int jl = (u_up - 1 >= 1) ? left[ro + u_up - 1] : -1;
        const int jr = (u_up + 1 <= Uu - 1) ? right[ro + u_up + 1] : Uu;
        if (jl >= 1 && jr <= Uu - 1) {
            const float dl = depth_up[ro + jl], dr = depth_up[ro + jr];
            if (dl == dl && dr == dr) {                 /* the reference's !is_nan test */
                if (n == 0) { mn = dl; mx = dl; }
                mn = fminf(mn, dl); mx = fmaxf(mx, dl);
                mn = fminf(mn, dr); mx = fmaxf(mx, dr);
                n += 2;
            }
        }
    }
    const size_t o = ((size_t)s * Vd + v) * (size_t)Ud + u;
    if (n > 1) { dmin_map[o] = mn; dmax_map[o] = mx; }
    else if (write_default) { dmin_map[o] = dmin_default; dmax_map[o] = dmax_default; }
}

/* A [S][rows][U] map seen through a window of rows: the rank's own rows [r0, r0 + rows) in `mid` (plane stride
 * mid_rows * U) plus H halo rows above / below them in `top` / `bot` (planes of H rows each: the neighbours' rows,
 * pushed into this rank's arena).  A whole map is the window r0 = 0, rows = all, no halo. */
template <typename T> struct svu_view { const T* mid; const T* top; const T* bot; int r0, rows, H; };
template <typename T>
__device__ __forceinline__ const T* svu_row(const svu_view<T>& w, int s, int r, int U)
{
    if (r < w.r0) return w.top + ((size_t)s * w.H + (r - (w.r0 - w.H))) * (size_t)U;
    if (r >= w.r0 + w.rows) return w.bot + ((size_t)s * w.H + (r - (w.r0 + w.rows))) * (size_t)U;
    return w.mid + ((size_t)s * w.rows + (r - w.r0)) * (size_t)U;
}
template <typename T> static svu_view<T> svu_whole(const T* p, int rows) { svu_view<T> w; w.mid = p; w.top = w.bot = nullptr; w.r0 = 0; w.rows = rows; w.H = 0; return w; }

/* source row of the bilinear / nearest upsampling of fuse_disp_maps for destination row y (host and device agree) */
static __host__ __device__ inline void fuse_src_rows(int y, int Vs, int Vd, int* iy0, int* iy1, float* fyo, int* ny)
{
    const double sy = (double)Vs / Vd;
    float fy = (float)((y + 0.5) * sy - 0.5);
    int iy = (int)floorf(fy); fy -= iy;
    if (iy < 0) { iy = 0; fy = 0.f; }
    if (iy >= Vs - 1) { iy = Vs - 1; fy = 0.f; }
    *iy0 = iy; *iy1 = (iy + 1 < Vs - 1) ? iy + 1 : Vs - 1; *fyo = fy;
    const int n = (int)floor(y * sy);
    *ny = n < Vs - 1 ? n : Vs - 1;
}

/* ---- fuse one level: map = valid ? disp : bilinear(map_down); mask = valid | nearest(mask_down) ----
 * cv::resize INTER_LINEAR float32 (half-pixel centres, edge clamp, horizontal then vertical
 * interpolation) and INTER_NEAREST (src = min(floor(dst * scale), n - 1)).
 */
__global__ void fuse_level_kernel(const svu_view<float> map_down, const svu_view<uint8_t> mask_down,
                                  int Vs, int Us, const float* __restrict__ disp, const uint8_t* __restrict__ valid,
                                  int Vd, int Ud, float* __restrict__ map_out, uint8_t* __restrict__ mask_out,
                                  int y0, int Vd_loc)
{
    /* map_down / mask_down: windows of the coarser level (Vs x Us); disp / valid / outputs: the Vd_loc local rows
     * from global row y0 on (Vd = global rows of the finer level) */
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int yl = blockIdx.y, s = blockIdx.z;
    const int y = yl + y0;
    if (x >= Ud) return;
    const size_t o = ((size_t)s * Vd_loc + yl) * (size_t)Ud + x;
    const double sx = (double)Us / Ud;
    int iy, y1, ny; float fy;
    fuse_src_rows(y, Vs, Vd, &iy, &y1, &fy, &ny);
    const int nx = min((int)floor(x * sx), Us - 1);
    const uint8_t mup = svu_row(mask_down, s, ny, Us)[nx];
    const uint8_t vl = valid[o];
    mask_out[o] = vl | mup;
    if (vl) { map_out[o] = disp[o]; return; }
    float fx = (float)((x + 0.5) * sx - 0.5);
    int ix = (int)floorf(fx); fx -= ix;
    if (ix < 0) { ix = 0; fx = 0.f; }
    if (ix >= Us - 1) { ix = Us - 1; fx = 0.f; }
    const float* r0 = svu_row(map_down, s, iy, Us);
    const float* r1 = svu_row(map_down, s, y1, Us);
    const int x1 = min(ix + 1, Us - 1);
    const float a1 = fx, a0 = 1.f - a1, b1 = fy, b0 = 1.f - b1;
    float h0 = r0[ix] * a0; { float t = r0[x1] * a1; h0 = h0 + t; }
    float h1 = r1[ix] * a0; { float t = r1[x1] * a1; h1 = h1 + t; }
    float r = h0 * b0; { float t = h1 * b1; r = r + t; }
    map_out[o] = 0.0f + r;
}

__device__ __forceinline__ void sort2(float& a, float& b) { float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }

/* cv::medianBlur(float32, 3): 3x3 median, BORDER_REPLICATE */
__global__ void median3x3_kernel(const svu_view<float> src, int V, int U, float* __restrict__ dst, int v0, int V_loc)
{
    /* src: window of the V-row map; dst: the V_loc local rows from global row v0 on */
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y + v0, s = blockIdx.z;
    if (u >= U) return;
    float w[9]; int n = 0;
#pragma unroll
    for (int dv = -1; dv <= 1; ++dv) {
        const int vv = min(max(v + dv, 0), V - 1);
        const float* p = svu_row(src, s, vv, U);
#pragma unroll
        for (int du = -1; du <= 1; ++du) {
            const int uu = min(max(u + du, 0), U - 1);
            w[n++] = p[uu];
        }
    }
    /* median-of-9 exchange network */
    sort2(w[1], w[2]); sort2(w[4], w[5]); sort2(w[7], w[8]);
    sort2(w[0], w[1]); sort2(w[3], w[4]); sort2(w[6], w[7]);
    sort2(w[1], w[2]); sort2(w[4], w[5]); sort2(w[7], w[8]);
    sort2(w[0], w[3]); sort2(w[5], w[8]); sort2(w[4], w[7]);
    sort2(w[3], w[6]); sort2(w[1], w[4]); sort2(w[2], w[5]);
    sort2(w[4], w[7]); sort2(w[4], w[2]); sort2(w[6], w[4]);
    sort2(w[4], w[2]);
    dst[((size_t)s * V_loc + (v - v0)) * (size_t)U + u] = w[4];
}
