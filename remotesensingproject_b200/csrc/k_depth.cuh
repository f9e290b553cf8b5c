/*
 * k_depth.cuh — the hot kernel: radiance sampling along EPI lines, mean shift,
 * score S(u,d), argmax and the disparity confidence, for the pixels of one line
 * s_hat that are still to be computed.
 *
 * Replaces compute_1D_depth_epi (rslf_depth_computation_core.hpp:480-661) with
 * Interpolation1DLinear::interpolate_mat (rslf_interpolation.hpp:155-193),
 * BandwidthKernel::evaluate_mat (src/rslf_kernels.cpp:16-26, 39-54) and the
 * multi-channel multiply / divide helpers (src/rslf_depth_computation_core.cpp:25-51),
 * as driven by compute_1D_depth_epi_pile (core.hpp:772-875).
 *
 * Work decomposition
 *   - compact_kernel ANDs the line's "remaining" mask with its edge mask in
 *     place (core.hpp:510-513) and appends the surviving pixels to a list
 *     (the reference's findNonZero, :516).  No host round trip: the depth kernel
 *     reads the list length from device memory.
 *   - depth_kernel is persistent: one warp per block, blocks stride over
 *     "warp items" = (pixel, chunk of 32*H hypotheses).  A lane owns H
 *     consecutive hypotheses.  Sums over views s run sequentially in ascending
 *     s inside a lane, exactly like cv::reduce over rows, so scores are
 *     bit-identical to the CPU path; no cross-lane arithmetic touches a score.
 *   - the warp's radiances r[s][c][d] live in shared memory ([s][c][32*H]
 *     floats, one LDS.32/64/128 per lane, conflict-free); out-of-image samples
 *     hold a large finite sentinel instead of NaN: K = max(1 - |x/h|^2, 0) is 0
 *     for it, as for NaN in the reference, and r*K contributes +0.
 *   - pixels with more hypotheses than one warp item holds are merged by the
 *     last-arriving warp (arrival counter per pixel), first maximum winning
 *     like cv::minMaxLoc.
 *
 * Roofline: FP32 issue bound.  Per (pixel, hypothesis, view) sample and mean-
 * shift iteration the reference's separately rounded operations are, for RGB,
 * 3 subtractions, 6 + 3 multiplications, 7 additions and 1 max = 19 FP32
 * pipe instructions + 1 FMNMX (9 + 1 for gray); none may be fused (parity).
 */
#pragma once
#include "rslf_common.cuh"

#define RSLF_RAD_SENTINEL 1.0e18f

struct __align__(16) rslf_partial {
    float mx; int idx; float dv; float rb[3]; double sum;
};

struct depth_args {
    const float* epi; int V, S, U, D; int s_hat; float slope; float inv; int iters;
    const int* items; const int* count;
    const float* dmin_map; const float* dmax_map; float dmin_c, dmax_c;
    float* ce; uint8_t* emask; float* cd; float* depth; float* rbar;   /* planes of line s_hat */
    float raw_thr;
    int chunks; rslf_partial* partials; int* arrive;
};

/* remaining &= emask (in place) and list the survivors; remaining == nullptr: list emask. */
__global__ void compact_kernel(const uint8_t* __restrict__ emask, uint8_t* __restrict__ remaining, int n,
                               int* __restrict__ items, int* __restrict__ count,
                               unsigned long long* __restrict__ total)
{
    const int lane = threadIdx.x & 31;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
        int i = base + lane;
        uint8_t m = 0;
        if (i < n) {
            m = emask[i];
            if (remaining) { m &= remaining[i]; remaining[i] = m; }
        }
        unsigned bal = __ballot_sync(0xffffffffu, m != 0);
        if (bal) {
            int pos = 0;
            if (lane == 0) {
                int c = __popc(bal);
                pos = atomicAdd(count, c);
                atomicAdd(total, (unsigned long long)c);
            }
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (m) items[pos + __popc(bal & ((1u << lane) - 1u))] = i;
        }
    }
}

template <int H> struct rad_vec;
template <> struct rad_vec<1> { typedef float type; };
template <> struct rad_vec<2> { typedef float2 type; };
template <> struct rad_vec<4> { typedef float4 type; };

template <int H> __device__ __forceinline__ void rad_store(float* p, const float (&v)[H]);
template <> __device__ __forceinline__ void rad_store<1>(float* p, const float (&v)[1]) { *p = v[0]; }
template <> __device__ __forceinline__ void rad_store<2>(float* p, const float (&v)[2]) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}
template <> __device__ __forceinline__ void rad_store<4>(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <int H> __device__ __forceinline__ void rad_load(const float* p, float (&v)[H]);
template <> __device__ __forceinline__ void rad_load<1>(const float* p, float (&v)[1]) { v[0] = *p; }
template <> __device__ __forceinline__ void rad_load<2>(const float* p, float (&v)[2]) {
    float2 t = *reinterpret_cast<const float2*>(p); v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void rad_load<4>(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}

#define DEPTH_UNR 4      /* views per unrolled block; the radiance rows are padded to a multiple of it */

static __host__ __device__ inline int depth_padded_views(int S) { return (S + DEPTH_UNR - 1) / DEPTH_UNR * DEPTH_UNR; }

template <int C, int H, bool NONNEG>
__global__ void __launch_bounds__(32)
depth_kernel(const depth_args a)
{
    extern __shared__ float4 rad_raw[];
    float* rad = reinterpret_cast<float*>(rad_raw);         /* [Spad][C][32*H]; rows >= S hold the sentinel */
    const int lane = threadIdx.x;
    const int S = a.S, U = a.U, D = a.D;
    const int Spad = depth_padded_views(S);
    constexpr int W = 32 * H;
    const long long total = (long long)(*a.count) * a.chunks;
    const float inv = a.inv;

    for (long long w = blockIdx.x; w < total; w += gridDim.x) {
        const int item = (int)(w / a.chunks);
        const int chunk = (int)(w - (long long)item * a.chunks);
        const int pix = a.items[item];
        const int v = pix / U, u = pix - v * U;
        const float dmin = a.dmin_map ? a.dmin_map[pix] : a.dmin_c;
        const float dmax = a.dmax_map ? a.dmax_map[pix] : a.dmax_c;
        const float* epi = a.epi + (size_t)v * S * (size_t)U * C;
        const int dbase = chunk * W + lane * H;

        /* D[d] = dmin + d * (dmax - dmin) / (dim_d - 1)   (core.hpp:547-548) */
        float Dv[H];
        {
            const float range = dmax - dmin;
            const float den = (float)(D - 1);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float t = (float)(dbase + h) * range;
                t = t / den;
                Dv[h] = dmin + t;
            }
        }
        /* ---- radiances: I = (s_hat - s) * D * slope + u, linear interpolation (core.hpp:550-552,
         *      interp.hpp:155-193); card_R = number of in-image views per hypothesis.
         *      DEPTH_UNR views per step, all gathers of a step issued before the first use;
         *      the loads use clamped indices so that they need no branch. ---- */
        float card[H];
#pragma unroll
        for (int h = 0; h < H; ++h) card[h] = 0.f;
        const float uf = (float)u;
        __syncwarp();
        for (int sb = 0; sb < Spad; sb += DEPTH_UNR) {
            float e0[DEPTH_UNR][H][C], e1[DEPTH_UNR][H][C], tt[DEPTH_UNR][H];
            bool ok[DEPTH_UNR][H];
#pragma unroll
            for (int j = 0; j < DEPTH_UNR; ++j) {
                const int s = sb + j;
                const float k = (float)(a.s_hat - s);
                const float* row = epi + (size_t)min(s, S - 1) * U * C;
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    float I = k * Dv[h];
                    I = I * a.slope;
                    I = I + uf;
                    const float fl = floorf(I);
                    const int i0 = (int)fl;
                    const int i1 = i0 + ((I != fl) ? 1 : 0);            /* ceil */
                    ok[j][h] = !(i0 < 0 || i1 > U - 1) && (I == I) && (s < S);
                    tt[j][h] = I - (float)i0;
                    const int c0 = min(max(i0, 0), U - 1), c1 = min(max(i1, 0), U - 1);
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        e0[j][h][c] = __ldg(row + (size_t)c0 * C + c);
                        e1[j][h][c] = __ldg(row + (size_t)c1 * C + c);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < DEPTH_UNR; ++j) {
                float val[C][H];
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const float t = tt[j][h];
                    const float omt = 1.f - t;
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const float p = omt * e0[j][h][c];
                        const float q = t * e1[j][h][c];
                        val[c][h] = ok[j][h] ? (p + q) : RSLF_RAD_SENTINEL;
                    }
                    card[h] += ok[j][h] ? 1.0f : 0.0f;
                }
#pragma unroll
                for (int c = 0; c < C; ++c) rad_store<H>(rad + ((size_t)(sb + j) * C + c) * W + lane * H, val[c]);
            }
        }
        __syncwarp();
        /* ---- mean shift (core.hpp:577-610): r_bar <- row s_hat, then iterate.  The loop runs over
         *      the padded rows in blocks of DEPTH_UNR with the next block's LDS issued before the
         *      current block's arithmetic; sentinel rows add K = 0 and r*K = +0, which leaves every
         *      partial sum bit-identical. ---- */
        float rb[C][H];
#pragma unroll
        for (int c = 0; c < C; ++c) rad_load<H>(rad + ((size_t)a.s_hat * C + c) * W + lane * H, rb[c]);
        float sK[H];
#pragma unroll
        for (int h = 0; h < H; ++h) sK[h] = 0.f;
        const float* rp = rad + lane * H;
        for (int it = 0; it < a.iters; ++it) {
            float sR[C][H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                sK[h] = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) sR[c][h] = 0.f;
            }
            float nx[DEPTH_UNR][C][H];
#pragma unroll
            for (int j = 0; j < DEPTH_UNR; ++j)
#pragma unroll
                for (int c = 0; c < C; ++c) rad_load<H>(rp + ((size_t)j * C + c) * W, nx[j][c]);
            for (int sb = 0; sb < Spad; sb += DEPTH_UNR) {
                float r[DEPTH_UNR][C][H];
#pragma unroll
                for (int j = 0; j < DEPTH_UNR; ++j)
#pragma unroll
                    for (int c = 0; c < C; ++c)
#pragma unroll
                        for (int h = 0; h < H; ++h) r[j][c][h] = nx[j][c][h];
                const int sn = min(sb + DEPTH_UNR, Spad - DEPTH_UNR);       /* last block: harmless re-read */
#pragma unroll
                for (int j = 0; j < DEPTH_UNR; ++j)
#pragma unroll
                    for (int c = 0; c < C; ++c) rad_load<H>(rp + ((size_t)(sn + j) * C + c) * W, nx[j][c]);
#pragma unroll
                for (int j = 0; j < DEPTH_UNR; ++j) {
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        float b;
                        if (C == 1) {
                            const float x = r[j][0][h] - rb[0][h];              /* core.hpp:591 */
                            b = (inv * x) * x;                                  /* kern.cpp:21: multiply(src, src, scale) */
                        } else {
                            const float x0 = r[j][0][h] - rb[0][h];
                            const float x1 = r[j][C > 1 ? 1 : 0][h] - rb[C > 1 ? 1 : 0][h];
                            const float x2 = r[j][C > 2 ? 2 : 0][h] - rb[C > 2 ? 2 : 0][h];
                            const float b0 = (inv * x0) * x0;                   /* kern.cpp:43 */
                            const float b1 = (inv * x1) * x1;
                            const float b2 = (inv * x2) * x2;
                            b = (b0 + b1) + b2;                                 /* kern.cpp:47-49 */
                        }
                        const float kk = fmaxf(1.0f - b, 0.f);                  /* kern.cpp:23-25 / 51-53 */
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const float r0 = NONNEG ? r[j][c][h] : fmaxf(r[j][c][h], 0.f);   /* core.hpp:580 */
                            const float p = r0 * kk;                            /* core.cpp:28 / 33-37 */
                            sR[c][h] = sR[c][h] + p;                            /* core.hpp:602 */
                        }
                        sK[h] = sK[h] + kk;                                     /* core.hpp:603 */
                    }
                }
            }
            /* r_bar = sum_rK / sum_K (x / 0 = 0 as in OpenCV 3), then max(., 0) (core.hpp:606-609) */
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float den = sK[h];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float q = (den != 0.f) ? (sR[c][h] / den) : 0.f;
                    rb[c][h] = (q > 0.f) ? q : 0.f;
                }
            }
        }
        /* ---- score = sum_K(last iteration) / card_R, max(., 0) (core.hpp:616-622);
         *      lane-local first maximum, then warp argmax with the lowest index on ties ---- */
        float best = -1.f; int bidx = 0x7fffffff; double sum = 0.0;
        float bdv = 0.f, brb[C];
#pragma unroll
        for (int c = 0; c < C; ++c) brb[c] = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            if (dbase + h < D) {
                const float q = sK[h] / card[h];
                const float sc = (q > 0.f) ? q : 0.f;
                sum += (double)sc;
                if (sc > best) {
                    best = sc; bidx = dbase + h; bdv = Dv[h];
#pragma unroll
                    for (int c = 0; c < C; ++c) brb[c] = rb[c][h];
                }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, off);
            const float odv = __shfl_xor_sync(0xffffffffu, bdv, off);
            float orb[C];
#pragma unroll
            for (int c = 0; c < C; ++c) orb[c] = __shfl_xor_sync(0xffffffffu, brb[c], off);
            sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (ob > best || (ob == best && oi < bidx)) {
                best = ob; bidx = oi; bdv = odv;
#pragma unroll
                for (int c = 0; c < C; ++c) brb[c] = orb[c];
            }
        }
        /* ---- merge the chunks of this pixel (last arriver), then the outputs (core.hpp:630-657) ---- */
        bool finalise = true;
        if (a.chunks > 1) {
            if (lane == 0) {
                rslf_partial p;
                p.mx = best; p.idx = bidx; p.dv = bdv; p.sum = sum;
#pragma unroll
                for (int c = 0; c < 3; ++c) p.rb[c] = (c < C) ? brb[c < C ? c : 0] : 0.f;
                a.partials[(size_t)item * a.chunks + chunk] = p;
                __threadfence();
                const int old = atomicAdd(a.arrive + item, 1);
                finalise = (old == a.chunks - 1);
                if (finalise) {
                    __threadfence();
                    best = -1.f; bidx = 0x7fffffff; sum = 0.0;
                    const volatile rslf_partial* pp = a.partials + (size_t)item * a.chunks;
                    for (int k = 0; k < a.chunks; ++k) {
                        const float m = pp[k].mx;
                        sum += pp[k].sum;
                        if (m > best) {
                            best = m; bidx = pp[k].idx; bdv = pp[k].dv;
#pragma unroll
                            for (int c = 0; c < C; ++c) brb[c] = pp[k].rb[c];
                        }
                    }
                    a.arrive[item] = 0;
                }
            }
        }
        if (lane == 0 && finalise) {
            const double maxVal = (double)best;
            if (maxVal > (double)a.raw_thr) {
                a.depth[pix] = bdv;
                const double mean = sum / (double)D;                     /* cv::mean (core.hpp:641) */
                a.cd[pix] = (float)((double)a.ce[pix] * fabs(maxVal - mean));
#pragma unroll
                for (int c = 0; c < C; ++c) a.rbar[(size_t)pix * C + c] = brb[c];
            } else {
                a.ce[pix] = 0.f;                                         /* core.hpp:655-656 */
                a.emask[pix] = 0;
            }
        }
    }
}

struct depth_plan { int H; int blocks_per_sm; size_t smem; int chunks; };

static inline size_t depth_smem_bytes(int S, int C, int H) { return (size_t)depth_padded_views(S) * C * 32 * H * sizeof(float); }

/* Chooses hypotheses per lane: the widest H that still leaves >= 8 resident warps per SM,
 * else the H with the most resident hypotheses.  RSLF_DEPTH_H overrides (experiments). */
static depth_plan plan_depth(const rslf_ctx* ctx, int S, int C, int D)
{
    const size_t budget = 200 * 1024;     /* leave some of the 228 KB to L1 for the EPI gathers */
    int bestH = 1; long bestScore = -1; int bestW = 1;
    const char* env = getenv("RSLF_DEPTH_H");
    int forced = env ? atoi(env) : 0;
    for (int H = 4; H >= 1; H >>= 1) {
        if (forced && H != forced) continue;
        size_t sm = depth_smem_bytes(S, C, H);
        if (sm > ctx->smem_optin) continue;
        if (!forced && H > 1 && 32 * (H / 2) >= D) continue;        /* would leave lanes idle */
        int w = (int)(budget / (sm + 1024));
        if (w < 1) w = 1;
        if (w > 32) w = 32;
        long score = (w >= 8) ? (1000000L * H) : (long)w * H;
        if (score > bestScore) { bestScore = score; bestH = H; bestW = w; }
    }
    depth_plan p;
    p.H = bestH; p.blocks_per_sm = bestW; p.smem = depth_smem_bytes(S, C, bestH);
    p.chunks = rslf_div_up(D, 32 * bestH);
    return p;
}

template <int C, int H, bool NONNEG>
static int launch_depth_t(rslf_ctx* ctx, const depth_args& a, const depth_plan& p)
{
    auto kern = depth_kernel<C, H, NONNEG>;
    /* shared-memory carve-out: what the resident warps need, the rest stays L1 for the EPI gathers */
    static int configured_carve = -1;
    int carve = (int)((p.blocks_per_sm * (p.smem + 1024) * 100 + 228 * 1024 - 1) / (228 * 1024));
    if (carve > 100) carve = 100;
    if (configured_carve != carve) {
        RSLF_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        RSLF_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        configured_carve = carve;
    }
    int occ = 0;
    RSLF_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, p.smem));
    if (occ < 1) occ = 1;
    int bps = occ < p.blocks_per_sm ? occ : p.blocks_per_sm;
    kern<<<ctx->num_sm * bps, 32, p.smem, ctx->stream>>>(a);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    ctx->timing.depth_launches += 1;
    return RSLF_OK;
}

static int launch_depth(rslf_ctx* ctx, int C, bool nonneg, const depth_args& a, const depth_plan& p)
{
#define RSLF_DEPTH_CASE(CC, HH)                                                      \
    if (C == CC && p.H == HH)                                                        \
        return nonneg ? launch_depth_t<CC, HH, true>(ctx, a, p) : launch_depth_t<CC, HH, false>(ctx, a, p);
    RSLF_DEPTH_CASE(1, 1) RSLF_DEPTH_CASE(1, 2) RSLF_DEPTH_CASE(1, 4)
    RSLF_DEPTH_CASE(3, 1) RSLF_DEPTH_CASE(3, 2) RSLF_DEPTH_CASE(3, 4)
#undef RSLF_DEPTH_CASE
    snprintf(ctx->err, sizeof(ctx->err), "unsupported channel count %d", C);
    return RSLF_ERR_UNSUPPORTED;
}
