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
 *   - EPI slicing: in view s the hypotheses of a warp item read one contiguous
 *     segment of the EPI scanline (v, s).  The lanes issue one TMA bulk copy
 *     per view (cp.async.bulk + mbarrier complete_tx; SASS UBLKCP / SYNCS) for
 *     ALL views of the item at once, straight into the shared-memory row that
 *     will hold that view's radiances; one mbarrier wait later every segment
 *     is resident, so the HBM/L2 latency is paid once per item instead of once
 *     per view.  Each row is then converted in place: the lanes read their two
 *     neighbouring pixels, synchronise, and overwrite the row with the
 *     interpolated radiances, four views at a time, transposed to
 *     [c][h][lane][view] so that one LDS.128 of a lane fetches a channel of
 *     four consecutive views in the mean shift (conflict-free).
 *   - the first RV views can stay in REGISTERS instead (a first staging round
 *     through the same rows): fewer shared-memory rows per warp, hence more
 *     resident warps when S*C is large, and no LDS for those views.
 *   - out-of-image samples hold a large finite sentinel instead of NaN:
 *     K = max(1 - |x/h|^2, 0) is 0 for it, as for NaN in the reference, and
 *     r*K contributes +0.
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
#include "k_balance.cuh"

#define RSLF_RAD_SENTINEL 1.0e18f

/* Staging of the scanline segments: 0 = one TMA bulk copy per view (UBLKCP, issued lane by lane through the
 * uniform datapath), 1 = 16-byte cp.async copies (LDGSTS), half a warp per view, all lanes in parallel. */
#ifndef RSLF_STAGE_CPASYNC
#define RSLF_STAGE_CPASYNC 0
#endif
/* 1 = the converted radiance blocks are packed at a uniform stride of DEPTH_UNR * C * 32 * H floats (block b's
 * destination only overlaps the staging sub-areas of blocks <= b, which have been consumed by then), so the
 * mean shift needs no block-offset look-ups. */
#ifndef RSLF_COMPACT_BLOCKS
#define RSLF_COMPACT_BLOCKS 1
#endif
/* unroll factor of the mean-shift loop over shared-memory block pairs (1 = 8 views per loop iteration) */
#ifndef RSLF_MS_UNROLL
#define RSLF_MS_UNROLL 1
#endif
#ifndef RSLF_PREFETCH_ITEM
#define RSLF_PREFETCH_ITEM 0
#endif
#define RSLF_PRAGMA_(x) _Pragma(#x)
#define RSLF_PRAGMA(x) RSLF_PRAGMA_(x)

struct __align__(16) rslf_partial {
    float mx; int idx; float dv; float rb[3]; double sum;
};

struct depth_args {
    const float* epi; int V, S, U, D; int s_hat; float slope; float inv; int iters;
    float negzero;                /* -0.0f, opaque to the compiler: addend of the packed multiplies */
    int wpv_q16;                  /* pixels a chunk's hypotheses spread per view step, 16.16 fixed point (row sizing) */
    int reg_last;                 /* register-resident views: 0 = the first RV views, 1 = the last RV (padded) views */
    const int4* rec;              /* work records of the pass: {pixel index in the stack `epi`, dmin, dmax, C_e} (k_balance.cuh) */
    const int* count;             /* number of records (direct mode) */
    int* queue;                   /* work queue of the launch: next unclaimed warp item (zero when the kernel starts) */
    int pix_off;                  /* direct mode: record pixel - pix_off = pixel index in the result planes below */
    float* ce; uint8_t* emask; float* cd; float* depth; float* rbar;   /* planes of line s_hat (direct mode) */
    float raw_thr;
    int chunks; rslf_partial* partials; int* arrive;
    depth_bal bal;                /* pass-balanced multi-GPU mode (bal.n > 0): records / results of all ranks */
    rslf_decision* log; int* log_count; int log_cap; int level;   /* diagnostics: decision log (rslf_cuda_set_decision_log) */
};

/* one warp item from the launch's work queue (lane 0 claims, the warp learns) */
__device__ __forceinline__ int depth_claim(int* queue, int lane)
{
    int w = 0;
    if (lane == 0) w = atomicAdd(queue, 1);
    return __shfl_sync(0xffffffffu, w, 0);
}

/* Outputs of one pixel (core.hpp:630-657), written by one lane: the result planes in direct mode, a result
 * record in the owner's list in balanced mode (applied by bal_apply_kernel). */
template <int C>
__device__ __forceinline__ void depth_emit(const depth_args& a, int rec_pix, float ce, int owner, int ridx, float best, int bidx,
                                           float bdv, double sum, const float (&brb)[C])
{
    const double maxVal = (double)best;
    const bool ok = maxVal > (double)a.raw_thr;
    const double mean = sum / (double)a.D;                         /* cv::mean (core.hpp:641) */
    const float cdv = (float)((double)ce * fabs(maxVal - mean));
    if (a.log) {
        const int at = atomicAdd(a.log_count, 1);
        if (at < a.log_cap) {
            rslf_decision r;
            r.level = a.level; r.s_hat = a.s_hat; r.pix = rec_pix; r.index = bidx; r.score = best;
            r.rbar[0] = brb[0]; r.rbar[1] = brb[C > 1 ? 1 : 0]; r.rbar[2] = brb[C > 2 ? 2 : 0];
            r.disp_conf = cdv; r.disparity = bdv; r.accepted = ok ? 1 : 0; r.reserved = 0;
            a.log[at] = r;
        }
    }
    if (a.bal.n) {
        float4* r = a.bal.res[owner] + 2 * (size_t)ridx;
        r[0] = make_float4(bdv, cdv, brb[0], brb[C > 1 ? 1 : 0]);
        r[1] = make_float4(brb[C > 2 ? 2 : 0], __int_as_float(ok ? 1 : 0), 0.f, 0.f);
        return;
    }
    const int pix = rec_pix - a.pix_off;
    if (ok) {
        a.depth[pix] = bdv;
        a.cd[pix] = cdv;
#pragma unroll
        for (int c = 0; c < C; ++c) a.rbar[(size_t)pix * C + c] = brb[c];
    } else {
        a.ce[pix] = 0.f;                                           /* core.hpp:655-656 */
        a.emask[pix] = 0;
    }
}

/* end of a depth kernel in balanced mode: once every block's result stores are visible system-wide, the last
 * block raises this rank's done flag on every peer */
__device__ __forceinline__ void depth_bal_finish(const depth_args& a)
{
    if (!a.bal.n) return;
    __threadfence_system();
    bal_last_block_signal(a.bal.blocks_done, a.bal.done, a.bal.n, a.bal.seq, 0u);
}

/* remaining &= emask (in place) and list the survivors (core.hpp:510-516); remaining == nullptr: list emask.
 * Every listed pixel gets its work record (k_balance.cuh).  rowwork (optional): per level-0 row of this rank,
 * the number of pixels evaluated so far in the run — row v of this level counts for level-0 row
 * ((v + v0) << shift) - v0_base.  Balanced multi-GPU mode (bal_n > 0): the last block publishes the count. */
struct compact_args {
    const uint8_t* emask; uint8_t* remaining; int n;
    int* items; int4* rec; int* count; unsigned long long* total;
    unsigned* rowwork; int U, v0, shift, v0_base, rows_base;
    int V, border_lo, border_hi, select;
    const float* dmin_map; const float* dmax_map; float dmin_c, dmax_c; const float* ce;
    int pix_off;                                   /* record pixel = local pixel + pix_off (first row of the block in the stack) */
    int bal_n; unsigned seq; unsigned long long* counts_peer[RSLF_MAX_PEERS]; int* blocks_done;
};

__global__ void compact_kernel(const compact_args a)
{
    /* select 0: every row; 1: only the border rows (v < border_lo or v >= V - border_hi); 2: only the others */
    const int lane = threadIdx.x & 31, n = a.n;
    for (int base = (blockIdx.x * blockDim.x + threadIdx.x) - lane; base < n; base += gridDim.x * blockDim.x) {
        int i = base + lane;
        uint8_t m = 0;
        if (i < n && a.select) {
            const int v = i / a.U;
            const bool border = (v < a.border_lo) || (v >= a.V - a.border_hi);
            if (border != (a.select == 1)) i = n;                  /* not this launch's row */
        }
        if (i < n) {
            m = a.emask[i];
            if (a.remaining) { m &= a.remaining[i]; a.remaining[i] = m; }
        }
        unsigned bal = __ballot_sync(0xffffffffu, m != 0);
        if (bal) {
            int pos = 0;
            if (lane == 0) {
                int c = __popc(bal);
                pos = atomicAdd(a.count, c);
                atomicAdd(a.total, (unsigned long long)c);
            }
            pos = __shfl_sync(0xffffffffu, pos, 0);
            if (m) {
                const int at = pos + __popc(bal & ((1u << lane) - 1u));
                a.items[at] = i;
                int4 r;
                r.x = i + a.pix_off;
                r.y = __float_as_int(a.dmin_map ? a.dmin_map[i] : a.dmin_c);
                r.z = __float_as_int(a.dmax_map ? a.dmax_map[i] : a.dmax_c);
                r.w = __float_as_int(a.ce[i]);
                a.rec[at] = r;
            }
            if (a.rowwork && m) {
                int r = (((i / a.U) + a.v0) << a.shift) - a.v0_base;
                r = min(max(r, 0), a.rows_base - 1);
                atomicAdd(a.rowwork + r, 1u);
            }
        }
    }
    if (a.bal_n) {
        /* records and list are read by the peers once the count is published */
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const int old = atomicAdd(a.blocks_done, 1);
            if (old == (int)gridDim.x - 1) {
                *a.blocks_done = 0;
                __threadfence_system();
                const unsigned cnt = (unsigned)*reinterpret_cast<volatile int*>(a.count);
                const unsigned long long v = ((unsigned long long)a.seq << 32) | (unsigned long long)cnt;
                for (int r = 0; r < a.bal_n; ++r) *reinterpret_cast<volatile unsigned long long*>(a.counts_peer[r]) = v;
            }
        }
    }
}

#define DEPTH_UNR 4      /* views per unrolled block of the mean shift; rows are padded to a multiple of it */
#define DEPTH_MAX_ROW_FLOATS 1024

static __host__ __device__ inline int depth_padded_views(int S) { return (S + DEPTH_UNR - 1) / DEPTH_UNR * DEPTH_UNR; }

/*
 * Shared-memory rows.  Row r hosts, in staging round 0 (only if RV > 0), the scanline segment of view r
 * (r < RV), and in round 1 the segment and then the radiances of view RV + r.  Its size is the larger of
 * the radiance row (C * 32 * H floats) and the longest segment it may receive: a chunk's hypotheses spread
 * over ceil(|s_hat - s| * wpv) + 3 pixels in view s (wpv in 16.16 fixed point so that host and device
 * agree to the bit), plus 4 floats of alignment slack.
 */
static __host__ __device__ inline int depth_segment_floats(int s, int S, int s_hat, int wpv_q16, int C)
{
    if (s >= S) return 0;
    long long k = s_hat > s ? s_hat - s : s - s_hat;
    long long span = ((k * (long long)wpv_q16 + 65535) >> 16) + 3;
    long long f = (span * C + 4 + 3) & ~3LL;
    return (int)(f > DEPTH_MAX_ROW_FLOATS ? DEPTH_MAX_ROW_FLOATS : f);
}
/* Shared memory is organised in blocks of DEPTH_UNR views, one table of block offsets per staging round.
 * Round 0 (only if RV > 0) stages the RV register-resident views, round 1 the other Spad - RV views; both
 * rounds use the same area from its start, so the area is as large as the larger round.  The register views
 * are the first RV views (reg_last = 0) or the last RV views of the padded range (reg_last = 1): the host
 * picks the end that is farther from s_hat, whose segments are the longest, so that the views that stay in
 * shared memory are the ones with short segments (a pass at the rim of the stack then needs hardly more shared
 * memory than the centre pass, and as many warps stay resident).  While staged, view j of a block owns the
 * sub-area [j * pitch, (j + 1) * pitch) with pitch = max(C * 32 * H, longest segment of the block); after the
 * conversion the block's first DEPTH_UNR * C * 32 * H floats hold the radiances transposed as
 * [c][h][lane][view j], so that ONE LDS.128 of a lane fetches a channel of all four views. */
static __host__ __device__ inline int depth_round_base(int round, int Spad, int RV, int reg_last)
{
    return round == 0 ? (reg_last ? Spad - RV : 0) : (reg_last ? 0 : RV);
}
static __host__ __device__ inline int depth_block_pitch(int round, int b, int S, int Spad, int RV, int reg_last, int s_hat,
                                                        int wpv_q16, int C, int W)
{
    const int base = depth_round_base(round, Spad, RV, reg_last);
    const int nviews = round == 0 ? RV : Spad - RV;
    int f = C * W;
    for (int j = 0; j < DEPTH_UNR; ++j) {
        const int r = b * DEPTH_UNR + j;
        if (r < nviews) { int g = depth_segment_floats(base + r, S, s_hat, wpv_q16, C); f = g > f ? g : f; }
    }
    return f;
}
static __host__ __device__ inline int depth_num_rows(int Spad, int RV) { return (Spad - RV) > RV ? (Spad - RV) : RV; }

/* ---- mbarrier / TMA bulk-copy primitives (PTX; SASS: SYNCS.*, UBLKCP) ---- */
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    /* try_wait suspends the warp for a hardware-defined time slice; a copy that never lands (a protocol
     * bug) traps instead of hanging the device */
    unsigned done = 0;
    for (int spin = 0; spin < (1 << 20); ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

/*
 * Segment of scanline (v, s) that the chunk's hypotheses touch: pixels floor(min I) .. ceil(max I) with
 * I = (s_hat - s) * D * slope + u at the chunk's first and last hypothesis (every float operation involved
 * is monotone, so all other hypotheses fall in between), clipped to the image.  lo / hi are returned;
 * lo > hi: nothing to stage (padding view, or the whole chunk leaves the image in this view).
 */
__device__ __forceinline__ void view_span(const depth_args& a, int s, float uf, float Dlo, float Dhi, int& lo_i, int& hi_i)
{
    const float k = (float)(a.s_hat - s);
    float Ia = k * Dlo; Ia = Ia * a.slope; Ia = Ia + uf;
    float Ib = k * Dhi; Ib = Ib * a.slope; Ib = Ib + uf;
    const float lo = floorf(fminf(Ia, Ib)), hi = ceilf(fmaxf(Ia, Ib));
    lo_i = 1; hi_i = 0;
    if (s >= a.S || !(lo <= hi)) return;                     /* padding view or NaN */
    lo_i = (int)fmaxf(lo, 0.f); hi_i = (int)fminf(hi, (float)(a.U - 1));
}

/* One view of the mean-shift sums for the lane's H hypotheses (core.hpp:591-603, kern.cpp:16-26 / 39-54,
 * core.cpp:25-38): K = max(1 - |(r - r_bar)/h|^2, 0); sum_K += K; sum_rK += max(r, 0) * K. */
template <int C, int H, bool NONNEG, bool FAST = false>
__device__ __forceinline__ void ms_accumulate(const float (&r)[C][H], const float (&rb)[C][H], float inv,
                                              float (&sR)[C][H], float (&sK)[H])
{
    if constexpr (FAST) {
        /* Contracted form (opt-in, rslf_cuda_set_fast_math): the same sums with fused multiply-adds, 6 C + 2 -> 4 C
         * operations per view and hypothesis (RGB: 20 -> 12 instructions).  NOT bit-identical to the reference: every
         * fused operation drops one intermediate rounding, scores move by a few 1e-7 relative, inside the north star's
         * tolerance (indices where the winning margin exceeds 1e-5, 1e-4 relative on scores / confidences). */
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float t;
            {
                const float x0 = r[0][h] - rb[0][h];
                t = x0 * x0;
            }
#pragma unroll
            for (int c = 1; c < C; ++c) {
                const float x = r[c][h] - rb[c][h];
                t = __fmaf_rn(x, x, t);
            }
            const float kk = fmaxf(__fmaf_rn(-inv, t, 1.0f), 0.f);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float r0 = NONNEG ? r[c][h] : fmaxf(r[c][h], 0.f);
                sR[c][h] = __fmaf_rn(r0, kk, sR[c][h]);
            }
            sK[h] = sK[h] + kk;
        }
    } else {
#pragma unroll
    for (int h = 0; h < H; ++h) {
        float b;
        if (C == 1) {
            const float x = r[0][h] - rb[0][h];
            b = (inv * x) * x;                                  /* multiply(src, src, scale) */
        } else {
            const float x0 = r[0][h] - rb[0][h];
            const float x1 = r[C > 1 ? 1 : 0][h] - rb[C > 1 ? 1 : 0][h];
            const float x2 = r[C > 2 ? 2 : 0][h] - rb[C > 2 ? 2 : 0][h];
            const float b0 = (inv * x0) * x0;
            const float b1 = (inv * x1) * x1;
            const float b2 = (inv * x2) * x2;
            b = (b0 + b1) + b2;                                 /* reduce over channels: (c0 + c1) + c2 */
        }
        const float kk = fmaxf(1.0f - b, 0.f);
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float r0 = NONNEG ? r[c][h] : fmaxf(r[c][h], 0.f);
            const float p = r0 * kk;
            sR[c][h] = sR[c][h] + p;
        }
        sK[h] = sK[h] + kk;
    }
    }
}


/* ---- packed FP32x2 arithmetic (Blackwell FFMA2 / FADD2) ------------------------------------------------
 * sm_100a issues two separately rounded FP32 operations per lane with one instruction; a micro-benchmark
 * whose second operands are loop constants (served by the operand-reuse cache) reaches 73.9 T multiply /
 * add operations per second against 36.2 T with scalar instructions (k_peak.cuh).  In the mean shift the
 * operands are fresh register pairs, a packed instruction then reads 4 - 6 registers and the register
 * file's two banks make it issue over 2 - 3 cycles: measured on B200 the packed variant is bit-identical
 * to the scalar one and NOT faster (C3: 1347 ms against 1321 ms in the depth kernel), so it is compiled
 * out by default (RSLF_PACKED_FP32=1 re-enables it; tests pass either way).  Each half of a packed
 * operation is an ordinary round-to-nearest-even operation.
 * CAUTION: ptxas 12.9 contracts a packed multiply that feeds a packed add into one FFMA2 even with
 * --fmad=false and explicit .rn (seen in SASS), which would change the rounding.  The multiply is therefore
 * written as fma(a, b, nz) with nz = -0.0 read from the kernel arguments: ptxas cannot know the value, so
 * it stays a genuine FFMA2 that nothing can be contracted with, and a * b + (-0) rounds exactly like
 * a * b (including the sign of a zero product).
 */
#ifndef RSLF_PACKED_FP32
#define RSLF_PACKED_FP32 0      /* measured: no faster than scalar in this kernel, see below */
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b, f32x2 nz) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
    return r;
}

/*
 * Two consecutive views (va: view s, vb: view s + 1) of the mean-shift sums, same arithmetic and the same
 * accumulation order as two calls of ms_accumulate, issued as packed operations:
 *   H even   : pairs of neighbouring hypotheses, every operation packed;
 *   C=3, H=1 : (c0, c1) of one view form a pair, c2 of the two views form a pair;
 *   C=1, H=1 : the two views form a pair (the two accumulations stay scalar and ordered).
 */
template <int C, int H, bool NONNEG, bool FAST = false>
__device__ __forceinline__ void ms_accumulate_pair(const float (&va)[C][H], const float (&vb)[C][H], const float (&rb)[C][H],
                                                   float inv, f32x2 NZ, float (&sR)[C][H], float (&sK)[H])
{
    if constexpr (FAST) {
        ms_accumulate<C, H, NONNEG, true>(va, rb, inv, sR, sK);
        ms_accumulate<C, H, NONNEG, true>(vb, rb, inv, sR, sK);
    } else {
#if RSLF_PACKED_FP32
    const f32x2 INV = pk2(inv, inv), ONE = pk2(1.0f, 1.0f);
    if (H % 2 == 0) {
#pragma unroll
        for (int view = 0; view < 2; ++view) {
            const float (&r)[C][H] = view ? vb : va;
#pragma unroll
            for (int hp = 0; hp < H / 2; ++hp) {
                const int h0 = 2 * hp, h1 = (2 * hp + 1) % H;
                f32x2 R[C], B[C];
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    R[c] = pk2(r[c][h0], r[c][h1]);
                    const f32x2 X = sub2(R[c], pk2(rb[c][h0], rb[c][h1]));
                    B[c] = mul2(mul2(INV, X, NZ), X, NZ);
                }
                f32x2 Bs = B[0];
                if (C == 3) Bs = add2(add2(B[0], B[C > 1 ? 1 : 0]), B[C > 2 ? 2 : 0]);
                float k0, k1;
                upk2(sub2(ONE, Bs), k0, k1);
                k0 = fmaxf(k0, 0.f); k1 = fmaxf(k1, 0.f);
                const f32x2 K = pk2(k0, k1);
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    f32x2 R0 = R[c];
                    if (!NONNEG) R0 = pk2(fmaxf(r[c][h0], 0.f), fmaxf(r[c][h1], 0.f));
                    const f32x2 A = add2(pk2(sR[c][h0], sR[c][h1]), mul2(R0, K, NZ));
                    upk2(A, sR[c][h0], sR[c][h1]);
                }
                upk2(add2(pk2(sK[h0], sK[h1]), K), sK[h0], sK[h1]);
            }
        }
    } else if (C == 3) {
        constexpr int c1 = C > 1 ? 1 : 0, c2 = C > 2 ? 2 : 0;
        const f32x2 RB01 = pk2(rb[0][0], rb[c1][0]), RB2 = pk2(rb[c2][0], rb[c2][0]);
        const f32x2 R01a = pk2(va[0][0], va[c1][0]), R01b = pk2(vb[0][0], vb[c1][0]), R2 = pk2(va[c2][0], vb[c2][0]);
        const f32x2 X01a = sub2(R01a, RB01), X01b = sub2(R01b, RB01), X2 = sub2(R2, RB2);
        const f32x2 B01a = mul2(mul2(INV, X01a, NZ), X01a, NZ), B01b = mul2(mul2(INV, X01b, NZ), X01b, NZ), B2 = mul2(mul2(INV, X2, NZ), X2, NZ);
        float b0a, b1a, b0b, b1b;
        upk2(B01a, b0a, b1a); upk2(B01b, b0b, b1b);
        const float sa = b0a + b1a, sb = b0b + b1b;                  /* (c0 + c1), then + c2 */
        float ka, kb;
        upk2(sub2(ONE, add2(pk2(sa, sb), B2)), ka, kb);
        ka = fmaxf(ka, 0.f); kb = fmaxf(kb, 0.f);
        f32x2 Q01a = R01a, Q01b = R01b, Q2 = R2;
        if (!NONNEG) {
            Q01a = pk2(fmaxf(va[0][0], 0.f), fmaxf(va[c1][0], 0.f)); Q01b = pk2(fmaxf(vb[0][0], 0.f), fmaxf(vb[c1][0], 0.f));
            Q2 = pk2(fmaxf(va[c2][0], 0.f), fmaxf(vb[c2][0], 0.f));
        }
        f32x2 A01 = pk2(sR[0][0], sR[c1][0]);
        A01 = add2(A01, mul2(Q01a, pk2(ka, ka), NZ));                    /* view s first, then view s + 1 */
        A01 = add2(A01, mul2(Q01b, pk2(kb, kb), NZ));
        upk2(A01, sR[0][0], sR[c1][0]);
        float p2a, p2b;
        upk2(mul2(Q2, pk2(ka, kb), NZ), p2a, p2b);
        sR[c2][0] = sR[c2][0] + p2a; sR[c2][0] = sR[c2][0] + p2b;
        sK[0] = sK[0] + ka; sK[0] = sK[0] + kb;
    } else {
        const f32x2 R = pk2(va[0][0], vb[0][0]);
        const f32x2 X = sub2(R, pk2(rb[0][0], rb[0][0]));
        float ka, kb;
        upk2(sub2(ONE, mul2(mul2(INV, X, NZ), X, NZ)), ka, kb);
        ka = fmaxf(ka, 0.f); kb = fmaxf(kb, 0.f);
        f32x2 Q = R;
        if (!NONNEG) Q = pk2(fmaxf(va[0][0], 0.f), fmaxf(vb[0][0], 0.f));
        float pa, pb;
        upk2(mul2(Q, pk2(ka, kb), NZ), pa, pb);
        sR[0][0] = sR[0][0] + pa; sR[0][0] = sR[0][0] + pb;
        sK[0] = sK[0] + ka; sK[0] = sK[0] + kb;
    }
#else
    ms_accumulate<C, H, NONNEG>(va, rb, inv, sR, sK);
    ms_accumulate<C, H, NONNEG>(vb, rb, inv, sR, sK);
#endif
    }
}

/*
 * In-place conversion of the staged rows of one round, DEPTH_UNR views per step (all segment reads of
 * the step, one warp synchronisation, then the radiance stores): I = (s_hat - s) * D * slope + u
 * (core.hpp:550-552); floor / ceil neighbours, in-image test and linear interpolation
 * (interp.hpp:171-185).  Views >= S (padding) and out-of-image samples become the sentinel.
 * ufl = u as float, NaN for lanes whose hypothesis index is >= D (they then fail every test).
 */
/* float offset of radiance block b inside the area */
template <int C, int H>
__device__ __forceinline__ int rad_block(const int* blk_off, int b)
{
#if RSLF_COMPACT_BLOCKS
    return b * (DEPTH_UNR * C * 32 * H);
#else
    return blk_off[b];
#endif
}

template <int C, int H, bool FALLBACK>
__device__ __forceinline__ void convert_rows(const depth_args& a, const int4* meta, const int* blk_off, float* rows, int vbase, int nviews,
                                             int lane, long long row0, const float (&ufl)[H], float Um1f,
                                             const float (&Dv)[H], int (&cardi)[H])
{
#pragma unroll 1
    for (int rb0 = 0; rb0 < nviews; rb0 += DEPTH_UNR) {
        float val[DEPTH_UNR][C][H];
#pragma unroll
        for (int j = 0; j < DEPTH_UNR; ++j) {
            const int s = vbase + rb0 + j;
            const int4 m = meta[rb0 + j];
            const float* row = rows + m.z;
            const float k = (float)(a.s_hat - s);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float I = k * Dv[h];
                I = I * a.slope;
                I = I + ufl[h];
                const float fl = floorf(I);
                const int i0 = (int)fl;
                const int ne = (I != fl) ? 1 : 0;                   /* ceil(I) = floor(I) + ne */
                /* i0 >= 0 and i1 <= U-1  <=>  0 <= I <= U-1 (NaN fails both) */
                const bool ok = (I >= 0.f) && (I <= Um1f) && (s < a.S);
                const float t = I - fl;                             /* = I - float(i0) */
                const float omt = 1.f - t;
                const int p0 = m.x + i0 * C, p1 = p0 + ne * C;
                float e0[C], e1[C];
                if (FALLBACK && m.w && ok && (p0 < 0 || p1 + C > m.y)) {
                    const float* g = a.epi + (row0 + (long long)s * a.U) * C;
#pragma unroll
                    for (int c = 0; c < C; ++c) { e0[c] = __ldg(g + (size_t)i0 * C + c); e1[c] = __ldg(g + (size_t)(i0 + ne) * C + c); }
                } else {
                    const int q0 = ok ? p0 : 0, q1 = ok ? p1 : 0;
#pragma unroll
                    for (int c = 0; c < C; ++c) { e0[c] = row[q0 + c]; e1[c] = row[q1 + c]; }
                }
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float p = omt * e0[c];
                    const float q = t * e1[c];
                    val[j][c][h] = ok ? (p + q) : RSLF_RAD_SENTINEL;
                }
                cardi[h] += ok ? 1 : 0;
            }
        }
        __syncwarp();                                               /* every lane has read the segments */
        /* radiances of the block, transposed: [c][h][lane][view j] */
        float* blk = rows + rad_block<C, H>(blk_off, rb0 / DEPTH_UNR);
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int h = 0; h < H; ++h)
                *reinterpret_cast<float4*>(blk + ((c * H + h) * 32 + lane) * DEPTH_UNR) =
                    make_float4(val[0][c][h], val[1][c][h], val[2][c][h], val[3][c][h]);
    }
}

/* Register budget.  The register file is split per SM sub-partition (512 registers per lane each) and
 * the one-warp blocks are spread over the four sub-partitions, so k resident warps per sub-partition
 * may use 512 / k registers: 255 (k = 2), 168 (3), 128 (4), 96 (5), 80 (6), 64 (8). */
#define DEPTH_REG_ESTIMATE(C, H, RV) ((RV) * (C) * (H) + 48 + 36 * (H))
template <int C, int H, int RV> struct depth_min_blocks {
    enum { est = DEPTH_REG_ESTIMATE(C, H, RV), wps = 512 / est, value = 4 * (wps > 8 ? 8 : (wps < 1 ? 1 : wps)) };
};

template <int C, int H, bool NONNEG, int RV>
__global__ void __launch_bounds__(32, depth_min_blocks<C, H, RV>::value)
depth_kernel(const depth_args a)
{
    extern __shared__ float4 smem_raw[];
    const int lane = threadIdx.x;
    const int S = a.S, U = a.U, D = a.D;
    const int Spad = max(depth_padded_views(S), RV);
    const int nrows = depth_num_rows(Spad, RV);
    const int nblk = (Spad - RV) / DEPTH_UNR;               /* blocks of DEPTH_UNR shared-memory views */
    constexpr int W = 32 * H;
    /* [mbarrier, 16 B][block offsets of round 0: nb0 + 1 ints, padded to 16 B][block offsets of round 1: nblk + 1 ints,
     *  padded][per-row staging records: nrows int4][staging area / radiance blocks] */
    constexpr int nb0 = RV / DEPTH_UNR;
    const int rlast = a.reg_last;
    int* blk_off0 = reinterpret_cast<int*>(smem_raw) + 4;
    int* blk_off1 = blk_off0 + ((nb0 + 1 + 3) & ~3);
    int4* meta = reinterpret_cast<int4*>(blk_off1 + ((nblk + 1 + 3) & ~3));
    float* rows = reinterpret_cast<float*>(meta + nrows);
    const unsigned bar = smem_u32(smem_raw);
    const float inv = a.inv;
    const f32x2 NZ = pk2(a.negzero, a.negzero);          /* -0.0 the compiler cannot see (see mul2) */
    const float Um1f = (float)(U - 1);

    /* prologue: mbarrier and the block offset tables (exclusive prefix sums of the block sizes) */
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
#pragma unroll 1
    for (int round = (RV > 0 ? 0 : 1); round < 2; ++round) {
        int* tab = round ? blk_off1 : blk_off0;
        const int nb = round ? nblk : nb0;
        int carry = 0;
        for (int base = 0; base < nb; base += 32) {
            const int b = base + lane;
            int f = (b < nb) ? DEPTH_UNR * depth_block_pitch(round, b, S, Spad, RV, rlast, a.s_hat, a.wpv_q16, C, W) : 0;
            int incl = f;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, off);
                if (lane >= off) incl += t;
            }
            if (b < nb) tab[b] = carry + incl - f;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) tab[nb] = carry;
    }
    __syncwarp();
    unsigned phase = 0;                                     /* mbarrier phase parity */

    /* this launch's warp items: (record, chunk) pairs, claimed from the launch's work queue.  Balanced mode: the
     * records of this rank's share of the pass, fetched from their owners. */
    bal_slice sl; sl.lo = 0; sl.n = 0; sl.pre_excl = 0; sl.pre_incl = 0;
    int total;
    if (a.bal.n) { sl = bal_begin(a.bal, lane); total = sl.n * a.chunks; }
    else total = (*a.count) * a.chunks;
    int w = depth_claim(a.queue, lane);
    while (w < total) {
        int w_next = 0;
        if (lane == 0) w_next = atomicAdd(a.queue, 1);          /* claimed now, needed after this item */
        const int item = w / a.chunks;
        const int chunk = w - item * a.chunks;
        int owner = 0, ridx = item;
        int4 rec;
        if (a.bal.n) { bal_locate(sl, sl.lo + item, a.bal.n, lane, owner, ridx); rec = ld_cv_int4(a.bal.rec[owner] + ridx); }
        else rec = a.rec[item];
        const int pix = rec.x;
        const float dmin = __int_as_float(rec.y), dmax = __int_as_float(rec.z);
        /* dmin == dmax (bounds of a coarser level that met, ftc.hpp:201-294): all D hypotheses are the same EPI
         * line, so every score equals the first one: d* = 0, C_d = 0.  Chunk 0 evaluates it, the others are void. */
        const bool flat = (a.chunks > 1) && (dmin == dmax);
        if (flat && chunk > 0) { w = __shfl_sync(0xffffffffu, w_next, 0); continue; }
        const int v = pix / U, u = pix - v * U;
        const int dbase = chunk * W + lane * H;
        const float uf = (float)u;
        float ufl[H];                                          /* u, or NaN for padding hypotheses (d >= D) */
#pragma unroll
        for (int h = 0; h < H; ++h) ufl[h] = (dbase + h < D) ? uf : __int_as_float(0x7fc00000);
        /* D[d] = dmin + d * (dmax - dmin) / (dim_d - 1)   (core.hpp:547-548) */
        float Dv[H], Dlo, Dhi;
        {
            const float range = dmax - dmin;
            const float den = (float)(D - 1);
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float t = (float)(dbase + h) * range;
                t = t / den;
                Dv[h] = dmin + t;
            }
            /* first and last real hypothesis of the chunk: they bound every lane's EPI line */
            float t = (float)(chunk * W) * range; t = t / den; Dlo = dmin + t;
            t = (float)(min(chunk * W + W, D) - 1) * range; t = t / den; Dhi = dmin + t;
        }
        const long long row0 = (long long)v * S * U;           /* pixel index of (v, s = 0, u = 0) */
        int cardi[H];
        float rb[C][H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            cardi[h] = 0;
#pragma unroll
            for (int c = 0; c < C; ++c) rb[c][h] = 0.f;
        }
        float rr[RV > 0 ? RV : 1][C][H];

        /* ---- radiances (interp.hpp:155-193) and card_R.  Round 0 (RV > 0): the RV register views (the first or
         *      the last RV views, a.reg_last) -> registers; round 1: the other views -> shared-memory blocks.  Per round: the lanes issue one TMA bulk
         *      copy per view into that view's row, wait once, then convert the rows in place. ---- */
#pragma unroll 1
        for (int round = (RV > 0 ? 0 : 1); round < 2; ++round) {
            const int vbase = depth_round_base(round, Spad, RV, rlast);
            const int nviews = round ? (Spad - RV) : RV;
            const int* blk_off = round ? blk_off1 : blk_off0;
            bool any_cut = false;
            __syncwarp();
            {
                /* staging record of row r: x = float offset of the scanline's pixel 0 inside the staged
                 * segment, y = floats staged (multiple of 4), z = row offset, w = 1 if the segment was cut */
                unsigned bytes = 0, cut = 0;
                for (int r = lane; r < nviews; r += 32) {
                    int lo, hi;
                    view_span(a, vbase + r, uf, Dlo, Dhi, lo, hi);
                    const int bo = blk_off[r / DEPTH_UNR], pitch = (blk_off[r / DEPTH_UNR + 1] - bo) / DEPTH_UNR;
                    int4 m; m.x = 0; m.y = 0; m.z = bo + (r % DEPTH_UNR) * pitch; m.w = 0;
                    if (lo <= hi) {
                        const int mis = ((int)((row0 + (long long)(vbase + r) * U + lo) & 3LL) * C) & 3;
                        const int want = ((hi + 1 - lo) * C + mis + 3) & ~3;
                        const int cap = pitch;
                        m.x = mis - lo * C; m.y = min(want, cap); m.w = want > cap;
                        bytes += 4u * (unsigned)m.y; cut |= (unsigned)m.w;
                    }
                    meta[r] = m;
                }
                any_cut = __any_sync(0xffffffffu, cut != 0);
#if RSLF_STAGE_CPASYNC
                (void)bytes;
                __syncwarp();                                           /* the staging records are read by other lanes */
                /* half a warp per view, lane q of the half copies the 16-byte pieces q, q + 16, ... of the segment */
                const int half = lane >> 4, l16 = lane & 15;
#pragma unroll 1
                for (int r0 = 0; r0 < nviews; r0 += 2) {
                    const int r = r0 + half;
                    const int4 m = meta[r < nviews ? r : r0];
                    const int n4 = (r < nviews) ? (m.y >> 2) : 0;
                    const float* src = a.epi + ((row0 + (long long)(vbase + r) * U) * C - m.x);
                    const unsigned dst = smem_u32(rows + m.z);
                    for (int q = l16; q < n4; q += 16) cp_async16(dst + 16u * (unsigned)q, src + 4 * q);
                }
            }
            cp_async_wait_all();
            __syncwarp();
#else
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, off);
                if (lane == 0) mbar_arrive_expect_tx(bar, bytes);
                for (int r = lane; r < nviews; r += 32) {
                    const int4 m = meta[r];
                    /* segment start in the stack: (row0 + s*U) * C - m.x floats, 16-byte aligned by construction */
                    if (m.y > 0)
                        tma_bulk_g2s(smem_u32(rows + m.z), a.epi + ((row0 + (long long)(vbase + r) * U) * C - m.x), 4u * (unsigned)m.y, bar);
                }
            }
            mbar_wait(bar, phase);
            phase ^= 1u;
#endif
            /* in-place conversion of the rows; the variant with the global-memory fallback is only needed when a
             * segment was cut to its row (user-edited bounds wider than the global range) */
            if (any_cut) convert_rows<C, H, true>(a, meta, blk_off, rows, vbase, nviews, lane, row0, ufl, Um1f, Dv, cardi);
            else convert_rows<C, H, false>(a, meta, blk_off, rows, vbase, nviews, lane, row0, ufl, Um1f, Dv, cardi);
            __syncwarp();
            /* r_bar <- radiances of view s_hat (core.hpp:577) */
            if (a.s_hat >= vbase && a.s_hat < vbase + nviews) {
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int h = 0; h < H; ++h)
                        rb[c][h] = rows[rad_block<C, H>(blk_off, (a.s_hat - vbase) / DEPTH_UNR) + ((c * H + h) * 32 + lane) * DEPTH_UNR + (a.s_hat - vbase) % DEPTH_UNR];
            }
            if (RV > 0 && round == 0) {
#pragma unroll
                for (int b = 0; b < RV / DEPTH_UNR; ++b)
#pragma unroll
                    for (int c = 0; c < C; ++c)
#pragma unroll
                        for (int h = 0; h < H; ++h) {
                            const float4 t = *reinterpret_cast<const float4*>(rows + rad_block<C, H>(blk_off, b) + ((c * H + h) * 32 + lane) * DEPTH_UNR);
                            rr[b * DEPTH_UNR + 0][c][h] = t.x; rr[b * DEPTH_UNR + 1][c][h] = t.y;
                            rr[b * DEPTH_UNR + 2][c][h] = t.z; rr[b * DEPTH_UNR + 3][c][h] = t.w;
                        }
            }
        }
        float card[H];
#pragma unroll
        for (int h = 0; h < H; ++h) card[h] = (float)cardi[h];
        /* ---- mean shift (core.hpp:577-610).  Shared-memory rows are consumed in blocks of DEPTH_UNR
         *      views with two register buffers in ping-pong: the LDS of the next block are in flight
         *      while the current one is accumulated.  Sentinel rows add K = 0 and r*K = +0, which
         *      leaves every partial sum bit-identical. ---- */
        float sK[H];
#pragma unroll
        for (int h = 0; h < H; ++h) sK[h] = 0.f;
        for (int it = 0; it < a.iters; ++it) {
            float sR[C][H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                sK[h] = 0.f;
#pragma unroll
                for (int c = 0; c < C; ++c) sR[c][h] = 0.f;
            }
            float ba[DEPTH_UNR][C][H], bb[DEPTH_UNR][C][H];
            auto load_block = [&](const float* blk, float (&dst)[DEPTH_UNR][C][H]) {
#pragma unroll
                for (int c = 0; c < C; ++c)
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        const float4 t = *reinterpret_cast<const float4*>(blk + (c * H + h) * 32 * DEPTH_UNR);
                        dst[0][c][h] = t.x; dst[1][c][h] = t.y; dst[2][c][h] = t.z; dst[3][c][h] = t.w;
                    }
            };
#if RSLF_COMPACT_BLOCKS
            constexpr int BLK = DEPTH_UNR * C * 32 * H;                 /* floats per radiance block */
            const float* const p0 = rows + lane * DEPTH_UNR;
            const float* const plast = p0 + (nblk > 0 ? nblk - 1 : 0) * BLK;
#define RSLF_BLOCK_PTR(bi) (p0 + (bi) * BLK)
#else
#define RSLF_BLOCK_PTR(bi) (rows + blk_off1[bi] + lane * DEPTH_UNR)
#endif
            if (nblk > 0) load_block(RSLF_BLOCK_PTR(0), ba);
            /* views in ascending s (the order of cv::reduce): register views first, or last (a.reg_last) */
#pragma unroll 1
            for (int part = 0; part < 2; ++part) {
                if ((part == 1) == (rlast != 0)) {
#pragma unroll
                    for (int j = 0; j < RV; j += 2) ms_accumulate_pair<C, H, NONNEG>(rr[j], rr[RV > 1 ? j + 1 : 0], rb, inv, NZ, sR, sK);
                } else {
#if RSLF_COMPACT_BLOCKS
                    /* two blocks per step, running pointer, the prefetch of the step after the last clamped to the last block */
                    const float* p = p0;
RSLF_PRAGMA(unroll RSLF_MS_UNROLL)
                    for (int n = nblk >> 1; n > 0; --n) {
                        load_block(p + BLK, bb);
#pragma unroll
                        for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG>(ba[j], ba[j + 1], rb, inv, NZ, sR, sK);
                        p += 2 * BLK;
                        load_block(p < plast ? p : plast, ba);
#pragma unroll
                        for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG>(bb[j], bb[j + 1], rb, inv, NZ, sR, sK);
                    }
                    if (nblk & 1) {
#pragma unroll
                        for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG>(ba[j], ba[j + 1], rb, inv, NZ, sR, sK);
                    }
#else
                    int bi = 0;
                    for (; bi + 2 <= nblk; bi += 2) {
                        load_block(RSLF_BLOCK_PTR(bi + 1), bb);
#pragma unroll
                        for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG>(ba[j], ba[j + 1], rb, inv, NZ, sR, sK);
                        load_block(RSLF_BLOCK_PTR(min(bi + 2, nblk - 1)), ba);   /* last: harmless re-read */
#pragma unroll
                        for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG>(bb[j], bb[j + 1], rb, inv, NZ, sR, sK);
                    }
                    if (bi < nblk) {
#pragma unroll
                        for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG>(ba[j], ba[j + 1], rb, inv, NZ, sR, sK);
                    }
#endif
                }
            }
#undef RSLF_BLOCK_PTR
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
        if (flat) {
            sum = (double)best * (double)D;                              /* D equal scores: their double sum is exact */
        } else if (a.chunks > 1) {
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
        if (lane == 0 && finalise) depth_emit<C>(a, pix, __int_as_float(rec.w), owner, ridx, best, bidx, bdv, sum, brb);
        w = __shfl_sync(0xffffffffu, w_next, 0);
    }
    depth_bal_finish(a);
}

struct depth_plan { int H; int RV; int reg_last; int blocks_per_sm; size_t smem; int chunks; int wpv_q16; };

/* pixels a chunk of 32*H hypotheses spreads per view step, 16.16 fixed point, rounded up */
static inline int depth_wpv_q16(int D, int H, float dmin, float dmax, float slope)
{
    double per_view = std::fabs((double)dmax - (double)dmin) * std::fabs((double)slope) * (32.0 * H - 1.0) / (double)(D - 1);
    if (!(per_view < 1024.0)) per_view = 1024.0;
    return (int)std::ceil(per_view * 65536.0);
}

static inline size_t depth_smem_bytes(int S, int C, int H, int RV, int reg_last, int s_hat, int wpv_q16)
{
    int spad = depth_padded_views(S);
    if (spad < RV) spad = RV;
    const int nrows = depth_num_rows(spad, RV), nb0 = RV / DEPTH_UNR, nb1 = (spad - RV) / DEPTH_UNR;
    size_t fl = 4 + ((nb0 + 1 + 3) & ~3) + ((nb1 + 1 + 3) & ~3) + 4 * (size_t)nrows;   /* mbarrier, block offsets, staging records */
    size_t area0 = 0, area1 = 0;
    for (int b = 0; b < nb0; ++b) area0 += (size_t)DEPTH_UNR * depth_block_pitch(0, b, S, spad, RV, reg_last, s_hat, wpv_q16, C, 32 * H);
    for (int b = 0; b < nb1; ++b) area1 += (size_t)DEPTH_UNR * depth_block_pitch(1, b, S, spad, RV, reg_last, s_hat, wpv_q16, C, 32 * H);
    return (fl + (area0 > area1 ? area0 : area1)) * sizeof(float);
}

/* (H, RV) variants that are instantiated */
static const int k_depth_variants[][2] = {{1, 0}, {1, 16}, {1, 32}, {2, 0}, {2, 16}, {4, 0}};

/* resident warps per SM allowed by the register file (see depth_min_blocks) */
static inline int depth_reg_warps(int C, int H, int RV)
{
    int wps = 512 / DEPTH_REG_ESTIMATE(C, H, RV);
    return 4 * (wps > 8 ? 8 : (wps < 1 ? 1 : wps));
}

/*
 * Chooses hypotheses per lane (H) and register-resident views (RV).  Measured on B200 (profiles/):
 * beyond 8 resident warps per SM (2 per scheduler) more residency buys nothing, while every
 * register-resident view saves its LDS and shrinks the warp's shared memory; wide lanes (H = 2, 4) pay
 * off only for gray stacks, whose rows are small.  So: RGB -> H = 1 and the largest RV that still leaves
 * >= 8 warps; gray -> the widest H that D fills, RV = 0.  Among equals, more resident warps win.
 * RSLF_DEPTH_H / RSLF_DEPTH_RV force a variant (experiments, tests).
 */
static depth_plan plan_depth(const rslf_ctx* ctx, int S, int C, int D, int s_hat, float dmin, float dmax, float slope)
{
    const size_t budget = 224 * 1024;     /* the gathers go through TMA, L1 needs no share */
    const char* eh = getenv("RSLF_DEPTH_H");
    const char* er = getenv("RSLF_DEPTH_RV");
    const int fH = eh ? atoi(eh) : 0, fRV = er ? atoi(er) : -1;
    const char* el = getenv("RSLF_DEPTH_REG_LAST");
    const int fL = el ? atoi(el) : -1;
    depth_plan best; best.H = 1; best.RV = 0; best.reg_last = 0; best.blocks_per_sm = 1;
    best.wpv_q16 = depth_wpv_q16(D, 1, dmin, dmax, slope);
    best.smem = depth_smem_bytes(S, C, 1, 0, 0, s_hat, best.wpv_q16);
    double bestScore = -1;
    for (auto& var : k_depth_variants) {
        const int H = var[0], RV = var[1];
        if (fH && H != fH) continue;
        if (fRV >= 0 && RV != fRV) continue;
        if (!fH && H > 1 && 32 * (H / 2) >= D) continue;            /* would leave lanes idle */
        if (fRV < 0 && RV > 0 && RV > depth_padded_views(S) - DEPTH_UNR) continue;
        if (!fH && fRV < 0 && C == 3 && H != 1) continue;
        if (!fH && fRV < 0 && C == 1 && RV != 0) continue;
        const int wpv = depth_wpv_q16(D, H, dmin, dmax, slope);
        /* register views at the end of the stack that is farther from s_hat (the smaller shared-memory need) */
        size_t sm = depth_smem_bytes(S, C, H, RV, 0, s_hat, wpv);
        int rl = 0;
        if (RV > 0 && fL != 0) {
            const size_t sm1 = depth_smem_bytes(S, C, H, RV, 1, s_hat, wpv);
            if (sm1 < sm || fL == 1) { sm = sm1; rl = 1; }
        }
        if (sm > ctx->smem_optin) continue;
        int w_smem = (int)(budget / (sm + 1024));
        int w_regs = depth_reg_warps(C, H, RV);
        int w = std::min(std::min(w_smem, w_regs), 32);
        if (w < 1) w = 1;
        double score = 1000.0 * std::min(w, 8) + 10.0 * RV + 100.0 * H + 0.1 * w;
        if (score > bestScore) { bestScore = score; best.H = H; best.RV = RV; best.reg_last = rl; best.blocks_per_sm = w; best.smem = sm; best.wpv_q16 = wpv; }
    }
    best.chunks = rslf_div_up(D, 32 * best.H);
    return best;
}

template <int C, int H, bool NONNEG, int RV>
static int launch_depth_t(rslf_ctx* ctx, const depth_args& a, const depth_plan& p)
{
    auto kern = depth_kernel<C, H, NONNEG, RV>;
    /* function attributes are per device: one flag per device a process drives (contexts on up to 64 devices) */
    static unsigned long long configured = 0;
    const unsigned long long dev_bit = 1ULL << (ctx->device & 63);
    if (!(configured & dev_bit)) {
        RSLF_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin));
        RSLF_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured |= dev_bit;
    }
    int occ = 0;
    RSLF_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, p.smem));
    if (occ < 1) occ = 1;
    kern<<<ctx->num_sm * occ, 32, p.smem, ctx->stream>>>(a);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    ctx->timing.depth_launches += 1;
    return RSLF_OK;
}

static int launch_depth(rslf_ctx* ctx, int C, bool nonneg, const depth_args& a, const depth_plan& p)
{
#define RSLF_DEPTH_CASE(CC, HH, RR)                                                          \
    if (C == CC && p.H == HH && p.RV == RR)                                                  \
        return nonneg ? launch_depth_t<CC, HH, true, RR>(ctx, a, p) : launch_depth_t<CC, HH, false, RR>(ctx, a, p);
    RSLF_DEPTH_CASE(1, 1, 0) RSLF_DEPTH_CASE(1, 1, 16) RSLF_DEPTH_CASE(1, 1, 32)
    RSLF_DEPTH_CASE(1, 2, 0) RSLF_DEPTH_CASE(1, 2, 16) RSLF_DEPTH_CASE(1, 4, 0)
    RSLF_DEPTH_CASE(3, 1, 0) RSLF_DEPTH_CASE(3, 1, 16) RSLF_DEPTH_CASE(3, 1, 32)
    RSLF_DEPTH_CASE(3, 2, 0) RSLF_DEPTH_CASE(3, 2, 16) RSLF_DEPTH_CASE(3, 4, 0)
#undef RSLF_DEPTH_CASE
    snprintf(ctx->err, sizeof(ctx->err), "no depth kernel variant for C=%d H=%d RV=%d", C, p.H, p.RV);
    return RSLF_ERR_UNSUPPORTED;
}
