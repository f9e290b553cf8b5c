/*
 * k_depth_tm.cuh — the depth kernel with Blackwell TENSOR MEMORY as a third tier of radiance storage.
 *
 * Same arithmetic, same operation order and the same outputs as depth_kernel (k_depth.cuh; it replaces the same
 * reference code: compute_1D_depth_epi core.hpp:480-661, interp.hpp:155-193, kern.cpp:16-54).  What changes is
 * where the radiances of a warp item (32 hypotheses x S views x C channels, 38 KB for C3) live during the ten
 * mean-shift iterations.  depth_kernel has registers + shared memory, which caps residency at 8 warps per SM
 * (2 per scheduler) for C3; with only two warps a scheduler idles whenever both are in a TMA wait / conversion /
 * prologue phase (ncu: issue slots 78 % busy).  The 256 KB of tensor memory per SM are otherwise unused on this
 * path (no MMA), and tools/probes/tmem_probe.cu measured that tcgen05.ld feeds the mean-shift arithmetic as
 * well as LDS.128 does (12 warps/SM: 1.59e12 against 1.63e12 view-samples/s, issue peak 1.68e12).  So:
 *
 *   - one CTA of 12 warps per SM (384 threads, <= 170 registers), every warp an independent worker with its own
 *     shared-memory region and mbarrier, striding over the warp items exactly like depth_kernel's one-warp blocks;
 *   - the CTA allocates all 512 TMEM columns (tcgen05.alloc); warp w owns the lane quarter w % 4 — the only
 *     lanes its tcgen05.ld / tcgen05.st can reach — and the 128 columns of group w / 4;
 *   - per item the views are split by distance from s_hat: the farthest 16 views -> registers, the next
 *     TV <= 128 / (4 C) * 4 views -> tensor memory (column = 4 C * block + 4 c + view-in-block, so that one
 *     tcgen05.ld.32x32b.x4 fetches a channel of four consecutive views, the same shape as the LDS.128 path),
 *     the nearest views -> shared memory (their segments are the shortest, so the staging area is small);
 *   - staging is done in rounds that share one area: [register views] [tensor-memory views, as many rounds as
 *     the area needs] [shared-memory views, converted in place].  The host packs the rounds (depth_tm_layout,
 *     passed in the kernel arguments) so that host and device agree by construction.
 *
 * Sums over views still run in ascending s inside a lane (registers / tensor memory / shared memory in that
 * order, or reversed when the far end is the high end), so every score is bit-identical to depth_kernel's
 * and to the CPU reference.
 */
#pragma once
#include "k_depth.cuh"

#define DEPTH_TM_WARPS 12
#define DEPTH_TM_RV 16
#define DEPTH_TM_MAX_BLOCKS 64          /* views / 4 */
#define DEPTH_TM_MAX_ROUNDS 8
#define DEPTH_TM_COLS_PER_WARP 128

enum { TM_KIND_REG = 0, TM_KIND_TMEM = 1, TM_KIND_SMEM = 2 };

struct depth_tm_layout {
    int TV, SV;                                          /* views in tensor memory / shared memory (multiples of 4) */
    int reg_last;                                        /* 0: [reg][tmem][smem] in ascending s, 1: [smem][tmem][reg] */
    int reg_b0, tm_b0, sm_b0;                            /* first 4-view block of each storage class */
    int nrounds;
    unsigned char round_b0[DEPTH_TM_MAX_ROUNDS], round_nb[DEPTH_TM_MAX_ROUNDS], round_kind[DEPTH_TM_MAX_ROUNDS];
    unsigned short blk_off4[DEPTH_TM_MAX_BLOCKS];        /* staging offset of the block inside its round's area, in units of 4 floats */
    unsigned short blk_pitch[DEPTH_TM_MAX_BLOCKS];       /* floats per staged view of the block (multiple of 4) */
    int area_floats, meta_views, warp_bytes;
};

/* Host: packs the views of a pass into storage classes and staging rounds.  Returns false if the pass does not fit. */
static inline bool depth_tm_build(depth_tm_layout& L, int S, int C, int s_hat, int wpv_q16, size_t smem_limit)
{
    const int RV = DEPTH_TM_RV, W = 32;
    const int Spad = depth_padded_views(S);
    const int NB = Spad / DEPTH_UNR;
    if (NB > DEPTH_TM_MAX_BLOCKS || Spad < RV + DEPTH_UNR) return false;
    int tvmax = DEPTH_TM_COLS_PER_WARP / (DEPTH_UNR * C) * DEPTH_UNR;
    depth_tm_layout best; bool have = false;
    for (int rl = 0; rl < 2; ++rl) {
        depth_tm_layout T; memset(&T, 0, sizeof(T));
        T.reg_last = rl;
        T.TV = std::min(tvmax, Spad - RV);
        T.SV = Spad - RV - T.TV;
        const int nbR = RV / DEPTH_UNR, nbT = T.TV / DEPTH_UNR, nbS = T.SV / DEPTH_UNR;
        if (!rl) { T.reg_b0 = 0; T.tm_b0 = nbR; T.sm_b0 = nbR + nbT; }
        else { T.sm_b0 = 0; T.tm_b0 = nbS; T.reg_b0 = nbS + nbT; }
        auto maxseg = [&](int b) {
            int f = 4;
            for (int j = 0; j < DEPTH_UNR; ++j) { int g = depth_segment_floats(b * DEPTH_UNR + j, S, s_hat, wpv_q16, C); f = g > f ? g : f; }
            return f;
        };
        /* shared-memory round (in place: every view row at least as large as its radiance row) */
        int areaS = 0;
        for (int i = 0; i < nbS; ++i) {
            const int b = T.sm_b0 + i, p = std::max(C * W, maxseg(b));
            T.blk_off4[b] = (unsigned short)(areaS / 4); T.blk_pitch[b] = (unsigned short)p; areaS += DEPTH_UNR * p;
        }
        /* register round (converted in place, then copied to registers) */
        int areaR = 0;
        for (int i = 0; i < nbR; ++i) {
            const int b = T.reg_b0 + i, p = std::max(C * W, maxseg(b));
            T.blk_off4[b] = (unsigned short)(areaR / 4); T.blk_pitch[b] = (unsigned short)p; areaR += DEPTH_UNR * p;
        }
        int area = std::max(areaS, areaR);
        /* tensor-memory rounds: greedy packing into the area (grown if a single block needs more) */
        for (int i = 0; i < nbT; ++i) area = std::max(area, DEPTH_UNR * maxseg(T.tm_b0 + i));
        /* the area may as well be as large as twelve warps allow: fewer tensor-memory rounds, i.e. fewer TMA waits */
        {
            const int meta_cap = std::max(RV, std::max(T.TV, T.SV));
            const long long per_warp = ((long long)smem_limit - 64) / DEPTH_TM_WARPS - 16 - 16LL * meta_cap;
            const int cap = (int)std::min<long long>(per_warp / 4, 4LL * 65535) & ~3;
            if (cap > area) area = cap;
        }
        int nr = 0, meta = std::max(RV, T.SV);
        auto push_round = [&](int b0, int nb, int kind) {
            if (nb <= 0) return true;
            if (nr >= DEPTH_TM_MAX_ROUNDS) return false;
            T.round_b0[nr] = (unsigned char)b0; T.round_nb[nr] = (unsigned char)nb; T.round_kind[nr] = (unsigned char)kind; ++nr;
            meta = std::max(meta, nb * DEPTH_UNR);
            return true;
        };
        bool ok = true;
        auto tm_rounds = [&]() {
            int i = 0;
            while (i < nbT && ok) {
                int used = 0, first = i;
                while (i < nbT) {
                    const int b = T.tm_b0 + i, p = maxseg(b);
                    if (used + DEPTH_UNR * p > area) break;
                    T.blk_off4[b] = (unsigned short)(used / 4); T.blk_pitch[b] = (unsigned short)p; used += DEPTH_UNR * p; ++i;
                }
                ok = ok && push_round(T.tm_b0 + first, i - first, TM_KIND_TMEM);
            }
        };
        /* staging order = ascending blocks is not required; registers first keeps their live range simple */
        ok = ok && push_round(T.reg_b0, nbR, TM_KIND_REG);
        tm_rounds();
        ok = ok && push_round(T.sm_b0, nbS, TM_KIND_SMEM);
        if (!ok) continue;
        T.nrounds = nr; T.area_floats = area; T.meta_views = meta;
        T.warp_bytes = 16 + 16 * meta + 4 * area;
        T.warp_bytes = (T.warp_bytes + 15) & ~15;
        if (area / 4 > 65535) continue;
        if ((size_t)T.warp_bytes * DEPTH_TM_WARPS + 64 > smem_limit) continue;
        /* fewer staging rounds first, then the smaller footprint */
        if (!have || T.nrounds < best.nrounds || (T.nrounds == best.nrounds && T.warp_bytes < best.warp_bytes)) { best = T; have = true; }
    }
    if (!have) return false;
    if ((size_t)best.warp_bytes * DEPTH_TM_WARPS + 64 > smem_limit) return false;
    L = best;
    return true;
}

/* ---- tensor-memory primitives (PTX; SASS: UTCALLOC-family, STTM, LDTM) ---- */
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float& a, float& b, float& c, float& d)
{
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(taddr));
    a = __uint_as_float(r0); b = __uint_as_float(r1); c = __uint_as_float(r2); d = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)),
                 "r"(__float_as_uint(c)), "r"(__float_as_uint(d)) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

/*
 * Conversion of the staged views of one round, DEPTH_UNR views per step (see convert_rows in k_depth.cuh for the
 * arithmetic: core.hpp:550-552, interp.hpp:171-185).  TO_TMEM: the radiances of block i of the round go to the
 * tensor-memory columns tcol + 4 C i ...; otherwise to rows + 4 C 32 i (packed, in place: block i's destination only
 * overlaps staging areas of blocks <= i).  rb receives the radiances of view s_hat when the round holds it.
 */
template <int C, bool FALLBACK, bool TO_TMEM>
__device__ __forceinline__ void tm_convert_round(const depth_args& a, const int4* meta, float* rows, uint32_t tcol, int vbase, int nviews,
                                                 int lane, long long row0, float ufl, float Um1f, float Dv, int& cardi, float (&rb)[C][1])
{
#pragma unroll 1
    for (int rb0 = 0; rb0 < nviews; rb0 += DEPTH_UNR) {
        float val[DEPTH_UNR][C];
#pragma unroll
        for (int j = 0; j < DEPTH_UNR; ++j) {
            const int s = vbase + rb0 + j;
            const int4 m = meta[rb0 + j];
            const float* row = rows + m.z;
            const float k = (float)(a.s_hat - s);
            float I = k * Dv;
            I = I * a.slope;
            I = I + ufl;
            const float fl = floorf(I);
            const int i0 = (int)fl;
            const int ne = (I != fl) ? 1 : 0;
            const bool ok = (I >= 0.f) && (I <= Um1f) && (s < a.S);
            const float t = I - fl;
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
                val[j][c] = ok ? (p + q) : RSLF_RAD_SENTINEL;
            }
            cardi += ok ? 1 : 0;
        }
        /* r_bar <- radiances of view s_hat (core.hpp:577) */
        const int js = a.s_hat - (vbase + rb0);
        if (js >= 0 && js < DEPTH_UNR) {
#pragma unroll
            for (int j = 0; j < DEPTH_UNR; ++j)
                if (j == js) {
#pragma unroll
                    for (int c = 0; c < C; ++c) rb[c][0] = val[j][c];
                }
        }
        if (TO_TMEM) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                tmem_st4(tcol + (uint32_t)((rb0 / DEPTH_UNR) * DEPTH_UNR * C + c * DEPTH_UNR), val[0][c], val[1][c], val[2][c], val[3][c]);
        } else {
            __syncwarp();                                               /* every lane has read the segments */
            float* blk = rows + (rb0 / DEPTH_UNR) * (DEPTH_UNR * C * 32);
#pragma unroll
            for (int c = 0; c < C; ++c)
                *reinterpret_cast<float4*>(blk + (c * 32 + lane) * DEPTH_UNR) = make_float4(val[0][c], val[1][c], val[2][c], val[3][c]);
        }
    }
}

template <int C, bool NONNEG, bool FAST>
__global__ void __launch_bounds__(32 * DEPTH_TM_WARPS, 1)
depth_kernel_tm(const depth_args a, const depth_tm_layout L)
{
    extern __shared__ float4 smem_raw[];
    __shared__ uint32_t tmem_base_s;
    constexpr int H = 1, RV = DEPTH_TM_RV, W = 32;
    constexpr int BLK = DEPTH_UNR * C * 32;                 /* floats per packed radiance block in shared memory */
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = a.S, U = a.U, D = a.D;

    /* all 512 columns for this CTA (it is alone on its SM); the allocating warp also frees them */
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    /* this warp's window: lane quarter wid % 4 (bits 31:16 of a TMEM address), column group wid / 4 */
    const uint32_t tm = tmem_base_s + ((uint32_t)(32 * (wid & 3)) << 16) + (uint32_t)(DEPTH_TM_COLS_PER_WARP * (wid >> 2));

    char* base = reinterpret_cast<char*>(smem_raw) + (size_t)wid * L.warp_bytes;
    const unsigned bar = smem_u32(base);
    int4* meta = reinterpret_cast<int4*>(base + 16);
    float* rows = reinterpret_cast<float*>(meta + L.meta_views);
    if (lane == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0;

    const float inv = a.inv;
    const f32x2 NZ = pk2(a.negzero, a.negzero);
    const float Um1f = (float)(U - 1);
    const int nbT = L.TV / DEPTH_UNR, nbS = L.SV / DEPTH_UNR;

    /* warp items = (record, chunk) pairs claimed from the launch's work queue: a pass with few items spreads over the
     * whole GPU, and the void items of flat pixels (below) cost nothing.  Balanced mode: this rank's share of the pass,
     * records fetched from their owners (k_balance.cuh). */
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
        /* dmin == dmax: all D hypotheses are the same EPI line (see depth_kernel): chunk 0 evaluates it, the others are void */
        const bool flat = (a.chunks > 1) && (dmin == dmax);
        if (flat && chunk > 0) { w = __shfl_sync(0xffffffffu, w_next, 0); continue; }
        const int v = pix / U, u = pix - v * U;
        const int dbase = chunk * W + lane;
        const float uf = (float)u;
        const float ufl = (dbase < D) ? uf : __int_as_float(0x7fc00000);   /* NaN for padding hypotheses */
        float Dv, Dlo, Dhi;
        {
            const float range = dmax - dmin;
            const float den = (float)(D - 1);
            float t = (float)dbase * range; t = t / den; Dv = dmin + t;
            t = (float)(chunk * W) * range; t = t / den; Dlo = dmin + t;
            t = (float)(min(chunk * W + W, D) - 1) * range; t = t / den; Dhi = dmin + t;
        }
        const long long row0 = (long long)v * S * U;
        int cardi = 0;
        float rb[C][H];
#pragma unroll
        for (int c = 0; c < C; ++c) rb[c][0] = 0.f;
        float rr[RV][C][H];

        /* ---- staging rounds ---- */
#pragma unroll 1
        for (int round = 0; round < L.nrounds; ++round) {
            const int b0 = L.round_b0[round], nb = L.round_nb[round], kind = L.round_kind[round];
            const int vbase = b0 * DEPTH_UNR, nviews = nb * DEPTH_UNR;
            bool any_cut = false;
            __syncwarp();
            {
                unsigned bytes = 0, cut = 0;
                for (int r = lane; r < nviews; r += 32) {
                    int lo, hi;
                    view_span(a, vbase + r, uf, Dlo, Dhi, lo, hi);
                    const int b = b0 + r / DEPTH_UNR;
                    const int pitch = L.blk_pitch[b];
                    int4 m; m.x = 0; m.y = 0; m.z = 4 * (int)L.blk_off4[b] + (r % DEPTH_UNR) * pitch; m.w = 0;
                    if (lo <= hi) {
                        const int mis = ((int)((row0 + (long long)(vbase + r) * U + lo) & 3LL) * C) & 3;
                        const int want = ((hi + 1 - lo) * C + mis + 3) & ~3;
                        m.x = mis - lo * C; m.y = min(want, pitch); m.w = want > pitch;
                        bytes += 4u * (unsigned)m.y; cut |= (unsigned)m.w;
                    }
                    meta[r] = m;
                }
                any_cut = __any_sync(0xffffffffu, cut != 0);
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, off);
                if (lane == 0) mbar_arrive_expect_tx(bar, bytes);
                for (int r = lane; r < nviews; r += 32) {
                    const int4 m = meta[r];
                    if (m.y > 0)
                        tma_bulk_g2s(smem_u32(rows + m.z), a.epi + ((row0 + (long long)(vbase + r) * U) * C - m.x), 4u * (unsigned)m.y, bar);
                }
            }
            __syncwarp();                                   /* the staging records are read by every lane below */
            mbar_wait(bar, phase);
            phase ^= 1u;
            if (kind == TM_KIND_TMEM) {
                const uint32_t tcol = tm + (uint32_t)((b0 - L.tm_b0) * DEPTH_UNR * C);
                if (any_cut) tm_convert_round<C, true, true>(a, meta, rows, tcol, vbase, nviews, lane, row0, ufl, Um1f, Dv, cardi, rb);
                else tm_convert_round<C, false, true>(a, meta, rows, tcol, vbase, nviews, lane, row0, ufl, Um1f, Dv, cardi, rb);
                tmem_wait_st();
            } else {
                if (any_cut) tm_convert_round<C, true, false>(a, meta, rows, 0u, vbase, nviews, lane, row0, ufl, Um1f, Dv, cardi, rb);
                else tm_convert_round<C, false, false>(a, meta, rows, 0u, vbase, nviews, lane, row0, ufl, Um1f, Dv, cardi, rb);
                __syncwarp();
                if (kind == TM_KIND_REG) {
#pragma unroll
                    for (int b = 0; b < RV / DEPTH_UNR; ++b)
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const float4 t = *reinterpret_cast<const float4*>(rows + b * BLK + (c * 32 + lane) * DEPTH_UNR);
                            rr[b * DEPTH_UNR + 0][c][0] = t.x; rr[b * DEPTH_UNR + 1][c][0] = t.y;
                            rr[b * DEPTH_UNR + 2][c][0] = t.z; rr[b * DEPTH_UNR + 3][c][0] = t.w;
                        }
                }
            }
        }
        const float card = (float)cardi;

        /* ---- mean shift (core.hpp:577-610), views in ascending s: [registers][tensor memory][shared memory] or reversed ---- */
        float sK[H];
        sK[0] = 0.f;
        for (int it = 0; it < a.iters; ++it) {
            float sR[C][H];
            sK[0] = 0.f;
#pragma unroll
            for (int c = 0; c < C; ++c) sR[c][0] = 0.f;
            float ba[DEPTH_UNR][C][H], bb[DEPTH_UNR][C][H];
            auto lds_block = [&](const float* blk, float (&dst)[DEPTH_UNR][C][H]) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float4 t = *reinterpret_cast<const float4*>(blk + c * 32 * DEPTH_UNR);
                    dst[0][c][0] = t.x; dst[1][c][0] = t.y; dst[2][c][0] = t.z; dst[3][c][0] = t.w;
                }
            };
            auto ldtm_block = [&](int slot, float (&dst)[DEPTH_UNR][C][H]) {
#pragma unroll
                for (int c = 0; c < C; ++c)
                    tmem_ld4(tm + (uint32_t)(slot * DEPTH_UNR * C + c * DEPTH_UNR), dst[0][c][0], dst[1][c][0], dst[2][c][0], dst[3][c][0]);
            };
#pragma unroll 1
            for (int part = 0; part < 3; ++part) {
                const int kind = L.reg_last ? (2 - part) : part;
                if (kind == TM_KIND_REG) {
#pragma unroll
                    for (int j = 0; j < RV; j += 2) ms_accumulate_pair<C, H, NONNEG, FAST>(rr[j], rr[j + 1], rb, inv, NZ, sR, sK);
                } else if (kind == TM_KIND_TMEM) {
                    if (nbT > 0) {
                        ldtm_block(0, ba);
                        tmem_wait_ld();
                        int slot = 0;
#pragma unroll 1
                        for (int n = nbT >> 1; n > 0; --n) {
                            ldtm_block(slot + 1, bb);
#pragma unroll
                            for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG, FAST>(ba[j], ba[j + 1], rb, inv, NZ, sR, sK);
                            slot += 2;
                            tmem_wait_ld();
                            ldtm_block(slot < nbT ? slot : nbT - 1, ba);
#pragma unroll
                            for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG, FAST>(bb[j], bb[j + 1], rb, inv, NZ, sR, sK);
                            tmem_wait_ld();
                        }
                        if (nbT & 1) {
#pragma unroll
                            for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG, FAST>(ba[j], ba[j + 1], rb, inv, NZ, sR, sK);
                        }
                    }
                } else {
                    if (nbS > 0) {
                        const float* const p0 = rows + lane * DEPTH_UNR;
                        const float* const plast = p0 + (nbS - 1) * BLK;
                        const float* p = p0;
                        lds_block(p0, ba);
#pragma unroll 1
                        for (int n = nbS >> 1; n > 0; --n) {
                            lds_block(p + BLK, bb);
#pragma unroll
                            for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG, FAST>(ba[j], ba[j + 1], rb, inv, NZ, sR, sK);
                            p += 2 * BLK;
                            lds_block(p < plast ? p : plast, ba);
#pragma unroll
                            for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG, FAST>(bb[j], bb[j + 1], rb, inv, NZ, sR, sK);
                        }
                        if (nbS & 1) {
#pragma unroll
                            for (int j = 0; j < DEPTH_UNR; j += 2) ms_accumulate_pair<C, H, NONNEG, FAST>(ba[j], ba[j + 1], rb, inv, NZ, sR, sK);
                        }
                    }
                }
            }
            /* r_bar = sum_rK / sum_K (x / 0 = 0 as in OpenCV 3), then max(., 0) (core.hpp:606-609) */
            const float den = sK[0];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float q = (den != 0.f) ? (sR[c][0] / den) : 0.f;
                rb[c][0] = (q > 0.f) ? q : 0.f;
            }
        }
        /* ---- score, warp argmax (lowest index on ties), chunk merge, outputs: as in depth_kernel (core.hpp:616-657) ---- */
        float best = -1.f; int bidx = 0x7fffffff; double sum = 0.0;
        float bdv = 0.f, brb[C];
#pragma unroll
        for (int c = 0; c < C; ++c) brb[c] = 0.f;
        if (dbase < D) {
            const float q = sK[0] / card;
            const float sc = (q > 0.f) ? q : 0.f;
            sum += (double)sc;
            if (sc > best) {
                best = sc; bidx = dbase; bdv = Dv;
#pragma unroll
                for (int c = 0; c < C; ++c) brb[c] = rb[c][0];
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
    if (a.bal.n) __threadfence_system();
    /* every tcgen05.ld / st of the CTA is complete (each warp waited on its own); free the columns */
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "n"(512));
    if (a.bal.n) bal_last_block_signal(a.bal.blocks_done, a.bal.done, a.bal.n, a.bal.seq, 0u);
}

template <int C, bool NONNEG, bool FAST>
static int launch_depth_tm_t(rslf_ctx* ctx, const depth_args& a, const depth_tm_layout& L)
{
    auto kern = depth_kernel_tm<C, NONNEG, FAST>;
    /* function attributes are per device: one flag per device a process drives */
    static unsigned long long configured = 0;
    const unsigned long long dev_bit = 1ULL << (ctx->device & 63);
    if (!(configured & dev_bit)) {
        /* the kernel also has 16 bytes of static shared memory (the TMEM base address) */
        RSLF_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smem_optin - 64));
        RSLF_CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured |= dev_bit;
    }
    const size_t smem = (size_t)L.warp_bytes * DEPTH_TM_WARPS;
    kern<<<ctx->num_sm, 32 * DEPTH_TM_WARPS, smem, ctx->stream>>>(a, L);
    RSLF_CUDA_TRY(ctx, cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    ctx->timing.depth_launches += 1;
    return RSLF_OK;
}

/* fast: the contracted (FMA) mean shift, see ms_accumulate; the default is the reference's separately rounded arithmetic */
static int launch_depth_tm(rslf_ctx* ctx, int C, bool nonneg, const depth_args& a, const depth_tm_layout& L, bool fast = false)
{
    if (C == 3) {
        if (fast) return nonneg ? launch_depth_tm_t<3, true, true>(ctx, a, L) : launch_depth_tm_t<3, false, true>(ctx, a, L);
        return nonneg ? launch_depth_tm_t<3, true, false>(ctx, a, L) : launch_depth_tm_t<3, false, false>(ctx, a, L);
    }
    if (C == 1) {
        if (fast) return nonneg ? launch_depth_tm_t<1, true, true>(ctx, a, L) : launch_depth_tm_t<1, false, true>(ctx, a, L);
        return nonneg ? launch_depth_tm_t<1, true, false>(ctx, a, L) : launch_depth_tm_t<1, false, false>(ctx, a, L);
    }
    snprintf(ctx->err, sizeof(ctx->err), "no tensor-memory depth kernel for C=%d", C);
    return RSLF_ERR_UNSUPPORTED;
}
