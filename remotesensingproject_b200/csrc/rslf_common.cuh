/*
 * rslf_common.cuh — shared definitions of the sm_100a EPI depth library.
 *
 * Layouts (all dense, device memory):
 *   EPI stacks  [V][S][U][C] float32  (the reference's Vec<Mat> epis: V Mats of S x U x C,
 *                                      RSLightFields/src/rslf_io.cpp:194-227)
 *   maps        [S][V][U]   float32 / uint8 (masks are 0 / 255)
 * The whole library is compiled with -fmad=false: the reference rounds after
 * every OpenCV operation, so no multiply-add may be contracted.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/rslf_b200.h"

#define RSLF_MAX_LEVELS 24
#define RSLF_MAX_RANKS 16        /* ranks that can share peer memory (= RSLF_MAX_PEERS, k_balance.cuh) */

struct rslf_level {
    int V = 0, U = 0;            /* rows held by this rank, columns                           */
    int v0 = 0, Vtot = 0;        /* first global row of this rank, global rows of the level   */
    bool replicated = false;     /* multi-rank run: this (coarse) level is computed whole by every rank */
    float* raw = nullptr;        /* un-normalised stack of levels > 0 (float32, or uint8 for 8-bit input: raw8 aliases it); all Vtot rows */
    float* epi_full = nullptr;   /* normalised stack [Vtot][S][U][C]: EVERY rank holds all rows of every level */
    float* epi = nullptr;        /* = epi_full + v0 rows: the rows of this rank's block (the whole stack on one GPU) */
    float* ce = nullptr;         /* edge confidence            [S][V][U]                      */
    float* cd = nullptr;         /* disparity confidence       [S][V][U]                      */
    float* depth = nullptr;      /* best depth                 [S][V][U]                      */
    float* rbar = nullptr;       /* mean-shift radiance        [S][V][U][C]                   */
    float* dmin = nullptr;       /* per-pixel bounds           [S][V][U]  (levels > 0 / edit) */
    float* dmax = nullptr;
    uint8_t* emask = nullptr;    /* edge-confidence mask       [S][V][U]                      */
    uint8_t* remaining = nullptr;/* "still to compute" mask    [S][V][U]                      */
    uint8_t* valid = nullptr;    /* validity mask for bounds / fuse [S][V][U]                 */
    int* rowdark = nullptr;      /* [V][S] confident-and-dark pixels not yet painted, then [V] row sums */
    int* dark_lo = nullptr; int* dark_hi = nullptr;   /* [V][S] column range of those pixels (never shrinks: a bound) */
    float slope = 1.f;
    int nonneg = 1;              /* normalised stack has no negative value                    */
    bool have_bounds = false;
    bool bounds_const = false; float bounds_lo = 0.f, bounds_hi = 0.f;   /* dmin / dmax maps not filled: every pixel has these bounds */
    size_t cap_px = 0;           /* allocated S*V*U                                           */
    size_t cap_stack = 0;        /* allocated floats of epi (and raw, levels > 0)             */
    int C = 0;                   /* channels the maps were allocated for                      */
};

struct rslf_stage_clock {
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    struct span { int stage; size_t a, b; };
    std::vector<span> spans;
    std::vector<float> span_ms;
};

enum { ST_EDGE = 0, ST_DEPTH, ST_REDUCE, ST_MEDIAN, ST_PROP, ST_PYR, ST_COUNT };

struct rslf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sm = 0;
    size_t smem_optin = 0;
    char err[512] = {0};

    /* input */
    int V = 0, S = 0, U = 0, C = 0, cv_depth = RSLF_DEPTH_32F;
    float scale_factor = -1.f;
    void* raw_in = nullptr;      /* level-0 raw stack as uploaded (u8 or f32), owned unless borrowed: this rank's rows */
    bool raw_borrowed = false;
    size_t raw_cap = 0;
    bool have_input = false;
    float* open_ce = nullptr; uint8_t* open_mask = nullptr; uint8_t* open_tmp = nullptr; size_t open_cap = 0;   /* extended planes of a sharded opening */
    void* img_staging = nullptr; size_t img_staging_cap = 0;   /* rslf_cuda_upload_images: the image stack before the transposition */
    unsigned ring_next = 0;
    void* ring[2] = {nullptr, nullptr}; cudaEvent_t ring_ev[2] = {nullptr, nullptr};   /* pinned ring for pageable host images */
    cudaStream_t stream2 = nullptr; cudaEvent_t ev_img = nullptr;
    void* raw_full = nullptr;    /* multi-rank runs: all rows of the raw stack, gathered once per input */
    size_t raw_full_cap = 0;
    unsigned input_epoch = 1, raw_full_epoch = 0;
    float mm0[2] = {0.f, 0.f}; unsigned mm0_epoch = 0;   /* max / min of the uploaded float stack, per input */
    /* pipelined ingest (rslf_cuda_upload_epis_pipelined): level 0 was normalised / its edge confidence computed while the
     * stack was still arriving; valid for the input of that epoch (and, for the edge confidence, those parameters) */
    unsigned pre_norm_epoch = 0, pre_edge_epoch = 0; int pre_nonneg = 1; rslf_params pre_params;
    int v0 = 0, V_total = 0;     /* row shard: first global row, global rows (level 0)        */
    int row_starts[65] = {0};    /* level-0 shard table, world + 1 entries                    */
    bool have_shards = false;

    /* levels */
    int n_levels = 0;
    rslf_level lv[RSLF_MAX_LEVELS];

    /* scratch */
    int* items = nullptr;        /* compacted pixel list of a pass                            */
    int* count = nullptr;        /* list lengths, one slot per pass                           */
    int* queue = nullptr;        /* work-queue heads of the depth launches, one slot per pass */
    int4* rec = nullptr;         /* work records of the pass (k_balance.cuh); in the arena when peers read them */
    int4* rec_own = nullptr;     /* the cudaMalloc'ed record list of single-rank runs          */
    int* dev_err = nullptr;      /* raised by a kernel whose wait for a peer timed out         */
    rslf_decision* dlog = nullptr; int* dlog_count = nullptr; size_t dlog_cap = 0;   /* decision log (diagnostics) */
    unsigned long long* total_px = nullptr;  /* [0] computed pixels of sharded levels, [1] of replicated levels */
    float* filtered = nullptr;   /* selective-median output plane [V][U]                      */
    int* winner = nullptr;       /* propagation arbitration [S][V][U], INT_MAX when idle      */
    int* arrive = nullptr;       /* per-item chunk arrival counters                           */
    void* partials = nullptr;    /* per-(item, chunk) partial argmax records                  */
    size_t partials_cap = 0;
    int* nearest_l = nullptr;    /* bounds scratch [S][V][U]                                  */
    int* nearest_r = nullptr;
    float* minmax = nullptr;     /* [0]=max, [1]=min of a stack                               */
    float* fuse_a = nullptr; float* fuse_b = nullptr;  /* fuse ping-pong maps [S][V][U]       */
    uint8_t* fuse_ma = nullptr; uint8_t* fuse_mb = nullptr;
    float* out_map = nullptr; uint8_t* out_valid = nullptr;
    size_t scratch_px = 0;       /* S*V*U the scratch was sized for                           */
    size_t scratch_plane = 0;    /* V*U the scratch was sized for                             */
    int scratch_world = 0;       /* ranks the scratch was sized for                           */
    size_t share_cap = 0;        /* pixels of a pass one rank may have to evaluate (arrive / partials)     */
    unsigned* rowwork = nullptr; size_t rowwork_cap = 0;   /* pixels evaluated per level-0 row of this rank (last run) */
    void* l2_flush = nullptr;
    /* row-sharded runs: gathered planes of line s_hat for the cross-row median, staging */
    float* g_depth = nullptr; float* g_colour = nullptr; uint8_t* g_mask = nullptr;
    void* g_send = nullptr; void* g_recv = nullptr; size_t g_stage_cap = 0; size_t g_plane_cap = 0;
    float* g_raw = nullptr; size_t g_raw_cap = 0;        /* gathered raw stack of a level (downsample halo) */
    float* g_map = nullptr; float* g_map2 = nullptr; uint8_t* g_mk = nullptr; uint8_t* g_mk2 = nullptr;
    size_t g_map_cap = 0;        /* gathered fuse maps (two slots), S*Vtot*U each */
    float* pile_depth_raw = nullptr;  /* 1D pile: unfiltered depth plane                      */
    unsigned* colour_hist = nullptr;  /* coloured maps: three 65536-bin key histograms (k_colour.cuh) */
    uint8_t* colour_lut = nullptr;    /* coloured maps: 256 x 3 colour table                  */

    /* results state */
    int last_kind = 0;           /* 1 = pile, 2 = depth2d, 3 = fine-to-coarse                 */
    int pile_s_hat = 0;

    /* timing */
    rslf_timing timing;
    int stage_timing = 1;
    int criterion = 0;           /* 0: edge confidence gates propagation / validity (default build), 1: disparity confidence (as intended) */
    int fast_math = 0;           /* contracted (FMA) mean shift in the tensor-memory depth kernel; default: exact */
    rslf_stage_clock clk;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;

    /* peer memory over NVLink (CUDA IPC, rslf_comm.cuh): this rank's arena, the mapped arenas of all ranks */
    char* arena = nullptr; char* arena_peer[RSLF_MAX_RANKS] = {nullptr};
    size_t arena_bytes = 0, arena_cap_rec = 0; int arena_U = 0, arena_C = 0, arena_S = 0;
    unsigned fuse_seq = 0;       /* sequence number of the last fusion that sent halo rows */
    char* p2p_up = nullptr; char* p2p_dn = nullptr;   /* arenas of the ranks above / below */
    size_t p2p_area = 0;         /* bytes of one halo area (2 rows of depth, colour, mask) */
    unsigned p2p_seq = 0; int p2p_state = 0;   /* 0 = not set up, 1 = ready, -1 = unavailable */
    unsigned bal_seq = 0;        /* sequence number of the last pass-balanced pass */
    int balance = 1;             /* pass-balanced depth kernel in row-sharded runs (RSLF_BALANCE=0: lock-step row blocks) */
    int* p2p_done = nullptr;     /* block completion counters: [0] push, [1] compact, [2] depth, [3] apply, [4] fuse halo */
    /* NCCL (resolved lazily with dlopen; see rslf_comm.cuh) */
    void* nccl_lib = nullptr;
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
};

#define RSLF_CUDA_TRY(ctx, call)                                                          \
    do {                                                                                  \
        cudaError_t _e = (call);                                                          \
        if (_e != cudaSuccess) {                                                          \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s -> %s", __FILE__,         \
                     __LINE__, #call, cudaGetErrorString(_e));                            \
            return RSLF_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

#define RSLF_TRY(call)                    \
    do {                                  \
        int _r = (call);                  \
        if (_r != RSLF_OK) return _r;     \
    } while (0)

static inline int rslf_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

/*
 * The reference compares float(sqrt(double sum of squares)) < eps
 * (rslf_types.cpp:86-91).  sqrt and the float rounding are monotone, so the
 * test equals  sum < T  for the smallest double T with float(sqrt(T)) >= eps.
 * Found by bisection over the (ordered) bit patterns of positive doubles.
 */
static inline double rslf_sq_threshold(float eps) {
    if (!(eps > 0.f)) return 0.0;          /* nothing is < eps */
    uint64_t lo = 0, hi = 0x7ff0000000000000ULL;   /* f(0) true, f(inf) false */
    while (hi - lo > 1) {
        uint64_t mid = lo + (hi - lo) / 2;
        double s; memcpy(&s, &mid, 8);
        bool lt = (float)__builtin_sqrt(s) < eps;
        if (lt) lo = mid; else hi = mid;
    }
    double T; memcpy(&T, &hi, 8);
    return T;
}

/* device helpers ---------------------------------------------------------- */
#ifdef __CUDACC__
/* norm<float>(x) = float(abs(x) * 1.73205080757) (rslf_types.cpp:80-84) */
__device__ __forceinline__ bool rslf_norm1_lt(float x, float eps) {
    return (float)((double)fabsf(x) * 1.73205080757) < eps;
}
/* norm<Vec3f> < eps via the squared-sum threshold (see rslf_sq_threshold) */
__device__ __forceinline__ bool rslf_norm3_lt(float x, float y, float z, double T) {
    double s = (double)x * (double)x;
    s = fma((double)y, (double)y, s);      /* the product of two floats is exact in double */
    s = fma((double)z, (double)z, s);
    return s < T;
}
template <int C>
__device__ __forceinline__ bool rslf_norm_diff_lt(const float* a, const float* b, float eps, double T) {
    if (C == 1) return rslf_norm1_lt(a[0] - b[0], eps);
    return rslf_norm3_lt(a[0] - b[0], a[1] - b[1], a[2] - b[2], T);
}
/* cv::BORDER_REFLECT_101 */
__device__ __forceinline__ int rslf_reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p : 2 * n - 2 - p;
    return p;
}
/* cv::BORDER_REFLECT */
__device__ __forceinline__ int rslf_reflect(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = (p < 0) ? -p - 1 : 2 * n - 1 - p;
    return p;
}
#endif
