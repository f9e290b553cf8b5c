/*
 * rslf_b200.h — C ABI of the B200-native EPI depth-estimation path.
 *
 * This is the drop-in boundary behind the reference's C++ depth-computer /
 * fine-to-coarse classes (RSLightFields).  Every entry point names the reference
 * interface it replaces (paths relative to /root/reference/RSLightFields).
 * Plain pointers and sizes only; no C++ / torch types cross this boundary.
 *
 * Conventions
 *   - all functions return 0 on success or a negative rslf error code
 *     (rslf_cuda_strerror turns it into text); there is NO CPU fallback:
 *     without a usable CUDA device every compute call fails with RSLF_ERR_CUDA.
 *   - dimension names follow the reference: s = view, v = image row, u = column,
 *     d = disparity hypothesis, C = channels (1 or 3).
 *   - host images use OpenCV's Mat layout: row-major, interleaved channels,
 *     `step` bytes between rows (cv::Mat::data / cv::Mat::step).
 *   - per-(s,v,u) maps are S planes of V x U, densely packed ([S][V][U]),
 *     float32 or uint8 (masks are 0 / 255 like OpenCV comparison results).
 *   - one host thread per rslf_ctx; a ctx owns one CUDA device + one stream.
 */
#ifndef RSLF_B200_H
#define RSLF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* OpenCV depth codes accepted by rslf_cuda_upload_epis (CV_8U = 0, CV_16U = 2, CV_32F = 5). */
#define RSLF_DEPTH_8U  0
#define RSLF_DEPTH_16U 2
#define RSLF_DEPTH_32F 5

enum {
    RSLF_OK             =  0,
    RSLF_ERR_CUDA       = -1,   /* CUDA runtime error; see rslf_cuda_last_error_text */
    RSLF_ERR_ARG        = -2,   /* invalid argument                                   */
    RSLF_ERR_STATE      = -3,   /* call order violated (e.g. run before upload)       */
    RSLF_ERR_UNSUPPORTED= -4,   /* valid in the reference, not implemented here       */
    RSLF_ERR_NCCL       = -5,   /* NCCL error                                         */
    RSLF_ERR_NOMEM      = -6,
    RSLF_ERR_PEER       = -7    /* a peer GPU of a row-sharded run did not answer in time  */
};

/*
 * Algorithm parameters.  Mirrors rslf::Depth1DParameters<T>
 * (include/rslf_depth_computation_core.hpp:66-142; defaults :16-31).
 * The interpolation class is always Interpolation1DLinear and the kernel class
 * BandwidthKernel(h) (core.hpp:76-78) — the only ones the reference selects.
 */
typedef struct rslf_params {
    float edge_score_threshold;      /* par_edge_score_threshold   0.02 */
    float line_score_threshold;      /* par_line_score_threshold   0.02 (unused: flag off) */
    float disp_score_threshold;      /* par_disp_score_threshold   0.01 (unused: flag off) */
    float raw_score_threshold;       /* par_raw_score_threshold    0    */
    int   mean_shift_max_iter;       /* par_mean_shift_max_iter    10   */
    int   edge_confidence_filter_size;   /* 9 */
    int   edge_confidence_opening_type;  /* cv::MORPH_ELLIPSE = 2 */
    int   edge_confidence_opening_size;  /* 1 => opening disabled (core.hpp:759) */
    int   median_filter_size;        /* par_median_filter_size     5    */
    float median_filter_epsilon;     /* 0.1  */
    float propagation_epsilon;       /* 0.1  */
    float slope_factor;              /* par_slope_factor 1.0 (set per pyramid level, ftc.hpp:139) */
    int   cut_shadows;               /* par_cut_shadows true */
    float shadow_level;              /* (float)(0.05 * 1.73205080757) */
    float kernel_h;                  /* _BANDWIDTH_KERNEL_PARAMETER 0.2 (kern.hpp:43) */
} rslf_params;

/* Fills *p with the reference defaults (core.hpp:16-31, 74-99). */
void rslf_params_default(rslf_params* p);

/* Per-call stage timings (CUDA events on the ctx stream) and work counters. */
typedef struct rslf_timing {
    float  ms_total;          /* whole run() on the device                      */
    float  ms_edge;           /* edge-confidence kernels                         */
    float  ms_depth;          /* sampling + mean-shift + score kernels           */
    float  ms_reduce;         /* argmax / confidence kernels                     */
    float  ms_median;         /* selective median                                */
    float  ms_propagate;      /* propagation                                     */
    float  ms_pyramid;        /* downsample + normalise + bounds + fuse + median */
    float  ms_h2d;            /* upload (host->device copies + normalisation)    */
    float  ms_d2h;            /* result download                                 */
    double computed_pixels;   /* sum over levels and passes of pixels evaluated  */
    double samples;           /* computed_pixels x D x S  (the BASELINE metric)  */
    double depth_launches;    /* launches of the dominant (mean-shift) kernel    */
    double kernel_launches;   /* all kernel launches of the call                 */
    int    levels;            /* pyramid levels run                              */
    int    passes;            /* s_hat passes run (all levels)                   */
} rslf_timing;

typedef struct rslf_ctx rslf_ctx;

/* One evaluated pixel of a run (diagnostics, rslf_cuda_set_decision_log): what the depth kernel decided for pixel
 * `pix` (= v * U + u in the level's image) of line s_hat of pyramid level `level` (core.hpp:630-657). */
typedef struct rslf_decision {
    int   level, s_hat, pix;
    int   index;            /* d*: first maximum of the score over the hypotheses                */
    float score;            /* its score                                                         */
    float rbar[3];          /* mean-shift radiance of that hypothesis                            */
    float disp_conf;        /* C_d                                                               */
    float disparity;        /* D[d*]                                                             */
    int   accepted;         /* score > par_raw_score_threshold                                   */
    int   reserved;
} rslf_decision;

/* ---- lifetime ------------------------------------------------------------ */
int  rslf_cuda_create(int device, rslf_ctx** out);
void rslf_cuda_destroy(rslf_ctx* ctx);
const char* rslf_cuda_strerror(int code);
/* Text of the last CUDA/NCCL failure seen by this ctx ("" if none). */
const char* rslf_cuda_last_error_text(const rslf_ctx* ctx);
int  rslf_cuda_last_timing(const rslf_ctx* ctx, rslf_timing* out);
/* ABI version, bumped on any signature change. */
int  rslf_cuda_abi_version(void);

/* Which confidence gates the propagation sources and the validity maps.
 * 0 (default) = the edge confidence: the reference as written — its `#elseif` lines (rslf_depth_computation_core.hpp:1099,
 *     rslf_depth_computation.hpp:903) are not preprocessor directives, so the default build always takes the `#else` branch;
 * 1 = the disparity confidence C_d > par_disp_score_threshold: the reference AS INTENDED with
 *     -D_USE_DISP_CONFIDENCE_SCORE once `#elseif` reads `#elif` (core.hpp:1097-1098, dc.hpp:901-902; report section
 *     "Disparity confidence score").  As written that macro does not compile.
 * 2 (line confidence, core.hpp:1032-1081) -> RSLF_ERR_UNSUPPORTED. */
int  rslf_cuda_set_confidence_criterion(rslf_ctx* ctx, int criterion);

/* ---- input ---------------------------------------------------------------
 * Replaces the input handling of the computers' constructors
 * (rslf_depth_computation.hpp:425-477 Depth1DComputer_pile,
 *  :651-704 Depth2DComputer; rslf_fine_to_coarse.hpp:103-159 FineToCoarse):
 * deep-copies the V EPIs (each S rows x U cols x C channels) to the device and
 * normalises to float32: 8U -> x*(1/255); otherwise x*(1/scale) with
 * scale = epi_scale_factor if >= 0 else the global max over the stack.
 * epi_ptrs[v] = cv::Mat::data of EPI v, row_step_bytes = cv::Mat::step.
 */
int rslf_cuda_upload_epis(rslf_ctx* ctx, const void* const* epi_ptrs,
                          int V, int S, int U, int C, int cv_depth,
                          size_t row_step_bytes, float epi_scale_factor);
/* Ingest fused with the input handling and overlapped with the first stage: the EPIs are uploaded in
 * chunks; while the next chunk crosses PCIe, the chunk that has arrived is normalised to float32
 * (rslf_depth_computation.hpp:463-477) and, if `params` is given, its edge confidence and mask are computed for all
 * views (rslf_depth_computation_core.hpp:901-931) on a second stream.  The next Depth2DComputer / FineToCoarse run on
 * this input with the same edge parameters starts from those maps.  Falls back to rslf_cuda_upload_epis when the scale
 * is only known after the upload (non-8-bit input with epi_scale_factor < 0), for row-sharded contexts and when the
 * morphological opening is on. */
int rslf_cuda_upload_epis_pipelined(rslf_ctx* ctx, const void* const* epi_ptrs,
                                    int V, int S, int U, int C, int cv_depth,
                                    size_t row_step_bytes, float epi_scale_factor, const rslf_params* params);
/* Same, but the raw stack already lives on this ctx's device as one dense
 * [V][S][U][C] array (used to time the path with HBM-resident inputs). */
int rslf_cuda_set_epis_device(rslf_ctx* ctx, const void* d_epis,
                              int V, int S, int U, int C, int cv_depth,
                              float epi_scale_factor);
/* Device-side EPI building: replaces rslf::build_epis_from_imgs
 * (src/rslf_io.cpp:194-227).  img_ptrs[s] = one V x U x C image per view. */
int rslf_cuda_upload_images(rslf_ctx* ctx, const void* const* img_ptrs,
                            int S, int V, int U, int C, int cv_depth,
                            size_t row_step_bytes, float epi_scale_factor);

/* ---- Depth1DComputer_pile<T>::run() --------------------------------------
 * (rslf_depth_computation.hpp:513-565): edge confidence of line s_hat,
 * per-pixel depth on confident pixels, selective median.  s_hat < 0 or >= S
 * selects floor(S/2) (dc.hpp:489-498).  Output pointers are host buffers
 * (V x U, dense); any of them may be NULL.  best_depth is the median-filtered
 * map (core.hpp:892 rebinds the member).  rbar is V x U x C.
 */
int rslf_cuda_depth1d_pile(rslf_ctx* ctx, float dmin, float dmax, int dim_d,
                           int s_hat, const rslf_params* params,
                           float* best_depth_vu, float* edge_conf_vu,
                           uint8_t* edge_mask_vu, float* disp_conf_vu,
                           float* rbar_vuc);
/* compute only / copy only halves of the call above */
int rslf_cuda_depth1d_pile_run(rslf_ctx* ctx, float dmin, float dmax, int dim_d,
                               int s_hat, const rslf_params* params);
int rslf_cuda_depth1d_pile_get(rslf_ctx* ctx, float* best_depth_vu,
                               float* edge_conf_vu, uint8_t* edge_mask_vu,
                               float* disp_conf_vu, float* rbar_vuc);
/* The unfiltered argmax depths of the last pile run: m_best_depth_v_u as it is
 * before core.hpp:881-892 replaces it by the median-filtered map.  A one-row pile read
 * back this way is Depth1DComputer<T>::run() (rslf_depth_computation.hpp:26-68, 319-363:
 * one EPI, one line, no median) — what the facade's Depth1DComputer does. */
int rslf_cuda_depth1d_pile_get_raw_depth(rslf_ctx* ctx, float* raw_depth_vu);

/* ---- Depth2DComputer<T>::run() -------------------------------------------
 * (rslf_depth_computation.hpp:748-805 -> core.hpp:901-1133): edge confidence on
 * all lines, then the sequential s_hat passes with propagation.  dmin_svu /
 * dmax_svu are optional per-(s,v,u) host maps (edit_dmin()/edit_dmax(),
 * dc.hpp:211-213); NULL means the constant dmin / dmax.
 * Outputs = the public members m_best_depth_s_v_u, m_edge_confidence_s_v_u,
 * m_edge_confidence_mask_s_v_u, m_disp_confidence_s_v_u, m_rbar_s_v_u
 * (dc.hpp:217-224); any may be NULL.
 */
int rslf_cuda_depth2d(rslf_ctx* ctx, float dmin, float dmax, int dim_d,
                      const rslf_params* params,
                      const float* dmin_svu, const float* dmax_svu,
                      float* best_depth_svu, float* edge_conf_svu,
                      uint8_t* edge_mask_svu, float* disp_conf_svu,
                      float* rbar_svuc);
int rslf_cuda_depth2d_run(rslf_ctx* ctx, float dmin, float dmax, int dim_d,
                          const rslf_params* params,
                          const float* dmin_svu, const float* dmax_svu);
int rslf_cuda_depth2d_get(rslf_ctx* ctx, float* best_depth_svu,
                          float* edge_conf_svu, uint8_t* edge_mask_svu,
                          float* disp_conf_svu, float* rbar_svuc);
/* Depth2DComputer::get_valid_depths_mask_s_v_u (dc.hpp:893-915). */
int rslf_cuda_depth2d_get_valid_mask(rslf_ctx* ctx, int accept_all,
                                     const rslf_params* params,
                                     uint8_t* valid_svu);

/* ---- FineToCoarse<T>::run() + get_results() ------------------------------
 * (rslf_fine_to_coarse.hpp:103-322, src/rslf_fine_to_coarse_core.cpp:14-135):
 * builds the pyramid (while V > 10 && U > 10 && level < max_pyr_depth), runs a
 * Depth2D computation per level fine -> coarse, derives per-pixel [dmin,dmax]
 * of the next level, fuses coarse -> fine and applies the 3x3 median.
 * out_map_svu: S x V x U float32; out_valid_svu: S x V x U uint8.
 */
int rslf_cuda_fine_to_coarse(rslf_ctx* ctx, float dmin, float dmax, int dim_d,
                             const rslf_params* params, int max_pyr_depth,
                             int accept_all_last_scale,
                             float* out_map_svu, uint8_t* out_valid_svu);
int rslf_cuda_fine_to_coarse_run(rslf_ctx* ctx, float dmin, float dmax, int dim_d,
                                 const rslf_params* params, int max_pyr_depth,
                                 int accept_all_last_scale);
int rslf_cuda_fine_to_coarse_get(rslf_ctx* ctx, float* out_map_svu,
                                 uint8_t* out_valid_svu);
/* Number of levels of the last fine-to-coarse run and the size of level p. */
int rslf_cuda_fine_to_coarse_level_dims(const rslf_ctx* ctx, int level,
                                        int* V, int* U);
/* Per-level Depth2D results of the last fine-to-coarse run (for parity tests;
 * the reference exposes them through m_computers[p]'s public members). */
int rslf_cuda_fine_to_coarse_get_level(rslf_ctx* ctx, int level,
                                       float* best_depth_svu, float* edge_conf_svu,
                                       uint8_t* edge_mask_svu, float* disp_conf_svu,
                                       float* dmin_svu, float* dmax_svu);

/* FineToCoarse::get_coloured_depth_maps(plots, cv_colormap, saturate) (rslf_fine_to_coarse.hpp:324-377) of the last
 * fine-to-coarse run: ImageConverter_uchar::fit on the fused map of view round(S / 2) (src/rslf_plot.cpp:66-98; the 2 %
 * and 98 % order statistics), copy_and_scale to uchar (:100-107), colour table, pixels black where invalid or (with
 * params->cut_shadows) darker than _SHADOW_NORMALIZED_LEVEL.  lut_bgr_256x3: the table cv::applyColorMap would use
 * (apply it to the ramp 0..255 once; OpenCV's tables are not embedded here).  out_bgr_svu3: S x V x U x 3 uint8.
 * fit_min_max (optional): the fitted range.  saturate = 0 (mean + 12 std) returns RSLF_ERR_UNSUPPORTED. */
int rslf_cuda_fine_to_coarse_get_coloured(rslf_ctx* ctx, const uint8_t* lut_bgr_256x3, int saturate,
                                          const rslf_params* params, uint8_t* out_bgr_svu3,
                                          double* fit_min_max);

/* ---- free functions of the reference (each one its own entry point) ------ */
/* rslf::downsample_EPIs (src/rslf_fine_to_coarse_core.cpp:14-60), float32 stacks.
 * in: dense [V][S][U][C]; out: dense [V2][S][U2][C], V2 = cvRound(V/2). */
int rslf_cuda_downsample_epis(rslf_ctx* ctx, const float* in_epis, int V, int S,
                              int U, int C, float* out_epis, int* V2, int* U2);
/* The same for CV_8U stacks, which stay 8-bit between pyramid levels: OpenCV's integer Gaussian
 * ([8,28,56,72,56,28,8]/256 twice, (acc + 2^15) >> 16) and integer half-size resize, bit-exact to cv2. */
int rslf_cuda_downsample_epis_u8(rslf_ctx* ctx, const uint8_t* in_epis, int V, int S,
                                 int U, int C, uint8_t* out_epis, int* V2, int* U2);
/* The same for CV_16U stacks (they stay 16-bit): the same integer Gaussian and integer half-size resize
 * ((a+b+c+d+2) >> 2: OpenCV's own code; an IPP build of OpenCV rounds the 2x2 ties to even instead). */
int rslf_cuda_downsample_epis_u16(rslf_ctx* ctx, const uint16_t* in_epis, int V, int S,
                                  int U, int C, uint16_t* out_epis, int* V2, int* U2);
/* rslf::fuse_disp_maps (src/rslf_fine_to_coarse_core.cpp:69-135).  disp_p[p] and
 * valid_p[p] are dense [S][V_p][U_p] host maps, finest first. */
int rslf_cuda_fuse_disp_maps(rslf_ctx* ctx, int levels, int S, const int* Vp,
                             const int* Up, const float* const* disp_p,
                             const uint8_t* const* valid_p,
                             float* out_map_svu, uint8_t* out_valid_svu);
/* The bound-propagation step of FineToCoarse<T>::run (rslf_fine_to_coarse.hpp:201-294):
 * per-pixel [dmin, dmax] of level p+1 (Vd x Ud) from the depths and validity of
 * level p (Vu x Uu).  dmin_map / dmax_map are in/out ([S][Vd][Ud], pre-filled
 * with the global bounds; pixels without two valid neighbours keep them). */
int rslf_cuda_set_bounds(rslf_ctx* ctx, const float* depth_up_svu,
                         const uint8_t* valid_up_svu, int S, int Vu, int Uu,
                         int Vd, int Ud, float* dmin_map_svu, float* dmax_map_svu);
/* rslf::selective_median_filter (core.hpp:663-718) on the uploaded EPIs. */
int rslf_cuda_selective_median(rslf_ctx* ctx, const float* src_vu,
                               const uint8_t* mask_vu, int s_hat, int size,
                               float epsilon, float* dst_vu);
/* rslf::compute_1D_edge_confidence_pile (core.hpp:728-770) for one line s. */
int rslf_cuda_edge_confidence(rslf_ctx* ctx, int s, const rslf_params* params,
                              float* edge_conf_vu, uint8_t* edge_mask_vu);

/* ---- multi-GPU (row sharding, one process per GPU) -----------------------
 * The reference's only parallel axis is the image row v (OpenMP `parallel for`,
 * core.hpp:743, :799, :1088).  Rows are split in contiguous blocks, rank r
 * holding the global rows [row_starts[r], row_starts[r+1]); every rank uploads
 * only its block and gets only its block of every result map.  Pyramid level p
 * is sharded while every boundary is a multiple of 2^p and every rank keeps
 * >= 2 rows; coarser levels are computed whole by every rank.  Exchanges on the
 * path: the two rows next to each block of the depth / mask / colour planes of
 * line s_hat for the cross-row selective median of each pass (core.hpp:698-709)
 * — peer-to-peer stores over NVLink (CUDA IPC), NCCL all-gather as fallback —,
 * the raw rows a level's 7x7 blur reads beyond the block (ftc_core.cpp:37),
 * the fused maps between fuse levels and the max of the input normalisation
 * (dc.hpp:442-460), all NCCL.
 */
int rslf_cuda_nccl_unique_id(void* id128 /* 128 bytes out */);
int rslf_cuda_comm_init(rslf_ctx* ctx, const void* id128, int rank, int world);
/* row_starts: n_ranks + 1 entries, row_starts[0] = 0, row_starts[n_ranks] = total rows. */
int rslf_cuda_set_row_shards(rslf_ctx* ctx, const int* row_starts, int n_ranks);

/* Pixels evaluated per image row of this rank's block during the last run (every pass and level, coarse
 * levels mapped back to level-0 rows): the work profile for balancing the row blocks of the next run. */
int rslf_cuda_get_row_work(rslf_ctx* ctx, unsigned* rows_out /* V entries */);

/* ---- measurement helpers -------------------------------------------------- */
/* Measured FP32 FADD/FMUL issue rate of this device in Gop/s (non-FMA, the
 * instruction mix of the mean-shift kernel); used as roofline denominator. */
int rslf_cuda_measure_fp32_peak(rslf_ctx* ctx, double* gops_nofma, double* gflops_fma);
/* The same separately rounded multiply / add mix issued as packed FP32x2 instructions (FFMA2 / FADD2). */
int rslf_cuda_measure_fp32x2_peak(rslf_ctx* ctx, double* gops_nofma_packed);
/* Host-side planning of the tensor-memory depth kernel for one pass (no device needed): how the S views of a
 * warp item are split over registers / tensor memory / shared memory and packed into staging rounds.
 * out[0..7] = {fits (0/1), register views, tensor-memory views, shared-memory views, registers at the high end
 * (0/1), staging rounds, shared-memory bytes per warp, warps per SM}; rounds (optional, 3 ints per round) =
 * {first 4-view block, number of blocks, destination kind 0 registers / 1 tensor memory / 2 shared memory};
 * blocks (optional, 2 ints per 4-view block) = {staging offset in floats inside its round, floats per staged view}. */
int rslf_plan_depth_tm(int S, int C, int D, int s_hat, float dmin, float dmax, float slope, size_t smem_limit,
                       int* out8, int* rounds, int* blocks);
/* Opt-in contracted arithmetic for the mean shift (core.hpp:577-625) of RGB stacks: fused multiply-adds instead of the
 * reference's separately rounded OpenCV operations (20 -> 12 instructions per view, hypothesis and iteration).  Results
 * are then NOT bit-identical to the reference: scores move by a few 1e-7 relative, which stays inside the tolerance the
 * port is specified to (same disparity index wherever the winning score margin exceeds 1e-5, 1e-4 relative on scores and
 * confidences); a pixel inside that margin may pick the neighbouring hypothesis and propagate it.  Default: off (exact).
 * The environment variable RSLF_FAST_MATH=1 sets it at context creation. */
int rslf_cuda_set_fast_math(rslf_ctx* ctx, int on);
/* ---- host images (cv::Mat storage) ------------------------------------------------------------------------
 * The reference's classes own S result Mats per map (rslf_depth_computation.hpp:170-228, rslf_fine_to_coarse.hpp:301-322).
 * These getters write plane s of a result straight into the caller's image s (pointer = cv::Mat::data, step =
 * cv::Mat::step), nullptr arrays / entries are skipped.  Page-locked destinations (rslf_host_alloc, cudaHostAlloc,
 * cudaHostRegister) are filled by one strided DMA per image; pageable ones (an ordinary cv::Mat) through a ring of
 * two pinned buffers, the host thread emptying one while the copy engine fills the other. */
int rslf_cuda_depth2d_get_mats(rslf_ctx* ctx, float* const* best_depth_s, float* const* edge_conf_s,
                               uint8_t* const* edge_mask_s, float* const* disp_conf_s, float* const* rbar_s,
                               size_t step_f32, size_t step_u8, size_t step_rbar);
int rslf_cuda_fine_to_coarse_get_mats(rslf_ctx* ctx, float* const* out_map_s, size_t map_step,
                                      uint8_t* const* out_valid_s, size_t valid_step);
/* Depth2DComputer::get_epis (rslf_depth_computation.hpp:207): the normalised float32 EPIs of the last run, V images
 * of S rows x U x C. */
int rslf_cuda_get_epis(rslf_ctx* ctx, float* const* epi_v, size_t step);
/* Page-locked host memory for image storage that is uploaded / downloaded repeatedly (full PCIe rate, asynchronous). */
void* rslf_host_alloc(size_t bytes);
void  rslf_host_free(void* p);

/* Host-only helper (no device needed): which part of a pass rank `rank` evaluates in the pass-balanced multi-GPU mode,
 * given the lengths of every rank's work list: first pixel and number of pixels in the concatenated lists, and how many
 * of them come from each owner (from_owner[world], may be NULL). */
int rslf_balance_share(const int* counts, int world, int rank, long long* first, int* count, int* from_owner);
/* Diagnostics: records every pixel decision of the following runs (up to `capacity` records; 0 switches the log
 * off) so that a test can replay them through the CPU oracle — the way the contracted (fast_math) mode is checked
 * against the tolerance it is specified to (tests/test_gpu_tolerance.py).  rslf_cuda_get_decision_log copies the
 * records of the last run out and returns their number in *count (clipped to max_records in out). */
int rslf_cuda_set_decision_log(rslf_ctx* ctx, size_t capacity);
int rslf_cuda_get_decision_log(rslf_ctx* ctx, rslf_decision* out, size_t max_records, size_t* count);
/* CUDA events inside a run: 0 = none (only ms_total is measured), 1 (default) = around the launches of the dominant
 * depth kernel (ms_depth, the roofline's denominator; two records per s_hat pass), 2 = around every stage (ms_edge ...
 * ms_pyramid; about twelve records per pass, which serialise the stream: +2 % on one GPU, +10 % on eight). */
int rslf_cuda_set_stage_timing(rslf_ctx* ctx, int level);
/* Diagnostics: the event spans of the last run in issue order (stage 0 edge, 1 depth, 2 reduce / compaction, 3 median,
 * 4 propagation, 5 pyramid; milliseconds): with level 2 one entry per stage and s_hat pass. */
int rslf_cuda_last_spans(const rslf_ctx* ctx, int* stage, float* ms, size_t max_spans, size_t* count);
/* Row-sharded runs: 1 (default) = the pixels of every s_hat pass are split evenly over the ranks whatever rows they
 * lie in (every rank holds the whole EPI stack, the result maps stay sharded by rows; work / result records travel
 * through peer memory over NVLink); 0 = lock-step row blocks: every rank evaluates the pixels of its own rows (the
 * reference's OpenMP axis, core.hpp:799).  Results are identical either way.  RSLF_BALANCE=0 sets it at creation. */
int rslf_cuda_set_balance(rslf_ctx* ctx, int on);
/* Writes >126 MB on the device so the next timed step starts with a cold L2. */
int rslf_cuda_flush_l2(rslf_ctx* ctx);
int rslf_cuda_sync(rslf_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* RSLF_B200_H */
