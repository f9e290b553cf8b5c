/*
 * ref_driver.cpp — C entry points around the REFERENCE'S OWN classes and functions, compiled from the
 * reference's sources where they lie (/root/reference/RSLightFields/{include,src}) against the stand-in OpenCV
 * headers of oracle/cvshim.  Output: oracle/_ref/librslf_ref.so (oracle/Makefile, target `ref`).
 *
 * TEST INFRASTRUCTURE ONLY: loaded by oracle/ref.py for tests/test_oracle_vs_reference.py (which pins the
 * restated oracle against the reference's own control flow), for the golden fixtures of tests/golden/ref_*.npz
 * and for the CPU legs of bench.py.  No reference source is copied into this repository: this file only calls
 *   rslf::Depth1DComputer<T>       (include/rslf_depth_computation.hpp:26-68, 254-363)
 *   rslf::Depth1DComputer_pile<T>  (include/rslf_depth_computation.hpp:93-143, 424-565)
 *   rslf::Depth2DComputer<T>       (include/rslf_depth_computation.hpp:166-229, 652-915)
 *   rslf::FineToCoarse<T>          (include/rslf_fine_to_coarse.hpp:26-81, 103-322)
 *   rslf::downsample_EPIs, rslf::fuse_disp_maps (src/rslf_fine_to_coarse_core.cpp:14-135)
 *   rslf::compute_1D_edge_confidence_pile, rslf::selective_median_filter (include/rslf_depth_computation_core.hpp:663-770)
 * The result members of the computers are private in the reference; they are read here with the usual
 * `#define private public` of white-box tests, which changes no behaviour.
 *
 * Layouts at this boundary: stacks [V][S][U][C] (the reference's Vec<Mat> epis, one S x U matrix per row v),
 * maps [S][V][U] (1D pile: [V][U]).
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>

#include <opencv2/core/core.hpp>
#include <opencv2/imgproc/imgproc.hpp>

#define private public
#define protected public
#include <rslf_fine_to_coarse.hpp>
#undef private
#undef protected

#include "../include/rslf_b200.h"

namespace {

using rslf::Mat;
template <typename T> using RVec = rslf::Vec<T>;

/* the reference prints progress to std::cout; keep the test logs quiet */
struct QuietCout {
    std::streambuf* old;
    std::ostringstream sink;
    QuietCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~QuietCout() { std::cout.rdbuf(old); }
};

/* Vec<Mat> epis from a dense [V][S][U][C] stack (copied: the computers keep / convert their own copies anyway) */
RVec<Mat> make_epis(const void* raw, int cv_depth, int V, int S, int U, int C)
{
    const int type = CV_MAKETYPE(cv_depth, C);
    const size_t esz = (cv_depth == CV_8U ? 1 : cv_depth == CV_16U ? 2 : 4) * (size_t)C;
    RVec<Mat> epis;
    for (int v = 0; v < V; ++v) {
        Mat m(S, U, type);
        std::memcpy(m.data, (const uint8_t*)raw + (size_t)v * S * U * esz, (size_t)S * U * esz);
        epis.push_back(m);
    }
    return epis;
}

template <typename T>
void fill_params(rslf::Depth1DParameters<T>& q, const rslf_params* P)
{
    if (!P) return;
    q.par_edge_score_threshold = P->edge_score_threshold;
    q.par_line_score_threshold = P->line_score_threshold;
    q.par_disp_score_threshold = P->disp_score_threshold;
    q.par_raw_score_threshold = P->raw_score_threshold;
    q.par_mean_shift_max_iter = (float)P->mean_shift_max_iter;
    q.par_edge_confidence_filter_size = P->edge_confidence_filter_size;
    q.par_edge_confidence_opening_type = P->edge_confidence_opening_type;
    q.par_edge_confidence_opening_size = P->edge_confidence_opening_size;
    q.par_median_filter_size = P->median_filter_size;
    q.par_median_filter_epsilon = P->median_filter_epsilon;
    q.par_propagation_epsilon = P->propagation_epsilon;
    q.par_slope_factor = P->slope_factor;
    q.par_cut_shadows = P->cut_shadows != 0;
    q.par_shadow_level = P->shadow_level;
    delete q.par_kernel_class;
    q.par_kernel_class = new rslf::BandwidthKernel<T>(P->kernel_h);
}

void put_f32(const Mat& m, float* dst, int C = 1)
{
    if (!dst) return;
    for (int y = 0; y < m.rows; ++y) std::memcpy(dst + (size_t)y * m.cols * C, m.ptr<float>(y), (size_t)m.cols * C * sizeof(float));
}
void put_u8(const Mat& m, uint8_t* dst)
{
    if (!dst) return;
    for (int y = 0; y < m.rows; ++y) std::memcpy(dst + (size_t)y * m.cols, m.ptr<uchar>(y), (size_t)m.cols);
}
void put_f32_stack(const RVec<Mat>& ms, float* dst, int C = 1)
{
    if (!dst) return;
    for (size_t s = 0; s < ms.size(); ++s) put_f32(ms[s], dst + s * (size_t)ms[s].rows * ms[s].cols * C, C);
}
void put_u8_stack(const RVec<Mat>& ms, uint8_t* dst)
{
    if (!dst) return;
    for (size_t s = 0; s < ms.size(); ++s) put_u8(ms[s], dst + s * (size_t)ms[s].rows * ms[s].cols);
}

template <typename T>
int pile_t(const void* raw, int cv_depth, int V, int S, int U, int C, float scale, float dmin, float dmax, int D, int s_hat,
           const rslf_params* P, float* best_depth, float* edge_conf, uint8_t* edge_mask, float* disp_conf, float* rbar)
{
    rslf::Depth1DParameters<T> q;
    fill_params(q, P);
    RVec<Mat> epis = make_epis(raw, cv_depth, V, S, U, C);
    rslf::Depth1DComputer_pile<T> comp(epis, dmin, dmax, D, s_hat, scale, q);
    comp.run();
    put_f32(comp.m_best_depth_v_u, best_depth);
    put_f32(comp.m_edge_confidence_v_u, edge_conf);
    put_u8(comp.m_edge_confidence_mask_v_u, edge_mask);
    put_f32(comp.m_disp_confidence_v_u, disp_conf);
    put_f32(comp.m_rbar_v_u, rbar, C);
    delete q.par_interpolation_class;
    delete q.par_kernel_class;
    return comp.get_s_hat();
}

/* Depth1DComputer: one EPI, one line, no median (dc.hpp:254-363) */
template <typename T>
int single_t(const void* raw, int cv_depth, int S, int U, int C, float scale, float dmin, float dmax, int D, int s_hat,
             const rslf_params* P, float* best_depth, float* edge_conf, uint8_t* edge_mask, float* disp_conf, float* rbar)
{
    rslf::Depth1DParameters<T> q;
    fill_params(q, P);
    RVec<Mat> epis = make_epis(raw, cv_depth, 1, S, U, C);
    rslf::Depth1DComputer<T> comp(epis[0], dmin, dmax, D, s_hat, scale, q);
    comp.run();
    put_f32(comp.m_best_depth_u, best_depth);
    put_f32(comp.m_edge_confidence_u, edge_conf);
    put_u8(comp.m_edge_confidence_mask_u, edge_mask);
    put_f32(comp.m_disp_confidence_u, disp_conf);
    put_f32(comp.m_rbar_u, rbar, C);
    delete q.par_interpolation_class;
    delete q.par_kernel_class;
    return comp.m_s_hat;
}

template <typename T>
void depth2d_t(const void* raw, int cv_depth, int V, int S, int U, int C, float scale, float dmin, float dmax, int D,
               const rslf_params* P, const float* dmin_svu, const float* dmax_svu, int accept_all, float* best_depth,
               float* edge_conf, uint8_t* edge_mask, float* disp_conf, float* rbar, uint8_t* valid)
{
    rslf::Depth1DParameters<T> q;
    fill_params(q, P);
    RVec<Mat> epis = make_epis(raw, cv_depth, V, S, U, C);
    rslf::Depth2DComputer<T> comp(epis, dmin, dmax, D, scale, q, false);
    if (dmin_svu && dmax_svu) {
        RVec<Mat>& mn = comp.edit_dmin();
        RVec<Mat>& mx = comp.edit_dmax();
        for (int s = 0; s < S; ++s)
            for (int v = 0; v < V; ++v) {
                std::memcpy(mn[s].ptr<float>(v), dmin_svu + ((size_t)s * V + v) * U, (size_t)U * sizeof(float));
                std::memcpy(mx[s].ptr<float>(v), dmax_svu + ((size_t)s * V + v) * U, (size_t)U * sizeof(float));
            }
    }
    comp.set_accept_all(accept_all != 0);
    comp.run();
    put_f32_stack(comp.m_best_depth_s_v_u, best_depth);
    put_f32_stack(comp.m_edge_confidence_s_v_u, edge_conf);
    put_u8_stack(comp.m_edge_confidence_mask_s_v_u, edge_mask);
    put_f32_stack(comp.m_disp_confidence_s_v_u, disp_conf);
    put_f32_stack(comp.m_rbar_s_v_u, rbar, C);
    if (valid) put_u8_stack(comp.get_valid_depths_mask_s_v_u(), valid);
    delete q.par_interpolation_class;
    delete q.par_kernel_class;
}

template <typename T>
int ftc_t(const void* raw, int cv_depth, int V, int S, int U, int C, float scale, float dmin, float dmax, int D,
          const rslf_params* P, int max_pyr_depth, int accept_all_last, float* out_map, uint8_t* out_valid,
          float* const* lv_depth, float* const* lv_ce, uint8_t* const* lv_mask, float* const* lv_cd,
          float* const* lv_dmin, float* const* lv_dmax, int* lv_V, int* lv_U, double* seconds_run)
{
    rslf::Depth1DParameters<T> q;
    fill_params(q, P);
    RVec<Mat> epis = make_epis(raw, cv_depth, V, S, U, C);
    rslf::FineToCoarse<T> ftc(epis, dmin, dmax, D, scale, q, max_pyr_depth, accept_all_last != 0);
    const auto t0 = std::chrono::steady_clock::now();
    ftc.run();
    RVec<Mat> maps, valids;
    ftc.get_results(maps, valids);
    const auto t1 = std::chrono::steady_clock::now();
    if (seconds_run) *seconds_run = std::chrono::duration<double>(t1 - t0).count();
    put_f32_stack(maps, out_map);
    put_u8_stack(valids, out_valid);
    const int L = (int)ftc.m_computers.size();
    for (int p = 0; p < L; ++p) {
        rslf::Depth2DComputer<T>* c = ftc.m_computers[p];
        if (lv_V) lv_V[p] = c->m_best_depth_s_v_u[0].rows;
        if (lv_U) lv_U[p] = c->m_best_depth_s_v_u[0].cols;
        if (lv_depth && lv_depth[p]) put_f32_stack(c->m_best_depth_s_v_u, lv_depth[p]);
        if (lv_ce && lv_ce[p]) put_f32_stack(c->m_edge_confidence_s_v_u, lv_ce[p]);
        if (lv_mask && lv_mask[p]) put_u8_stack(c->m_edge_confidence_mask_s_v_u, lv_mask[p]);
        if (lv_cd && lv_cd[p]) put_f32_stack(c->m_disp_confidence_s_v_u, lv_cd[p]);
        if (lv_dmin && lv_dmin[p]) put_f32_stack(c->edit_dmin(), lv_dmin[p]);
        if (lv_dmax && lv_dmax[p]) put_f32_stack(c->edit_dmax(), lv_dmax[p]);
    }
    delete q.par_interpolation_class;
    delete q.par_kernel_class;
    return L;
}

/* FineToCoarse ctor + run + get_coloured_depth_maps (ftc.hpp:324-377): ImageConverter_uchar::fit on the fused map of
 * view round(S / 2) (rslf_plot.cpp:66-98), copy_and_scale (:100-107), applyColorMap, invalid and shadow pixels black. */
template <typename T>
int ftc_coloured_t(const void* raw, int cv_depth, int V, int S, int U, int C, float scale, float dmin, float dmax, int D,
                   const rslf_params* P, const uint8_t* lut_bgr, int saturate, uint8_t* out_bgr_svu3)
{
    QuietCout quiet;
    rslf::Depth1DParameters<T> q;
    fill_params(q, P);
    RVec<Mat> epis = make_epis(raw, cv_depth, V, S, U, C);
    rslf::FineToCoarse<T> ftc(epis, dmin, dmax, D, scale, q);
    ftc.run();
    cv::cvshim_colormap_lut() = lut_bgr;
    RVec<Mat> plots;
    ftc.get_coloured_depth_maps(plots, cv::COLORMAP_JET, saturate != 0);
    cv::cvshim_colormap_lut() = nullptr;
    for (size_t s = 0; s < plots.size(); ++s)
        for (int v = 0; v < plots[s].rows; ++v)
            std::memcpy(out_bgr_svu3 + ((s * V + v) * (size_t)U) * 3, plots[s].ptr(v), (size_t)U * 3);
    delete q.par_interpolation_class;
    delete q.par_kernel_class;
    return (int)plots.size();
}

}  // namespace

extern "C" {

/* FineToCoarse::get_coloured_depth_maps with the colour table lut_bgr[256][3].  out: [S][V][U][3] uint8 (BGR). */
int ref_fine_to_coarse_coloured(const void* raw, int cv_depth, int V, int S, int U, int C, float scale, float dmin, float dmax,
                                int D, const rslf_params* P, const uint8_t* lut_bgr, int saturate, uint8_t* out_bgr_svu3)
{
    if (C == 1) return ftc_coloured_t<float>(raw, cv_depth, V, S, U, C, scale, dmin, dmax, D, P, lut_bgr, saturate, out_bgr_svu3);
    if (C == 3) return ftc_coloured_t<cv::Vec3f>(raw, cv_depth, V, S, U, C, scale, dmin, dmax, D, P, lut_bgr, saturate, out_bgr_svu3);
    return -1;
}

int ref_num_threads(void) { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }

/* Depth1DComputer_pile ctor + run.  raw: [V][S][U][C], cv_depth 0 (CV_8U), 2 (CV_16U) or 5 (CV_32F).  Returns the s_hat used. */
int ref_depth1d_pile(const void* raw, int cv_depth, int V, int S, int U, int C, float scale, float dmin, float dmax, int D,
                     int s_hat, const rslf_params* P, float* best_depth, float* edge_conf, uint8_t* edge_mask,
                     float* disp_conf, float* rbar)
{
    QuietCout q;
    if (C == 1) return pile_t<float>(raw, cv_depth, V, S, U, C, scale, dmin, dmax, D, s_hat, P, best_depth, edge_conf, edge_mask, disp_conf, rbar);
    if (C == 3) return pile_t<cv::Vec3f>(raw, cv_depth, V, S, U, C, scale, dmin, dmax, D, s_hat, P, best_depth, edge_conf, edge_mask, disp_conf, rbar);
    return -1;
}

/* Depth1DComputer ctor + run on ONE EPI raw: [S][U][C].  Outputs: U values (rbar: U x C).  Returns the s_hat used. */
int ref_depth1d(const void* raw, int cv_depth, int S, int U, int C, float scale, float dmin, float dmax, int D, int s_hat,
                const rslf_params* P, float* best_depth, float* edge_conf, uint8_t* edge_mask, float* disp_conf, float* rbar)
{
    QuietCout q;
    if (C == 1) return single_t<float>(raw, cv_depth, S, U, C, scale, dmin, dmax, D, s_hat, P, best_depth, edge_conf, edge_mask, disp_conf, rbar);
    if (C == 3) return single_t<cv::Vec3f>(raw, cv_depth, S, U, C, scale, dmin, dmax, D, s_hat, P, best_depth, edge_conf, edge_mask, disp_conf, rbar);
    return -1;
}

/* Depth2DComputer ctor (+ optional per-pixel bounds, set_accept_all) + run + get_valid_depths_mask_s_v_u */
int ref_depth2d(const void* raw, int cv_depth, int V, int S, int U, int C, float scale, float dmin, float dmax, int D,
                const rslf_params* P, const float* dmin_svu, const float* dmax_svu, int accept_all, float* best_depth,
                float* edge_conf, uint8_t* edge_mask, float* disp_conf, float* rbar, uint8_t* valid)
{
    QuietCout q;
    if (C == 1) depth2d_t<float>(raw, cv_depth, V, S, U, C, scale, dmin, dmax, D, P, dmin_svu, dmax_svu, accept_all, best_depth, edge_conf, edge_mask, disp_conf, rbar, valid);
    else if (C == 3) depth2d_t<cv::Vec3f>(raw, cv_depth, V, S, U, C, scale, dmin, dmax, D, P, dmin_svu, dmax_svu, accept_all, best_depth, edge_conf, edge_mask, disp_conf, rbar, valid);
    else return -1;
    return 0;
}

/* FineToCoarse ctor + run + get_results; per-level maps optional.  Returns the number of pyramid levels. */
int ref_fine_to_coarse(const void* raw, int cv_depth, int V, int S, int U, int C, float scale, float dmin, float dmax, int D,
                       const rslf_params* P, int max_pyr_depth, int accept_all_last, float* out_map, uint8_t* out_valid,
                       float* const* lv_depth, float* const* lv_ce, uint8_t* const* lv_mask, float* const* lv_cd,
                       float* const* lv_dmin, float* const* lv_dmax, int* lv_V, int* lv_U, double* seconds_run)
{
    QuietCout q;
    if (C == 1) return ftc_t<float>(raw, cv_depth, V, S, U, C, scale, dmin, dmax, D, P, max_pyr_depth, accept_all_last, out_map, out_valid, lv_depth, lv_ce, lv_mask, lv_cd, lv_dmin, lv_dmax, lv_V, lv_U, seconds_run);
    if (C == 3) return ftc_t<cv::Vec3f>(raw, cv_depth, V, S, U, C, scale, dmin, dmax, D, P, max_pyr_depth, accept_all_last, out_map, out_valid, lv_depth, lv_ce, lv_mask, lv_cd, lv_dmin, lv_dmax, lv_V, lv_U, seconds_run);
    return -1;
}

/* rslf::downsample_EPIs.  out: [V2][S][U2][C] of the input's depth; V2 / U2 are returned. */
int ref_downsample(const void* raw, int cv_depth, int V, int S, int U, int C, void* out, int* V2, int* U2)
{
    RVec<Mat> epis = make_epis(raw, cv_depth, V, S, U, C), down;
    rslf::downsample_EPIs(epis, down);
    const size_t esz = (cv_depth == CV_8U ? 1 : cv_depth == CV_16U ? 2 : 4) * (size_t)C;
    if (V2) *V2 = (int)down.size();
    if (U2) *U2 = down.empty() ? 0 : down[0].cols;
    if (out)
        for (size_t v = 0; v < down.size(); ++v)
            for (int s = 0; s < S; ++s)
                std::memcpy((uint8_t*)out + ((v * S + s) * (size_t)down[0].cols) * esz, down[v].ptr(s), (size_t)down[0].cols * esz);
    return 0;
}

/* rslf::fuse_disp_maps.  disp / valid: per level (finest first) [S][V_p][U_p]. */
int ref_fuse(int levels, int S, const int* Vp, const int* Up, const float* const* disp, const uint8_t* const* valid,
             float* out_map, uint8_t* out_valid)
{
    RVec<RVec<Mat> > dp(levels), vp(levels);
    for (int p = 0; p < levels; ++p)
        for (int s = 0; s < S; ++s) {
            Mat d(Vp[p], Up[p], CV_32FC1), m(Vp[p], Up[p], CV_8UC1);
            std::memcpy(d.data, disp[p] + (size_t)s * Vp[p] * Up[p], (size_t)Vp[p] * Up[p] * sizeof(float));
            std::memcpy(m.data, valid[p] + (size_t)s * Vp[p] * Up[p], (size_t)Vp[p] * Up[p]);
            dp[p].push_back(d); vp[p].push_back(m);
        }
    RVec<Mat> om, ov;
    rslf::fuse_disp_maps(dp, vp, om, ov);
    put_f32_stack(om, out_map);
    put_u8_stack(ov, out_valid);
    return 0;
}

/* rslf::compute_1D_edge_confidence_pile on line s of a NORMALISED float stack */
int ref_edge_confidence(const float* epis_norm, int V, int S, int U, int C, int s, const rslf_params* P, float* ce, uint8_t* mask)
{
    RVec<Mat> epis = make_epis(epis_norm, CV_32F, V, S, U, C);
    Mat ce_m = cv::Mat::zeros(V, U, CV_32FC1), mask_m;
    const int thr = omp_get_max_threads();
    if (C == 1) {
        rslf::Depth1DParameters<float> q; fill_params(q, P);
        RVec<rslf::BufferDepth1D<float>*> bufs;
        for (int t = 0; t < thr; ++t) bufs.push_back(new rslf::BufferDepth1D<float>(S, 2, U, CV_32FC1, q));
        rslf::compute_1D_edge_confidence_pile<float>(epis, s, ce_m, mask_m, q, bufs);
        for (auto b : bufs) delete b;
        delete q.par_interpolation_class; delete q.par_kernel_class;
    } else if (C == 3) {
        rslf::Depth1DParameters<cv::Vec3f> q; fill_params(q, P);
        RVec<rslf::BufferDepth1D<cv::Vec3f>*> bufs;
        for (int t = 0; t < thr; ++t) bufs.push_back(new rslf::BufferDepth1D<cv::Vec3f>(S, 2, U, CV_32FC3, q));
        rslf::compute_1D_edge_confidence_pile<cv::Vec3f>(epis, s, ce_m, mask_m, q, bufs);
        for (auto b : bufs) delete b;
        delete q.par_interpolation_class; delete q.par_kernel_class;
    } else return -1;
    put_f32(ce_m, ce);
    put_u8(mask_m, mask);
    return 0;
}

/* rslf::selective_median_filter on plane s_hat */
int ref_selective_median(const float* src_vu, const uint8_t* mask_vu, const float* epis_norm, int V, int S, int U, int C,
                         int s_hat, int size, float eps, float* dst_vu)
{
    RVec<Mat> epis = make_epis(epis_norm, CV_32F, V, S, U, C);
    Mat src(V, U, CV_32FC1), mask(V, U, CV_8UC1), dst;
    std::memcpy(src.data, src_vu, (size_t)V * U * sizeof(float));
    std::memcpy(mask.data, mask_vu, (size_t)V * U);
    if (C == 1) rslf::selective_median_filter<float>(src, dst, epis, s_hat, size, mask, eps);
    else if (C == 3) rslf::selective_median_filter<cv::Vec3f>(src, dst, epis, s_hat, size, mask, eps);
    else return -1;
    put_f32(dst, dst_vu);
    return 0;
}

}  // extern "C"
