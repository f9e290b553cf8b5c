/*
 * rslf_oracle.cpp — CPU restatement of RSLightFields' dense EPI depth path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
 * remotesensingproject_b200/ or its CUDA library) may include, link or call
 * this file; it is used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker and the timed CPU arm.
 *
 * PARITY PIN STATUS: the reference's repository holds no golden vectors or
 * asserting tests, and it cannot be linked against the real OpenCV here (no
 * OpenCV C++ headers/libs in the image, only the cv2 4.13 Python wheel).  This
 * restatement is pinned two ways:
 *  (1) against the REFERENCE'S OWN SOURCES, compiled where they lie against a
 *      stand-in for the OpenCV calls they make (oracle/cvshim, oracle/ref_driver.cpp
 *      -> oracle/_ref/librslf_ref.so): every map of Depth1DComputer_pile,
 *      Depth2DComputer and FineToCoarse bit-identical (tests/test_oracle_vs_reference.py,
 *      fixtures tests/golden/ref_*.npz).  This pins all control flow of the path;
 *  (2) the OpenCV primitives themselves (which (1) takes from the stand-in) against
 *      the real library: a cv2 4.13 replay of the reference's exact call sequence
 *      (oracle/cv2_mirror.py, tests/golden/px_*, down_*, fuse_*, tests/test_oracle_vs_cv2.py).
 * What stays unpinned: OpenCV 3.x-vs-4.x differences inside primitives.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference/RSLightFields).  All arithmetic is float32 with one rounding
 * per reference operation (build with -ffp-contract=off, no -ffast-math), in
 * the reference's operation order; the CUDA kernels use the same order so the
 * two agree bit for bit.
 *
 * Layouts: EPI stacks are dense [V][S][U][C] float32; maps are [S][V][U].
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/rslf_b200.h"

namespace {

struct Dims { int V, S, U, C; };

inline size_t epi_off(const Dims& g, int v, int s, int u) {
    return (((size_t)v * g.S + s) * g.U + u) * g.C;
}

/* types.cpp:80-84 — norm<float>: abs(x) * sqrt(3) evaluated in double, returned as float */
inline float norm1(float x) { return (float)((double)std::fabs(x) * 1.73205080757); }
/* types.cpp:86-91 — norm<Vec3f> = cv::norm(Vec3f): sqrt of the double-accumulated squares */
inline float norm3(float x, float y, float z) {
    double s = 0.0;
    s += (double)x * (double)x;
    s += (double)y * (double)y;
    s += (double)z * (double)z;
    return (float)std::sqrt(s);
}
inline float norm_diff(const float* a, const float* b, int C) {
    if (C == 1) return norm1(a[0] - b[0]);
    return norm3(a[0] - b[0], a[1] - b[1], a[2] - b[2]);
}
inline float norm_px(const float* a, int C) {
    if (C == 1) return norm1(a[0]);
    return norm3(a[0], a[1], a[2]);
}

/* cv::BORDER_REFLECT_101 (gfedcb|abcdefgh|gfedcba) */
inline int reflect101(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * n - 2 - p;
    }
    return p;
}
/* cv::BORDER_REFLECT (fedcba|abcdefgh|hgfedcb) */
inline int reflect(int p, int n) {
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p - 1;
        else p = 2 * n - 1 - p;
    }
    return p;
}

/*
 * compute_1D_edge_confidence (core.hpp:426-478) for line s of EPI row `row`
 * (U x C floats).  8 x filter2D{+1 centre, -1 at j} with BORDER_REFLECT_101
 * (core.hpp:449-458), squares summed channel by channel (core.cpp:6-23) into a
 * zero-initialised accumulator, shadow cut (core.hpp:464-474), threshold (:476).
 */
void edge_confidence_row(const float* row, int U, int C, const rslf_params& P,
                         float* ce, uint8_t* mask) {
    const int fs = P.edge_confidence_filter_size;
    const int centre = (fs - 1) / 2;
    for (int u = 0; u < U; ++u) ce[u] = 0.f;
    for (int j = 0; j < fs; ++j) {
        if (j == centre) continue;
        for (int c = 0; c < C; ++c) {
            for (int u = 0; u < U; ++u) {
                int q = reflect101(u + j - centre, U);
                float diff = row[(size_t)u * C + c] - row[(size_t)q * C + c];
                float sq = diff * diff;
                ce[u] = ce[u] + sq;
            }
        }
    }
    if (P.cut_shadows) {
        for (int u = 0; u < U; ++u)
            if (norm_px(row + (size_t)u * C, C) < P.shadow_level) ce[u] = 0.f;
    }
    for (int u = 0; u < U; ++u) mask[u] = (ce[u] > P.edge_score_threshold) ? 255 : 0;
}

/* Scratch for one pixel's hypothesis sweep, SoA over d so the loops vectorise. */
struct PixelScratch {
    std::vector<float> rad;     /* [S][C][D] raw radiances (NaN where invalid)   */
    std::vector<float> rad0;    /* [S][C][D] un-nanified: max(r, 0)              */
    std::vector<float> Dv, card, rbar, sumK, sumrK, score;
    std::vector<float> x;       /* [C][D] */
    void resize(int S, int C, int D) {
        rad.resize((size_t)S * C * D); rad0.resize((size_t)S * C * D);
        Dv.resize(D); card.resize(D); rbar.resize((size_t)C * D); sumK.resize(D);
        sumrK.resize((size_t)C * D); score.resize(D); x.resize((size_t)C * D);
    }
};

/*
 * One pixel of compute_1D_depth_epi (core.hpp:527-658): index matrix (:540-552),
 * Interpolation1DLinear::interpolate_mat (interp.hpp:155-193), mean shift
 * (:577-610) with BandwidthKernel::evaluate_mat (kern.cpp:16-26 / 39-54) and the
 * multi-channel multiply/divide helpers (core.cpp:25-51), score (:616-625).
 * Leaves score[d], rbar[c][d], Dv[d] in the scratch.
 */
/* Which confidence gates the propagation and the validity maps: 0 = the edge confidence (the reference as written:
 * the `#elseif` lines of core.hpp:1099 / dc.hpp:903 are not directives, so the default build always takes the `#else`
 * branch), 1 = the disparity confidence C_d > par_disp_score_threshold (the reference AS INTENDED with
 * -D_USE_DISP_CONFIDENCE_SCORE once `#elseif` reads `#elif`: core.hpp:1097-1098, dc.hpp:901-902).  orc_set_criterion. */
int g_criterion = 0;

/* REPLAY of recorded decisions (orc_replay_*): checks a run whose arithmetic is only specified to a tolerance — the
 * CUDA path's contracted (FMA) mode — against the reference's rule (north star): same disparity index wherever the
 * winning score margin exceeds margin_tol (1e-5), scores / r_bar / C_d within rel_tol (1e-4).  For every pixel this
 * oracle evaluates, the recorded decision of the same (level, line, pixel) is looked up; it must pick a hypothesis whose
 * exact score is within margin_tol of the exact maximum, and its floats must be within rel_tol of the exact ones for
 * that hypothesis.  The decision is then ADOPTED (index, r_bar, C_d, acceptance), so that everything downstream —
 * median, propagation, the remaining masks and thereby the set of pixels evaluated later, bounds, fusion — follows the
 * recorded run exactly: any remaining difference between the two runs' maps is a genuine error, not tolerance. */
struct Replay {
    bool on = false;
    std::unordered_map<uint64_t, const rslf_decision*> map;
    float margin_tol = 1e-5f, rel_tol = 1e-4f;
    int level = 0;
    long long looked_up = 0, missing = 0, index_changed = 0, bad_margin = 0, bad_score = 0, bad_rbar = 0, bad_cd = 0, bad_disp = 0;
    double worst_margin = 0, worst_rel = 0;
} g_replay;
inline uint64_t replay_key(int level, int s_hat, int pix) { return ((uint64_t)level << 56) | ((uint64_t)s_hat << 36) | (uint64_t)(uint32_t)pix; }
inline bool close_rel(float a, float b, float rel, float abs_floor) { return std::fabs((double)a - (double)b) <= (double)rel * std::max((double)std::fabs(b), (double)abs_floor); }

/* Optional statistics (off by default; orc_ms_stats_*): at which mean-shift iteration r_bar reaches a bitwise
 * fixed point, per hypothesis and per group of 32 consecutive hypotheses (a GPU warp).  Once r_bar(t+1) == r_bar(t)
 * every later iteration and the score repeat exactly, which is what the CUDA kernel's early exit relies on. */
bool g_ms_stats = false;
long long g_ms_hist_lane[33], g_ms_hist_warp[33], g_ms_flat_pixels, g_ms_pixels;

void pixel_scores(const float* epi /* [S][U][C] */, int S, int U, int C, int D,
                  int s_hat, int u, float dmin, float dmax, const rslf_params& P,
                  PixelScratch& w) {
    const float nanf_ = std::numeric_limits<float>::quiet_NaN();
    float* Dv = w.Dv.data();
    /* core.hpp:547-548: D[d] = dmin + d * (dmax - dmin) / (dim_d - 1) */
    {
        float range = dmax - dmin;
        for (int d = 0; d < D; ++d) {
            float t = (float)d * range;
            t = t / (float)(D - 1);
            Dv[d] = dmin + t;
        }
    }
    /* kern.hpp:43: inv_m_h_sq = 1.0 / (h*h) (float product, double divide, float store);
     * kern.cpp:21 uses 3 * inv for 1 channel. */
    float hh = P.kernel_h * P.kernel_h;
    float inv = (float)(1.0 / (double)hh);
    if (C == 1) inv = (float)(3 * inv);
    const float slope = P.slope_factor;
    const float uf = (float)u;
    float* card = w.card.data();
    for (int d = 0; d < D; ++d) card[d] = 0.f;
    /* I = S*D; I *= slope; I += u (core.hpp:550-552), then interpolate_mat */
    for (int s = 0; s < S; ++s) {
        const float k = (float)(s_hat - s);
        const float* row = epi + (size_t)s * U * C;
        float* r = &w.rad[(size_t)s * C * D];
        float* r0 = &w.rad0[(size_t)s * C * D];
        for (int d = 0; d < D; ++d) {
            float I = k * Dv[d];
            I = I * slope;
            I = I + uf;
            float fl = std::floor(I), ce = std::ceil(I);
            /* (int) casts of out-of-int-range floats are UB in the reference too;
             * realistic indices are far inside the int range. */
            int i0 = (int)fl, i1 = (int)ce;
            if (!(i0 < 0 || i1 > U - 1)) {
                float t = I - (float)i0;
                float omt = 1 - t;
                for (int c = 0; c < C; ++c) {
                    float a = omt * row[(size_t)i0 * C + c];
                    float b = t * row[(size_t)i1 * C + c];
                    float val = a + b;
                    r[(size_t)c * D + d] = val;
                    r0[(size_t)c * D + d] = (val > 0.f) ? val : 0.f; /* cv::max(r, 0): NaN/neg -> 0 (core.hpp:580) */
                }
                card[d] += 1.0f;
            } else {
                for (int c = 0; c < C; ++c) {
                    r[(size_t)c * D + d] = nanf_;
                    r0[(size_t)c * D + d] = 0.f;
                }
            }
        }
    }
    /* r_bar <- radiances.row(s_hat) (core.hpp:577) */
    float* rbar = w.rbar.data();
    std::memcpy(rbar, &w.rad[(size_t)s_hat * C * D], sizeof(float) * C * D);
    float* sumK = w.sumK.data();
    float* sumrK = w.sumrK.data();
    const int iters = P.mean_shift_max_iter;
    std::vector<int> fixed_at;                           /* statistics only: iterations a hypothesis needs */
    if (g_ms_stats) fixed_at.assign(D, iters);
    for (int it = 0; it < iters; ++it) {
        for (int d = 0; d < D; ++d) sumK[d] = 0.f;
        for (int i = 0; i < C * D; ++i) sumrK[i] = 0.f;
        for (int s = 0; s < S; ++s) {
            const float* r = &w.rad[(size_t)s * C * D];
            const float* r0 = &w.rad0[(size_t)s * C * D];
            if (C == 1) {
                for (int d = 0; d < D; ++d) {
                    float x = r[d] - rbar[d];            /* core.hpp:591 */
                    float a = inv * x;                   /* kern.cpp:21 multiply(src,src,scale): (scale*a)*b */
                    float b = a * x;
                    float kk = 1.0f - b;                 /* kern.cpp:23 */
                    kk = (kk > 0.f) ? kk : 0.f;          /* kern.cpp:25 max(.,0): NaN -> 0 */
                    float p = r0[d] * kk;                /* core.cpp:28 */
                    sumrK[d] = sumrK[d] + p;             /* core.hpp:602 reduce over s, ascending */
                    sumK[d] = sumK[d] + kk;              /* core.hpp:603 */
                }
            } else {
                for (int d = 0; d < D; ++d) {
                    float x0 = r[d] - rbar[d];
                    float x1 = r[D + d] - rbar[D + d];
                    float x2 = r[2 * D + d] - rbar[2 * D + d];
                    float b0 = (inv * x0) * x0;          /* kern.cpp:43 */
                    float b1 = (inv * x1) * x1;
                    float b2 = (inv * x2) * x2;
                    float sum = (b0 + b1) + b2;          /* kern.cpp:47-49 reduce over channels */
                    float kk = 1.0f - sum;               /* kern.cpp:51 */
                    kk = (kk > 0.f) ? kk : 0.f;          /* kern.cpp:53 */
                    float p0 = r0[d] * kk;               /* core.cpp:33-37 */
                    float p1 = r0[D + d] * kk;
                    float p2 = r0[2 * D + d] * kk;
                    sumrK[d] = sumrK[d] + p0;
                    sumrK[D + d] = sumrK[D + d] + p1;
                    sumrK[2 * D + d] = sumrK[2 * D + d] + p2;
                    sumK[d] = sumK[d] + kk;
                }
            }
        }
        /* core.hpp:606-609: r_bar = sum_rK / sum_K (OpenCV 3 divide: x/0 = 0), then max(.,0) */
        if (g_ms_stats) std::memcpy(w.x.data(), rbar, sizeof(float) * C * D);
        for (int c = 0; c < C; ++c)
            for (int d = 0; d < D; ++d) {
                float den = sumK[d];
                float q = (den != 0.f) ? (sumrK[(size_t)c * D + d] / den) : 0.f;
                rbar[(size_t)c * D + d] = (q > 0.f) ? q : 0.f;
            }
        if (g_ms_stats) {
            for (int d = 0; d < D; ++d) {
                if (fixed_at[d] < iters) continue;
                bool same = true;
                for (int c = 0; c < C; ++c) {
                    uint32_t a, b;
                    std::memcpy(&a, &w.x[(size_t)c * D + d], 4); std::memcpy(&b, &rbar[(size_t)c * D + d], 4);
                    same = same && (a == b);
                }
                if (same) fixed_at[d] = it + 1;          /* iterations executed when the sweep stops here */
            }
        }
    }
    if (g_ms_stats && iters > 0 && iters <= 32) {
        long long hl[33] = {0}, hw[33] = {0};
        for (int d0 = 0; d0 < D; d0 += 32) {
            int mx = 0;
            for (int d = d0; d < std::min(D, d0 + 32); ++d) { hl[fixed_at[d]]++; mx = std::max(mx, fixed_at[d]); }
            hw[mx]++;
        }
        for (int i = 0; i <= iters; ++i) {
            if (hl[i]) {
#pragma omp atomic
                g_ms_hist_lane[i] += hl[i];
            }
            if (hw[i]) {
#pragma omp atomic
                g_ms_hist_warp[i] += hw[i];
            }
        }
#pragma omp atomic
        g_ms_pixels += 1;
        if (dmin == dmax) {
#pragma omp atomic
            g_ms_flat_pixels += 1;
        }
    }
    /* core.hpp:616-622: K is re-evaluated from the stale r - r_bar of the last
     * iteration, i.e. it equals that iteration's K; score = sum_K / card_R, max(.,0). */
    float* score = w.score.data();
    if (iters <= 0) {
        /* r_m_r_bar is an empty Mat in that case in the reference; not reachable with defaults */
        for (int d = 0; d < D; ++d) score[d] = 0.f;
    } else {
        for (int d = 0; d < D; ++d) {
            float q = sumK[d] / card[d];
            score[d] = (q > 0.f) ? q : 0.f;
        }
    }
}

/*
 * compute_1D_depth_epi (core.hpp:480-661) for one EPI row v: AND of the masks
 * (:510-513), per-pixel sweep, argmax / confidences (:630-657).
 * remaining == nullptr means "no mask" (the pile computer, dc.hpp:516).
 * margin (optional) receives best-minus-second-best score per computed pixel.
 */
long depth_row(const float* epi, int S, int U, int C, int D, int s_hat,
               const float* dmin_u, const float* dmax_u, float dmin_c, float dmax_c,
               float* ce, uint8_t* emask, float* cd, float* depth, float* rbar_out,
               uint8_t* remaining, const rslf_params& P, PixelScratch& w,
               float* margin, int32_t* best_idx, int v = 0) {
    long computed = 0;
    for (int u = 0; u < U; ++u) {
        uint8_t m;
        if (remaining) { remaining[u] = remaining[u] & emask[u]; m = remaining[u]; }
        else m = emask[u];
        if (!m) continue;
        ++computed;
        float dmn = dmin_u ? dmin_u[u] : dmin_c;
        float dmx = dmax_u ? dmax_u[u] : dmax_c;
        pixel_scores(epi, S, U, C, D, s_hat, u, dmn, dmx, P, w);
        /* minMaxLoc: first maximum (core.hpp:634) */
        int best = 0; float mx = w.score[0];
        for (int d = 1; d < D; ++d) if (w.score[d] > mx) { mx = w.score[d]; best = d; }
        if (margin) {
            float second = -1.f;
            for (int d = 0; d < D; ++d) if (d != best && w.score[d] > second) second = w.score[d];
            margin[u] = mx - second;
        }
        if (best_idx) best_idx[u] = best;
        if (g_replay.on) {
            /* adopt the recorded decision of this pixel after checking it against the tolerance (see Replay) */
            auto it = g_replay.map.find(replay_key(g_replay.level, s_hat, v * U + u));
#pragma omp atomic
            g_replay.looked_up += 1;
            if (it == g_replay.map.end()) {
#pragma omp atomic
                g_replay.missing += 1;
            } else {
                const rslf_decision& r = *it->second;
                const int gi = (r.index >= 0 && r.index < D) ? r.index : best;
                const float gap = mx - w.score[gi];                        /* >= 0: exact score margin of the recorded choice */
                bool okm = gap <= g_replay.margin_tol, oks = close_rel(r.score, w.score[gi], g_replay.rel_tol, 1e-3f), okr = true;
                for (int c = 0; c < C; ++c) okr = okr && close_rel(r.rbar[c], w.rbar[(size_t)c * D + gi], g_replay.rel_tol, 1e-2f);
                double sum = 0.0;
                for (int d = 0; d < D; ++d) sum += (double)w.score[d];
                const float cd_exact = (float)((double)ce[u] * std::fabs((double)w.score[gi] - sum / (double)D));
                const bool accepted = r.accepted != 0;
                /* C_d = C_e |max - mean|: a difference of two scores, so its tolerance is rel_tol of the scores' scale */
                const bool okc = !accepted || std::fabs((double)r.disp_conf - (double)cd_exact) <= (double)g_replay.rel_tol * std::max((double)ce[u] * (double)std::max(mx, 1e-3f), 1e-9);
                const bool okd = r.disparity == w.Dv[gi];
#pragma omp critical(replay_stats)
                {
                    if (gi != best) g_replay.index_changed += 1;
                    if (!okm) g_replay.bad_margin += 1;
                    if (!oks) g_replay.bad_score += 1;
                    if (!okr) g_replay.bad_rbar += 1;
                    if (!okc) g_replay.bad_cd += 1;
                    if (!okd) g_replay.bad_disp += 1;
                    g_replay.worst_margin = std::max(g_replay.worst_margin, (double)gap);
                    if (w.score[gi] > 1e-3f) g_replay.worst_rel = std::max(g_replay.worst_rel, std::fabs((double)r.score - (double)w.score[gi]) / (double)w.score[gi]);
                }
                if (accepted) {
                    depth[u] = w.Dv[gi];
                    cd[u] = r.disp_conf;
                    for (int c = 0; c < C; ++c) rbar_out[(size_t)u * C + c] = r.rbar[c];
                } else {
                    ce[u] = 0.f; emask[u] = 0;
                }
                continue;
            }
        }
        double maxVal = (double)mx;
        if (maxVal > (double)P.raw_score_threshold) {
            depth[u] = w.Dv[best];
            double sum = 0.0;                                   /* cv::mean: double accumulation (:641) */
            for (int d = 0; d < D; ++d) sum += (double)w.score[d];
            double mean = sum / (double)D;
            cd[u] = (float)((double)ce[u] * std::fabs(maxVal - mean));
            for (int c = 0; c < C; ++c) rbar_out[(size_t)u * C + c] = w.rbar[(size_t)c * D + best];
        } else {
            ce[u] = 0.f;                                        /* core.hpp:655-656 */
            emask[u] = 0;
        }
    }
    return computed;
}

/*
 * selective_median_filter (core.hpp:663-718).  epis: the whole stack; colours
 * are taken on line s_hat.  dst is zero where the mask is unset (:679).
 */
void selective_median(const float* src, float* dst, const float* epis, const Dims& g,
                      int s_hat, int size, const uint8_t* mask, float eps) {
    const int V = g.V, U = g.U, C = g.C;
    const int width = (size - 1) / 2;
    std::fill(dst, dst + (size_t)V * U, 0.f);
#pragma omp parallel for schedule(dynamic, 4)
    for (int v = 0; v < V; ++v) {
        std::vector<float> buf;
        buf.reserve((size_t)size * size);
        for (int u = 0; u < U; ++u) {
            if (!mask[(size_t)v * U + u]) continue;
            buf.clear();
            const float* pc = epis + epi_off(g, v, s_hat, u);
            for (int k = std::max(0, v - width); k < std::min(V, v + width + 1); ++k)
                for (int l = std::max(0, u - width); l < std::min(U, u + width + 1); ++l)
                    if (mask[(size_t)k * U + l] &&
                        norm_diff(pc, epis + epi_off(g, k, s_hat, l), C) < eps)
                        buf.push_back(src[(size_t)k * U + l]);
            std::nth_element(buf.begin(), buf.begin() + buf.size() / 2, buf.end());
            dst[(size_t)v * U + u] = buf[buf.size() / 2];
        }
    }
}

/* cv::getStructuringElement(shape, Size(k, k)) with the default anchor (k/2, k/2): 0 = MORPH_RECT, 1 = MORPH_CROSS,
 * 2 = MORPH_ELLIPSE (row i spans c -+ cvRound(c * sqrt((r^2 - (i-r)^2) / r^2)), r = c = k/2).  k x k bytes, 0 / 1. */
std::vector<uint8_t> structuring_element(int shape, int k) {
    std::vector<uint8_t> K((size_t)k * k, 0);
    const int r = k / 2, c = k / 2;
    const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < k; ++i) {
        int j1 = 0, j2 = 0;
        if (shape == 0 || (shape == 1 && i == r)) j2 = k;
        else if (shape == 1) { j1 = c; j2 = c + 1; }
        else {
            const int dy = i - r;
            if (std::abs(dy) <= r) {
                const int dx = (int)std::nearbyint(c * std::sqrt((r * r - dy * dy) * inv_r2));
                j1 = std::max(c - dx, 0); j2 = std::min(c + dx + 1, k);
            }
        }
        for (int j = j1; j < j2; ++j) K[(size_t)i * k + j] = 1;
    }
    return K;
}

/* cv::morphologyEx(mask, mask, MORPH_OPEN, kernel) (core.hpp:759-769): erosion then dilation of a V x U uint8 image,
 * min / max over the element's non-zero offsets (ky - k/2, kx - k/2); positions outside the image do not take part
 * (BORDER_CONSTANT with morphologyDefaultBorderValue). */
void morph_open(uint8_t* mask, int V, int U, int shape, int k) {
    const std::vector<uint8_t> K = structuring_element(shape, k);
    const int a = k / 2;
    std::vector<uint8_t> tmp((size_t)V * U);
    for (int phase = 0; phase < 2; ++phase) {
        const uint8_t* in = phase == 0 ? mask : tmp.data();
        uint8_t* out = phase == 0 ? tmp.data() : mask;
#pragma omp parallel for schedule(static)
        for (int y = 0; y < V; ++y)
            for (int x = 0; x < U; ++x) {
                int acc = phase == 0 ? 255 : 0;
                for (int ky = 0; ky < k; ++ky) {
                    const int yy = y + ky - a;
                    if (yy < 0 || yy >= V) continue;
                    for (int kx = 0; kx < k; ++kx) {
                        const int xx = x + kx - a;
                        if (!K[(size_t)ky * k + kx] || xx < 0 || xx >= U) continue;
                        const int v = in[(size_t)yy * U + xx];
                        acc = phase == 0 ? std::min(acc, v) : std::max(acc, v);
                    }
                }
                out[(size_t)y * U + x] = (uint8_t)acc;
            }
    }
}

/* compute_1D_edge_confidence_pile (core.hpp:728-770): every row, then the optional morphological opening of the
 * V x U mask (size > 1; disabled by default, core.hpp:29) */
void edge_confidence_plane(const float* epis, const Dims& g, int s, const rslf_params& P,
                           float* ce, uint8_t* mask) {
#pragma omp parallel for schedule(static)
    for (int v = 0; v < g.V; ++v)
        edge_confidence_row(epis + epi_off(g, v, s, 0), g.U, g.C, P,
                            ce + (size_t)v * g.U, mask + (size_t)v * g.U);
    if (P.edge_confidence_opening_size > 1)
        morph_open(mask, g.V, g.U, P.edge_confidence_opening_type, P.edge_confidence_opening_size);
}

/* The s_hat visiting order of compute_2D_depth_epi (core.hpp:953, 981-990). */
std::vector<int> visiting_order(int S) {
    int s_hat = (int)std::floor(S / 2.0);
    std::vector<int> order;
    order.push_back(s_hat);
    for (int off = 1; off < S - s_hat; ++off) {
        order.push_back(s_hat + off);
        if (s_hat - off > -1) order.push_back(s_hat - off);
    }
    return order;
}

/*
 * Depth2DComputer::run (dc.hpp:748-805) = compute_2D_edge_confidence
 * (core.hpp:901-931) + compute_2D_depth_epi (core.hpp:933-1133).
 * All maps [S][V][U]; depth/cd/rbar must be zero-initialised by the caller
 * (dc.hpp:741-744; the reference leaves cd/ce uninitialised = fresh zero pages).
 */
void depth2d(const float* epis, const Dims& g, int D, float dmin_c, float dmax_c,
             const float* dmin_svu, const float* dmax_svu, const rslf_params& P,
             float* ce, uint8_t* emask, float* cd, float* depth, float* rbar,
             double* computed_per_pass /* S entries or null */, double* total_computed) {
    const int V = g.V, S = g.S, U = g.U, C = g.C;
    const size_t plane = (size_t)V * U;
    for (int s = 0; s < S; ++s)
        edge_confidence_plane(epis, g, s, P, ce + s * plane, emask + s * plane);
    std::vector<uint8_t> remaining(emask, emask + (size_t)S * plane);   /* core.hpp:958-963 */
    std::vector<float> filtered(plane);
    std::vector<int> order = visiting_order(S);
    double total = 0;
    int pass = 0;
    for (int s_hat : order) {
        float* ce_p = ce + s_hat * plane;
        uint8_t* em_p = emask + s_hat * plane;
        float* cd_p = cd + s_hat * plane;
        float* dp_p = depth + s_hat * plane;
        float* rb_p = rbar + s_hat * plane * C;
        uint8_t* rem_p = remaining.data() + s_hat * plane;
        long computed = 0;
        /* compute_1D_depth_epi_pile (core.hpp:799-875) */
#pragma omp parallel reduction(+ : computed)
        {
            PixelScratch w; w.resize(S, C, D);
#pragma omp for schedule(dynamic, 1)
            for (int v = 0; v < V; ++v) {
                computed += depth_row(epis + epi_off(g, v, 0, 0), S, U, C, D, s_hat,
                                      dmin_svu ? dmin_svu + s_hat * plane + (size_t)v * U : nullptr,
                                      dmax_svu ? dmax_svu + s_hat * plane + (size_t)v * U : nullptr,
                                      dmin_c, dmax_c,
                                      ce_p + (size_t)v * U, em_p + (size_t)v * U, cd_p + (size_t)v * U,
                                      dp_p + (size_t)v * U, rb_p + (size_t)v * U * C,
                                      rem_p + (size_t)v * U, P, w, nullptr, nullptr, v);
            }
        }
        total += (double)computed;
        if (computed_per_pass) computed_per_pass[pass] = (double)computed;
        ++pass;
        /* core.hpp:881-892: the filtered map only rebinds the local header, so
         * depth[s_hat] keeps the unfiltered values; propagation reads `filtered`. */
        selective_median(dp_p, filtered.data(), epis, g, s_hat, P.median_filter_size, em_p,
                         P.median_filter_epsilon);
        /* propagation (core.hpp:1086-1129); rows independent, u ascending, s ascending */
#pragma omp parallel for schedule(dynamic, 4)
        for (int v = 0; v < V; ++v) {
            for (int u = 0; u < U; ++u) {
                if (g_criterion == 1) {                      /* core.hpp:1097-1098 (as intended) */
                    if (!(cd_p[(size_t)v * U + u] > P.disp_score_threshold)) continue;
                } else if (!em_p[(size_t)v * U + u]) continue;           /* core.hpp:1101-1102 */
                float cur = filtered[(size_t)v * U + u];
                const float* rb = rb_p + ((size_t)v * U + u) * C;
                for (int s = 0; s < S; ++s) {
                    float t = cur * (float)(s_hat - s);
                    t = t * P.slope_factor;
                    int q = u + (int)std::round(t);
                    if (q > -1 && q < U && remaining[s * plane + (size_t)v * U + q] &&
                        norm_diff(epis + epi_off(g, v, s, q), rb, C) < P.propagation_epsilon) {
                        depth[s * plane + (size_t)v * U + q] = cur;
                        remaining[s * plane + (size_t)v * U + q] = 0;
                        cd[s * plane + (size_t)v * U + q] = cd_p[(size_t)v * U + u];
                    }
                }
            }
        }
    }
    if (total_computed) *total_computed = total;
}

/* cvRound: round half to even */
inline int cv_round(double x) { return (int)std::nearbyint(x); }

/*
 * downsample_EPIs (ftc_core.cpp:14-60) for float32 stacks: per view,
 * GaussianBlur 7x7 sigma 0 BORDER_REFLECT (kernel [1,3.5,7,9,7,3.5,1]/32,
 * separable: rows then columns), then resize fx=fy=0.5 INTER_LINEAR, which at
 * exactly 1/2 scale is OpenCV's 2x2 area average with clamped source indices.
 */
void downsample(const float* in, const Dims& g, float* out, int V2, int U2) {
    const int V = g.V, S = g.S, U = g.U, C = g.C;
    const float k0 = 9.f / 32.f, k1 = 7.f / 32.f, k2 = 3.5f / 32.f, k3 = 1.f / 32.f;
    Dims go{V2, S, U2, C};
#pragma omp parallel
    {
        std::vector<float> hbuf((size_t)V * U * C), blur((size_t)V * U * C);
#pragma omp for schedule(dynamic, 1)
        for (int s = 0; s < S; ++s) {
            /* row (horizontal) pass — symmetric form k0*x0 + k1*(x-1+x1) + k2*(x-2+x2) + k3*(x-3+x3) */
            for (int v = 0; v < V; ++v)
                for (int u = 0; u < U; ++u)
                    for (int c = 0; c < C; ++c) {
                        auto px = [&](int q) { return in[epi_off(g, v, s, reflect(q, U)) + c]; };
                        float acc = k0 * px(u);
                        float t1 = px(u - 1) + px(u + 1); t1 = k1 * t1; acc = acc + t1;
                        float t2 = px(u - 2) + px(u + 2); t2 = k2 * t2; acc = acc + t2;
                        float t3 = px(u - 3) + px(u + 3); t3 = k3 * t3; acc = acc + t3;
                        hbuf[((size_t)v * U + u) * C + c] = acc;
                    }
            /* column (vertical) pass */
            for (int v = 0; v < V; ++v)
                for (int u = 0; u < U; ++u)
                    for (int c = 0; c < C; ++c) {
                        auto px = [&](int q) { return hbuf[((size_t)reflect(q, V) * U + u) * C + c]; };
                        float acc = k0 * px(v);
                        float t1 = px(v - 1) + px(v + 1); t1 = k1 * t1; acc = acc + t1;
                        float t2 = px(v - 2) + px(v + 2); t2 = k2 * t2; acc = acc + t2;
                        float t3 = px(v - 3) + px(v + 3); t3 = k3 * t3; acc = acc + t3;
                        blur[((size_t)v * U + u) * C + c] = acc;
                    }
            /* 2x2 area average, source index clamped at odd edges */
            for (int v = 0; v < V2; ++v) {
                int r0 = std::min(2 * v, V - 1), r1 = std::min(2 * v + 1, V - 1);
                for (int u = 0; u < U2; ++u) {
                    int c0 = std::min(2 * u, U - 1), c1 = std::min(2 * u + 1, U - 1);
                    for (int c = 0; c < C; ++c) {
                        float a = blur[((size_t)r0 * U + c0) * C + c];
                        float b = blur[((size_t)r0 * U + c1) * C + c];
                        float cc = blur[((size_t)r1 * U + c0) * C + c];
                        float d = blur[((size_t)r1 * U + c1) * C + c];
                        float sum = ((a + b) + cc) + d;
                        out[epi_off(go, v, s, u) + c] = sum * 0.25f;
                    }
                }
            }
        }
    }
}

/*
 * downsample_EPIs (ftc_core.cpp:14-60) for 8-bit and 16-bit stacks, which keep their depth between pyramid
 * levels (ftc.hpp:142-147 normalises each level from the un-normalised stack).  OpenCV's integer paths, bit-exact
 * against cv2 (oracle/cv2_mirror.py, tests/golden/down_u8_*.npz; tests/test_oracle_vs_cv2.py for CV_16U):
 * GaussianBlur 7x7 sigma 0 = integer kernel [8,28,56,72,56,28,8] (sum 256) along rows, then along columns,
 * result (acc + 2^15) >> 16 (the exact value rounded half up), BORDER_REFLECT; resize fx=fy=0.5 =
 * (a+b+c+d+2) >> 2, and where an odd dimension leaves only one source row / column the exact mean of the
 * available pixels rounded half to even.  (CV_16U resize: OpenCV's own code; an IPP build rounds the 2x2 ties
 * half to even instead.)  Unsigned 32-bit accumulators: 256 * 256 * 65535 + 2^15 < 2^32.
 */
template <typename T>
void downsample_int(const T* in, const Dims& g, T* out, int V2, int U2) {
    const int V = g.V, S = g.S, U = g.U, C = g.C;
    static const uint32_t K[7] = {8, 28, 56, 72, 56, 28, 8};
    Dims go{V2, S, U2, C};
#pragma omp parallel
    {
        std::vector<uint32_t> hbuf((size_t)V * U * C);
        std::vector<T> blur((size_t)V * U * C);
#pragma omp for schedule(dynamic, 1)
        for (int s = 0; s < S; ++s) {
            for (int v = 0; v < V; ++v)
                for (int u = 0; u < U; ++u)
                    for (int c = 0; c < C; ++c) {
                        uint32_t acc = 0;
                        for (int j = 0; j < 7; ++j) acc += K[j] * (uint32_t)in[epi_off(g, v, s, reflect(u + j - 3, U)) + c];
                        hbuf[((size_t)v * U + u) * C + c] = acc;
                    }
            for (int v = 0; v < V; ++v)
                for (int u = 0; u < U; ++u)
                    for (int c = 0; c < C; ++c) {
                        uint32_t acc = 0;
                        for (int j = 0; j < 7; ++j) acc += K[j] * hbuf[((size_t)reflect(v + j - 3, V) * U + u) * C + c];
                        blur[((size_t)v * U + u) * C + c] = (T)((acc + 32768u) >> 16);
                    }
            for (int v = 0; v < V2; ++v) {
                const int nr = (2 * v + 1 < V) ? 2 : 1;
                for (int u = 0; u < U2; ++u) {
                    const int nc = (2 * u + 1 < U) ? 2 : 1;
                    for (int c = 0; c < C; ++c) {
                        uint32_t sum = 0;
                        for (int a = 0; a < nr; ++a)
                            for (int b = 0; b < nc; ++b) sum += blur[((size_t)(2 * v + a) * U + (2 * u + b)) * C + c];
                        const int n = nr * nc;
                        uint32_t r;
                        if (n == 4) r = (sum + 2) >> 2;
                        else if (n == 2) r = (sum + ((sum >> 1) & 1)) >> 1;      /* half to even */
                        else r = sum;
                        out[epi_off(go, v, s, u) + c] = (T)r;
                    }
                }
            }
        }
    }
}
void downsample_u8(const uint8_t* in, const Dims& g, uint8_t* out, int V2, int U2) { downsample_int<uint8_t>(in, g, out, V2, U2); }
void downsample_u16(const uint16_t* in, const Dims& g, uint16_t* out, int V2, int U2) { downsample_int<uint16_t>(in, g, out, V2, U2); }

/*
 * FineToCoarse::run bound propagation (ftc.hpp:201-294) from level p (up) to
 * level p+1 (down).  Left scan tests indices u_up-1 .. 1 (index 0 never),
 * right scan u_up+1 .. U_up-1.
 */
void set_bounds(const float* depth_up, const uint8_t* valid_up, int S, int Vu, int Uu,
                int Vd, int Ud, float* dmin_map, float* dmax_map) {
#pragma omp parallel for schedule(static)
    for (int s = 0; s < S; ++s) {
        const float* dep = depth_up + (size_t)s * Vu * Uu;
        const uint8_t* val = valid_up + (size_t)s * Vu * Uu;
        for (int v = 0; v < Vd; ++v)
            for (int u = 0; u < Ud; ++u) {
                float cand[4]; int n = 0;
                int v_up = std::min(2 * v, Vu - 1);
                int u_up = std::min(2 * u, Uu - 1);
                for (int line = 0; line < 2; ++line) {
                    int vv = v_up + line;
                    if (line == 1 && !(vv < Vu)) break;
                    bool hl = false, hr = false; float dl = 0.f, dr = 0.f;
                    int ul = u_up;
                    while (ul > 1) { ul -= 1; if (val[(size_t)vv * Uu + ul] > 0) { dl = dep[(size_t)vv * Uu + ul]; hl = true; break; } }
                    int ur = u_up;
                    while (ur < Uu - 1) { ur += 1; if (val[(size_t)vv * Uu + ur] > 0) { dr = dep[(size_t)vv * Uu + ur]; hr = true; break; } }
                    /* a NaN depth would also fail the reference's !is_nan test */
                    if (hl && hr && dl == dl && dr == dr) { cand[n++] = dl; cand[n++] = dr; }
                }
                if (n > 1) {
                    float mn = cand[0], mx = cand[0];
                    for (int i = 1; i < n; ++i) { mn = std::min(mn, cand[i]); mx = std::max(mx, cand[i]); }
                    dmin_map[((size_t)s * Vd + v) * Ud + u] = mn;
                    dmax_map[((size_t)s * Vd + v) * Ud + u] = mx;
                }
            }
    }
}

/* cv::resize(..., dsize, INTER_LINEAR) float32 upscaling: half-pixel centres, edge clamp;
 * horizontal pass then vertical pass like OpenCV's HResizeLinear/VResizeLinear. */
void resize_linear(const float* src, int Vs, int Us, float* dst, int Vd, int Ud) {
    std::vector<int> xi(Ud), yi(Vd);
    std::vector<float> xa(Ud), ya(Vd);
    double sx = (double)Us / Ud, sy = (double)Vs / Vd;
    for (int x = 0; x < Ud; ++x) {
        float fx = (float)((x + 0.5) * sx - 0.5);
        int ix = (int)std::floor(fx); fx -= ix;
        if (ix < 0) { ix = 0; fx = 0.f; }
        if (ix >= Us - 1) { ix = Us - 1; fx = 0.f; }
        xi[x] = ix; xa[x] = fx;
    }
    for (int y = 0; y < Vd; ++y) {
        float fy = (float)((y + 0.5) * sy - 0.5);
        int iy = (int)std::floor(fy); fy -= iy;
        if (iy < 0) { iy = 0; fy = 0.f; }
        if (iy >= Vs - 1) { iy = Vs - 1; fy = 0.f; }
        yi[y] = iy; ya[y] = fy;
    }
    for (int y = 0; y < Vd; ++y) {
        int y0 = yi[y], y1 = std::min(y0 + 1, Vs - 1);
        float b1 = ya[y], b0 = 1.f - b1;
        for (int x = 0; x < Ud; ++x) {
            int x0 = xi[x], x1 = std::min(x0 + 1, Us - 1);
            float a1 = xa[x], a0 = 1.f - a1;
            float h0 = src[(size_t)y0 * Us + x0] * a0; { float t = src[(size_t)y0 * Us + x1] * a1; h0 = h0 + t; }
            float h1 = src[(size_t)y1 * Us + x0] * a0; { float t = src[(size_t)y1 * Us + x1] * a1; h1 = h1 + t; }
            float o = h0 * b0; { float t = h1 * b1; o = o + t; }
            dst[(size_t)y * Ud + x] = o;
        }
    }
}

/* cv::resize(..., dsize, INTER_NEAREST): src = min(floor(dst * scale), n-1) */
void resize_nearest_u8(const uint8_t* src, int Vs, int Us, uint8_t* dst, int Vd, int Ud) {
    double sx = (double)Us / Ud, sy = (double)Vs / Vd;
    for (int y = 0; y < Vd; ++y) {
        int iy = std::min((int)std::floor(y * sy), Vs - 1);
        for (int x = 0; x < Ud; ++x) {
            int ix = std::min((int)std::floor(x * sx), Us - 1);
            dst[(size_t)y * Ud + x] = src[(size_t)iy * Us + ix];
        }
    }
}

inline float med3(float a, float b, float c) { return std::max(std::min(a, b), std::min(std::max(a, b), c)); }

/* cv::medianBlur(float32, 3): 3x3 median, BORDER_REPLICATE */
void median3x3(const float* src, float* dst, int V, int U) {
    for (int v = 0; v < V; ++v)
        for (int u = 0; u < U; ++u) {
            float w[9]; int n = 0;
            for (int dv = -1; dv <= 1; ++dv)
                for (int du = -1; du <= 1; ++du) {
                    int vv = std::min(std::max(v + dv, 0), V - 1);
                    int uu = std::min(std::max(u + du, 0), U - 1);
                    w[n++] = src[(size_t)vv * U + uu];
                }
            std::nth_element(w, w + 4, w + 9);
            dst[(size_t)v * U + u] = w[4];
        }
}

/* fuse_disp_maps (ftc_core.cpp:69-135) */
void fuse(int levels, int S, const int* Vp, const int* Up, const float* const* disp,
          const uint8_t* const* valid, float* out_map, uint8_t* out_valid) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int s = 0; s < S; ++s) {
        int p = levels - 1;
        std::vector<float> map_down(disp[p] + (size_t)s * Vp[p] * Up[p], disp[p] + (size_t)(s + 1) * Vp[p] * Up[p]);
        std::vector<uint8_t> mask_down(valid[p] + (size_t)s * Vp[p] * Up[p], valid[p] + (size_t)(s + 1) * Vp[p] * Up[p]);
        for (; p > 0; --p) {
            int Vn = Vp[p - 1], Un = Up[p - 1];
            std::vector<float> map_up((size_t)Vn * Un);
            std::vector<uint8_t> mask_up((size_t)Vn * Un);
            resize_linear(map_down.data(), Vp[p], Up[p], map_up.data(), Vn, Un);
            resize_nearest_u8(mask_down.data(), Vp[p], Up[p], mask_up.data(), Vn, Un);
            const float* dn = disp[p - 1] + (size_t)s * Vn * Un;
            const uint8_t* vn = valid[p - 1] + (size_t)s * Vn * Un;
            map_down.assign((size_t)Vn * Un, 0.f);
            mask_down.assign((size_t)Vn * Un, 0);
            for (size_t i = 0; i < (size_t)Vn * Un; ++i) {
                map_down[i] = vn[i] ? dn[i] : (0.0f + map_up[i]);   /* setTo(0, invalid); add(.., up, .., invalid) */
                mask_down[i] = vn[i] | mask_up[i];
            }
        }
        median3x3(map_down.data(), out_map + (size_t)s * Vp[0] * Up[0], Vp[0], Up[0]);
        std::memcpy(out_valid + (size_t)s * Vp[0] * Up[0], mask_down.data(), (size_t)Vp[0] * Up[0]);
    }
}

/*
 * FineToCoarse::get_coloured_depth_maps (ftc.hpp:324-377) after get_results: map / valid [S][V][U] fused results,
 * epis0 = the normalised level-0 stack (m_computers[0]->get_epis()), lut = 256 x 3 colour table (cv::applyColorMap).
 *  - ImageConverter_uchar::fit (rslf_plot.cpp:66-98) on the map of view (int)std::round(S / 2.0): saturate -> the
 *    values of rank floor(0.02 N) and floor(0.98 N) of the sorted plane; else min = true minimum,
 *    max = min(mean + 12 std, true maximum) (cv::meanStdDev, double sums);
 *  - copy_and_scale (:100-107): float alpha = 255.0 / (max - min); convertTo(CV_8U, alpha, -alpha * min) =
 *    saturate_cast<uchar>(cvRound(x * float(alpha) + float(beta))), multiply and add rounded separately (OpenCV 3.x
 *    cvtScale_<float, uchar, float>; a 4.x build with FMA dispatch fuses them: DESIGN.md section 5);
 *  - colour table, then black where the validity mask is 0 and, with par_cut_shadows, where
 *    norm(E(s, u)) < _SHADOW_NORMALIZED_LEVEL (the macro core.hpp:30, not par_shadow_level).
 * out: [S][V][U][3] uint8; fit_min_max (optional): the fitted min and max.
 */
void colour_maps(const float* map, const uint8_t* valid, const float* epis0, const Dims& g, const uint8_t* lut,
                 bool saturate, bool cut_shadows, uint8_t* out, double* fit_min_max) {
    const int S = g.S, V = g.V, U = g.U, C = g.C;
    const size_t plane = (size_t)V * U;
    const int s_fit = (int)std::round(S / 2.0);
    const float* fitp = map + (size_t)s_fit * plane;
    double mn, mx;
    if (saturate) {
        std::vector<float> sorted(fitp, fitp + plane);
        std::sort(sorted.begin(), sorted.end());
        mn = sorted[(size_t)std::floor(0.02 * (double)plane)];
        mx = sorted[(size_t)std::floor(0.98 * (double)plane)];
    } else {
        double s = 0, sq = 0, tmin = fitp[0], tmax = fitp[0];
        for (size_t i = 0; i < plane; ++i) {
            const double v = fitp[i];
            s += v; sq += v * v; tmin = std::min(tmin, v); tmax = std::max(tmax, v);
        }
        const double mu = s / (double)plane, sd = std::sqrt(std::max(sq / (double)plane - mu * mu, 0.0));
        mn = tmin; mx = std::min(mu + 12 * sd, tmax);
    }
    if (fit_min_max) { fit_min_max[0] = mn; fit_min_max[1] = mx; }
    const float alpha = (float)(255.0 / (mx - mn));
    const double beta_d = -alpha * mn;                       /* float * double -> double (rslf_plot.cpp:106) */
    const float a = (float)(double)alpha, b = (float)beta_d;
    const float shadow = (float)(0.05 * 1.73205080757);      /* im_norm < _SHADOW_NORMALIZED_LEVEL on a CV_32F image */
#pragma omp parallel for schedule(static)
    for (int s = 0; s < S; ++s)
        for (int v = 0; v < V; ++v)
            for (int u = 0; u < U; ++u) {
                const size_t o = (size_t)s * plane + (size_t)v * U + u;
                float x = map[o] * a;
                x = x + b;
                int r = (int)std::nearbyint(x);              /* cvRound, then saturate_cast<uchar> */
                r = r < 0 ? 0 : r > 255 ? 255 : r;
                bool black = valid[o] == 0;
                if (cut_shadows && norm_px(epis0 + epi_off(g, v, s, u), C) < shadow) black = true;
                for (int c = 0; c < 3; ++c) out[o * 3 + c] = black ? 0 : lut[3 * r + c];
            }
}

/* Input normalisation of the computers' ctors (dc.hpp:442-477, 668-704). */
float normalise(const void* raw, int cv_depth, size_t n, float scale_factor, float* out) {
    if (cv_depth == RSLF_DEPTH_8U) {
        const uint8_t* p = (const uint8_t*)raw;
        const float a = (float)(1.0 / 255.0);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)n; ++i) out[i] = (float)p[i] * a;
        return 255.f;
    }
    if (cv_depth == RSLF_DEPTH_16U) {
        /* minMaxLoc -> (float)max (dc.hpp:453-456); convertTo(CV_32F, 1.0 / scale): float(x) * float(alpha) */
        const uint16_t* p = (const uint16_t*)raw;
        float sf = scale_factor;
        if (sf < 0) {
            float mx = sf;
#pragma omp parallel for reduction(max : mx) schedule(static)
            for (long long i = 0; i < (long long)n; ++i) mx = std::max(mx, (float)p[i]);
            sf = mx;
        }
        const float a = (float)(1.0 / (double)sf);
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)n; ++i) out[i] = (float)p[i] * a;
        return sf;
    }
    const float* p = (const float*)raw;
    float sf = scale_factor;
    if (sf < 0) {
        float mx = sf;
#pragma omp parallel for reduction(max : mx) schedule(static)
        for (long long i = 0; i < (long long)n; ++i) mx = std::max(mx, p[i]);
        sf = mx;
    }
    const float a = (float)(1.0 / (double)sf);
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) out[i] = p[i] * a;
    return sf;
}

}  // namespace

extern "C" {

/* statistics of the mean-shift fixed point (see pixel_scores): enable + reset, read back */
void orc_ms_stats_enable(int on) {
    g_ms_stats = on != 0;
    std::memset(g_ms_hist_lane, 0, sizeof(g_ms_hist_lane)); std::memset(g_ms_hist_warp, 0, sizeof(g_ms_hist_warp));
    g_ms_flat_pixels = 0; g_ms_pixels = 0;
}
void orc_ms_stats_read(long long* lane33, long long* warp33, long long* pixels, long long* flat_pixels) {
    std::memcpy(lane33, g_ms_hist_lane, sizeof(g_ms_hist_lane)); std::memcpy(warp33, g_ms_hist_warp, sizeof(g_ms_hist_warp));
    *pixels = g_ms_pixels; *flat_pixels = g_ms_flat_pixels;
}

/* replay of recorded decisions (see Replay): load `n` records, run any of the entry points below, read the report */
void orc_replay_begin(const rslf_decision* recs, long long n, float margin_tol, float rel_tol) {
    g_replay = Replay();
    g_replay.map.reserve((size_t)n * 2);
    for (long long i = 0; i < n; ++i) g_replay.map[replay_key(recs[i].level, recs[i].s_hat, recs[i].pix)] = recs + i;
    g_replay.margin_tol = margin_tol; g_replay.rel_tol = rel_tol; g_replay.on = true;
}
/* out[0..9] = looked up, missing, index changed, bad margin, bad score, bad r_bar, bad C_d, bad disparity value,
 * records loaded, 0; worst[0..1] = largest exact margin of an adopted choice, largest relative score error */
void orc_replay_end(long long* out10, double* worst2) {
    out10[0] = g_replay.looked_up; out10[1] = g_replay.missing; out10[2] = g_replay.index_changed; out10[3] = g_replay.bad_margin;
    out10[4] = g_replay.bad_score; out10[5] = g_replay.bad_rbar; out10[6] = g_replay.bad_cd; out10[7] = g_replay.bad_disp;
    out10[8] = (long long)g_replay.map.size(); out10[9] = 0;
    worst2[0] = g_replay.worst_margin; worst2[1] = g_replay.worst_rel;
    g_replay = Replay();
}

void orc_set_criterion(int c) { g_criterion = c; }
int orc_get_criterion(void) { return g_criterion; }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void orc_params_default(rslf_params* p) {
    p->edge_score_threshold = 0.02f; p->line_score_threshold = 0.02f;
    p->disp_score_threshold = 0.01f; p->raw_score_threshold = 0.f;
    p->mean_shift_max_iter = 10; p->edge_confidence_filter_size = 9;
    p->edge_confidence_opening_type = 2; p->edge_confidence_opening_size = 1;
    p->median_filter_size = 5; p->median_filter_epsilon = 0.1f; p->propagation_epsilon = 0.1f;
    p->slope_factor = 1.0f; p->cut_shadows = 1; p->shadow_level = (float)(0.05 * 1.73205080757);
    p->kernel_h = 0.2f;
}

float orc_normalise(const void* raw, int cv_depth, int V, int S, int U, int C, float scale_factor, float* out) {
    return normalise(raw, cv_depth, (size_t)V * S * U * C, scale_factor, out);
}

void orc_edge_confidence(const float* epis, int V, int S, int U, int C, int s, const rslf_params* P,
                         float* ce_vu, uint8_t* mask_vu) {
    Dims g{V, S, U, C};
    edge_confidence_plane(epis, g, s, *P, ce_vu, mask_vu);
}

/* scores[d], rbar[d][c], Dv[d] of one pixel (for fine-grained checks) */
void orc_pixel_scores(const float* epi_su, int S, int U, int C, int D, int s_hat, int u, float dmin, float dmax,
                      const rslf_params* P, float* scores, float* rbar_dc, float* dvals) {
    PixelScratch w; w.resize(S, C, D);
    pixel_scores(epi_su, S, U, C, D, s_hat, u, dmin, dmax, *P, w);
    for (int d = 0; d < D; ++d) {
        scores[d] = w.score[d]; dvals[d] = w.Dv[d];
        for (int c = 0; c < C; ++c) rbar_dc[(size_t)d * C + c] = w.rbar[(size_t)c * D + d];
    }
}

/* getStructuringElement / morphologyEx(MORPH_OPEN) as the edge-mask opening uses them (core.hpp:759-769) */
void orc_structuring_element(int shape, int k, uint8_t* out_kk) {
    std::vector<uint8_t> K = structuring_element(shape, k);
    std::memcpy(out_kk, K.data(), K.size());
}
void orc_morph_open(uint8_t* mask_vu, int V, int U, int shape, int k) { morph_open(mask_vu, V, U, shape, k); }

void orc_selective_median(const float* src_vu, const uint8_t* mask_vu, const float* epis, int V, int S, int U, int C,
                          int s_hat, int size, float eps, float* dst_vu) {
    Dims g{V, S, U, C};
    selective_median(src_vu, dst_vu, epis, g, s_hat, size, mask_vu, eps);
}

/*
 * Depth1DComputer_pile::run (dc.hpp:513-565).  Outputs V x U (rbar V x U x C),
 * zero-initialised here.  best_depth is the filtered map.  margin / best_idx
 * (optional) are diagnostic: score margin and argmax index per computed pixel.
 * Returns the number of pixels evaluated.
 */
double orc_depth1d_pile(const float* epis, int V, int S, int U, int C, float dmin, float dmax, int D, int s_hat,
                        const rslf_params* Pp, float* best_depth, float* ce, uint8_t* emask, float* cd,
                        float* rbar, float* raw_depth /* optional: unfiltered */, float* margin, int32_t* best_idx) {
    const rslf_params& P = *Pp;
    Dims g{V, S, U, C};
    if (s_hat < 0 || s_hat > S - 1) s_hat = (int)std::floor((0.0 + S) / 2);
    const size_t plane = (size_t)V * U;
    std::vector<float> depth(plane, 0.f);
    std::fill(cd, cd + plane, 0.f);
    std::fill(rbar, rbar + plane * C, 0.f);
    if (margin) std::fill(margin, margin + plane, 0.f);
    if (best_idx) std::fill(best_idx, best_idx + plane, -1);
    edge_confidence_plane(epis, g, s_hat, P, ce, emask);
    long computed = 0;
#pragma omp parallel reduction(+ : computed)
    {
        PixelScratch w; w.resize(S, C, D);
#pragma omp for schedule(dynamic, 1)
        for (int v = 0; v < V; ++v)
            computed += depth_row(epis + epi_off(g, v, 0, 0), S, U, C, D, s_hat, nullptr, nullptr, dmin, dmax,
                                  ce + (size_t)v * U, emask + (size_t)v * U, cd + (size_t)v * U,
                                  depth.data() + (size_t)v * U, rbar + (size_t)v * U * C, nullptr, P, w,
                                  margin ? margin + (size_t)v * U : nullptr,
                                  best_idx ? best_idx + (size_t)v * U : nullptr, v);
    }
    if (raw_depth) std::memcpy(raw_depth, depth.data(), plane * sizeof(float));
    selective_median(depth.data(), best_depth, epis, g, s_hat, P.median_filter_size, emask, P.median_filter_epsilon);
    return (double)computed;
}

/* Depth2DComputer::run.  Maps [S][V][U]; zero-initialised here.  Returns pixels evaluated. */
double orc_depth2d(const float* epis, int V, int S, int U, int C, float dmin, float dmax, int D,
                   const rslf_params* P, const float* dmin_svu, const float* dmax_svu,
                   float* best_depth, float* ce, uint8_t* emask, float* cd, float* rbar,
                   double* computed_per_pass) {
    Dims g{V, S, U, C};
    const size_t n = (size_t)S * V * U;
    std::fill(best_depth, best_depth + n, 0.f);
    std::fill(cd, cd + n, 0.f);
    std::fill(rbar, rbar + n * C, 0.f);
    double total = 0;
    depth2d(epis, g, D, dmin, dmax, dmin_svu, dmax_svu, *P, ce, emask, cd, best_depth, rbar, computed_per_pass, &total);
    return total;
}

int orc_visiting_order(int S, int* order) {
    std::vector<int> o = visiting_order(S);
    for (size_t i = 0; i < o.size(); ++i) order[i] = o[i];
    return (int)o.size();
}

int orc_half_size(int n) { return cv_round(n * 0.5); }

void orc_downsample(const float* in, int V, int S, int U, int C, float* out) {
    Dims g{V, S, U, C};
    downsample(in, g, out, cv_round(V * 0.5), cv_round(U * 0.5));
}

void orc_downsample_u8(const uint8_t* in, int V, int S, int U, int C, uint8_t* out) {
    Dims g{V, S, U, C};
    downsample_u8(in, g, out, cv_round(V * 0.5), cv_round(U * 0.5));
}

void orc_downsample_u16(const uint16_t* in, int V, int S, int U, int C, uint16_t* out) {
    Dims g{V, S, U, C};
    downsample_u16(in, g, out, cv_round(V * 0.5), cv_round(U * 0.5));
}

void orc_set_bounds(const float* depth_up, const uint8_t* valid_up, int S, int Vu, int Uu, int Vd, int Ud,
                    float* dmin_map, float* dmax_map) {
    set_bounds(depth_up, valid_up, S, Vu, Uu, Vd, Ud, dmin_map, dmax_map);
}

void orc_fuse(int levels, int S, const int* Vp, const int* Up, const float* const* disp, const uint8_t* const* valid,
              float* out_map, uint8_t* out_valid) {
    fuse(levels, S, Vp, Up, disp, valid, out_map, out_valid);
}

void orc_colour_maps(const float* map_svu, const uint8_t* valid_svu, const float* epis0_norm, int V, int S, int U, int C,
                     const uint8_t* lut_bgr, int saturate, int cut_shadows, uint8_t* out_bgr, double* fit_min_max) {
    Dims g{V, S, U, C};
    colour_maps(map_svu, valid_svu, epis0_norm, g, lut_bgr, saturate != 0, cut_shadows != 0, out_bgr, fit_min_max);
}

void orc_resize_linear(const float* src, int Vs, int Us, float* dst, int Vd, int Ud) { resize_linear(src, Vs, Us, dst, Vd, Ud); }
void orc_median3x3(const float* src, float* dst, int V, int U) { median3x3(src, dst, V, U); }

/* Number of pyramid levels FineToCoarse builds (ftc.hpp:130) and their sizes. */
int orc_pyramid_dims(int V, int U, int max_pyr_depth, int* Vp, int* Up) {
    if (max_pyr_depth < 1) max_pyr_depth = std::numeric_limits<int>::max();
    int n = 0;
    while (V > 10 && U > 10 && n < max_pyr_depth) {
        Vp[n] = V; Up[n] = U; ++n;
        V = cv_round(V * 0.5); U = cv_round(U * 0.5);
    }
    return n;
}

/*
 * FineToCoarse ctor + run + get_results (ftc.hpp:103-322) for a float32 or 8U
 * level-0 stack `raw` ([V][S][U][C]); 8-bit stacks stay 8-bit between levels.
 * level_out (optional): per level p, pointers to caller-allocated
 * [depth, ce, cd, dmin, dmax] float maps and [emask] — see python wrapper.
 * Returns total pixels evaluated; samples_out = sum over levels computed*D*S.
 */
double orc_fine_to_coarse(const void* raw, int cv_depth, int V, int S, int U, int C, float scale_factor,
                          float dmin, float dmax, int D, const rslf_params* Pp, int max_pyr_depth,
                          int accept_all_last, float* out_map, uint8_t* out_valid,
                          float** lvl_depth, float** lvl_ce, uint8_t** lvl_emask, float** lvl_cd,
                          float** lvl_dmin, float** lvl_dmax, double* computed_per_level) {
    int Vp[32], Up[32];
    int levels = orc_pyramid_dims(V, U, max_pyr_depth, Vp, Up);
    if (levels == 0) return 0;
    if (cv_depth != RSLF_DEPTH_32F && cv_depth != RSLF_DEPTH_8U && cv_depth != RSLF_DEPTH_16U) return -1;
    std::vector<std::vector<float>> depth(levels), ce(levels), cd(levels), dmn(levels), dmx(levels);
    std::vector<std::vector<uint8_t>> emask(levels), valid(levels);
    std::vector<float> raw_cur, raw_next, norm, rbar;
    std::vector<uint8_t> raw8_cur, raw8_next;
    std::vector<uint16_t> raw16_cur, raw16_next;
    size_t n0 = (size_t)V * S * U * C;
    const void* cur_raw = raw;
    double total = 0;
    for (int p = 0; p < levels; ++p) {
        g_replay.level = p;
        Dims g{Vp[p], S, Up[p], C};
        size_t n = (size_t)S * Vp[p] * Up[p];
        norm.resize((size_t)Vp[p] * S * Up[p] * C);
        normalise(cur_raw, cv_depth, norm.size(), scale_factor, norm.data());
        rslf_params P = *Pp;
        P.slope_factor = (float)((0.0 + Up[p]) / Up[0]);          /* ftc.hpp:139 */
        depth[p].assign(n, 0.f); ce[p].assign(n, 0.f); cd[p].assign(n, 0.f); emask[p].assign(n, 0);
        rbar.assign(n * C, 0.f);
        if (p == 0) { dmn[p].assign(n, dmin); dmx[p].assign(n, dmax); }
        double comp = 0;
        depth2d(norm.data(), g, D, dmin, dmax, dmn[p].data(), dmx[p].data(), P, ce[p].data(), emask[p].data(),
                cd[p].data(), depth[p].data(), rbar.data(), nullptr, &comp);
        total += comp;
        if (computed_per_level) computed_per_level[p] = comp;
        /* validity (dc.hpp:893-915): C_e > thr, or everything at an accept-all level */
        valid[p].resize(n);
        bool accept_all = accept_all_last && (p == levels - 1);
        for (size_t i = 0; i < n; ++i)
            valid[p][i] = accept_all ? (ce[p][i] > -1.f ? 255 : 0)
                          : g_criterion == 1 ? (cd[p][i] > P.disp_score_threshold ? 255 : 0)      /* dc.hpp:901-902 */
                                             : (ce[p][i] > P.edge_score_threshold ? 255 : 0);    /* dc.hpp:905-906 */
        if (p + 1 < levels) {
            /* next level's raw stack (ftc.hpp:146) and bounds (ftc.hpp:201-294) */
            if (cv_depth == RSLF_DEPTH_8U) {
                raw8_next.resize((size_t)Vp[p + 1] * S * Up[p + 1] * C);
                downsample_u8((const uint8_t*)cur_raw, g, raw8_next.data(), Vp[p + 1], Up[p + 1]);
            } else if (cv_depth == RSLF_DEPTH_16U) {
                raw16_next.resize((size_t)Vp[p + 1] * S * Up[p + 1] * C);
                downsample_u16((const uint16_t*)cur_raw, g, raw16_next.data(), Vp[p + 1], Up[p + 1]);
            } else {
                raw_next.resize((size_t)Vp[p + 1] * S * Up[p + 1] * C);
                downsample((const float*)cur_raw, g, raw_next.data(), Vp[p + 1], Up[p + 1]);
            }
            size_t nn = (size_t)S * Vp[p + 1] * Up[p + 1];
            dmn[p + 1].assign(nn, dmin); dmx[p + 1].assign(nn, dmax);
            /* bounds use level p's validity with m_accept_all as set: only the last level accepts all */
            set_bounds(depth[p].data(), valid[p].data(), S, Vp[p], Up[p], Vp[p + 1], Up[p + 1], dmn[p + 1].data(),
                       dmx[p + 1].data());
            if (cv_depth == RSLF_DEPTH_8U) { raw8_cur.swap(raw8_next); cur_raw = raw8_cur.data(); }
            else if (cv_depth == RSLF_DEPTH_16U) { raw16_cur.swap(raw16_next); cur_raw = raw16_cur.data(); }
            else { raw_cur.swap(raw_next); cur_raw = raw_cur.data(); }
        }
    }
    (void)n0;
    std::vector<const float*> dp(levels); std::vector<const uint8_t*> vp(levels);
    for (int p = 0; p < levels; ++p) { dp[p] = depth[p].data(); vp[p] = valid[p].data(); }
    fuse(levels, S, Vp, Up, dp.data(), vp.data(), out_map, out_valid);
    for (int p = 0; p < levels; ++p) {
        size_t n = (size_t)S * Vp[p] * Up[p];
        if (lvl_depth && lvl_depth[p]) std::memcpy(lvl_depth[p], depth[p].data(), n * 4);
        if (lvl_ce && lvl_ce[p]) std::memcpy(lvl_ce[p], ce[p].data(), n * 4);
        if (lvl_emask && lvl_emask[p]) std::memcpy(lvl_emask[p], emask[p].data(), n);
        if (lvl_cd && lvl_cd[p]) std::memcpy(lvl_cd[p], cd[p].data(), n * 4);
        if (lvl_dmin && lvl_dmin[p]) std::memcpy(lvl_dmin[p], dmn[p].data(), n * 4);
        if (lvl_dmax && lvl_dmax[p]) std::memcpy(lvl_dmax[p], dmx[p].data(), n * 4);
    }
    return total;
}

}  // extern "C"
