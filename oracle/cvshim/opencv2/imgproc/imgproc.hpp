/*
 * cvshim imgproc — filter2D, GaussianBlur, resize, medianBlur as RSLightFields' depth path calls them
 * (see ../core/core.hpp for what this stand-in is).  TEST INFRASTRUCTURE ONLY.
 *
 * Each function states the arithmetic it follows; the same arithmetic is restated in oracle/rslf_oracle.cpp,
 * where it is pinned against the cv2 4.13 wheel (tests/test_oracle_vs_cv2.py, tests/golden/).
 * Unsupported argument combinations abort instead of approximating.
 */
#ifndef CVSHIM_IMGPROC_HPP
#define CVSHIM_IMGPROC_HPP
#include "../core/core.hpp"

namespace cv {

namespace detail {
/* cv::borderInterpolate */
inline int border(int p, int n, int type)
{
    if ((unsigned)p < (unsigned)n) return p;
    if (type == BORDER_REPLICATE) return p < 0 ? 0 : n - 1;
    if (type == BORDER_REFLECT || type == BORDER_REFLECT_101) {
        const int delta = type == BORDER_REFLECT_101 ? 1 : 0;
        if (n == 1) return 0;
        do {
            if (p < 0) p = -p - 1 + delta;
            else p = n - 1 - (p - n) - delta;
        } while ((unsigned)p >= (unsigned)n);
        return p;
    }
    shim_abort("border type");
}
inline int cv_round(double x) { return (int)std::nearbyint(x); }     /* cvRound: half to even */
}  // namespace detail

/* cv::filter2D on CV_32F with a CV_32F kernel (correlation): dst(x) = delta + sum over the kernel's NON-ZERO
 * coefficients, in scan order, of k * src(x + j - anchor), accumulated in float (FilterEngine / Filter2D). */
inline void filter2D(const Mat& src, Mat& dst, int ddepth, const Mat& kernel, Point anchor = Point(-1, -1), double delta = 0,
                     int borderType = BORDER_DEFAULT)
{
    shim_check(src.depth() == CV_32F && (ddepth == -1 || ddepth == CV_32F) && kernel.type() == CV_32FC1, "filter2D: CV_32F");
    Mat S = src, K = kernel;
    const int ax = anchor.x < 0 ? K.cols / 2 : anchor.x, ay = anchor.y < 0 ? K.rows / 2 : anchor.y;
    const int cn = S.channels();
    Mat out(S.rows, S.cols, S.type());
    for (int y = 0; y < S.rows; ++y)
        for (int x = 0; x < S.cols; ++x)
            for (int c = 0; c < cn; ++c) {
                float acc = (float)delta;
                for (int ky = 0; ky < K.rows; ++ky)
                    for (int kx = 0; kx < K.cols; ++kx) {
                        const float k = K.at<float>(ky, kx);
                        if (k == 0.f) continue;
                        const int yy = detail::border(y + ky - ay, S.rows, borderType), xx = detail::border(x + kx - ax, S.cols, borderType);
                        const float t = k * S.ptr<float>(yy)[xx * cn + c];
                        acc = acc + t;
                    }
                out.ptr<float>(y)[x * cn + c] = acc;
            }
    if (dst.rows == out.rows && dst.cols == out.cols && dst.type() == out.type() && dst.data) out.copyTo(dst);
    else dst = out;
}

/* cv::GaussianBlur, ksize 7x7, sigma 0 (-> sigma 1.4, the fixed small-kernel table [0.03125, 0.109375, 0.21875,
 * 0.28125, ...]), separable: rows, then columns.
 * CV_32F: symmetric row / column filters, k0*x0 + k1*(x-1 + x+1) + k2*(x-2 + x+2) + k3*(x-3 + x+3) in float.
 * CV_8U / CV_16U: fixed-point kernel [8, 28, 56, 72, 56, 28, 8] / 256 in both passes, result (acc + 2^15) >> 16. */
inline void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT)
{
    shim_check(ksize.width == 7 && ksize.height == 7 && sigmaX == 0 && sigmaY == 0, "GaussianBlur: 7x7, sigma 0");
    Mat S = src;
    const int V = S.rows, U = S.cols, cn = S.channels();
    Mat out(V, U, S.type());
    if (S.depth() == CV_32F) {
        const float k0 = 9.f / 32.f, k1 = 7.f / 32.f, k2 = 3.5f / 32.f, k3 = 1.f / 32.f;
        Mat h(V, U, S.type());
        for (int v = 0; v < V; ++v)
            for (int u = 0; u < U; ++u)
                for (int c = 0; c < cn; ++c) {
                    auto px = [&](int q) { return S.ptr<float>(v)[detail::border(q, U, borderType) * cn + c]; };
                    float acc = k0 * px(u);
                    float t1 = px(u - 1) + px(u + 1); t1 = k1 * t1; acc = acc + t1;
                    float t2 = px(u - 2) + px(u + 2); t2 = k2 * t2; acc = acc + t2;
                    float t3 = px(u - 3) + px(u + 3); t3 = k3 * t3; acc = acc + t3;
                    h.ptr<float>(v)[u * cn + c] = acc;
                }
        for (int v = 0; v < V; ++v)
            for (int u = 0; u < U; ++u)
                for (int c = 0; c < cn; ++c) {
                    auto px = [&](int q) { return h.ptr<float>(detail::border(q, V, borderType))[u * cn + c]; };
                    float acc = k0 * px(v);
                    float t1 = px(v - 1) + px(v + 1); t1 = k1 * t1; acc = acc + t1;
                    float t2 = px(v - 2) + px(v + 2); t2 = k2 * t2; acc = acc + t2;
                    float t3 = px(v - 3) + px(v + 3); t3 = k3 * t3; acc = acc + t3;
                    out.ptr<float>(v)[u * cn + c] = acc;
                }
    } else if (S.depth() == CV_8U || S.depth() == CV_16U) {
        /* the same fixed-point kernel for both integer depths (pinned against cv2 4.13 incl. exact ties);
         * unsigned 32-bit: 256 * 256 * 65535 + 2^15 < 2^32 */
        const bool wide = S.depth() == CV_16U;
        static const uint32_t K[7] = {8, 28, 56, 72, 56, 28, 8};
        auto px = [&](int v, int i) -> uint32_t { return wide ? (uint32_t)S.ptr<ushort>(v)[i] : (uint32_t)S.ptr<uchar>(v)[i]; };
        std::vector<uint32_t> h((size_t)V * U * cn);
        for (int v = 0; v < V; ++v)
            for (int u = 0; u < U; ++u)
                for (int c = 0; c < cn; ++c) {
                    uint32_t acc = 0;
                    for (int j = 0; j < 7; ++j) acc += K[j] * px(v, detail::border(u + j - 3, U, borderType) * cn + c);
                    h[((size_t)v * U + u) * cn + c] = acc;
                }
        for (int v = 0; v < V; ++v)
            for (int u = 0; u < U; ++u)
                for (int c = 0; c < cn; ++c) {
                    uint32_t acc = 0;
                    for (int j = 0; j < 7; ++j) acc += K[j] * h[((size_t)detail::border(v + j - 3, V, borderType) * U + u) * cn + c];
                    const uint32_t r = (acc + 32768u) >> 16;
                    if (wide) out.ptr<ushort>(v)[u * cn + c] = (ushort)r; else out.ptr<uchar>(v)[u * cn + c] = (uchar)r;
                }
    } else shim_abort("GaussianBlur on this depth");
    dst = out;
}

/* cv::resize.
 * INTER_LINEAR with fx = fy = 0.5 (dsize = cvRound(n * 0.5)): OpenCV's exact-half fast path, the 2x2 mean with the source
 *   index clamped at an odd edge — CV_32F: ((a + b) + c) + d, then * 0.25f; CV_8U: (a + b + c + d + 2) >> 2, and the mean of
 *   the available pixels rounded half to even where only one source row / column exists.
 * INTER_LINEAR to a larger dsize, CV_32FC1: half-pixel centres, edge clamp, horizontal pass then vertical pass.
 * INTER_NEAREST: src = min(floor(dst * scale), n - 1). */
inline void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR)
{
    Mat S = src;
    const int Vs = S.rows, Us = S.cols, cn = S.channels();
    if (dsize.width == 0 || dsize.height == 0) {
        shim_check(fx > 0 && fy > 0, "resize: scale");
        dsize = Size(detail::cv_round(Us * fx), detail::cv_round(Vs * fy));
    }
    const int Vd = dsize.height, Ud = dsize.width;
    Mat out(Vd, Ud, S.type());
    if (interpolation == INTER_NEAREST) {
        const double sx = (double)Us / Ud, sy = (double)Vs / Vd;
        const size_t es = S.elemSize();
        for (int y = 0; y < Vd; ++y) {
            const int iy = std::min((int)std::floor(y * sy), Vs - 1);
            for (int x = 0; x < Ud; ++x) {
                const int ix = std::min((int)std::floor(x * sx), Us - 1);
                std::memcpy(out.ptr(y) + x * es, S.ptr(iy) + ix * es, es);
            }
        }
    } else if (interpolation == INTER_LINEAR && fx == 0.5 && fy == 0.5) {
        if (S.depth() == CV_32F) {
            for (int v = 0; v < Vd; ++v) {
                const int r0 = std::min(2 * v, Vs - 1), r1 = std::min(2 * v + 1, Vs - 1);
                for (int u = 0; u < Ud; ++u) {
                    const int c0 = std::min(2 * u, Us - 1), c1 = std::min(2 * u + 1, Us - 1);
                    for (int c = 0; c < cn; ++c) {
                        const float a = S.ptr<float>(r0)[c0 * cn + c], b = S.ptr<float>(r0)[c1 * cn + c];
                        const float cc = S.ptr<float>(r1)[c0 * cn + c], d = S.ptr<float>(r1)[c1 * cn + c];
                        float sum = a + b; sum = sum + cc; sum = sum + d;
                        out.ptr<float>(v)[u * cn + c] = sum * 0.25f;
                    }
                }
            }
        } else if (S.depth() == CV_8U || S.depth() == CV_16U) {
            /* CV_16U: OpenCV's own code path (an IPP build rounds the 2x2 ties half to even instead) */
            const bool wide = S.depth() == CV_16U;
            for (int v = 0; v < Vd; ++v) {
                const int nr = (2 * v + 1 < Vs) ? 2 : 1;
                for (int u = 0; u < Ud; ++u) {
                    const int nc = (2 * u + 1 < Us) ? 2 : 1;
                    for (int c = 0; c < cn; ++c) {
                        int sum = 0;
                        for (int a = 0; a < nr; ++a)
                            for (int b = 0; b < nc; ++b) {
                                const int yy = std::min(2 * v + a, Vs - 1), ii = std::min(2 * u + b, Us - 1) * cn + c;
                                sum += wide ? (int)S.ptr<ushort>(yy)[ii] : (int)S.ptr<uchar>(yy)[ii];
                            }
                        const int n = nr * nc;
                        int r;
                        if (n == 4) r = (sum + 2) >> 2;
                        else if (n == 2) r = (sum + ((sum >> 1) & 1)) >> 1;
                        else r = sum;
                        if (wide) out.ptr<ushort>(v)[u * cn + c] = (ushort)r; else out.ptr<uchar>(v)[u * cn + c] = (uchar)r;
                    }
                }
            }
        } else shim_abort("resize 0.5 on this depth");
    } else if (interpolation == INTER_LINEAR) {
        shim_check(S.type() == CV_32FC1, "resize INTER_LINEAR to dsize: CV_32FC1");
        std::vector<int> xi(Ud), yi(Vd);
        std::vector<float> xa(Ud), ya(Vd);
        const double sx = (double)Us / Ud, sy = (double)Vs / Vd;
        for (int x = 0; x < Ud; ++x) {
            float f = (float)((x + 0.5) * sx - 0.5);
            int ix = (int)std::floor(f); f -= ix;
            if (ix < 0) { ix = 0; f = 0.f; }
            if (ix >= Us - 1) { ix = Us - 1; f = 0.f; }
            xi[x] = ix; xa[x] = f;
        }
        for (int y = 0; y < Vd; ++y) {
            float f = (float)((y + 0.5) * sy - 0.5);
            int iy = (int)std::floor(f); f -= iy;
            if (iy < 0) { iy = 0; f = 0.f; }
            if (iy >= Vs - 1) { iy = Vs - 1; f = 0.f; }
            yi[y] = iy; ya[y] = f;
        }
        for (int y = 0; y < Vd; ++y) {
            const int y0 = yi[y], y1 = std::min(y0 + 1, Vs - 1);
            const float b1 = ya[y], b0 = 1.f - b1;
            for (int x = 0; x < Ud; ++x) {
                const int x0 = xi[x], x1 = std::min(x0 + 1, Us - 1);
                const float a1 = xa[x], a0 = 1.f - a1;
                float h0 = S.at<float>(y0, x0) * a0; { const float t = S.at<float>(y0, x1) * a1; h0 = h0 + t; }
                float h1 = S.at<float>(y1, x0) * a0; { const float t = S.at<float>(y1, x1) * a1; h1 = h1 + t; }
                float o = h0 * b0; { const float t = h1 * b1; o = o + t; }
                out.at<float>(y, x) = o;
            }
        }
    } else shim_abort("resize interpolation");
    dst = out;
}

/* cv::medianBlur(CV_32FC1, 3): 3x3 median, BORDER_REPLICATE */
inline void medianBlur(const Mat& src, Mat& dst, int ksize)
{
    shim_check(ksize == 3 && src.type() == CV_32FC1, "medianBlur: 3x3 CV_32FC1");
    Mat S = src;
    const int V = S.rows, U = S.cols;
    Mat out(V, U, S.type());
    for (int v = 0; v < V; ++v)
        for (int u = 0; u < U; ++u) {
            float w[9]; int n = 0;
            for (int dv = -1; dv <= 1; ++dv)
                for (int du = -1; du <= 1; ++du)
                    w[n++] = S.at<float>(std::min(std::max(v + dv, 0), V - 1), std::min(std::max(u + du, 0), U - 1));
            std::nth_element(w, w + 4, w + 9);
            out.at<float>(v, u) = w[4];
        }
    dst = out;
}

/* cv::getStructuringElement (anchor at the centre, ksize / 2): MORPH_RECT all ones; MORPH_CROSS the anchor's row and
 * column; MORPH_ELLIPSE row i covers [c - dx, c + dx] with dx = cvRound(c * sqrt((r^2 - (i - r)^2) / r^2)), r = c = ksize / 2.
 * Pinned against cv2 4.13 for sizes 1..15 (tests/test_oracle_vs_cv2.py). */
inline Mat getStructuringElement(int shape, Size ksize, Point anchor = Point(-1, -1))
{
    shim_check(shape == MORPH_RECT || shape == MORPH_CROSS || shape == MORPH_ELLIPSE, "getStructuringElement: shape");
    shim_check(anchor.x == -1 && anchor.y == -1, "getStructuringElement: default anchor");
    const int ax = ksize.width / 2, ay = ksize.height / 2;
    const int r = ksize.height / 2, c = ksize.width / 2;
    const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    Mat K(ksize.height, ksize.width, CV_8UC1);
    for (int i = 0; i < ksize.height; ++i) {
        int j1 = 0, j2 = 0;
        if (shape == MORPH_RECT || (shape == MORPH_CROSS && i == ay)) j2 = ksize.width;
        else if (shape == MORPH_CROSS) { j1 = ax; j2 = j1 + 1; }
        else {
            const int dy = i - r;
            if (std::abs(dy) <= r) {
                const int dx = detail::cv_round(c * std::sqrt((r * r - dy * dy) * inv_r2));
                j1 = std::max(c - dx, 0); j2 = std::min(c + dx + 1, ksize.width);
            }
        }
        uchar* row = K.ptr<uchar>(i);
        for (int j = 0; j < ksize.width; ++j) row[j] = (j >= j1 && j < j2) ? 1 : 0;
    }
    return K;
}

/* cv::morphologyEx(MORPH_OPEN) on CV_8UC1, one iteration, default anchor, BORDER_CONSTANT with
 * morphologyDefaultBorderValue(): erosion then dilation, dst(y, x) = min / max over the element's non-zero (ky, kx) of
 * src(y + ky - ay, x + kx - ax), positions outside the image ignored.  Pinned against cv2 4.13 (odd and even sizes). */
inline void morphologyEx(const Mat& src, Mat& dst, int op, const Mat& kernel)
{
    shim_check(op == MORPH_OPEN && src.type() == CV_8UC1 && kernel.type() == CV_8UC1, "morphologyEx: MORPH_OPEN on CV_8UC1");
    Mat S = src, K = kernel;
    const int V = S.rows, U = S.cols, ay = K.rows / 2, ax = K.cols / 2;
    auto pass = [&](const Mat& in, bool erode) {
        Mat out(V, U, CV_8UC1);
        for (int y = 0; y < V; ++y)
            for (int x = 0; x < U; ++x) {
                int acc = erode ? 255 : 0;
                for (int ky = 0; ky < K.rows; ++ky)
                    for (int kx = 0; kx < K.cols; ++kx) {
                        if (!K.ptr<uchar>(ky)[kx]) continue;
                        const int yy = y + ky - ay, xx = x + kx - ax;
                        if (yy < 0 || yy >= V || xx < 0 || xx >= U) continue;
                        const int v = in.ptr<uchar>(yy)[xx];
                        acc = erode ? std::min(acc, v) : std::max(acc, v);
                    }
                out.ptr<uchar>(y)[x] = (uchar)acc;
            }
        return out;
    };
    Mat out = pass(pass(S, true), false);
    if (dst.rows == V && dst.cols == U && dst.type() == CV_8UC1 && dst.data) out.copyTo(dst);
    else dst = out;
}
/* cv::applyColorMap: CV_8UC1 -> CV_8UC3 through a 256-entry BGR table.  OpenCV's tables are data of the real library:
 * the driver hands in the table it was given (cv2.applyColorMap of a 0..255 ramp, tests/golden/colormap_jet.npy). */
inline const uchar*& cvshim_colormap_lut() { static const uchar* lut = nullptr; return lut; }
inline void applyColorMap(const Mat& src, Mat& dst, int)
{
    const uchar* lut = cvshim_colormap_lut();
    shim_check(lut != nullptr, "applyColorMap: no colour table set");
    shim_check(src.type() == CV_8UC1, "applyColorMap: CV_8UC1");
    Mat S = src, out(S.rows, S.cols, CV_8UC3);
    for (int y = 0; y < S.rows; ++y)
        for (int x = 0; x < S.cols; ++x) {
            const uchar* e = lut + 3 * (int)S.ptr<uchar>(y)[x];
            uchar* o = out.ptr<uchar>(y) + 3 * x;
            o[0] = e[0]; o[1] = e[1]; o[2] = e[2];
        }
    dst = out;
}
inline void cvtColor(const Mat&, Mat&, int, int = 0) { shim_abort("cvtColor"); }
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) { shim_abort("line"); }

}  // namespace cv
#endif
