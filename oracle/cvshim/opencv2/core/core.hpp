/*
 * cvshim — a minimal stand-in for the part of the OpenCV 3.x C++ API that RSLightFields' depth path uses.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/).  The image this project is built in has no OpenCV C++ headers or
 * libraries, so the reference (/root/reference/RSLightFields, every header of which includes <opencv2/...>)
 * cannot be compiled against the real library.  With this directory on the include path the reference's OWN
 * sources compile unmodified, from where they lie, into oracle/_ref/librslf_ref.so (oracle/Makefile target
 * `ref`); the reference then supplies all the control flow of the path (pass order, masks, propagation,
 * bounds, pyramid, fusion) and this file supplies the OpenCV primitives it calls.
 *
 * Semantics: cv::Mat is a reference-counted 2-D array header with row/col views; assignment of an expression
 * (`m = a > t`, `m = a * b`) writes into the existing buffer when size and type already match, like
 * cv::MatExpr does — the reference relies on that to fill rows of larger maps.  Float arithmetic is float32
 * with one rounding per OpenCV operation, in the order OpenCV 3.x uses (stated per function below and pinned
 * against the cv2 4.13 wheel by tests/test_oracle_vs_cv2.py through the oracle, which restates the same
 * primitives).  Functions that the depth path never reaches (display, file I/O, DFT, colour maps, morphology)
 * are declared so that the headers compile and abort if they are ever called.
 */
#ifndef CVSHIM_CORE_HPP
#define CVSHIM_CORE_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <vector>

typedef unsigned char uchar;
typedef unsigned short ushort;

#define CV_CN_SHIFT 3
#define CV_DEPTH_MAX (1 << CV_CN_SHIFT)
#define CV_MAT_DEPTH_MASK (CV_DEPTH_MAX - 1)
#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAT_DEPTH(flags) ((flags) & CV_MAT_DEPTH_MASK)
#define CV_MAKETYPE(depth, cn) (CV_MAT_DEPTH(depth) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_MAT_CN(flags) ((((flags) >> CV_CN_SHIFT) & 511) + 1)
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16UC3 CV_MAKETYPE(CV_16U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_GRAY2RGB 8
#define CV_WINDOW_NORMAL 0

namespace cv {

[[noreturn]] inline void shim_abort(const char* what)
{
    std::fprintf(stderr, "cvshim: %s is outside the depth path and not implemented in the stand-in\n", what);
    std::abort();
}
inline void shim_check(bool ok, const char* what)
{
    if (!ok) { std::fprintf(stderr, "cvshim: assertion failed: %s\n", what); std::abort(); }
}

/* ---------------------------------------------------------------- small value types */
template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(N == 3, "3 values"); val[0] = a; val[1] = b; val[2] = c; }
    Vec(T a, T b) { static_assert(N == 2, "2 values"); val[0] = a; val[1] = b; }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
};
/* cv::Matx / cv::Vec arithmetic: element by element, saturate_cast<T> of the result (float: one rounding) */
template <typename T, int N> inline Vec<T, N> operator+(const Vec<T, N>& a, const Vec<T, N>& b) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = (T)(a[i] + b[i]); return r; }
template <typename T, int N> inline Vec<T, N> operator-(const Vec<T, N>& a, const Vec<T, N>& b) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = (T)(a[i] - b[i]); return r; }
template <typename T, int N> inline Vec<T, N> operator*(const Vec<T, N>& a, float s) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = (T)(a[i] * s); return r; }
template <typename T, int N> inline Vec<T, N> operator*(float s, const Vec<T, N>& a) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = (T)(a[i] * s); return r; }
template <typename T, int N> inline Vec<T, N> operator*(const Vec<T, N>& a, double s) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = (T)(a[i] * s); return r; }
template <typename T, int N> inline Vec<T, N> operator*(double s, const Vec<T, N>& a) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = (T)(a[i] * s); return r; }
template <typename T, int N> inline Vec<T, N> operator*(const Vec<T, N>& a, int s) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = (T)(a[i] * s); return r; }
template <typename T, int N> inline Vec<T, N> operator*(int s, const Vec<T, N>& a) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = (T)(a[i] * s); return r; }
template <typename T, int N> inline Vec<T, N> operator/(const Vec<T, N>& a, float s) { Vec<T, N> r; for (int i = 0; i < N; ++i) r[i] = (T)(a[i] / s); return r; }
template <typename T, int N> inline Vec<T, N>& operator+=(Vec<T, N>& a, const Vec<T, N>& b) { for (int i = 0; i < N; ++i) a[i] = (T)(a[i] + b[i]); return a; }
template <typename T, int N> inline bool operator==(const Vec<T, N>& a, const Vec<T, N>& b) { for (int i = 0; i < N; ++i) if (!(a[i] == b[i])) return false; return true; }
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 2> Vec2f;
typedef Vec<uchar, 3> Vec3b;
typedef Vec<ushort, 3> Vec3w;
/* cv::norm(Vec): normL2Sqr<T, double> — squares accumulated in double in index order, sqrt in double */
template <typename T, int N> inline double norm(const Vec<T, N>& v)
{
    double s = 0.0;
    for (int i = 0; i < N; ++i) s += (double)v[i] * (double)v[i];
    return std::sqrt(s);
}

struct Scalar {
    double val[4];
    Scalar() { val[0] = val[1] = val[2] = val[3] = 0; }
    Scalar(double v0) { val[0] = v0; val[1] = val[2] = val[3] = 0; }
    Scalar(double v0, double v1, double v2 = 0, double v3 = 0) { val[0] = v0; val[1] = v1; val[2] = v2; val[3] = v3; }
    static Scalar all(double v) { return Scalar(v, v, v, v); }
    double& operator[](int i) { return val[i]; }
    const double& operator[](int i) const { return val[i]; }
};
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size& o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size& o) const { return !(*this == o); }
    int area() const { return width * height; }
};
struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Range {
    int start, end;
    Range() : start(0), end(0) {}
    Range(int s, int e) : start(s), end(e) {}
    static Range all() { return Range(std::numeric_limits<int>::min(), std::numeric_limits<int>::max()); }
    bool is_all() const { return start == std::numeric_limits<int>::min(); }
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { REDUCE_SUM = 0, REDUCE_AVG = 1, REDUCE_MAX = 2, REDUCE_MIN = 3 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };
enum { CMP_EQ = 0, CMP_GT = 1, CMP_GE = 2, CMP_LT = 3, CMP_LE = 4, CMP_NE = 5 };
enum { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };
enum { MORPH_ERODE = 0, MORPH_DILATE = 1, MORPH_OPEN = 2, MORPH_CLOSE = 3 };
enum { COLORMAP_AUTUMN = 0, COLORMAP_BONE = 1, COLORMAP_JET = 2 };
/* OpenCV 3.x C-API constants still used by the reference (rslf_plot.cpp:71) */
#define CV_SORT_EVERY_ROW 0
#define CV_SORT_EVERY_COLUMN 1
#define CV_SORT_ASCENDING 0
#define CV_SORT_DESCENDING 16
enum { ROTATE_90_CLOCKWISE = 0, ROTATE_180 = 1, ROTATE_90_COUNTERCLOCKWISE = 2 };
enum { DFT_INVERSE = 1, DFT_SCALE = 2, DFT_ROWS = 4, DFT_COMPLEX_OUTPUT = 16, DFT_REAL_OUTPUT = 32 };
enum { WINDOW_NORMAL = 0, WINDOW_AUTOSIZE = 1 };
enum { COLOR_GRAY2RGB = 8, COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8 };
enum { SORT_EVERY_ROW = 0, SORT_EVERY_COLUMN = 1, SORT_ASCENDING = 0, SORT_DESCENDING = 16 };
enum { IMREAD_UNCHANGED = -1, IMREAD_GRAYSCALE = 0, IMREAD_COLOR = 1, IMREAD_ANYDEPTH = 2, IMREAD_ANYCOLOR = 4 };

class Mat;
class MatExpr;

/* the `size` member of cv::Mat: callable, comparable */
struct MatSize {
    const Mat* m;
    explicit MatSize(const Mat* m_) : m(m_) {}
    Size operator()() const;
    bool operator==(const MatSize& o) const;
    bool operator!=(const MatSize& o) const { return !(*this == o); }
    int operator[](int i) const;
};

class Mat {
public:
    int rows, cols;
    uchar* data;
    size_t step;                 /* bytes between rows */
    MatSize size;

    Mat() : rows(0), cols(0), data(nullptr), step(0), size(this), type_(0) {}
    Mat(int r, int c, int type) : rows(0), cols(0), data(nullptr), step(0), size(this), type_(0) { create(r, c, type); }
    Mat(Size s, int type) : rows(0), cols(0), data(nullptr), step(0), size(this), type_(0) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, const Scalar& s) : rows(0), cols(0), data(nullptr), step(0), size(this), type_(0) { create(r, c, type); setTo(s); }
    Mat(Size sz, int type, const Scalar& s) : rows(0), cols(0), data(nullptr), step(0), size(this), type_(0) { create(sz.height, sz.width, type); setTo(s); }
    /* user-owned data (no reference counting), like cv::Mat(rows, cols, type, void*, step) */
    Mat(int r, int c, int type, void* user, size_t step_ = 0)
        : rows(r), cols(c), data((uchar*)user), step(step_ ? step_ : (size_t)c * esz(type)), size(this), type_(type) {}
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), data(o.data), step(o.step), size(this), type_(o.type_), buf_(o.buf_) {}
    Mat(const MatExpr& e);
    Mat& operator=(const Mat& o)
    {
        if (this != &o) { rows = o.rows; cols = o.cols; data = o.data; step = o.step; type_ = o.type_; buf_ = o.buf_; }
        return *this;
    }
    Mat& operator=(const MatExpr& e);
    Mat& operator=(const Scalar& s) { setTo(s); return *this; }

    static size_t esz1(int type) { static const size_t t[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return t[CV_MAT_DEPTH(type)]; }
    static size_t esz(int type) { return esz1(type) * CV_MAT_CN(type); }

    /* cv::Mat::create: keeps the buffer when the geometry and type already match */
    void create(int r, int c, int type)
    {
        if (data && rows == r && cols == c && type_ == type) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * esz(type);
        const size_t bytes = (size_t)r * step;
        /* zero-filled: the reference accumulates into maps it never initialises (dc.hpp:497, :724-731) and relies on
         * large allocations being fresh zero pages; the stand-in makes that deterministic for small test sizes too */
        buf_ = std::shared_ptr<uchar>(bytes ? new uchar[bytes]() : nullptr, std::default_delete<uchar[]>());
        data = buf_.get();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { rows = cols = 0; data = nullptr; step = 0; buf_.reset(); }

    int type() const { return type_; }
    int depth() const { return CV_MAT_DEPTH(type_); }
    int channels() const { return CV_MAT_CN(type_); }
    size_t elemSize() const { return esz(type_); }
    size_t elemSize1() const { return esz1(type_); }
    size_t total() const { return (size_t)rows * cols; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return rows <= 1 || step == (size_t)cols * elemSize(); }

    Mat row(int y) const { Mat m(*this); m.rows = 1; m.data = data + (size_t)y * step; return m; }
    Mat col(int x) const { Mat m(*this); m.cols = 1; m.data = data + (size_t)x * elemSize(); return m; }
    Mat rowRange(int a, int b) const { Mat m(*this); m.rows = b - a; m.data = data + (size_t)a * step; return m; }
    Mat colRange(int a, int b) const { Mat m(*this); m.cols = b - a; m.data = data + (size_t)a * elemSize(); return m; }
    Mat rowRange(const Range& r) const { return r.is_all() ? *this : rowRange(r.start, r.end); }
    Mat colRange(const Range& r) const { return r.is_all() ? *this : colRange(r.start, r.end); }
    Mat operator()(const Range& rr, const Range& cr) const { return rowRange(rr).colRange(cr); }
    Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
    Mat operator()(const Range* ranges) const { return rowRange(ranges[0]).colRange(ranges[1]); }

    template <typename T> T* ptr(int y = 0) { return reinterpret_cast<T*>(data + (size_t)y * step); }
    template <typename T> const T* ptr(int y = 0) const { return reinterpret_cast<const T*>(data + (size_t)y * step); }
    uchar* ptr(int y = 0) { return data + (size_t)y * step; }
    const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }
    template <typename T> T& at(int y, int x) { return reinterpret_cast<T*>(data + (size_t)y * step)[x]; }
    template <typename T> const T& at(int y, int x) const { return reinterpret_cast<const T*>(data + (size_t)y * step)[x]; }
    /* single index: element i of a row or column vector (or of a continuous matrix in scan order) */
    template <typename T> T& at(int i)
    {
        if (rows == 1) return at<T>(0, i);
        if (cols == 1) return at<T>(i, 0);
        return at<T>(i / cols, i % cols);
    }
    template <typename T> const T& at(int i) const { return const_cast<Mat*>(this)->at<T>(i); }
    template <typename T> T& at(Point p) { return at<T>(p.y, p.x); }
    template <typename T> const T& at(Point p) const { return at<T>(p.y, p.x); }

    Mat clone() const { Mat m; copyTo(m); return m; }
    void copyTo(Mat& dst) const
    {
        if (dst.data == data && dst.rows == rows && dst.cols == cols && dst.step == step) return;
        dst.create(rows, cols, type_);
        const size_t rb = (size_t)cols * elemSize();
        for (int y = 0; y < rows; ++y) std::memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, rb);
    }
    void copyTo(Mat&& dst) const { Mat& d = dst; copyTo(d); }   /* row(...).copyTo(view.row(...)) */
    void copyTo(Mat& dst, const Mat& mask) const
    {
        shim_check(mask.type() == CV_8UC1 && mask.rows == rows && mask.cols == cols, "copyTo mask geometry");
        if (dst.rows != rows || dst.cols != cols || dst.type_ != type_) { dst.create(rows, cols, type_); dst.setTo(Scalar(0)); }
        const size_t es = elemSize();
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x)
                if (mask.at<uchar>(y, x)) std::memcpy(dst.data + (size_t)y * dst.step + x * es, data + (size_t)y * step + x * es, es);
    }
    void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const;
    Mat& setTo(const Scalar& s, const Mat& mask = Mat());
    /* same data, new channel count / row count (the matrix must be continuous when the row count changes) */
    Mat reshape(int cn, int new_rows = 0) const
    {
        Mat m(*this);
        if (cn == 0) cn = channels();
        const size_t row_elems1 = (size_t)cols * channels();          /* scalars per row */
        if (new_rows == 0 || new_rows == rows) {
            shim_check(row_elems1 % cn == 0, "reshape: channels");
            m.cols = (int)(row_elems1 / cn);
        } else {
            shim_check(isContinuous(), "reshape: continuous");
            const size_t tot1 = row_elems1 * rows;
            shim_check(tot1 % ((size_t)new_rows * cn) == 0, "reshape: rows");
            m.rows = new_rows;
            m.cols = (int)(tot1 / ((size_t)new_rows * cn));
            m.step = (size_t)m.cols * cn * elemSize1();
        }
        m.type_ = CV_MAKETYPE(depth(), cn);
        return m;
    }
    /* cv::Mat::forEach: f(element, position) over all elements (sequential here) */
    template <typename T, typename F> void forEach(const F& f)
    {
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) { const int pos[2] = {y, x}; f(at<T>(y, x), pos); }
    }
    static MatExpr zeros(int r, int c, int type);
    static MatExpr zeros(Size s, int type);
    static MatExpr ones(int r, int c, int type);
    MatExpr mul(const Mat& o, double scale = 1.0) const;
    MatExpr t() const;

private:
    int type_;
    std::shared_ptr<uchar> buf_;
};

inline Size MatSize::operator()() const { return Size(m->cols, m->rows); }
inline bool MatSize::operator==(const MatSize& o) const { return m->rows == o.m->rows && m->cols == o.m->cols; }
inline int MatSize::operator[](int i) const { return i == 0 ? m->rows : m->cols; }

template <typename T> class Mat_ : public Mat {
public:
    Mat_() : Mat() {}
    Mat_(const Mat& m) : Mat() { shim_abort("Mat_<T>(Mat)"); (void)m; }
};

/* a deferred operation that knows how to write itself into a destination (cv::MatExpr) */
class MatExpr {
public:
    std::function<void(Mat&)> eval;
    explicit MatExpr(std::function<void(Mat&)> f) : eval(std::move(f)) {}
};
inline Mat::Mat(const MatExpr& e) : rows(0), cols(0), data(nullptr), step(0), size(this), type_(0) { e.eval(*this); }
inline Mat& Mat::operator=(const MatExpr& e) { e.eval(*this); return *this; }

/* ---------------------------------------------------------------- element-wise helpers */
namespace detail {
template <typename F> inline void for_rows(const Mat& a, F f) { for (int y = 0; y < a.rows; ++y) f(y); }
inline void same(const Mat& a, const Mat& b, const char* what)
{
    shim_check(a.rows == b.rows && a.cols == b.cols && a.type() == b.type(), what);
}
/* dst may alias a source: element-wise ops read element i before writing element i, which is safe */
template <typename Op> inline void binary_f32(const Mat& a, const Mat& b, Mat& dst, Op op, const char* what)
{
    same(a, b, what);
    shim_check(a.depth() == CV_32F, what);
    Mat A = a, B = b;                     /* keep the sources alive if dst is reallocated */
    dst.create(A.rows, A.cols, A.type());
    const int n = A.cols * A.channels();
    for (int y = 0; y < A.rows; ++y) {
        const float* pa = A.ptr<float>(y); const float* pb = B.ptr<float>(y); float* pd = dst.ptr<float>(y);
        for (int i = 0; i < n; ++i) pd[i] = op(pa[i], pb[i]);
    }
}
template <typename Op> inline void unary_f32(const Mat& a, Mat& dst, Op op, const char* what)
{
    shim_check(a.depth() == CV_32F, what);
    Mat A = a;
    dst.create(A.rows, A.cols, A.type());
    const int n = A.cols * A.channels();
    for (int y = 0; y < A.rows; ++y) {
        const float* pa = A.ptr<float>(y); float* pd = dst.ptr<float>(y);
        for (int i = 0; i < n; ++i) pd[i] = op(pa[i]);
    }
}
}  // namespace detail

/* ---------------------------------------------------------------- core array operations */
/* cv::add / subtract / multiply / divide on CV_32F: one float operation per element.
 * multiply with scale != 1: (float(scale) * a) * b, left to right (OpenCV 3.x mul_<float, float>).
 * divide: b == 0 gives 0 (OpenCV 3.x div_<float>). */
inline void add(const Mat& a, const Mat& b, Mat& dst) { detail::binary_f32(a, b, dst, [](float x, float y) { return x + y; }, "add"); }
inline void subtract(const Mat& a, const Mat& b, Mat& dst) { detail::binary_f32(a, b, dst, [](float x, float y) { return x - y; }, "subtract"); }
inline void multiply(const Mat& a, const Mat& b, Mat& dst, double scale = 1.0)
{
    const float s = (float)scale;
    if (s == 1.0f) detail::binary_f32(a, b, dst, [](float x, float y) { return x * y; }, "multiply");
    else detail::binary_f32(a, b, dst, [s](float x, float y) { float t = s * x; return t * y; }, "multiply");
}
inline void multiply(const Mat& a, double s, Mat& dst)
{
    const float v = (float)s;
    detail::unary_f32(a, dst, [v](float x) { return x * v; }, "multiply scalar");
}
inline void multiply(const Mat& a, double s, Mat&& dst) { Mat& d = dst; multiply(a, s, d); }
inline void divide(const Mat& a, const Mat& b, Mat& dst)
{
    detail::binary_f32(a, b, dst, [](float x, float y) { return y != 0.f ? x / y : 0.f; }, "divide");
}
inline void add(const Mat& a, const Scalar& s, Mat& dst)
{
    const float v = (float)s[0];
    detail::unary_f32(a, dst, [v](float x) { return x + v; }, "add scalar");
}
inline void add(const Mat& a, double s, Mat& dst) { add(a, Scalar(s), dst); }
inline void subtract(const Scalar& s, const Mat& a, Mat& dst)
{
    const float v = (float)s[0];
    detail::unary_f32(a, dst, [v](float x) { return v - x; }, "scalar subtract");
}
inline void subtract(double s, const Mat& a, Mat& dst) { subtract(Scalar(s), a, dst); }
inline void subtract(const Mat& a, const Scalar& s, Mat& dst)
{
    const float v = (float)s[0];
    detail::unary_f32(a, dst, [v](float x) { return x - v; }, "subtract scalar");
}
/* masked add: dst(i) = a(i) + b(i) where mask(i) != 0, other elements of dst keep their value */
inline void add(const Mat& a, const Mat& b, Mat& dst, const Mat& mask)
{
    if (mask.empty()) { add(a, b, dst); return; }
    detail::same(a, b, "masked add");
    shim_check(mask.type() == CV_8UC1 && mask.rows == a.rows && mask.cols == a.cols, "masked add: mask");
    Mat A = a, B = b;
    if (dst.rows != A.rows || dst.cols != A.cols || dst.type() != A.type()) { dst.create(A.rows, A.cols, A.type()); dst.setTo(Scalar(0)); }
    const int cn = A.channels();
    if (A.depth() == CV_32F) {
        for (int y = 0; y < A.rows; ++y)
            for (int x = 0; x < A.cols; ++x)
                if (mask.at<uchar>(y, x))
                    for (int c = 0; c < cn; ++c) dst.ptr<float>(y)[x * cn + c] = A.ptr<float>(y)[x * cn + c] + B.ptr<float>(y)[x * cn + c];
    } else if (A.depth() == CV_8U) {
        for (int y = 0; y < A.rows; ++y)
            for (int x = 0; x < A.cols; ++x)
                if (mask.at<uchar>(y, x))
                    for (int c = 0; c < cn; ++c) {
                        int v = (int)A.ptr<uchar>(y)[x * cn + c] + (int)B.ptr<uchar>(y)[x * cn + c];
                        dst.ptr<uchar>(y)[x * cn + c] = (uchar)(v > 255 ? 255 : v);
                    }
    } else shim_abort("masked add on this depth");
}
inline void add(const Mat& a, double s, Mat&& dst, const Mat& mask)
{
    /* cv::add(expr, 0.0, view, mask): scalar second operand, destination is a row view */
    Mat b(a.rows, a.cols, a.type(), Scalar::all(s));
    Mat& d = dst;
    add(a, b, d, mask);
}
inline void add(const Mat& a, double s, Mat& dst, const Mat& mask)
{
    Mat b(a.rows, a.cols, a.type(), Scalar::all(s));
    add(a, b, dst, mask);
}
/* cv::max(src, scalar): the vector path of OpenCV returns the SECOND operand when the first is NaN
 * (maxps), which is what the reference relies on ("this operation removes nan's", kern.cpp:25, :53) */
inline void max(const Mat& a, const Scalar& s, Mat& dst)
{
    const float v = (float)s[0];
    detail::unary_f32(a, dst, [v](float x) { return x > v ? x : v; }, "max scalar");
}
inline void max(const Mat& a, double s, Mat& dst) { max(a, Scalar(s), dst); }
inline void max(const Mat& a, const Mat& b, Mat& dst) { detail::binary_f32(a, b, dst, [](float x, float y) { return x > y ? x : y; }, "max"); }
inline void min(const Mat& a, double s, Mat& dst)
{
    const float v = (float)s;
    detail::unary_f32(a, dst, [v](float x) { return x < v ? x : v; }, "min scalar");
}
/* cv::pow: power 2 is multiply(src, src) in OpenCV */
inline void pow(const Mat& a, double p, Mat& dst)
{
    shim_check(p == 2.0, "pow: only power 2 is on the path");
    multiply(a, a, dst);
}
inline void sqrt(const Mat& a, Mat& dst) { detail::unary_f32(a, dst, [](float x) { return std::sqrt(x); }, "sqrt"); }

/* cv::compare with a scalar: the scalar is converted to the array's depth; result CV_8UC1 0 / 255 */
inline void compare(const Mat& a, double s, Mat& dst, int op)
{
    shim_check(a.channels() == 1, "compare: single channel");
    Mat A = a;
    dst.create(A.rows, A.cols, CV_8UC1);
    for (int y = 0; y < A.rows; ++y) {
        uchar* pd = dst.ptr<uchar>(y);
        for (int x = 0; x < A.cols; ++x) {
            bool r = false;
            if (A.depth() == CV_32F) {
                const float v = A.ptr<float>(y)[x], t = (float)s;
                r = op == CMP_EQ ? v == t : op == CMP_GT ? v > t : op == CMP_GE ? v >= t : op == CMP_LT ? v < t : op == CMP_LE ? v <= t : v != t;
            } else if (A.depth() == CV_8U) {
                const double v = A.ptr<uchar>(y)[x];
                r = op == CMP_EQ ? v == s : op == CMP_GT ? v > s : op == CMP_GE ? v >= s : op == CMP_LT ? v < s : op == CMP_LE ? v <= s : v != s;
            } else shim_abort("compare on this depth");
            pd[x] = r ? 255 : 0;
        }
    }
}
inline void bitwise_and(const Mat& a, const Mat& b, Mat& dst)
{
    detail::same(a, b, "bitwise_and");
    Mat A = a, B = b;
    dst.create(A.rows, A.cols, A.type());
    const size_t n = (size_t)A.cols * A.elemSize();
    for (int y = 0; y < A.rows; ++y) for (size_t i = 0; i < n; ++i) dst.ptr(y)[i] = A.ptr(y)[i] & B.ptr(y)[i];
}
inline void bitwise_or(const Mat& a, const Mat& b, Mat& dst)
{
    detail::same(a, b, "bitwise_or");
    Mat A = a, B = b;
    dst.create(A.rows, A.cols, A.type());
    const size_t n = (size_t)A.cols * A.elemSize();
    for (int y = 0; y < A.rows; ++y) for (size_t i = 0; i < n; ++i) dst.ptr(y)[i] = A.ptr(y)[i] | B.ptr(y)[i];
}
inline void bitwise_not(const Mat& a, Mat& dst)
{
    Mat A = a;
    dst.create(A.rows, A.cols, A.type());
    const size_t n = (size_t)A.cols * A.elemSize();
    for (int y = 0; y < A.rows; ++y) for (size_t i = 0; i < n; ++i) dst.ptr(y)[i] = (uchar)~A.ptr(y)[i];
}
/* cv::findNonZero: scan order (rows, then columns) */
inline void findNonZero(const Mat& m, std::vector<Point>& out)
{
    shim_check(m.type() == CV_8UC1, "findNonZero: CV_8UC1");
    out.clear();
    for (int y = 0; y < m.rows; ++y) for (int x = 0; x < m.cols; ++x) if (m.at<uchar>(y, x)) out.push_back(Point(x, y));
}
inline int countNonZero(const Mat& m)
{
    int n = 0;
    for (int y = 0; y < m.rows; ++y) for (int x = 0; x < m.cols; ++x) if (m.at<uchar>(y, x)) ++n;
    return n;
}
/* cv::repeat: tile ny x nx times */
inline void repeat(const Mat& src, int ny, int nx, Mat& dst)
{
    Mat S = src;
    dst.create(S.rows * ny, S.cols * nx, S.type());
    const size_t rb = (size_t)S.cols * S.elemSize();
    for (int y = 0; y < dst.rows; ++y)
        for (int k = 0; k < nx; ++k) std::memcpy(dst.ptr(y) + k * rb, S.ptr(y % S.rows), rb);
}
/* cv::reduce(SUM) on CV_32F -> CV_32F.  dim 0 (reduceR_<float, float>): the accumulator row starts as row 0
 * and rows are added in ascending order, in float.  dim 1: per row and channel, columns added in ascending order. */
inline void reduce(const Mat& src, Mat& dst, int dim, int rtype, int dtype = -1)
{
    (void)dtype;
    shim_check(rtype == REDUCE_SUM && src.depth() == CV_32F, "reduce: SUM of CV_32F");
    Mat S = src;
    const int cn = S.channels();
    if (dim == 0) {
        Mat out(1, S.cols, S.type());
        const int n = S.cols * cn;
        float* acc = out.ptr<float>(0);
        for (int i = 0; i < n; ++i) acc[i] = S.ptr<float>(0)[i];
        for (int y = 1; y < S.rows; ++y) { const float* p = S.ptr<float>(y); for (int i = 0; i < n; ++i) acc[i] = acc[i] + p[i]; }
        dst = out;
    } else {
        Mat out(S.rows, 1, S.type());
        for (int y = 0; y < S.rows; ++y) {
            const float* p = S.ptr<float>(y);
            for (int k = 0; k < cn; ++k) {
                float a0 = p[k];
                for (int x = 1; x < S.cols; ++x) a0 = a0 + p[x * cn + k];
                out.ptr<float>(y)[k] = a0;
            }
        }
        dst = out;
    }
}
/* cv::minMaxLoc: first occurrence of the extreme values in scan order */
inline void minMaxLoc(const Mat& m, double* minVal, double* maxVal = nullptr, Point* minLoc = nullptr, Point* maxLoc = nullptr)
{
    /* several channels are accepted when no location is asked for (the array is then scanned as scalars) */
    shim_check(!m.empty() && (m.channels() == 1 || (!minLoc && !maxLoc)), "minMaxLoc: single channel");
    double mn = 0, mx = 0; Point pmn(0, 0), pmx(0, 0); bool first = true;
    const int ncol = m.cols * m.channels();
    for (int y = 0; y < m.rows; ++y)
        for (int x = 0; x < ncol; ++x) {
            double v;
            switch (m.depth()) {
                case CV_32F: v = m.ptr<float>(y)[x]; break;
                case CV_8U: v = m.ptr<uchar>(y)[x]; break;
                case CV_16U: v = m.ptr<ushort>(y)[x]; break;
                case CV_64F: v = m.ptr<double>(y)[x]; break;
                default: shim_abort("minMaxLoc on this depth");
            }
            if (first) { mn = mx = v; pmn = pmx = Point(x, y); first = false; continue; }
            if (v < mn) { mn = v; pmn = Point(x, y); }
            if (v > mx) { mx = v; pmx = Point(x, y); }
        }
    if (minVal) *minVal = mn;
    if (maxVal) *maxVal = mx;
    if (minLoc) *minLoc = pmn;
    if (maxLoc) *maxLoc = pmx;
}
/* cv::mean: sum accumulated in double in scan order / number of elements */
inline Scalar mean(const Mat& m)
{
    Scalar s;
    const int cn = m.channels();
    shim_check(m.depth() == CV_32F && cn <= 4, "mean: CV_32F");
    for (int y = 0; y < m.rows; ++y)
        for (int x = 0; x < m.cols; ++x)
            for (int c = 0; c < cn; ++c) s[c] += (double)m.ptr<float>(y)[x * cn + c];
    const double n = (double)m.total();
    for (int c = 0; c < cn; ++c) s[c] = n > 0 ? s[c] / n : 0.0;
    return s;
}
inline void split(const Mat& src, std::vector<Mat>& planes)
{
    const int cn = src.channels();
    const int d = src.depth();
    const size_t e1 = src.elemSize1();
    planes.resize(cn);
    for (int c = 0; c < cn; ++c) {
        planes[c].create(src.rows, src.cols, CV_MAKETYPE(d, 1));
        if (e1 == 4) {
            for (int y = 0; y < src.rows; ++y) {
                const uint32_t* q = src.ptr<uint32_t>(y);
                uint32_t* o = planes[c].ptr<uint32_t>(y);
                for (int x = 0; x < src.cols; ++x) o[x] = q[(size_t)x * cn + c];
            }
            continue;
        }
        for (int y = 0; y < src.rows; ++y)
            for (int x = 0; x < src.cols; ++x) std::memcpy(planes[c].ptr(y) + x * e1, src.ptr(y) + ((size_t)x * cn + c) * e1, e1);
    }
}
inline void split(const Mat& src, Mat* planes)
{
    std::vector<Mat> v; split(src, v);
    for (size_t c = 0; c < v.size(); ++c) planes[c] = v[c];
}
inline void merge(const Mat* planes, size_t n, Mat& dst)
{
    shim_check(n > 0, "merge");
    std::vector<Mat> P(planes, planes + n);       /* keep the planes alive if dst aliases one */
    const int d = P[0].depth();
    const size_t e1 = P[0].elemSize1();
    for (size_t c = 0; c < n; ++c) shim_check(P[c].channels() == 1 && P[c].rows == P[0].rows && P[c].cols == P[0].cols && P[c].depth() == d, "merge planes");
    dst.create(P[0].rows, P[0].cols, CV_MAKETYPE(d, (int)n));
    if (e1 == 4) {                                   /* 32-bit elements: typed loops */
        for (int y = 0; y < dst.rows; ++y) {
            uint32_t* o = dst.ptr<uint32_t>(y);
            for (size_t c = 0; c < n; ++c) {
                const uint32_t* q = P[c].ptr<uint32_t>(y);
                for (int x = 0; x < dst.cols; ++x) o[(size_t)x * n + c] = q[x];
            }
        }
        return;
    }
    for (size_t c = 0; c < n; ++c)
        for (int y = 0; y < dst.rows; ++y)
            for (int x = 0; x < dst.cols; ++x) std::memcpy(dst.ptr(y) + ((size_t)x * n + c) * e1, P[c].ptr(y) + x * e1, e1);
}
inline void merge(const std::vector<Mat>& planes, Mat& dst) { merge(planes.data(), planes.size(), dst); }
/* cv::gemm for CV_32F (GEMMSingleMul<float, double>): products accumulated in double, result cast to float */
inline void gemm_f32(const Mat& a, const Mat& b, Mat& dst)
{
    shim_check(a.type() == CV_32FC1 && b.type() == CV_32FC1 && a.cols == b.rows, "gemm: CV_32FC1");
    Mat A = a, B = b;
    Mat out(A.rows, B.cols, CV_32FC1);
    for (int i = 0; i < A.rows; ++i)
        for (int j = 0; j < B.cols; ++j) {
            double s = 0.0;
            for (int k = 0; k < A.cols; ++k) s += (double)A.at<float>(i, k) * (double)B.at<float>(k, j);
            out.at<float>(i, j) = (float)s;
        }
    if (dst.rows == out.rows && dst.cols == out.cols && dst.type() == out.type() && dst.data) out.copyTo(dst);
    else dst = out;
}
inline void transpose(const Mat& src, Mat& dst)
{
    Mat S = src;
    Mat out(S.cols, S.rows, S.type());
    const size_t es = S.elemSize();
    for (int y = 0; y < S.rows; ++y) for (int x = 0; x < S.cols; ++x) std::memcpy(out.ptr(x) + y * es, S.ptr(y) + x * es, es);
    dst = out;
}
inline void vconcat(const Mat& a, const Mat& b, Mat& dst)
{
    shim_check(a.cols == b.cols && a.type() == b.type(), "vconcat");
    Mat A = a, B = b;
    Mat out(A.rows + B.rows, A.cols, A.type());
    const size_t rb = (size_t)A.cols * A.elemSize();
    for (int y = 0; y < A.rows; ++y) std::memcpy(out.ptr(y), A.ptr(y), rb);
    for (int y = 0; y < B.rows; ++y) std::memcpy(out.ptr(A.rows + y), B.ptr(y), rb);
    dst = out;
}
inline void hconcat(const Mat& a, const Mat& b, Mat& dst)
{
    shim_check(a.rows == b.rows && a.type() == b.type(), "hconcat");
    Mat A = a, B = b;
    Mat out(A.rows, A.cols + B.cols, A.type());
    const size_t ra = (size_t)A.cols * A.elemSize(), rb = (size_t)B.cols * B.elemSize();
    for (int y = 0; y < A.rows; ++y) { std::memcpy(out.ptr(y), A.ptr(y), ra); std::memcpy(out.ptr(y) + ra, B.ptr(y), rb); }
    dst = out;
}

/* Mat::convertTo: dst = saturate_cast<T>(src * alpha + beta); float -> float in float (cvtScale_<float, float, float>),
 * 8-bit / 16-bit -> float in float as well (cvtScale_<uchar, float, float>) */
inline void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const
{
    Mat S = *this;
    const int ddepth = rtype < 0 ? depth() : CV_MAT_DEPTH(rtype);
    const int cn = channels();
    Mat out;
    if (dst.data == data && ddepth == depth()) out = dst; else out.create(rows, cols, CV_MAKETYPE(ddepth, cn));
    const int n = cols * cn;
    const float a = (float)alpha, b = (float)beta;
    const bool plain = (alpha == 1.0 && beta == 0.0);
    for (int y = 0; y < rows; ++y) {
        for (int i = 0; i < n; ++i) {
            float v;
            switch (S.depth()) {
                case CV_8U: v = (float)S.ptr<uchar>(y)[i]; break;
                case CV_16U: v = (float)S.ptr<ushort>(y)[i]; break;
                case CV_32F: v = S.ptr<float>(y)[i]; break;
                case CV_32S: v = (float)S.ptr<int>(y)[i]; break;
                default: shim_abort("convertTo from this depth");
            }
            if (!plain) { v = v * a; v = v + b; }
            switch (ddepth) {
                case CV_32F: out.ptr<float>(y)[i] = v; break;
                case CV_8U: { int r = (int)std::nearbyint(v); out.ptr<uchar>(y)[i] = (uchar)(r < 0 ? 0 : r > 255 ? 255 : r); break; }
                default: shim_abort("convertTo to this depth");
            }
        }
    }
    dst = out;
}
inline Mat& Mat::setTo(const Scalar& s, const Mat& mask)
{
    if (!mask.empty()) shim_check(mask.type() == CV_8UC1 && mask.rows == rows && mask.cols == cols, "setTo mask");
    const int cn = channels();
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            if (!mask.empty() && !mask.at<uchar>(y, x)) continue;
            for (int c = 0; c < cn; ++c) {
                const double v = s[c < 4 ? c : 3];
                switch (depth()) {
                    case CV_32F: ptr<float>(y)[x * cn + c] = (float)v; break;
                    case CV_8U: { int r = (int)std::nearbyint(v); ptr<uchar>(y)[x * cn + c] = (uchar)(r < 0 ? 0 : r > 255 ? 255 : r); break; }
                    case CV_32S: ptr<int>(y)[x * cn + c] = (int)std::nearbyint(v); break;
                    case CV_64F: ptr<double>(y)[x * cn + c] = v; break;
                    case CV_16U: { int r = (int)std::nearbyint(v); ptr<ushort>(y)[x * cn + c] = (ushort)(r < 0 ? 0 : r > 65535 ? 65535 : r); break; }
                    default: shim_abort("setTo on this depth");
                }
            }
        }
    return *this;
}
inline MatExpr Mat::zeros(int r, int c, int type) { return MatExpr([=](Mat& d) { d.create(r, c, type); d.setTo(Scalar::all(0)); }); }
inline MatExpr Mat::zeros(Size s, int type) { return zeros(s.height, s.width, type); }
inline MatExpr Mat::ones(int r, int c, int type) { return MatExpr([=](Mat& d) { d.create(r, c, type); d.setTo(Scalar(1)); }); }
inline MatExpr Mat::mul(const Mat& o, double scale) const { Mat a = *this, b = o; return MatExpr([=](Mat& d) { multiply(a, b, d, scale); }); }
inline MatExpr Mat::t() const { Mat a = *this; return MatExpr([=](Mat& d) { transpose(a, d); }); }

/* ---------------------------------------------------------------- expression operators */
inline MatExpr operator*(const Mat& a, const Mat& b) { return MatExpr([=](Mat& d) { gemm_f32(a, b, d); }); }
inline MatExpr operator*(const Mat& a, double s) { return MatExpr([=](Mat& d) { a.convertTo(d, -1, s, 0.0); }); }
inline MatExpr operator*(double s, const Mat& a) { return a * s; }
inline MatExpr operator+(const Mat& a, const Mat& b) { return MatExpr([=](Mat& d) { add(a, b, d); }); }
inline MatExpr operator-(const Mat& a, const Mat& b) { return MatExpr([=](Mat& d) { subtract(a, b, d); }); }
inline MatExpr operator/(const Mat& a, const Mat& b) { return MatExpr([=](Mat& d) { divide(a, b, d); }); }
inline MatExpr operator+(const Mat& a, double s) { return MatExpr([=](Mat& d) { add(a, Scalar(s), d); }); }
inline MatExpr operator-(const Mat& a, double s) { return MatExpr([=](Mat& d) { subtract(a, Scalar(s), d); }); }
inline MatExpr operator>(const Mat& a, double s) { return MatExpr([=](Mat& d) { compare(a, s, d, CMP_GT); }); }
inline MatExpr operator>=(const Mat& a, double s) { return MatExpr([=](Mat& d) { compare(a, s, d, CMP_GE); }); }
inline MatExpr operator<(const Mat& a, double s) { return MatExpr([=](Mat& d) { compare(a, s, d, CMP_LT); }); }
inline MatExpr operator<=(const Mat& a, double s) { return MatExpr([=](Mat& d) { compare(a, s, d, CMP_LE); }); }
inline MatExpr operator==(const Mat& a, double s) { return MatExpr([=](Mat& d) { compare(a, s, d, CMP_EQ); }); }
inline MatExpr operator!=(const Mat& a, double s) { return MatExpr([=](Mat& d) { compare(a, s, d, CMP_NE); }); }
inline MatExpr operator/(const MatExpr& a, const MatExpr& b) { Mat A = a, B = b; return A / B; }
/* a *= s is a.convertTo(a, -1, s) in OpenCV: x * float(s) + 0 in float */
inline Mat& operator*=(Mat& a, double s) { a.convertTo(a, -1, s, 0.0); return a; }
inline Mat& operator+=(Mat& a, double s) { add(a, Scalar(s), a); return a; }
inline Mat& operator-=(Mat& a, double s) { subtract(a, Scalar(s), a); return a; }
inline Mat& operator+=(Mat& a, const Mat& b) { add(a, b, a); return a; }
inline Mat& operator-=(Mat& a, const Mat& b) { subtract(a, b, a); return a; }
inline Mat& operator/=(Mat& a, const Mat& b) { divide(a, b, a); return a; }
inline Mat& operator/=(Mat& a, double s) { a.convertTo(a, -1, 1.0 / s, 0.0); return a; }

inline std::ostream& operator<<(std::ostream& os, const Mat& m) { return os << "[Mat " << m.rows << "x" << m.cols << "]"; }
inline std::ostream& operator<<(std::ostream& os, const Size& s) { return os << "[" << s.width << " x " << s.height << "]"; }
template <typename T> inline std::ostream& operator<<(std::ostream& os, const Point_<T>& p) { return os << "[" << p.x << ", " << p.y << "]"; }

/* ---------------------------------------------------------------- declared, outside the depth path */
/* cv::meanStdDev of a single-channel CV_32F image: double sums, std = sqrt(max(sum(x^2) / N - mean^2, 0)) */
inline void meanStdDev(const Mat& m, Scalar& mean, Scalar& stddev)
{
    shim_check(m.type() == CV_32FC1 && !m.empty(), "meanStdDev: CV_32FC1");
    double s = 0, sq = 0;
    for (int y = 0; y < m.rows; ++y)
        for (int x = 0; x < m.cols; ++x) { const double v = m.ptr<float>(y)[x]; s += v; sq += v * v; }
    const double n = (double)m.rows * m.cols, mu = s / n;
    mean = Scalar(mu); stddev = Scalar(std::sqrt(std::max(sq / n - mu * mu, 0.0)));
}
/* cv::sort: every column of a CV_32FC1 matrix, ascending (rslf_plot.cpp:71) */
inline void sort(const Mat& src, Mat& dst, int flags)
{
    shim_check(src.type() == CV_32FC1 && flags == (CV_SORT_EVERY_COLUMN + CV_SORT_ASCENDING), "sort: columns of CV_32FC1, ascending");
    Mat S = src, out(S.rows, S.cols, CV_32FC1);
    std::vector<float> col(S.rows);
    for (int x = 0; x < S.cols; ++x) {
        for (int y = 0; y < S.rows; ++y) col[y] = S.ptr<float>(y)[x];
        std::sort(col.begin(), col.end());
        for (int y = 0; y < S.rows; ++y) out.ptr<float>(y)[x] = col[y];
    }
    dst = out;
}
inline int getOptimalDFTSize(int) { shim_abort("getOptimalDFTSize"); }
inline void dft(const Mat&, Mat&, int = 0, int = 0) { shim_abort("dft"); }
inline void idft(const Mat&, Mat&, int = 0, int = 0) { shim_abort("idft"); }
inline void copyMakeBorder(const Mat&, Mat&, int, int, int, int, int, const Scalar& = Scalar()) { shim_abort("copyMakeBorder"); }
inline void rotate(const Mat&, Mat&, int) { shim_abort("rotate"); }
inline void flip(const Mat&, Mat&, int) { shim_abort("flip"); }
inline void normalize(const Mat&, Mat&, double = 1, double = 0, int = 4, int = -1) { shim_abort("normalize"); }
inline void mulSpectrums(const Mat&, const Mat&, Mat&, int, bool = false) { shim_abort("mulSpectrums"); }
inline void magnitude(const Mat&, const Mat&, Mat&) { shim_abort("magnitude"); }

class FileNode {
public:
    template <typename T> void operator>>(T&) const { shim_abort("FileStorage"); }
    bool empty() const { return true; }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) { shim_abort("FileStorage"); }
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const std::string&) const { shim_abort("FileStorage"); }
    FileNode operator[](const char*) const { shim_abort("FileStorage"); }
    template <typename T> FileStorage& operator<<(const T&) { shim_abort("FileStorage"); }
};

}  // namespace cv

#endif
