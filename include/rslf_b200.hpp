/*
 * rslf_b200.hpp — C++ host side of the B200-native EPI depth path.
 *
 * Header-only mirror of the reference's depth-computer / fine-to-coarse classes
 * (RSLightFields/include/rslf_depth_computation.hpp: Depth1DComputer_pile :93-143,
 * Depth2DComputer :166-229; rslf_fine_to_coarse.hpp: FineToCoarse :26-81;
 * parameter block rslf_depth_computation_core.hpp:66-142) on top of the C ABI of
 * rslf_b200.h.  Same class names, constructor arguments, run() / getters and
 * result members, in namespace rslf_b200 so both can live in one program; a
 * maintainer switches a call site by changing the namespace.
 *
 * cv::Mat: when <opencv2/core.hpp> is available the classes take and return
 * cv::Mat / std::vector<cv::Mat> exactly like the reference.  This build image
 * has no OpenCV C++ package, so otherwise a minimal layout-compatible stand-in
 * (rows, cols, type, data, step, reference-counted storage) is used.
 *
 * Errors: the reference has no error codes of its own (only cv::Exception from
 * OpenCV asserts); here every failing C-ABI call throws rslf_b200::Error carrying
 * the code and the library's message.  There is no CPU fallback.
 */
#ifndef RSLF_B200_HPP
#define RSLF_B200_HPP

#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "rslf_b200.h"

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>) && !defined(RSLF_B200_NO_OPENCV)
#include <opencv2/core.hpp>
#define RSLF_B200_HAVE_OPENCV 1
#if __has_include(<opencv2/imgproc.hpp>)
#include <opencv2/imgproc.hpp>              /* cv::applyColorMap, for get_coloured_depth_maps' colour table */
#define RSLF_B200_HAVE_OPENCV_IMGPROC 1
#endif
#endif
#endif

namespace rslf_b200
{

#ifdef RSLF_B200_HAVE_OPENCV
using Mat = cv::Mat;
using Vec3f = cv::Vec3f;
#else
/* ---- minimal cv::Mat stand-in (layout-compatible subset) ---- */
#ifndef CV_8U
#define CV_8U 0
#define CV_16U 2
#define CV_32F 5
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) (((depth) & 7) + (((cn) - 1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_16UC3 CV_MAKETYPE(CV_16U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#endif
struct Vec3f { float val[3]; float& operator[](int i) { return val[i]; } const float& operator[](int i) const { return val[i]; } };

class Mat
{
public:
    int rows = 0, cols = 0;
    unsigned char* data = nullptr;
    size_t step = 0;                       /* bytes between rows */

    Mat() {}
    Mat(int a_rows, int a_cols, int a_type) { create(a_rows, a_cols, a_type); }
    /* header over user memory (not owned), like cv::Mat(rows, cols, type, data, step) */
    Mat(int a_rows, int a_cols, int a_type, void* a_data, size_t a_step = 0)
        : rows(a_rows), cols(a_cols), data((unsigned char*)a_data), m_type(a_type)
    {
        step = a_step ? a_step : (size_t)a_cols * elemSize();
    }
    void create(int a_rows, int a_cols, int a_type)
    {
        rows = a_rows; cols = a_cols; m_type = a_type;
        step = (size_t)cols * elemSize();
        /* like cv::Mat::create: uninitialised storage */
        m_store = std::shared_ptr<unsigned char>(new unsigned char[step * (size_t)rows + 1], std::default_delete<unsigned char[]>());
        data = m_store.get();
    }
    static Mat zeros(int a_rows, int a_cols, int a_type) { Mat m(a_rows, a_cols, a_type); std::memset(m.data, 0, m.step * (size_t)a_rows); return m; }
    /* storage in page-locked memory (rslf_host_alloc): the result images of the computers, filled by DMA */
    static Mat pinned(int a_rows, int a_cols, int a_type)
    {
        Mat m; m.rows = a_rows; m.cols = a_cols; m.m_type = a_type;
        m.step = (size_t)a_cols * m.elemSize();
        void* p = rslf_host_alloc(m.step * (size_t)a_rows);
        if (!p) { m.create(a_rows, a_cols, a_type); return m; }
        std::memset(p, 0, m.step * (size_t)a_rows);
        m.m_store = std::shared_ptr<unsigned char>((unsigned char*)p, [](unsigned char* q) { rslf_host_free(q); });
        m.data = m.m_store.get();
        return m;
    }
    int type() const { return m_type; }
    int depth() const { return m_type & 7; }
    int channels() const { return (m_type >> CV_CN_SHIFT) + 1; }
    size_t elemSize1() const { return depth() == CV_8U ? 1 : depth() == CV_16U ? 2 : 4; }
    size_t elemSize() const { return elemSize1() * channels(); }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    bool isContinuous() const { return step == (size_t)cols * elemSize(); }
    template <typename T> T* ptr(int r = 0) { return (T*)(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + (size_t)r * step); }
    template <typename T> T& at(int r, int c) { return ptr<T>(r)[c]; }
    template <typename T> const T& at(int r, int c) const { return ptr<T>(r)[c]; }
    Mat clone() const
    {
        Mat m(rows, cols, m_type);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
private:
    int m_type = 0;
    std::shared_ptr<unsigned char> m_store;
};
#endif /* RSLF_B200_HAVE_OPENCV */

template <typename T> using Vec = std::vector<T>;

class Error : public std::runtime_error
{
public:
    Error(int a_code, const std::string& a_what) : std::runtime_error(a_what), code(a_code) {}
    int code;
};

/* channel count of the template argument: float -> 1, Vec3f -> 3 */
template <typename DataType> struct channels_of;
template <> struct channels_of<float> { enum { value = 1 }; };
template <> struct channels_of<Vec3f> { enum { value = 3 }; };

/*
 * rslf::Depth1DParameters<DataType> (core.hpp:66-142).  The interpolation class
 * is always the linear one and the kernel class BandwidthKernel(h) — the only
 * ones the reference instantiates (core.hpp:76-78) — so the two class pointers
 * are replaced by the bandwidth h.
 */
template <typename DataType>
struct Depth1DParameters
{
    Depth1DParameters()
    {
        rslf_params p;
        rslf_params_default(&p);
        par_edge_score_threshold = p.edge_score_threshold;
        par_line_score_threshold = p.line_score_threshold;
        par_disp_score_threshold = p.disp_score_threshold;
        par_raw_score_threshold = p.raw_score_threshold;
        par_mean_shift_max_iter = (float)p.mean_shift_max_iter;
        par_edge_confidence_filter_size = p.edge_confidence_filter_size;
        par_edge_confidence_opening_type = p.edge_confidence_opening_type;
        par_edge_confidence_opening_size = p.edge_confidence_opening_size;
        par_median_filter_size = p.median_filter_size;
        par_median_filter_epsilon = p.median_filter_epsilon;
        par_propagation_epsilon = p.propagation_epsilon;
        par_slope_factor = p.slope_factor;
        par_cut_shadows = p.cut_shadows != 0;
        par_shadow_level = p.shadow_level;
        par_kernel_bandwidth = p.kernel_h;
    }
    float par_edge_score_threshold;
    float par_line_score_threshold;
    float par_disp_score_threshold;
    float par_raw_score_threshold;
    float par_mean_shift_max_iter;        /* a float in the reference too (core.hpp:115) */
    int par_edge_confidence_filter_size;
    int par_edge_confidence_opening_type;
    int par_edge_confidence_opening_size;
    int par_median_filter_size;
    float par_median_filter_epsilon;
    float par_propagation_epsilon;
    float par_slope_factor;
    bool par_cut_shadows;
    float par_shadow_level;
    float par_kernel_bandwidth;           /* BandwidthKernel(h), rslf_kernels.hpp:43 */

    static Depth1DParameters& get_default() { static Depth1DParameters s_default; return s_default; }

    rslf_params to_abi() const
    {
        rslf_params p;
        p.edge_score_threshold = par_edge_score_threshold; p.line_score_threshold = par_line_score_threshold;
        p.disp_score_threshold = par_disp_score_threshold; p.raw_score_threshold = par_raw_score_threshold;
        p.mean_shift_max_iter = (int)par_mean_shift_max_iter;
        p.edge_confidence_filter_size = par_edge_confidence_filter_size;
        p.edge_confidence_opening_type = par_edge_confidence_opening_type;
        p.edge_confidence_opening_size = par_edge_confidence_opening_size;
        p.median_filter_size = par_median_filter_size; p.median_filter_epsilon = par_median_filter_epsilon;
        p.propagation_epsilon = par_propagation_epsilon; p.slope_factor = par_slope_factor;
        p.cut_shadows = par_cut_shadows ? 1 : 0; p.shadow_level = par_shadow_level; p.kernel_h = par_kernel_bandwidth;
        return p;
    }
};

namespace detail
{
/* one rslf_ctx with RAII + error translation */
class Device
{
public:
    explicit Device(int a_device = 0) { check(rslf_cuda_create(a_device, &m_ctx), "rslf_cuda_create"); }
    ~Device() { if (m_ctx) rslf_cuda_destroy(m_ctx); }
    Device(const Device&) = delete;
    Device& operator=(const Device&) = delete;
    rslf_ctx* get() const { return m_ctx; }
    void check(int a_rc, const char* a_what) const
    {
        if (a_rc == RSLF_OK) return;
        std::string msg = std::string(a_what) + ": " + rslf_cuda_strerror(a_rc);
        if (m_ctx) { msg += ": "; msg += rslf_cuda_last_error_text(m_ctx); }
        throw Error(a_rc, msg);
    }
    /* the reference ctors' input handling (dc.hpp:425-477, 651-704): V Mats of S x U x C */
    void upload(const Vec<Mat>& a_epis, float a_scale, int a_channels, int& V, int& S, int& U, const rslf_params* a_params = nullptr)
    {
        if (a_epis.empty() || a_epis[0].empty()) throw Error(RSLF_ERR_ARG, "empty EPI vector");
        V = (int)a_epis.size(); S = a_epis[0].rows; U = a_epis[0].cols;
        const int depth = a_epis[0].depth();
        std::vector<const void*> ptrs(V);
        for (int v = 0; v < V; ++v) {
            const Mat& m = a_epis[v];
            if (m.rows != S || m.cols != U || m.channels() != a_channels || m.depth() != depth || (size_t)m.step != (size_t)a_epis[0].step)
                throw Error(RSLF_ERR_ARG, "EPIs must share size, type and step, and match the class's channel count");
            ptrs[v] = m.data;
        }
        /* with the computer's parameters: pipelined ingest (normalisation and level-0 edge confidence overlap the upload) */
        if (a_params) check(rslf_cuda_upload_epis_pipelined(m_ctx, ptrs.data(), V, S, U, a_channels, depth, (size_t)a_epis[0].step, a_scale, a_params),
                            "rslf_cuda_upload_epis_pipelined");
        else check(rslf_cuda_upload_epis(m_ctx, ptrs.data(), V, S, U, a_channels, depth, (size_t)a_epis[0].step, a_scale),
                   "rslf_cuda_upload_epis");
    }
    /* Row-sharded run, one process per GPU: this object was built from rows [row_starts[rank], row_starts[rank + 1])
     * of the light field; nccl_id = the 128 bytes of rslf_b200::nccl_unique_id() made by rank 0 and sent to all ranks. */
    bool sharded = false;
    void shard(const std::vector<int>& a_row_starts, int a_rank, const void* a_nccl_id)
    {
        const int world = (int)a_row_starts.size() - 1;
        sharded = true;
        check(rslf_cuda_comm_init(m_ctx, a_nccl_id, a_rank, world), "rslf_cuda_comm_init");
        check(rslf_cuda_set_row_shards(m_ctx, a_row_starts.data(), world), "rslf_cuda_set_row_shards");
    }
private:
    rslf_ctx* m_ctx = nullptr;
};

/* S result images (uninitialised: the library overwrites every row).  Ordinary pageable storage, like the
 * reference's result Mats: the library fills the image rows itself through its pinned ring
 * (rslf_cuda_*_get_mats), with no intermediate copy on this side.  Callers that keep their images across runs can
 * allocate them page-locked (rslf_host_alloc; Mat::pinned in the stand-in) and are then served by plain DMA. */
inline Vec<Mat> planes(int S, int V, int U, int type)
{
    Vec<Mat> v(S);
    for (auto& m : v) m = Mat(V, U, type);
    return v;
}

/* The reference constructs one computer object per light field.  A device context owns some tens of gigabytes of
 * device buffers and the pinned transfer ring, which it re-uses from run to run when the geometry repeats, so the
 * contexts are pooled per process: a computer borrows one for its lifetime and hands it back in its destructor
 * (contexts that joined a row-sharded communicator are not shared). */
class DevicePool
{
public:
    static std::shared_ptr<Device> acquire(int a_device)
    {
        DevicePool& p = instance();
        std::unique_ptr<Device> d;
        {
            std::lock_guard<std::mutex> lock(p.m_mutex);
            for (size_t i = 0; i < p.m_free.size(); ++i)
                if (p.m_free[i].first == a_device) { d = std::move(p.m_free[i].second); p.m_free.erase(p.m_free.begin() + i); break; }
        }
        if (!d) d.reset(new Device(a_device));
        rslf_cuda_set_confidence_criterion(d->get(), 0);
        rslf_cuda_set_fast_math(d->get(), 0);
        Device* raw = d.release();
        return std::shared_ptr<Device>(raw, [a_device](Device* q) {
            if (q->sharded) { delete q; return; }
            DevicePool& pool = instance();
            std::lock_guard<std::mutex> lock(pool.m_mutex);
            if (pool.m_free.size() < 4) pool.m_free.emplace_back(a_device, std::unique_ptr<Device>(q));
            else delete q;
        });
    }
private:
    static DevicePool& instance() { static DevicePool s_pool; return s_pool; }
    std::mutex m_mutex;
    std::vector<std::pair<int, std::unique_ptr<Device>>> m_free;
};
template <typename T> inline std::vector<T*> data_ptrs(Vec<Mat>& a_mats) { std::vector<T*> p(a_mats.size()); for (size_t i = 0; i < a_mats.size(); ++i) p[i] = a_mats[i].template ptr<T>(); return p; }
/* dense [n][rows][row_elems] staging <- n Mats (the small free-function wrappers) */
template <typename T>
inline std::vector<T> gather(const Vec<Mat>& a_mats, int rows, size_t a_row_elems)
{
    std::vector<T> buf(a_mats.size() * (size_t)rows * a_row_elems);
    for (size_t s = 0; s < a_mats.size(); ++s)
        for (int v = 0; v < rows; ++v)
            std::memcpy(buf.data() + ((size_t)s * rows + v) * a_row_elems, a_mats[s].template ptr<T>(v), a_row_elems * sizeof(T));
    return buf;
}
/* dense [S][V][U] staging <-> S Mats */
template <typename T>
inline void scatter(const std::vector<T>& a_buf, Vec<Mat>& a_mats, int V, size_t a_row_elems)
{
    for (size_t s = 0; s < a_mats.size(); ++s)
        for (int v = 0; v < V; ++v)
            std::memcpy(a_mats[s].template ptr<T>(v), a_buf.data() + ((size_t)s * V + v) * a_row_elems, a_row_elems * sizeof(T));
}
} // namespace detail

/* ---------------------------------------------------------------- Depth1DComputer */
/* rslf::Depth1DComputer<DataType> (dc.hpp:26-68, ctor :254-317, run :319-363): ONE EPI (S x U x C), edge confidence and
 * depth of line s_hat, no selective median.  On the device it runs as a one-row pile (the per-line results coincide
 * except for the median, so the unfiltered depth is read back); the opening of the pile wrapper (core.hpp:759) does
 * not apply because the reference calls compute_1D_edge_confidence directly (dc.hpp:339). */
template <typename DataType>
class Depth1DComputer
{
public:
    Depth1DComputer(const Mat& epi, float dmin, float dmax, int dim_d, int s_hat = -1, float epi_scale_factor = -1,
                    const Depth1DParameters<DataType>& parameters = Depth1DParameters<DataType>::get_default(), int device = 0)
        : m_parameters(parameters), m_device(detail::DevicePool::acquire(device)), m_dim_d(dim_d), m_dmin(dmin), m_dmax(dmax)
    {
        int one = 0;
        m_device->upload(Vec<Mat>(1, epi), epi_scale_factor, channels_of<DataType>::value, one, m_dim_s, m_dim_u);
        m_s_hat = (s_hat < 0 || s_hat > m_dim_s - 1) ? (int)std::floor((0.0 + m_dim_s) / 2) : s_hat;      /* dc.hpp:296-305 */
    }
    void run()
    {
        const int C = channels_of<DataType>::value;
        rslf_params p = m_parameters.to_abi();
        p.edge_confidence_opening_size = 1;
        m_best_depth_u = Mat::zeros(1, m_dim_u, CV_32FC1);
        m_edge_confidence_u = Mat::zeros(1, m_dim_u, CV_32FC1);
        m_edge_confidence_mask_u = Mat::zeros(1, m_dim_u, CV_8UC1);
        m_disp_confidence_u = Mat::zeros(1, m_dim_u, CV_32FC1);
        m_rbar_u = Mat::zeros(1, m_dim_u, CV_MAKETYPE(CV_32F, C));
        m_device->check(rslf_cuda_depth1d_pile(m_device->get(), m_dmin, m_dmax, m_dim_d, m_s_hat, &p, nullptr,
                                               m_edge_confidence_u.template ptr<float>(),
                                               m_edge_confidence_mask_u.template ptr<unsigned char>(),
                                               m_disp_confidence_u.template ptr<float>(), m_rbar_u.template ptr<float>()),
                        "rslf_cuda_depth1d_pile");
        m_device->check(rslf_cuda_depth1d_pile_get_raw_depth(m_device->get(), m_best_depth_u.template ptr<float>()),
                        "rslf_cuda_depth1d_pile_get_raw_depth");
    }
    int get_s_hat() const { return m_s_hat; }

    /* private in the reference (dc.hpp:53-60); exposed here because they are the results */
    Mat m_edge_confidence_u, m_edge_confidence_mask_u, m_disp_confidence_u, m_rbar_u, m_best_depth_u;

private:
    const Depth1DParameters<DataType>& m_parameters;
    std::shared_ptr<detail::Device> m_device;
    int m_dim_d, m_dim_s = 0, m_dim_u = 0, m_s_hat = 0;
    float m_dmin, m_dmax;
};
using Depth1DComputer_1ch = Depth1DComputer<float>;
using Depth1DComputer_3ch = Depth1DComputer<Vec3f>;

/* ---------------------------------------------------------------- Depth1DComputer_pile */
template <typename DataType>
class Depth1DComputer_pile
{
public:
    Depth1DComputer_pile(const Vec<Mat>& epis, float dmin, float dmax, int dim_d, int s_hat = -1,
                         float epi_scale_factor = -1,
                         const Depth1DParameters<DataType>& parameters = Depth1DParameters<DataType>::get_default(),
                         int device = 0)
        : m_parameters(parameters), m_device(detail::DevicePool::acquire(device)), m_dim_d(dim_d), m_dmin(dmin), m_dmax(dmax)
    {
        m_device->upload(epis, epi_scale_factor, channels_of<DataType>::value, m_dim_v, m_dim_s, m_dim_u);
        /* s_hat defaults to floor(S/2) (dc.hpp:489-498) */
        m_s_hat = (s_hat < 0 || s_hat > m_dim_s - 1) ? (int)std::floor((0.0 + m_dim_s) / 2) : s_hat;
    }
    void run()
    {
        const int C = channels_of<DataType>::value;
        rslf_params p = m_parameters.to_abi();
        m_best_depth_v_u = Mat::zeros(m_dim_v, m_dim_u, CV_32FC1);
        m_edge_confidence_v_u = Mat::zeros(m_dim_v, m_dim_u, CV_32FC1);
        m_edge_confidence_mask_v_u = Mat::zeros(m_dim_v, m_dim_u, CV_8UC1);
        m_disp_confidence_v_u = Mat::zeros(m_dim_v, m_dim_u, CV_32FC1);
        m_rbar_v_u = Mat::zeros(m_dim_v, m_dim_u, CV_MAKETYPE(CV_32F, C));
        m_device->check(rslf_cuda_depth1d_pile(m_device->get(), m_dmin, m_dmax, m_dim_d, m_s_hat, &p,
                                               m_best_depth_v_u.template ptr<float>(), m_edge_confidence_v_u.template ptr<float>(),
                                               m_edge_confidence_mask_v_u.template ptr<unsigned char>(),
                                               m_disp_confidence_v_u.template ptr<float>(), m_rbar_v_u.template ptr<float>()),
                        "rslf_cuda_depth1d_pile");
    }
    int get_s_hat() { return m_s_hat; }
    rslf_timing get_timing() const { rslf_timing t; rslf_cuda_last_timing(m_device->get(), &t); return t; }

    /* private in the reference (dc.hpp:124-140); exposed here because they are the results */
    Mat m_edge_confidence_v_u, m_edge_confidence_mask_v_u, m_disp_confidence_v_u, m_rbar_v_u, m_best_depth_v_u;

private:
    const Depth1DParameters<DataType>& m_parameters;     /* held by reference, as in the reference (dc.hpp:142) */
    std::shared_ptr<detail::Device> m_device;
    int m_dim_d, m_dim_v = 0, m_dim_s = 0, m_dim_u = 0, m_s_hat = 0;
    float m_dmin, m_dmax;
};
using Depth1DComputer_pile_1ch = Depth1DComputer_pile<float>;
using Depth1DComputer_pile_3ch = Depth1DComputer_pile<Vec3f>;

/* ---------------------------------------------------------------- Depth2DComputer */
template <typename DataType>
class Depth2DComputer
{
public:
    Depth2DComputer(const Vec<Mat>& epis, float dmin, float dmax, int dim_d, float epi_scale_factor = -1,
                    const Depth1DParameters<DataType>& parameters = Depth1DParameters<DataType>::get_default(),
                    bool verbose = true, int device = 0)
        : m_parameters(parameters), m_device(detail::DevicePool::acquire(device)), m_dim_d(dim_d), m_dmin(dmin), m_dmax(dmax),
          m_verbose(verbose)
    {
        const rslf_params p = m_parameters.to_abi();
        m_device->upload(epis, epi_scale_factor, channels_of<DataType>::value, m_dim_v, m_dim_s, m_dim_u, &p);
    }
    void run()
    {
        const int C = channels_of<DataType>::value;
        const int S = m_dim_s, V = m_dim_v, U = m_dim_u;
        rslf_params p = m_parameters.to_abi();
        const size_t px = (size_t)S * V * U;
        std::vector<float> lo, hi;                 /* user-edited bounds (edit_dmin / edit_dmax) travel as dense maps */
        if (!m_dmin_s_v_u.empty()) {
            lo.resize(px); hi.resize(px);
            for (int s = 0; s < S; ++s)
                for (int v = 0; v < V; ++v) {
                    std::memcpy(&lo[((size_t)s * V + v) * U], m_dmin_s_v_u[s].template ptr<float>(v), U * sizeof(float));
                    std::memcpy(&hi[((size_t)s * V + v) * U], m_dmax_s_v_u[s].template ptr<float>(v), U * sizeof(float));
                }
        }
        m_device->check(rslf_cuda_depth2d_run(m_device->get(), m_dmin, m_dmax, m_dim_d, &p, lo.empty() ? nullptr : lo.data(),
                                              hi.empty() ? nullptr : hi.data()), "rslf_cuda_depth2d_run");
        m_best_depth_s_v_u = detail::planes(S, V, U, CV_32FC1);
        m_edge_confidence_s_v_u = detail::planes(S, V, U, CV_32FC1);
        m_disp_confidence_s_v_u = detail::planes(S, V, U, CV_32FC1);
        m_edge_confidence_mask_s_v_u = detail::planes(S, V, U, CV_8UC1);
        m_rbar_s_v_u = detail::planes(S, V, U, CV_MAKETYPE(CV_32F, C));
        /* the library writes every plane straight into its Mat (no staging copy on this side) */
        auto pd = detail::data_ptrs<float>(m_best_depth_s_v_u), pe = detail::data_ptrs<float>(m_edge_confidence_s_v_u),
             pc = detail::data_ptrs<float>(m_disp_confidence_s_v_u), pr = detail::data_ptrs<float>(m_rbar_s_v_u);
        auto pm = detail::data_ptrs<unsigned char>(m_edge_confidence_mask_s_v_u);
        m_device->check(rslf_cuda_depth2d_get_mats(m_device->get(), pd.data(), pe.data(), pm.data(), pc.data(), pr.data(),
                                                   (size_t)m_best_depth_s_v_u[0].step, (size_t)m_edge_confidence_mask_s_v_u[0].step,
                                                   (size_t)m_rbar_s_v_u[0].step), "rslf_cuda_depth2d_get_mats");
    }
    const Vec<Mat>& get_depths_s_v_u() { return m_best_depth_s_v_u; }
    /* the normalised float32 EPIs the computer works on (dc.hpp:207) */
    const Vec<Mat>& get_epis()
    {
        if (m_epis.empty()) {
            const int C = channels_of<DataType>::value;
            m_epis.resize(m_dim_v);
            for (auto& m : m_epis) m = Mat::zeros(m_dim_s, m_dim_u, CV_MAKETYPE(CV_32F, C));
            auto pe = detail::data_ptrs<float>(m_epis);
            m_device->check(rslf_cuda_get_epis(m_device->get(), pe.data(), (size_t)m_epis[0].step), "rslf_cuda_get_epis");
        }
        return m_epis;
    }
    /* row-sharded run (one process per GPU): see detail::Device::shard; call before run() */
    void set_row_shards(const std::vector<int>& a_row_starts, int a_rank, const void* a_nccl_id) { m_device->shard(a_row_starts, a_rank, a_nccl_id); }
    /* C_e > threshold, or everything when accept_all is set (dc.hpp:893-915) */
    Vec<Mat> get_valid_depths_mask_s_v_u()
    {
        rslf_params p = m_parameters.to_abi();
        std::vector<unsigned char> m((size_t)m_dim_s * m_dim_v * m_dim_u);
        m_device->check(rslf_cuda_depth2d_get_valid_mask(m_device->get(), m_accept_all ? 1 : 0, &p, m.data()),
                        "rslf_cuda_depth2d_get_valid_mask");
        Vec<Mat> out = detail::planes(m_dim_s, m_dim_v, m_dim_u, CV_8UC1);
        detail::scatter(m, out, m_dim_v, (size_t)m_dim_u);
        return out;
    }
    /* per-pixel bounds (dc.hpp:211-213): allocated on first use, filled with the global bounds */
    Vec<Mat>& edit_dmin() { ensure_bounds(); return m_dmin_s_v_u; }
    Vec<Mat>& edit_dmax() { ensure_bounds(); return m_dmax_s_v_u; }
    void set_accept_all(bool b) { m_accept_all = b; }
    rslf_timing get_timing() const { rslf_timing t; rslf_cuda_last_timing(m_device->get(), &t); return t; }

    Vec<Mat> m_edge_confidence_s_v_u, m_edge_confidence_mask_s_v_u, m_disp_confidence_s_v_u, m_rbar_s_v_u, m_best_depth_s_v_u;

private:
    void ensure_bounds()
    {
        if (!m_dmin_s_v_u.empty()) return;
        m_dmin_s_v_u = detail::planes(m_dim_s, m_dim_v, m_dim_u, CV_32FC1);
        m_dmax_s_v_u = detail::planes(m_dim_s, m_dim_v, m_dim_u, CV_32FC1);
        for (int s = 0; s < m_dim_s; ++s)
            for (int v = 0; v < m_dim_v; ++v)
                for (int u = 0; u < m_dim_u; ++u) {
                    m_dmin_s_v_u[s].template at<float>(v, u) = m_dmin;
                    m_dmax_s_v_u[s].template at<float>(v, u) = m_dmax;
                }
    }
    const Depth1DParameters<DataType>& m_parameters;
    std::shared_ptr<detail::Device> m_device;
    int m_dim_d, m_dim_v = 0, m_dim_s = 0, m_dim_u = 0;
    float m_dmin, m_dmax;
    Vec<Mat> m_dmin_s_v_u, m_dmax_s_v_u, m_epis;
    bool m_accept_all = false;
    bool m_verbose;
};
using Depth2DComputer_1ch = Depth2DComputer<float>;
using Depth2DComputer_3ch = Depth2DComputer<Vec3f>;

/* ---------------------------------------------------------------- FineToCoarse */
template <typename DataType>
class FineToCoarse
{
public:
    FineToCoarse(const Vec<Mat>& a_epis, float a_d_min, float a_d_max, int a_dim_d, float a_epi_scale_factor = -1,
                 const Depth1DParameters<DataType>& a_parameters = Depth1DParameters<DataType>::get_default(),
                 int a_max_pyr_depth = -1, bool a_accept_all_last_scale = true, int a_device = 0)
        : m_parameters(a_parameters), m_device(detail::DevicePool::acquire(a_device)), m_dim_d(a_dim_d), m_dmin(a_d_min),
          m_dmax(a_d_max), m_max_pyr_depth(a_max_pyr_depth), m_accept_all_last_scale(a_accept_all_last_scale)
    {
        const rslf_params p = m_parameters.to_abi();
        m_device->upload(a_epis, a_epi_scale_factor, channels_of<DataType>::value, m_dim_v, m_dim_s, m_dim_u, &p);
    }
    void run()
    {
        rslf_params p = m_parameters.to_abi();
        m_device->check(rslf_cuda_fine_to_coarse_run(m_device->get(), m_dmin, m_dmax, m_dim_d, &p, m_max_pyr_depth,
                                                     m_accept_all_last_scale ? 1 : 0), "rslf_cuda_fine_to_coarse_run");
    }
    /* fused disparity maps and validity masks, one per view (ftc.hpp:301-322) */
    void get_results(Vec<Mat>& a_out_map_s_v_u, Vec<Mat>& a_out_validity_s_v_u)
    {
        a_out_map_s_v_u = detail::planes(m_dim_s, m_dim_v, m_dim_u, CV_32FC1);
        a_out_validity_s_v_u = detail::planes(m_dim_s, m_dim_v, m_dim_u, CV_8UC1);
        auto pm = detail::data_ptrs<float>(a_out_map_s_v_u);
        auto pv = detail::data_ptrs<unsigned char>(a_out_validity_s_v_u);
        m_device->check(rslf_cuda_fine_to_coarse_get_mats(m_device->get(), pm.data(), (size_t)a_out_map_s_v_u[0].step, pv.data(),
                                                          (size_t)a_out_validity_s_v_u[0].step), "rslf_cuda_fine_to_coarse_get_mats");
    }
    /* row-sharded run (one process per GPU): see detail::Device::shard; call before run() */
    void set_row_shards(const std::vector<int>& a_row_starts, int a_rank, const void* a_nccl_id) { m_device->shard(a_row_starts, a_rank, a_nccl_id); }
    /* 0: the reference as written (edge confidence gates propagation / validity), 1: -D_USE_DISP_CONFIDENCE_SCORE as intended */
    void set_confidence_criterion(int a_criterion) { m_device->check(rslf_cuda_set_confidence_criterion(m_device->get(), a_criterion), "rslf_cuda_set_confidence_criterion"); }
    /* coloured disparity maps, one CV_8UC3 (BGR) image per view (ftc.hpp:324-377).  With OpenCV the colour table is
     * the one cv::applyColorMap(…, a_cv_colormap) uses; without it the caller supplies the 256 x 3 table. */
    void get_coloured_depth_maps(Vec<Mat>& a_out_plot_depth_s_v_u, int a_cv_colormap = 2 /* cv::COLORMAP_JET */,
                                 bool a_saturate = true, const unsigned char* a_lut_bgr_256x3 = nullptr)
    {
        unsigned char lut[768];
        if (a_lut_bgr_256x3) std::memcpy(lut, a_lut_bgr_256x3, sizeof(lut));
        else {
#ifdef RSLF_B200_HAVE_OPENCV_IMGPROC
            cv::Mat ramp(256, 1, CV_8UC1), table;
            for (int i = 0; i < 256; ++i) ramp.at<unsigned char>(i) = (unsigned char)i;
            cv::applyColorMap(ramp, table, a_cv_colormap);
            std::memcpy(lut, table.data, sizeof(lut));
#else
            (void)a_cv_colormap;
            throw Error(RSLF_ERR_ARG, "get_coloured_depth_maps: built without OpenCV's imgproc, pass the colour table");
#endif
        }
        const size_t px = (size_t)m_dim_s * m_dim_v * m_dim_u;
        std::vector<unsigned char> bgr(px * 3);
        const rslf_params p = m_parameters.to_abi();
        m_device->check(rslf_cuda_fine_to_coarse_get_coloured(m_device->get(), lut, a_saturate ? 1 : 0, &p, bgr.data(), nullptr),
                        "rslf_cuda_fine_to_coarse_get_coloured");
        a_out_plot_depth_s_v_u = detail::planes(m_dim_s, m_dim_v, m_dim_u, CV_8UC3);
        detail::scatter(bgr, a_out_plot_depth_s_v_u, m_dim_v, (size_t)m_dim_u * 3);
    }
    rslf_timing get_timing() const { rslf_timing t; rslf_cuda_last_timing(m_device->get(), &t); return t; }

private:
    const Depth1DParameters<DataType>& m_parameters;
    std::shared_ptr<detail::Device> m_device;
    int m_dim_d, m_dim_v = 0, m_dim_s = 0, m_dim_u = 0;
    float m_dmin, m_dmax;
    int m_max_pyr_depth;
    bool m_accept_all_last_scale;
};
using FineToCoarse_1ch = FineToCoarse<float>;
using FineToCoarse_3ch = FineToCoarse<Vec3f>;

/* ---------------------------------------------------------------- free functions
 * The reference's L2 entry points (rslf_depth_computation_core.hpp:236-375, rslf_fine_to_coarse_core.hpp:28-49) with
 * their Mat signatures.  BufferDepth1D (core.hpp:154-230) is per-thread scratch of the CPU path and has no content
 * here; the type exists so that call sites compile unchanged.  Every call runs on a process-wide device context. */
template <typename DataType> struct BufferDepth1D {};

/* 128-byte NCCL id for row-sharded runs: made by one rank, sent to the others by the caller's own means */
inline std::vector<unsigned char> nccl_unique_id()
{
    std::vector<unsigned char> id(128);
    const int rc = rslf_cuda_nccl_unique_id(id.data());
    if (rc != RSLF_OK) throw Error(rc, std::string("rslf_cuda_nccl_unique_id: ") + rslf_cuda_strerror(rc));
    return id;
}

namespace detail
{
inline Device& shared_device() { static Device s_dev(0); return s_dev; }
}

/* compute_1D_edge_confidence_pile (core.hpp:728-770): C_e and its mask on line a_s of every EPI; EPIs normalised float32 */
template <typename DataType>
void compute_1D_edge_confidence_pile(const Vec<Mat>& a_epis, int a_s, Mat& a_edge_confidence_v_u, Mat& a_edge_confidence_mask_v_u,
                                     const Depth1DParameters<DataType>& a_parameters, Vec<BufferDepth1D<DataType>*>* = nullptr)
{
    detail::Device& d = detail::shared_device();
    int V, S, U;
    d.upload(a_epis, 1.f, channels_of<DataType>::value, V, S, U);
    const rslf_params p = a_parameters.to_abi();
    a_edge_confidence_v_u = Mat::zeros(V, U, CV_32FC1);
    a_edge_confidence_mask_v_u = Mat::zeros(V, U, CV_8UC1);
    d.check(rslf_cuda_edge_confidence(d.get(), a_s, &p, a_edge_confidence_v_u.template ptr<float>(),
                                      a_edge_confidence_mask_v_u.template ptr<unsigned char>()), "rslf_cuda_edge_confidence");
}
/* compute_1D_edge_confidence (core.hpp:426-478): one EPI.  The opening belongs to the pile wrapper (core.hpp:759). */
template <typename DataType>
void compute_1D_edge_confidence(const Mat& a_epi, int a_s, Mat& a_edge_confidence_u, Mat& a_edge_confidence_mask_u,
                                const Depth1DParameters<DataType>& a_parameters, BufferDepth1D<DataType>* = nullptr)
{
    Depth1DParameters<DataType> q = a_parameters;
    q.par_edge_confidence_opening_size = 1;
    compute_1D_edge_confidence_pile<DataType>(Vec<Mat>(1, a_epi), a_s, a_edge_confidence_u, a_edge_confidence_mask_u, q);
}
/* compute_2D_edge_confidence (core.hpp:901-931): every line of every EPI */
template <typename DataType>
void compute_2D_edge_confidence(const Vec<Mat>& a_epis, Vec<Mat>& a_edge_confidence_s_v_u, Vec<Mat>& a_edge_confidence_mask_s_v_u,
                                const Depth1DParameters<DataType>& a_parameters, Vec<BufferDepth1D<DataType>*>* = nullptr)
{
    const int S = a_epis.empty() ? 0 : a_epis[0].rows;
    a_edge_confidence_s_v_u.resize(S); a_edge_confidence_mask_s_v_u.resize(S);
    for (int s = 0; s < S; ++s)
        compute_1D_edge_confidence_pile<DataType>(a_epis, s, a_edge_confidence_s_v_u[s], a_edge_confidence_mask_s_v_u[s], a_parameters);
}
/* selective_median_filter (core.hpp:663-718) */
template <typename DataType>
void selective_median_filter(const Mat& a_src, Mat& a_dst, const Vec<Mat>& a_epis, int a_s_hat, int a_size, const Mat& a_mask_v_u, float a_epsilon)
{
    detail::Device& d = detail::shared_device();
    int V, S, U;
    d.upload(a_epis, 1.f, channels_of<DataType>::value, V, S, U);
    const std::vector<float> src = detail::gather<float>(Vec<Mat>(1, a_src), V, (size_t)U);
    const std::vector<unsigned char> mask = detail::gather<unsigned char>(Vec<Mat>(1, a_mask_v_u), V, (size_t)U);
    a_dst = Mat::zeros(V, U, CV_32FC1);
    d.check(rslf_cuda_selective_median(d.get(), src.data(), mask.data(), a_s_hat, a_size, a_epsilon, a_dst.template ptr<float>()),
            "rslf_cuda_selective_median");
}
/* compute_2D_depth_epi (core.hpp:933-1133): all lines of all EPIs with per-pixel bounds.  The edge confidence is
 * recomputed on the device from the same EPIs and parameters (the reference takes it as an input that its caller
 * made with compute_2D_edge_confidence, dc.hpp:771-781); the line-confidence maps are left untouched (flag off). */
template <typename DataType>
void compute_2D_depth_epi(const Vec<Mat>& a_epis, const Vec<Mat>& a_dmin_s_v_u, const Vec<Mat>& a_dmax_s_v_u, int a_dim_d,
                          Vec<Mat>& a_edge_confidence_s_v_u, Vec<Mat>& a_edge_confidence_mask_s_v_u, Vec<Mat>& a_disp_confidence_s_v_u,
                          Vec<Mat>& /* a_line_confidence_s_v_u */, Vec<Mat>& a_best_depth_s_v_u, Vec<Mat>& a_rbar_s_v_u,
                          const Depth1DParameters<DataType>& a_parameters, Vec<BufferDepth1D<DataType>*>* = nullptr, bool = true)
{
    detail::Device& d = detail::shared_device();
    const int C = channels_of<DataType>::value;
    int V, S, U;
    d.upload(a_epis, 1.f, C, V, S, U);
    const rslf_params p = a_parameters.to_abi();
    const std::vector<float> lo = detail::gather<float>(a_dmin_s_v_u, V, (size_t)U), hi = detail::gather<float>(a_dmax_s_v_u, V, (size_t)U);
    d.check(rslf_cuda_depth2d_run(d.get(), 0.f, 0.f, a_dim_d, &p, lo.data(), hi.data()), "rslf_cuda_depth2d_run");
    a_best_depth_s_v_u = detail::planes(S, V, U, CV_32FC1);
    a_edge_confidence_s_v_u = detail::planes(S, V, U, CV_32FC1);
    a_disp_confidence_s_v_u = detail::planes(S, V, U, CV_32FC1);
    a_edge_confidence_mask_s_v_u = detail::planes(S, V, U, CV_8UC1);
    a_rbar_s_v_u = detail::planes(S, V, U, CV_MAKETYPE(CV_32F, C));
    auto pd = detail::data_ptrs<float>(a_best_depth_s_v_u), pe = detail::data_ptrs<float>(a_edge_confidence_s_v_u),
         pc = detail::data_ptrs<float>(a_disp_confidence_s_v_u), pr = detail::data_ptrs<float>(a_rbar_s_v_u);
    auto pm = detail::data_ptrs<unsigned char>(a_edge_confidence_mask_s_v_u);
    d.check(rslf_cuda_depth2d_get_mats(d.get(), pd.data(), pe.data(), pm.data(), pc.data(), pr.data(), (size_t)a_best_depth_s_v_u[0].step,
                                       (size_t)a_edge_confidence_mask_s_v_u[0].step, (size_t)a_rbar_s_v_u[0].step), "rslf_cuda_depth2d_get_mats");
}
/* downsample_EPIs (rslf_fine_to_coarse_core.cpp:14-60): CV_32F, CV_8U and CV_16U stacks keep their depth */
inline void downsample_EPIs(const Vec<Mat>& in_epis, Vec<Mat>& out_epis)
{
    detail::Device& d = detail::shared_device();
    if (in_epis.empty() || in_epis[0].empty()) throw Error(RSLF_ERR_ARG, "empty EPI vector");
    const int V = (int)in_epis.size(), S = in_epis[0].rows, U = in_epis[0].cols, C = in_epis[0].channels(), depth = in_epis[0].depth();
    int V2 = 0, U2 = 0;
    auto run = [&](auto tag) {
        using T = decltype(tag);
        const std::vector<T> in = detail::gather<T>(in_epis, S, (size_t)U * C);
        std::vector<T> out(((size_t)V / 2 + 1) * S * ((size_t)U / 2 + 1) * C);
        int rc;
        if (sizeof(T) == 4) rc = rslf_cuda_downsample_epis(d.get(), (const float*)in.data(), V, S, U, C, (float*)out.data(), &V2, &U2);
        else if (sizeof(T) == 1) rc = rslf_cuda_downsample_epis_u8(d.get(), (const unsigned char*)in.data(), V, S, U, C, (unsigned char*)out.data(), &V2, &U2);
        else rc = rslf_cuda_downsample_epis_u16(d.get(), (const unsigned short*)in.data(), V, S, U, C, (unsigned short*)out.data(), &V2, &U2);
        d.check(rc, "rslf_cuda_downsample_epis");
        out_epis.assign(V2, Mat());
        for (int v = 0; v < V2; ++v) {
            out_epis[v] = Mat::zeros(S, U2, CV_MAKETYPE(depth, C));
            for (int s = 0; s < S; ++s)
                std::memcpy(out_epis[v].template ptr<T>(s), out.data() + (((size_t)v * S + s) * U2) * C, (size_t)U2 * C * sizeof(T));
        }
    };
    if (depth == CV_32F) run(float());
    else if (depth == CV_8U) run((unsigned char)0);
    else if (depth == CV_16U) run((unsigned short)0);
    else throw Error(RSLF_ERR_UNSUPPORTED, "downsample_EPIs: CV_8U, CV_16U or CV_32F");
}
/* fuse_disp_maps (rslf_fine_to_coarse_core.cpp:69-135): pyramids finest first */
inline void fuse_disp_maps(const Vec<Vec<Mat>>& in_disp_pyr_p_s_v_u, const Vec<Vec<Mat>>& in_validity_indicators_p_s_v_u,
                           Vec<Mat>& out_map_s_v_u, Vec<Mat>& out_validity_s_v_u)
{
    detail::Device& d = detail::shared_device();
    const int L = (int)in_disp_pyr_p_s_v_u.size();
    if (L < 1 || (int)in_validity_indicators_p_s_v_u.size() != L || in_disp_pyr_p_s_v_u[0].empty()) throw Error(RSLF_ERR_ARG, "empty pyramid");
    const int S = (int)in_disp_pyr_p_s_v_u[0].size();
    std::vector<int> Vp(L), Up(L);
    std::vector<std::vector<float>> dd(L); std::vector<std::vector<unsigned char>> vv(L);
    std::vector<const float*> dp(L); std::vector<const unsigned char*> vp(L);
    for (int p = 0; p < L; ++p) {
        Vp[p] = in_disp_pyr_p_s_v_u[p][0].rows; Up[p] = in_disp_pyr_p_s_v_u[p][0].cols;
        dd[p] = detail::gather<float>(in_disp_pyr_p_s_v_u[p], Vp[p], (size_t)Up[p]);
        vv[p] = detail::gather<unsigned char>(in_validity_indicators_p_s_v_u[p], Vp[p], (size_t)Up[p]);
        dp[p] = dd[p].data(); vp[p] = vv[p].data();
    }
    const size_t px = (size_t)S * Vp[0] * Up[0];
    std::vector<float> om(px); std::vector<unsigned char> ov(px);
    d.check(rslf_cuda_fuse_disp_maps(d.get(), L, S, Vp.data(), Up.data(), dp.data(), vp.data(), om.data(), ov.data()), "rslf_cuda_fuse_disp_maps");
    out_map_s_v_u = detail::planes(S, Vp[0], Up[0], CV_32FC1);
    out_validity_s_v_u = detail::planes(S, Vp[0], Up[0], CV_8UC1);
    detail::scatter(om, out_map_s_v_u, Vp[0], (size_t)Up[0]);
    detail::scatter(ov, out_validity_s_v_u, Vp[0], (size_t)Up[0]);
}

} // namespace rslf_b200

#endif /* RSLF_B200_HPP */
