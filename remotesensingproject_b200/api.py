"""Host-side mirror of the reference's depth-computer / fine-to-coarse classes.

Same class names, constructor arguments and result members as
RSLightFields/include/rslf_depth_computation.hpp (Depth1DComputer_pile :93-143,
Depth2DComputer :166-229) and rslf_fine_to_coarse.hpp (FineToCoarse :26-81),
with numpy arrays standing in for cv::Mat (identical memory layout: row-major,
interleaved channels).  Everything is computed by the CUDA library
(librslf_b200.so) through the C ABI of include/rslf_b200.h; there is no CPU
fallback — without the library or without a CUDA device the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RSLF_B200_LIB: developer hook to time an alternative build of the same library (A/B experiments)
LIB_PATH = os.environ.get("RSLF_B200_LIB") or os.path.join(_HERE, "librslf_b200.so")

RSLF_DEPTH_8U = 0
RSLF_DEPTH_16U = 2
RSLF_DEPTH_32F = 5
_DEPTH_OF = {np.dtype(np.uint8): RSLF_DEPTH_8U, np.dtype(np.uint16): RSLF_DEPTH_16U, np.dtype(np.float32): RSLF_DEPTH_32F}


class RslfError(RuntimeError):
    pass


class Params(C.Structure):
    """rslf_params == rslf::Depth1DParameters<T> (rslf_depth_computation_core.hpp:66-142)."""
    _fields_ = [
        ("edge_score_threshold", C.c_float), ("line_score_threshold", C.c_float),
        ("disp_score_threshold", C.c_float), ("raw_score_threshold", C.c_float),
        ("mean_shift_max_iter", C.c_int), ("edge_confidence_filter_size", C.c_int),
        ("edge_confidence_opening_type", C.c_int), ("edge_confidence_opening_size", C.c_int),
        ("median_filter_size", C.c_int), ("median_filter_epsilon", C.c_float),
        ("propagation_epsilon", C.c_float), ("slope_factor", C.c_float),
        ("cut_shadows", C.c_int), ("shadow_level", C.c_float), ("kernel_h", C.c_float),
    ]


class Decision(C.Structure):
    """rslf_decision (include/rslf_b200.h)."""
    _fields_ = [("level", C.c_int), ("s_hat", C.c_int), ("pix", C.c_int), ("index", C.c_int), ("score", C.c_float),
                ("rbar", C.c_float * 3), ("disp_conf", C.c_float), ("disparity", C.c_float), ("accepted", C.c_int),
                ("reserved", C.c_int)]


class Timing(C.Structure):
    """rslf_timing (include/rslf_b200.h)."""
    _fields_ = [
        ("ms_total", C.c_float), ("ms_edge", C.c_float), ("ms_depth", C.c_float), ("ms_reduce", C.c_float),
        ("ms_median", C.c_float), ("ms_propagate", C.c_float), ("ms_pyramid", C.c_float),
        ("ms_h2d", C.c_float), ("ms_d2h", C.c_float),
        ("computed_pixels", C.c_double), ("samples", C.c_double), ("depth_launches", C.c_double),
        ("kernel_launches", C.c_double), ("levels", C.c_int), ("passes", C.c_int),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    """Loads librslf_b200.so (built by __graft_entry__.build() / csrc/Makefile). Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RslfError("CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.rslf_cuda_strerror.restype = C.c_char_p
        _lib.rslf_cuda_last_error_text.restype = C.c_char_p
        _lib.rslf_cuda_last_error_text.argtypes = [C.c_void_p]
    return _lib


def default_params(**kw):
    p = Params()
    lib().rslf_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _fp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _bp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_uint8))


class Context:
    """One rslf_ctx: a CUDA device, a stream and the device buffers of one light field."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().rslf_cuda_create(int(device), C.byref(self._h))
        if rc != 0:
            raise RslfError("rslf_cuda_create(device=%d) failed: %s (no CUDA device? there is no CPU fallback)"
                            % (device, lib().rslf_cuda_strerror(rc).decode()))
        self.device = device
        self.dims = None
        self._keep = None

    def close(self):
        if self._h:
            lib().rslf_cuda_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc, what):
        if rc != 0:
            raise RslfError("%s failed: %s: %s" % (what, lib().rslf_cuda_strerror(rc).decode(),
                                                   lib().rslf_cuda_last_error_text(self._h).decode()))

    # ---- input ----------------------------------------------------------------
    def upload_epis(self, epis, epi_scale_factor=-1.0, params=None):
        """epis: the reference's Vec<Mat> — a [V][S][U][C] (or [V][S][U]) array or a list of V
        [S][U][C] arrays; uint8, uint16 or float32 host memory (numpy, or a pinned CPU torch tensor).
        params given: pipelined ingest (rslf_cuda_upload_epis_pipelined) — the stack is normalised and its level-0 edge
        confidence computed chunk by chunk while the rest is still being uploaded."""
        mats = _as_mats(epis)
        V = len(mats)
        S, U, Cc = mats[0].shape
        depth = _DEPTH_OF[mats[0].dtype]
        ptrs = (C.c_void_p * V)(*[m.ctypes.data for m in mats])
        step = mats[0].strides[0]
        if params is not None:
            self.check(lib().rslf_cuda_upload_epis_pipelined(self._h, ptrs, V, S, U, Cc, depth, C.c_size_t(step),
                                                             C.c_float(epi_scale_factor), C.byref(params)), "rslf_cuda_upload_epis_pipelined")
        else:
            self.check(lib().rslf_cuda_upload_epis(self._h, ptrs, V, S, U, Cc, depth, C.c_size_t(step),
                                                   C.c_float(epi_scale_factor)), "rslf_cuda_upload_epis")
        self.dims = (V, S, U, Cc)
        self._keep = None

    def upload_images(self, imgs, epi_scale_factor=-1.0):
        """imgs: S images [V][U][C] (rslf::build_epis_from_imgs input, rslf_io.cpp:194-227)."""
        mats = _as_mats(imgs)
        S = len(mats)
        V, U, Cc = mats[0].shape
        depth = _DEPTH_OF[mats[0].dtype]
        ptrs = (C.c_void_p * S)(*[m.ctypes.data for m in mats])
        self.check(lib().rslf_cuda_upload_images(self._h, ptrs, S, V, U, Cc, depth, C.c_size_t(mats[0].strides[0]),
                                                 C.c_float(epi_scale_factor)), "rslf_cuda_upload_images")
        self.dims = (V, S, U, Cc)
        self._keep = None

    def set_epis_device(self, t, epi_scale_factor=-1.0):
        """t: a CUDA torch tensor [V][S][U][C], float32, uint16 or uint8, contiguous, on this ctx's device."""
        if not (t.is_cuda and t.is_contiguous() and t.dim() == 4):
            raise RslfError("set_epis_device needs a contiguous 4-d CUDA tensor")
        V, S, U, Cc = t.shape
        depth = {1: RSLF_DEPTH_8U, 2: RSLF_DEPTH_16U}.get(t.element_size(), RSLF_DEPTH_32F)
        self.check(lib().rslf_cuda_set_epis_device(self._h, C.c_void_p(t.data_ptr()), V, S, U, Cc, depth,
                                                   C.c_float(epi_scale_factor)), "rslf_cuda_set_epis_device")
        self.dims = (V, S, U, Cc)
        self._keep = t

    # ---- runs -------------------------------------------------------------------
    def depth1d_pile_run(self, dmin, dmax, dim_d, s_hat, params):
        self.check(lib().rslf_cuda_depth1d_pile_run(self._h, C.c_float(dmin), C.c_float(dmax), int(dim_d), int(s_hat),
                                                    C.byref(params)), "rslf_cuda_depth1d_pile_run")

    def depth1d_pile_get(self, want_raw=False):
        V, S, U, Cc = self.dims
        out = dict(best_depth=np.empty((V, U), np.float32), edge_conf=np.empty((V, U), np.float32),
                   edge_mask=np.empty((V, U), np.uint8), disp_conf=np.empty((V, U), np.float32),
                   rbar=np.empty((V, U, Cc), np.float32))
        self.check(lib().rslf_cuda_depth1d_pile_get(self._h, _fp(out["best_depth"]), _fp(out["edge_conf"]),
                                                    _bp(out["edge_mask"]), _fp(out["disp_conf"]), _fp(out["rbar"])),
                   "rslf_cuda_depth1d_pile_get")
        if want_raw:
            out["raw_depth"] = np.empty((V, U), np.float32)
            self.check(lib().rslf_cuda_depth1d_pile_get_raw_depth(self._h, _fp(out["raw_depth"])),
                       "rslf_cuda_depth1d_pile_get_raw_depth")
        return out

    def depth2d_run(self, dmin, dmax, dim_d, params, dmin_svu=None, dmax_svu=None):
        self.check(lib().rslf_cuda_depth2d_run(self._h, C.c_float(dmin), C.c_float(dmax), int(dim_d), C.byref(params),
                                               _fp(dmin_svu), _fp(dmax_svu)), "rslf_cuda_depth2d_run")

    def depth2d_get(self, outs=None):
        V, S, U, Cc = self.dims
        out = outs or dict(best_depth=np.empty((S, V, U), np.float32), edge_conf=np.empty((S, V, U), np.float32),
                           edge_mask=np.empty((S, V, U), np.uint8), disp_conf=np.empty((S, V, U), np.float32),
                           rbar=np.empty((S, V, U, Cc), np.float32))
        self.check(lib().rslf_cuda_depth2d_get(self._h, _fp(out.get("best_depth")), _fp(out.get("edge_conf")),
                                               _bp(out.get("edge_mask")), _fp(out.get("disp_conf")),
                                               _fp(out.get("rbar"))), "rslf_cuda_depth2d_get")
        return out

    def depth2d_valid_mask(self, accept_all, params):
        V, S, U, Cc = self.dims
        m = np.empty((S, V, U), np.uint8)
        self.check(lib().rslf_cuda_depth2d_get_valid_mask(self._h, int(bool(accept_all)), C.byref(params), _bp(m)),
                   "rslf_cuda_depth2d_get_valid_mask")
        return m

    def fine_to_coarse_run(self, dmin, dmax, dim_d, params, max_pyr_depth=-1, accept_all_last_scale=True):
        self.check(lib().rslf_cuda_fine_to_coarse_run(self._h, C.c_float(dmin), C.c_float(dmax), int(dim_d),
                                                      C.byref(params), int(max_pyr_depth),
                                                      int(bool(accept_all_last_scale))),
                   "rslf_cuda_fine_to_coarse_run")

    def fine_to_coarse_get(self, out_map=None, out_valid=None):
        V, S, U, Cc = self.dims
        if out_map is None:
            out_map = np.empty((S, V, U), np.float32)
        if out_valid is None:
            out_valid = np.empty((S, V, U), np.uint8)
        self.check(lib().rslf_cuda_fine_to_coarse_get(self._h, _fp(out_map), _bp(out_valid)),
                   "rslf_cuda_fine_to_coarse_get")
        return out_map, out_valid

    def fine_to_coarse_coloured(self, lut_bgr, params, saturate=True):
        """FineToCoarse::get_coloured_depth_maps of the last run -> ([S][V][U][3] uint8 BGR, (fit min, fit max))."""
        V, S, U, Cc = self.dims
        lut = np.ascontiguousarray(lut_bgr, np.uint8)
        if lut.size != 768:
            raise RslfError("the colour table must hold 256 x 3 bytes")
        out = np.empty((S, V, U, 3), np.uint8)
        mm = (C.c_double * 2)()
        self.check(lib().rslf_cuda_fine_to_coarse_get_coloured(self._h, _bp(lut), int(bool(saturate)), C.byref(params),
                                                               _bp(out), mm), "rslf_cuda_fine_to_coarse_get_coloured")
        return out, (mm[0], mm[1])

    def fine_to_coarse_levels(self):
        V, S, U, Cc = self.dims
        n = lib().rslf_cuda_fine_to_coarse_level_dims(self._h, 0, None, None)
        if n < 0:
            self.check(n, "rslf_cuda_fine_to_coarse_level_dims")
        levels = []
        for p in range(n):
            v = C.c_int()
            u = C.c_int()
            lib().rslf_cuda_fine_to_coarse_level_dims(self._h, p, C.byref(v), C.byref(u))
            Vp, Up = v.value, u.value
            d = dict(best_depth=np.empty((S, Vp, Up), np.float32), edge_conf=np.empty((S, Vp, Up), np.float32),
                     edge_mask=np.empty((S, Vp, Up), np.uint8), disp_conf=np.empty((S, Vp, Up), np.float32),
                     dmin=np.empty((S, Vp, Up), np.float32), dmax=np.empty((S, Vp, Up), np.float32))
            self.check(lib().rslf_cuda_fine_to_coarse_get_level(self._h, p, _fp(d["best_depth"]), _fp(d["edge_conf"]),
                                                                _bp(d["edge_mask"]), _fp(d["disp_conf"]),
                                                                _fp(d["dmin"]), _fp(d["dmax"])),
                       "rslf_cuda_fine_to_coarse_get_level")
            levels.append(d)
        return levels

    # ---- free functions ---------------------------------------------------------
    def edge_confidence(self, s, params):
        V, S, U, Cc = self.dims
        ce = np.empty((V, U), np.float32)
        m = np.empty((V, U), np.uint8)
        self.check(lib().rslf_cuda_edge_confidence(self._h, int(s), C.byref(params), _fp(ce), _bp(m)),
                   "rslf_cuda_edge_confidence")
        return ce, m

    def selective_median(self, src, mask, s_hat, size=5, eps=0.1):
        V, S, U, Cc = self.dims
        src = np.ascontiguousarray(src, np.float32)
        mask = np.ascontiguousarray(mask, np.uint8)
        dst = np.empty((V, U), np.float32)
        self.check(lib().rslf_cuda_selective_median(self._h, _fp(src), _bp(mask), int(s_hat), int(size), C.c_float(eps),
                                                    _fp(dst)), "rslf_cuda_selective_median")
        return dst

    def downsample_epis(self, raw):
        raw = np.ascontiguousarray(raw)
        V, S, U, Cc = raw.shape
        v2 = C.c_int()
        u2 = C.c_int()
        if raw.dtype == np.uint8:
            out = np.empty((int(np.rint(V * 0.5)), S, int(np.rint(U * 0.5)), Cc), np.uint8)
            self.check(lib().rslf_cuda_downsample_epis_u8(self._h, _bp(raw), V, S, U, Cc, _bp(out), C.byref(v2),
                                                          C.byref(u2)), "rslf_cuda_downsample_epis_u8")
            return out
        if raw.dtype == np.uint16:
            out = np.empty((int(np.rint(V * 0.5)), S, int(np.rint(U * 0.5)), Cc), np.uint16)
            self.check(lib().rslf_cuda_downsample_epis_u16(self._h, C.c_void_p(raw.ctypes.data), V, S, U, Cc,
                                                           C.c_void_p(out.ctypes.data), C.byref(v2), C.byref(u2)),
                       "rslf_cuda_downsample_epis_u16")
            return out
        raw = np.ascontiguousarray(raw, np.float32)
        out = np.empty((int(np.rint(V * 0.5)), S, int(np.rint(U * 0.5)), Cc), np.float32)
        self.check(lib().rslf_cuda_downsample_epis(self._h, _fp(raw), V, S, U, Cc, _fp(out), C.byref(v2), C.byref(u2)),
                   "rslf_cuda_downsample_epis")
        assert (v2.value, u2.value) == (out.shape[0], out.shape[2])
        return out

    def set_bounds(self, depth_up, valid_up, Vd, Ud, dmin, dmax):
        depth_up = np.ascontiguousarray(depth_up, np.float32)
        valid_up = np.ascontiguousarray(valid_up, np.uint8)
        S, Vu, Uu = depth_up.shape
        dmn = np.full((S, Vd, Ud), dmin, np.float32)
        dmx = np.full((S, Vd, Ud), dmax, np.float32)
        self.check(lib().rslf_cuda_set_bounds(self._h, _fp(depth_up), _bp(valid_up), S, Vu, Uu, int(Vd), int(Ud),
                                              _fp(dmn), _fp(dmx)), "rslf_cuda_set_bounds")
        return dmn, dmx

    def fuse_disp_maps(self, disp_p, valid_p):
        L = len(disp_p)
        disp_p = [np.ascontiguousarray(d, np.float32) for d in disp_p]
        valid_p = [np.ascontiguousarray(v, np.uint8) for v in valid_p]
        S = disp_p[0].shape[0]
        Vp = (C.c_int * L)(*[d.shape[1] for d in disp_p])
        Up = (C.c_int * L)(*[d.shape[2] for d in disp_p])
        dp = (C.POINTER(C.c_float) * L)(*[_fp(d) for d in disp_p])
        vp = (C.POINTER(C.c_uint8) * L)(*[_bp(v) for v in valid_p])
        out_map = np.empty(disp_p[0].shape, np.float32)
        out_valid = np.empty(disp_p[0].shape, np.uint8)
        self.check(lib().rslf_cuda_fuse_disp_maps(self._h, L, S, Vp, Up, dp, vp, _fp(out_map), _bp(out_valid)),
                   "rslf_cuda_fuse_disp_maps")
        return out_map, out_valid

    # ---- misc -------------------------------------------------------------------
    def timing(self):
        t = Timing()
        self.check(lib().rslf_cuda_last_timing(self._h, C.byref(t)), "rslf_cuda_last_timing")
        return t.as_dict()

    def set_stage_timing(self, level):
        """0: no CUDA events inside a run, 1 (default): around the depth kernel only, 2: around every stage."""
        lib().rslf_cuda_set_stage_timing(self._h, int(level))

    def set_decision_log(self, capacity):
        """Diagnostics: record every pixel decision of the following runs (0: off)."""
        self.check(lib().rslf_cuda_set_decision_log(self._h, C.c_size_t(int(capacity))), "rslf_cuda_set_decision_log")
        self._dlog_cap = int(capacity)

    def decision_log(self):
        """(ctypes array of Decision, n) of the last run."""
        buf = (Decision * max(1, self._dlog_cap))()
        n = C.c_size_t(0)
        self.check(lib().rslf_cuda_get_decision_log(self._h, buf, C.c_size_t(self._dlog_cap), C.byref(n)), "rslf_cuda_get_decision_log")
        if n.value > self._dlog_cap:
            raise RslfError("decision log overflow: %d records, capacity %d" % (n.value, self._dlog_cap))
        return buf, n.value

    def set_confidence_criterion(self, name):
        """"edge" (the reference's default build) or "disp" (-D_USE_DISP_CONFIDENCE_SCORE as intended: C_d > par_disp_score_threshold
        gates propagation and validity)."""
        self.check(lib().rslf_cuda_set_confidence_criterion(self._h, {"edge": 0, "disp": 1, "line": 2}[name]),
                   "rslf_cuda_set_confidence_criterion")

    def set_balance(self, on):
        self.check(lib().rslf_cuda_set_balance(self._h, int(bool(on))), "rslf_cuda_set_balance")

    def last_spans(self):
        """[(stage id, ms)] of the last run in issue order (see rslf_cuda_last_spans)."""
        n = C.c_size_t(0)
        self.check(lib().rslf_cuda_last_spans(self._h, None, None, C.c_size_t(0), C.byref(n)), "rslf_cuda_last_spans")
        st = (C.c_int * max(1, n.value))()
        ms = (C.c_float * max(1, n.value))()
        self.check(lib().rslf_cuda_last_spans(self._h, st, ms, C.c_size_t(n.value), C.byref(n)), "rslf_cuda_last_spans")
        return [(st[i], ms[i]) for i in range(n.value)]

    def set_fast_math(self, on):
        """Opt-in contracted (FMA) mean shift for RGB stacks: faster, within the specified tolerance, not bit-identical."""
        self.check(lib().rslf_cuda_set_fast_math(self._h, int(bool(on))), "rslf_cuda_set_fast_math")

    def sync(self):
        self.check(lib().rslf_cuda_sync(self._h), "rslf_cuda_sync")

    def flush_l2(self):
        self.check(lib().rslf_cuda_flush_l2(self._h), "rslf_cuda_flush_l2")

    def measure_fp32_peak(self):
        a = C.c_double()
        b = C.c_double()
        self.check(lib().rslf_cuda_measure_fp32_peak(self._h, C.byref(a), C.byref(b)), "rslf_cuda_measure_fp32_peak")
        return a.value, b.value

    def measure_fp32x2_peak(self):
        a = C.c_double()
        self.check(lib().rslf_cuda_measure_fp32x2_peak(self._h, C.byref(a)), "rslf_cuda_measure_fp32x2_peak")
        return a.value

    def row_work(self):
        """Pixels evaluated per local image row in the last run (numpy uint32 [V])."""
        V = self.dims[0]
        out = np.zeros(V, np.uint32)
        self.check(lib().rslf_cuda_get_row_work(self._h, out.ctypes.data_as(C.POINTER(C.c_uint))), "rslf_cuda_get_row_work")
        return out

    def comm_init(self, uid_bytes, rank, world):
        buf = (C.c_char * 128).from_buffer_copy(uid_bytes)
        self.check(lib().rslf_cuda_comm_init(self._h, buf, int(rank), int(world)), "rslf_cuda_comm_init")

    def set_row_shards(self, row_starts):
        """row_starts: world + 1 global row boundaries (see remotesensingproject_b200.shard.row_shards)."""
        n = len(row_starts) - 1
        arr = (C.c_int * (n + 1))(*[int(x) for x in row_starts])
        self.check(lib().rslf_cuda_set_row_shards(self._h, arr, n), "rslf_cuda_set_row_shards")


def nccl_unique_id():
    buf = (C.c_char * 128)()
    rc = lib().rslf_cuda_nccl_unique_id(buf)
    if rc != 0:
        raise RslfError("rslf_cuda_nccl_unique_id failed: %s" % lib().rslf_cuda_strerror(rc).decode())
    return bytes(buf)


def _as_mats(epis):
    """Normalises the accepted inputs to a list of [rows][cols][C] host arrays sharing dtype / step."""
    if hasattr(epis, "numpy") and not isinstance(epis, np.ndarray):      # CPU torch tensor (possibly pinned)
        epis = epis.numpy()
    if isinstance(epis, np.ndarray):
        if epis.ndim == 3:
            epis = epis[..., None]
        if epis.ndim != 4:
            raise RslfError("expected [V][S][U][C] or [V][S][U]")
        if epis.dtype not in _DEPTH_OF:
            raise RslfError("only uint8 (CV_8U), uint16 (CV_16U) and float32 (CV_32F) inputs are implemented")
        if not epis[0].flags["C_CONTIGUOUS"]:
            epis = np.ascontiguousarray(epis)
        return [epis[v] for v in range(epis.shape[0])]
    mats = []
    for m in epis:
        if hasattr(m, "numpy") and not isinstance(m, np.ndarray):
            m = m.numpy()
        if m.ndim == 2:
            m = m[..., None]
        if m.dtype not in _DEPTH_OF:
            raise RslfError("only uint8 (CV_8U), uint16 (CV_16U) and float32 (CV_32F) inputs are implemented")
        if m.strides[2] != m.itemsize or m.strides[1] != m.itemsize * m.shape[2]:
            m = np.ascontiguousarray(m)
        mats.append(m)
    if len({(m.shape, m.dtype, m.strides[0]) for m in mats}) != 1:
        raise RslfError("all EPIs must share shape, dtype and row step")
    return mats


# ---------------------------------------------------------------------------
# Mirrors of the reference classes
# ---------------------------------------------------------------------------
class Depth1DComputer:
    """rslf::Depth1DComputer<T> (rslf_depth_computation.hpp:26-68, ctor :254-317, run :319-363): ONE EPI [S][U][C],
    edge confidence and depth of line s_hat, no selective median.  The reference keeps the results private; they are
    members here (m_best_depth_u, m_edge_confidence_u, m_edge_confidence_mask_u, m_disp_confidence_u, m_rbar_u).
    Runs as a one-row pile on the device: the per-line results of the two classes coincide except for the median."""

    def __init__(self, epi, dmin, dmax, dim_d, s_hat=-1, epi_scale_factor=-1.0, parameters=None, device=0, ctx=None):
        self.m_ctx = ctx or Context(device)
        # compute_1D_edge_confidence is called directly (dc.hpp:339): the opening of the pile wrapper does not apply
        self.m_parameters = Params.from_buffer_copy(bytes(parameters)) if parameters is not None else default_params()
        self.m_parameters.edge_confidence_opening_size = 1
        epi = np.asarray(epi)
        if epi.ndim == 2:
            epi = epi[..., None]
        self.m_ctx.upload_epis(epi[None], epi_scale_factor)
        S = epi.shape[0]
        self.m_dim_d = dim_d
        self.m_dmin, self.m_dmax = dmin, dmax
        if s_hat < 0 or s_hat > S - 1:
            s_hat = int(np.floor((0.0 + S) / 2))
        self.m_s_hat = s_hat

    def run(self):
        self.m_ctx.depth1d_pile_run(self.m_dmin, self.m_dmax, self.m_dim_d, self.m_s_hat, self.m_parameters)
        r = self.m_ctx.depth1d_pile_get(want_raw=True)
        self.m_best_depth_u = r["raw_depth"][0]
        self.m_edge_confidence_u = r["edge_conf"][0]
        self.m_edge_confidence_mask_u = r["edge_mask"][0]
        self.m_disp_confidence_u = r["disp_conf"][0]
        self.m_rbar_u = r["rbar"][0]
        return self


class Depth1DComputer_pile:
    """rslf::Depth1DComputer_pile<T> (rslf_depth_computation.hpp:93-143, ctor :425-511, run :513-565)."""

    def __init__(self, epis, dmin, dmax, dim_d, s_hat=-1, epi_scale_factor=-1.0, parameters=None, device=0, ctx=None):
        self.m_ctx = ctx or Context(device)
        self.m_parameters = parameters or default_params()
        if hasattr(epis, "is_cuda") and epis.is_cuda:
            self.m_ctx.set_epis_device(epis, epi_scale_factor)
        else:
            self.m_ctx.upload_epis(epis, epi_scale_factor)
        V, S, U, Cc = self.m_ctx.dims
        self.m_dim_d = dim_d
        self.m_dmin, self.m_dmax = dmin, dmax
        if s_hat < 0 or s_hat > S - 1:
            s_hat = int(np.floor((0.0 + S) / 2))
        self.m_s_hat = s_hat

    def get_s_hat(self):
        return self.m_s_hat

    def run(self, fetch=True):
        self.m_ctx.depth1d_pile_run(self.m_dmin, self.m_dmax, self.m_dim_d, self.m_s_hat, self.m_parameters)
        if fetch:
            r = self.m_ctx.depth1d_pile_get(want_raw=True)
            self.m_best_depth_v_u = r["best_depth"]
            self.m_edge_confidence_v_u = r["edge_conf"]
            self.m_edge_confidence_mask_v_u = r["edge_mask"]
            self.m_disp_confidence_v_u = r["disp_conf"]
            self.m_rbar_v_u = r["rbar"]
            self.m_raw_depth_v_u = r["raw_depth"]
        return self


class Depth2DComputer:
    """rslf::Depth2DComputer<T> (rslf_depth_computation.hpp:166-229, ctor :651-746, run :748-805)."""

    def __init__(self, epis, dmin, dmax, dim_d, epi_scale_factor=-1.0, parameters=None, verbose=True, device=0,
                 ctx=None):
        self.m_ctx = ctx or Context(device)
        self.m_parameters = parameters or default_params()
        if hasattr(epis, "is_cuda") and epis.is_cuda:
            self.m_ctx.set_epis_device(epis, epi_scale_factor)
        else:
            self.m_ctx.upload_epis(epis, epi_scale_factor, params=self.m_parameters)      # pipelined ingest
        self.m_dim_d = dim_d
        self.m_dmin, self.m_dmax = dmin, dmax
        self.m_dmin_s_v_u = None
        self.m_dmax_s_v_u = None
        self.m_accept_all = False
        self.m_verbose = verbose

    def edit_dmin(self):
        if self.m_dmin_s_v_u is None:
            V, S, U, Cc = self.m_ctx.dims
            self.m_dmin_s_v_u = np.full((S, V, U), self.m_dmin, np.float32)
            self.m_dmax_s_v_u = np.full((S, V, U), self.m_dmax, np.float32)
        return self.m_dmin_s_v_u

    def edit_dmax(self):
        self.edit_dmin()
        return self.m_dmax_s_v_u

    def set_accept_all(self, b):
        self.m_accept_all = bool(b)

    def run(self, fetch=True):
        self.m_ctx.depth2d_run(self.m_dmin, self.m_dmax, self.m_dim_d, self.m_parameters, self.m_dmin_s_v_u,
                               self.m_dmax_s_v_u)
        if fetch:
            r = self.m_ctx.depth2d_get()
            self.m_best_depth_s_v_u = r["best_depth"]
            self.m_edge_confidence_s_v_u = r["edge_conf"]
            self.m_edge_confidence_mask_s_v_u = r["edge_mask"]
            self.m_disp_confidence_s_v_u = r["disp_conf"]
            self.m_rbar_s_v_u = r["rbar"]
        return self

    def get_depths_s_v_u(self):
        return self.m_best_depth_s_v_u

    def get_valid_depths_mask_s_v_u(self):
        return self.m_ctx.depth2d_valid_mask(self.m_accept_all, self.m_parameters)


class FineToCoarse:
    """rslf::FineToCoarse<T> (rslf_fine_to_coarse.hpp:26-81, ctor :103-159, run :171-299, get_results :301-322)."""

    def __init__(self, epis, d_min, d_max, dim_d, epi_scale_factor=-1.0, parameters=None, max_pyr_depth=-1,
                 accept_all_last_scale=True, device=0, ctx=None):
        self.m_ctx = ctx or Context(device)
        self.m_parameters = parameters or default_params()
        if hasattr(epis, "is_cuda") and epis.is_cuda:
            self.m_ctx.set_epis_device(epis, epi_scale_factor)
        else:
            self.m_ctx.upload_epis(epis, epi_scale_factor, params=self.m_parameters)      # pipelined ingest
        self.m_dim_d = dim_d
        self.m_dmin, self.m_dmax = d_min, d_max
        self.m_max_pyr_depth = max_pyr_depth
        self.m_accept_all_last_scale = accept_all_last_scale

    def run(self):
        self.m_ctx.fine_to_coarse_run(self.m_dmin, self.m_dmax, self.m_dim_d, self.m_parameters, self.m_max_pyr_depth,
                                      self.m_accept_all_last_scale)
        return self

    def get_results(self, out_map=None, out_valid=None):
        """-> (out_map_s_v_u [S][V][U] float32, out_validity_s_v_u [S][V][U] uint8)."""
        return self.m_ctx.fine_to_coarse_get(out_map, out_valid)

    def get_coloured_depth_maps(self, cv_colormap=2, saturate=True, lut_bgr=None):
        """FineToCoarse::get_coloured_depth_maps(out, cv_colormap = COLORMAP_JET, saturate) (ftc.hpp:324-377) ->
        [S][V][U][3] uint8 (BGR).  The colour table is OpenCV's: taken from cv2.applyColorMap of the ramp 0..255 unless
        a 256 x 3 table is given."""
        if lut_bgr is None:
            import cv2
            lut_bgr = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(256, 1), int(cv_colormap)).reshape(256, 3)
        return self.m_ctx.fine_to_coarse_coloured(lut_bgr, self.m_parameters, saturate)[0]

    def get_levels(self):
        return self.m_ctx.fine_to_coarse_levels()
