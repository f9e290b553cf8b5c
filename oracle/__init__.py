"""ctypes binding of the CPU oracle (oracle/rslf_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


class Params(C.Structure):
    """Mirror of rslf_params (include/rslf_b200.h)."""
    _fields_ = [
        ("edge_score_threshold", C.c_float), ("line_score_threshold", C.c_float),
        ("disp_score_threshold", C.c_float), ("raw_score_threshold", C.c_float),
        ("mean_shift_max_iter", C.c_int), ("edge_confidence_filter_size", C.c_int),
        ("edge_confidence_opening_type", C.c_int), ("edge_confidence_opening_size", C.c_int),
        ("median_filter_size", C.c_int), ("median_filter_epsilon", C.c_float),
        ("propagation_epsilon", C.c_float), ("slope_factor", C.c_float),
        ("cut_shadows", C.c_int), ("shadow_level", C.c_float), ("kernel_h", C.c_float),
    ]


def build(force=False):
    src = os.path.join(_HERE, "rslf_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_normalise.restype = C.c_float
        _lib.orc_depth1d_pile.restype = C.c_double
        _lib.orc_depth2d.restype = C.c_double
        _lib.orc_fine_to_coarse.restype = C.c_double
    return _lib


def default_params(**kw):
    p = Params()
    lib().orc_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _f(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _b(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_uint8))


def _c32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _cv_depth(a):
    """OpenCV depth code of a raw stack: CV_8U = 0, CV_16U = 2, CV_32F = 5 (anything else is converted to float32)."""
    return {np.dtype(np.uint8): 0, np.dtype(np.uint16): 2}.get(a.dtype, 5)


class Decision(C.Structure):
    """rslf_decision (include/rslf_b200.h)."""
    _fields_ = [("level", C.c_int), ("s_hat", C.c_int), ("pix", C.c_int), ("index", C.c_int), ("score", C.c_float),
                ("rbar", C.c_float * 3), ("disp_conf", C.c_float), ("disparity", C.c_float), ("accepted", C.c_int),
                ("reserved", C.c_int)]


class replay:
    """with oracle.replay(records, n): ... runs the oracle entry points with the recorded decisions adopted after the
    tolerance check (see `Replay` in rslf_oracle.cpp); .report holds the counters afterwards."""

    def __init__(self, records, n, margin_tol=1e-5, rel_tol=1e-4):
        self.records, self.n, self.margin_tol, self.rel_tol = records, int(n), margin_tol, rel_tol
        self.report = None

    def __enter__(self):
        lib().orc_replay_begin(self.records, C.c_longlong(self.n), C.c_float(self.margin_tol), C.c_float(self.rel_tol))
        return self

    def __exit__(self, *exc):
        out = (C.c_longlong * 10)()
        worst = (C.c_double * 2)()
        lib().orc_replay_end(out, worst)
        keys = ("looked_up", "missing", "index_changed", "bad_margin", "bad_score", "bad_rbar", "bad_cd", "bad_disparity", "records")
        self.report = dict(zip(keys, out[:9]))
        self.report["worst_margin"], self.report["worst_rel_score"] = worst[0], worst[1]
        return False


def set_criterion(name):
    """"edge": the reference as written (default build); "disp": -D_USE_DISP_CONFIDENCE_SCORE as intended (`#elseif` -> `#elif`):
    propagation sources and validity maps are gated by C_d > par_disp_score_threshold (core.hpp:1097-1098, dc.hpp:901-902)."""
    lib().orc_set_criterion({"edge": 0, "disp": 1}[name])


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def normalise(raw, scale_factor=-1.0):
    """Computer ctor input normalisation.  raw: [V][S][U][C] uint8 or float32."""
    raw = np.ascontiguousarray(raw)
    V, S, U, Cc = raw.shape
    depth = _cv_depth(raw)
    if depth == 5:
        raw = _c32(raw)
    out = np.empty(raw.shape, np.float32)
    lib().orc_normalise(raw.ctypes.data_as(C.c_void_p), depth, V, S, U, Cc, C.c_float(scale_factor), _f(out))
    return out


def edge_confidence(epis, s, params=None):
    epis = _c32(epis)
    V, S, U, Cc = epis.shape
    p = params or default_params()
    ce = np.zeros((V, U), np.float32)
    mask = np.zeros((V, U), np.uint8)
    lib().orc_edge_confidence(_f(epis), V, S, U, Cc, int(s), C.byref(p), _f(ce), _b(mask))
    return ce, mask


def structuring_element(shape, k):
    """cv::getStructuringElement(shape, (k, k)) as the edge-mask opening builds it (core.hpp:762-766)."""
    out = np.zeros((k, k), np.uint8)
    lib().orc_structuring_element(int(shape), int(k), _b(out))
    return out


def morph_open(mask, shape, k):
    """cv::morphologyEx(mask, mask, MORPH_OPEN, getStructuringElement(shape, (k, k))) (core.hpp:768)."""
    out = np.ascontiguousarray(mask, np.uint8).copy()
    V, U = out.shape
    lib().orc_morph_open(_b(out), V, U, int(shape), int(k))
    return out


def pixel_scores(epi, s_hat, u, dmin, dmax, D, params=None):
    """epi: [S][U][C] normalised.  Returns scores[D], rbar[D][C], dvals[D]."""
    epi = _c32(epi)
    S, U, Cc = epi.shape
    p = params or default_params()
    scores = np.zeros(D, np.float32)
    rbar = np.zeros((D, Cc), np.float32)
    dv = np.zeros(D, np.float32)
    lib().orc_pixel_scores(_f(epi), S, U, Cc, D, int(s_hat), int(u), C.c_float(dmin), C.c_float(dmax),
                           C.byref(p), _f(scores), _f(rbar), _f(dv))
    return scores, rbar, dv


def selective_median(src, mask, epis, s_hat, size=5, eps=0.1):
    epis = _c32(epis)
    V, S, U, Cc = epis.shape
    src = _c32(src)
    mask = np.ascontiguousarray(mask, np.uint8)
    dst = np.zeros((V, U), np.float32)
    lib().orc_selective_median(_f(src), _b(mask), _f(epis), V, S, U, Cc, int(s_hat), int(size), C.c_float(eps), _f(dst))
    return dst


def depth1d_pile(epis, dmin, dmax, D, s_hat=-1, params=None):
    """Depth1DComputer_pile::run on a normalised stack.  Returns a dict of maps."""
    epis = _c32(epis)
    V, S, U, Cc = epis.shape
    p = params or default_params()
    out = dict(
        best_depth=np.zeros((V, U), np.float32), edge_conf=np.zeros((V, U), np.float32),
        edge_mask=np.zeros((V, U), np.uint8), disp_conf=np.zeros((V, U), np.float32),
        rbar=np.zeros((V, U, Cc), np.float32), raw_depth=np.zeros((V, U), np.float32),
        margin=np.zeros((V, U), np.float32), best_idx=np.zeros((V, U), np.int32))
    n = lib().orc_depth1d_pile(_f(epis), V, S, U, Cc, C.c_float(dmin), C.c_float(dmax), int(D), int(s_hat),
                               C.byref(p), _f(out["best_depth"]), _f(out["edge_conf"]), _b(out["edge_mask"]),
                               _f(out["disp_conf"]), _f(out["rbar"]), _f(out["raw_depth"]), _f(out["margin"]),
                               out["best_idx"].ctypes.data_as(C.POINTER(C.c_int32)))
    out["computed_pixels"] = n
    return out


def depth1d(epi, dmin, dmax, D, s_hat=-1, params=None):
    """Depth1DComputer::run (dc.hpp:319-363) on ONE normalised EPI [S][U][C]: compute_1D_edge_confidence (no opening:
    that belongs to the pile wrapper, core.hpp:759) and compute_1D_depth_epi of line s_hat; no selective median, so
    best_depth is the argmax disparity itself."""
    p = Params.from_buffer_copy(bytes(params)) if params is not None else default_params()
    p.edge_confidence_opening_size = 1
    r = depth1d_pile(_c32(epi)[None], dmin, dmax, D, s_hat=s_hat, params=p)
    return dict(best_depth=r["raw_depth"][0], edge_conf=r["edge_conf"][0], edge_mask=r["edge_mask"][0],
                disp_conf=r["disp_conf"][0], rbar=r["rbar"][0], computed_pixels=r["computed_pixels"])


def visiting_order(S):
    buf = (C.c_int * (S + 1))()
    n = lib().orc_visiting_order(int(S), buf)
    return list(buf[:n])


def depth2d(epis, dmin, dmax, D, params=None, dmin_svu=None, dmax_svu=None):
    """Depth2DComputer::run on a normalised stack."""
    epis = _c32(epis)
    V, S, U, Cc = epis.shape
    p = params or default_params()
    out = dict(
        best_depth=np.zeros((S, V, U), np.float32), edge_conf=np.zeros((S, V, U), np.float32),
        edge_mask=np.zeros((S, V, U), np.uint8), disp_conf=np.zeros((S, V, U), np.float32),
        rbar=np.zeros((S, V, U, Cc), np.float32))
    per_pass = np.zeros(S + 1, np.float64)
    if dmin_svu is not None:
        dmin_svu = _c32(dmin_svu)
        dmax_svu = _c32(dmax_svu)
    n = lib().orc_depth2d(_f(epis), V, S, U, Cc, C.c_float(dmin), C.c_float(dmax), int(D), C.byref(p),
                          _f(dmin_svu), _f(dmax_svu), _f(out["best_depth"]), _f(out["edge_conf"]),
                          _b(out["edge_mask"]), _f(out["disp_conf"]), _f(out["rbar"]),
                          per_pass.ctypes.data_as(C.POINTER(C.c_double)))
    out["computed_pixels"] = n
    out["computed_per_pass"] = per_pass[:len(visiting_order(S))]
    return out


def half_size(n):
    return lib().orc_half_size(int(n))


def downsample(raw):
    raw = np.ascontiguousarray(raw)
    if raw.dtype == np.uint8:
        V, S, U, Cc = raw.shape
        out = np.zeros((half_size(V), S, half_size(U), Cc), np.uint8)
        lib().orc_downsample_u8(_b(raw), V, S, U, Cc, _b(out))
        return out
    if raw.dtype == np.uint16:
        V, S, U, Cc = raw.shape
        out = np.zeros((half_size(V), S, half_size(U), Cc), np.uint16)
        lib().orc_downsample_u16(raw.ctypes.data_as(C.c_void_p), V, S, U, Cc, out.ctypes.data_as(C.c_void_p))
        return out
    raw = _c32(raw)
    V, S, U, Cc = raw.shape
    out = np.zeros((half_size(V), S, half_size(U), Cc), np.float32)
    lib().orc_downsample(_f(raw), V, S, U, Cc, _f(out))
    return out


def set_bounds(depth_up, valid_up, Vd, Ud, dmin, dmax):
    depth_up = _c32(depth_up)
    valid_up = np.ascontiguousarray(valid_up, np.uint8)
    S, Vu, Uu = depth_up.shape
    dmn = np.full((S, Vd, Ud), dmin, np.float32)
    dmx = np.full((S, Vd, Ud), dmax, np.float32)
    lib().orc_set_bounds(_f(depth_up), _b(valid_up), S, Vu, Uu, int(Vd), int(Ud), _f(dmn), _f(dmx))
    return dmn, dmx


def fuse(disp_p, valid_p):
    """fuse_disp_maps.  disp_p / valid_p: lists (finest first) of [S][V_p][U_p]."""
    L = len(disp_p)
    disp_p = [_c32(d) for d in disp_p]
    valid_p = [np.ascontiguousarray(v, np.uint8) for v in valid_p]
    S = disp_p[0].shape[0]
    Vp = (C.c_int * L)(*[d.shape[1] for d in disp_p])
    Up = (C.c_int * L)(*[d.shape[2] for d in disp_p])
    dp = (C.POINTER(C.c_float) * L)(*[_f(d) for d in disp_p])
    vp = (C.POINTER(C.c_uint8) * L)(*[_b(v) for v in valid_p])
    out_map = np.zeros(disp_p[0].shape, np.float32)
    out_valid = np.zeros(disp_p[0].shape, np.uint8)
    lib().orc_fuse(L, S, Vp, Up, dp, vp, _f(out_map), _b(out_valid))
    return out_map, out_valid


def colour_maps(out_map, out_valid, epis0_norm, lut_bgr, saturate=True, cut_shadows=True):
    """FineToCoarse::get_coloured_depth_maps (ftc.hpp:324-377) from the fused results of fine_to_coarse():
    out_map / out_valid [S][V][U], epis0_norm = normalise(raw) [V][S][U][C], lut_bgr = 256 x 3 colour table.
    Returns ([S][V][U][3] uint8, (fit min, fit max))."""
    out_map = _c32(out_map)
    out_valid = np.ascontiguousarray(out_valid, np.uint8)
    epis0_norm = _c32(epis0_norm)
    lut = np.ascontiguousarray(lut_bgr, np.uint8).reshape(256, 3)
    V, S, U, Cc = epis0_norm.shape
    out = np.zeros((S, V, U, 3), np.uint8)
    mm = np.zeros(2, np.float64)
    lib().orc_colour_maps(_f(out_map), _b(out_valid), _f(epis0_norm), V, S, U, Cc, _b(lut), int(bool(saturate)),
                          int(bool(cut_shadows)), _b(out), mm.ctypes.data_as(C.POINTER(C.c_double)))
    return out, (float(mm[0]), float(mm[1]))


def resize_linear(src, Vd, Ud):
    src = _c32(src)
    dst = np.zeros((Vd, Ud), np.float32)
    lib().orc_resize_linear(_f(src), src.shape[0], src.shape[1], _f(dst), int(Vd), int(Ud))
    return dst


def median3x3(src):
    src = _c32(src)
    dst = np.zeros_like(src)
    lib().orc_median3x3(_f(src), _f(dst), src.shape[0], src.shape[1])
    return dst


def pyramid_dims(V, U, max_pyr_depth=-1):
    Vp = (C.c_int * 32)()
    Up = (C.c_int * 32)()
    n = lib().orc_pyramid_dims(int(V), int(U), int(max_pyr_depth), Vp, Up)
    return [(Vp[i], Up[i]) for i in range(n)]


def fine_to_coarse(raw, dmin, dmax, D, scale_factor=-1.0, params=None, max_pyr_depth=-1,
                   accept_all_last=True, want_levels=True):
    """FineToCoarse ctor + run + get_results on a raw [V][S][U][C] stack."""
    raw = np.ascontiguousarray(raw)
    V, S, U, Cc = raw.shape
    depth = _cv_depth(raw)
    if depth == 5:
        raw = _c32(raw)
    p = params or default_params()
    dims = pyramid_dims(V, U, max_pyr_depth)
    L = len(dims)
    out_map = np.zeros((S, V, U), np.float32)
    out_valid = np.zeros((S, V, U), np.uint8)
    levels = []
    for (Vp, Up) in dims:
        levels.append(dict(
            best_depth=np.zeros((S, Vp, Up), np.float32), edge_conf=np.zeros((S, Vp, Up), np.float32),
            edge_mask=np.zeros((S, Vp, Up), np.uint8), disp_conf=np.zeros((S, Vp, Up), np.float32),
            dmin=np.zeros((S, Vp, Up), np.float32), dmax=np.zeros((S, Vp, Up), np.float32)))

    def arr(key, fn, ty):
        return (C.POINTER(ty) * L)(*[fn(l[key]) for l in levels])
    per_level = np.zeros(max(L, 1), np.float64)
    n = lib().orc_fine_to_coarse(
        raw.ctypes.data_as(C.c_void_p), depth, V, S, U, Cc, C.c_float(scale_factor), C.c_float(dmin),
        C.c_float(dmax), int(D), C.byref(p), int(max_pyr_depth), int(bool(accept_all_last)), _f(out_map),
        _b(out_valid), arr("best_depth", _f, C.c_float), arr("edge_conf", _f, C.c_float),
        arr("edge_mask", _b, C.c_uint8), arr("disp_conf", _f, C.c_float), arr("dmin", _f, C.c_float),
        arr("dmax", _f, C.c_float), per_level.ctypes.data_as(C.POINTER(C.c_double)))
    if n < 0:
        raise ValueError("oracle: unsupported input for fine_to_coarse")
    return dict(map=out_map, valid=out_valid, levels=levels, dims=dims, computed_pixels=n,
                computed_per_level=per_level[:L], samples=float(n) * D * S)
