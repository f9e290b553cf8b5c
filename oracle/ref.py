"""ctypes binding of oracle/_ref/librslf_ref.so: the REFERENCE'S OWN sources (/root/reference/RSLightFields,
compiled where they lie) against the stand-in OpenCV headers of oracle/cvshim (see oracle/ref_driver.cpp).

TEST INFRASTRUCTURE ONLY.  Used by tests/test_oracle_vs_reference.py to pin the restated oracle against the
reference's control flow, by tests/golden/make_ref_golden.py to write fixtures, and by the CPU legs of bench.py.
The library is built in the CPU container (the reference tree does not exist on the GPU box) and travels with
the repository snapshot; `available()` says whether it is there.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import Params, default_params  # noqa: F401  (same rslf_params layout)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "librslf_ref.so")
# the reference built with -D_USE_DISP_CONFIDENCE_SCORE after `#elseif` -> `#elif` (oracle/Makefile, target ref_cd)
LIB_CD_PATH = os.path.join(_HERE, "_ref", "librslf_ref_cd.so")
REFERENCE_ROOT = "/root/reference/RSLightFields"
_lib = None
_lib_cd = None
_criterion = "edge"


def set_criterion(name):
    """"edge": the default build; "disp": the disparity-confidence build (see LIB_CD_PATH)."""
    global _criterion
    assert name in ("edge", "disp")
    _criterion = name


def available_cd():
    return os.path.exists(LIB_CD_PATH) or os.path.isdir(REFERENCE_ROOT)


def build(force=False):
    """Compiles the reference from its own tree when that tree is present; otherwise uses the prebuilt file."""
    if os.path.isdir(REFERENCE_ROOT) and (force or not os.path.exists(LIB_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"] + (["-B"] if force else []))
    return LIB_PATH


def available():
    return os.path.exists(LIB_PATH) or os.path.isdir(REFERENCE_ROOT)


def lib():
    global _lib, _lib_cd
    if _criterion == "disp":
        if _lib_cd is None:
            if os.path.isdir(REFERENCE_ROOT) and not os.path.exists(LIB_CD_PATH):
                subprocess.check_call(["make", "-C", _HERE, "-s", "ref_cd"])
            if not os.path.exists(LIB_CD_PATH):
                raise RuntimeError("oracle/_ref/librslf_ref_cd.so is missing and %s is not present to build it" % REFERENCE_ROOT)
            _lib_cd = C.CDLL(LIB_CD_PATH)
        return _lib_cd
    if _lib is None:
        build()
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("oracle/_ref/librslf_ref.so is missing and %s is not present to build it" % REFERENCE_ROOT)
        _lib = C.CDLL(LIB_PATH)
    return _lib


def num_threads():
    return lib().ref_num_threads()


def set_num_threads(n):
    lib().ref_set_num_threads(int(n))


def _f(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))


def _b(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_uint8))


def _raw(raw):
    raw = np.ascontiguousarray(raw)
    if raw.dtype not in (np.uint8, np.uint16):
        raw = np.ascontiguousarray(raw, dtype=np.float32)
    return raw, {np.dtype(np.uint8): 0, np.dtype(np.uint16): 2}.get(raw.dtype, 5)      # CV_8U, CV_16U, CV_32F


def depth1d(raw_epi, dmin, dmax, D, s_hat=-1, scale_factor=-1.0, params=None):
    """rslf::Depth1DComputer<T>(epi, dmin, dmax, D, s_hat, scale, params).run() on ONE raw EPI [S][U][C]
    (dc.hpp:254-363): edge confidence and depth of line s_hat, no median."""
    raw, depth = _raw(np.asarray(raw_epi)[None])
    _, S, U, Cc = raw.shape
    p = params or default_params()
    out = dict(best_depth=np.zeros(U, np.float32), edge_conf=np.zeros(U, np.float32), edge_mask=np.zeros(U, np.uint8),
               disp_conf=np.zeros(U, np.float32), rbar=np.zeros((U, Cc), np.float32))
    rc = lib().ref_depth1d(raw.ctypes.data_as(C.c_void_p), depth, S, U, Cc, C.c_float(scale_factor), C.c_float(dmin),
                           C.c_float(dmax), int(D), int(s_hat), C.byref(p), _f(out["best_depth"]), _f(out["edge_conf"]),
                           _b(out["edge_mask"]), _f(out["disp_conf"]), _f(out["rbar"]))
    if rc < 0:
        raise ValueError("reference: unsupported channel count")
    out["s_hat"] = rc
    return out


def depth1d_pile(raw, dmin, dmax, D, s_hat=-1, scale_factor=-1.0, params=None):
    """rslf::Depth1DComputer_pile<T>(epis, dmin, dmax, D, s_hat, scale, params).run() on a RAW [V][S][U][C] stack."""
    raw, depth = _raw(raw)
    V, S, U, Cc = raw.shape
    p = params or default_params()
    out = dict(best_depth=np.zeros((V, U), np.float32), edge_conf=np.zeros((V, U), np.float32),
               edge_mask=np.zeros((V, U), np.uint8), disp_conf=np.zeros((V, U), np.float32),
               rbar=np.zeros((V, U, Cc), np.float32))
    rc = lib().ref_depth1d_pile(raw.ctypes.data_as(C.c_void_p), depth, V, S, U, Cc, C.c_float(scale_factor), C.c_float(dmin),
                                C.c_float(dmax), int(D), int(s_hat), C.byref(p), _f(out["best_depth"]), _f(out["edge_conf"]),
                                _b(out["edge_mask"]), _f(out["disp_conf"]), _f(out["rbar"]))
    if rc < 0:
        raise ValueError("reference: unsupported channel count")
    out["s_hat"] = rc
    return out


def depth2d(raw, dmin, dmax, D, scale_factor=-1.0, params=None, dmin_svu=None, dmax_svu=None, accept_all=False):
    """rslf::Depth2DComputer<T>: ctor, optional per-pixel bounds through edit_dmin / edit_dmax, run, validity masks."""
    raw, depth = _raw(raw)
    V, S, U, Cc = raw.shape
    p = params or default_params()
    out = dict(best_depth=np.zeros((S, V, U), np.float32), edge_conf=np.zeros((S, V, U), np.float32),
               edge_mask=np.zeros((S, V, U), np.uint8), disp_conf=np.zeros((S, V, U), np.float32),
               rbar=np.zeros((S, V, U, Cc), np.float32), valid=np.zeros((S, V, U), np.uint8))
    if dmin_svu is not None:
        dmin_svu = np.ascontiguousarray(dmin_svu, np.float32)
        dmax_svu = np.ascontiguousarray(dmax_svu, np.float32)
    rc = lib().ref_depth2d(raw.ctypes.data_as(C.c_void_p), depth, V, S, U, Cc, C.c_float(scale_factor), C.c_float(dmin),
                           C.c_float(dmax), int(D), C.byref(p), _f(dmin_svu), _f(dmax_svu), int(bool(accept_all)),
                           _f(out["best_depth"]), _f(out["edge_conf"]), _b(out["edge_mask"]), _f(out["disp_conf"]),
                           _f(out["rbar"]), _b(out["valid"]))
    if rc < 0:
        raise ValueError("reference: unsupported channel count")
    return out


def fine_to_coarse(raw, dmin, dmax, D, scale_factor=-1.0, params=None, max_pyr_depth=-1, accept_all_last=True, dims=None):
    """rslf::FineToCoarse<T>: ctor + run + get_results.  dims: the (V_p, U_p) of the levels (to size the per-level
    outputs); when None only the fused maps are returned."""
    raw, depth = _raw(raw)
    V, S, U, Cc = raw.shape
    p = params or default_params()
    out_map = np.zeros((S, V, U), np.float32)
    out_valid = np.zeros((S, V, U), np.uint8)
    L = len(dims) if dims else 0
    levels = []
    for (Vp, Up) in (dims or []):
        levels.append(dict(best_depth=np.zeros((S, Vp, Up), np.float32), edge_conf=np.zeros((S, Vp, Up), np.float32),
                           edge_mask=np.zeros((S, Vp, Up), np.uint8), disp_conf=np.zeros((S, Vp, Up), np.float32),
                           dmin=np.zeros((S, Vp, Up), np.float32), dmax=np.zeros((S, Vp, Up), np.float32)))

    def arr(key, fn, ty):
        return (C.POINTER(ty) * max(L, 1))(*[fn(l[key]) for l in levels]) if L else None
    lvV = (C.c_int * 32)()
    lvU = (C.c_int * 32)()
    secs = C.c_double(0.0)
    n = lib().ref_fine_to_coarse(raw.ctypes.data_as(C.c_void_p), depth, V, S, U, Cc, C.c_float(scale_factor), C.c_float(dmin),
                                 C.c_float(dmax), int(D), C.byref(p), int(max_pyr_depth), int(bool(accept_all_last)),
                                 _f(out_map), _b(out_valid), arr("best_depth", _f, C.c_float), arr("edge_conf", _f, C.c_float),
                                 arr("edge_mask", _b, C.c_uint8), arr("disp_conf", _f, C.c_float), arr("dmin", _f, C.c_float),
                                 arr("dmax", _f, C.c_float), lvV, lvU, C.byref(secs))
    if n < 0:
        raise ValueError("reference: unsupported channel count")
    if dims is not None and [(lvV[i], lvU[i]) for i in range(n)] != list(dims):
        raise ValueError("reference built levels %s, caller expected %s" % ([(lvV[i], lvU[i]) for i in range(n)], dims))
    return dict(map=out_map, valid=out_valid, levels=levels, dims=[(lvV[i], lvU[i]) for i in range(n)], seconds_run=secs.value)


def fine_to_coarse_coloured(raw, dmin, dmax, D, lut_bgr, scale_factor=-1.0, params=None, saturate=True):
    """rslf::FineToCoarse<T>: ctor, run, get_coloured_depth_maps(plots, COLORMAP_JET, saturate) with the colour table
    lut_bgr (256 x 3) standing in for OpenCV's.  Returns [S][V][U][3] uint8 (BGR)."""
    raw, depth = _raw(raw)
    V, S, U, Cc = raw.shape
    p = params or default_params()
    lut = np.ascontiguousarray(lut_bgr, np.uint8).reshape(256, 3)
    out = np.zeros((S, V, U, 3), np.uint8)
    n = lib().ref_fine_to_coarse_coloured(raw.ctypes.data_as(C.c_void_p), depth, V, S, U, Cc, C.c_float(scale_factor),
                                          C.c_float(dmin), C.c_float(dmax), int(D), C.byref(p), _b(lut),
                                          int(bool(saturate)), _b(out))
    if n != S:
        raise ValueError("reference: get_coloured_depth_maps returned %d maps" % n)
    return out


def downsample(raw):
    """rslf::downsample_EPIs on a RAW stack (uint8 stays uint8)."""
    raw, depth = _raw(raw)
    V, S, U, Cc = raw.shape
    V2 = C.c_int(0)
    U2 = C.c_int(0)
    lib().ref_downsample(raw.ctypes.data_as(C.c_void_p), depth, V, S, U, Cc, None, C.byref(V2), C.byref(U2))
    out = np.zeros((V2.value, S, U2.value, Cc), raw.dtype)
    lib().ref_downsample(raw.ctypes.data_as(C.c_void_p), depth, V, S, U, Cc, out.ctypes.data_as(C.c_void_p), None, None)
    return out


def fuse(disp_p, valid_p):
    """rslf::fuse_disp_maps.  disp_p / valid_p: lists (finest first) of [S][V_p][U_p]."""
    L = len(disp_p)
    disp_p = [np.ascontiguousarray(d, np.float32) for d in disp_p]
    valid_p = [np.ascontiguousarray(v, np.uint8) for v in valid_p]
    S = disp_p[0].shape[0]
    Vp = (C.c_int * L)(*[d.shape[1] for d in disp_p])
    Up = (C.c_int * L)(*[d.shape[2] for d in disp_p])
    dp = (C.POINTER(C.c_float) * L)(*[_f(d) for d in disp_p])
    vp = (C.POINTER(C.c_uint8) * L)(*[_b(v) for v in valid_p])
    out_map = np.zeros(disp_p[0].shape, np.float32)
    out_valid = np.zeros(disp_p[0].shape, np.uint8)
    lib().ref_fuse(L, S, Vp, Up, dp, vp, _f(out_map), _b(out_valid))
    return out_map, out_valid


def edge_confidence(epis_norm, s, params=None):
    """rslf::compute_1D_edge_confidence_pile on line s of a normalised float stack."""
    epis = np.ascontiguousarray(epis_norm, np.float32)
    V, S, U, Cc = epis.shape
    p = params or default_params()
    ce = np.zeros((V, U), np.float32)
    mask = np.zeros((V, U), np.uint8)
    lib().ref_edge_confidence(_f(epis), V, S, U, Cc, int(s), C.byref(p), _f(ce), _b(mask))
    return ce, mask


def selective_median(src, mask, epis_norm, s_hat, size=5, eps=0.1):
    """rslf::selective_median_filter."""
    epis = np.ascontiguousarray(epis_norm, np.float32)
    V, S, U, Cc = epis.shape
    src = np.ascontiguousarray(src, np.float32)
    mask = np.ascontiguousarray(mask, np.uint8)
    dst = np.zeros((V, U), np.float32)
    lib().ref_selective_median(_f(src), _b(mask), _f(epis), V, S, U, Cc, int(s_hat), int(size), C.c_float(eps), _f(dst))
    return dst
