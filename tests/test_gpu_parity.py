"""GPU parity: the CUDA path, called through the C ABI (librslf_b200.so), against
the CPU oracle on the same seeded inputs.  Integer / mask / index results must be
identical; float maps are compared exactly as well (the kernels keep the
reference's operation order and rounding), with the north star's 1e-4 relative
tolerance as the documented fallback bound for scores and confidences.
"""
import os

import numpy as np
import pytest

import oracle
from remotesensingproject_b200 import api
from remotesensingproject_b200.synth import make_light_field_np

pytestmark = pytest.mark.gpu

RTOL = 1e-4          # north-star tolerance for scores / confidences


def lf(S, V, U, C, seed, dmin=-1.0, dmax=2.0, **kw):
    epis, _ = make_light_field_np(S, V, U, C, dmin=dmin, dmax=dmax, seed=seed, layers=5, **kw)
    return epis


def assert_maps_equal(gpu, ref, keys, exact=True):
    for k in keys:
        a, b = gpu[k], ref[k]
        assert a.shape == b.shape, k
        if a.dtype == np.uint8 or exact:
            bad = np.flatnonzero(a.ravel() != b.ravel())
            assert bad.size == 0, "%s: %d / %d differ, first at %d: %r vs %r" % (
                k, bad.size, a.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]])
        else:
            np.testing.assert_allclose(a, b, rtol=RTOL, atol=1e-7, err_msg=k)


# --------------------------------------------------------------------------- edge confidence
@pytest.mark.parametrize("C", [1, 3])
@pytest.mark.parametrize("U", [9, 70, 1100])
def test_edge_confidence(gpu_ctx, C, U):
    epis = lf(5, 6, U, C, seed=11 + U)
    p = api.default_params()
    gpu_ctx.upload_epis(epis, epi_scale_factor=1.0)
    for s in (0, 2, 4):
        ce, m = gpu_ctx.edge_confidence(s, p)
        ce_o, m_o = oracle.edge_confidence(oracle.normalise(epis, 1.0), s)
        np.testing.assert_array_equal(m, m_o)
        np.testing.assert_array_equal(ce, ce_o)


def test_edge_confidence_golden(gpu_ctx, golden_dir):
    for name in ("px_c3_s9", "px_c1_s8"):
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        gpu_ctx.upload_epis(g["epi"][None], epi_scale_factor=1.0)
        ce, m = gpu_ctx.edge_confidence(int(g["s_hat"]), api.default_params())
        np.testing.assert_array_equal(m[0], g["mask"])
        np.testing.assert_allclose(ce[0], g["ce"], rtol=1e-6, atol=1e-7)


# --------------------------------------------------------------------------- Depth1DComputer_pile
PILE_CASES = [
    # S, V, U, C, D, forced H (0 = planner's choice), forced register-resident views (-1 = planner's choice)
    (9, 8, 64, 3, 16, 0, -1),
    (8, 8, 64, 1, 40, 0, -1),      # even S
    (7, 6, 50, 3, 100, 0, -1),
    (7, 6, 50, 1, 300, 0, -1),     # several chunks of hypotheses (last-arriver merge)
    (5, 6, 40, 3, 70, 1, 0),       # H=1, three chunks, D not a multiple of 32, all views in shared memory
    (6, 5, 33, 3, 64, 2, 0),
    (7, 5, 33, 3, 130, 4, 0),
    (21, 5, 40, 3, 40, 1, 16),     # 16 views in registers, 5 (+3 padding) in shared memory
    (37, 4, 40, 3, 40, 1, 32),
    (50, 4, 40, 3, 33, 1, 48),
    (19, 4, 40, 3, 70, 2, 16),
    (9, 4, 40, 3, 20, 1, 16),      # fewer views than register slots
    (20, 4, 40, 1, 40, 1, 16),
    (33, 4, 40, 1, 70, 2, 16),
    (51, 3, 40, 1, 40, 1, 48),
]


@pytest.mark.parametrize("S,V,U,C,D,H,RV", PILE_CASES)
def test_depth1d_pile(gpu_ctx, monkeypatch, S, V, U, C, D, H, RV):
    monkeypatch.delenv("RSLF_DEPTH_H", raising=False)
    monkeypatch.delenv("RSLF_DEPTH_RV", raising=False)
    if H:
        monkeypatch.setenv("RSLF_DEPTH_H", str(H))
    if RV >= 0:
        monkeypatch.setenv("RSLF_DEPTH_RV", str(RV))
    epis = lf(S, V, U, C, seed=100 + S + D)
    comp = api.Depth1DComputer_pile(epis, -1.0, 2.0, D, epi_scale_factor=1.0, ctx=gpu_ctx).run()
    ref = oracle.depth1d_pile(oracle.normalise(epis, 1.0), -1.0, 2.0, D)
    gpu = dict(best_depth=comp.m_best_depth_v_u, edge_conf=comp.m_edge_confidence_v_u,
               edge_mask=comp.m_edge_confidence_mask_v_u, disp_conf=comp.m_disp_confidence_v_u,
               rbar=comp.m_rbar_v_u, raw_depth=comp.m_raw_depth_v_u)
    assert ref["computed_pixels"] > 0
    assert gpu_ctx.timing()["computed_pixels"] == ref["computed_pixels"]
    assert_maps_equal(gpu, ref, ["edge_mask", "edge_conf", "raw_depth", "best_depth", "rbar", "disp_conf"])


# register-resident views at either end of the stack (the planner picks the end farther from s_hat), forced both ways
REG_LAST_CASES = [
    # S, C, D, H, RV, reg_last, s_hat
    (21, 3, 40, 1, 16, 0, 2), (21, 3, 40, 1, 16, 1, 2), (21, 3, 40, 1, 16, 1, 18), (21, 3, 40, 1, 16, -1, 0),
    (37, 3, 40, 1, 32, 1, 3), (37, 3, 40, 1, 32, 0, 35), (37, 3, 70, 1, 32, -1, 36), (38, 3, 33, 1, 32, 1, 19),
    (19, 3, 70, 2, 16, 1, 1), (33, 1, 70, 2, 16, 1, 30), (20, 1, 40, 1, 16, 1, 0), (9, 3, 20, 1, 16, 1, 4),
]


@pytest.mark.parametrize("S,C,D,H,RV,RL,s_hat", REG_LAST_CASES)
def test_depth1d_pile_register_views_at_either_end(gpu_ctx, monkeypatch, S, C, D, H, RV, RL, s_hat):
    monkeypatch.setenv("RSLF_DEPTH_H", str(H))
    monkeypatch.setenv("RSLF_DEPTH_RV", str(RV))
    monkeypatch.delenv("RSLF_DEPTH_REG_LAST", raising=False)
    if RL >= 0:
        monkeypatch.setenv("RSLF_DEPTH_REG_LAST", str(RL))
    epis = lf(S, 4, 48, C, seed=300 + S + D + s_hat)
    comp = api.Depth1DComputer_pile(epis, -1.0, 2.0, D, s_hat=s_hat, epi_scale_factor=1.0, ctx=gpu_ctx).run()
    ref = oracle.depth1d_pile(oracle.normalise(epis, 1.0), -1.0, 2.0, D, s_hat=s_hat)
    gpu = dict(best_depth=comp.m_best_depth_v_u, edge_conf=comp.m_edge_confidence_v_u,
               edge_mask=comp.m_edge_confidence_mask_v_u, disp_conf=comp.m_disp_confidence_v_u,
               rbar=comp.m_rbar_v_u, raw_depth=comp.m_raw_depth_v_u)
    assert ref["computed_pixels"] > 0
    assert_maps_equal(gpu, ref, ["edge_mask", "edge_conf", "raw_depth", "best_depth", "rbar", "disp_conf"])


# tensor-memory variant of the depth kernel (k_depth_tm.cuh): 16 views in registers, up to 128 / (4 C) * 4 views in
# TMEM, the rest in shared memory; RSLF_DEPTH_TMEM=2 selects it for every stack it fits
TMEM_CASES = [
    # S, C, D, s_hat (-1 = centre), U
    (24, 3, 40, -1, 48), (37, 3, 70, 3, 48), (37, 3, 33, 35, 48), (60, 3, 40, -1, 40), (60, 3, 100, 0, 40), (60, 3, 64, 59, 40),
    (100, 3, 48, -1, 64), (100, 3, 40, 99, 64), (100, 3, 40, 2, 64), (21, 3, 20, 10, 40), (19, 3, 20, 4, 40),
    (40, 1, 70, -1, 48), (150, 1, 40, 7, 40), (200, 1, 36, -1, 40),
]


@pytest.mark.parametrize("S,C,D,s_hat,U", TMEM_CASES)
def test_depth1d_pile_tensor_memory_variant(gpu_ctx, monkeypatch, S, C, D, s_hat, U):
    monkeypatch.delenv("RSLF_DEPTH_H", raising=False)
    monkeypatch.delenv("RSLF_DEPTH_RV", raising=False)
    monkeypatch.setenv("RSLF_DEPTH_TMEM", "2")
    epis = lf(S, 3, U, C, seed=500 + S + D, dmin=-1.0, dmax=1.5)
    comp = api.Depth1DComputer_pile(epis, -1.0, 1.5, D, s_hat=s_hat, epi_scale_factor=1.0, ctx=gpu_ctx).run()
    ref = oracle.depth1d_pile(oracle.normalise(epis, 1.0), -1.0, 1.5, D, s_hat=s_hat)
    gpu = dict(best_depth=comp.m_best_depth_v_u, edge_conf=comp.m_edge_confidence_v_u,
               edge_mask=comp.m_edge_confidence_mask_v_u, disp_conf=comp.m_disp_confidence_v_u,
               rbar=comp.m_rbar_v_u, raw_depth=comp.m_raw_depth_v_u)
    assert ref["computed_pixels"] > 0
    assert_maps_equal(gpu, ref, ["edge_mask", "edge_conf", "raw_depth", "best_depth", "rbar", "disp_conf"])


def test_fine_to_coarse_tensor_memory_variant(gpu_ctx, monkeypatch):
    """Whole pipeline with the tensor-memory kernel: per-pixel bounds (coarse levels), negative radiances."""
    monkeypatch.setenv("RSLF_DEPTH_TMEM", "2")
    for C, shift in ((3, 0.0), (1, 0.0), (3, -0.15)):
        epis = lf(28, 30, 60, C, seed=640 + C) + np.float32(shift)
        f = api.FineToCoarse(epis, -1.0, 2.0, 40, epi_scale_factor=1.0, ctx=gpu_ctx).run()
        m, v = f.get_results()
        r = oracle.fine_to_coarse(epis, -1.0, 2.0, 40, scale_factor=1.0)
        np.testing.assert_array_equal(v, r["valid"])
        np.testing.assert_array_equal(m, r["map"])


def test_depth1d_pile_uint8_and_max_scale(gpu_ctx):
    epis = lf(7, 6, 48, 3, seed=5)
    u8 = np.clip(np.rint(epis * 255.0), 0, 255).astype(np.uint8)
    comp = api.Depth1DComputer_pile(u8, -1.0, 2.0, 24, ctx=gpu_ctx).run()
    ref = oracle.depth1d_pile(oracle.normalise(u8), -1.0, 2.0, 24)
    np.testing.assert_array_equal(comp.m_best_depth_v_u, ref["best_depth"])
    np.testing.assert_array_equal(comp.m_edge_confidence_mask_v_u, ref["edge_mask"])
    # float input in a SkySat-like range, epi_scale_factor = -1 -> divide by the stack maximum (dc.hpp:442-460)
    sky = (epis * 255.0 + 8.0).astype(np.float32)
    comp = api.Depth1DComputer_pile(sky, -1.0, 2.0, 24, epi_scale_factor=-1.0, ctx=gpu_ctx).run()
    ref = oracle.depth1d_pile(oracle.normalise(sky, -1.0), -1.0, 2.0, 24)
    np.testing.assert_array_equal(comp.m_best_depth_v_u, ref["best_depth"])
    np.testing.assert_array_equal(comp.m_disp_confidence_v_u, ref["disp_conf"])


def test_depth1d_pile_negative_values_and_other_line(gpu_ctx):
    """Negative radiances take the max(r, 0) branch (core.hpp:580); s_hat other than the centre."""
    epis = lf(6, 5, 40, 3, seed=9) - np.float32(0.2)
    comp = api.Depth1DComputer_pile(epis, -1.0, 1.0, 20, s_hat=1, epi_scale_factor=1.0, ctx=gpu_ctx).run()
    ref = oracle.depth1d_pile(oracle.normalise(epis, 1.0), -1.0, 1.0, 20, s_hat=1)
    np.testing.assert_array_equal(comp.m_edge_confidence_mask_v_u, ref["edge_mask"])
    np.testing.assert_array_equal(comp.m_raw_depth_v_u, ref["raw_depth"])
    np.testing.assert_array_equal(comp.m_rbar_v_u, ref["rbar"])
    np.testing.assert_array_equal(comp.m_disp_confidence_v_u, ref["disp_conf"])


def test_pixel_scores_golden(gpu_ctx, golden_dir):
    """Against the cv2 replay of the reference's per-pixel OpenCV sequence: argmax index and depth value."""
    for name in ("px_c3_s9", "px_c1_s8", "px_c3_s5_eq"):
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        D = int(g["D"])
        comp = api.Depth1DComputer_pile(g["epi"][None], float(g["dmin"]), float(g["dmax"]), D,
                                        s_hat=int(g["s_hat"]), epi_scale_factor=1.0, ctx=gpu_ctx).run()
        for i, u in enumerate(g["us"]):
            if not g["mask"][u]:
                continue
            sc = g["score"][i]
            margin = np.sort(sc)[-1] - np.sort(sc)[-2] if D > 1 else 1.0
            if margin > 1e-5 or name.endswith("_eq"):
                assert comp.m_raw_depth_v_u[0, u] == g["dvals"][i][int(g["best"][i])]
            np.testing.assert_allclose(comp.m_rbar_v_u[0, u], g["rbar"][i][int(g["best"][i])], rtol=RTOL, atol=1e-6)
            cd = np.float32(np.float64(g["ce"][u]) * abs(float(g["maxv"][i]) - float(g["mean"][i])))
            np.testing.assert_allclose(comp.m_disp_confidence_v_u[0, u], cd, rtol=RTOL, atol=1e-7)


# --------------------------------------------------------------------------- Depth2DComputer
@pytest.mark.parametrize("S,V,U,C,D", [(7, 9, 48, 3, 24), (6, 9, 48, 1, 24), (5, 4, 130, 3, 40)])
def test_depth2d(gpu_ctx, S, V, U, C, D):
    epis = lf(S, V, U, C, seed=31 + S + C)
    comp = api.Depth2DComputer(epis, -1.0, 2.0, D, epi_scale_factor=1.0, ctx=gpu_ctx).run()
    ref = oracle.depth2d(oracle.normalise(epis, 1.0), -1.0, 2.0, D)
    gpu = dict(best_depth=comp.m_best_depth_s_v_u, edge_conf=comp.m_edge_confidence_s_v_u,
               edge_mask=comp.m_edge_confidence_mask_s_v_u, disp_conf=comp.m_disp_confidence_s_v_u,
               rbar=comp.m_rbar_s_v_u)
    assert gpu_ctx.timing()["computed_pixels"] == ref["computed_pixels"]
    assert_maps_equal(gpu, ref, ["edge_mask", "edge_conf", "best_depth", "disp_conf", "rbar"])
    valid = comp.get_valid_depths_mask_s_v_u()
    np.testing.assert_array_equal(valid, (ref["edge_conf"] > np.float32(0.02)).astype(np.uint8) * 255)


def test_depth2d_per_pixel_bounds(gpu_ctx):
    S, V, U, C, D = 5, 6, 40, 3, 16
    epis = lf(S, V, U, C, seed=77)
    rng = np.random.default_rng(3)
    comp = api.Depth2DComputer(epis, -1.0, 2.0, D, epi_scale_factor=1.0, ctx=gpu_ctx)
    lo = rng.uniform(-1.0, 0.5, (S, V, U)).astype(np.float32)
    hi = (lo + rng.uniform(0.0, 1.5, (S, V, U))).astype(np.float32)
    hi[0, 0, :8] = lo[0, 0, :8]                       # dmin == dmax -> argmax index 0
    comp.edit_dmin()[:] = lo
    comp.edit_dmax()[:] = hi
    comp.run()
    ref = oracle.depth2d(oracle.normalise(epis, 1.0), -1.0, 2.0, D, dmin_svu=lo, dmax_svu=hi)
    np.testing.assert_array_equal(comp.m_best_depth_s_v_u, ref["best_depth"])
    np.testing.assert_array_equal(comp.m_disp_confidence_s_v_u, ref["disp_conf"])
    np.testing.assert_array_equal(comp.m_edge_confidence_mask_s_v_u, ref["edge_mask"])


def test_depth2d_bounds_wider_than_global_range(gpu_ctx):
    """edit_dmin / edit_dmax maps far wider than the constructor's range: the staged scanline segments are
    cut to their rows and the kernel's global-memory fallback has to give the same radiances."""
    S, V, U, C, D = 9, 4, 200, 3, 40
    epis = lf(S, V, U, C, seed=78, dmin=-3.0, dmax=3.0)
    comp = api.Depth2DComputer(epis, 0.0, 0.25, D, epi_scale_factor=1.0, ctx=gpu_ctx)
    comp.edit_dmin()[:] = np.float32(-20.0)
    comp.edit_dmax()[:] = np.float32(20.0)
    comp.run()
    lo = np.full((S, V, U), -20.0, np.float32)
    hi = np.full((S, V, U), 20.0, np.float32)
    ref = oracle.depth2d(oracle.normalise(epis, 1.0), 0.0, 0.25, D, dmin_svu=lo, dmax_svu=hi)
    np.testing.assert_array_equal(comp.m_best_depth_s_v_u, ref["best_depth"])
    np.testing.assert_array_equal(comp.m_disp_confidence_s_v_u, ref["disp_conf"])
    np.testing.assert_array_equal(comp.m_rbar_s_v_u, ref["rbar"])


@pytest.mark.parametrize("C", [3, 1])
def test_depth2d_bounds_tensor_memory_variant(gpu_ctx, monkeypatch, C):
    """Per-pixel bounds (incl. dmin == dmax) and bounds far wider than the constructor's range (segments cut to their
    staging rows -> global-memory fallback) through the tensor-memory kernel."""
    monkeypatch.setenv("RSLF_DEPTH_TMEM", "2")
    S, V, U, D = 24, 3, 120, 40
    epis = lf(S, V, U, C, seed=91 + C, dmin=-2.0, dmax=2.0)
    rng = np.random.default_rng(5)
    lo = rng.uniform(-1.0, 0.5, (S, V, U)).astype(np.float32)
    hi = (lo + rng.uniform(0.0, 1.5, (S, V, U))).astype(np.float32)
    hi[0, 0, :8] = lo[0, 0, :8]
    for (a, b, gmin, gmax) in ((lo, hi, -1.0, 2.0),
                               (np.full((S, V, U), -9.0, np.float32), np.full((S, V, U), 9.0, np.float32), 0.0, 0.25)):
        comp = api.Depth2DComputer(epis, gmin, gmax, D, epi_scale_factor=1.0, ctx=gpu_ctx)
        comp.edit_dmin()[:] = a
        comp.edit_dmax()[:] = b
        comp.run()
        ref = oracle.depth2d(oracle.normalise(epis, 1.0), gmin, gmax, D, dmin_svu=a, dmax_svu=b)
        np.testing.assert_array_equal(comp.m_edge_confidence_mask_s_v_u, ref["edge_mask"])
        np.testing.assert_array_equal(comp.m_best_depth_s_v_u, ref["best_depth"])
        np.testing.assert_array_equal(comp.m_disp_confidence_s_v_u, ref["disp_conf"])
        np.testing.assert_array_equal(comp.m_rbar_s_v_u, ref["rbar"])


# --------------------------------------------------------------------------- FineToCoarse
@pytest.mark.parametrize("S,V,U,C,D,scale", [(5, 44, 60, 3, 16, 1.0), (4, 27, 90, 1, 24, -1.0)])
def test_fine_to_coarse(gpu_ctx, S, V, U, C, D, scale):
    epis = lf(S, V, U, C, seed=55 + C)
    if scale < 0:
        epis = (epis * 255.0 + 8.0).astype(np.float32)      # every level normalised by its own maximum
    ftc = api.FineToCoarse(epis, -1.0, 2.0, D, epi_scale_factor=scale, ctx=gpu_ctx).run()
    out_map, out_valid = ftc.get_results()
    ref = oracle.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=scale)
    levels = ftc.get_levels()
    assert [(l["best_depth"].shape[1], l["best_depth"].shape[2]) for l in levels] == ref["dims"]
    assert len(levels) >= 2
    for p, (lg, lr) in enumerate(zip(levels, ref["levels"])):
        assert_maps_equal(lg, lr, ["edge_mask", "edge_conf", "dmin", "dmax", "best_depth", "disp_conf"])
    np.testing.assert_array_equal(out_valid, ref["valid"])
    np.testing.assert_array_equal(out_map, ref["map"])
    assert gpu_ctx.timing()["samples"] == ref["samples"]


def test_fine_to_coarse_max_depth_and_no_accept_all(gpu_ctx):
    epis = lf(4, 30, 44, 3, seed=8)
    ftc = api.FineToCoarse(epis, -1.0, 2.0, 12, epi_scale_factor=1.0, max_pyr_depth=2, accept_all_last_scale=False,
                           ctx=gpu_ctx).run()
    out_map, out_valid = ftc.get_results()
    ref = oracle.fine_to_coarse(epis, -1.0, 2.0, 12, scale_factor=1.0, max_pyr_depth=2, accept_all_last=False)
    assert len(ftc.get_levels()) == 2
    np.testing.assert_array_equal(out_valid, ref["valid"])
    np.testing.assert_array_equal(out_map, ref["map"])


# --------------------------------------------------------------------------- free functions
@pytest.mark.parametrize("C,size", [(3, 5), (1, 5), (3, 3), (1, 7), (3, 1)])
def test_selective_median(gpu_ctx, C, size):
    S, V, U = 3, 17, 45
    epis = lf(S, V, U, C, seed=21)
    rng = np.random.default_rng(C * 10 + size)
    src = rng.uniform(-1, 4, (V, U)).astype(np.float32)
    src[rng.random((V, U)) < 0.3] = np.float32(1.5)          # ties
    mask = (rng.random((V, U)) < 0.7).astype(np.uint8) * 255
    gpu_ctx.upload_epis(epis, epi_scale_factor=1.0)
    out = gpu_ctx.selective_median(src, mask, 1, size=size, eps=0.1)
    ref = oracle.selective_median(src, mask, oracle.normalise(epis, 1.0), 1, size=size, eps=0.1)
    np.testing.assert_array_equal(out, ref)


def test_downsample_golden_and_oracle(gpu_ctx, golden_dir):
    for name in ("down_c3_odd", "down_c1_even", "down_c1_135"):
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        out = gpu_ctx.downsample_epis(g["raw"])
        np.testing.assert_array_equal(out, oracle.downsample(g["raw"]))
        np.testing.assert_allclose(out, g["out"], rtol=5e-7, atol=0)      # cv2.GaussianBlur + cv2.resize
    rng = np.random.default_rng(1)
    raw = rng.random((75, 2, 131, 3), dtype=np.float32)                   # several tiles, odd sizes
    np.testing.assert_array_equal(gpu_ctx.downsample_epis(raw), oracle.downsample(raw))


def test_downsample_uint8_golden_and_oracle(gpu_ctx, golden_dir):
    """8-bit stacks: OpenCV's integer blur / resize paths, bit-exact against the cv2 fixtures."""
    for name in ("down_u8_c3_odd", "down_u8_c1_135", "down_u8_c1_even"):
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        out = gpu_ctx.downsample_epis(g["raw"])
        assert out.dtype == np.uint8
        np.testing.assert_array_equal(out, g["out"])
        np.testing.assert_array_equal(out, oracle.downsample(g["raw"]))
    rng = np.random.default_rng(2)
    raw = rng.integers(0, 256, (75, 2, 131, 3), dtype=np.uint8)
    np.testing.assert_array_equal(gpu_ctx.downsample_epis(raw), oracle.downsample(raw))


def test_fine_to_coarse_uint8_pyramid(gpu_ctx):
    """CV_8U input through the whole pyramid (levels stay 8-bit, each normalised by 1/255)."""
    epis = lf(5, 46, 70, 3, seed=61)
    u8 = np.clip(np.rint(epis * 255.0), 0, 255).astype(np.uint8)
    ftc = api.FineToCoarse(u8, -1.0, 2.0, 16, ctx=gpu_ctx).run()
    out_map, out_valid = ftc.get_results()
    ref = oracle.fine_to_coarse(u8, -1.0, 2.0, 16)
    assert len(ftc.get_levels()) == len(ref["dims"]) >= 3
    for lg, lr in zip(ftc.get_levels(), ref["levels"]):
        assert_maps_equal(lg, lr, ["edge_mask", "edge_conf", "dmin", "dmax", "best_depth", "disp_conf"])
    np.testing.assert_array_equal(out_valid, ref["valid"])
    np.testing.assert_array_equal(out_map, ref["map"])


def test_set_bounds(gpu_ctx):
    rng = np.random.default_rng(12)
    S, Vu, Uu = 3, 23, 77
    depth = rng.uniform(-1, 4, (S, Vu, Uu)).astype(np.float32)
    valid = (rng.random((S, Vu, Uu)) < 0.15).astype(np.uint8) * 255
    valid[0, 0] = 0                       # a row without any valid pixel
    valid[0, 1] = 0
    valid[0, 1, 0] = 255                  # only index 0 valid: never tested on the left (ftc.hpp:222-236)
    valid[1, 2] = 255
    Vd, Ud = oracle.half_size(Vu), oracle.half_size(Uu)
    a = gpu_ctx.set_bounds(depth, valid, Vd, Ud, -1.0, 4.0)
    b = oracle.set_bounds(depth, valid, Vd, Ud, -1.0, 4.0)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1], b[1])


def test_fuse_golden_and_oracle(gpu_ctx, golden_dir):
    g = np.load(os.path.join(golden_dir, "fuse_3lvl.npz"))
    disp, valid = [g["d0"], g["d1"], g["d2"]], [g["v0"], g["v1"], g["v2"]]
    m, k = gpu_ctx.fuse_disp_maps(disp, valid)
    mo, ko = oracle.fuse(disp, valid)
    np.testing.assert_array_equal(k, ko)
    np.testing.assert_array_equal(m, mo)
    np.testing.assert_array_equal(k, g["out_valid"])                      # cv2.resize / medianBlur replay
    np.testing.assert_allclose(m, g["out_map"], rtol=0, atol=5e-6)


# --------------------------------------------------------------------------- behaviour at the boundary
def test_errors_are_loud():
    ctx = api.Context(0)
    with pytest.raises(api.RslfError):
        ctx.depth1d_pile_run(-1.0, 1.0, 8, -1, api.default_params())      # nothing uploaded
    ctx.upload_epis(np.zeros((3, 3, 12, 3), np.float32), 1.0)
    with pytest.raises(api.RslfError):
        ctx.depth1d_pile_run(-1.0, 1.0, 1, -1, api.default_params())      # dim_d < 2
    with pytest.raises(api.RslfError):
        ctx.fine_to_coarse_get()                                          # no run yet
    ctx.close()


def test_all_dark_and_flat_inputs(gpu_ctx):
    """No confident pixel at all (flat / dark light field): empty work lists must be handled."""
    epis = np.full((6, 5, 40, 3), 0.01, np.float32)
    comp = api.Depth2DComputer(epis, -1.0, 2.0, 8, epi_scale_factor=1.0, ctx=gpu_ctx).run()
    assert not comp.m_edge_confidence_mask_s_v_u.any()
    assert not comp.m_best_depth_s_v_u.any()
    assert gpu_ctx.timing()["computed_pixels"] == 0


def test_device_resident_input_matches_host_upload(gpu_ctx):
    import torch
    epis = lf(5, 6, 40, 3, seed=2)
    a = api.Depth2DComputer(epis, -1.0, 2.0, 12, epi_scale_factor=1.0, ctx=gpu_ctx).run().m_best_depth_s_v_u.copy()
    t = torch.from_numpy(epis).cuda()
    b = api.Depth2DComputer(t, -1.0, 2.0, 12, epi_scale_factor=1.0, ctx=gpu_ctx).run().m_best_depth_s_v_u
    np.testing.assert_array_equal(a, b)


def test_upload_images_builds_epis(gpu_ctx):
    """rslf::build_epis_from_imgs (rslf_io.cpp:194-227) on the device."""
    epis = lf(4, 5, 36, 3, seed=4)                       # [V][S][U][C]
    imgs = np.ascontiguousarray(epis.transpose(1, 0, 2, 3))
    gpu_ctx.upload_epis(epis, 1.0)
    p = api.default_params()
    ce_a, m_a = gpu_ctx.edge_confidence(2, p)
    gpu_ctx.upload_images(imgs, 1.0)
    ce_b, m_b = gpu_ctx.edge_confidence(2, p)
    np.testing.assert_array_equal(ce_a, ce_b)
    np.testing.assert_array_equal(m_a, m_b)


def test_row_work_profile(gpu_ctx):
    """rslf_cuda_get_row_work: per-row counts of evaluated pixels, all passes and levels, sum = computed pixels."""
    epis = lf(5, 44, 60, 3, seed=58)
    api.FineToCoarse(epis, -1.0, 2.0, 16, epi_scale_factor=1.0, ctx=gpu_ctx).run()
    w = gpu_ctx.row_work()
    assert w.shape == (44,) and int(w.sum()) == int(gpu_ctx.timing()["computed_pixels"])
    ref = oracle.fine_to_coarse(epis, -1.0, 2.0, 16, scale_factor=1.0)
    assert int(w.sum()) == int(ref["computed_pixels"])


# --------------------------------------------------------------------------- against the reference's own code
def _ref_cases():
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_ref_golden.py")
    spec = importlib.util.spec_from_file_location("make_ref_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", ["ref_pile_c3", "ref_pile_c1_u8", "ref_2d_c3", "ref_ftc_c1", "ref_ftc_c3_u8"])
def test_against_reference_fixtures(gpu_ctx, golden_dir, name):
    """tests/golden/ref_*.npz were written by the reference's own sources (compiled against oracle/cvshim):
    every map of the CUDA path must be bit-identical to them."""
    mod = _ref_cases()
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    out = mod.run_case(name, g["epis"], side="gpu", ctx=gpu_ctx)
    for k in g.files:
        if k != "epis":
            np.testing.assert_array_equal(out[k], g[k], err_msg="%s:%s" % (name, k))


@pytest.mark.parametrize("S,V,U,C,D,as_u8", [(12, 48, 96, 3, 64, False), (9, 52, 110, 1, 48, True), (16, 40, 72, 3, 40, True)])
def test_against_reference_build(gpu_ctx, S, V, U, C, D, as_u8):
    """The reference build itself (oracle/_ref/librslf_ref.so travels with the snapshot): larger fine-to-coarse runs than
    the fixtures hold, all levels and the fused result bit-identical."""
    from oracle import ref
    if not os.path.exists(ref.LIB_PATH):
        pytest.skip("oracle/_ref/librslf_ref.so was not shipped")
    epis = lf(S, V, U, C, seed=700 + V)
    if as_u8:
        epis = np.clip(np.rint(epis * 255.0), 0, 255).astype(np.uint8)
    dims = oracle.pyramid_dims(V, U)
    r = ref.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=-1.0 if as_u8 else 1.0, dims=dims)
    f = api.FineToCoarse(epis, -1.0, 2.0, D, epi_scale_factor=-1.0 if as_u8 else 1.0, ctx=gpu_ctx).run()
    m, v = f.get_results()
    np.testing.assert_array_equal(v, r["valid"])
    np.testing.assert_array_equal(m, r["map"])
    levels = f.get_levels()
    assert len(levels) == len(dims)
    for p, lv in enumerate(levels):
        for k in ("edge_mask", "best_depth", "disp_conf"):
            np.testing.assert_array_equal(lv[k], r["levels"][p][k], err_msg="level %d %s" % (p, k))


def test_row_sharded_run_equals_single_gpu():
    """Two ranks on two GPUs (torchrun + NCCL): every output block is bit-identical to the one-GPU result.
    Skipped on boxes with a single GPU (the driver's -m gpu run); tools/multigpu_check.py is the same check."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tools", "multigpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multi-GPU parity OK" in r.stdout
