"""GPU parity of the rows SURVEY.md 8(f) marks "next" (dormant options of the reference's parameter block and
input formats either side of the hot path), through the C ABI against the CPU oracle.  Same bar as
tests/test_gpu_parity.py: every map identical.
"""
import os

import numpy as np
import pytest

import oracle
from remotesensingproject_b200 import api
from remotesensingproject_b200.synth import make_light_field_np

pytestmark = pytest.mark.gpu

MAPS = ["edge_mask", "edge_conf", "best_depth", "rbar", "disp_conf"]


def lf(S, V, U, C, seed, dmin=-1.0, dmax=2.0, **kw):
    epis, _ = make_light_field_np(S, V, U, C, dmin=dmin, dmax=dmax, seed=seed, layers=5, **kw)
    return epis


def same(a, b, tag):
    assert a.shape == b.shape, tag
    bad = np.flatnonzero(a.ravel() != b.ravel())
    assert bad.size == 0, "%s: %d / %d differ, first at %d: %r vs %r" % (
        tag, bad.size, a.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]])


# --------------------------------------------------------------------------- morphological opening of the edge mask
OPENINGS = [(2, 3), (0, 2), (1, 5), (2, 4), (2, 7), (0, 3)]      # (cv::MorphShapes, size): core.hpp:759-769


@pytest.mark.parametrize("shape,size", OPENINGS)
@pytest.mark.parametrize("C", [1, 3])
def test_edge_mask_opening_plane(gpu_ctx, shape, size, C):
    epis = lf(5, 19, 150, C, seed=60 + size)
    kw = dict(edge_confidence_opening_type=shape, edge_confidence_opening_size=size, edge_score_threshold=0.2)  # ragged masks
    p, po = api.default_params(**kw), oracle.default_params(**kw)
    gpu_ctx.upload_epis(epis, epi_scale_factor=1.0)
    for s in (0, 3):
        ce, m = gpu_ctx.edge_confidence(s, p)
        ce_o, m_o = oracle.edge_confidence(oracle.normalise(epis, 1.0), s, po)
        same(m, m_o, "opened mask")
        same(ce, ce_o, "C_e")
        _, m_plain = oracle.edge_confidence(oracle.normalise(epis, 1.0), s, oracle.default_params(edge_score_threshold=0.2))
        assert (m_o != m_plain).any() and m_o.any()


def test_edge_mask_opening_thin_images(gpu_ctx):
    """Images thinner than the element: positions outside the image take no part (morphologyDefaultBorderValue)."""
    for (V, U) in ((1, 40), (2, 33), (12, 9)):
        epis = lf(3, V, U, 3, seed=V + U)
        for shape, size in ((0, 5), (2, 5), (1, 4)):
            p = api.default_params(edge_confidence_opening_type=shape, edge_confidence_opening_size=size)
            po = oracle.default_params(edge_confidence_opening_type=shape, edge_confidence_opening_size=size)
            gpu_ctx.upload_epis(epis, epi_scale_factor=1.0)
            _, m = gpu_ctx.edge_confidence(1, p)
            _, m_o = oracle.edge_confidence(oracle.normalise(epis, 1.0), 1, po)
            same(m, m_o, "opened mask %dx%d shape %d size %d" % (V, U, shape, size))


@pytest.mark.parametrize("shape,size", [(2, 3), (0, 2), (1, 5)])
def test_opening_through_the_computers(gpu_ctx, shape, size):
    """The opened masks drive the pile, every pass of the 2D computer (incl. the dark-row bookkeeping of the
    propagation, counted on the opened mask) and every level of the pyramid."""
    kw = dict(edge_confidence_opening_type=shape, edge_confidence_opening_size=size)
    p, po = api.default_params(**kw), oracle.default_params(**kw)
    epis = lf(7, 14, 60, 3, seed=40 + size, dark_fraction=0.15)
    comp = api.Depth1DComputer_pile(epis, -1.0, 2.0, 24, epi_scale_factor=1.0, parameters=p, ctx=gpu_ctx).run()
    ref = oracle.depth1d_pile(oracle.normalise(epis, 1.0), -1.0, 2.0, 24, params=po)
    gpu = dict(best_depth=comp.m_best_depth_v_u, edge_conf=comp.m_edge_confidence_v_u,
               edge_mask=comp.m_edge_confidence_mask_v_u, disp_conf=comp.m_disp_confidence_v_u, rbar=comp.m_rbar_v_u)
    for k in MAPS:
        same(gpu[k], ref[k], "pile " + k)
    c2 = api.Depth2DComputer(epis, -1.0, 2.0, 24, epi_scale_factor=1.0, parameters=p, ctx=gpu_ctx).run()
    r2 = oracle.depth2d(oracle.normalise(epis, 1.0), -1.0, 2.0, 24, params=po)
    g2 = dict(best_depth=c2.m_best_depth_s_v_u, edge_conf=c2.m_edge_confidence_s_v_u,
              edge_mask=c2.m_edge_confidence_mask_s_v_u, disp_conf=c2.m_disp_confidence_s_v_u, rbar=c2.m_rbar_s_v_u)
    assert gpu_ctx.timing()["computed_pixels"] == r2["computed_pixels"]
    for k in MAPS:
        same(g2[k], r2[k], "2d " + k)
    # confident-and-dark pixels (no shadow cut, wide propagation epsilon): sources without r_bar paint them, and the
    # rows that still hold such targets are tracked from the OPENED mask
    kd = dict(kw, cut_shadows=0, propagation_epsilon=0.25)
    c2 = api.Depth2DComputer(epis, -1.0, 2.0, 24, epi_scale_factor=1.0, parameters=api.default_params(**kd), ctx=gpu_ctx).run()
    r2 = oracle.depth2d(oracle.normalise(epis, 1.0), -1.0, 2.0, 24, params=oracle.default_params(**kd))
    g2 = dict(best_depth=c2.m_best_depth_s_v_u, edge_conf=c2.m_edge_confidence_s_v_u,
              edge_mask=c2.m_edge_confidence_mask_s_v_u, disp_conf=c2.m_disp_confidence_s_v_u, rbar=c2.m_rbar_s_v_u)
    assert gpu_ctx.timing()["computed_pixels"] == r2["computed_pixels"]
    for k in MAPS:
        same(g2[k], r2[k], "2d dark " + k)
    epis = lf(5, 24, 64, 1, seed=41 + size)
    f = api.FineToCoarse(epis, -1.0, 2.0, 16, epi_scale_factor=1.0, parameters=p, ctx=gpu_ctx).run()
    m, v = f.get_results()
    r3 = oracle.fine_to_coarse(epis, -1.0, 2.0, 16, scale_factor=1.0, params=po)
    same(v, r3["valid"], "ftc valid")
    same(m, r3["map"], "ftc map")


def test_opening_rejects_what_it_cannot_do(gpu_ctx):
    gpu_ctx.upload_epis(lf(3, 4, 40, 3, seed=1), epi_scale_factor=1.0)
    with pytest.raises(api.RslfError):
        gpu_ctx.edge_confidence(1, api.default_params(edge_confidence_opening_size=33))     # element larger than 31
    with pytest.raises(api.RslfError):
        gpu_ctx.edge_confidence(1, api.default_params(edge_confidence_opening_type=7, edge_confidence_opening_size=3))


# --------------------------------------------------------------------------- CV_16U input
def u16_stack(S, V, U, C, seed, peak):
    return np.clip(np.rint(lf(S, V, U, C, seed=seed) * peak), 0, 65535).astype(np.uint16)


@pytest.mark.parametrize("V,U,C", [(13, 27, 3), (20, 32, 1), (9, 135, 1), (40, 70, 3), (2, 3, 3)])
def test_downsample_uint16(gpu_ctx, V, U, C):
    """downsample_EPIs on CV_16U stacks (integer Gaussian, (a+b+c+d+2) >> 2), incl. exact ties of the blur."""
    rng = np.random.default_rng(V * 1000 + U)
    raw = rng.integers(0, 65536, size=(V, 3, U, C)).astype(np.uint16)
    got = gpu_ctx.downsample_epis(raw)
    want = oracle.downsample(raw)
    assert got.dtype == np.uint16
    same(got, want, "u16 downsample")
    ties = (raw // 128 * 128).astype(np.uint16)
    same(gpu_ctx.downsample_epis(ties), oracle.downsample(ties), "u16 downsample, tie values")
    top = np.full((V, 2, U, C), 65535, np.uint16)                       # accumulators at their maximum
    same(gpu_ctx.downsample_epis(top), oracle.downsample(top), "u16 downsample, saturated")


@pytest.mark.parametrize("C,scale", [(3, -1.0), (1, -1.0), (3, 4095.0)])
def test_uint16_input_through_the_computers(gpu_ctx, C, scale):
    """dc.hpp:442-477: x float(1 / scale), scale = stack maximum (scale < 0) or the given factor."""
    raw = u16_stack(7, 10, 56, C, seed=70 + C, peak=4095.0 if scale > 0 else 60000.0)
    norm = oracle.normalise(raw, scale)
    comp = api.Depth1DComputer_pile(raw, -1.0, 2.0, 24, epi_scale_factor=scale, ctx=gpu_ctx).run()
    ref = oracle.depth1d_pile(norm, -1.0, 2.0, 24)
    gpu = dict(best_depth=comp.m_best_depth_v_u, edge_conf=comp.m_edge_confidence_v_u,
               edge_mask=comp.m_edge_confidence_mask_v_u, disp_conf=comp.m_disp_confidence_v_u, rbar=comp.m_rbar_v_u)
    assert ref["computed_pixels"] > 0
    for k in MAPS:
        same(gpu[k], ref[k], "u16 pile " + k)
    c2 = api.Depth2DComputer(raw, -1.0, 2.0, 24, epi_scale_factor=scale, ctx=gpu_ctx).run()
    r2 = oracle.depth2d(norm, -1.0, 2.0, 24)
    g2 = dict(best_depth=c2.m_best_depth_s_v_u, edge_conf=c2.m_edge_confidence_s_v_u,
              edge_mask=c2.m_edge_confidence_mask_s_v_u, disp_conf=c2.m_disp_confidence_s_v_u, rbar=c2.m_rbar_s_v_u)
    for k in MAPS:
        same(g2[k], r2[k], "u16 2d " + k)


@pytest.mark.parametrize("S,V,U,C,scale", [(5, 44, 70, 3, -1.0), (4, 27, 135, 1, -1.0), (5, 24, 50, 3, 4095.0)])
def test_uint16_fine_to_coarse(gpu_ctx, S, V, U, C, scale):
    """The pyramid stays 16-bit between levels (ftc.hpp:142-147); every level is normalised from its own raw stack."""
    raw = u16_stack(S, V, U, C, seed=V, peak=4095.0 if scale > 0 else 60000.0)
    f = api.FineToCoarse(raw, -1.0, 2.0, 16, epi_scale_factor=scale, ctx=gpu_ctx).run()
    m, v = f.get_results()
    r = oracle.fine_to_coarse(raw, -1.0, 2.0, 16, scale_factor=scale)
    assert len(r["dims"]) >= 2 and r["computed_pixels"] > 0
    assert gpu_ctx.timing()["computed_pixels"] == r["computed_pixels"]
    for p, (g, o) in enumerate(zip(f.get_levels(), r["levels"])):
        for k in ("edge_mask", "edge_conf", "dmin", "dmax", "best_depth", "disp_conf"):
            same(g[k], o[k], "u16 level %d %s" % (p, k))
    same(v, r["valid"], "u16 ftc valid")
    same(m, r["map"], "u16 ftc map")


def test_uint16_images_upload(gpu_ctx):
    """rslf_cuda_upload_images (build_epis_from_imgs, rslf_io.cpp:194-227) with 16-bit frames."""
    raw = u16_stack(5, 8, 40, 3, seed=3, peak=50000.0)
    imgs = np.ascontiguousarray(raw.transpose(1, 0, 2, 3))               # [S][V][U][C]
    gpu_ctx.upload_images(imgs, epi_scale_factor=-1.0)
    ce, m = gpu_ctx.edge_confidence(2, api.default_params())
    ce_o, m_o = oracle.edge_confidence(oracle.normalise(raw, -1.0), 2)
    same(m, m_o, "u16 images mask")
    same(ce, ce_o, "u16 images C_e")


# --------------------------------------------------------------------------- coloured disparity maps
@pytest.mark.parametrize("S,V,U,C,kw", [(6, 24, 64, 3, {}), (5, 45, 90, 1, {}), (7, 24, 64, 3, dict(cut_shadows=0)),
                                        (2, 24, 40, 1, {}), (5, 130, 300, 3, {})])
def test_coloured_depth_maps(gpu_ctx, golden_dir, S, V, U, C, kw):
    """FineToCoarse::get_coloured_depth_maps (ftc.hpp:324-377): quantile fit by radix select, uchar scaling, colour
    table, invalid / shadow pixels black."""
    import os
    lut = np.load(os.path.join(golden_dir, "colormap_jet.npy"))
    epis = lf(S, V, U, C, seed=50 + V, dark_fraction=0.2)
    p, po = api.default_params(**kw), oracle.default_params(**kw)
    f = api.FineToCoarse(epis, -1.0, 2.0, 16, epi_scale_factor=1.0, parameters=p, ctx=gpu_ctx).run()
    got, fit = gpu_ctx.fine_to_coarse_coloured(lut, p, saturate=True)
    o = oracle.fine_to_coarse(epis, -1.0, 2.0, 16, scale_factor=1.0, params=po)
    want, fit_o = oracle.colour_maps(o["map"], o["valid"], oracle.normalise(epis, 1.0), lut, cut_shadows=po.cut_shadows)
    assert fit == fit_o, (fit, fit_o)
    same(got, want, "coloured maps")
    same(f.get_coloured_depth_maps(lut_bgr=lut), want, "coloured maps (mirror class)")
    m, v = f.get_results()                                   # the fused results are untouched by the colouring
    same(m, o["map"], "ftc map after colouring")
    same(v, o["valid"], "ftc valid after colouring")


def test_coloured_depth_maps_negative_disparities(gpu_ctx, golden_dir):
    """Order statistics over negative values (the integer keys of the radix select reverse their order)."""
    import os
    lut = np.load(os.path.join(golden_dir, "colormap_jet.npy"))
    epis = lf(4, 24, 48, 3, seed=9, dmin=-3.0, dmax=-1.0)
    f = api.FineToCoarse(epis, -3.0, -1.0, 12, epi_scale_factor=1.0, ctx=gpu_ctx).run()
    got, fit = gpu_ctx.fine_to_coarse_coloured(lut, api.default_params())
    o = oracle.fine_to_coarse(epis, -3.0, -1.0, 12, scale_factor=1.0)
    want, fit_o = oracle.colour_maps(o["map"], o["valid"], oracle.normalise(epis, 1.0), lut)
    assert fit == fit_o and fit[0] < 0
    same(got, want, "coloured maps, negative disparities")


def test_coloured_depth_maps_rejects(gpu_ctx, golden_dir):
    import os
    lut = np.load(os.path.join(golden_dir, "colormap_jet.npy"))
    epis = lf(4, 24, 48, 3, seed=9)
    api.FineToCoarse(epis, -1.0, 2.0, 12, epi_scale_factor=1.0, ctx=gpu_ctx).run()
    with pytest.raises(api.RslfError):
        gpu_ctx.fine_to_coarse_coloured(lut, api.default_params(), saturate=False)       # mean + 12 std fit: not implemented
    api.Depth2DComputer(epis, -1.0, 2.0, 12, epi_scale_factor=1.0, ctx=gpu_ctx).run()
    with pytest.raises(api.RslfError):
        gpu_ctx.fine_to_coarse_coloured(lut, api.default_params())                       # no fine-to-coarse result


# --------------------------------------------------------------------------- opt-in contracted arithmetic
def _pile_maps(comp):
    return dict(best_depth=comp.m_best_depth_v_u, edge_conf=comp.m_edge_confidence_v_u,
                edge_mask=comp.m_edge_confidence_mask_v_u, disp_conf=comp.m_disp_confidence_v_u,
                rbar=comp.m_rbar_v_u, raw_depth=comp.m_raw_depth_v_u)


@pytest.mark.parametrize("S,C,D,s_hat,U", [(24, 3, 40, -1, 96), (60, 3, 100, 0, 64), (100, 3, 128, -1, 64), (37, 3, 33, 35, 48)])
def test_fast_math_stays_within_the_specified_tolerance(gpu_ctx, monkeypatch, S, C, D, s_hat, U):
    """rslf_cuda_set_fast_math: fused multiply-adds in the mean shift of the tensor-memory kernel.  Not bit-identical,
    but inside the tolerance the port is specified to: same masks, the same disparity index wherever the reference's
    winning score margin exceeds 1e-5, 1e-4 relative on the confidences."""
    monkeypatch.delenv("RSLF_DEPTH_H", raising=False)
    monkeypatch.delenv("RSLF_DEPTH_RV", raising=False)
    monkeypatch.delenv("RSLF_DEPTH_TMEM", raising=False)
    import ctypes
    fits = (ctypes.c_int * 8)()
    api.lib().rslf_plan_depth_tm(S, C, D, s_hat if s_hat >= 0 else S // 2, ctypes.c_float(-1.0), ctypes.c_float(1.5),
                                 ctypes.c_float(1.0), ctypes.c_size_t(227 * 1024 - 64), fits, None, None)
    assert fits[0] == 1, "the case must run on the tensor-memory kernel (the only one with a contracted variant)"
    epis = lf(S, 6, U, C, seed=900 + S + D, dmin=-1.0, dmax=1.5)
    ref = oracle.depth1d_pile(oracle.normalise(epis, 1.0), -1.0, 1.5, D, s_hat=s_hat)
    exact = _pile_maps(api.Depth1DComputer_pile(epis, -1.0, 1.5, D, s_hat=s_hat, epi_scale_factor=1.0, ctx=gpu_ctx).run())
    gpu_ctx.set_fast_math(True)
    try:
        fast = _pile_maps(api.Depth1DComputer_pile(epis, -1.0, 1.5, D, s_hat=s_hat, epi_scale_factor=1.0, ctx=gpu_ctx).run())
    finally:
        gpu_ctx.set_fast_math(False)
    for k in ("edge_mask", "edge_conf", "raw_depth", "disp_conf", "rbar"):
        same(exact[k], ref[k], "exact mode " + k)                       # the default path is untouched by the switch
    same(fast["edge_mask"], ref["edge_mask"], "fast mode mask")
    same(fast["edge_conf"], ref["edge_conf"], "fast mode C_e")
    computed = ref["best_idx"] >= 0
    safe = computed & (ref["margin"] > 1e-5)
    assert safe.sum() > 0.5 * computed.sum() > 0
    same(fast["raw_depth"][safe], ref["raw_depth"][safe], "fast mode disparity where the margin exceeds 1e-5")
    agree = computed & (fast["raw_depth"] == ref["raw_depth"])
    np.testing.assert_allclose(fast["disp_conf"][agree], ref["disp_conf"][agree], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(fast["rbar"][agree], ref["rbar"][agree], rtol=1e-4, atol=1e-6)
    assert (fast["disp_conf"] != ref["disp_conf"]).any(), "the contracted kernel did not run"


# The whole pipeline in contracted mode (propagation, pyramid, fusion) is checked by tests/test_gpu_tolerance.py: the
# oracle replays the recorded decisions and every map must then be identical.


# --------------------------------------------------------------------------- Depth1DComputer (one EPI)
@pytest.mark.parametrize("S,U,C,D,s_hat,kw", [(9, 64, 3, 16, -1, {}), (8, 64, 1, 40, -1, {}), (7, 50, 3, 100, 2, {}),
                                                (12, 33, 1, 33, 5, dict(edge_confidence_opening_size=3)),
                                                (24, 96, 3, 40, -1, {})])
def test_single_epi_computer(gpu_ctx, S, U, C, D, s_hat, kw):
    """rslf::Depth1DComputer<T> (dc.hpp:254-363): one EPI, one line, no median; the opening parameter is ignored."""
    epi = lf(S, 3, U, C, seed=100 + S + D)[1]
    comp = api.Depth1DComputer(epi, -1.0, 2.0, D, s_hat=s_hat, epi_scale_factor=1.0, parameters=api.default_params(**kw),
                               ctx=gpu_ctx).run()
    o = oracle.depth1d(oracle.normalise(epi[None], 1.0)[0], -1.0, 2.0, D, s_hat=s_hat, params=oracle.default_params(**kw))
    assert o["computed_pixels"] > 0
    same(comp.m_edge_confidence_mask_u, o["edge_mask"], "1d mask")
    same(comp.m_edge_confidence_u, o["edge_conf"], "1d C_e")
    same(comp.m_best_depth_u, o["best_depth"], "1d depth")
    same(comp.m_disp_confidence_u, o["disp_conf"], "1d C_d")
    same(comp.m_rbar_u, o["rbar"], "1d rbar")


# --------------------------------------------------------------------------- disparity-confidence criterion
@pytest.mark.parametrize("S,V,U,C,D,thr", [(7, 24, 64, 3, 24, 0.01), (6, 44, 70, 1, 32, 0.01), (24, 23, 96, 3, 40, 0.05), (5, 23, 50, 3, 40, 0.0)])
def test_disp_confidence_criterion(gpu_ctx, S, V, U, C, D, thr):
    """rslf_cuda_set_confidence_criterion(1): the reference as intended with -D_USE_DISP_CONFIDENCE_SCORE (propagation and
    validity gated by C_d > par_disp_score_threshold).  Checked against the oracle's restatement, which
    tests/test_oracle_vs_reference.py pins to the reference built that way, and against that build itself when shipped."""
    from oracle import ref
    epis, _ = make_light_field_np(S, V, U, C, dmin=-1.0, dmax=2.0, seed=500 + S + V, layers=5)
    p = api.default_params(disp_score_threshold=thr)
    po = oracle.default_params(disp_score_threshold=thr)
    gpu_ctx.set_confidence_criterion("disp")
    oracle.set_criterion("disp")
    try:
        f = api.FineToCoarse(epis, -1.0, 2.0, D, epi_scale_factor=1.0, parameters=p, ctx=gpu_ctx).run()
        out_map, out_valid = f.get_results()
        levels = f.get_levels()
        o = oracle.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=1.0, params=po)
        np.testing.assert_array_equal(out_valid, o["valid"])
        np.testing.assert_array_equal(out_map, o["map"])
        for g, l in zip(levels, o["levels"]):
            for k in ("edge_mask", "best_depth", "dmin", "dmax"):
                np.testing.assert_array_equal(g[k], l[k], err_msg=k)
        assert gpu_ctx.timing()["computed_pixels"] == o["computed_pixels"]
        c = api.Depth2DComputer(epis, -1.0, 2.0, D, epi_scale_factor=1.0, parameters=p, ctx=gpu_ctx).run()
        o2 = oracle.depth2d(oracle.normalise(epis, 1.0), -1.0, 2.0, D, params=po)
        np.testing.assert_array_equal(c.m_best_depth_s_v_u, o2["best_depth"])
        np.testing.assert_array_equal(c.get_valid_depths_mask_s_v_u(), (o2["disp_conf"] > np.float32(thr)).astype(np.uint8) * 255)
        if os.path.exists(ref.LIB_CD_PATH):
            ref.set_criterion("disp")
            r = ref.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=1.0, params=po)
            np.testing.assert_array_equal(out_map, r["map"])
            np.testing.assert_array_equal(out_valid, r["valid"])
    finally:
        gpu_ctx.set_confidence_criterion("edge")
        oracle.set_criterion("edge")
        ref.set_criterion("edge")
    with pytest.raises(api.RslfError):
        gpu_ctx.set_confidence_criterion("line")


# --------------------------------------------------------------------------- pipelined ingest
@pytest.mark.parametrize("dtype,scale", [(np.float32, 1.0), (np.uint8, -1.0), (np.uint16, 300.0), (np.float32, -1.0)])
def test_pipelined_ingest_equals_plain_upload(gpu_ctx, dtype, scale):
    """rslf_cuda_upload_epis_pipelined normalises the stack and computes the level-0 edge confidence chunk by chunk while
    the rest is uploaded; the run that follows must give exactly what a plain upload gives — also when the parameters
    change between the upload and the run (the precomputed maps are then discarded), and for a second run."""
    epis, _ = make_light_field_np(7, 40, 90, 3, dmin=-1.0, dmax=2.0, seed=321, layers=5)
    if dtype == np.uint8:
        epis = np.clip(np.rint(epis * 255.0), 0, 255).astype(np.uint8)
    elif dtype == np.uint16:
        epis = np.clip(np.rint(epis * 300.0), 0, 65535).astype(np.uint16)
    p = api.default_params()
    ref = oracle.fine_to_coarse(epis, -1.0, 2.0, 24, scale_factor=scale)
    gpu_ctx.upload_epis(epis, scale)                                        # plain
    gpu_ctx.fine_to_coarse_run(-1.0, 2.0, 24, p)
    m0, v0 = gpu_ctx.fine_to_coarse_get()
    gpu_ctx.upload_epis(epis, scale, params=p)                              # pipelined
    gpu_ctx.fine_to_coarse_run(-1.0, 2.0, 24, p)
    m1, v1 = gpu_ctx.fine_to_coarse_get()
    gpu_ctx.fine_to_coarse_run(-1.0, 2.0, 24, p)                            # again on the same input
    m2, v2 = gpu_ctx.fine_to_coarse_get()
    for m, v in ((m0, v0), (m1, v1), (m2, v2)):
        np.testing.assert_array_equal(v, ref["valid"])
        np.testing.assert_array_equal(m, ref["map"])
    q = api.default_params(edge_score_threshold=0.05)
    gpu_ctx.upload_epis(epis, scale, params=p)
    gpu_ctx.fine_to_coarse_run(-1.0, 2.0, 24, q)                            # other edge parameters than at the upload
    m3, v3 = gpu_ctx.fine_to_coarse_get()
    r3 = oracle.fine_to_coarse(epis, -1.0, 2.0, 24, scale_factor=scale, params=oracle.default_params(edge_score_threshold=0.05))
    np.testing.assert_array_equal(m3, r3["map"])
    np.testing.assert_array_equal(v3, r3["valid"])
    c = api.Depth1DComputer_pile(epis, -1.0, 2.0, 24, epi_scale_factor=scale, ctx=gpu_ctx).run()     # another computer afterwards
    pile = oracle.depth1d_pile(oracle.normalise(epis, scale), -1.0, 2.0, 24)
    np.testing.assert_array_equal(c.m_best_depth_v_u, pile["best_depth"])
