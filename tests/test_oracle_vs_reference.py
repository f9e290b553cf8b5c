"""Pins the restated CPU oracle (oracle/rslf_oracle.cpp) against the REFERENCE'S OWN code: the sources under
/root/reference/RSLightFields compiled, where they lie, against the stand-in OpenCV headers of oracle/cvshim
(oracle/ref_driver.cpp -> oracle/_ref/librslf_ref.so).  The reference supplies every piece of control flow of the
path (pass order, masks, propagation, bound propagation, pyramid, fusion); all maps must be bit-identical.

Also checks both against the fixtures the reference build wrote (tests/golden/ref_*.npz, made by
tests/golden/make_ref_golden.py), so that the pin survives on machines where the reference tree is absent.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import ref
from remotesensingproject_b200.synth import make_light_field_np

needs_ref = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/librslf_ref.so not built and no reference tree")


def lf(S, V, U, C, seed, dmin=-1.0, dmax=2.0, **kw):
    epis, _ = make_light_field_np(S, V, U, C, dmin=dmin, dmax=dmax, seed=seed, layers=5, **kw)
    return epis


def assert_same(a, b, keys, tag=""):
    for k in keys:
        x, y = a[k], b[k]
        assert x.shape == y.shape, (tag, k)
        bad = np.flatnonzero(x.ravel() != y.ravel())
        assert bad.size == 0, "%s %s: %d / %d differ, first at %d: %r vs %r" % (
            tag, k, bad.size, x.size, bad[0], x.ravel()[bad[0]], y.ravel()[bad[0]])


MAPS = ["edge_mask", "edge_conf", "best_depth", "disp_conf", "rbar"]

PILE = [
    # S, V, U, C, D, s_hat
    (9, 8, 64, 3, 16, -1), (8, 8, 64, 1, 40, -1), (7, 6, 50, 3, 100, -1), (7, 6, 50, 1, 300, 2),
    (5, 6, 40, 3, 70, 0), (21, 5, 40, 3, 40, 20), (12, 4, 33, 1, 33, 5),
]


@needs_ref
@pytest.mark.parametrize("S,V,U,C,D,s_hat", PILE)
def test_pile(S, V, U, C, D, s_hat):
    epis = lf(S, V, U, C, seed=100 + S + D)
    r = ref.depth1d_pile(epis, -1.0, 2.0, D, s_hat=s_hat, scale_factor=1.0)
    o = oracle.depth1d_pile(oracle.normalise(epis, 1.0), -1.0, 2.0, D, s_hat=s_hat)
    assert o["computed_pixels"] > 0
    assert_same(r, o, MAPS, "pile")


@needs_ref
@pytest.mark.parametrize("S,U,C,D,s_hat,kw", [(9, 64, 3, 16, -1, {}), (8, 64, 1, 40, -1, {}), (7, 50, 3, 100, 2, {}),
                                                (12, 33, 1, 33, 5, dict(edge_confidence_opening_size=3))])
def test_single_epi_computer(S, U, C, D, s_hat, kw):
    """rslf::Depth1DComputer<T> (dc.hpp:254-363): one EPI, one line, no median, no opening (the parameter is ignored)."""
    epi = lf(S, 3, U, C, seed=100 + S + D)[1]
    p = oracle.default_params(**kw)
    r = ref.depth1d(epi, -1.0, 2.0, D, s_hat=s_hat, scale_factor=1.0, params=p)
    o = oracle.depth1d(oracle.normalise(epi[None], 1.0)[0], -1.0, 2.0, D, s_hat=s_hat, params=p)
    assert o["computed_pixels"] > 0
    assert_same(r, o, MAPS, "single EPI")
    u8 = np.clip(np.rint(epi * 255.0), 0, 255).astype(np.uint8)
    assert_same(ref.depth1d(u8, -1.0, 2.0, D, s_hat=s_hat), oracle.depth1d(oracle.normalise(u8[None])[0], -1.0, 2.0, D, s_hat=s_hat),
                MAPS, "single EPI, 8-bit")


@needs_ref
def test_pile_input_normalisation():
    """8-bit input (x 1/255), float input divided by the stack maximum (scale < 0) or by a given factor (dc.hpp:442-477)."""
    epis = lf(7, 6, 48, 3, seed=5)
    u8 = np.clip(np.rint(epis * 255.0), 0, 255).astype(np.uint8)
    assert_same(ref.depth1d_pile(u8, -1.0, 2.0, 24), oracle.depth1d_pile(oracle.normalise(u8), -1.0, 2.0, 24), MAPS, "u8")
    sky = (epis * 255.0 + 8.0).astype(np.float32)
    assert_same(ref.depth1d_pile(sky, -1.0, 2.0, 24, scale_factor=-1.0),
                oracle.depth1d_pile(oracle.normalise(sky, -1.0), -1.0, 2.0, 24), MAPS, "max-scaled")
    assert_same(ref.depth1d_pile(sky, -1.0, 2.0, 24, scale_factor=300.0),
                oracle.depth1d_pile(oracle.normalise(sky, 300.0), -1.0, 2.0, 24), MAPS, "factor")


@needs_ref
def test_pile_negative_radiances_and_parameters():
    epis = lf(6, 5, 40, 3, seed=9) - np.float32(0.2)
    assert_same(ref.depth1d_pile(epis, -1.0, 1.0, 20, s_hat=1, scale_factor=1.0),
                oracle.depth1d_pile(oracle.normalise(epis, 1.0), -1.0, 1.0, 20, s_hat=1), MAPS, "negative")
    epis = lf(9, 6, 60, 1, seed=21)
    p = oracle.default_params(mean_shift_max_iter=3, kernel_h=0.35, edge_score_threshold=0.05, raw_score_threshold=0.4,
                              median_filter_size=3, median_filter_epsilon=0.05, edge_confidence_filter_size=5,
                              cut_shadows=0)
    assert_same(ref.depth1d_pile(epis, -2.0, 2.0, 50, scale_factor=1.0, params=p),
                oracle.depth1d_pile(oracle.normalise(epis, 1.0), -2.0, 2.0, 50, params=p), MAPS, "params")


@needs_ref
@pytest.mark.parametrize("S,V,U,C,D", [(6, 24, 64, 3, 32), (7, 12, 50, 1, 24), (11, 16, 80, 3, 20), (4, 9, 41, 3, 33)])
def test_depth2d(S, V, U, C, D):
    epis = lf(S, V, U, C, seed=7 + S)
    r = ref.depth2d(epis, -1.0, 2.0, D, scale_factor=1.0)
    o = oracle.depth2d(oracle.normalise(epis, 1.0), -1.0, 2.0, D)
    assert_same(r, o, MAPS, "2d")
    # validity masks of get_valid_depths_mask_s_v_u (dc.hpp:893-915): C_e > threshold
    np.testing.assert_array_equal(r["valid"], np.where(o["edge_conf"] > np.float32(0.02), 255, 0).astype(np.uint8))


@needs_ref
def test_depth2d_per_pixel_bounds_and_propagation_epsilon():
    S, V, U, C, D = 7, 10, 56, 3, 24
    epis = lf(S, V, U, C, seed=33)
    rng = np.random.default_rng(1)
    dmn = rng.uniform(-1.0, 0.0, size=(S, V, U)).astype(np.float32)
    dmx = (dmn + rng.uniform(0.0, 2.5, size=(S, V, U))).astype(np.float32)
    dmx[0, 0, :8] = dmn[0, 0, :8]                  # dmin == dmax
    p = oracle.default_params(propagation_epsilon=0.03)
    r = ref.depth2d(epis, -1.0, 2.0, D, scale_factor=1.0, params=p, dmin_svu=dmn, dmax_svu=dmx)
    o = oracle.depth2d(oracle.normalise(epis, 1.0), -1.0, 2.0, D, params=p, dmin_svu=dmn, dmax_svu=dmx)
    assert_same(r, o, MAPS, "2d bounds")


FTC = [
    # S, V, U, C, D, uint8 input, scale factor
    (6, 24, 64, 3, 32, False, 1.0),          # 2 levels
    (5, 45, 90, 1, 24, False, -1.0),         # 3 levels, odd sizes (45 -> 22 -> 11), stack maximum
    (5, 44, 70, 3, 16, True, -1.0),          # 8-bit pyramid (integer blur / resize), 3 levels
    (4, 27, 135, 1, 20, True, -1.0),         # cvRound half-to-even sizes (135 -> 68, 27 -> 14)
]


@needs_ref
@pytest.mark.parametrize("S,V,U,C,D,as_u8,scale", FTC)
def test_fine_to_coarse(S, V, U, C, D, as_u8, scale):
    epis = lf(S, V, U, C, seed=50 + V)
    if as_u8:
        epis = np.clip(np.rint(epis * 255.0), 0, 255).astype(np.uint8)
    elif scale < 0:
        epis = (epis * 200.0 + 3.0).astype(np.float32)
    dims = oracle.pyramid_dims(V, U)
    r = ref.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=scale, dims=dims)
    o = oracle.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=scale)
    assert r["dims"] == dims and len(dims) >= 2
    for p in range(len(dims)):
        assert_same(r["levels"][p], o["levels"][p], ["edge_mask", "edge_conf", "dmin", "dmax", "best_depth", "disp_conf"], "level %d" % p)
    assert_same(r, o, ["valid", "map"], "fused")


@needs_ref
def test_fine_to_coarse_options():
    epis = lf(5, 40, 64, 3, seed=77)
    dims = oracle.pyramid_dims(40, 64, 2)
    r = ref.fine_to_coarse(epis, -1.0, 2.0, 16, scale_factor=1.0, max_pyr_depth=2, accept_all_last=False, dims=dims)
    o = oracle.fine_to_coarse(epis, -1.0, 2.0, 16, scale_factor=1.0, max_pyr_depth=2, accept_all_last=False)
    assert len(dims) == 2
    assert_same(r, o, ["valid", "map"], "options")


@needs_ref
def test_free_functions():
    """downsample_EPIs, fuse_disp_maps, compute_1D_edge_confidence_pile, selective_median_filter."""
    for (S, V, U, C) in [(3, 13, 27, 3), (4, 20, 32, 1), (2, 9, 135, 1)]:
        epis = lf(S, V, U, C, seed=V)
        np.testing.assert_array_equal(ref.downsample(epis), oracle.downsample(epis))
        u8 = np.clip(np.rint(epis * 255.0), 0, 255).astype(np.uint8)
        np.testing.assert_array_equal(ref.downsample(u8), oracle.downsample(u8))
        for s in (0, S - 1):
            ce_r, m_r = ref.edge_confidence(epis, s)
            ce_o, m_o = oracle.edge_confidence(epis, s)
            np.testing.assert_array_equal(m_r, m_o)
            np.testing.assert_array_equal(ce_r, ce_o)
        rng = np.random.default_rng(V)
        src = rng.uniform(-1, 2, size=(V, U)).astype(np.float32)
        mask = np.where(rng.random((V, U)) < 0.6, 255, 0).astype(np.uint8)
        got = ref.selective_median(src, mask, epis, S // 2)
        want = oracle.selective_median(src, mask, epis, S // 2)
        np.testing.assert_array_equal(got[mask > 0], want[mask > 0])
    rng = np.random.default_rng(3)
    sizes = [(21, 37), (10, 18), (5, 9)]
    disp = [rng.uniform(-1, 3, size=(2, v, u)).astype(np.float32) for (v, u) in sizes]
    valid = [np.where(rng.random((2, v, u)) < 0.5, 255, 0).astype(np.uint8) for (v, u) in sizes]
    m_r, v_r = ref.fuse(disp, valid)
    m_o, v_o = oracle.fuse(disp, valid)
    np.testing.assert_array_equal(v_r, v_o)
    np.testing.assert_array_equal(m_r, m_o)


@needs_ref
@pytest.mark.parametrize("shape,size", [(2, 3), (0, 2), (1, 5), (2, 4)])
def test_edge_mask_opening(shape, size):
    """par_edge_confidence_opening_size > 1: morphological opening of every V x U edge mask (core.hpp:759-769),
    through the pile, the 2D computer and the pyramid."""
    p = oracle.default_params(edge_confidence_opening_type=shape, edge_confidence_opening_size=size)
    epis = lf(7, 14, 60, 3, seed=40 + size)
    ce_r, m_r = ref.edge_confidence(epis, 3, params=p)
    ce_o, m_o = oracle.edge_confidence(epis, 3, params=p)
    _, m_plain = oracle.edge_confidence(epis, 3)
    np.testing.assert_array_equal(m_r, m_o)
    np.testing.assert_array_equal(ce_r, ce_o)
    assert (m_o != m_plain).any() and m_o.any()
    if size % 2:                         # even elements are anchored off-centre: cv2's opening then shifts the mask
        assert (m_o <= m_plain).all()
    assert_same(ref.depth1d_pile(epis, -1.0, 2.0, 24, scale_factor=1.0, params=p),
                oracle.depth1d_pile(oracle.normalise(epis, 1.0), -1.0, 2.0, 24, params=p), MAPS, "pile opening")
    assert_same(ref.depth2d(epis, -1.0, 2.0, 24, scale_factor=1.0, params=p),
                oracle.depth2d(oracle.normalise(epis, 1.0), -1.0, 2.0, 24, params=p), MAPS, "2d opening")
    epis = lf(5, 24, 64, 1, seed=41 + size)
    dims = oracle.pyramid_dims(24, 64)
    r = ref.fine_to_coarse(epis, -1.0, 2.0, 16, scale_factor=1.0, params=p, dims=dims)
    o = oracle.fine_to_coarse(epis, -1.0, 2.0, 16, scale_factor=1.0, params=p)
    assert_same(r, o, ["valid", "map"], "ftc opening")


@needs_ref
@pytest.mark.parametrize("S,V,U,C,scale", [(5, 44, 70, 3, -1.0), (4, 27, 135, 1, -1.0), (5, 24, 50, 3, 4095.0)])
def test_16_bit_input(S, V, U, C, scale):
    """CV_16U stacks: x float(1 / scale) with scale = the stack maximum or a given factor (dc.hpp:442-477); the pyramid
    stays 16-bit between levels (ftc.hpp:142-147; integer Gaussian and half-size resize)."""
    epis = lf(S, V, U, C, seed=V)
    u16 = np.clip(np.rint(epis * (4095.0 if scale > 0 else 60000.0)), 0, 65535).astype(np.uint16)
    np.testing.assert_array_equal(ref.downsample(u16), oracle.downsample(u16))
    assert_same(ref.depth1d_pile(u16, -1.0, 2.0, 24, scale_factor=scale),
                oracle.depth1d_pile(oracle.normalise(u16, scale), -1.0, 2.0, 24), MAPS, "u16 pile")
    dims = oracle.pyramid_dims(V, U)
    r = ref.fine_to_coarse(u16, -1.0, 2.0, 16, scale_factor=scale, dims=dims)
    o = oracle.fine_to_coarse(u16, -1.0, 2.0, 16, scale_factor=scale)
    assert len(dims) >= 2 and o["computed_pixels"] > 0
    for p in range(len(dims)):
        assert_same(r["levels"][p], o["levels"][p], ["edge_mask", "edge_conf", "dmin", "dmax", "best_depth", "disp_conf"], "u16 level %d" % p)
    assert_same(r, o, ["valid", "map"], "u16 fused")


@needs_ref
@pytest.mark.parametrize("S,V,U,C,saturate,kw", [
    (6, 24, 64, 3, True, {}), (5, 45, 90, 1, True, {}), (4, 24, 50, 3, False, {}),
    (7, 24, 64, 3, True, dict(cut_shadows=0)), (2, 24, 40, 1, True, {})])
def test_coloured_depth_maps(golden_dir, S, V, U, C, saturate, kw):
    """FineToCoarse::get_coloured_depth_maps (ftc.hpp:324-377; ImageConverter_uchar, rslf_plot.cpp:66-107) of the
    reference build, with OpenCV's JET table handed to the stand-in applyColorMap."""
    lut = np.load(os.path.join(golden_dir, "colormap_jet.npy"))
    epis = lf(S, V, U, C, seed=50 + V, dark_fraction=0.2)
    p = oracle.default_params(**kw)
    r = ref.fine_to_coarse_coloured(epis, -1.0, 2.0, 16, lut, scale_factor=1.0, params=p, saturate=saturate)
    o = oracle.fine_to_coarse(epis, -1.0, 2.0, 16, scale_factor=1.0, params=p)
    c, (mn, mx) = oracle.colour_maps(o["map"], o["valid"], oracle.normalise(epis, 1.0), lut, saturate=saturate,
                                     cut_shadows=p.cut_shadows)
    assert mx > mn and c.any()
    np.testing.assert_array_equal(r, c)


GOLDEN = ["ref_pile_c3", "ref_pile_c1_u8", "ref_2d_c3", "ref_ftc_c1", "ref_ftc_c3_u8"]


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_reference_fixtures(golden_dir, name):
    """The fixtures written by the reference build (tests/golden/make_ref_golden.py): works without the reference."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_ref_golden", os.path.join(golden_dir, "make_ref_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    run_case = mod.run_case
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    out = run_case(name, g["epis"], oracle_side=True)
    for k in g.files:
        if k == "epis":
            continue
        np.testing.assert_array_equal(out[k], g[k], err_msg="%s:%s" % (name, k))


@pytest.mark.parametrize("name", ["ref_opening_pile_c3", "ref_u16_ftc_c3", "ref_coloured_c3", "ref_single_epi_c1"])
def test_oracle_matches_reference_fixtures_of_the_late_rows(golden_dir, name):
    """Fixtures written by the reference build for the rows added late (edge-mask opening, CV_16U input, coloured maps,
    the single-EPI computer): the pin survives where the reference tree is absent."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_ref_golden", os.path.join(golden_dir, "make_ref_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    out = mod.run_late_case(name, g["epis"], side="oracle")
    assert set(out) == set(g.files) - {"epis"}
    for k in out:
        np.testing.assert_array_equal(out[k], g[k], err_msg="%s:%s" % (name, k))



# --------------------------------------------------------------------------- the disparity-confidence criterion
needs_ref_cd = pytest.mark.skipif(not ref.available_cd(), reason="oracle/_ref/librslf_ref_cd.so not built and no reference tree")


@pytest.fixture
def disp_criterion():
    oracle.set_criterion("disp")
    ref.set_criterion("disp")
    yield
    oracle.set_criterion("edge")
    ref.set_criterion("edge")


@needs_ref_cd
@pytest.mark.parametrize("S,V,U,C,D,thr", [(7, 24, 64, 3, 24, 0.01), (6, 44, 70, 1, 32, 0.01), (9, 12, 80, 3, 16, 0.05), (5, 23, 50, 3, 40, 0.0)])
def test_disp_confidence_criterion_as_intended(disp_criterion, S, V, U, C, D, thr):
    """The reference with -D_USE_DISP_CONFIDENCE_SCORE does not compile as written (`#elseif` is no directive); built
    with `#elseif` -> `#elif` (oracle/Makefile, target ref_cd) it gates the propagation sources and the validity maps
    by C_d > par_disp_score_threshold (core.hpp:1097-1098, dc.hpp:901-902).  The oracle's restatement of that variant
    must equal it bit for bit, and it must differ from the default build (otherwise the test tests nothing)."""
    epis = lf(S, V, U, C, seed=500 + S + V)
    p = oracle.default_params(disp_score_threshold=thr)
    o = oracle.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=1.0, params=p)
    r = ref.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=1.0, params=p, dims=o["dims"])
    assert_same(r, o, ["map", "valid"], "ftc disp")
    for lo, lr in zip(o["levels"], r["levels"]):
        assert_same(lr, lo, ["edge_mask", "edge_conf", "best_depth", "disp_conf", "dmin", "dmax"], "ftc disp level")
    o2 = oracle.depth2d(oracle.normalise(epis, 1.0), -1.0, 2.0, D, params=p)
    r2 = ref.depth2d(epis, -1.0, 2.0, D, scale_factor=1.0, params=p)
    assert_same(r2, o2, MAPS, "2d disp")
    oracle.set_criterion("edge")
    e2 = oracle.depth2d(oracle.normalise(epis, 1.0), -1.0, 2.0, D, params=p)
    oracle.set_criterion("disp")
    if thr > 0:
        assert e2["computed_pixels"] != o2["computed_pixels"] or not np.array_equal(e2["best_depth"], o2["best_depth"])
