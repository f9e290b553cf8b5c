"""Pins the CPU oracle (oracle/rslf_oracle.cpp) against the cv2 replay of the
reference's OpenCV call sequence (oracle/cv2_mirror.py -> tests/golden/*.npz).

Masks and argmax indices must agree exactly; floats to 1e-6: cv2 4.13's
multiply(a, a, scale) keeps a double intermediate where OpenCV 3.x (the
reference's version) rounds per float operation, so the last ulp may differ.
"""
import glob
import os

import numpy as np
import pytest

import oracle


def _cases(golden_dir, pat):
    return sorted(glob.glob(os.path.join(golden_dir, pat)))


HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", _cases(HERE, "px_*.npz"), ids=os.path.basename)
def test_edge_confidence_and_pixel_scores(path):
    g = np.load(path)
    epi = g["epi"]
    s_hat = int(g["s_hat"])
    ce, mask = oracle.edge_confidence(epi[None], s_hat)
    np.testing.assert_array_equal(mask[0], g["mask"])
    np.testing.assert_allclose(ce[0], g["ce"], rtol=1e-6, atol=1e-7)
    for i, u in enumerate(g["us"]):
        sc, rb, dv = oracle.pixel_scores(epi, s_hat, int(u), float(g["dmin"]), float(g["dmax"]), int(g["D"]))
        np.testing.assert_array_equal(dv, g["dvals"][i])
        np.testing.assert_allclose(sc, g["score"][i], rtol=0, atol=1e-6)
        np.testing.assert_allclose(rb, g["rbar"][i], rtol=0, atol=1e-6)
        # first maximum, like cv::minMaxLoc
        assert int(np.argmax(sc)) == int(g["best"][i])
        assert abs(float(sc.max()) - float(g["maxv"][i])) < 1e-6
        assert abs(float(sc.astype(np.float64).mean()) - float(g["mean"][i])) < 1e-6


@pytest.mark.parametrize("path", [p for p in _cases(HERE, "down_*.npz") if "_u8_" not in p], ids=os.path.basename)
def test_downsample(path):
    g = np.load(path)
    out = oracle.downsample(g["raw"])
    assert out.shape == g["out"].shape
    np.testing.assert_allclose(out, g["out"], rtol=5e-7, atol=0)


@pytest.mark.parametrize("path", _cases(HERE, "down_u8_*.npz"), ids=os.path.basename)
def test_downsample_uint8_is_bit_exact(path):
    g = np.load(path)
    out = oracle.downsample(g["raw"])
    assert out.dtype == np.uint8
    np.testing.assert_array_equal(out, g["out"])


def test_fuse(golden_dir):
    g = np.load(os.path.join(golden_dir, "fuse_3lvl.npz"))
    m, k = oracle.fuse([g["d0"], g["d1"], g["d2"]], [g["v0"], g["v1"], g["v2"]])
    np.testing.assert_array_equal(k, g["out_valid"])
    np.testing.assert_allclose(m, g["out_map"], rtol=0, atol=5e-6)


def test_visiting_order_quirk():
    # even S: line 0 is never a s_hat (core.hpp:981-990); odd S: every line once
    assert oracle.visiting_order(6) == [3, 4, 2, 5, 1]
    assert oracle.visiting_order(5) == [2, 3, 1, 4, 0]
    assert sorted(oracle.visiting_order(9)) == list(range(9))


def test_pyramid_dims_round_half_even():
    # cvRound(n * 0.5): 135 -> 68, 75 -> 38, 45 -> 22, 33 -> 16
    assert [oracle.half_size(n) for n in (135, 75, 45, 33, 17, 9)] == [68, 38, 22, 16, 8, 4]
    assert oracle.pyramid_dims(600, 1200) == [(600, 1200), (300, 600), (150, 300), (75, 150), (38, 75), (19, 38)]
    assert oracle.pyramid_dims(1080, 1920)[-1] == (17, 30)


def test_constants():
    p = oracle.default_params()
    assert np.float32(p.shadow_level) == np.float32(0.05 * 1.73205080757)
    assert np.float32(1.0 / float(np.float32(0.2) * np.float32(0.2))) == np.float32(24.999998)


def test_structuring_element_and_opening_match_cv2():
    """The optional opening of the edge mask (core.hpp:759-769): getStructuringElement and
    morphologyEx(MORPH_OPEN) of the real OpenCV (cv2 4.13), odd and even sizes, thin images."""
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    for shape in (cv2.MORPH_RECT, cv2.MORPH_CROSS, cv2.MORPH_ELLIPSE):
        for k in range(1, 16):
            np.testing.assert_array_equal(oracle.structuring_element(shape, k), cv2.getStructuringElement(shape, (k, k)))
    rng = np.random.default_rng(11)
    for shape in (0, 1, 2):
        for k in (2, 3, 4, 5, 7, 9):
            for (V, U) in ((1, 9), (7, 1), (6, 11), (23, 37)):
                m = np.where(rng.random((V, U)) < 0.75, 255, 0).astype(np.uint8)
                want = cv2.morphologyEx(m, cv2.MORPH_OPEN, cv2.getStructuringElement(shape, (k, k)))
                np.testing.assert_array_equal(oracle.morph_open(m, shape, k), want, err_msg="shape %d k %d %dx%d" % (shape, k, V, U))


def test_downsample_uint16_matches_cv2():
    """CV_16U pyramids (they stay 16-bit, ftc.hpp:142-147): GaussianBlur 7x7 + resize 0.5 of the real OpenCV, incl.
    exact .5 ties of the blur (values that are multiples of 128), odd sizes and images smaller than the kernel.
    IPP is switched off: OpenCV's own 2x2 resize rounds ties up ((a+b+c+d+2) >> 2), the IPP build of the wheel rounds
    them to even; the blur is identical either way."""
    cv2 = pytest.importorskip("cv2")
    cv2.setNumThreads(1)
    ipp = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        rng = np.random.default_rng(5)
        for t in range(14):
            V, U, S = int(rng.integers(2, 40)), int(rng.integers(2, 40)), 2
            C = int(rng.choice([1, 3]))
            raw = rng.integers(0, 65536, size=(V, S, U, C)).astype(np.uint16)
            if t % 3 == 0:
                raw = (raw // 128 * 128).astype(np.uint16)
            out = oracle.downsample(raw)
            assert out.dtype == np.uint16
            for s in range(S):
                b = cv2.GaussianBlur(np.ascontiguousarray(raw[:, s]), (7, 7), 0, borderType=cv2.BORDER_REFLECT)
                r = cv2.resize(b, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_LINEAR).reshape(out[:, s].shape)
                np.testing.assert_array_equal(out[:, s], r, err_msg="%dx%dx%d" % (V, U, C))
    finally:
        cv2.ipp.setUseIPP(ipp)


def test_coloured_maps_match_cv2_sort_and_colormap(golden_dir):
    """ImageConverter_uchar::fit's cv::sort quantiles and cv::applyColorMap(COLORMAP_JET) of the real OpenCV; the
    committed colour table (tests/golden/colormap_jet.npy) is the one cv2 produces."""
    cv2 = pytest.importorskip("cv2")
    import math
    from remotesensingproject_b200.synth import make_light_field_np
    lut = np.load(os.path.join(golden_dir, "colormap_jet.npy"))
    ramp = np.arange(256, dtype=np.uint8).reshape(256, 1)
    np.testing.assert_array_equal(cv2.applyColorMap(ramp, cv2.COLORMAP_JET).reshape(256, 3), lut)
    S = 6
    epis, _ = make_light_field_np(S, 48, 96, 3, dmin=-1.0, dmax=2.0, seed=3, layers=5, dark_fraction=0.2)
    o = oracle.fine_to_coarse(epis, -1.0, 2.0, 16, scale_factor=1.0)
    norm = oracle.normalise(epis, 1.0)
    c, (mn, mx) = oracle.colour_maps(o["map"], o["valid"], norm, lut)
    plane = o["map"][int(math.floor(S / 2.0 + 0.5))]
    srt = cv2.sort(plane.reshape(-1, 1), cv2.SORT_EVERY_COLUMN + cv2.SORT_ASCENDING)
    assert mn == float(srt[int(math.floor(0.02 * plane.size)), 0]) and mx == float(srt[int(math.floor(0.98 * plane.size)), 0])
    alpha = np.float32(255.0 / (mx - mn))
    beta = np.float32(-float(alpha) * mn)
    for s in range(S):
        u8 = np.clip(np.rint(o["map"][s] * alpha + beta), 0, 255).astype(np.uint8)      # convertTo(CV_8U, alpha, beta)
        col = cv2.applyColorMap(u8, cv2.COLORMAP_JET)
        col[o["valid"][s] == 0] = 0
        nrm = np.sqrt((norm[:, s].astype(np.float64) ** 2).sum(-1)).astype(np.float32)
        col[nrm < np.float32(0.05 * 1.73205080757)] = 0
        np.testing.assert_array_equal(col, c[s])
