"""The C++ host side (include/rslf_b200.hpp: the reference's class API over the C ABI)
run as a compiled program and compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from remotesensingproject_b200.synth import make_light_field_np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "facade_demo")
    libdir = os.path.join(ROOT, "remotesensingproject_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "facade_demo.cpp"), "-o", exe,
                           "-L" + libdir, "-lrslf_b200", "-Wl,-rpath," + libdir])
    return exe


def test_facade_compiles_without_gpu(tmp_path):
    _build(tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("C", [3, 1])
def test_facade_matches_oracle(tmp_path, C):
    exe = _build(tmp_path)
    S, V, U, D = 5, 24, 40, 12
    epis, _ = make_light_field_np(S, V, U, C, dmin=-1.0, dmax=2.0, seed=90 + C, layers=5)
    inp = str(tmp_path / "in.bin")
    epis.tofile(inp)
    out = str(tmp_path / "out")
    lut = np.load(os.path.join(ROOT, "tests", "golden", "colormap_jet.npy"))
    lut_path = str(tmp_path / "jet.bin")
    lut.tofile(lut_path)
    r = subprocess.run([exe, inp, str(V), str(S), str(U), str(C), str(D), "-1.0", "2.0", out, lut_path],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "expected failure" in r.stdout
    norm = oracle.normalise(epis, 1.0)
    one = oracle.depth1d(norm[0], -1.0, 2.0, D)
    np.testing.assert_array_equal(np.fromfile(out + "_1d_depth.bin", np.float32), one["best_depth"])
    np.testing.assert_array_equal(np.fromfile(out + "_1d_cd.bin", np.float32), one["disp_conf"])
    pile = oracle.depth1d_pile(norm, -1.0, 2.0, D)
    np.testing.assert_array_equal(np.fromfile(out + "_pile_depth.bin", np.float32).reshape(V, U), pile["best_depth"])
    np.testing.assert_array_equal(np.fromfile(out + "_pile_mask.bin", np.uint8).reshape(V, U), pile["edge_mask"])
    d2 = oracle.depth2d(norm, -1.0, 2.0, D)
    np.testing.assert_array_equal(np.fromfile(out + "_2d_depth.bin", np.float32).reshape(S, V, U), d2["best_depth"])
    np.testing.assert_array_equal(np.fromfile(out + "_2d_cd.bin", np.float32).reshape(S, V, U), d2["disp_conf"])
    np.testing.assert_array_equal(np.fromfile(out + "_2d_valid.bin", np.uint8).reshape(S, V, U),
                                  (d2["edge_conf"] > np.float32(0.02)).astype(np.uint8) * 255)
    ftc = oracle.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=1.0)
    np.testing.assert_array_equal(np.fromfile(out + "_ftc_map.bin", np.float32).reshape(S, V, U), ftc["map"])
    np.testing.assert_array_equal(np.fromfile(out + "_ftc_valid.bin", np.uint8).reshape(S, V, U), ftc["valid"])
    bgr, _ = oracle.colour_maps(ftc["map"], ftc["valid"], norm, lut)
    np.testing.assert_array_equal(np.fromfile(out + "_ftc_bgr.bin", np.uint8).reshape(S, V, U, 3), bgr)
    # the free functions and get_epis
    np.testing.assert_array_equal(np.fromfile(out + "_epis.bin", np.float32).reshape(V, S, U, C), norm)
    ce, mask = oracle.edge_confidence(norm, 1)
    np.testing.assert_array_equal(np.fromfile(out + "_fn_ce.bin", np.float32).reshape(V, U), ce)
    np.testing.assert_array_equal(np.fromfile(out + "_fn_mask.bin", np.uint8).reshape(V, U), mask)
    med = oracle.selective_median(d2["best_depth"][2], d2["edge_mask"][2], norm, 2, size=5, eps=0.1)
    np.testing.assert_array_equal(np.fromfile(out + "_fn_median.bin", np.float32).reshape(V, U), med)
    half = oracle.downsample(norm)
    np.testing.assert_array_equal(np.fromfile(out + "_fn_down.bin", np.float32).reshape(half.shape), half)
    d2b = oracle.depth2d(half, -1.0, 2.0, D)
    np.testing.assert_array_equal(np.fromfile(out + "_fn_l1_depth.bin", np.float32).reshape(d2b["best_depth"].shape), d2b["best_depth"])
    v0 = (d2["edge_conf"] > np.float32(0.02)).astype(np.uint8) * 255
    v1 = np.full(d2b["edge_mask"].shape, 255, np.uint8)
    fm, fv = oracle.fuse([d2["best_depth"], d2b["best_depth"]], [v0, v1])
    np.testing.assert_array_equal(np.fromfile(out + "_fn_fuse_map.bin", np.float32).reshape(S, V, U), fm)
    np.testing.assert_array_equal(np.fromfile(out + "_fn_fuse_valid.bin", np.uint8).reshape(S, V, U), fv)
    np.testing.assert_array_equal(np.fromfile(out + "_fn_2d_depth.bin", np.float32).reshape(S, V, U), d2["best_depth"])
