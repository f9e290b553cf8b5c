"""GPU parity AT THE GEOMETRY THE BENCH TIMES: bands of rows of the BASELINE.json configurations (all views, all
columns, the configuration's D), so that the very launches bench.py measures — for C3 the tensor-memory depth kernel
at S = 100, U = 1920, D = 256 (8 chunks of hypotheses, two tensor-memory staging rounds, +-200 px segments), per-pixel
bounds on the coarser level, flat (dmin == dmax) pixels — are compared with the reference.

The checker is the reference's own sources (oracle/_ref/librslf_ref.so, compiled in the CPU container from
/root/reference and shipped) when present, else the restated oracle, which tests/test_oracle_vs_reference.py pins
bit-exact to it.  Every level map and the fused map must be identical (assert_array_equal), C_d within the north
star's 1e-4 relative (its mean over the hypotheses is a double sum in another order)."""
import os

import numpy as np
import pytest

import bench
import oracle
from oracle import ref
from remotesensingproject_b200 import api

pytestmark = pytest.mark.gpu

DMIN, DMAX = bench.DMIN, bench.DMAX


def band(name, rows):
    cfg = bench.CONFIGS[name]
    return cfg, bench.cpu_sample(cfg, rows)


def checker_ftc(epis, cfg, dims):
    if os.path.exists(ref.LIB_PATH):
        ref.set_num_threads(os.cpu_count() or 1)
        return ref.fine_to_coarse(epis, DMIN, DMAX, cfg["D"], scale_factor=cfg["scale"], dims=dims), "reference build"
    oracle.set_num_threads(os.cpu_count() or 1)
    return oracle.fine_to_coarse(epis, DMIN, DMAX, cfg["D"], scale_factor=cfg["scale"]), "oracle"


def checker_2d(epis, cfg):
    if os.path.exists(ref.LIB_PATH):
        ref.set_num_threads(os.cpu_count() or 1)
        return ref.depth2d(epis, DMIN, DMAX, cfg["D"], scale_factor=cfg["scale"]), "reference build"
    oracle.set_num_threads(os.cpu_count() or 1)
    return oracle.depth2d(oracle.normalise(epis, cfg["scale"]), DMIN, DMAX, cfg["D"]), "oracle"


def same(a, b, what):
    assert a.shape == b.shape, what
    bad = np.flatnonzero(a.ravel() != b.ravel())
    assert bad.size == 0, "%s: %d / %d differ, first at %d: %r vs %r" % (what, bad.size, a.size, bad[0], a.ravel()[bad[0]], b.ravel()[bad[0]])


def check_ftc(gpu_ctx, name, rows):
    cfg, epis = band(name, rows)
    f = api.FineToCoarse(epis, DMIN, DMAX, cfg["D"], epi_scale_factor=cfg["scale"], ctx=gpu_ctx).run()
    out_map, out_valid = f.get_results()
    levels = f.get_levels()
    dims = [(l["best_depth"].shape[1], l["best_depth"].shape[2]) for l in levels]
    r, who = checker_ftc(epis, cfg, dims)
    assert len(r["levels"]) == len(levels) >= 2, "the band must span at least two pyramid levels"
    flat = 0
    for p, (g, o) in enumerate(zip(levels, r["levels"])):
        for k in ("edge_mask", "edge_conf", "dmin", "dmax", "best_depth"):
            same(g[k], o[k], "%s level %d %s vs %s" % (name, p, k, who))
        np.testing.assert_allclose(g["disp_conf"], o["disp_conf"], rtol=1e-4, atol=1e-7, err_msg="level %d C_d" % p)
        flat += int(np.count_nonzero((o["dmin"] == o["dmax"]) & (o["edge_mask"] != 0)))
    same(out_valid, r["valid"], name + " fused validity vs " + who)
    same(out_map, r["map"], name + " fused map vs " + who)
    return gpu_ctx.timing(), flat


def test_c3_band_fine_to_coarse(gpu_ctx):
    """C3 (100 views x 1920 columns RGB, D = 256): 24 rows -> levels of 24 and 12 rows."""
    t, flat = check_ftc(gpu_ctx, "c3", 24)
    assert t["levels"] == 2 and t["passes"] == 2 * 99
    assert flat > 0, "the coarser level is expected to hold flat (dmin == dmax) pixels: the shortcut must be exercised"


def test_c3_band_depth2d(gpu_ctx):
    """C3 geometry through Depth2DComputer (single scale): every map of the class."""
    cfg, epis = band("c3", 12)
    c = api.Depth2DComputer(epis, DMIN, DMAX, cfg["D"], epi_scale_factor=cfg["scale"], ctx=gpu_ctx).run()
    r, who = checker_2d(epis, cfg)
    same(c.m_edge_confidence_mask_s_v_u, r["edge_mask"], "edge mask vs " + who)
    same(c.m_edge_confidence_s_v_u, r["edge_conf"], "C_e vs " + who)
    same(c.m_best_depth_s_v_u, r["best_depth"], "disparity vs " + who)
    same(c.m_rbar_s_v_u, r["rbar"], "r_bar vs " + who)
    np.testing.assert_allclose(c.m_disp_confidence_s_v_u, r["disp_conf"], rtol=1e-4, atol=1e-7)


def test_c1_band_depth2d(gpu_ctx):
    """C1 (50 views x 960 columns RGB, D = 64, single scale)."""
    cfg, epis = band("c1", 32)
    c = api.Depth2DComputer(epis, DMIN, DMAX, cfg["D"], epi_scale_factor=cfg["scale"], ctx=gpu_ctx).run()
    r, who = checker_2d(epis, cfg)
    same(c.m_edge_confidence_mask_s_v_u, r["edge_mask"], "edge mask vs " + who)
    same(c.m_best_depth_s_v_u, r["best_depth"], "disparity vs " + who)
    same(c.m_rbar_s_v_u, r["rbar"], "r_bar vs " + who)
    np.testing.assert_allclose(c.m_disp_confidence_s_v_u, r["disp_conf"], rtol=1e-4, atol=1e-7)


def test_c2_band_fine_to_coarse(gpu_ctx):
    """C2 (30 views x 1200 columns gray in the SkySat range, D = 128, scale = -1: per-level maximum)."""
    t, _ = check_ftc(gpu_ctx, "c2", 48)
    assert t["levels"] == 3


def test_c5_field_band_fine_to_coarse(gpu_ctx):
    """C5 geometry (one of the batched fields: 50 views x 1280 columns RGB, D = 128)."""
    check_ftc(gpu_ctx, "c5", 24)


def test_c4_band_fine_to_coarse(gpu_ctx):
    """C4 geometry (200 views x 3840 columns RGB, D = 512): 12 rows, one level; S = 200 exceeds what the
    tensor-memory planner takes, so this is the shared-memory depth kernel with 16 chunks."""
    cfg, epis = band("c4", 12)
    f = api.FineToCoarse(epis, DMIN, DMAX, cfg["D"], epi_scale_factor=cfg["scale"], ctx=gpu_ctx).run()
    out_map, out_valid = f.get_results()
    r, who = checker_ftc(epis, cfg, [(12, cfg["U"])])
    same(out_valid, r["valid"], "c4 fused validity vs " + who)
    same(out_map, r["map"], "c4 fused map vs " + who)
