"""The contracted (fast_math) mode against the tolerance the path is specified to (BASELINE.json north_star):
bit-exact masks and disparity indices wherever the reference's winning score margin exceeds 1e-5, <= 1e-4 relative
on scores and confidences.

Method (oracle REPLAY, oracle/rslf_oracle.cpp `Replay`): the CUDA run records every pixel decision (index, score, r_bar,
C_d); the exact CPU oracle then re-runs the same input and, for every pixel IT evaluates, looks the recorded decision up,
checks it — the recorded index must be a hypothesis whose EXACT score is within 1e-5 of the exact maximum, the recorded
floats within 1e-4 relative of the exact ones for that hypothesis — and adopts it.  Everything downstream (selective
median, propagation and the remaining masks, hence WHICH pixels are evaluated in later passes, bound propagation,
fusion) is exact arithmetic in both runs, so with the decisions adopted every map of every level and the fused map must
be IDENTICAL, and no evaluated pixel may lack a record.  A single decision outside the tolerance, or any other
discrepancy of the pipeline, fails the test; nothing is "allowed to differ a little".
"""
import numpy as np
import pytest

import bench
import oracle
from remotesensingproject_b200 import api
from remotesensingproject_b200.synth import make_light_field_np

pytestmark = pytest.mark.gpu


def replay_check(gpu_ctx, epis, dmin, dmax, D, scale, fast, expect_changed=None):
    S = epis.shape[1]
    gpu_ctx.set_decision_log(int(epis.shape[0] * epis.shape[2] * S * 1.5) + 1024)
    gpu_ctx.set_fast_math(fast)
    try:
        f = api.FineToCoarse(epis, dmin, dmax, D, epi_scale_factor=scale, ctx=gpu_ctx).run()
        out_map, out_valid = f.get_results()
        levels = f.get_levels()
        recs, n = gpu_ctx.decision_log()
        assert n == gpu_ctx.timing()["computed_pixels"]
    finally:
        gpu_ctx.set_fast_math(False)
        gpu_ctx.set_decision_log(0)
    oracle.set_num_threads(__import__("os").cpu_count() or 1)
    with oracle.replay(recs, n) as rp:
        o = oracle.fine_to_coarse(epis, dmin, dmax, D, scale_factor=scale)
    r = rp.report
    assert r["looked_up"] == n == r["records"], r            # the two runs evaluated the same pixels, each once
    for k in ("missing", "bad_margin", "bad_score", "bad_rbar", "bad_cd", "bad_disparity"):
        assert r[k] == 0, r
    assert r["worst_margin"] <= 1e-5 and r["worst_rel_score"] <= 1e-4, r
    for p, (g, l) in enumerate(zip(levels, o["levels"])):
        for k in ("edge_mask", "edge_conf", "dmin", "dmax", "best_depth", "disp_conf"):
            np.testing.assert_array_equal(g[k], l[k], err_msg="level %d %s" % (p, k))
    np.testing.assert_array_equal(out_valid, o["valid"])
    np.testing.assert_array_equal(out_map, o["map"])
    if expect_changed is not None:
        assert (r["index_changed"] > 0) == expect_changed, r
    return r


def test_exact_mode_replays_without_a_single_change(gpu_ctx):
    """Sanity of the method: the default (exact) mode adopts nothing — every recorded index is the oracle's own."""
    epis, _ = make_light_field_np(26, 24, 96, 3, dmin=-1.0, dmax=2.0, seed=77, layers=5)
    r = replay_check(gpu_ctx, epis, -1.0, 2.0, 40, 1.0, fast=False, expect_changed=False)
    assert r["worst_rel_score"] == 0.0


@pytest.mark.parametrize("S,V,U,D,seed", [(26, 24, 96, 40, 77), (40, 46, 128, 72, 78), (24, 23, 200, 33, 79)])
def test_fast_math_within_the_specified_tolerance(gpu_ctx, S, V, U, D, seed):
    epis, _ = make_light_field_np(S, V, U, 3, dmin=-1.0, dmax=2.0, seed=seed, layers=5)
    replay_check(gpu_ctx, epis, -1.0, 2.0, D, 1.0, fast=True)


def test_fast_math_c3_band(gpu_ctx):
    """The geometry the bench times (C3: 100 views x 1920 columns RGB, D = 256), 24 rows = two pyramid levels."""
    cfg = bench.CONFIGS["c3"]
    epis = bench.cpu_sample(cfg, 24)
    r = replay_check(gpu_ctx, epis, bench.DMIN, bench.DMAX, cfg["D"], cfg["scale"], fast=True)
    print("c3 band, fast_math: %d decisions, %d adopted a different index (all within 1e-5 of the exact maximum), "
          "worst margin %.3g, worst relative score error %.3g" % (r["records"], r["index_changed"], r["worst_margin"], r["worst_rel_score"]))


def test_fast_math_is_refused_where_it_would_be_a_no_op(gpu_ctx):
    """Gray stacks take the shared-memory depth kernel, which has no contracted variant: the run must say so."""
    epis, _ = make_light_field_np(9, 12, 64, 1, dmin=-1.0, dmax=2.0, seed=5, layers=5)
    gpu_ctx.set_fast_math(True)
    try:
        with pytest.raises(api.RslfError):
            api.FineToCoarse(epis, -1.0, 2.0, 16, epi_scale_factor=1.0, ctx=gpu_ctx).run()
    finally:
        gpu_ctx.set_fast_math(False)
