"""GPU parity of the rows SURVEY.md 8(f) marks "next" (dormant options of the reference's parameter block and
input formats either side of the hot path), through the C ABI against the CPU oracle.  Same bar as
tests/test_gpu_parity.py: every map identical.
"""
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
