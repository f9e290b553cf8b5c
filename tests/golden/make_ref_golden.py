"""Writes tests/golden/ref_*.npz: outputs of the REFERENCE'S OWN code (oracle/_ref/librslf_ref.so = the sources under
/root/reference/RSLightFields compiled against oracle/cvshim) on small seeded inputs.  Run in the CPU container:

    python tests/golden/make_ref_golden.py

`run_case` is shared with the tests so that the oracle and the CUDA path are run with exactly the same arguments:
side "ref" = the reference build, "oracle" = oracle/rslf_oracle.cpp, "gpu" = the CUDA library through api.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> (kind, S, V, U, C, uint8 input, scale factor, dmin, dmax, D, seed)
CASES = {
    "ref_pile_c3": ("pile", 9, 6, 48, 3, False, 1.0, -1.0, 2.0, 40, 901),
    "ref_pile_c1_u8": ("pile", 8, 6, 48, 1, True, -1.0, -1.0, 2.0, 33, 902),
    "ref_2d_c3": ("2d", 5, 8, 40, 3, False, 1.0, -1.0, 2.0, 24, 903),
    "ref_ftc_c1": ("ftc", 4, 26, 44, 1, False, -1.0, -1.0, 2.0, 20, 904),
    "ref_ftc_c3_u8": ("ftc", 3, 24, 40, 3, True, -1.0, -1.0, 2.0, 16, 905),
}
PILE_KEYS = ["best_depth", "edge_conf", "edge_mask", "disp_conf", "rbar"]


def make_input(name):
    from remotesensingproject_b200.synth import make_light_field_np
    kind, S, V, U, C, as_u8, scale, dmin, dmax, D, seed = CASES[name]
    epis, _ = make_light_field_np(S, V, U, C, dmin=dmin, dmax=dmax, seed=seed, layers=5)
    if as_u8:
        return np.clip(np.rint(epis * 255.0), 0, 255).astype(np.uint8)
    if scale < 0:
        return (epis * 200.0 + 3.0).astype(np.float32)
    return epis


def run_case(name, epis, side="ref", oracle_side=False, ctx=None):
    kind, S, V, U, C, as_u8, scale, dmin, dmax, D, seed = CASES[name]
    if oracle_side:
        side = "oracle"
    if side == "ref":
        from oracle import ref
        if kind == "pile":
            r = ref.depth1d_pile(epis, dmin, dmax, D, scale_factor=scale)
            return {k: r[k] for k in PILE_KEYS}
        if kind == "2d":
            r = ref.depth2d(epis, dmin, dmax, D, scale_factor=scale)
            return {k: r[k] for k in PILE_KEYS + ["valid"]}
        r = ref.fine_to_coarse(epis, dmin, dmax, D, scale_factor=scale)
        return dict(map=r["map"], valid=r["valid"])
    if side == "oracle":
        import oracle
        if kind == "pile":
            r = oracle.depth1d_pile(oracle.normalise(epis, scale), dmin, dmax, D)
            return {k: r[k] for k in PILE_KEYS}
        if kind == "2d":
            r = oracle.depth2d(oracle.normalise(epis, scale), dmin, dmax, D)
            out = {k: r[k] for k in PILE_KEYS}
            out["valid"] = np.where(r["edge_conf"] > np.float32(0.02), 255, 0).astype(np.uint8)
            return out
        r = oracle.fine_to_coarse(epis, dmin, dmax, D, scale_factor=scale)
        return dict(map=r["map"], valid=r["valid"])
    from remotesensingproject_b200 import api
    if kind == "pile":
        c = api.Depth1DComputer_pile(epis, dmin, dmax, D, epi_scale_factor=scale, ctx=ctx).run()
        return dict(best_depth=c.m_best_depth_v_u, edge_conf=c.m_edge_confidence_v_u, edge_mask=c.m_edge_confidence_mask_v_u,
                    disp_conf=c.m_disp_confidence_v_u, rbar=c.m_rbar_v_u)
    if kind == "2d":
        c = api.Depth2DComputer(epis, dmin, dmax, D, epi_scale_factor=scale, ctx=ctx).run()
        return dict(best_depth=c.m_best_depth_s_v_u, edge_conf=c.m_edge_confidence_s_v_u,
                    edge_mask=c.m_edge_confidence_mask_s_v_u, disp_conf=c.m_disp_confidence_s_v_u, rbar=c.m_rbar_s_v_u,
                    valid=c.get_valid_depths_mask_s_v_u())
    f = api.FineToCoarse(epis, dmin, dmax, D, epi_scale_factor=scale, ctx=ctx).run()
    m, v = f.get_results()
    return dict(map=m, valid=v)


# ---- rows added after the first fixtures: options and entry points either side of the hot path --------------------------
# name -> (kind, S, V, U, C, D, seed); inputs are float32 in [0, 1] unless the kind says otherwise
LATE_CASES = {
    "ref_opening_pile_c3": ("opening", 7, 14, 60, 3, 24, 911),      # MORPH_ELLIPSE 3x3 opening of the edge mask (core.hpp:759-769)
    "ref_u16_ftc_c3": ("u16", 5, 44, 70, 3, 16, 912),               # CV_16U stack scaled by its maximum, 16-bit pyramid
    "ref_coloured_c3": ("coloured", 6, 24, 64, 3, 16, 913),         # FineToCoarse::get_coloured_depth_maps (ftc.hpp:324-377)
    "ref_single_epi_c1": ("single", 8, 1, 64, 1, 40, 914),          # Depth1DComputer on one EPI (dc.hpp:254-363)
}


def make_late_input(name):
    from remotesensingproject_b200.synth import make_light_field_np
    kind, S, V, U, C, D, seed = LATE_CASES[name]
    epis, _ = make_light_field_np(S, V, U, C, dmin=-1.0, dmax=2.0, seed=seed, layers=5, dark_fraction=0.2)
    if kind == "u16":
        return np.clip(np.rint(epis * 60000.0), 0, 65535).astype(np.uint16)
    return epis


def run_late_case(name, epis, side="ref"):
    """side "ref": the reference build; "oracle": oracle/rslf_oracle.cpp.  (The CUDA path is compared with the oracle on
    the same rows in tests/test_gpu_widening.py.)"""
    import oracle
    from oracle import ref
    kind, S, V, U, C, D, seed = LATE_CASES[name]
    here = os.path.dirname(os.path.abspath(__file__))
    if kind == "opening":
        p = oracle.default_params(edge_confidence_opening_type=2, edge_confidence_opening_size=3)
        r = ref.depth1d_pile(epis, -1.0, 2.0, D, scale_factor=1.0, params=p) if side == "ref" else \
            oracle.depth1d_pile(oracle.normalise(epis, 1.0), -1.0, 2.0, D, params=p)
        return {k: r[k] for k in PILE_KEYS}
    if kind == "u16":
        r = ref.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=-1.0, dims=oracle.pyramid_dims(V, U)) if side == "ref" else \
            oracle.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=-1.0)
        return dict(map=r["map"], valid=r["valid"])
    if kind == "coloured":
        lut = np.load(os.path.join(here, "colormap_jet.npy"))
        if side == "ref":
            return dict(bgr=ref.fine_to_coarse_coloured(epis, -1.0, 2.0, D, lut, scale_factor=1.0))
        o = oracle.fine_to_coarse(epis, -1.0, 2.0, D, scale_factor=1.0)
        return dict(bgr=oracle.colour_maps(o["map"], o["valid"], oracle.normalise(epis, 1.0), lut)[0])
    r = ref.depth1d(epis[0], -1.0, 2.0, D, scale_factor=1.0) if side == "ref" else \
        oracle.depth1d(oracle.normalise(epis, 1.0)[0], -1.0, 2.0, D)
    return {k: r[k] for k in PILE_KEYS}


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    for name in LATE_CASES:
        epis = make_late_input(name)
        out = run_late_case(name, epis, side="ref")
        path = os.path.join(here, name + ".npz")
        np.savez_compressed(path, epis=epis, **out)
        print("%s: %d bytes" % (path, os.path.getsize(path)))
    for name in CASES:
        epis = make_input(name)
        out = run_case(name, epis, side="ref")
        path = os.path.join(here, name + ".npz")
        np.savez_compressed(path, epis=epis, **out)
        print("%s: %d bytes, %d confident pixels" % (path, os.path.getsize(path), int((out.get("edge_mask", out.get("valid")) > 0).sum())))


if __name__ == "__main__":
    main()
