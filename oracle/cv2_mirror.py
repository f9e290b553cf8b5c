"""cv2 mirror of the reference's OpenCV call sequence — pins the oracle.

TEST INFRASTRUCTURE ONLY.  The reference (RSLightFields) cannot be built here
(needs OpenCV 3.x C++), so the strongest available pin is to replay, call by
call, the OpenCV operations its hot path issues, through the cv2 4.13 wheel of
this image.  `python oracle/cv2_mirror.py` regenerates tests/golden/*.npz from
seeded inputs; tests/test_oracle_vs_cv2.py checks oracle/rslf_oracle.cpp against
them (masks / indices exactly, floats to 1e-6 — cv2 4.13's multiply(a,a,scale)
uses a double intermediate where OpenCV 3.x rounds per float op).

Each function cites the reference lines it mirrors (paths relative to
/root/reference/RSLightFields).
"""
import os
import sys

import cv2
import numpy as np

cv2.setNumThreads(1)
F = np.float32


def edge_confidence_row(epi, s, filter_size=9, cut_shadows=True, shadow_level=F(0.05 * 1.73205080757),
                        thr=F(0.02)):
    """compute_1D_edge_confidence (core.hpp:426-478 + core.cpp:6-23). epi: [S][U][C] float32."""
    S, U, C = epi.shape
    row = np.ascontiguousarray(epi[s:s + 1])            # 1 x U x C
    if C == 1:
        row = row[:, :, 0]
    ce = np.zeros((1, U), F)
    centre = (filter_size - 1) // 2
    for j in range(filter_size):
        if j == centre:
            continue
        k = np.zeros((1, filter_size), F)
        k[0, centre] = 1.0
        k[0, j] = -1.0
        tmp = cv2.filter2D(row, -1, k, anchor=(-1, -1), delta=0, borderType=cv2.BORDER_REFLECT_101)
        if C == 1:
            ce += cv2.pow(tmp, 2)
        else:
            for ch in cv2.split(tmp):
                ce += cv2.pow(ch, 2)
    if cut_shadows:
        for u in range(U):
            if C == 1:
                n = F(abs(epi[s, u, 0]) * 1.73205080757)
            else:
                n = F(cv2.norm(epi[s, u].astype(F)))
            if n < shadow_level:
                ce[0, u] = 0.0
    mask = (ce > thr).astype(np.uint8) * 255
    return ce[0], mask[0]


def interpolate_mat(epi, I):
    """Interpolation1DLinear::interpolate_mat (interp.hpp:155-193)."""
    S, U, C = epi.shape
    D = I.shape[1]
    res = np.full((S, D, C), np.nan, F)
    card = np.zeros(D, F)
    for r in range(S):
        for c in range(D):
            x = I[r, c]
            i0 = int(np.floor(x))
            i1 = int(np.ceil(x))
            t = F(x - F(i0))
            if not (i0 < 0 or i1 > U - 1):
                res[r, c] = F(F(1) - t) * epi[r, i0] + t * epi[r, i1]
                card[c] += 1.0
    return res, card


def pixel_scores(epi, s_hat, u, dmin, dmax, D, h=F(0.2), iters=10, slope=F(1.0)):
    """One pixel of compute_1D_depth_epi (core.hpp:527-625) replayed with cv2 ops."""
    S, U, C = epi.shape
    Scol = np.array([[F(s_hat - s)] for s in range(S)], F)
    Dv = np.zeros((1, D), F)
    for d in range(D):
        Dv[0, d] = F(dmin) + F(F(d) * F(F(dmax) - F(dmin))) / F(D - 1)
    I = cv2.gemm(Scol, Dv, 1.0, None, 0.0)               # I = S * D
    I = (I * F(slope)).astype(F)                          # I *= slope
    I = cv2.add(I, float(u))                              # I += u
    rad, card = interpolate_mat(epi, I)
    inv = F(1.0 / float(F(h) * F(h)))
    rad_cv = rad if C == 3 else rad[:, :, 0]
    rbar = rad_cv[s_hat:s_hat + 1].copy()
    rad0 = cv2.max(rad_cv, 0.0)
    K = None
    for _ in range(iters):
        rb = cv2.repeat(rbar, S, 1)
        rm = cv2.subtract(rad_cv, rb)
        if C == 1:
            K = cv2.multiply(rm, rm, scale=float(3 * inv))
        else:
            sq = cv2.multiply(rm, rm, scale=float(inv))
            K = cv2.reduce(sq.reshape(S * D, 3), 1, cv2.REDUCE_SUM).reshape(S, D)
        K = cv2.subtract(1.0, K)
        K = cv2.max(K, 0.0)
        Kv = K if C == 1 else cv2.merge([K, K, K])
        rK = cv2.multiply(rad0, Kv)
        s_rK = cv2.reduce(rK, 0, cv2.REDUCE_SUM)
        s_K = cv2.reduce(K, 0, cv2.REDUCE_SUM)
        s_Kv = s_K if C == 1 else cv2.merge([s_K, s_K, s_K])
        with np.errstate(all="ignore"):
            rbar = cv2.divide(s_rK, s_Kv)
        rbar = cv2.max(rbar, 0.0)
    s_K = cv2.reduce(K, 0, cv2.REDUCE_SUM)
    score = cv2.divide(s_K, card.reshape(1, D))
    score = cv2.max(score, 0.0)
    _, maxVal, _, maxIdx = cv2.minMaxLoc(score)
    mean = cv2.mean(score)[0]
    return dict(score=score[0], rbar=rbar.reshape(D, C), dvals=Dv[0], best=maxIdx[0], max=maxVal, mean=mean)


def downsample(raw):
    """downsample_EPIs (ftc_core.cpp:14-60). raw: [V][S][U][C]."""
    V, S, U, C = raw.shape
    imgs = []
    for s in range(S):
        img = np.ascontiguousarray(raw[:, s])
        if C == 1:
            img = img[:, :, 0]
        b = cv2.GaussianBlur(img, (7, 7), 0, 0, borderType=cv2.BORDER_REFLECT)
        r = cv2.resize(b, None, fx=0.5, fy=0.5, interpolation=cv2.INTER_LINEAR)
        imgs.append(r.reshape(r.shape[0], r.shape[1], C))
    return np.ascontiguousarray(np.stack(imgs, 1))


def fuse(disp_p, valid_p):
    """fuse_disp_maps (ftc_core.cpp:69-135)."""
    S = disp_p[0].shape[0]
    L = len(disp_p)
    out_map, out_valid = [], []
    for s in range(S):
        md = disp_p[L - 1][s]
        mk = valid_p[L - 1][s]
        for p in range(L - 1, 0, -1):
            size = (disp_p[p - 1].shape[2], disp_p[p - 1].shape[1])
            up = cv2.resize(md, size, interpolation=cv2.INTER_LINEAR)
            mup = cv2.resize(mk, size, interpolation=cv2.INTER_NEAREST)
            md = disp_p[p - 1][s].copy()
            inval = (valid_p[p - 1][s] == 0)
            md[inval] = 0.0
            md = cv2.add(md, up, dst=md, mask=inval.astype(np.uint8) * 255)
            mk = cv2.bitwise_or(valid_p[p - 1][s], mup)
        out_map.append(cv2.medianBlur(md, 3))
        out_valid.append(mk)
    return np.stack(out_map), np.stack(out_valid)


def _synthetic_epi(rng, S, U, C, d_true):
    """A small EPI with two slanted textures so that scores are informative."""
    tex = rng.random((U + 64, C)).astype(F)
    k = np.array([0.25, 0.5, 0.25], F)
    for _ in range(2):
        tex = np.stack([np.convolve(tex[:, c], k, mode="same") for c in range(C)], 1).astype(F)
    s_hat = S // 2
    epi = np.zeros((S, U, C), F)
    for s in range(S):
        x = np.arange(U) - (s_hat - s) * d_true + 32
        i0 = np.floor(x).astype(int)
        t = (x - i0).astype(F)[:, None]
        epi[s] = tex[i0] * (1 - t) + tex[i0 + 1] * t
    epi[:, U // 2:, :] *= F(0.5)
    epi[:, :4, :] = F(0.01)            # dark pixels -> shadow cut
    return epi


def generate(outdir):
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(4242)
    cases = {}
    # --- per-pixel scores + edge confidence, C in {1,3}, even/odd S ---------
    for name, (S, U, C, D, dmin, dmax, d_true) in {
        "px_c3_s9": (9, 48, 3, 12, -1.0, 4.0, 1.3),
        "px_c1_s8": (8, 40, 1, 16, -2.0, 2.0, -0.7),
        "px_c3_s5_eq": (5, 33, 3, 6, 0.5, 0.5, 0.5),       # dmin == dmax
    }.items():
        epi = _synthetic_epi(rng, S, U, C, d_true)
        s_hat = S // 2
        ce, mask = edge_confidence_row(epi, s_hat)
        us = [0, 1, 5, U // 3, U // 2, U - 2, U - 1]
        sc = [pixel_scores(epi, s_hat, u, dmin, dmax, D) for u in us]
        cases[name] = dict(
            epi=epi, s_hat=s_hat, D=D, dmin=F(dmin), dmax=F(dmax), us=np.array(us), ce=ce, mask=mask,
            score=np.stack([x["score"] for x in sc]), rbar=np.stack([x["rbar"] for x in sc]),
            dvals=np.stack([x["dvals"] for x in sc]), best=np.array([x["best"] for x in sc]),
            maxv=np.array([x["max"] for x in sc]), mean=np.array([x["mean"] for x in sc]))
    for name, c in cases.items():
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **c)
    # --- pyramid ops -----------------------------------------------------------
    for name, (V, S, U, C) in {"down_c3_odd": (23, 3, 37, 3), "down_c1_even": (24, 2, 30, 1),
                               "down_c1_135": (27, 2, 135, 1)}.items():
        raw = (rng.random((V, S, U, C)).astype(F) * F(250.0) + F(8.0))
        np.savez_compressed(os.path.join(outdir, name + ".npz"), raw=raw, out=downsample(raw))
    dims = [(23, 37), (12, 18), (6, 9)]
    S = 2
    disp_p = [rng.uniform(-1, 4, (S, v, u)).astype(F) for v, u in dims]
    valid_p = [((rng.random((S, v, u)) > 0.5).astype(np.uint8) * 255) for v, u in dims]
    valid_p[-1][:] = 255
    m, k = fuse(disp_p, valid_p)
    np.savez_compressed(os.path.join(outdir, "fuse_3lvl.npz"), d0=disp_p[0], d1=disp_p[1], d2=disp_p[2],
                        v0=valid_p[0], v1=valid_p[1], v2=valid_p[2], out_map=m, out_valid=k)
    # 8-bit stacks take OpenCV's integer blur / resize paths (own generator so the fixtures above keep their stream)
    rng8 = np.random.default_rng(777)
    for name, (V, S, U, C) in {"down_u8_c3_odd": (23, 2, 37, 3), "down_u8_c1_135": (27, 2, 135, 1),
                               "down_u8_c1_even": (24, 2, 30, 1)}.items():
        raw = rng8.integers(0, 256, (V, S, U, C), dtype=np.uint8)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), raw=raw, out=downsample(raw))
    print("wrote fixtures to", outdir)


if __name__ == "__main__":
    generate(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
