"""Synthetic 3D light fields (layered fronto-parallel scene) for tests and bench.

There is no network for datasets, so every workload of BASELINE.json is rendered
from this seeded generator.  A scene is L layers with
disparities in [dmin+0.25, dmax-0.25]; view s shows the layer point
x = u - (s_hat - s) * d_l, i.e. the reference's EPI line convention
u' = u + (s_hat - s) * d (rslf_depth_computation_core.hpp:542-552).  The nearest
layer (largest d) wins; textures are interpolated linearly.

Textures come from numpy's default_rng (host, deterministic); the per-view
rendering runs in torch on the requested device so the big configurations can
be produced on the GPU in seconds.
"""
import math

import numpy as np
import torch


def _blur1d(x, sigma, dim):
    r = max(1, int(math.ceil(3 * sigma)))
    k = torch.arange(-r, r + 1, dtype=torch.float32)
    k = torch.exp(-0.5 * (k / sigma) ** 2)
    k = (k / k.sum()).to(x.device)
    x = x.movedim(dim, -1)
    shp = x.shape
    y = torch.nn.functional.pad(x.reshape(-1, 1, shp[-1]), (r, r), mode="replicate")
    y = torch.nn.functional.conv1d(y, k.view(1, 1, -1))
    return y.reshape(shp).movedim(-1, dim)


def make_light_field(S, V, U, C, dmin=-1.0, dmax=4.0, seed=20130, layers=12, rect_density=1.0,
                     dark_fraction=0.05, value_range=(0.0, 1.0), device="cpu", dtype=torch.float32,
                     layout="vsuc"):
    """Returns (epis, layer_disparities).

    epis: torch tensor [V][S][U][C] (layout="vsuc", the reference's Vec<Mat> epis,
    rslf_io.cpp:194-227) or [S][V][U][C] (layout="svuc", the image stack) with
    values in value_range.  rect_density scales the number of piecewise-constant
    rectangles per layer, i.e. the edge density (a declared workload parameter).
    """
    rng = np.random.default_rng(seed)
    dev = torch.device(device)
    s_hat = S // 2
    reach = max(abs(dmin), abs(dmax)) * max(s_hat, S - 1 - s_hat)
    pad = int(math.ceil(reach)) + 3
    W = U + 2 * pad
    disps = np.sort(rng.uniform(dmin + 0.25, dmax - 0.25, size=layers)).astype(np.float32)
    tex = []
    sup = []
    for l in range(layers):
        t = torch.from_numpy(rng.random((V, W, C), dtype=np.float32)).to(dev)
        t = _blur1d(_blur1d(t, 1.5, 0), 1.5, 1)
        t = (t - 0.5) * 2.5 + 0.5          # restore contrast lost to the low-pass
        n_rect = max(1, int(rect_density * V * W / 1500))
        hs = rng.integers(8, 65, size=n_rect)
        ws = rng.integers(8, 65, size=n_rect)
        v0 = rng.integers(0, max(1, V), size=n_rect)
        u0 = rng.integers(0, max(1, W), size=n_rect)
        cols = rng.random((n_rect, C), dtype=np.float32)
        for i in range(n_rect):
            t[v0[i]:v0[i] + hs[i], u0[i]:u0[i] + ws[i], :] = (
                0.5 * t[v0[i]:v0[i] + hs[i], u0[i]:u0[i] + ws[i], :] + 0.5 * torch.from_numpy(cols[i]).to(dev))
        t = t.clamp_(0.0, 1.0)
        # dark patches to exercise the shadow cut (core.hpp:464-474)
        n_dark = max(1, int(dark_fraction * V * W / (48 * 48)))
        for i in range(n_dark):
            a = int(rng.integers(0, max(1, V)))
            b = int(rng.integers(0, max(1, W)))
            t[a:a + 48, b:b + 48, :] *= 0.03
        m = torch.zeros((V, W), dtype=torch.float32, device=dev)
        if l == 0:
            m[:] = 1.0
        else:
            n_sup = max(1, int(V * W / 30000))
            for i in range(n_sup):
                a = int(rng.integers(0, max(1, V)))
                b = int(rng.integers(0, max(1, W)))
                hh = int(rng.integers(max(8, V // 12), max(9, V // 3)))
                ww = int(rng.integers(max(8, U // 12), max(9, U // 3)))
                m[a:a + hh, b:b + ww] = 1.0
        tex.append(t)
        sup.append(m)
    lo, hi = value_range
    out = torch.empty((S, V, U, C), dtype=dtype, device=dev)
    uu = torch.arange(U, dtype=torch.float32, device=dev)
    for s in range(S):
        img = torch.zeros((V, U, C), dtype=torch.float32, device=dev)
        for l in range(layers):           # far -> near, nearer layers overwrite
            x = uu - (s_hat - s) * float(disps[l]) + pad
            x0 = torch.floor(x)
            t = (x - x0).view(1, U, 1)
            i0 = x0.long().clamp_(0, W - 2)
            a = tex[l][:, i0, :]
            b = tex[l][:, i0 + 1, :]
            val = a * (1 - t) + b * t
            ms = sup[l][:, torch.round(x).long().clamp_(0, W - 1)]
            img = torch.where(ms.unsqueeze(-1) > 0.5, val, img)
        out[s] = (img * (hi - lo) + lo).to(dtype)
    if layout == "svuc":
        return out, disps
    return out.permute(1, 0, 2, 3).contiguous(), disps


def make_light_field_np(*args, **kw):
    """numpy float32 [V][S][U][C] version (host)."""
    kw.setdefault("device", "cpu")
    epis, disps = make_light_field(*args, **kw)
    return epis.numpy(), disps
