#!/usr/bin/env python
"""Histogram of the mean-shift iteration at which r_bar reaches a bitwise fixed point (oracle statistics), on a band
of a bench configuration: how many of the ten iterations an exact early exit saves, per hypothesis and per warp
(32 consecutive hypotheses).  Test/analysis tooling only (uses the oracle)."""
import argparse, ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c3")
ap.add_argument("--rows", type=int, default=8)
a = ap.parse_args()
cfg = bench.CONFIGS[a.config]
epis = bench.cpu_sample(cfg, a.rows)
L = oracle.lib()
L.orc_ms_stats_enable(1)
samples, secs = bench.run_cpu(cfg, epis, "port")
lane = (C.c_longlong * 33)(); warp = (C.c_longlong * 33)(); px = C.c_longlong(); flat = C.c_longlong()
L.orc_ms_stats_read(lane, warp, C.byref(px), C.byref(flat))
L.orc_ms_stats_enable(0)
lane = np.array(lane[:11], dtype=np.float64); warp = np.array(warp[:11], dtype=np.float64)
it = np.arange(11)
print(json.dumps({"config": a.config, "rows": a.rows, "pixels": px.value, "flat_pixels": flat.value, "cpu_s": round(secs, 1),
                  "lane_hist": (lane / lane.sum()).round(4).tolist(), "warp_hist": (warp / warp.sum()).round(4).tolist(),
                  "mean_iters_lane": float((lane * it).sum() / lane.sum()),
                  "mean_iters_warp": float((warp * it).sum() / warp.sum())}))
