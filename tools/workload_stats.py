#!/usr/bin/env python
"""Sizes of the intermediate clouds of the bench workload (hdl64_1800, 50k+200k map) — run on the GPU box."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alego_pkg
alego = alego_pkg.load()
P = alego.default_params(1)
seed = 100
w = alego.SynthWorld(seed=seed)
corner, surf = w.make_map(50000, 200000, seed=seed, radius=100.0)
g = alego.Alego(P, n_seq=1)
g.pipeline_config(lm_every=1)
g.lm_set_map(0, corner, surf)
for t in range(4):
    scan = w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=1000 * seed + t)
    buf, n = g.pack_scans([scan])
    g.pipeline_step(buf, n)
    M = len(g.debug("segmentedCloudColInd"))
    sr, er = g.debug("startRingIndex"), g.debug("endRingIndex")
    ring_n = (er - sr + 11)
    stats = {k: len(g.debug(k)) for k in ("sharp_idx", "less_sharp_idx", "flat_idx", "less_flat", "outlier_cloud", "lm_corner_ds", "lm_surf_ds", "lm_outlier_ds", "lm_surf_total_ds")}
    edge = g.debug("lm_edge").reshape(-1, 10)
    plane = g.debug("lm_plane").reshape(-1, 8)
    print("sweep", t, "points", len(scan), "M", M, "ring max/mean", ring_n.max(), int(ring_n.mean()), stats,
          "edge corr", int((edge[:, 0] != 0).sum()), "plane corr", int((plane[:, 0] != 0).sum()),
          "lm_trace", len(g.debug("lm_trace")) // 7, "lo_trace", len(g.debug("lo_trace")) // 7)
