#!/usr/bin/env python
"""Timings of the SURVEY §8f rows N2 (adjustDistortion) and N4 (loop-closure ICP) on the GPU box, next to the CPU restatement
(oracle, one core).  Per-kernel times come from the library's CUDA-event profile (alego_profile_*), the call times are host
wall clock around the C-ABI call (host clouds in, result out).  Prints one JSON object."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import alego_pkg
from oracle import binding as ob
import test_next_rows as T

alego = alego_pkg.load()
out = {}
ONLY = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else "all"  # icp | distortion | all
QUICK = "--quick" in sys.argv  # one repetition, small batch: for runs under ncu

P = alego.default_params(1)
rng = np.random.default_rng(8)
if ONLY in ("all", "icp"):
    # ---- N4: latest keyframe against a history cloud as detectLoopClosure builds it: VoxelGrid(1.0) of the neighbouring keyframes
    #      (laserMapping.cpp:41,811) — here ~60 k voxels of a 240 m x 240 m scene; source = 0.4 m voxels within 60 m ----------------
    dense = T.icp_scene(rng, 400000, ext=120.0)
    tgt, _ = ob.voxel_grid(dense, 1.0)
    near = dense[(np.abs(dense[:, 0]) < 60) & (np.abs(dense[:, 1]) < 60)]
    src, _ = ob.voxel_grid(near[::2], 0.4)
    src, _ = T.misalign(src, [0.02, -0.005, 0.004], np.array([0.3, -0.2, 0.05]))
    g = alego.Alego(P, n_seq=1)
    g.lc_icp(src, tgt)  # allocation + warm-up
    g.profile_enable(True)
    g.profile_reset()
    ts = []
    for _ in range(1 if QUICK else 5):
        t0 = time.perf_counter()
        r = g.lc_icp(src, tgt)
        ts.append(time.perf_counter() - t0)
    prof = g.profile()
    g.profile_enable(False)
    t0 = time.perf_counter()
    w = ob.icp(src, tgt, exact_sums=False)
    t_cpu = time.perf_counter() - t0
    n_it, ms_it = prof.get("icp_iterate", (0, 0.0))
    out["icp"] = {
        "source_points": len(src), "target_points": len(tgt), "iterations": r["iterations"], "state": r["state"],
        "gpu_call_ms_median": round(1e3 * float(np.median(ts)), 3),
        "gpu_iterate_kernel_ms_avg": round(ms_it / max(n_it, 1), 4), "gpu_iterate_launches": n_it,
        "gpu_kernels_ms": {k: round(v[1] / max(v[0], 1), 4) for k, v in prof.items()},
        "cpu_oracle_ms": round(1e3 * t_cpu, 1), "cpu_iterations": w["iterations"],
        "translation_diff_m": float(np.abs(r["T"][:3, 3] - w["T"][:3, 3]).max()),
    }

if ONLY in ("all", "distortion"):
    # ---- N2: 64 x 1800 sweeps, 32 sequences ------------------------------------------------------------------------------------
    B = 8 if QUICK else 32
    scans = []
    for s in range(B):
        wd = alego.SynthWorld(seed=s % 8)
        sc = wd.render(P, alego.trajectory_pose(s // 8, seed=s % 8), noise_seed=500 + s)
        h = np.degrees(-np.arctan2(sc[:, 1], sc[:, 0]) + 2 * np.pi) % 360.0
        scans.append(sc[h < 300.0])  # a 300-degree sweep: the whole cloud is visited (tests/test_next_rows.py explains)
    g2 = alego.Alego(P, n_seq=B)
    buf, n = g2.pack_scans(scans)
    qa = [T.make_queue(rng, 10.0 + b) for b in range(B)]
    t_scan = np.array([10.0 + b for b in range(B)])
    g2.ip_process(buf, n)
    M = [len(g2.ip_get(b, labels=False)["segmented_cloud"]) for b in range(B)]
    g2.lo_adjust_distortion(t_scan, [x[0] for x in qa], [x[1] for x in qa], [x[2] for x in qa])
    g2.profile_enable(True)
    g2.profile_reset()
    for _ in range(1 if QUICK else 5):
        g2.ip_process(buf, n)
        n_adj, _ = g2.lo_adjust_distortion(t_scan, [x[0] for x in qa], [x[1] for x in qa], [x[2] for x in qa])
    prof = g2.profile()
    info = g2.ip_get(0, labels=False)
    g2.ip_process(buf, n)
    info = g2.ip_get(0, labels=False)
    t0 = time.perf_counter()
    ob.adjust_distortion(info["segmented_cloud"], info["segmentedCloudColInd"], info["startOrientation"], info["endOrientation"],
                         P.horizon_scan, t_scan[0], qa[0][0], qa[0][1], qa[0][2])
    t_cpu = time.perf_counter() - t0
    k_n, k_ms = prof.get("lo_adjust_apply", (0, 0.0))
    out["adjust_distortion"] = {
        "n_seq": B, "points_per_sequence": int(np.mean(M)), "points_adjusted_per_sequence": int(np.mean(n_adj)),
        "gpu_apply_kernel_ms_per_launch": round(k_ms / max(k_n, 1), 4),
        "gpu_walk_kernel_ms_per_launch": round(prof.get("lo_adjust_walk", (0, 0.0))[1] / max(prof.get("lo_adjust_walk", (1, 0.0))[0], 1), 4),
        "algorithmic_GBps": round(sum(M) * (16 + 16 + 4) / max(k_ms / max(k_n, 1), 1e-9) / 1e6, 1),
        "cpu_oracle_ms_one_sequence": round(1e3 * t_cpu, 3),
    }
print(json.dumps(out, indent=1))
