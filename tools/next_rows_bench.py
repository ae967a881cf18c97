#!/usr/bin/env python
"""Timings of the SURVEY §8f rows N1 (local-map assembly), N2 (adjustDistortion) and N4 (loop-closure ICP) on the GPU box, next to the CPU restatement
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
ONLY = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else "all"  # icp | distortion | assemble | all
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
if ONLY in ("all", "assemble"):
    # ---- N1: a 50-keyframe window of 64 x 1800 sweeps (laserMapping.cpp:206-243, 316-319): every keyframe = the three downsampled
    #      clouds LaserMapping stores for a frame (:503-510) and its pose; the window is transformed, concatenated and voxelised ----
    P64 = alego.default_params(alego.PRESET_HDL64_1800)
    seed = 3
    wd = alego.SynthWorld(seed=seed)
    g3 = alego.Alego(P64, n_seq=1)
    empty = np.zeros((0, 4), np.float32)
    g3.lm_set_map(0, empty, empty)
    g3.pipeline_config(lm_every=1)
    K = 8 if QUICK else 50
    ck, sk, ok_, poses6 = [], [], [], []
    for t in range(K):
        scan = wd.render(P64, alego.trajectory_pose(t, seed=seed), noise_seed=700 + t)
        buf, n = g3.pack_scans([scan])
        g3.pipeline_step(buf, n)
        c, s_, o_ = g3.lm_get_downsampled(0)
        ck.append(c); sk.append(s_); ok_.append(o_)
        x, y, z, yaw = alego.trajectory_pose(t, seed=seed)
        poses6.append(np.array([x, y, z, 0.0, 0.0, yaw], np.float32))  # ground-truth key poses: a consistent window
    poses6 = np.stack(poses6)
    g3.lm_assemble_map(0, ck, sk, ok_, poses6)  # scratch growth + warm-up
    g3.synchronize()
    g3.profile_enable(True)
    g3.profile_reset()
    ts = []
    for _ in range(1 if QUICK else 5):
        t0 = time.perf_counter()
        g3.lm_assemble_map(0, ck, sk, ok_, poses6)
        g3.synchronize()
        ts.append(time.perf_counter() - t0)
    prof = g3.profile()
    g3.profile_enable(False)
    gc, gs = g3.lm_get_map(0)
    t0 = time.perf_counter()
    cm, sm, _ = ob.lm_assemble_map(ck, sk, ok_, poses6, P64.lm_corner_leaf, P64.lm_surf_leaf, stable=False)
    t_cpu = time.perf_counter() - t0
    out["assemble_map"] = {
        "keyframes": K, "corner_points_in": int(sum(len(c) for c in ck)), "surf_points_in": int(sum(len(a) + len(b) for a, b in zip(sk, ok_))),
        "corner_points_out": len(gc), "surf_points_out": len(gs),
        "gpu_call_ms_median": round(1e3 * float(np.median(ts)), 3),
        "gpu_kernels_ms_per_call": {k: round(v[1] / max(len(ts), 1), 4) for k, v in prof.items()},
        "cpu_oracle_ms": round(1e3 * t_cpu, 1),
        "bit_exact_vs_cpu": bool(np.array_equal(gc, cm) and np.array_equal(gs, sm)),
        "note": "gpu_call = host clouds (pageable) in, H2D of the window, transform, VoxelGrid in pcl's record order, map left on the "
                "device; runs once per saved keyframe (laserMapping.cpp:491-545), not per sweep",
    }
print(json.dumps(out, indent=1))
