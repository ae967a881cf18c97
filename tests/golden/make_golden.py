#!/usr/bin/env python
"""Generates tests/golden/*.npz — golden input/output vectors of the ORACLE (oracle/alego_oracle.cpp) on seeded
synthetic inputs.  The reference itself has no golden vectors and cannot run here (parity unpinned), so
these files pin the oracle: tests/test_golden.py checks (CPU) that the oracle still reproduces them and (GPU) that the
CUDA path reproduces them through the C ABI.

    python tests/golden/make_golden.py        # rewrites the fixtures (only when the oracle changes on purpose)

BASELINE.json configs covered: cfg1 (single 16x1800 scan: IP + features), cfg2 (LaserOdometry scan-to-scan on 16x1800
consecutive sweeps, 5+5 and README's 5+10 iterations), cfg3-shaped small case (LaserMapping scan-to-map against a local
map; the full 50k+200k map is generated, not stored — tests/test_gpu_parity.py covers it against the live oracle).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import alego_pkg  # noqa: E402

IP_KEYS = ["range_mat", "ground_mat", "label_mat", "startRingIndex", "endRingIndex", "segmentedCloudGroundFlag",
           "segmentedCloudColInd", "segmentedCloudRange", "outlier_cloud"]
FEAT_KEYS = ["cloud_curvature", "cloud_neighbor_picked", "cloud_label", "cloud_sort_idx", "sharp_idx", "less_sharp_idx", "flat_idx",
             "less_flat_stable"]
LO_KEYS = ["lo_params", "t_w_cur", "r_w_cur", "lo_surf_corr", "lo_corner_corr", "lo_trace"]
LM_KEYS = ["lm_params", "t_map2laser", "t_map2odom", "r_map2odom", "lm_corner_sel", "lm_surf_sel", "lm_trace", "lm_corner_ds",
           "lm_surf_total_ds"]


def main():
    alego = alego_pkg.load()
    from oracle import binding as ob
    P = alego.default_params(alego.PRESET_VLP16_1800)
    seed = 0
    w = alego.SynthWorld(seed=seed)
    scans = [w.render(P, alego.trajectory_pose(t, speed=0.25, yaw_rate=0.02, seed=seed), noise_seed=900 + t) for t in range(3)]

    # ---- cfg1: one scan, IP + features
    o = ob.Oracle(P, stable_voxel=True)
    assert o.ip(scans[0]) == 0
    o.lo_features()
    out = {"scan": scans[0]}
    for k in IP_KEYS + FEAT_KEYS:
        out[k] = np.asarray(o.get(k))
    M = len(out["segmentedCloudColInd"])
    for k in ("cloud_curvature", "cloud_neighbor_picked", "cloud_label", "cloud_sort_idx"):
        out[k] = out[k][5:M - 5]   # only [5, M-5) is defined by the reference (laserOdometry.cpp:122-129)
    out["range_mat"] = np.where(out["range_mat"] == np.finfo(np.float64).max, np.finfo(np.float32).max, out["range_mat"]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "cfg1_vlp16_ip_features.npz"), **out)

    # ---- cfg2: scan-to-scan over the three consecutive sweeps, code default 5+5 and README's 5+10
    out = {"scan1": scans[1], "scan2": scans[2]}
    for tag, ci in (("5_5", 5), ("5_10", 10)):
        Q = P.copy()
        Q.lo_corner_iters = ci
        o = ob.Oracle(Q, lm_every=0, stable_voxel=True)
        for t, s in enumerate(scans):
            o.ip(s)
            o.lo_features()
            o.lo_scan2scan()
            rep = o.report("lo")
            if t > 0:
                for k in LO_KEYS:
                    out["%s_t%d_%s" % (tag, t, k)] = np.asarray(o.get(k))
                out["%s_t%d_report" % (tag, t)] = np.array([rep["n_corner"], rep["n_surf"], rep["iterations"]], np.int32)
    np.savez_compressed(os.path.join(HERE, "cfg2_vlp16_scan2scan.npz"), **out)

    # ---- cfg3 (small): scan-to-map, features of sweep 0 from the oracle front end, 4k corner + 20k surf map
    corner_map, surf_map = w.make_map(4000, 20000, seed=seed, radius=60.0)
    o = ob.Oracle(P, stable_voxel=True)
    o.ip(scans[0])
    o.lo_features()
    corner, surf, outl = o.get("less_sharp"), o.get("less_flat_stable"), o.get("outlier_cloud")
    yaw = np.deg2rad(1.0)
    R0 = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
    t0 = np.array([0.15, -0.12, 0.05])
    x0 = np.array([0.15, -0.12, 0.05, 0.0, 0.0, yaw])
    out = {"corner_map": corner_map, "surf_map": surf_map, "corner": corner, "surf": surf, "outlier": outl, "t_odom": t0, "r_odom": R0, "x0": x0}
    for tag, iters in (("2x20", (2, 20)), ("1x10", (1, 10))):   # code default and BASELINE cfg3's "10 LM iters"
        Q = P.copy()
        Q.lm_outer_iters, Q.lm_max_iters = iters
        o2 = ob.Oracle(Q, stable_voxel=True)
        o2.lm_set_map(corner_map, surf_map)
        o2.lm_set_scan(corner, surf, outl)
        o2.lm_set_odom(t0, R0)
        o2.lm_set_params(x0)
        o2.lm_scan2map()
        rep = o2.report("lm")
        for k in LM_KEYS:
            out["%s_%s" % (tag, k)] = np.asarray(o2.get(k))
        out["%s_report" % tag] = np.array([rep["n_corner"], rep["n_surf"], rep["iterations"]], np.int32)
    np.savez_compressed(os.path.join(HERE, "cfg3_small_scan2map.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
