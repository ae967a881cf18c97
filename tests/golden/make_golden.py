#!/usr/bin/env python
"""Generates tests/golden/*.npz — golden input/output vectors on seeded synthetic inputs.  The reference holds no golden
vectors of its own, so they are produced HERE from the reference's own code: every array the reference build
(oracle/_ref: src/imageProjection.cpp, laserOdometry.cpp, laserMapping.cpp compiled unmodified against stand-in third-party
headers) exposes is asserted equal to the port oracle's (oracle/alego_oracle.cpp) before it is written — bit-exact for the
integer / float32 arrays, 1e-9 for the double solver state; what the reference computes but never materialises (feature
INDEX lists, correspondence index lists) is taken from the port, whose gathered clouds are asserted equal to the clouds the
reference publishes.  PCL's VoxelGrid record order (std::sort on the voxel index alone) is the literal one everywhere.
tests/test_golden.py checks (CPU) that the oracle still reproduces the files and (GPU) that the CUDA path reproduces them
through the C ABI.  Needs /root/reference (for oracle/_ref); run in the build container only.

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
             "less_flat"]
LO_KEYS = ["lo_params", "t_w_cur", "r_w_cur", "lo_surf_corr", "lo_corner_corr", "lo_trace"]
LM_KEYS = ["lm_params", "t_map2laser", "t_map2odom", "r_map2odom", "lm_corner_sel", "lm_surf_sel", "lm_trace", "lm_corner_ds",
           "lm_surf_total_ds"]


def main():
    alego = alego_pkg.load()
    from oracle import binding as ob, ref_binding as rb
    P = alego.default_params(alego.PRESET_VLP16_1800)
    V = "vlp16_1800"

    def same(a, b, what, tol=0.0):
        a, b = np.asarray(a), np.asarray(b)
        ok = a.shape == b.shape and (np.array_equal(a, b) if tol == 0.0 else np.abs(a - b).max() < tol)
        assert ok, "reference build and port disagree on " + what

    seed = 0
    w = alego.SynthWorld(seed=seed)
    scans = [w.render(P, alego.trajectory_pose(t, speed=0.25, yaw_rate=0.02, seed=seed), noise_seed=900 + t) for t in range(3)]

    # ---- cfg1: one scan, IP + features
    o = ob.Oracle(P, stable_voxel=False)
    assert o.ip(scans[0]) == 0
    o.lo_features()
    rip, rlo = rb.RefImageProjection(V), rb.RefLaserOdometry(V)
    assert rip.process(scans[0]) == 0 and rlo.process(rip) == 0
    for k in IP_KEYS:
        same(np.asarray(rip.get(k)).ravel(), np.asarray(o.get(k)).ravel(), k)
    Mr = len(rip.get("segmentedCloudColInd"))
    for k in ("cloud_curvature", "cloud_neighbor_picked", "cloud_label", "cloud_sort_idx"):
        same(rlo.get(k)[5:Mr - 5], o.get(k)[5:Mr - 5], k)
    seg = o.get("segmented_cloud")
    for cloud, idx in (("sharp", "sharp_idx"), ("less_sharp", "less_sharp_idx"), ("flat", "flat_idx")):
        same(rlo.get(cloud), seg[o.get(idx)].reshape(-1, 4), cloud)
    same(rlo.get("less_flat"), o.get("less_flat"), "less_flat")
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
        o = ob.Oracle(Q, lm_every=0, stable_voxel=False)
        rip, rlo = rb.RefImageProjection(V), rb.RefLaserOdometry(V)
        for t, s in enumerate(scans):
            o.ip(s)
            o.lo_features()
            o.lo_scan2scan()
            rep = o.report("lo")
            if ci == 5:  # the reference's own iteration counts (laserOdometry.cpp:415,489); 5+10 is the README's variant, port only
                assert rip.process(s) == 0 and rlo.process(rip) == 0
                same(rlo.get("lo_params"), o.get("lo_params"), "lo_params", 1e-9)
                same(rlo.get("t_w_cur"), o.get("t_w_cur"), "t_w_cur", 1e-9)
                if t > 0:
                    same(rlo.get("lo_trace"), o.get("lo_trace"), "lo_trace", 1e-9)
                    assert int(rlo.get("lo_solve_iterations").sum()) == rep["iterations"]
            if t > 0:
                for k in LO_KEYS:
                    out["%s_t%d_%s" % (tag, t, k)] = np.asarray(o.get(k))
                out["%s_t%d_report" % (tag, t)] = np.array([rep["n_corner"], rep["n_surf"], rep["iterations"]], np.int32)
    np.savez_compressed(os.path.join(HERE, "cfg2_vlp16_scan2scan.npz"), **out)

    # ---- cfg3 (small): scan-to-map, features of sweep 0 from the oracle front end, 4k corner + 20k surf map
    corner_map, surf_map = w.make_map(4000, 20000, seed=seed, radius=60.0)
    o = ob.Oracle(P, stable_voxel=False)
    o.ip(scans[0])
    o.lo_features()
    corner, surf, outl = o.get("less_sharp"), o.get("less_flat"), o.get("outlier_cloud")
    yaw = np.deg2rad(1.0)
    R0 = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1.0]])
    t0 = np.array([0.15, -0.12, 0.05])
    x0 = np.array([0.15, -0.12, 0.05, 0.0, 0.0, yaw])
    out = {"corner_map": corner_map, "surf_map": surf_map, "corner": corner, "surf": surf, "outlier": outl, "t_odom": t0, "r_odom": R0, "x0": x0}
    for tag, iters in (("2x20", (2, 20)), ("1x10", (1, 10))):   # code default and BASELINE cfg3's "10 LM iters"
        Q = P.copy()
        Q.lm_outer_iters, Q.lm_max_iters = iters
        o2 = ob.Oracle(Q, stable_voxel=False)
        o2.lm_set_map(corner_map, surf_map)
        o2.lm_set_scan(corner, surf, outl)
        o2.lm_set_odom(t0, R0)
        o2.lm_set_params(x0)
        o2.lm_scan2map()
        rep = o2.report("lm")
        if iters == (2, 20):  # the reference's own loop bounds (laserMapping.cpp:360,470)
            rlm = rb.RefLaserMapping(V)
            assert rlm.scan2map(corner_map, surf_map, corner, surf, outl, t0, [np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)], x0) == 0
            same(rlm.get("lm_params"), o2.get("lm_params"), "lm_params", 1e-9)
            same(rlm.get("lm_trace"), o2.get("lm_trace"), "lm_trace", 1e-9)
            same(rlm.get("lm_corner_ds"), o2.get("lm_corner_ds"), "lm_corner_ds")
            same(rlm.get("lm_surf_total_ds"), o2.get("lm_surf_total_ds"), "lm_surf_total_ds")
            assert int(rlm.get("lm_solve_iterations").sum()) == rep["iterations"]
        for k in LM_KEYS:
            out["%s_%s" % (tag, k)] = np.asarray(o2.get(k))
        out["%s_report" % tag] = np.array([rep["n_corner"], rep["n_surf"], rep["iterations"]], np.int32)
    np.savez_compressed(os.path.join(HERE, "cfg3_small_scan2map.npz"), **out)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
