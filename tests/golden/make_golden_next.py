#!/usr/bin/env python
"""Generates tests/golden/n2_adjust_distortion.npz and n4_loop_closure_icp.npz — golden vectors of the ORACLE's restatements of
LaserOdometry::adjustDistortion (laserOdometry.cpp:557-657) and of the ICP of performLoopClosure (row N4,
laserMapping.cpp:667-688) on seeded inputs.  Inputs and outputs are both stored, so the files pin the oracle and serve the GPU
path as fixtures.  Separate from make_golden.py so that the older fixtures stay byte-identical.

    python tests/golden/make_golden_next.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import alego_pkg
    from oracle import binding as ob
    import test_next_rows as T
    alego = alego_pkg.load()
    # ---- N2: a 300-degree VLP-16 sweep (every ring is visited) and a wrapped IMU queue
    P = alego.default_params(alego.PRESET_VLP16_1800)
    w = alego.SynthWorld(seed=9)
    scan = w.render(P, alego.trajectory_pose(0, seed=9), noise_seed=909)
    h = np.degrees(-np.arctan2(scan[:, 1], scan[:, 0]) + 2 * np.pi) % 360.0
    scan = np.ascontiguousarray(scan[h < 300.0])
    o = ob.Oracle(P)
    assert o.ip(scan) == 0
    seg, col = o.get("segmented_cloud"), o.get("segmentedCloudColInd")
    rng = np.random.default_rng(77)
    scan_time = 1234.5
    q, last, first = T.make_queue(rng, scan_time, wrap_at=170)
    out, visited, it = ob.adjust_distortion(seg, col, float(o.get("startOrientation")), float(o.get("endOrientation")), P.horizon_scan,
                                            scan_time, q, last, first)
    assert visited == len(seg)
    np.savez_compressed(os.path.join(HERE, "n2_adjust_distortion.npz"), scan=scan, queue=q, ptr_last=np.int32(last),
                        ptr_last_iter=np.int32(first), scan_time=np.float64(scan_time), segmented_cloud=seg, col=col,
                        adjusted=out, visited=np.int32(visited), ptr_last_iter_out=np.int32(it))
    print("n2_adjust_distortion.npz: %d points, max shift %.3f m" % (len(seg), np.abs(out[:, :3] - seg[:, :3]).max()))
    # ---- N4: keyframe against history cloud
    rng = np.random.default_rng(78)
    tgt = T.icp_scene(rng, 1500)
    src, R = T.misalign(T.icp_scene(rng, 500), [0.02, -0.008, 0.005], np.array([0.25, 0.2, -0.04]))
    src[:6, 2] += rng.uniform(8, 15, 6).astype(np.float32)
    r = ob.icp(src, tgt, exact_sums=True)
    rf = ob.icp(src, tgt, exact_sums=False)
    np.savez_compressed(os.path.join(HERE, "n4_loop_closure_icp.npz"), source=src, target=tgt, T=r["T"], fitness=np.float64(r["fitness"]),
                        iterations=np.int32(r["iterations"]), state=np.int32(r["state"]), trace=r["trace"], T_float_sums=rf["T"],
                        iterations_float_sums=np.int32(rf["iterations"]))
    print("n4_loop_closure_icp.npz: %d -> %d points, %d iterations (%s), fitness %.4f" % (len(src), len(tgt), r["iterations"],
                                                                                      ob.ICP_STATES[r["state"]], r["fitness"]))


if __name__ == "__main__":
    main()
