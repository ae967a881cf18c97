#!/usr/bin/env python
"""Generates tests/golden/n1_local_map.npz — golden vectors of the ORACLE's local-map assembly (SURVEY §8f row N1,
laserMapping.cpp:194-323) on seeded keyframes: inputs (keyframe clouds + poses) and outputs (keyframe matrices,
corner_from_map_ds_, surf_from_map_ds_).  Separate from make_golden.py so that the older fixtures stay byte-identical.

    python tests/golden/make_golden_n1.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def main():
    from oracle import binding as ob
    rng = np.random.default_rng(2024)
    K = 4
    sizes = [(60, 400, 30), (45, 380, 0), (0, 410, 25), (70, 350, 40)]  # ragged, with empty clouds
    poses = (rng.uniform(-1, 1, (K, 6)) * np.array([6, 6, 0.3, 0.03, 0.03, 2.0])).astype(np.float32)
    out = {"poses6": poses}
    ck, sk, okf = [], [], []
    for k, (nc, ns, no) in enumerate(sizes):
        c = rng.uniform(-15, 15, (nc, 4)).astype(np.float32)
        s = rng.uniform(-15, 15, (ns, 4)).astype(np.float32)
        o = rng.uniform(-15, 15, (no, 4)).astype(np.float32)
        s[:, 2] = rng.normal(-1.7, 0.05, ns).astype(np.float32)  # mostly a ground sheet: many points per 0.8 m voxel
        ck.append(c); sk.append(s); okf.append(o)
        out["corner%d" % k], out["surf%d" % k], out["outlier%d" % k] = c, s, o
    cm, sm, M = ob.lm_assemble_map(ck, sk, okf, poses, 0.4, 0.8, stable=False)
    out["corner_from_map_ds"], out["surf_from_map_ds"], out["matrices"] = cm, sm, M
    np.savez_compressed(os.path.join(HERE, "n1_local_map.npz"), **out)
    print("n1_local_map.npz: %d corner, %d surf map points" % (len(cm), len(sm)))


if __name__ == "__main__":
    main()
