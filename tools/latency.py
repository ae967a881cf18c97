#!/usr/bin/env python
"""Per-sweep latency of the synchronous alego_pipeline_step (host sweep in, poses out) for small batches, eager launches vs
CUDA-graph replay (alego_pipeline_config options bit 1).  Run on the GPU box."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import alego_pkg

alego = alego_pkg.load()
P = alego.default_params(1)
seed, T = 100, 60
w = alego.SynthWorld(seed=seed)
corner, surf = w.make_map(50000, 200000, seed=seed, radius=100.0)
sweeps = [w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=1000 * seed + t) for t in range(T)]
out = {}
for n_seq in (1, 8):
    for graphs in (False, True):
        g = alego.Alego(P, n_seq=n_seq)
        g.set_point_stride(3)
        for b in range(n_seq):
            g.lm_set_map(b, corner, surf)
        g.pipeline_config(lm_every=1, graphs=graphs)
        buf = alego.pinned_empty((n_seq, g.max_points, 3), np.float32)
        n = np.zeros(n_seq, np.int32)
        ts = []
        for t in range(T):
            b_, n_ = g.pack_scans([sweeps[t]] * n_seq)
            buf[:] = b_
            n[:] = n_
            t0 = time.perf_counter()
            poses = g.pipeline_step(buf, n)
            ts.append(time.perf_counter() - t0)
        g.close()
        out["n_seq=%d %s" % (n_seq, "graph" if graphs else "eager")] = {"median_ms": round(1e3 * float(np.median(ts[10:])), 4),
                                                                          "p95_ms": round(1e3 * float(np.percentile(ts[10:], 95)), 4)}
print(json.dumps(out, indent=1))
