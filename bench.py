#!/usr/bin/env python
"""bench.py — scans/sec of the A-LeGO-LOAM hot path (ImageProjection -> LaserOdometry -> LaserMapping) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--preset ...] [--total-seq S]

One "step" = one sweep of every sequence of the batch through IP -> LO -> LM.  Each GPU (one process per GPU, torchrun for N > 1)
owns `--n-seq` INDEPENDENT sequences (weak scaling; no collective on the data path — the sequences share nothing, SURVEY.md §8e;
torch.distributed only carries the barrier, the max-over-ranks of the device-timed region and the per-rank timing gather).
`--total-seq S` instead fixes the whole job at S sequences split round-robin over the ranks (strong scaling, BASELINE config 5:
8 independent 64x2048 sequences on 1/2/4/8 GPUs).  Prints ONE JSON line (rank 0).

  value         scans/s with the sweeps already resident in HBM when the timed region starts
  e2e           scans/s through alego_pipeline_submit/_collect with HOST (pinned) buffers: H2D of every sweep and D2H of the poses
                inside the timed region
  roofline      dominant kernel: algorithmic bytes / CUDA-event duration vs the measured HBM peak
  kernels       every kernel's share of the step (CUDA events, same workload, separate pass)
  parity_check  AFTER the timed regions: the poses / feature indices the timed passes left behind, for one batch slot of every
                unique sequence, against the CPU reference chain run over the same sweeps
  cpu_baseline  the reference's CPU path on a bounded sample, 1 core, with its own per-stage timers (LM ms/iter included)

--impl reference times the reference's CPU path on all host cores, one independent sequence per core, every step a bounded sample
(`--ref-sweeps` consecutive sweeps per core).  CPU path = oracle/_ref (the reference's own sources compiled against stand-in
third-party headers, kind "reference") when those libraries exist, else the port oracle (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESETS = {"vlp16_1800": 0, "hdl64_1800": 1, "hdl64_2048": 2}
N_UNIQUE = 8  # distinct synthetic sequences (worlds + trajectories); batch slots reuse them round-robin
METRIC = "scans/sec on 64x1800 sweeps IP+LO+LM"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--preset", default="hdl64_1800", choices=sorted(PRESETS))
    ap.add_argument("--n-seq", type=int, default=256, help="independent sequences per GPU (weak scaling)")
    ap.add_argument("--total-seq", type=int, default=0, help="if > 0: sequences of the WHOLE job, split over the ranks (strong scaling)")
    ap.add_argument("--lm-every", type=int, default=1, help="LaserMapping on every k-th sweep (reference: 2)")
    ap.add_argument("--map-corner", type=int, default=50000)
    ap.add_argument("--map-surf", type=int, default=200000)
    ap.add_argument("--cpu-sweeps", type=int, default=60, help="bounded CPU sample of the cpu_baseline leg: sweeps of one sequence")
    ap.add_argument("--ref-sweeps", type=int, default=4, help="--impl reference: consecutive sweeps per core per step")
    ap.add_argument("--point-stride", type=int, default=3, choices=[3, 4],
                    help="floats per input point: 3 = packed x,y,z (the path never reads the sensor intensity), 4 = x,y,z,intensity")
    ap.add_argument("--map-order", default="voxel", choices=["voxel", "generator"],
                    help="structure of the local-map clouds: 'voxel' = pcl::VoxelGrid output (one point per occupied voxel, ascending "
                         "voxel index, x fastest) like the reference's corner_from_map_ds_ / surf_from_map_ds_ (laserMapping.cpp:316-319); "
                         "'generator' = the synthetic generator's structure-by-structure order, several points per voxel")
    ap.add_argument("--graphs", action="store_true", help="CUDA-graph replay of the pass (launch-latency regime: few sequences per GPU)")
    ap.add_argument("--cpu-kind", default="auto", choices=["auto", "reference", "port"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    return ap.parse_args()


def workload_name(args):
    return "%s IP+LO+LM, local map %dk corner + %dk surf re-indexed every mapped sweep, lm_every=%d" % (
        args.preset, args.map_corner // 1000, args.map_surf // 1000, args.lm_every)


def config_block(args, P):
    """The workload description BOTH arms print (only values that follow from the command line, so the two lines carry the same
    dict; everything measured at run time goes to "run")."""
    world = max(args.gpus, 1)
    total = args.total_seq if args.total_seq > 0 else args.n_seq * world
    return {"workload": workload_name(args), "preset": args.preset, "image": "%dx%d" % (P.n_scan, P.horizon_scan),
            "sequences_total": total, "scans_per_step": total, "scaling_mode": "strong" if args.total_seq > 0 else "weak",
            "point_stride_floats": args.point_stride, "map_order": args.map_order,
            "l2_policy": "inputs larger than L2 at the default batch (one step = %d sweeps of ~%.1f MB each); below ~100 sequences per "
                         "GPU every step still reads fresh sweeps, never the previous step's" % (
                             total // world, P.n_scan * P.horizon_scan * 0.93 * 4 * args.point_stride / 1e6),
            "lm_iters": "%d outer x <=%d LM" % (P.lm_outer_iters, P.lm_max_iters),
            "lo_iters": "%d surf + %d corner" % (P.lo_surf_iters, P.lo_corner_iters)}


def voxel_grid_output(cloud, leaf, n_keep, seed):
    """Give a cloud the structure of pcl::VoxelGrid OUTPUT — what the reference's corner_from_map_ds_ / surf_from_map_ds_ are
    (laserMapping.cpp:316-319): exactly one point per occupied voxel, in ascending voxel index ijk0 + ijk1*dx + ijk2*dx*dy (x fastest).
    The first point of every voxel is kept, then voxels are dropped at random (seeded) down to n_keep points."""
    if len(cloud) == 0:
        return cloud
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(cloud[:, :3] * inv).astype(np.int64)
    ijk -= ijk.min(axis=0)
    d = ijk.max(axis=0) + 1
    key = ijk[:, 0] + ijk[:, 1] * d[0] + ijk[:, 2] * d[0] * d[1]
    _, first = np.unique(key, return_index=True)  # ascending key, first occurrence
    if len(first) > n_keep:
        first = np.sort(np.random.default_rng(seed).choice(first, n_keep, replace=False))
        first = first[np.argsort(key[first], kind="stable")]
    return np.ascontiguousarray(cloud[first])


def make_sequences(alego, P, n_sweeps, n_corner, n_surf, ids=None, map_order="voxel"):
    """Seeded sequences (world + trajectory + local map consistent with the world), n_sweeps consecutive sweeps each.  `ids` are
    GLOBAL sequence ids; the seed depends on the id only (sharding.sequence_seed), never on the rank or the world size."""
    from alego_b200 import sharding
    ids = list(range(N_UNIQUE)) if ids is None else list(ids)

    def one(gid):
        seed = sharding.sequence_seed(gid)
        w = alego.SynthWorld(seed=seed)
        sweeps = [w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=1000 * seed + t) for t in range(n_sweeps)]
        if map_order == "voxel":  # over-generate, then one point per voxel (VoxelGrid output), exactly the requested sizes
            corner, surf = w.make_map(int(n_corner * 1.5), int(n_surf * 1.6), seed=seed, radius=100.0)
            corner = voxel_grid_output(corner, P.lm_corner_leaf, n_corner, seed)
            surf = voxel_grid_output(surf, P.lm_surf_leaf, n_surf, seed + 1)
        else:
            corner, surf = w.make_map(n_corner, n_surf, seed=seed, radius=100.0)
        return {"id": gid, "sweeps": sweeps, "map_corner": corner, "map_surf": surf}

    with ThreadPoolExecutor(max_workers=min(len(ids), os.cpu_count() or 1)) as ex:  # the generator is C++ behind ctypes (no GIL)
        return list(ex.map(one, ids))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                smax = float(f[2])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[1]))
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# launch tag (what the per-kernel profile is keyed by) -> kernel function of the ncu launch list, and which of its launches: a
# function launched for several purposes per step (the ordering kernels) is identified by its longest launch
TRAFFIC_KEY = {
    "lo_lfv_order": ("vox_order_warp", "longest"), "lm_voxel_order_3": ("vox_order_wide", "longest"),
    "maprows_fill_map_surf": ("mr_fill", "mean"), "maprows_bbox_map_surf": ("mr_bbox", "mean"), "maprows_clear_map_surf": ("mr_clear", "mean"),
    "lm_knn_surf": ("lm_knn_rows", "mean"), "lm_knn_corner": ("lm_knn", "mean"), "lm_solve": ("lm_solve<128, 2>", "mean"),
    "lm_fit_surf": ("lm_fit<0>", "mean"), "lm_fit_corner": ("lm_fit<1>", "mean"),
    "lo_assoc_surf": ("lo_assoc<1, 8>", "mean"), "lo_assoc_corner": ("lo_assoc<0, 32>", "mean"),
}


def ncu_traffic(kernel, n_seq, preset, stride):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu launch list of this same command
    line (profiles/*_ncu_traffic.json, newest first); None when no capture matches the configuration."""
    import glob
    func, which = TRAFFIC_KEY.get(kernel, (kernel, "mean"))
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_traffic.json")), reverse=True):
        try:
            d = json.load(open(f))
        except Exception:
            continue
        if d.get("n_seq") == n_seq and d.get("preset") == preset and d.get("point_stride") == stride:
            v = d.get("dram_bytes_longest_launch" if which == "longest" else "dram_bytes_per_launch", {}).get(func)
            if isinstance(v, (int, float)):
                return float(v), os.path.basename(f)
    return None, None


# ------------------------------------------------------------------------------------------------------------------------------
# The reference's CPU path (checker / baseline only — never on the product path)
# ------------------------------------------------------------------------------------------------------------------------------
def cpu_kind(args):
    """"reference" = oracle/_ref (the reference's own translation units, compiled here against stand-in third-party headers);
    "port" = oracle/alego_oracle.cpp."""
    if args.cpu_kind != "auto":
        return args.cpu_kind
    from oracle import ref_binding
    return "reference" if ref_binding.available(args.preset) else "port"


class CpuChain:
    """IP -> LO -> LM of ONE sequence on the CPU against a fixed local map, LaserMapping on every lm_every-th sweep — what
    alego_pipeline_step does on the device.  kind "reference": the three nodelets of oracle/_ref chained like their topics;
    kind "port": the oracle's pipeline_step."""

    def __init__(self, kind, P, preset, map_corner, map_surf, lm_every):
        self.kind, self.lm_every, self.count = kind, lm_every, 0
        self.cm, self.sm = map_corner, map_surf
        self.stage_ms = np.zeros(3)      # IP, LO (features + scan-to-scan), LM — accumulated
        self.lm_detail_ms = np.zeros(4)  # the reference's own TicToc: downsample, kd-tree build, association, solver
        self.lm_iters = 0
        self.lm_calls = 0
        if kind == "reference":
            from oracle import ref_binding as rb
            self.ip, self.lo, self.lm = rb.RefImageProjection(preset), rb.RefLaserOdometry(preset), rb.RefLaserMapping(preset)
        else:
            from oracle import binding as ob
            self.o = ob.Oracle(P, lm_every=lm_every, stable_voxel=False)
            self.o.lm_set_map(map_corner, map_surf)

    def step(self, scan):
        if self.kind == "port":
            self.o.pipeline_step(scan)
            t = self.o.get("timings_ms")
            self.stage_ms += [t[0], t[1] + t[2], t[3] if self.lm_every and self.count % self.lm_every == 0 else 0.0]
            if self.lm_every and self.count % self.lm_every == 0:
                self.lm_iters += self.o.report("lm")["iterations"]
                self.lm_calls += 1
            self.count += 1
            return
        t0 = time.perf_counter()
        self.ip.process(scan)
        t1 = time.perf_counter()
        self.lo.process(self.ip)
        t2 = time.perf_counter()
        if self.lm_every and self.count % self.lm_every == 0:
            odom = self.lo.get("odom_lidar") if self.count > 0 else np.array([0, 0, 0, 1.0, 0, 0, 0])
            self.lm.scan2map(self.cm, self.sm, self.lo.get("corner_last"), self.lo.get("surf_last"), self.ip.get("outlier_cloud"),
                             odom[:3], odom[3:])
            self.lm_detail_ms += self.lm.get("lm_timing_ms")
            self.lm_iters += int(self.lm.get("lm_solve_iterations").sum())
            self.lm_calls += 1
        t3 = time.perf_counter()
        self.stage_ms += [(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3]
        self.count += 1

    def state(self):
        """what the parity check compares: LaserOdometry's and LaserMapping's parameter blocks"""
        if self.kind == "port":
            return {"lo_params": self.o.get("lo_params"), "lm_params": self.o.get("lm_params"),
                    "less_sharp": self.o.get("less_sharp"), "n_seg": len(self.o.get("segmentedCloudColInd"))}
        return {"lo_params": self.lo.get("lo_params"), "lm_params": self.lm.get("lm_params") if self.lm_calls else np.zeros(6),
                "less_sharp": self.lo.get("less_sharp"), "n_seg": len(self.ip.get("segmentedCloudColInd"))}


def pingpong(n, t):
    """index into n consecutive sweeps walked forwards, then backwards, ... (consecutive sweeps stay consecutive)"""
    if n <= 1:
        return 0
    period = 2 * (n - 1)
    t %= period
    return t if t < n else period - t


def cpu_baseline_block(args, P, seq, kind):
    """rank 0, N = 1: the CPU chain on ONE core over a bounded sample of sequence 0, with per-stage times and the reference's own
    LaserMapping timers next to the device's lm block."""
    chain = CpuChain(kind, P, args.preset, seq["map_corner"], seq["map_surf"], args.lm_every)
    n = len(seq["sweeps"])
    t0 = time.perf_counter()
    for t in range(args.cpu_sweeps):
        chain.step(seq["sweeps"][pingpong(n, t)])
    dt = time.perf_counter() - t0
    per = chain.stage_ms / max(chain.count, 1)
    out = {"value": chain.count / dt, "unit": "scans/s", "cores": 1, "kind": kind,
           "sample": "%d consecutive sweeps of sequence 0 through IP+LO+LM (%s), 1 thread" % (
               chain.count, "oracle/_ref: the reference's own sources, stand-in PCL/Ceres" if kind == "reference" else "port oracle"),
           "stage_ms_per_scan": {"ip": round(per[0], 3), "lo": round(per[1], 3), "lm": round(per[2], 3)},
           # the reference runs its three stages as three nodes / threads (laserOdometry.cpp:76, laserMapping.cpp:95): pipelined over
           # 3 cores one sequence advances at 1 / max(stage)
           "three_core_pipelined_scans_per_s": round(1e3 / max(per.max(), 1e-9), 2)}
    if chain.lm_calls:
        lm = {"iters_per_scan2map": chain.lm_iters / chain.lm_calls}
        if kind == "reference":
            d = chain.lm_detail_ms / chain.lm_calls
            lm.update({"downsample_ms": round(d[0], 3), "kdtree_build_ms": round(d[1], 3), "association_ms": round(d[2], 3),
                       "solver_ms": round(d[3], 3), "ms_per_iter": round(chain.lm_detail_ms[3] / max(chain.lm_iters, 1), 4),
                       "source": "the reference's own TicToc log lines (laserMapping.cpp:344,358,464,476), mean per scan2MapOptimization"})
        else:
            lm.update({"scan2map_ms": round(chain.stage_ms[2] / chain.lm_calls, 3),
                       "ms_per_iter_upper_bound": round(chain.stage_ms[2] / max(chain.lm_iters, 1), 4)})
        out["lm"] = lm
    return out


def _ref_worker(conn, kind, P_bytes, preset, seq, lm_every, n_sweeps):
    """one host core of --impl reference: an independent sequence, n_sweeps consecutive sweeps per step"""
    import alego_pkg
    alego = alego_pkg.load()
    P = alego.AlegoParams.from_buffer_copy(P_bytes)
    chain = CpuChain(kind, P, preset, seq["map_corner"], seq["map_surf"], lm_every)
    n, t = len(seq["sweeps"]), 0
    conn.send("ready")
    while True:
        cmd = conn.recv()
        if cmd == "stop":
            break
        t0 = time.perf_counter()
        for _ in range(n_sweeps):
            chain.step(seq["sweeps"][pingpong(n, t)])
            t += 1
        conn.send(time.perf_counter() - t0)
    conn.send((chain.stage_ms.tolist(), chain.count, chain.lm_detail_ms.tolist(), chain.lm_iters, chain.lm_calls))


def run_reference(args, alego, P, rank):
    """--impl reference: the CPU path on all host cores; every core runs an independent sequence; EXACTLY W + K steps, each a
    bounded sample of `--ref-sweeps` consecutive sweeps per core."""
    if rank != 0:
        return
    import multiprocessing as mp
    kind = cpu_kind(args)
    cores = os.cpu_count() or 1
    K, W, n_sw = args.steps, args.warmup, args.ref_sweeps
    seqs = make_sequences(alego, P, 12, args.map_corner, args.map_surf, map_order=args.map_order)
    ctx = mp.get_context("fork")
    pipes, procs = [], []
    for c in range(cores):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_ref_worker, args=(b, kind, bytes(P), args.preset, seqs[c % N_UNIQUE], args.lm_every, n_sw), daemon=True)
        p.start()
        pipes.append(a)
        procs.append(p)
    for a in pipes:
        a.recv()
    step_s = []
    for it in range(W + K):
        t0 = time.perf_counter()
        for a in pipes:
            a.send("go")
        for a in pipes:
            a.recv()
        if it >= W:
            step_s.append(time.perf_counter() - t0)
    for a in pipes:
        a.send("stop")
    stats = [a.recv() for a in pipes]
    for p in procs:
        p.join(timeout=10)
    total_s = float(np.sum(step_s))
    value = cores * n_sw * K / total_s
    stage = np.sum([s[0] for s in stats], axis=0) / max(sum(s[1] for s in stats), 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": total_s / K * 1e3, "higher_is_better": True, "scaling": "strong" if args.total_seq > 0 else "weak",
        "vs_baseline": None, "dtype": "f32 geometry / f64 solver", "data": "synthetic",
        "config": config_block(args, P),
        "cpu_baseline": {"value": value, "unit": "scans/s", "cores": cores, "kind": kind,
                         "sample": "every step = %d host cores x %d consecutive sweeps (one independent sequence per core, state carried "
                                   "from step to step); %d warm-up + %d timed steps" % (cores, n_sw, W, K),
                         "stage_ms_per_scan": {"ip": round(stage[0], 3), "lo": round(stage[1], 3), "lm": round(stage[2], 3)}},
        "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------------------------
# byte models and the LM half of the metric
# ------------------------------------------------------------------------------------------------------------------------------
def algorithmic_bytes(name, st):
    pts, cells, kept = st["points"], st["cells"], st["kept"]
    pb = 4.0 * st["stride"]  # bytes per input point
    qc, qs = st.get("lm_corner_queries", 0.0), st.get("lm_surf_queries", 0.0)
    table = {
        "ip_project": (pb + 4.0) * pts,                # read the point + 4 B winner atomic
        # winner + gathered point -> cloud(16) + range(4) + ground(1) + cell flags(1)
        "ip_image": 4.0 * cells + pb * pts + 22.0 * cells,
        "ccl_rows": (1 + 4 + 8) * cells,               # flags -> parent + zeroed component statistics
        "ccl_merge": 5.0 * cells,                      # flags + the parent entries of joined cells (minimum)
        "ccl_flatten": 8.0 * cells,
        "ccl_strip": 8.0 * cells,                      # §8 d4's 8 B / cell for the whole labelling (K3+K4)
        "ip_rowcount": 13.0 * cells,
        "ip_compact": 13.0 * cells + (16 + 4 + 25.0) * kept,
        "ip_label": 8.0 * cells,
        "lo_curv_occl": 21.0 * kept,                   # 4+4 read, 4+1+4+4 written per segmented point
        # counting sort of the local map into the grid, rebuilt every mapped sweep (the reference's kd-tree builds)
        "grid_count_map_surf": 20.0 * st["map_surf_pts"],    # 16 B point + 4 B counter
        "grid_fill_map_surf": 36.0 * st["map_surf_pts"],     # 16 B point + 4 B cursor + 16 B sorted copy
        "grid_count_map_corner": 20.0 * st["map_corner_pts"],
        "grid_fill_map_corner": 36.0 * st["map_corner_pts"],
        # voxel-row index of the surf map (VoxelGrid-output maps): bounding box = read the points; fill = read the points + one
        # 8 B entry per touched word (<= one per point); clear = the table
        "maprows_bbox_map_surf": 16.0 * st["map_surf_pts"],
        "maprows_fill_map_surf": 24.0 * st["map_surf_pts"],
        # std::sort's partition phase over the (voxel, point) records of the per-ring less-flat lists: read + write each 8 B
        # record once (the ~8 partition levels in between are on-chip work when the list fits in L1 / shared memory)
        "lo_lfv_order": 16.0 * kept,
        "lo_lfv_keys": (16.0 + 4.0 + 8.0) * kept,      # point + label in, record out
        "lo_lfv_finish": (8.0 + 16.0) * kept + 16.0 * st.get("less_flat_pts", 0.0),
        # §8 d4 K14/K15: 16 B query + 5 x 16 B neighbours + 40 B result per query (candidate buckets come on top: see "traffic")
        "lm_knn_corner": 136.0 * qc, "lm_knn_surf": 136.0 * qs,
        # the fit that follows: 5 indices + query + 5 neighbours in, one correspondence block out (80 B edge, 64 B plane)
        "lm_fit_corner": (20.0 + 16.0 + 80.0 + 80.0) * qc, "lm_fit_surf": (20.0 + 16.0 + 80.0 + 64.0) * qs,
        # K11/K12 (scan-to-scan): the same 56 B per residual block per pass
        "lo_solve_surf": 56.0 * 2.0 * st.get("lo_surf_resid_iters", 0.0), "lo_solve_corner": 56.0 * 2.0 * st.get("lo_corner_resid_iters", 0.0),
        # §8 d4 K16: 56 B per residual block per pass, 2 passes per LM iteration
        "lm_solve": 56.0 * 2.0 * st.get("lm_resid_iters", 0.0),
    }
    v = table.get(name)
    return v if v else None


def lm_block(kernels, lm_reports, lo_reports, B):
    """Second half of BASELINE.json's metric: scan-to-map LM ms/iter.  One LM iteration = one residual + Jacobian pass over the
    sequence's correspondences + the 6x6 trust-region step; lm_solve runs the iterations of all B sequences of the batch in
    lock-step (one CTA per sequence), so the per-launch time / iterations is the latency of ONE iteration of one sequence
    while B of them are in flight, and / B its amortised cost."""
    it = float(np.mean([r["iterations"] for r in lm_reports])) if lm_reports else 0.0
    solve_ms = kernels.get("lm_solve", {}).get("ms_per_launch")
    assoc_ms = sum(kernels[k]["ms_per_launch"] for k in ("lm_knn_corner", "lm_fit_corner", "lm_knn_surf", "lm_fit_surf") if k in kernels)
    index_ms = sum(v["ms_per_launch"] * v["launches_per_step"] for k, v in kernels.items()
                   if (k.startswith("grid_") and "_map_" in k) or k.startswith("maprows_"))
    out = {"iters_per_scan2map": it, "edge_correspondences": float(np.mean([r["n_corner"] for r in lm_reports])) if lm_reports else 0.0,
           "plane_correspondences": float(np.mean([r["n_surf"] for r in lm_reports])) if lm_reports else 0.0,
           "lo_iters_per_scan": float(np.mean([r["iterations"] for r in lo_reports])) if lo_reports else 0.0}
    if solve_ms and it > 0:
        out["ms_per_iter"] = solve_ms / it
        out["ms_per_iter_amortised_per_sequence"] = solve_ms / it / B
        out["association_ms_per_batch"] = assoc_ms
        out["map_index_build_ms_per_batch"] = index_ms
        out["unit"] = "ms per LM iteration (latency with %d sequences in lock-step); amortised = / %d" % (B, B)
    return out


def parity_check(alego, args, P, kind, seqs_by_slot, slots, snapshots, n_steps):
    """Outside every timed region: the CPU chain over the SAME sweeps the three passes consumed (3 x n_steps consecutive sweeps per
    sequence), compared at the end of each pass with what the device left behind in the chosen batch slots."""
    def one(slot):
        seq = seqs_by_slot[slot]
        chain = CpuChain(kind, P, args.preset, seq["map_corner"], seq["map_surf"], args.lm_every)
        res = []
        for t in range(3 * n_steps):
            chain.step(seq["sweeps"][t])
            if (t + 1) % n_steps == 0:
                res.append(chain.state())
        return res

    with ThreadPoolExecutor(max_workers=min(len(slots), os.cpu_count() or 1)) as ex:  # ctypes calls release the GIL
        cpu = list(ex.map(one, slots))
    worst_lo, worst_lm, idx_equal, n = 0.0, 0.0, True, 0
    for k, slot in enumerate(slots):
        for p in range(3):
            dev, ref = snapshots[p][slot], cpu[k][p]
            worst_lo = max(worst_lo, float(np.abs(dev["lo_params"] - ref["lo_params"]).max()))
            worst_lm = max(worst_lm, float(np.abs(dev["lm_params"] - ref["lm_params"]).max()))
            idx_equal = idx_equal and dev["n_seg"] == ref["n_seg"] and np.array_equal(dev["less_sharp"], ref["less_sharp"])
            n += 1
    return {"slots": [int(s) for s in slots], "comparisons": n, "passes": ["hbm-resident", "e2e submit/collect", "per-kernel profile"],
            "sweeps_per_sequence": 3 * n_steps, "against": kind, "max_lo_pose_err": worst_lo, "max_lm_pose_err": worst_lm,
            "max_pose_err": max(worst_lo, worst_lm), "less_sharp_and_segmentation_equal": bool(idx_equal), "tolerance": 1e-4,
            "ok": bool(idx_equal and max(worst_lo, worst_lm) < 1e-4)}


_JSON_FD = None


def own_stdout():
    """The contract is ONE JSON line on stdout.  Native code under this process (the compiled reference prints progress with
    std::cout / printf) shares file descriptor 1, so move everything that is not the JSON line to stderr: fd 1 is re-pointed at
    fd 2 for the whole run and the original stdout is kept for emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def main():
    args = parse()
    own_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import alego_pkg
    alego = alego_pkg.load()
    from alego_b200 import sharding
    P = alego.default_params(PRESETS[args.preset])

    if args.impl == "reference":
        run_reference(args, alego, P, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    n_visible = torch.cuda.device_count()
    if os.environ.get("ALEGO_BENCH_DEVICE_ORDER", "") == "identity":  # for the rank -> GPU comparison in profiles/ (default: interleaved)
        device = local_rank % n_visible
    else:
        device = sharding.device_for_local_rank(local_rank, n_visible, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    torch.cuda.set_device(device)
    numa = None
    try:  # keep the rank (and the pinned sweep buffers it first-touches) on the CPUs next to its GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(device))
        numa = sorted(os.sched_getaffinity(0))
        numa = "cpus %d-%d (%d)" % (numa[0], numa[-1], len(numa))
    except Exception as e:  # best effort
        numa = "unbound (%s)" % type(e).__name__
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, max(args.warmup, 3)
    n_steps = W + K
    if args.total_seq > 0:  # strong scaling: the job's sequences split round-robin, each rank renders its own
        my_ids = sharding.sequences_of_rank(args.total_seq, rank, world)
        B = len(my_ids)
        if B == 0:
            raise SystemExit("bench.py: --total-seq %d leaves rank %d without a sequence" % (args.total_seq, rank))
        unique_ids = my_ids[:N_UNIQUE] if len(my_ids) > N_UNIQUE else my_ids
    else:           # weak scaling: every rank runs the same N_UNIQUE worlds (identical work per rank: the max over ranks is not a seed lottery)
        B = args.n_seq
        unique_ids = list(range(min(N_UNIQUE, B)))
    U = len(unique_ids)
    seqs = make_sequences(alego, P, 4 * n_steps, args.map_corner, args.map_surf, ids=unique_ids, map_order=args.map_order)
    g = alego.Alego(P, n_seq=B, device=device)
    g.pipeline_config(lm_every=args.lm_every, rebuild_map_index_every_step=True, graphs=args.graphs)
    g.set_point_stride(args.point_stride)
    PS = args.point_stride
    for b in range(B):
        s = seqs[b % U]
        g.lm_set_map(b, s["map_corner"], s["map_surf"])
    Nmax = g.max_points
    # pinned host sweep buffers, refilled per pass: A = sweeps [0,n), B = [n,2n), C = [2n,3n) of every sequence,
    # so LaserOdometry / LaserMapping always see consecutive sweeps of a trajectory
    host = [alego.pinned_empty((B, Nmax, PS), np.float32) for _ in range(n_steps)]
    host_n = [np.zeros(B, np.int32) for _ in range(n_steps)]
    pts_per_step = []

    def fill(first_sweep):
        for t in range(n_steps):
            for u in range(U):
                sw = seqs[u]["sweeps"][first_sweep + t]
                host[t][u::U, :len(sw)] = sw[:, :PS]
                host_n[t][u::U] = len(sw)
            pts_per_step.append(int(host_n[t].sum()))

    # batch slots whose results are compared with the CPU chain afterwards: one per unique sequence, from the MIDDLE of the batch
    # (a kernel whose bounded-grid stride loop breaks at large B would show there, not in slot 0)
    slots = [u + U * ((B // U) // 2) for u in range(U)] if B >= U else list(range(B))
    slots = [s for s in slots if s < B]
    snapshots = []

    def snapshot():
        g.synchronize()
        snap = {}
        for s in slots:
            snap[s] = {"lo_params": g.debug("lo_params", s).copy(), "lm_params": g.debug("lm_params", s).copy(),
                       "less_sharp": g.debug("less_sharp", s).copy(), "n_seg": len(g.debug("segmentedCloudColInd", s))}
        # every slot of a unique sequence must carry the same answer (bit for bit: nothing on the path depends on the slot index)
        same = True
        for u in range(U):
            ref = g.debug("lm_params", u)
            for b in list(range(u, B, U))[:: max(1, (B // U) // 8)]:
                same = same and np.array_equal(g.debug("lm_params", b), ref) and np.array_equal(g.debug("lo_params", b), g.debug("lo_params", u))
        snap["_slots_identical"] = bool(same)
        snapshots.append(snap)

    # ---------------- pass A: inputs resident in HBM ----------------
    fill(0)
    for t in range(n_steps):
        g.stage_upload(t, host[t], host_n[t])
    for t in range(W):
        g.stage_select(t)
        g.pipeline_step(None, None, want_poses=False)
    barrier()
    sampler = ClockSampler(device)
    l0 = g.launch_count()
    t_wall0 = time.time()
    g.timer_mark(0)
    for t in range(W, W + K):
        g.stage_select(t)
        g.pipeline_step(None, None, want_poses=False)
    g.timer_mark(1)
    barrier()
    ms_dev = g.timer_elapsed_ms(0, 1)
    launches = g.launch_count() - l0
    snapshot()

    # H2D probe: what this box's PCIe link gives a plain pinned-memory copy of one step's sweeps (context for e2e, which is
    # link-bound once the pass itself is faster than the copy).  All ranks probe at the same time, like the e2e pass does.
    probe_src = torch.empty(int(host[0].nbytes), dtype=torch.uint8).pin_memory()
    probe_dst = torch.empty(int(host[0].nbytes), dtype=torch.uint8, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe_dst.copy_(probe_src, non_blocking=True)
    barrier()
    ev0.record()
    for _ in range(4):
        probe_dst.copy_(probe_src, non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    h2d_probe_gbs = 4 * probe_src.numel() / (ev0.elapsed_time(ev1) * 1e-3) / 1e9
    del probe_src, probe_dst

    # ---------------- pass B: end to end, host buffers in, poses out ----------------
    # alego_pipeline_submit / _collect: every sweep is copied from pinned HOST memory inside the timed region (on the copy
    # stream, overlapping the previous pass) and every step's poses are read back to the host
    fill(n_steps)
    for t in range(W):  # warm-up through the same asynchronous API (its staging buffers are allocated on first use)
        g.pipeline_submit(host[t], host_n[t])
        g.pipeline_collect()
    barrier()
    t_host0 = time.perf_counter()
    g.timer_mark(2)
    DEPTH = 3  # steps in flight (alego_pipeline_submit allows three)
    for t in range(W, W + K):
        if t - W >= DEPTH:
            poses = g.pipeline_collect()
        g.pipeline_submit(host[t], host_n[t])
    for _ in range(min(K, DEPTH)):
        poses = g.pipeline_collect()
    g.timer_mark(3)
    torch.cuda.synchronize()
    ms_e2e_host = (time.perf_counter() - t_host0) * 1e3
    barrier()
    # device events on the compute stream bracket the region; the first H2D runs on the copy stream, so the (slightly
    # larger) host wall clock between the two synchronisation points is taken when it exceeds the event time
    ms_e2e_dev = g.timer_elapsed_ms(2, 3)
    ms_e2e = max(ms_e2e_dev, ms_e2e_host)
    clocks = sampler.stop(t_wall0, time.time())  # SM clocks / throttle reasons from the start of pass A to the end of pass B
    snapshot()

    # ---------------- pass C: per-kernel CUDA events on the same workload ----------------
    fill(2 * n_steps)
    for t in range(n_steps):
        g.stage_upload(t, host[t], host_n[t])
    for t in range(W):
        g.stage_select(t)
        g.pipeline_step(None, None, want_poses=False)
    g.profile_enable(True)
    g.profile_reset()
    for t in range(W, W + K):
        g.stage_select(t)
        g.pipeline_step(None, None, want_poses=False)
    g.synchronize()
    prof = g.profile()
    g.profile_enable(False)
    snapshot()

    # ---------------- pass D: pass B again, untimed, with per-step timing events ----------------
    # where a step's time goes (copy engine vs SMs, and how they overlap); the next n_steps sweeps of every sequence
    fill(3 * n_steps)
    g.pipeline_timeline(True)
    tl = []
    for t in range(W, W + K):
        if t - W >= DEPTH:
            g.pipeline_collect(want_poses=False)
            tl.append(g.pipeline_timeline(True).copy())
        g.pipeline_submit(host[t], host_n[t])
    for _ in range(min(K, DEPTH)):
        g.pipeline_collect(want_poses=False)
        tl.append(g.pipeline_timeline(True).copy())
    g.pipeline_timeline(False)
    tl = np.array(tl, np.float64)
    timeline = None
    if len(tl) >= 3:
        body = tl[1:]  # the first step has nothing to overlap with
        timeline = {"h2d_ms": round(float(np.mean(body[:, 1] - body[:, 0])), 3),
                    "front_end_ms": round(float(np.mean(body[:, 3] - body[:, 2])), 3),
                    "submit_to_poses_ms": round(float(np.mean(body[:, 4] - body[:, 0])), 3),
                    "step_interval_ms": round(float((tl[-1, 4] - tl[0, 4]) / (len(tl) - 1)), 3),
                    "copy_engine_idle_between_steps_ms": round(float(np.mean(tl[1:, 0] - tl[:-1, 1])), 3),
                    "front_end_waits_for_copy_ms": round(float(np.mean(np.maximum(body[:, 2] - np.maximum(tl[:-1, 3], body[:, 1]), 0.0))), 3),
                    "last_steps_ms": [[round(float(v - tl[-4:][0, 0]), 3) for v in row] for row in tl[-4:]],
                    "note": "CUDA events per step of an extra, untimed run of the same loop: H2D copy (copy stream), front end = "
                            "ImageProjection + LaserOdometry (main stream), poses = after LaserMapping (side stream) + D2H; last_steps_ms rows = "
                            "[H2D starts, H2D done, front end starts, front end done, poses on host]"}

    kept = float(np.mean([len(g.debug("segmentedCloudColInd", b)) for b in range(U)])) * B
    lm_reports = [g.solve_report("lm", b) for b in range(U)]
    lo_reports = [g.solve_report("lo", b) for b in range(U)]
    lm_queries = (float(np.mean([len(g.debug("lm_corner_ds", b)) for b in range(U)])) * B,
                  float(np.mean([len(g.debug("lm_surf_total_ds", b)) for b in range(U)])) * B)

    # max over ranks for the headline; every rank's own figures for attribution
    ms_dev_max, ms_e2e_max = sharding.reduce_max_ms([ms_dev, ms_e2e], dist if world > 1 else None, device="cuda")
    per_rank = sharding.gather_rank_stats({"rank": rank, "device": device, "ms_dev_per_step": ms_dev / K, "ms_e2e_per_step": ms_e2e / K,
                                           "h2d_probe_gbs": round(h2d_probe_gbs, 1), "cpu_affinity": numa, "n_seq": B},
                                          dist if world > 1 else None)

    if rank == 0:
        total_seq = sum(r["n_seq"] for r in per_rank)
        value = total_seq * K / (ms_dev_max * 1e-3)
        e2e_value = total_seq * K / (ms_e2e_max * 1e-3)
        st = {"points": float(np.mean(pts_per_step)), "cells": float(B * P.n_scan * P.horizon_scan), "kept": kept, "B": B,
              "ground_rows": min(P.ground_scan_id + 1, P.n_scan), "R": P.n_scan, "stride": PS,
              "map_surf_pts": float(B * np.mean([len(q["map_surf"]) for q in seqs])),
              "map_corner_pts": float(B * np.mean([len(q["map_corner"]) for q in seqs])),
              "lm_corner_queries": lm_queries[0], "lm_surf_queries": lm_queries[1],
              "lm_resid_iters": B * float(np.mean([(r["n_corner"] + r["n_surf"]) * r["iterations"] for r in lm_reports])),
              # LaserOdometry: surf solve then corner solve, fixed iteration counts (laserOdometry.cpp:326, :410-421, :484-502)
              # (the report carries the iterations of both solves together: halves)
              "lo_surf_resid_iters": B * float(np.mean([r["n_surf"] * r["iterations"] / 2.0 for r in lo_reports])),
              "lo_corner_resid_iters": B * float(np.mean([r["n_corner"] * r["iterations"] / 2.0 for r in lo_reports]))}
        total_kernel_ms = sum(ms for _, ms in prof.values())
        kernels = {}
        for name, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
            per_launch_ms = ms / max(n, 1)
            ab = algorithmic_bytes(name, st)
            kernels[name] = {"launches_per_step": n / K, "ms_per_launch": round(per_launch_ms, 5), "share": round(ms / total_kernel_ms, 4),
                             "algorithmic_gbs": round(ab / (per_launch_ms * 1e-3) / 1e9, 1) if ab else None}
        peak, peak_src = measured_peak_gbs()
        dom = next(iter(kernels))
        # roofline of the dominant kernel that has a streaming byte model; plus the two kernels north_star names
        dom_stream = next((k for k in kernels if kernels[k]["algorithmic_gbs"] is not None), None)
        ab = algorithmic_bytes(dom_stream, st)
        achieved = kernels[dom_stream]["algorithmic_gbs"]
        traffic, traffic_src = ncu_traffic(dom_stream, B, args.preset, PS)
        roofline = {"bound": "hbm", "kernel": dom_stream, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                    "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": ab,
                    "dominant_kernel_overall": dom,
                    "named": {k: {"achieved": kernels[k]["algorithmic_gbs"], "frac": round(kernels[k]["algorithmic_gbs"] / peak, 4),
                                  "algorithmic_bytes_per_launch": algorithmic_bytes(k, st), "traffic": ncu_traffic(k, B, args.preset, PS)[0]}
                              for k in ("ip_project", "ip_image", "lo_curv_occl") if k in kernels}}
        dev_ms = [r["ms_dev_per_step"] for r in per_rank]
        e2e_ms = [r["ms_e2e_per_step"] for r in per_rank]
        line = {
            "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev_max / K, "higher_is_better": True, "scaling": "strong" if args.total_seq > 0 else "weak", "vs_baseline": None,
            "dtype": "f32 geometry / f64 solver", "data": "synthetic (seeded ray-cast sweeps, %d unique sequences reused round-robin over the batch; "
                                                          "the same sequences on every rank)" % U,
            "config": config_block(args, P),
            "run": {"n_seq_per_gpu": B, "points_per_scan": st["points"] / B, "graphs": bool(args.graphs),
                    "sweep_mb_per_step_per_gpu": round(st["points"] * 4 * PS / 1e6, 1),
                    "ms_dev_per_step_min_median_max": [round(float(f(dev_ms)), 4) for f in (np.min, np.median, np.max)],
                    "ms_e2e_per_step_min_median_max": [round(float(f(e2e_ms)), 4) for f in (np.min, np.median, np.max)],
                    "per_rank": per_rank},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "scans/s", "h2d_bytes_per_step": int(st["points"] * 4 * PS + B * 4), "d2h_bytes_per_step": B * 12 * 8,
                    "ms_per_step": ms_e2e_max / K, "host_wall_ms_per_step": ms_e2e_host / K, "device_event_ms_per_step": ms_e2e_dev / K,
                    "h2d_probe_gbs": round(h2d_probe_gbs, 1), "timeline": timeline,
                    "api": "alego_pipeline_submit/_collect, pinned host sweeps, 3 steps in flight (H2D of sweep t+1 overlaps the pass over sweep t)"},
            "gpu_launches": int(launches),
            "lm": lm_block(kernels, lm_reports, lo_reports, B),
            "roofline": roofline,
            "kernels": kernels,
            "kernel_ms_per_step": total_kernel_ms / K,
        }
        kind = cpu_kind(args)
        if not args.no_parity_check:
            seqs_by_slot = {s: seqs[s % U] for s in slots}
            pc = parity_check(alego, args, P, kind, seqs_by_slot, slots, snapshots, n_steps)
            pc["all_slots_of_a_sequence_identical"] = all(s["_slots_identical"] for s in snapshots)
            line["parity_check"] = pc
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_block(args, P, seqs[0], kind)
        else:
            line["cpu_baseline"] = None
        emit(line)
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
