#!/usr/bin/env python
"""bench.py — scans/sec of the A-LeGO-LOAM hot path (ImageProjection -> LaserOdometry -> LaserMapping) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one sweep of every sequence of the batch through IP -> LO -> LM.  Each GPU (one process per GPU,
torchrun for N > 1) owns `--n-seq` INDEPENDENT sequences (weak scaling; no collective on the data path — the
sequences share nothing; torch.distributed is only used for the barrier and the max-over-ranks
of the device-timed region).  Prints ONE JSON line (rank 0).

  value      scans/s with the sweeps already resident in HBM when the timed region starts
  e2e        scans/s through alego_pipeline_step with HOST (pinned) buffers: H2D of every sweep and D2H of the
             poses inside the timed region
  roofline   dominant kernel: algorithmic bytes / CUDA-event duration vs the measured HBM peak
  kernels    every kernel's share of the step (CUDA events, same workload, separate pass)
  cpu_baseline  the oracle (CPU restatement of the reference, kind "port") on a bounded sample, 1 core

--impl reference times the reference's CPU path (the oracle port — the reference itself cannot be built here,
see DESIGN.md) on all host cores, independent sequences in parallel processes.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PRESETS = {"vlp16_1800": 0, "hdl64_1800": 1, "hdl64_2048": 2}
N_UNIQUE = 8  # distinct synthetic sequences (worlds + trajectories); batch slots reuse them round-robin


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--preset", default="hdl64_1800", choices=sorted(PRESETS))
    ap.add_argument("--n-seq", type=int, default=256, help="independent sequences per GPU")
    ap.add_argument("--lm-every", type=int, default=1, help="LaserMapping on every k-th sweep (reference: 2)")
    ap.add_argument("--map-corner", type=int, default=50000)
    ap.add_argument("--map-surf", type=int, default=200000)
    ap.add_argument("--cpu-sweeps", type=int, default=40, help="bounded CPU sample: sweeps per sequence")
    ap.add_argument("--point-stride", type=int, default=3, choices=[3, 4],
                    help="floats per input point: 3 = packed x,y,z (the path never reads the sensor intensity), 4 = x,y,z,intensity")
    ap.add_argument("--map-order", default="voxel", choices=["voxel", "generator"],
                    help="order of the local-map clouds: 'voxel' = ascending PCL VoxelGrid index (x fastest), the order of the "
                         "reference's corner_from_map_ds_ / surf_from_map_ds_ (VoxelGrid outputs, laserMapping.cpp:316-319); "
                         "'generator' = the synthetic generator's structure-by-structure order")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(args):
    return "%s IP+LO+LM, local map %dk corner + %dk surf re-indexed every mapped sweep, lm_every=%d" % (
        args.preset, args.map_corner // 1000, args.map_surf // 1000, args.lm_every)


def voxel_order(cloud, leaf):
    """Reorder a cloud the way pcl::VoxelGrid emits its output: ascending voxel index ijk0 + ijk1*dx + ijk2*dx*dy (stable)."""
    if len(cloud) == 0:
        return cloud
    inv = np.float32(1.0) / np.float32(leaf)
    ijk = np.floor(cloud[:, :3] * inv).astype(np.int64)
    ijk -= ijk.min(axis=0)
    d = ijk.max(axis=0) + 1
    key = ijk[:, 0] + ijk[:, 1] * d[0] + ijk[:, 2] * d[0] * d[1]
    return np.ascontiguousarray(cloud[np.argsort(key, kind="stable")])


def make_sequences(alego, P, n_steps, n_corner, n_surf, rank=0, map_order="voxel"):
    """N_UNIQUE seeded sequences: n_steps consecutive sweeps each + a local map consistent with the world."""
    seqs = []
    for u in range(N_UNIQUE):
        seed = 100 + 17 * rank + u
        w = alego.SynthWorld(seed=seed)
        sweeps = [w.render(P, alego.trajectory_pose(t, seed=seed), noise_seed=1000 * seed + t) for t in range(n_steps)]
        corner, surf = w.make_map(n_corner, n_surf, seed=seed, radius=100.0)
        if map_order == "voxel":
            corner, surf = voxel_order(corner, P.lm_corner_leaf), voxel_order(surf, P.lm_surf_leaf)
        seqs.append({"sweeps": sweeps, "map_corner": corner, "map_surf": surf})
    return seqs


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


def ncu_traffic(kernel, n_seq, preset, stride):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture of this
    same command line (profiles/*_ncu_traffic.json, newest first); None when no capture matches the configuration."""
    import glob
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_traffic.json")), reverse=True):
        try:
            d = json.load(open(f))
        except Exception:
            continue
        if d.get("n_seq") == n_seq and d.get("preset") == preset and d.get("point_stride") == stride:
            v = d.get("dram_bytes_per_launch", {}).get(kernel)
            if isinstance(v, (int, float)):
                return float(v), os.path.basename(f)
    return None, None


def cpu_oracle_run(P_bytes, preset_id, sweeps, map_corner, map_surf, lm_every):
    """Time the oracle on one sequence (single thread). Returns seconds for len(sweeps) sweeps."""
    from oracle import binding as ob
    import alego_pkg
    alego = alego_pkg.load()
    P = alego.AlegoParams.from_buffer_copy(P_bytes)
    o = ob.Oracle(P, lm_every=lm_every, stable_voxel=False)
    o.lm_set_map(map_corner, map_surf)
    t0 = time.perf_counter()
    for s in sweeps:
        o.pipeline_step(s)
    dt = time.perf_counter() - t0
    return dt, o.get("timings_ms").tolist()


def _cpu_worker(args):
    return cpu_oracle_run(*args)


def run_reference(args, alego, P, rank, world):
    """--impl reference: the CPU path on all host cores; every core runs an independent sequence."""
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    n_sweeps = max(4, min(args.cpu_sweeps, 12))
    seqs = make_sequences(alego, P, n_sweeps, args.map_corner, args.map_surf, map_order=args.map_order)
    jobs = [(bytes(P), PRESETS[args.preset], seqs[c % N_UNIQUE]["sweeps"], seqs[c % N_UNIQUE]["map_corner"], seqs[c % N_UNIQUE]["map_surf"],
             args.lm_every) for c in range(cores)]
    ctx = mp.get_context("fork")
    step_times = []
    with ctx.Pool(cores) as pool:
        for it in range(args.warmup + args.steps):
            if it >= max(1, min(args.warmup, 1)) + min(args.steps, 3):  # bounded: the whole run stays within minutes
                break
            t0 = time.perf_counter()
            pool.map(_cpu_worker, jobs)
            dt = time.perf_counter() - t0
            if it >= min(args.warmup, 1):
                step_times.append(dt)
    per_step = float(np.mean(step_times))
    value = cores * n_sweeps / per_step
    line = {
        "impl": "reference", "metric": "scans/sec on 64x1800 sweeps IP+LO+LM", "value": value, "unit": "scans/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 geometry / f64 solver", "data": "synthetic",
        "config": {"workload": workload_name(args), "map_order": args.map_order,
                   "scans_per_step": cores * n_sweeps, "timed_steps_executed": len(step_times),
                   "lm_iters": "%d outer x <=%d LM" % (P.lm_outer_iters, P.lm_max_iters), "lo_iters": "%d surf + %d corner" % (P.lo_surf_iters, P.lo_corner_iters)},
        "cpu_baseline": {"value": value, "unit": "scans/s", "cores": cores, "kind": "port",
                         "sample": "%d cores x %d consecutive sweeps (one independent sequence per core) per step" % (cores, n_sweeps)},
        "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# algorithmic bytes per launch of the kernels that have a streaming model
def algorithmic_bytes(name, st):
    pts, cells, kept, B = st["points"], st["cells"], st["kept"], st["B"]
    pb = 4.0 * st["stride"]  # bytes per input point
    gfrac = st["ground_rows"] / st["R"]
    table = {
        "ip_project": (pb + 4.0) * pts,                # read the point + 4 B winner atomic
        # winner + gathered point -> cloud(16) + range(4) + ground(1) + cell flags(1)
        "ip_image": 4.0 * cells + pb * pts + 22.0 * cells,
        "ccl_rows": (1 + 4 + 8) * cells,               # flags -> parent + zeroed component statistics
        "ccl_merge": 5.0 * cells,                      # flags + the parent entries of joined cells (minimum)
        "ccl_flatten": 8.0 * cells,
        "ip_rowcount": 13.0 * cells,
        "ip_compact": 13.0 * cells + (16 + 4 + 25.0) * kept,
        "ip_label": 8.0 * cells,
        "lo_curv_occl": 21.0 * kept,                   # 4+4 read, 4+1+4+4 written per segmented point
        # counting sort of the local map into the hashed grid, rebuilt every mapped sweep (the reference's kd-tree builds)
        "grid_count_map_surf": 20.0 * st["map_surf_pts"],    # 16 B point + 4 B counter
        "grid_fill_map_surf": 36.0 * st["map_surf_pts"],     # 16 B point + 4 B cursor + 16 B sorted copy
        "grid_count_map_corner": 20.0 * st["map_corner_pts"],
        "grid_fill_map_corner": 36.0 * st["map_corner_pts"],
    }
    return table.get(name)


def lm_block(kernels, lm_reports, lo_reports, B):
    """Second half of BASELINE.json's metric: scan-to-map LM ms/iter.  One LM iteration = one residual + Jacobian pass over the
    sequence's correspondences + the 6x6 trust-region step; lm_solve runs the iterations of all B sequences of the batch in
    lock-step (one CTA per sequence), so the per-launch time / iterations is the latency of ONE iteration of one sequence
    while B of them are in flight, and / B its amortised cost."""
    it = float(np.mean([r["iterations"] for r in lm_reports])) if lm_reports else 0.0
    solve_ms = kernels.get("lm_solve", {}).get("ms_per_launch")
    assoc_ms = sum(kernels[k]["ms_per_launch"] for k in ("lm_knn_corner", "lm_fit_corner", "lm_knn_surf", "lm_fit_surf") if k in kernels)
    out = {"iters_per_scan2map": it, "edge_correspondences": float(np.mean([r["n_corner"] for r in lm_reports])) if lm_reports else 0.0,
           "plane_correspondences": float(np.mean([r["n_surf"] for r in lm_reports])) if lm_reports else 0.0,
           "lo_iters_per_scan": float(np.mean([r["iterations"] for r in lo_reports])) if lo_reports else 0.0}
    if solve_ms and it > 0:
        out["ms_per_iter"] = solve_ms / it
        out["ms_per_iter_amortised_per_sequence"] = solve_ms / it / B
        out["association_ms_per_batch"] = assoc_ms
        out["unit"] = "ms per LM iteration (latency with %d sequences in lock-step); amortised = / %d" % (B, B)
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import alego_pkg
    alego = alego_pkg.load()
    P = alego.default_params(PRESETS[args.preset])

    if args.impl == "reference":
        run_reference(args, alego, P, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = None
    try:  # keep the rank (and the pinned sweep buffers it first-touches) on the CPUs next to its GPU
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        numa = sorted(os.sched_getaffinity(0))
        numa = "cpus %d-%d (%d)" % (numa[0], numa[-1], len(numa))
    except Exception as e:  # best effort
        numa = "unbound (%s)" % type(e).__name__
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W, B = args.steps, max(args.warmup, 3), args.n_seq
    n_steps = W + K
    seqs = make_sequences(alego, P, 3 * n_steps, args.map_corner, args.map_surf, rank=rank, map_order=args.map_order)
    g = alego.Alego(P, n_seq=B, device=local_rank)
    g.pipeline_config(lm_every=args.lm_every, rebuild_map_index_every_step=True)
    g.set_point_stride(args.point_stride)
    PS = args.point_stride
    for b in range(B):
        s = seqs[b % N_UNIQUE]
        g.lm_set_map(b, s["map_corner"], s["map_surf"])
    Nmax = g.max_points
    # pinned host sweep buffers, refilled per pass: A = sweeps [0,n), B = [n,2n), C = [2n,3n) of every sequence,
    # so LaserOdometry / LaserMapping always see consecutive sweeps of a trajectory
    host = [alego.pinned_empty((B, Nmax, PS), np.float32) for _ in range(n_steps)]
    host_n = [np.zeros(B, np.int32) for _ in range(n_steps)]
    pts_per_step = []

    def fill(first_sweep):
        for t in range(n_steps):
            for u in range(min(N_UNIQUE, B)):
                sw = seqs[u]["sweeps"][first_sweep + t]
                host[t][u::N_UNIQUE, :len(sw)] = sw[:, :PS]
                host_n[t][u::N_UNIQUE] = len(sw)
            pts_per_step.append(int(host_n[t].sum()))

    # ---------------- pass A: inputs resident in HBM ----------------
    fill(0)
    for t in range(n_steps):
        g.stage_upload(t, host[t], host_n[t])
    for t in range(W):
        g.stage_select(t)
        g.pipeline_step(None, None, want_poses=False)
    barrier()
    sampler = ClockSampler(local_rank)
    l0 = g.launch_count()
    t_wall0 = time.time()
    g.timer_mark(0)
    for t in range(W, W + K):
        g.stage_select(t)
        g.pipeline_step(None, None, want_poses=False)
    g.timer_mark(1)
    barrier()
    t_wall1 = time.time()
    ms_dev = g.timer_elapsed_ms(0, 1)
    launches = g.launch_count() - l0

    # H2D probe: what this box's PCIe link gives a plain pinned-memory copy of one step's sweeps (context for e2e, which is
    # link-bound once the pass itself is faster than the copy)
    probe_src = torch.empty(int(host[0].nbytes), dtype=torch.uint8).pin_memory()
    probe_dst = torch.empty(int(host[0].nbytes), dtype=torch.uint8, device="cuda")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    probe_dst.copy_(probe_src, non_blocking=True)
    torch.cuda.synchronize()
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
    barrier()
    ms_e2e_host = (time.perf_counter() - t_host0) * 1e3
    # device events on the compute stream bracket the region; the first H2D runs on the copy stream, so the (slightly
    # larger) host wall clock between the two synchronisation points is taken when it exceeds the event time
    ms_e2e_dev = g.timer_elapsed_ms(2, 3)
    ms_e2e = max(ms_e2e_dev, ms_e2e_host)
    clocks = sampler.stop(t_wall0, time.time())  # SM clocks / throttle reasons from the start of pass A to the end of pass B

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
    kept = float(np.mean([len(g.debug("segmentedCloudColInd", b)) for b in range(min(B, N_UNIQUE))])) * B
    lm_reports = [g.solve_report("lm", b) for b in range(min(B, N_UNIQUE))]
    lo_reports = [g.solve_report("lo", b) for b in range(min(B, N_UNIQUE))]

    times = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(times[0]), float(times[1])

    if rank == 0:
        scans = B * K * world
        value = scans / (ms_dev * 1e-3)
        e2e_value = scans / (ms_e2e * 1e-3)
        st = {"points": float(np.mean(pts_per_step)), "cells": float(B * P.n_scan * P.horizon_scan), "kept": kept, "B": B,
              "ground_rows": min(P.ground_scan_id + 1, P.n_scan), "R": P.n_scan, "stride": PS,
              "map_surf_pts": float(B * np.mean([len(q["map_surf"]) for q in seqs])),
              "map_corner_pts": float(B * np.mean([len(q["map_corner"]) for q in seqs]))}
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
        line = {
            "metric": "scans/sec on 64x1800 sweeps IP+LO+LM", "value": value, "unit": "scans/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 geometry / f64 solver", "data": "synthetic (seeded ray-cast sweeps, %d unique sequences reused round-robin over the batch)" % N_UNIQUE,
            "config": {"workload": workload_name(args),
                       "n_seq_per_gpu": B, "scans_per_step": B * world, "points_per_scan": st["points"] / B,
                       "point_stride_floats": PS, "map_order": args.map_order, "rank0_cpu_affinity": numa,
                       "l2_policy": "inputs larger than L2 (%.0f MB of sweeps per step per GPU)" % (st["points"] * 4 * PS / 1e6),
                       "lm_iters": "%d outer x <=%d LM" % (P.lm_outer_iters, P.lm_max_iters), "lo_iters": "%d surf + %d corner" % (P.lo_surf_iters, P.lo_corner_iters)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "scans/s", "h2d_bytes_per_step": int(st["points"] * 4 * PS + B * 4), "d2h_bytes_per_step": B * 12 * 8,
                    "ms_per_step": ms_e2e / K, "host_wall_ms_per_step": ms_e2e_host / K, "device_event_ms_per_step": ms_e2e_dev / K,
                    "h2d_probe_gbs": round(h2d_probe_gbs, 1),
                    "api": "alego_pipeline_submit/_collect, pinned host sweeps, 3 steps in flight (H2D of sweep t+1 overlaps the pass over sweep t)"},
            "gpu_launches": int(launches),
            "lm": lm_block(kernels, lm_reports, lo_reports, B),
            "roofline": roofline,
            "kernels": kernels,
            "kernel_ms_per_step": total_kernel_ms / K,
        }
        if not args.no_cpu_baseline and world == 1:
            n_sw = args.cpu_sweeps
            s0 = seqs[0]
            sw = [s0["sweeps"][t % len(s0["sweeps"])] for t in range(min(n_sw, len(s0["sweeps"])))]
            dt, stage_ms = cpu_oracle_run(bytes(P), PRESETS[args.preset], sw, s0["map_corner"], s0["map_surf"], args.lm_every)
            line["cpu_baseline"] = {"value": len(sw) / dt, "unit": "scans/s", "cores": 1, "kind": "port",
                                    "sample": "%d consecutive sweeps of sequence 0 through the oracle (IP+LO+LM), 1 thread" % len(sw),
                                    "last_sweep_stage_ms": {"ip": stage_ms[0], "features": stage_ms[1], "scan2scan": stage_ms[2], "scan2map": stage_ms[3]}}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
