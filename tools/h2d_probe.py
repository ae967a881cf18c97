#!/usr/bin/env python
"""Concurrent pinned host->device copy bandwidth per GPU SET on a multi-GPU box (the end-to-end rate of this path is bounded by
sweep ingestion: 1.3 MB per sweep).  One process, one stream per GPU, the pinned source of every GPU allocated while the calling
thread is bound to that GPU's CPU set (NVML).  Prints `nvidia-smi topo -m`, the NUMA layout and a JSON table:
per set, the GB/s every member GPU reached while all members copied at once.

    python tools/h2d_probe.py [--mb 256] [--reps 6] > gpurun_out/h2d_probe.json
"""
import argparse
import json
import os
import subprocess
import sys

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256)
    ap.add_argument("--reps", type=int, default=6)
    args = ap.parse_args()
    n = torch.cuda.device_count()
    out = {"n_gpus": n, "mb": args.mb, "sets": []}
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["numactl", "-H"]):
        try:
            txt = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
        except Exception as e:  # tool not installed
            txt = "unavailable: %s" % e
        out[" ".join(cmd)] = txt.splitlines()[:60]
    try:
        import pynvml
        pynvml.nvmlInit()
    except Exception:
        pynvml = None
    all_cpus = os.sched_getaffinity(0)
    src, dst, aff = {}, {}, {}
    for d in range(n):
        if pynvml is not None:
            try:
                pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(d))
                c = sorted(os.sched_getaffinity(0))
                aff[d] = "%d-%d (%d)" % (c[0], c[-1], len(c))
            except Exception as e:
                aff[d] = "unbound (%s)" % type(e).__name__
        src[d] = torch.empty(args.mb << 20, dtype=torch.uint8).pin_memory()
        src[d].fill_(d)
        dst[d] = torch.empty(args.mb << 20, dtype=torch.uint8, device="cuda:%d" % d)
        os.sched_setaffinity(0, all_cpus)
    out["cpu_affinity"] = aff
    half = n // 2
    sets = [[0]]
    if n >= 2:
        sets += [[0, 1], [0, half]]
    if n >= 4:
        sets += [list(range(4)), [0, half, 1, half + 1], [0, 2, half, half + 2]]
    if n >= 8:
        sets += [list(range(8))]
    for gset in sets:
        streams = {d: torch.cuda.Stream(device=d) for d in gset}
        ev = {d: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for d in gset}
        for d in gset:  # warm-up
            with torch.cuda.stream(streams[d]):
                dst[d].copy_(src[d], non_blocking=True)
        for d in gset:
            torch.cuda.synchronize(d)
        for d in gset:
            with torch.cuda.stream(streams[d]):
                ev[d][0].record(streams[d])
        for _ in range(args.reps):
            for d in gset:
                with torch.cuda.stream(streams[d]):
                    dst[d].copy_(src[d], non_blocking=True)
        for d in gset:
            with torch.cuda.stream(streams[d]):
                ev[d][1].record(streams[d])
        for d in gset:
            torch.cuda.synchronize(d)
        gbs = {d: round(args.reps * (args.mb << 20) / (ev[d][0].elapsed_time(ev[d][1]) * 1e-3) / 1e9, 1) for d in gset}
        out["sets"].append({"gpus": gset, "gbs_per_gpu": gbs, "aggregate_gbs": round(sum(gbs.values()), 1)})
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
