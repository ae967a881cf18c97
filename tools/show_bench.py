#!/usr/bin/env python
"""Print a compact view of a bench.py JSON line (value, e2e, per-kernel ms)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e)
        continue
    print("%s: value %.0f %s  e2e %.0f  ms/step %.3f  kernel-sum %.3f  launches %s" % (
        f, d["value"], d["unit"], d["e2e"]["value"], d["ms_per_step"], d.get("kernel_ms_per_step", 0), d.get("gpu_launches")))
    if "roofline" in d:
        print("  roofline:", {k: d["roofline"][k] for k in ("kernel", "achieved", "frac", "traffic")}, d["roofline"].get("named"))
    for k, v in d.get("kernels", {}).items():
        print("  %-24s %8.4f ms  %5.1f%%  %s" % (k, v["ms_per_launch"] * v["launches_per_step"], 100 * v["share"],
                                                  "" if v["algorithmic_gbs"] is None else "%.0f GB/s" % v["algorithmic_gbs"]))
    if d.get("cpu_baseline"):
        print("  cpu_baseline:", d["cpu_baseline"]["value"], d["cpu_baseline"]["unit"], d["cpu_baseline"]["cores"], "core(s)")
