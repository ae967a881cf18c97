#!/usr/bin/env python
"""Launch list of `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` -> markdown table
(per kernel name: launches, mean device time, share of the captured window, mean DRAM bytes per launch and GB/s) and the traffic
JSON bench.py reads for roofline.traffic.
usage: ncu_launch_table.py launches.csv n_seq preset point_stride traffic.json > table.md"""
import csv
import json
import re
import sys

path, n_seq, preset, stride, out_json = sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), sys.argv[5]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
rows = rows[rows.index(hdr) + 1:]
iid, iname, imet, iunit, ival = (hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "nsecond": 1e-3,
         "msecond": 1e3, "second": 1e6}
launch = {}
for r in rows:
    d = launch.setdefault(r[iid], {"name": re.sub(r"\(.*", "", r[iname]).split("::")[-1]})
    d[r[imet]] = float(r[ival].replace(",", "")) * scale.get(r[iunit], 1)
agg = {}
for d in launch.values():
    a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "bytes": 0.0, "longest_us": -1.0, "longest_bytes": 0.0})
    a["n"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0.0)
    by = d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a["bytes"] += by
    if d.get("gpu__time_duration.sum", 0.0) > a["longest_us"]:  # a function launched for several purposes: its biggest job
        a["longest_us"], a["longest_bytes"] = d.get("gpu__time_duration.sum", 0.0), by
tot = sum(a["us"] for a in agg.values())
print("| kernel | launches | mean us | share of window | DRAM MB / launch | DRAM GB/s |\n|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    us, by = a["us"] / a["n"], a["bytes"] / a["n"]
    print("| %s | %d | %.1f | %.1f %% | %.1f | %.0f |" % (k, a["n"], us, 100 * a["us"] / tot, by / 1e6, by / (us * 1e-6) / 1e9 if us else 0))
print("\nwindow: %d launches, %.3f ms of kernel time (serialised, cold-cache: compare SHARES with bench.py's, not absolutes)" % (len(launch), tot / 1e3))
json.dump({"n_seq": n_seq, "preset": preset, "point_stride": stride, "source": path,
           "note": "mean dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed by kernel function name (launch-list capture)",
           "dram_bytes_per_launch": {k.replace("_kernel", ""): a["bytes"] / a["n"] for k, a in agg.items()},
           "dram_bytes_longest_launch": {k.replace("_kernel", ""): a["longest_bytes"] for k, a in agg.items()}}, open(out_json, "w"), indent=1)
