#!/usr/bin/env python
"""Per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of the uniquely named kernels of an
`ncu --set full ... --page raw --csv` export -> JSON that bench.py reads for roofline.traffic.
usage: ncu_traffic.py raw.csv n_seq preset point_stride > profiles/rNN_ncu_traffic.json"""
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(r, name):
    return float(r[idx[name]].replace(",", "")) * scale.get(units[idx[name]], 1)


acc = {}
for r in data:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).split("::")[-1].replace("_kernel", "")
    name = re.sub(r"<.*", "", name).replace("void ", "").strip()
    t = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    acc.setdefault(name, []).append(t)
MULTI_ROLE = ("grid_count", "grid_scan", "grid_fill", "lm_knn", "lm_fit", "lm_voxel", "lo_assoc", "lo_solve", "lo_pose")
out = {"n_seq": int(sys.argv[2]), "preset": sys.argv[3], "point_stride": int(sys.argv[4]), "source": sys.argv[1],
       "note": "kernels with one role per step: mean over the captured launches; the others (%s) serve several clouds / phases and are listed per launch in capture order" % ", ".join(MULTI_ROLE),
       "dram_bytes_per_launch": {k: (v if k in MULTI_ROLE else sum(v) / len(v)) for k, v in acc.items()}}
# the grid-build kernels run once per cloud and step; the four clouds differ in size by an order of magnitude each, so the roles
# are recognisable from the traffic itself: corner_last < surf_last < map_corner < map_surf
for k in ("grid_count", "grid_scan", "grid_fill"):
    v = out["dram_bytes_per_launch"].get(k)
    if isinstance(v, list) and len(v) >= 4:
        roles = ("corner_last", "surf_last", "map_corner", "map_surf")
        per_step = [sorted(v[i:i + 4]) for i in range(0, len(v) - len(v) % 4, 4)]
        for r, name in enumerate(roles):
            out["dram_bytes_per_launch"]["%s_%s" % (k, name)] = sum(p[r] for p in per_step) / len(per_step)
print(json.dumps(out, indent=1))
