#!/usr/bin/env python
"""usage: tools/ncu_sass_hot.py SOURCE_CSV KERNEL_INDEX [top] — per-SASS-line executed instructions and stall samples of one captured
launch from `ncu --page source --csv --print-source sass`; prints the instruction stream with counts (loop bodies stand out)."""
import csv
import sys

csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sys.argv[1])))
k = int(sys.argv[2])
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
a = starts[k]
b = starts[k + 1] if k + 1 < len(starts) else len(rows)
print(rows[a][1])
hdr = rows[a + 1]
ci = hdr.index("Instructions Executed")
cs = hdr.index("# Samples")
body = [r for r in rows[a + 2:b] if len(r) > ci]
tot = sum(int(r[ci] or 0) for r in body)
tots = sum(int(r[cs] or 0) for r in body)
print("warp instructions %d, samples %d, SASS lines %d" % (tot, tots, len(body)))
mode = sys.argv[3] if len(sys.argv) > 3 else "stream"
if mode == "stream":
    for i, r in enumerate(body):
        print("%4d %10s %6s  %s" % (i, r[ci], r[cs], r[1].strip()[:110]))
else:
    for r in sorted(body, key=lambda r: -int(r[cs] or 0))[:int(mode)]:
        print("%10s %6s  %s" % (r[ci], r[cs], r[1].strip()[:110]))
