#!/usr/bin/env python
"""Hot spots of an `ncu --page source --csv --print-source sass` export: per kernel, the SASS lines with the most stall
samples / executed instructions, grouped in address order so loops are recognisable.
usage: ncu_hot.py file.csv [kernel-substring] [top-N]"""
import csv
import io
import re
import sys

path = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = open(path).read()
parts = re.split(r'(?m)^(?="Kernel Name")', txt)
for part in parts:
    if not part.strip():
        continue
    rows = list(csv.reader(io.StringIO(part)))
    name = rows[0][1]
    if want not in name:
        continue
    hdr = rows[1]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [r for r in rows[2:] if len(r) > iex]
    tot_s = sum(int(r[isamp] or 0) for r in data)
    tot_e = sum(int(r[iex] or 0) for r in data)
    print("== %s\n   %d SASS lines, %d samples, %.2f M warp-instr" % (name[:100], len(data), tot_s, tot_e / 1e6))
    ranked = sorted(range(len(data)), key=lambda i: -int(data[i][isamp] or 0))[:topn]
    for i in sorted(ranked):
        r = data[i]
        print("  %5d  %5.1f%% samp  %6.2f%% exec  %s" % (i, 100.0 * int(r[isamp] or 0) / max(tot_s, 1), 100.0 * int(r[iex] or 0) / max(tot_e, 1), r[isrc].strip()[:90]))
