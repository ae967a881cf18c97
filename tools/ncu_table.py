#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export (one row per kernel launch) as a markdown table:
time, DRAM bytes, throughput percentages, occupancy, instruction count and the two dominant stall reasons."""
import csv
import re
import sys


def f(row, idx, name, default=0.0):
    try:
        return float(row[idx[name]].replace(",", ""))
    except Exception:
        return default


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    unit = {h: units[i] for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", h)]

    def to_bytes(name, row):
        v = f(row, idx, name)
        u = unit.get(name, "")
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    def to_us(name, row):
        v = f(row, idx, name)
        u = unit.get(name, "")
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1e-3)

    print("| # | kernel | time us | dram rd MB | dram wr MB | dram % | L2 % | L1 % | SM % | occ % | regs | warp-instr M | top stalls |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for k, r in enumerate(data):
        name = r[idx["Kernel Name"]]
        name = re.sub(r"\(.*", "", name).replace("(anonymous namespace)::", "")
        st = sorted(((f(r, idx, c), re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", c).group(1)) for c in stall_cols), reverse=True)
        top = ", ".join("%s %.1f" % (n, v) for v, n in st[:3] if n not in ("selected",))
        print("| %d | %s | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %.2f | %s |" % (
            k, name, to_us("gpu__time_duration.sum", r), to_bytes("dram__bytes_read.sum", r) / 1e6, to_bytes("dram__bytes_write.sum", r) / 1e6,
            f(r, idx, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), f(r, idx, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            f(r, idx, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"), f(r, idx, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
            f(r, idx, "sm__warps_active.avg.pct_of_peak_sustained_active"), int(f(r, idx, "launch__registers_per_thread")),
            f(r, idx, "smsp__inst_executed.sum") / 1e6, top))


if __name__ == "__main__":
    main(sys.argv[1])
