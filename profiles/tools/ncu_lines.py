#!/usr/bin/env python
"""Per-source-line warp-instruction counts of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep [min_percent] [group_spec.json]"""
import csv, io, json, subprocess, sys

rep = sys.argv[1]
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
for r in rows[2:3]:
    for k in ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
              "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_write.sum", "dram__bytes_read.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers"]:
        if k in h:
            print(f"{k:70} {r[h.index(k)]} {rows[1][h.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None
tot = 0
out = []
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            n = int(r[7]); ti = int(r[8])
        except ValueError:
            continue
        out.append((cur, int(r[0]), n, ti, r[1]))
        tot += n
print(f"total warp instructions (source page): {tot/1e6:.1f} M")
for f, l, n, ti, s in out:
    if n > tot * minpct / 100:
        print(f"{f[:16]:16} {l:4} {n/1e6:8.1f}M {100*n/tot:5.1f}% lanes={ti/max(n,1):4.1f} {s.strip()[:120]}")
if len(sys.argv) > 3:
    groups = json.load(open(sys.argv[3]))
    acc = {k: [0, 0] for k in groups}
    rest = 0
    for f, l, n, ti, s in out:
        for k, (gf, a, b) in groups.items():
            if f == gf and a <= l <= b:
                acc[k][0] += n; acc[k][1] += ti
                break
        else:
            rest += n
    for k, (n, ti) in acc.items():
        print(f"{k:24} {n/1e6:8.1f}M {100*n/tot:5.1f}% lanes {ti/max(n,1):.1f}")
    print(f"{'(other)':24} {rest/1e6:8.1f}M {100*rest/tot:5.1f}%")
