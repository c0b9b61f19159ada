"""Per-kernel table of one training step from an ncu launch list (gpu__time_duration.sum, --csv).
    python profiles/tools/launch_table.py gpurun_out/r2_train_launches.csv"""
import collections
import csv
import sys

with open(sys.argv[1]) as f:
    rows = list(csv.DictReader([l for l in f if not l.startswith("==")]))
adam = [i for i, r in enumerate(rows) if "adam_kernel" in r["Kernel Name"]]
step = rows[adam[-2] + 1:adam[-1] + 1]                       # the launches between two optimiser steps
agg = collections.defaultdict(lambda: [0, 0.0])
for r in step:
    name = r["Kernel Name"].split("(")[0].split("::")[-1][-48:]
    agg[name][0] += 1
    agg[name][1] += float(r["Metric Value"]) / 1000.0
tot = sum(v[1] for v in agg.values())
print(f"{len(step)} launches, {tot:.1f} us of kernel time (ncu: cold caches, serialised)")
print("| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1] / v[0]:.1f} | {v[1]:.1f} | {100 * v[1] / tot:.1f} % |")
if len(sys.argv) > 2:
    for r in step:
        print(r["Grid Size"], f"{float(r['Metric Value']) / 1000:.1f}", r["Kernel Name"][:40])
