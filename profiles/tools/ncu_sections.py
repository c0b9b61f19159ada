#!/usr/bin/env python
"""Per-section summary of one kernel from an .ncu-rep source page: warp-instructions, lanes, stall samples by reason.
usage: ncu_sections.py report.ncu-rep sections.json     (sections: {"name": ["file-prefix", first_line, last_line], ...})"""
import csv, io, json, subprocess, sys

rep, spec = sys.argv[1], json.load(open(sys.argv[2]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur, hdr = None, None
acc = {}
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    line = int(r[0])
    name = "other"
    for k, (pref, lo, hi) in spec.items():
        if cur and cur.startswith(pref) and lo <= line <= hi:
            name = k
            break
    a = acc.setdefault(name, {})
    for i, h in enumerate(hdr):
        if h in ("# Samples", "Instructions Executed", "Thread Instructions Executed") or (h.startswith("stall_") and "Not Issued" not in h):
            try:
                a[h] = a.get(h, 0) + float(r[i])
            except ValueError:
                pass
tot_s = sum(a.get("# Samples", 0) for a in acc.values())
tot_i = sum(a.get("Instructions Executed", 0) for a in acc.values())
print(f"{'section':16} {'instr%':>7} {'lanes':>6} {'samples%':>9}  top stalls")
for k, a in sorted(acc.items(), key=lambda kv: -kv[1].get("# Samples", 0)):
    st = sorted(((v, h[6:]) for h, v in a.items() if h.startswith("stall_")), reverse=True)[:5]
    ins = a.get("Instructions Executed", 0)
    print(f"{k:16} {100*ins/tot_i:7.1f} {a.get('Thread Instructions Executed',0)/max(ins,1):6.1f} {100*a.get('# Samples',0)/tot_s:9.1f}  "
          + ", ".join(f"{n} {100*v/max(a.get('# Samples',1),1):.0f}%" for v, n in st))
