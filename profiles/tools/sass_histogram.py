"""SASS opcode histogram of libxq_b200.so per kernel (cuobjdump -sass | c++filt): the opcodes that prove tcgen05 / TMEM /
bulk copies / mbarriers / programmatic dependent launch.    python profiles/tools/sass_histogram.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
so = os.path.join(ROOT, "xiangqi-alphazero_b200", "libxq_b200.so")
sass = subprocess.run(f"cuobjdump -sass {so} | c++filt", shell=True, capture_output=True, text=True).stdout
KEEP = re.compile(r"^(UTC\w+|LDTM|STTM|UBLKCP|UTMA\w+|SYNCS|ELECT|R2UR|BAR|ATOMG|ATOMS|RED|PREEXIT|ACQBULK|UCGABAR\w*|CGABAR\w*|MEMBAR|HMMA|FENCE)")
print("# SASS opcode histogram of xiangqi-alphazero_b200/libxq_b200.so (cuobjdump -sass, sm_100a), per kernel")
print("# UTCHMMA = tcgen05.mma (kind::f16 and kind::tf32 alike), UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, LDTM = tcgen05.ld,")
print("# UBLKCP = cp.async.bulk (1-D bulk copy through the TMA engine; no tensor maps: UTMALDG does not occur),")
print("# SYNCS.* = mbarrier ops, ELECT = elect.sync, R2UR = register -> uniform register moves,")
print("# PREEXIT = griddepcontrol.launch_dependents, ACQBULK = griddepcontrol.wait (programmatic dependent launch)\n")
name, count, hist = None, 0, collections.Counter()


def flush():
    if name:
        print(f"{name}   [{count} instructions]")
        print("    " + (", ".join(f"{k} {v}" for k, v in sorted(hist.items())) or "-"))


for line in sass.splitlines():
    m = re.match(r"\s*Function : (.*)", line)
    if m:
        flush()
        name, count, hist = m.group(1).split("(")[0], 0, collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        count += 1
        op = m.group(1)
        if KEEP.match(op):
            hist[".".join(op.split(".")[:3]) if op.startswith("SYNCS") else op.split(".")[0] if not op.startswith(("UTC", "BAR")) else ".".join(op.split(".")[:2])] += 1
flush()
