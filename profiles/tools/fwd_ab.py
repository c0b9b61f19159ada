"""Forward time of XiangqiNet(128,6) at batch 4096 for one setting of the XQ_NET_* switches (read once in xq_create):
run it once per setting on the same box and compare.  Prints ms per forward: 20 launches alone (3 repetitions) and a 400 ms train."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
import torch
import xq_native
import model as M

eng = xq_native.Engine(0)
torch.manual_seed(1)
net = M.B200Net(eng, M.XiangqiNet(int(os.environ.get("XQ_BENCH_CHANNELS", 128)), int(os.environ.get("XQ_BENCH_BLOCKS", 6))).eval(), max_batch=4096)
for _ in range(5):
    net.run()
torch.cuda.synchronize()
res = []
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        net.run()
    b.record()
    torch.cuda.synchronize()
    res.append(a.elapsed_time(b) / 20)
    torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(360):
    net.run()
b.record()
torch.cuda.synchronize()
print({k: os.environ.get(k) for k in ("XQ_NET_FORK", "XQ_NET_2CTA")}, "alone", [round(x, 4) for x in res], "train", round(a.elapsed_time(b) / 360, 4))
