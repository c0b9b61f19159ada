"""Forward time of XiangqiNet(128,6) on the bf16 kernels as a function of the batch (net.run(n)): the latency floor that
bounds evaluation (32 games per GPU) and the tail of self-play.  Prints ms per forward, 50 back-to-back launches each,
and the per-launch times of one forward at the smallest batch (CUDA events around each xq_net_gemm via set_timing)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
import torch
import xq_native
import model as M

eng = xq_native.Engine(0)
torch.manual_seed(1)
net = M.B200Net(eng, M.XiangqiNet(128, 6).eval(), max_batch=4096)
for _ in range(5):
    net.run()
torch.cuda.synchronize()
for n in (1, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
    for _ in range(3):
        net.run(n)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50):
        net.run(n)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 50
    print(f"batch {n:5d}: {ms * 1000:8.1f} us per forward, {n / ms / 1000:8.3f} M boards/s")
