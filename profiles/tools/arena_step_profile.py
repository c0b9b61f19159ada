"""Evaluation arena of configs[4] on one GPU (32 games, 100 simulations, XiangqiNet(128,6) x 2): wall time of a few plies
(CUDA events) -- and, under `ncu --metrics gpu__time_duration.sum`, the launch list of its lockstep steps.
    python profiles/tools/arena_step_profile.py [plies]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
import ctypes as C
import torch
import xq_native
import model as M
from arena import Arena
from selfplay_engine import SelfPlayEngine

plies = int(sys.argv[1]) if len(sys.argv) > 1 else 8
eng = xq_native.Engine(0)
torch.manual_seed(3)
a, b = M.XiangqiNet(128, 6).eval(), M.XiangqiNet(128, 6).eval()
ar = Arena(eng, a, b, 32, 100)
ar.sp.reset()
cfg = SelfPlayEngine.make_config(dict(num_simulations=100, c_puct=1.5, max_game_length=200, random_opening_moves=0, enable_resign=False),
                                 32, seed=0, add_noise=False, leaves_per_game=1)
def play(k):
    eng._check(eng.L.xq_arena_play(eng.h, C.byref(cfg), C.byref(ar.plan_new), C.byref(ar.plan_old), k, ar.move_log.data_ptr(), eng._stream()))
play(2)
torch.cuda.synchronize()
import time
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
t0 = time.perf_counter()
play(plies)
host_ms = (time.perf_counter() - t0) * 1e3          # time the host needs to ENQUEUE the plies (no synchronisation inside)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"{plies} plies x 101 lockstep steps of 32 games: {ms:.1f} ms = {ms / plies / 101 * 1000:.1f} us per step on the device; "
      f"the host enqueues them in {host_ms:.1f} ms = {host_ms / plies / 101 * 1000:.1f} us per step")
