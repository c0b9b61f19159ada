#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 self-play engine (one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload selfplay|movegen] [--impl reference]

Workloads (BASELINE.json `configs`):
  movegen   configs[1]: batched legal-move generation + in-check + feature planes over 1M
            random-playout positions per GPU; metric = legal-move positions/s.
  selfplay  configs[2]: 4096 concurrent self-play games x 800 MCTS simulations/move with the
            128ch x 6 ResNet evaluator; metric = MCTS simulations/s.  (The default.)
  iteration configs[4]: one full iteration (self-play, training, evaluation) of the standard preset
            with 1024 games per GPU; metric = seconds per iteration (bench_iteration.py).
  train     configs[4], training step: global batch 256, Adam, clip, XiangqiNet(128,6) on the
            device-resident replay ring; metric = training samples/s (bench_train.py).

A "step" is one pass of the hot path over one batch of synthetic input already resident in
HBM.  `value` is device-timed (CUDA events, barrier + synchronize on both sides, max over
ranks); `e2e` is the same work through the host-buffer C-ABI call (pinned host buffers, H2D and
D2H inside the timed region).  `--impl reference` times the reference's CPU implementation
on the host cores: for self-play the UNMODIFIED reference (baseline/_ref/training, a git-ignored mirror
of the reference's training/ directory: parallel_self_play() in CPU mode, bench_reference.py), for movegen the
reference's Cython engine compiled as-is (oracle/_ref); the C oracle port is the fallback and a cross-check key.
It is the only mode that executes anything under oracle/ besides `cpu_baseline`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "xiangqi-alphazero_b200")
sys.path.insert(0, PKG)

ALGO_BYTES_MOVEGEN_PLANES = 5563.4   # 90+1 read, 1+1+2*35.2 + 5400 written per position
POSITIONS_PER_GPU = 1_000_000


def movegen_config(world):
    """`config` of the movegen line: identical in the GPU arm and in --impl reference."""
    return {"workload": "movegen: configs[1], ordered legal moves + in-check + fp32 planes over 1M "
                        "device-generated random-playout positions per GPU",
            "positions_per_gpu": POSITIONS_PER_GPU,
            "l2": "outputs 5.66 GB/step >> 126 MB L2, no flush needed", "parallelism": f"games sharded x{world}, no collective"}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi, during the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------
# CPU reference arm (also used for cpu_baseline): the reference's Cython engine on host cores
# ------------------------------------------------------------------------------------------------
def _ref_worker(args):
    """One host process: legal moves + in-check for a slice of positions (the loop of SURVEY 8(d))."""
    boards, sides, use_ref = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import xq_oracle
    t0 = time.perf_counter()
    if use_ref:
        ref = xq_oracle.ref_engine()
        gen, chk = ref.cy_generate_legal_moves, ref.cy_is_in_check
        tot = 0
        for i in range(len(sides)):
            b = boards[i].reshape(10, 9)
            tot += len(gen(b, int(sides[i])))
            chk(b, int(sides[i]))
    else:
        _, n, _, _ = xq_oracle.movegen_batch(boards, sides)
        tot = int(n.sum())
    return time.perf_counter() - t0, tot


def cpu_movegen_rate(n_positions, procs):
    """positions/s of the CPU implementation with `procs` processes on a bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import multiprocessing as mp
    import numpy as np
    import xq_oracle
    use_ref = xq_oracle.ref_engine() is not None
    boards, sides = xq_oracle.random_playout_positions(20261018, n_positions)
    chunks = [(boards[i::procs], sides[i::procs], use_ref) for i in range(procs)]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_ref_worker, [(boards[:64], sides[:64], use_ref)] * procs)   # warm the workers
        t0 = time.perf_counter()
        pool.map(_ref_worker, chunks)
        wall = time.perf_counter() - t0
    return n_positions / wall, ("reference" if use_ref else "port"), wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_core = 60_000 if args.workload == "movegen" else 0
    n = per_core * cores
    rates = []
    kind = "port"
    for i in range(args.warmup + args.steps):
        r, kind, wall = cpu_movegen_rate(n, cores)
        if i >= args.warmup:
            rates.append((r, wall))
    value = sum(r for r, _ in rates) / len(rates)
    ms = 1e3 * sum(w for _, w in rates) / len(rates)
    sample = f"{n} random-playout positions per step ({per_core}/core), legal moves + in-check via the Python wrapper"
    line = {
        "impl": "reference", "metric": "legal_move_positions_per_sec", "value": value, "unit": "positions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": movegen_config(int(os.environ.get("WORLD_SIZE", str(args.gpus)))),
        "cpu_baseline": {"value": value, "unit": "positions/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# GPU arm: movegen workload
# ------------------------------------------------------------------------------------------------
def bench_movegen(args, rank, world, local_rank, dist):
    import numpy as np
    import torch
    import xq_native

    torch.cuda.set_device(local_rank)
    eng = xq_native.Engine(local_rank)
    N = POSITIONS_PER_GPU
    # synthetic input: device-side random legal playouts (different seed per rank = weak scaling)
    boards, sides, _, _ = eng.random_playouts(20261018 + rank, 5600)
    assert boards.shape[0] >= N, boards.shape
    boards, sides = boards[:N].contiguous(), sides[:N].contiguous()
    out = (torch.empty((N, 128), dtype=torch.int16, device=eng.dev), torch.empty((N,), dtype=torch.uint8, device=eng.dev),
           torch.empty((N,), dtype=torch.uint8, device=eng.dev), torch.empty((N, 15, 10, 9), dtype=torch.float32, device=eng.dev))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        eng.movegen(boards, sides, planes=True, out=out)
    torch.cuda.synchronize()
    eng.launch_count(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    torch.cuda.synchronize()
    t_all0 = torch.cuda.Event(enable_timing=True)
    t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record()
    for a, b in evs:
        a.record()
        eng.movegen(boards, sides, planes=True, out=out)   # outputs (5.4 GB) >> L2: every step streams to HBM
        b.record()
    t_all1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count()
    total_ms = t_all0.elapsed_time(t_all1)
    kern_ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    n_moves_mean = float(out[1].float().mean().item())

    # e2e: host buffers (pinned) through the host-pointer C-ABI call, copies inside the timed region
    NE = 262_144
    hb = torch.empty((NE, 90), dtype=torch.int8).pin_memory()
    hs = torch.empty((NE,), dtype=torch.int8).pin_memory()
    hb.copy_(boards[:NE].cpu())
    hs.copy_(sides[:NE].cpu())
    h_out = (torch.empty((NE, 128), dtype=torch.int16).pin_memory().numpy(), torch.empty((NE,), dtype=torch.uint8).pin_memory().numpy(),
             torch.empty((NE,), dtype=torch.uint8).pin_memory().numpy(), torch.empty((NE, 15, 10, 9), dtype=torch.float32).pin_memory().numpy())
    hbn, hsn = hb.numpy(), hs.numpy()
    eng.movegen_host(hbn, hsn, planes=True, out=h_out)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        eng.movegen_host(hbn, hsn, planes=True, out=h_out)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    h2d = NE * 91
    d2h = NE * (256 + 2 + 5400)
    # the same call with the planes as bits (176 instead of 5400 bytes per position over PCIe)
    import numpy as np
    hp_out = h_out[:3] + (torch.empty((NE, 44), dtype=torch.int32).pin_memory().numpy().view(np.uint32),)
    eng.movegen_host(hbn, hsn, planes="packed", out=hp_out)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.movegen_host(hbn, hsn, planes="packed", out=hp_out)
    e2e_packed_s = (time.perf_counter() - t0) / e2e_steps

    # reductions over ranks: max time
    if world > 1:
        t = torch.tensor([total_ms, kern_ms, e2e_s, e2e_packed_s], device=eng.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, kern_ms, e2e_s, e2e_packed_s = t.tolist()
    if rank != 0:
        return
    peaks, peak_kind = measured_peaks()
    ms_per_step = total_ms / args.steps
    value = world * N / (ms_per_step * 1e-3)
    achieved = ALGO_BYTES_MOVEGEN_PLANES * N / (kern_ms * 1e-3) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        per_core = 150_000
        r, kind, wall = cpu_movegen_rate(per_core * cores, cores)
        cpu = {"value": r, "unit": "positions/s", "cores": cores, "kind": kind,
               "sample": f"{per_core * cores} random-playout positions, legal moves + in-check (no planes), {wall:.1f} s"}
    line = {
        "metric": "legal_move_positions_per_sec", "value": value, "unit": "positions/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int8", "data": "synthetic",
        "config": movegen_config(world), "mean_legal_moves": n_moves_mean,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                     "frac": achieved / peaks["hbm_gbs"],
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch over 1M positions (ncu --set full,
                     # profiles/r1_movegen_ncu.md): 0.091 + 5.600 GB (thread kernel), 0.09 + 5.60 GB (warp kernel)
                     "traffic": 5.69e9 if N == 1_000_000 else None, "peak_source": peak_kind,
                     "kernel": {"warp": "movegen_kernel<true> (one warp per board)",
                                "thread": "movegen_tpb_kernel<true> (one thread per board)"}[eng.movegen_impl],
                     "algorithmic_bytes_per_position": ALGO_BYTES_MOVEGEN_PLANES,
                     "kernel_ms": kern_ms},
        "cpu_baseline": cpu,
        "e2e": {"value": world * NE / e2e_s, "unit": "positions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "positions_per_step": NE, "api": "xq_movegen_batch_host (float32 planes: 97 % of the bytes over PCIe)",
                "packed_planes": {"value": world * NE / e2e_packed_s, "unit": "positions/s", "h2d_bytes_per_step": h2d,
                                  "d2h_bytes_per_step": NE * (256 + 2 + 176), "api": "xq_movegen_batch_host_packed"}},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    emit(line)


class _QuietStdout:
    """Route fd 1 to stderr while the benchmark runs (NCCL and other native libraries print banners to
    stdout) and restore it for the single JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def restore(self):
        if self.saved is not None:
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)
            self.saved = None

    def __exit__(self, *exc):
        self.restore()
        return False


QUIET = None


def emit(line: dict):
    """Print the one JSON line on the real stdout."""
    q = QUIET or getattr(sys.modules.get("__main__"), "QUIET", None)   # bench_selfplay imports this file as a module
    if q is not None:
        q.restore()
    print(json.dumps(line), flush=True)


def main():
    global QUIET
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "movegen", "selfplay", "train", "iteration"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "selfplay" if os.path.exists(os.path.join(PKG, "selfplay_engine.py")) else "movegen"
    QUIET = _QuietStdout().__enter__()
    if args.impl == "reference":
        if args.workload == "train":
            sys.path.insert(0, ROOT)
            import bench_train
            return bench_train.run_reference(args)
        if args.workload == "selfplay":
            sys.path.insert(0, ROOT)
            import bench_selfplay
            return bench_selfplay.run_reference(args)
        return run_reference_arm(args)

    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's version/info lines off stdout (one JSON line only)
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.workload == "movegen":
            bench_movegen(args, rank, world, local_rank, dist)
        elif args.workload == "train":
            sys.path.insert(0, ROOT)
            import bench_train
            bench_train.run(args, rank, world, local_rank, dist)
        elif args.workload == "iteration":
            sys.path.insert(0, ROOT)
            import bench_iteration
            bench_iteration.run(args, rank, world, local_rank, dist)
        else:
            sys.path.insert(0, ROOT)
            import bench_selfplay
            bench_selfplay.run(args, rank, world, local_rank, dist)
    finally:
        if dist is not None:
            dist.destroy_process_group()
        if QUIET is not None:
            QUIET.restore()


if __name__ == "__main__":
    main()
