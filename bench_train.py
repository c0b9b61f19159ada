"""bench.py --workload train: the training step of BASELINE configs[4] (train.py:376-447) -- batch 256, Adam lr 2e-3
wd 1e-4, clip 1.0, XiangqiNet(128,6) -- on a device-resident replay ring of synthetic self-play records.

Step = one optimiser step on one GLOBAL minibatch of 256 samples.  One GPU (and dp_mode "replicate"): every layer, the loss
and every gradient run on the hand-written kernels of csrc/xq_tnet.cu (tnet.HandStep: tf32 tcgen05 contractions, plane-layout
BatchNorm / ReLU / residual kernels, replayed from a CUDA graph) followed by the clip + Adam kernels; XQ_TRAIN_HAND=0 runs the
round-1 step through torch / cuDNN / cuBLAS for comparison.  Under torchrun with dp_mode "shard" the minibatch is split
across the ranks (strong scaling, torch modules, peer-memory BatchNorm, gradient all-reduce overlapped with backward).
Metric = training samples per second.  Roofline line: the dominant kernel of the hand-written step, the tf32 implicit-GEMM
convolution tg_kernel (tensor bound; peak = half the measured bf16 rate, the tf32:bf16 ratio of the tensor core); the
HBM-bound Adam kernel is reported next to it.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))

BATCH = 256
CHANNELS = int(os.environ.get("XQ_BENCH_CHANNELS", 128))
BLOCKS = int(os.environ.get("XQ_BENCH_BLOCKS", 6))
RECORDS = 25000              # 50 000 logical samples = the reference's max_buffer_size


def synthetic_records(eng, n, seed):
    """n sample records from device-generated random-playout positions with random visit distributions."""
    import torch
    boards, sides, _, _ = eng.random_playouts(seed, n // 150 + 8)
    boards, sides = boards[:n].contiguous(), sides[:n].contiguous()
    acts, cnt, _, _ = eng.movegen(boards, sides)
    g = torch.Generator(device=eng.dev).manual_seed(seed)
    p = torch.rand((n, 128), generator=g, device=eng.dev)
    p = p * (torch.arange(128, device=eng.dev)[None, :] < cnt[:, None].long())
    p = p / p.sum(dim=1, keepdim=True).clamp_min(1e-9)
    rec = torch.zeros((n, 896), dtype=torch.uint8, device=eng.dev)
    rec[:, :90] = boards.view(torch.uint8)
    rec[:, 90] = sides.view(torch.uint8)
    rec[:, 91] = cnt
    rec[:, 128:384] = acts.contiguous().view(torch.uint8).reshape(n, 256)
    rec[:, 384:896] = p.float().contiguous().view(torch.uint8).reshape(n, 512)
    z = torch.randint(-1, 2, (n,), generator=g, device=eng.dev).float()
    keep = cnt > 0
    return rec[keep], z[keep]


def run(args, rank, world, local_rank, dist):
    import numpy as np
    import torch
    import bench
    sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
    import train as T
    from replay import policy_value_loss

    torch.cuda.set_device(local_rank)
    cfg = T.TrainingConfig()
    cfg.num_channels, cfg.num_res_blocks, cfg.batch_size = CHANNELS, BLOCKS, BATCH
    cfg.checkpoint_dir = "/tmp/xq_bench_train"
    cfg.hand_step = os.environ.get("XQ_TRAIN_HAND", "1") != "0"       # A/B switch: 0 = the torch / cuDNN / cuBLAS step of round 1
    if os.environ.get("XQ_TRAIN_DP_MODE"):
        cfg.dp_mode = os.environ["XQ_TRAIN_DP_MODE"]
    torch.manual_seed(20261018)
    tr = T.AlphaZeroTrainer(cfg)
    eng = tr.eng
    rec, z = synthetic_records(eng, RECORDS, 20261018)            # same records on every rank (same seed)
    tr.replay_buffer.append_raw(rec, z)
    n = len(tr.replay_buffer)
    tr.current_model.train()
    gen = torch.Generator().manual_seed(1)

    def step():
        gidx = torch.randint(0, n, (BATCH,), generator=gen)
        mine = tr._shard(gidx)
        hb = tr._hand.buffers(int(mine.numel())) if tr._hand is not None else None
        states, target, zz = tr.replay_buffer.batch(mine, out=(hb.states, hb.act, hb.prob, hb.n, hb.z) if hb else None)   # H2D: the index list
        if tr._hand is not None:                                    # the hand-written step (tnet.HandStep, csrc/xq_tnet.cu)
            pl, vl = tr._hand.step(states, target[0], target[1], target[2], zz, 1.0 / BATCH)
        else:
            logits, values = tr.current_model(states)
            pl, vl = policy_value_loss(eng, logits, values, target, zz, global_batch=BATCH)
            tr.optimizer.zero_grad()
            (pl + vl).backward()
        tr.optimizer.step()
        return pl, vl

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    eng.launch_count(reset=True)
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count()
    if tr._hand is not None:
        launches += args.steps * tr._hand.buffers(BATCH).launches     # the kernels replayed from the CUDA graph, counted at capture
    ms = e0.elapsed_time(e1)

    # dominant hand-written kernels in isolation (CUDA events on the launching stream, buffers >> L2: 4 x 100 MB)
    opt = tr.optimizer
    eng.set_timing(True)
    ks = {"adam": [], "sumsq": [], "loss": [], "batch": []}
    for _ in range(5):
        pl, vl = step()
        torch.cuda.synchronize()
        ks["adam"].append(eng.last_kernel_ms())                     # the last timed xq_* call of a step is xq_adam_step
    for _ in range(5):
        eng._check(eng.L.xq_grad_sumsq(eng.h, opt.flat_g.data_ptr(), opt.n, opt.partial.data_ptr(), int(opt.partial.numel()),
                                       opt.sumsq.data_ptr(), eng._stream()))
        torch.cuda.synchronize()
        ks["sumsq"].append(eng.last_kernel_ms())
        idx = torch.randint(0, n, (BATCH,), generator=gen)
        states, target, zz = tr.replay_buffer.batch(idx)
        torch.cuda.synchronize()
        ks["batch"].append(eng.last_kernel_ms())
        lg = torch.randn(BATCH, 8100, device=eng.dev)
        vv = torch.zeros(BATCH, 1, device=eng.dev)
        policy_value_loss(eng, lg, vv, target, zz)
        torch.cuda.synchronize()
        ks["loss"].append(eng.last_kernel_ms())
    k_ms = {k: float(np.median(v)) for k, v in ks.items()}
    conv = None
    if tr._hand is not None:
        # the dominant kernel of the hand-written step alone: one 3x3 tower convolution (fprop) at batch 256, 20 launches
        import tnet
        hs = tr._hand
        hb = hs.buffers(BATCH)
        kb, cc = CHANNELS // 32, CHANNELS // 4
        eng.set_timing(False)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):
            t0.record()
            for _ in range(20):
                tnet.tgemm(eng, hb.Ap[0], hb.R, kb, hs.img_f[1], 9, kb, False, hb.pairs, CHANNELS // 128, hb.n_rows, out=hb.Y[1], out_rows=hb.R,
                           out_chunks=cc)
            t1.record()
            torch.cuda.synchronize()
        conv_ms = t0.elapsed_time(t1) / 20
        conv_flops = 2.0 * 90 * 9 * CHANNELS * CHANNELS * BATCH          # algorithmic: real cells only (the 110-row layout executes 22 % more)
        conv = {"ms_per_launch": conv_ms, "algorithmic_flops_per_launch": conv_flops, "achieved_tflops": conv_flops / (conv_ms * 1e-3) / 1e12}
    eng.set_timing(False)

    # e2e: the public call a user makes -- AlphaZeroTrainer.train_network() over one epoch of a 4096-sample buffer;
    # per step the host sends the minibatch index list, at the end it reads the loss statistics back
    small = T.AlphaZeroTrainer.__new__(T.AlphaZeroTrainer)
    small.__dict__.update(tr.__dict__)
    from replay import DeviceReplayBuffer
    small.replay_buffer = DeviceReplayBuffer(eng, 4096)
    small.replay_buffer.append_raw(rec[:2048], z[:2048])
    small.config = T.TrainingConfig()
    small.config.__dict__.update(cfg.__dict__)
    small.config.num_epochs, small.config.min_buffer_size = 1, 1
    small.train_network()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = small.train_network()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_samples = len(small.replay_buffer)

    if world > 1:
        t = torch.tensor([ms, e2e_s], device=eng.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = t.tolist()
    if rank != 0:
        return
    peaks, peak_kind = bench.measured_peaks()
    ms_per_step = ms / args.steps
    nparam = opt.n
    adam_bytes = 28.0 * nparam                                       # read p, g, m, v; write p, m, v (float32)
    achieved = adam_bytes / (k_ms["adam"] * 1e-3) / 1e9
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_train_rate(os.cpu_count() or 1)
    hand = tr._hand is not None
    tf32_peak = peaks["bf16_tflops"] / 2.0                             # tf32 runs at half the bf16 rate on the tensor core (4096 vs 8192 ops/clk/SM)
    roof_adam = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                 "kernel": "adam_kernel (clip + weight decay + Adam over flat buffers)", "algorithmic_bytes_per_parameter": 28}
    line = {
        "metric": "train_samples_per_sec", "value": BATCH / (ms_per_step * 1e-3), "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (device-generated random-playout records)",
        "config": {"workload": f"train: configs[4] training step, global batch {BATCH}, XiangqiNet({CHANNELS},{BLOCKS}), "
                               + ("hand-written step (tf32 tcgen05 contractions, fp32 storage and accumulation, CUDA graph)" if hand else
                                  "torch modules (fp32 cuDNN / cuBLAS), hand-written BatchNorm / loss / Adam")
                               + f", Adam lr 2e-3 wd 1e-4, clip 1.0, replay ring of {n} logical samples in HBM",
                   "parameters": nparam, "l2": "flat optimiser buffers 4 x %.0f MB > 126 MB L2" % (nparam * 4 / 1e6),
                   "parallelism": (f"replicas x{world}: every rank runs the whole minibatch, no collective in the step" if hand else
                                   f"dp{world}: minibatch split across ranks, global-minibatch BatchNorm statistics, gradient all-reduce overlapped with backward")},
        "roofline": ({"bound": "tensor", "achieved": conv["achieved_tflops"], "peak": tf32_peak, "unit": "TFLOP/s",
                      "frac": conv["achieved_tflops"] / tf32_peak, "traffic": 15.05e6,
                      "peak_source": peak_kind + " bf16 burst / 2 (tf32 rate of the tensor core)",
                      "kernel": "tg_kernel: 3x3 tower convolution fprop / dgrad, tf32 tcgen05 implicit GEMM (24 launches per step)",
                      "dominant_kernel": conv, "second_kernel": roof_adam, "kernel_ms": k_ms,
                      "note": "110 work items of 256 rows on 148 SMs, one per CTA: the tensor pipe runs at the full tf32 rate while MMAs "
                              "are in flight (ncu: 18.4 k tensor cycles = the algorithmic count) and the rest of the launch is prologue, "
                              "first loads and the epilogue of the single item (profiles/r2_train_ncu.md); traffic = dram read + write of "
                              "one launch from ncu --set full"}
                     if hand else
                     dict(roof_adam, traffic=None, peak_source=peak_kind, kernel_ms=k_ms,
                          loss_kernel_gbs=BATCH * 8100 * 8 / (k_ms["loss"] * 1e-3) / 1e9,
                          sumsq_kernel_gbs=nparam * 4 / (k_ms["sumsq"] * 1e-3) / 1e9,
                          note="the conv forward/backward of this arm is torch/cuDNN (library); hand-written: batch builder, loss+gradient, "
                               "BatchNorm, gradient norm, clip+Adam")),
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_samples / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": BATCH * 8, "d2h_bytes_per_step": 16,
                "policy_loss": st.get("policy_loss")},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    bench.emit(line)


def train_block(local_rank=0, steps=30, torch_steps=10):
    """The f1 line of the default bench: one optimiser step of configs[4]'s training (batch 256, XiangqiNet(128,6)) on the
    hand-written kernels (tnet.HandStep + clip/Adam), and the same step through torch / cuDNN / cuBLAS on the same box."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
    import train as T
    from replay import policy_value_loss
    out = {}
    for name, hand, k in (("hand_written", True, steps), ("torch_cudnn", False, torch_steps)):
        cfg = T.TrainingConfig()
        cfg.num_channels, cfg.num_res_blocks, cfg.batch_size = CHANNELS, BLOCKS, BATCH
        cfg.checkpoint_dir = "/tmp/xq_bench_train"
        cfg.hand_step = hand
        torch.manual_seed(20261018)
        tr = T.AlphaZeroTrainer(cfg)
        eng = tr.eng
        rec, z = synthetic_records(eng, 4096, 20261018)
        tr.replay_buffer.append_raw(rec, z)
        n = len(tr.replay_buffer)
        tr.current_model.train()
        gen = torch.Generator().manual_seed(1)

        def step():
            idx = torch.randint(0, n, (BATCH,), generator=gen)
            hb = tr._hand.buffers(BATCH) if tr._hand is not None else None
            states, target, zz = tr.replay_buffer.batch(idx, out=(hb.states, hb.act, hb.prob, hb.n, hb.z) if hb else None)
            if tr._hand is not None:
                pl, vl = tr._hand.step(states, target[0], target[1], target[2], zz, 1.0 / BATCH)
            else:
                logits, values = tr.current_model(states)
                pl, vl = policy_value_loss(eng, logits, values, target, zz, global_batch=BATCH)
                tr.optimizer.zero_grad()
                (pl + vl).backward()
            tr.optimizer.step()
            return pl

        for _ in range(4):
            step()
        torch.cuda.synchronize()
        l0 = eng.launch_count(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            pl = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / k
        own = eng.launch_count() / k                                  # host-issued launches of this repo's kernels (batch builder, clip, Adam, ...)
        if tr._hand is not None:
            own += tr._hand.buffers(BATCH).launches                   # + the kernels inside the replayed CUDA graph (counted when it was captured)
        out[name] = {"ms_per_step": ms, "samples_per_s": BATCH / (ms * 1e-3), "steps": k, "policy_loss_last": float(pl),
                     "own_kernel_launches_per_step": own}
        del tr
        torch.cuda.empty_cache()
    out["speedup"] = out["torch_cudnn"]["ms_per_step"] / out["hand_written"]["ms_per_step"]
    out["workload"] = (f"configs[4] training step: global batch {BATCH}, XiangqiNet({CHANNELS},{BLOCKS}), Adam lr 2e-3 wd 1e-4, clip 1.0; "
                       "hand_written = every layer, the loss and every gradient on csrc/xq_tnet.cu (tf32 tcgen05 contractions, fp32 "
                       "elsewhere) replayed from a CUDA graph + clip/Adam kernels; torch_cudnn = the same step through the torch modules "
                       "(fp32 cuDNN / cuBLAS, hand-written BatchNorm / loss / Adam) as in round 1")
    return out


def cpu_train_rate(threads, steps=2):
    """The reference's train_network inner loop (train.py:397-423) on the host cores: torch CPU, dense targets."""
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
    from model import XiangqiNet
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    net = XiangqiNet(CHANNELS, BLOCKS).train()
    opt = torch.optim.Adam(net.parameters(), lr=0.002, weight_decay=1e-4)
    x = (torch.rand(BATCH, 15, 10, 9) < 0.05).float()
    pi = torch.softmax(torch.randn(BATCH, 8100), 1)
    z = torch.randint(-1, 2, (BATCH, 1)).float()

    def one():
        lg, v = net(x)
        loss = -torch.mean(torch.sum(pi * F.log_softmax(lg, dim=1), dim=1)) + F.mse_loss(v, z)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
    one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    wall = time.perf_counter() - t0
    return {"value": steps * BATCH / wall, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"{steps} optimiser steps of batch {BATCH}, fp32 torch XiangqiNet({CHANNELS},{BLOCKS}) on {threads} threads "
                      f"(the loop of train.py:397-423 with resident dense tensors), {wall:.1f} s"}


def run_reference(args):
    import bench
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    vals = []
    last = None
    for i in range(args.warmup + args.steps):
        last = cpu_train_rate(cores, steps=1)
        if i >= args.warmup:
            vals.append(last["value"])
    value = sum(vals) / len(vals)
    last["value"] = value
    bench.emit({"impl": "reference", "metric": "train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * BATCH / value, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"train: configs[4] training step on host cores, batch {BATCH}, XiangqiNet({CHANNELS},{BLOCKS})"},
                "cpu_baseline": last, "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0})
