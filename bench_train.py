"""bench.py --workload train: the training step of BASELINE configs[4] (train.py:376-447) -- batch 256, Adam lr 2e-3
wd 1e-4, clip 1.0, XiangqiNet(128,6) -- on a device-resident replay ring of synthetic self-play records.

Step = one optimiser step on one GLOBAL minibatch of 256 samples (split across the ranks under torchrun: strong
scaling, the reference's batch size is kept; global-minibatch BatchNorm (DPBatchNorm2d) + gradient all-reduce overlapped with backward).
Metric = training samples per second.  The dominant HAND-WRITTEN kernels of the step are HBM-bound streaming
kernels (clip+Adam over the flat buffers: 28 B per parameter; loss+gradient: 64.8 KB per sample); the roofline line
reports the Adam kernel, the conv forward/backward itself is torch/cuDNN (library code, not claimed).
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
    eng.set_timing(False)
    k_ms = {k: float(np.median(v)) for k, v in ks.items()}

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
    line = {
        "metric": "train_samples_per_sec", "value": BATCH / (ms_per_step * 1e-3), "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (device-generated random-playout records)",
        "config": {"workload": f"train: configs[4] training step, global batch {BATCH}, XiangqiNet({CHANNELS},{BLOCKS}) fp32 (TF32 convs as "
                               f"torch defaults), Adam lr 2e-3 wd 1e-4, clip 1.0, replay ring of {n} logical samples in HBM",
                   "parameters": nparam, "l2": "flat optimiser buffers 4 x %.0f MB > 126 MB L2" % (nparam * 4 / 1e6),
                   "parallelism": f"dp{world}: minibatch split across ranks, global-minibatch BatchNorm statistics, gradient all-reduce overlapped with backward"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                     "traffic": None, "peak_source": peak_kind, "kernel": "adam_kernel (clip + weight decay + Adam over flat buffers)",
                     "algorithmic_bytes_per_parameter": 28, "kernel_ms": k_ms,
                     "loss_kernel_gbs": BATCH * 8100 * 8 / (k_ms["loss"] * 1e-3) / 1e9,
                     "sumsq_kernel_gbs": nparam * 4 / (k_ms["sumsq"] * 1e-3) / 1e9,
                     "note": "the conv forward/backward of the step is torch/cuDNN (library); hand-written: batch builder, loss+gradient, "
                             "gradient norm, clip+Adam"},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_samples / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": BATCH * 8, "d2h_bytes_per_step": 16,
                "policy_loss": st.get("policy_loss")},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    bench.emit(line)


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
