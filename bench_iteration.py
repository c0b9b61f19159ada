"""bench.py --workload iteration: BASELINE configs[4] -- one full AlphaZero iteration on the GPU(s):
self-play -> mirror augmentation (by index) -> data-parallel training -> evaluation arena -> promotion + broadcast.

Configuration = the reference's `standard_train` preset (train.py:677-689: XiangqiNet(128,6), 200 simulations per move,
max_game_length 300, random_opening_moves 6, resign on; batch 256, 5 epochs, Adam lr 2e-3 wd 1e-4, clip 1.0,
buffer 50 000, eval 100 simulations, promote at >= 0.55) with the number of games scaled to the machine:
1024 self-play games and 32 evaluation games PER GPU (weak scaling in games; the training step keeps the reference's
global batch of 256, split across the ranks).  Step = one iteration; metric = seconds per iteration (lower is better);
the line also carries the phase times and self-play games / simulations per second.
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
GAMES_PER_GPU = int(os.environ.get("XQ_BENCH_ITER_GAMES", 1024))
EVAL_PER_GPU = int(os.environ.get("XQ_BENCH_ITER_EVAL", 32))


def run(args, rank, world, local_rank, dist):
    import torch
    import bench
    sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
    import train as T

    torch.cuda.set_device(local_rank)
    cfg = T.standard_train()
    cfg.num_games_per_iter = GAMES_PER_GPU * world
    cfg.eval_games = EVAL_PER_GPU * world
    cfg.selfplay_slots = GAMES_PER_GPU
    cfg.dp_mode = os.environ.get("XQ_BENCH_DP_MODE", "auto")                        # auto | shard (NCCL data parallel) | replicate
    cfg.hand_step = os.environ.get("XQ_TRAIN_HAND", "1") != "0"
    cfg.selfplay_leaves_per_game = int(os.environ.get("XQ_BENCH_SP_LEAVES", 1))     # > 1: opt-in virtual-loss search
    cfg.eval_leaves_per_game = int(os.environ.get("XQ_BENCH_EVAL_LEAVES", 1))
    cfg.checkpoint_dir = "/tmp/xq_bench_iter"
    torch.manual_seed(20261018)
    tr = T.AlphaZeroTrainer(cfg)
    eng = tr.eng

    def barrier():
        if world > 1:
            dist.barrier()

    def iteration():
        t = {}
        torch.cuda.synchronize(); barrier(); t0 = time.perf_counter()
        sp = tr.self_play()
        torch.cuda.synchronize(); barrier(); t1 = time.perf_counter()
        st = tr.train_network()
        torch.cuda.synchronize(); barrier(); t2 = time.perf_counter()
        ev = tr.evaluate()
        torch.cuda.synchronize(); barrier(); t3 = time.perf_counter()
        t.update(self_play_s=t1 - t0, train_s=t2 - t1, eval_s=t3 - t2, total_s=t3 - t0, games=sp["games"],
                 samples=sp["new_samples"], avg_plies=sp["avg_steps"], buffer=sp["buffer_size"], policy_loss=st.get("policy_loss"),
                 win_rate=ev["win_rate"])
        return t

    for _ in range(max(1, args.warmup)):
        iteration()
    eng.launch_count(reset=True)
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    e0.record()
    its = [iteration() for _ in range(args.steps)]
    e1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count()
    if world > 1:
        t = torch.tensor([ms], device=eng.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank != 0:
        return
    n = len(its)
    mean = lambda k: sum(i[k] for i in its) / n
    sims = sum(i["games"] * i["avg_plies"] for i in its) * cfg.num_simulations
    sp_s = sum(i["self_play_s"] for i in its)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import bench_selfplay
        import bench_train
        c_sp = bench_selfplay.cpu_selfplay_rate(os.cpu_count() or 1, sims=50)
        c_tr = bench_train.cpu_train_rate(os.cpu_count() or 1, steps=1)
        tr_samples = cfg.num_epochs * mean("buffer")
        proj = sims / n / c_sp["value"] + tr_samples / c_tr["value"]
        cpu = {"value": proj, "unit": "s/iteration", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"PROJECTED from measured host rates: {c_sp['value']:.0f} MCTS sims/s ({c_sp['sample']}) and "
                         f"{c_tr['value']:.0f} training samples/s ({c_tr['sample']}) applied to this iteration's "
                         f"{sims / n:.3g} simulations and {tr_samples:.0f} training samples (evaluation not counted)"}
    line = {
        "metric": "iteration_seconds", "value": ms / 1e3 / args.steps, "unit": "s/iteration", "n_gpus": world, "steps": args.steps,
        "warmup": max(1, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16 (self-play/evaluation forward) + f32 (training)", "data": "synthetic (random-init weights, self-generated games)",
        "config": {"workload": f"iteration: configs[4], standard_train preset, {GAMES_PER_GPU} self-play games + {EVAL_PER_GPU} evaluation games per GPU, "
                               f"XiangqiNet(128,6), 200 sims/move, 5 epochs x batch 256 over a 50 000-sample ring, eval 100 sims",
                   "games_per_iteration": cfg.num_games_per_iter, "eval_games": cfg.eval_games,
                   "parallelism": (f"games and evaluation pairs sharded x{world}; training dp{world} (global-minibatch BatchNorm exchanged over NVLink peer memory, gradient all-reduce split at the "
                                   f"policy FC weight and overlapped with backward)" if tr.dp_mode == "shard" else
                                   f"games and evaluation pairs sharded x{world}; training replicated on every rank (full minibatch, no collective in the step)"),
                   "dp_mode": tr.dp_mode, "train_step": "hand-written kernels (tnet.HandStep)" if tr._hand is not None else "torch modules", "selfplay_leaves_per_game": cfg.selfplay_leaves_per_game,
                   "eval_leaves_per_game": cfg.eval_leaves_per_game},
        "phases": {"self_play_s": mean("self_play_s"), "train_s": mean("train_s"), "eval_s": mean("eval_s"),
                   "games": mean("games"), "avg_plies": mean("avg_plies"), "new_samples": mean("samples"),
                   "selfplay_games_per_s": sum(i["games"] for i in its) / sp_s, "selfplay_sims_per_s": sims / sp_s,
                   "policy_loss": its[-1]["policy_loss"], "win_rate": its[-1]["win_rate"]},
        "roofline": None, "cpu_baseline": cpu,
        "e2e": {"value": ms / 1e3 / args.steps, "unit": "s/iteration", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "the iteration IS the public API call sequence (AlphaZeroTrainer.self_play / train_network / evaluate); samples never leave the device"},
        "gpu_launches": launches, "clocks": clocks,
    }
    bench.emit(line)
