"""bench.py --workload selfplay: BASELINE configs[2] -- 4096 concurrent self-play games x 800 MCTS
simulations per move, XiangqiNet(128, 6) with seeded random-init weights, c_puct 1.5, Dirichlet(0.3)
root noise, T = 1.0 below ply 20 else 0.3, random_opening_moves 6, resign (-0.9, 5 steps)
(train.py standard preset :677-689 + TrainingConfig defaults :64-78), on each GPU.

Step = one ply of every game: root evaluation + 800 x (select, tcgen05 forward, expand/backup) +
move selection, entirely on the device.  Metric = MCTS simulations per second (a simulation = one
iteration of mcts.py:126; terminal-leaf simulations count).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))

GAMES_PER_GPU = 4096
SIMS = 800
CHANNELS = int(os.environ.get("XQ_BENCH_CHANNELS", 128))   # XQ_BENCH_CHANNELS=256 XQ_BENCH_BLOCKS=20: BASELINE configs[3] network
BLOCKS = int(os.environ.get("XQ_BENCH_BLOCKS", 6))
REF_SIMS = 200                      # simulations per process per step in the CPU reference arm (bounded sample)


def flops_per_eval(c, r):
    """conv + linear MACs x 2 of XiangqiNet(c, r) (369 193 216 for 128x6, 4 301 360 896 for 256x20)."""
    return (2 * 90 * 9 * 15 * c + 2 * r * 2 * 90 * 9 * c * c + 2 * 90 * c * 32 + 2 * 2880 * 8100 + 2 * 90 * c * 4
            + 2 * (360 * 128 + 128))


FLOPS_PER_EVAL = flops_per_eval(CHANNELS, BLOCKS)
assert flops_per_eval(128, 6) == 369_193_216 and flops_per_eval(256, 20) == 4_301_360_896


def workload_config(world, games=None, sims=None):
    """`config` of the JSON line: a pure function of the workload and N, printed identically by the GPU arm and by
    --impl reference (the driver compares the two)."""
    games = GAMES_PER_GPU if games is None else games
    sims = SIMS if sims is None else sims
    name = "configs[3] (32768 games over 8 GPUs = 4096 per GPU)" if (CHANNELS, BLOCKS) == (256, 20) else "configs[2]"
    return {"workload": f"selfplay: {name}, {games} concurrent games/GPU x {sims} sims/move, XiangqiNet({CHANNELS},{BLOCKS}), "
                        f"c_puct 1.5, Dirichlet(0.3) root noise, step = one ply of every game",
            "games_per_gpu": games, "sims_per_move": sims,
            "l2": "per-step working set (trees + activations, > 1 GB) >> 126 MB L2",
            "parallelism": f"games sharded x{world}, no collective in self-play"}


class StdConfig:
    """standard_train preset of the reference (train.py:677-689) for the self-play keys."""
    num_simulations = SIMS
    c_puct = 1.5
    temperature_threshold = 20
    max_game_length = 300
    random_opening_moves = 6
    enable_resign = True
    resign_threshold = -0.9
    resign_check_steps = 5


def run(args, rank, world, local_rank, dist):
    import numpy as np
    import torch
    import bench
    import xq_native
    import model as M
    from selfplay_engine import SelfPlayEngine, SAMPLE_BYTES

    sims = int(os.environ.get("XQ_BENCH_SIMS", SIMS))
    games = int(os.environ.get("XQ_BENCH_GAMES", GAMES_PER_GPU))
    leaves = int(os.environ.get("XQ_BENCH_LEAVES", 1))   # > 1: opt-in multi-leaf (virtual loss) search, not the headline configuration
    torch.cuda.set_device(local_rank)
    eng = xq_native.Engine(local_rank)
    torch.manual_seed(20261018)                      # same random-init weights on every rank
    model = M.XiangqiNet(CHANNELS, BLOCKS).eval()
    cfgobj = StdConfig()
    cfgobj.num_simulations = sims
    total_plies = args.warmup + args.steps + 4
    sp = SelfPlayEngine(eng, model, n_slots=games, max_games=games * 4,
                        sample_capacity=games * (2 * total_plies + 8), node_capacity=games * (sims + 1) * 48, leaves_per_game=leaves)
    sp.reset()
    cfg = SelfPlayEngine.make_config(cfgobj, games * 4, seed=20261018 + rank, add_noise=True, leaves_per_game=leaves)

    def barrier():
        if world > 1:
            dist.barrier()

    sp.play(cfg, args.warmup)
    torch.cuda.synchronize()
    c0 = sp.counters()
    eng.launch_count(reset=True)
    sampler = bench.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    e0.record()
    sp.play(cfg, args.steps)                          # leaves live in HBM; the tree/activation working set (>> L2) is rewritten every step
    e1.record()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = eng.launch_count()
    ms = e0.elapsed_time(e1)
    c1 = sp.counters()
    sims_done = c1["sims"] - c0["sims"]
    evals_done = c1["evals"] - c0["evals"]
    if c1["error"]:
        raise RuntimeError(f"device error bits {c1['error']}")

    # forward-only timing of the dominant kernel chain (the evaluator) for the roofline line
    net = sp.net
    for _ in range(3):
        net.run()
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(20):
        net.run()
    f1.record()
    torch.cuda.synchronize()
    fwd_ms = f0.elapsed_time(f1) / 20
    # the dominant kernel alone: one residual-tower conv (layer 1 = first conv of block 0), 20 back-to-back launches
    for _ in range(3):
        net.run_layer(1)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(20):
        net.run_layer(1)
    k1.record()
    torch.cuda.synchronize()
    conv_ms = k0.elapsed_time(k1) / 20
    # the same chain for ~0.4 s: the SM clock settles under the power cap like it does inside a ply
    n_sus = max(50, int(400.0 / max(fwd_ms, 0.05)))
    f0.record()
    for _ in range(n_sus):
        net.run()
    f1.record()
    torch.cuda.synchronize()
    fwd_sus_ms = f0.elapsed_time(f1) / n_sus
    fwd_boards = net.max_batch                                       # = games x leaves_per_game

    # e2e, engine level: the same plies through the host-facing engine calls -- per step the host uploads the step's
    # control block and downloads that ply's sample records and counters (pinned host memory)
    h_raw = torch.empty((games * 2, SAMPLE_BYTES), dtype=torch.uint8).pin_memory().numpy()
    e2e_steps = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    s_before = sp.counters()
    d2h = 0
    for _ in range(e2e_steps):
        sp.play(cfg, 1)
        c = sp.counters()                                   # D2H: counters (synchronises)
        n_new = min(games, c["samples"])
        sp.fetch(c["samples"] - n_new, n_new, out=h_raw[:n_new])   # D2H: this ply's sample records + game results
        d2h += n_new * SAMPLE_BYTES + 14 * 8 + sp.max_games * 3
    e2e_s = time.perf_counter() - t0
    c2 = sp.counters()
    e2e_sims = c2["sims"] - s_before["sims"]

    # e2e through the reference's own entry point: parallel_selfplay.parallel_self_play(model, config) -> (data, stats)
    # with HOST weights in (the fp32 state_dict is folded, converted and uploaded inside the call) and HOST samples out
    # (sparse records downloaded, gathered over the ranks, densified on access).  Whole games from fresh openings,
    # truncated to a ply budget by config.max_game_length like the reference arm; a simulation count is not returned by
    # the contract, so it is counted the way the reference arm counts it: searched plies (samples / 2) x sims per move.
    api = None
    if not os.environ.get("XQ_BENCH_NO_API_E2E"):
        import parallel_selfplay as ps

        class ApiCfg(StdConfig):
            pass
        acfg = ApiCfg()
        acfg.num_simulations = sims
        acfg.max_game_length = 16                             # openings are 0..6 random plies: 10..16 searched plies per game
        acfg.num_games_per_iter = games * world
        acfg.selfplay_leaves_per_game = leaves
        host_model = M.XiangqiNet(CHANNELS, BLOCKS).eval()
        host_model.load_state_dict(model.state_dict())
        wcfg = ApiCfg()                                       # warm-up call (as train.py's second iteration sees it: the engine
        wcfg.num_simulations, wcfg.max_game_length = sims, 2  # of the first call is reused, only the weights are folded again)
        wcfg.num_games_per_iter, wcfg.selfplay_leaves_per_game = games * world, leaves
        ps.parallel_self_play(host_model, wcfg)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        data, stats = ps.parallel_self_play(host_model, acfg)
        dense = data[:256]                                    # the dense (planes, policy[8100], z) tuples of the contract
        api_s = time.perf_counter() - t0
        rec_bytes = int(getattr(data, "records", np.zeros((0, 896), np.uint8)).nbytes)
        wnet = ps._ENGINES[next(iter(ps._ENGINES))].net if ps._ENGINES else None
        h2d_w = int(sum(t.numel() * t.element_size() for t in wnet.keep)) if wnet is not None else 0
        api = {"value": stats["new_samples"] // 2 * sims / api_s, "seconds": api_s, "games": stats["games"],
               "new_samples": stats["new_samples"], "dense_tuples_materialised": len(dense),
               "h2d_bytes": h2d_w + 64, "d2h_bytes": rec_bytes // max(world, 1) + games * 3 + 14 * 8,
               "config": f"{games * world} games, max_game_length 16 (10-16 searched plies after the random opening), {sims} sims/move, after one warm-up call",
               "breakdown": dict(ps.LAST_TIMING)}
        del data, dense
        ps._ENGINES.clear()

    # secondary headline of BASELINE.json ("legal-move positions/sec"): K1 over 1M device-generated positions, with its
    # own end-to-end figure (host buffers through the C-ABI host call: float32 planes, and the 176-byte packed planes)
    mv = None
    try:
        mv = movegen_block(eng, rank, world, args)
    except Exception as ex:   # never let the secondary line break the headline
        mv = {"error": str(ex)[:300]}

    # BASELINE configs[3]'s network (256 channels x 20 blocks) on the same kernels: 1 warm-up + 2 timed plies
    c3 = None
    if (CHANNELS, BLOCKS) == (128, 6) and not os.environ.get("XQ_BENCH_NO_CONFIGS3"):
        try:
            del sp, net
            torch.cuda.empty_cache()
            c3 = configs3_block(rank, world, local_rank, dist, sims)
        except Exception as ex:
            c3 = {"error": str(ex)[:300]}

    # f1: the training step of configs[4] on the hand-written kernels next to the torch/cuDNN step
    f1 = None
    if world == 1 and (CHANNELS, BLOCKS) == (128, 6) and not os.environ.get("XQ_BENCH_NO_TRAIN"):
        try:
            import bench_train
            f1 = bench_train.train_block(local_rank)
        except Exception as ex:
            f1 = {"error": str(ex)[:300]}

    if world > 1:
        t = torch.tensor([ms, e2e_s, fwd_ms, conv_ms, api["seconds"] if api else 0.0], device=eng.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, fwd_ms, conv_ms, api_max = t.tolist()
        cnt = torch.tensor([sims_done, evals_done, e2e_sims], device=eng.dev, dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        sims_done, evals_done, e2e_sims = cnt.tolist()
        if api:
            api["value"] = api["new_samples"] // 2 * sims / api_max        # every rank returns the union of the samples
            api["seconds"] = api_max
    if rank != 0:
        return
    peaks, peak_kind = bench.measured_peaks()
    value = sims_done / (ms * 1e-3)
    peak_tf = peaks["bf16_tflops_sustained"]          # the forward runs inside a seconds-long step: sustained figure
    # achieved = algorithmic FLOPs of every network evaluation in the timed region / the region's device time,
    # i.e. the whole step (search kernels, launch gaps, power-capped clocks) is charged to the tensor kernels
    achieved_tf = FLOPS_PER_EVAL * evals_done / world / (ms * 1e-3) / 1e12
    isolated_tf = FLOPS_PER_EVAL * fwd_boards / (fwd_ms * 1e-3) / 1e12
    conv_flops = 2 * 90 * 9 * CHANNELS * CHANNELS * fwd_boards          # algorithmic: 90 cells x 9 taps x C x C MACs per board
    conv_tf = conv_flops / (conv_ms * 1e-3) / 1e12
    # ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one tower conv at batch 4096 (profiles/r2_net_ncu.md)
    conv_traffic = {(128, 4096): 142.5e6, (256, 4096): None}.get((CHANNELS, fwd_boards))
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, why = reference_selfplay_rate(steps=1, warmup=1, sims=sims, plies=1)
        if cpu is None:
            cpu = cpu_selfplay_rate(os.cpu_count() or 1, sims=min(sims, 100))
            cpu["fallback_reason"] = str(why)[:300]
    line = {
        "metric": "mcts_sims_per_sec", "value": value, "unit": "sims/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic (seeded random-init weights, self-generated games)",
        "config": workload_config(world, games, sims), "evals_in_region": evals_done, "leaves_per_game": leaves,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / peak_tf,
                     "traffic": conv_traffic,
                     "traffic_detail": {"unit": "MB per launch (ncu --set full, dram read + write, batch 4096, 128 channels)",
                                        "conv_kernel<128,8,0,3,0>": 142.5, "conv_kernel<128,8,0,3,0> (+residual)": 250.6,
                                        "fc_kernel<224>": 144.9, "algorithmic conv (in + out, + residual)": [188.7, 283.1],
                                        "note": "the write-back of a layer's output is partly absorbed by the 126 MB L2 before the next layer reads it",
                                        "source": "profiles/r2_net_ncu.md"},
                     "peak_source": peak_kind + " (sustained bf16)",
                     "kernel": f"conv_kernel x{2 * BLOCKS + 2} + fc_kernel (tcgen05 implicit-GEMM forward) + value_head_kernel",
                     "dominant_kernel": {"name": f"conv_kernel<{CHANNELS}->{CHANNELS}, 3x3> (residual tower, {2 * BLOCKS} launches per forward)",
                                         "ms_per_launch": conv_ms, "algorithmic_flops_per_launch": conv_flops,
                                         "achieved_tflops": conv_tf, "frac_of_burst_peak": conv_tf / peaks["bf16_tflops"],
                                         "timing": "20 back-to-back launches alone, CUDA events"},
                     "algorithmic_flops_per_eval": FLOPS_PER_EVAL, "forward_ms_isolated": fwd_ms, "forward_ms_back_to_back_400ms": fwd_sus_ms,
                     "forward_isolated_tflops": isolated_tf, "forward_isolated_frac_of_burst_peak": isolated_tf / peaks["bf16_tflops"],
                     "note": "achieved = FLOPs of all evaluations in the timed region / region time (search kernels and "
                             "power-capped clocks included); the isolated forward is timed back to back for 20 launches"},
        "cpu_baseline": cpu,
        "secondary": mv,
        "extra": {"configs3_256x20": c3, "f1_train_step": f1},
        "e2e": ({"value": api["value"], "unit": "sims/s", "api": "parallel_self_play",
                 "h2d_bytes_per_step": api["h2d_bytes"], "d2h_bytes_per_step": api["d2h_bytes"],
                 "step": "one parallel_self_play(model, config) call: " + api["config"], "seconds": api["seconds"],
                 "games": api["games"], "new_samples": api["new_samples"], "breakdown": api.get("breakdown"),
                 "engine_level": {"value": e2e_sims / e2e_s, "unit": "sims/s", "api": "SelfPlayEngine.play + counters + fetch per ply",
                                  "h2d_bytes_per_step": 64 + 160, "d2h_bytes_per_step": d2h // e2e_steps}}
                if api else
                {"value": e2e_sims / e2e_s, "unit": "sims/s", "api": "SelfPlayEngine.play + counters + fetch per ply",
                 "h2d_bytes_per_step": 64 + 160, "d2h_bytes_per_step": d2h // e2e_steps}),
        "gpu_launches": launches,
        "clocks": clocks,
    }
    bench.emit(line)


def movegen_block(eng, rank, world, args):
    """M2 of BASELINE.json (legal-move positions/s) with kernel rate, roofline share, end-to-end rates and the reference's
    Cython engine on the host cores."""
    import numpy as np
    import torch
    import bench
    N = 1_000_000
    pb, ps_, _, _ = eng.random_playouts(20261018 + rank, 5600)
    pb, ps_ = pb[:N].contiguous(), ps_[:N].contiguous()
    outb = (torch.empty((N, 128), dtype=torch.int16, device=eng.dev), torch.empty((N,), dtype=torch.uint8, device=eng.dev),
            torch.empty((N,), dtype=torch.uint8, device=eng.dev), torch.empty((N, 15, 10, 9), dtype=torch.float32, device=eng.dev))
    for _ in range(3):
        eng.movegen(pb, ps_, planes=True, out=outb)
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0.record()
    for _ in range(5):
        eng.movegen(pb, ps_, planes=True, out=outb)
    m1.record()
    torch.cuda.synchronize()
    mv_ms = m0.elapsed_time(m1) / 5
    gbs = 5563.4 * N / (mv_ms * 1e-3) / 1e9
    mv = {"metric": "legal_move_positions_per_sec", "value": N / (mv_ms * 1e-3), "unit": "positions/s (per GPU)",
          "kernel": {"warp": "movegen_kernel<true>", "thread": "movegen_tpb_kernel<true>"}[eng.movegen_impl]
                    + " (moves + in-check + fp32 planes)", "kernel_ms": mv_ms,
          "roofline": {"bound": "hbm", "achieved": gbs, "peak": bench.measured_peaks()[0]["hbm_gbs"], "unit": "GB/s",
                       "frac": gbs / bench.measured_peaks()[0]["hbm_gbs"], "traffic": 5.69e9,
                       "algorithmic_bytes_per_position": 5563.4}}
    del outb
    # end to end: pinned host buffers through xq_movegen_batch_host[_packed] (H2D + kernel + D2H in the timed region)
    NE = 262_144
    hb = torch.empty((NE, 90), dtype=torch.int8).pin_memory()
    hs = torch.empty((NE,), dtype=torch.int8).pin_memory()
    hb.copy_(pb[:NE].cpu())
    hs.copy_(ps_[:NE].cpu())
    hbn, hsn = hb.numpy(), hs.numpy()
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    base = (pin((NE, 128), torch.int16), pin((NE,), torch.uint8), pin((NE,), torch.uint8))
    for name, planes, pl, d2h in (("fp32_planes", True, pin((NE, 15, 10, 9), torch.float32), NE * (256 + 2 + 5400)),
                                  ("packed_planes", "packed", pin((NE, 44), torch.int32).view(np.uint32), NE * (256 + 2 + 176)),
                                  ("no_planes", False, None, NE * (256 + 2))):
        out = base + (pl,)
        eng.movegen_host(hbn, hsn, planes=planes, out=out)
        t0 = time.perf_counter()
        for _ in range(3):
            eng.movegen_host(hbn, hsn, planes=planes, out=out)
        dt = (time.perf_counter() - t0) / 3
        mv.setdefault("e2e", {})[name] = {"value": NE / dt, "unit": "positions/s", "h2d_bytes_per_step": NE * 91,
                                          "d2h_bytes_per_step": d2h, "positions_per_step": NE}
    mv["e2e"]["api"] = "xq_movegen_batch_host / xq_movegen_batch_host_packed (pinned host buffers, chunked copy/compute overlap)"
    del pb, ps_
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        r, kind, wall = bench.cpu_movegen_rate(40_000 * cores, cores)
        mv["cpu_baseline"] = {"value": r, "unit": "positions/s", "cores": cores, "kind": kind,
                              "sample": f"{40_000 * cores} random-playout positions, cy_generate_legal_moves + cy_is_in_check of the "
                                        f"reference's Cython engine compiled as-is (oracle/_ref), {cores} processes, {wall:.1f} s"}
    return mv


def configs3_block(rank, world, local_rank, dist, sims):
    """BASELINE configs[3]: XiangqiNet(256, 20), 4096 games per GPU (32768 over 8), Dirichlet root noise, 800 sims/move."""
    import torch
    import bench
    import xq_native
    import model as M
    from selfplay_engine import SelfPlayEngine
    C3, R3, games = 256, 20, GAMES_PER_GPU
    eng = xq_native.Engine(local_rank)
    torch.manual_seed(20261019)
    model = M.XiangqiNet(C3, R3).eval()
    cfgobj = StdConfig()
    cfgobj.num_simulations = sims
    sp = SelfPlayEngine(eng, model, n_slots=games, max_games=games * 2, sample_capacity=games * 8, node_capacity=games * (sims + 1) * 48)
    sp.reset()
    cfg = SelfPlayEngine.make_config(cfgobj, games * 2, seed=20261019 + rank, add_noise=True)
    sp.play(cfg, 1)
    torch.cuda.synchronize()
    c0 = sp.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    e0.record()
    sp.play(cfg, 2)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c1 = sp.counters()
    vals = torch.tensor([ms], device=eng.dev, dtype=torch.float64)
    cnt = torch.tensor([c1["sims"] - c0["sims"], c1["evals"] - c0["evals"]], device=eng.dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms = float(vals[0])
    sims_done, evals_done = cnt.tolist()
    peaks, _ = bench.measured_peaks()
    tf = flops_per_eval(C3, R3) * evals_done / world / (ms * 1e-3) / 1e12
    out = {"metric": "mcts_sims_per_sec", "value": sims_done / (ms * 1e-3), "unit": "sims/s", "n_gpus": world,
           "workload": f"configs[3]: {games} games/GPU x {sims} sims/move, XiangqiNet(256,20), Dirichlet(0.3) root noise, 1 warm-up + 2 timed plies",
           "ms_per_ply": ms / 2, "achieved_tflops": tf, "frac_of_sustained_peak": tf / peaks["bf16_tflops_sustained"],
           "frac_of_burst_peak": tf / peaks["bf16_tflops"],
           "note": "the sustained peak is cuBLAS bf16 8192^3 back to back for 4 s under the same power cap; a ratio above 1 means this path holds a higher rate than that GEMM did",
           "algorithmic_flops_per_eval": flops_per_eval(C3, R3), "error_bits": c1["error"]}
    del sp
    eng.close()
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (mcts.py search, fp32 torch forward, one thread per process --
# parallel_selfplay.py:_cpu_worker_entry) restated by the oracle; only the rules engine can be the
# reference's own compiled code (oracle/_ref), mcts.py/model.py cannot travel to the GPU box.
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    sims, seed = args
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    import xq_oracle
    torch.set_num_threads(1)

    class Block(nn.Module):
        def __init__(self, c):
            super().__init__()
            self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False)
            self.bn1 = nn.BatchNorm2d(c)
            self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False)
            self.bn2 = nn.BatchNorm2d(c)

        def forward(self, x):
            return F.relu(self.bn2(self.conv2(F.relu(self.bn1(self.conv1(x))))) + x)

    class Net(nn.Module):   # model.py:39-107
        def __init__(self, c=CHANNELS, r=BLOCKS):
            super().__init__()
            self.inp = nn.Sequential(nn.Conv2d(15, c, 3, padding=1, bias=False), nn.BatchNorm2d(c), nn.ReLU())
            self.blocks = nn.ModuleList([Block(c) for _ in range(r)])
            self.ph = nn.Sequential(nn.Conv2d(c, 32, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU(), nn.Flatten(), nn.Linear(2880, 8100))
            self.vh = nn.Sequential(nn.Conv2d(c, 4, 1, bias=False), nn.BatchNorm2d(4), nn.ReLU(), nn.Flatten(), nn.Linear(360, 128),
                                    nn.ReLU(), nn.Linear(128, 1), nn.Tanh())

        def forward(self, x):
            x = self.inp(x)
            for b in self.blocks:
                x = b(x)
            return self.ph(x), self.vh(x)

    torch.manual_seed(seed)
    net = Net().eval()

    def predict(board, player):   # model.py:109-124 on the oracle's planes
        with torch.no_grad():
            x = torch.from_numpy(xq_oracle.planes(board, player))[None]
            lg, v = net(x)
            return F.softmax(lg, dim=1)[0].numpy(), float(v.item())

    g = xq_oracle.OracleGame()
    rs = np.random.RandomState(seed)
    predict(g.board, 1)           # warm-up
    t0 = time.perf_counter()
    noise = rs.dirichlet([0.3] * len(g.get_legal_actions()))
    xq_oracle.mcts_search(g, sims, 1.5, predict, noise)
    return time.perf_counter() - t0


def _cpu_sample_text(procs, sims, wall):
    return (f"{procs} processes x one {sims}-simulation search from the start position, fp32 torch "
            f"XiangqiNet({CHANNELS},{BLOCKS}) 1 thread each (the reference's _cpu_worker_entry shape), {wall:.1f} s")


def cpu_selfplay_rate(procs, sims=100, pool=None):
    """The oracle PORT of the CPU path (C search of oracle/xq_oracle.c + a re-declared fp32 torch net): cross-check
    of the reference figure, and the fallback when baseline/_ref is absent."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(procs)
        pool.map(_cpu_worker, [(2, i) for i in range(procs)])           # start-up + import cost outside the timing
    try:
        t0 = time.perf_counter()
        pool.map(_cpu_worker, [(sims, 100 + i) for i in range(procs)])
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    rate = procs * sims / wall
    return {"value": rate, "unit": "sims/s", "cores": procs, "kind": "port", "sample": _cpu_sample_text(procs, sims, wall)}


def reference_selfplay_rate(steps=1, warmup=1, sims=SIMS, plies=1):
    """The reference ITSELF: baseline/_ref/training/parallel_selfplay.parallel_self_play in CPU mode, run by
    bench_reference.py in a clean process.  Returns the cpu_baseline object (kind "reference") and the raw record,
    or (None, why) when baseline/_ref is not there."""
    import subprocess
    cmd = [sys.executable, os.path.join(ROOT, "bench_reference.py"), "selfplay", "--steps", str(steps), "--warmup", str(warmup),
           "--sims", str(sims), "--plies", str(plies), "--channels", str(CHANNELS), "--blocks", str(BLOCKS)]
    env = {k: v for k, v in os.environ.items() if k not in ("PYTHONPATH",)}
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT)
    rec = None
    for ln in out.stdout.splitlines()[::-1]:
        if ln.startswith("{"):
            rec = json.loads(ln)
            break
    if out.returncode != 0 or rec is None or "unavailable" in rec or not rec.get("sims_total"):
        return None, (rec or {}).get("unavailable") or ("bench_reference.py failed: " + out.stderr[-300:].replace("\n", " | "))
    cpu = {"value": rec["sims_per_s"], "unit": "sims/s", "cores": rec["cores"], "kind": "reference",
           "workers": rec["workers"], "value_including_pool_startup": rec["sims_per_s_raw"],
           "pool_startup_s": rec["startup_median_s"],
           "sample": (f"unmodified {rec['module']} parallel_self_play() CPU mode: {rec['workers']} spawned workers "
                      f"(cpu_count-1 of {rec['cores']} cores, 1 torch thread each), one game per worker truncated to "
                      f"max_game_length={plies} ply from the start position (random_opening_moves=0), {sims} sims/move, "
                      f"Dirichlet noise on, XiangqiNet({CHANNELS},{BLOCKS}) fp32, Cython rules engine; {len(rec['steps'])} call(s), "
                      f"{rec['wall_total_s']:.1f} s wall; value = plies x sims / (wall - {rec['startup_median_s']:.1f} s pool "
                      f"start-up measured by the same call with a 0-ply budget)")}
    return cpu, rec


def run_reference(args):
    import bench
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sims = int(os.environ.get("XQ_BENCH_SIMS", SIMS))
    games = int(os.environ.get("XQ_BENCH_GAMES", GAMES_PER_GPU))
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cpu, rec = reference_selfplay_rate(steps=args.steps, warmup=max(1, args.warmup), sims=sims, plies=1)
    if cpu is None:                                   # baseline/_ref did not travel: fall back to the port, and say so
        procs = os.cpu_count() or 1
        cpu = cpu_selfplay_rate(procs, sims=REF_SIMS)
        cpu["fallback_reason"] = str(rec)[:300]
        ms = 1e3 * procs * REF_SIMS / cpu["value"]
    else:
        ms = 1e3 * rec["wall_total_s"] / max(1, len(rec["steps"]))
        try:                                          # cross-check key: the oracle port on the same cores
            port = cpu_selfplay_rate(os.cpu_count() or 1, sims=100)
            cpu["port_cross_check"] = {"value": port["value"], "sample": port["sample"]}
        except Exception as ex:
            cpu["port_cross_check"] = {"error": str(ex)[:200]}
    value = cpu["value"]
    line = {"impl": "reference", "metric": "mcts_sims_per_sec", "value": value, "unit": "sims/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic (seeded random-init weights, self-generated games)",
            "config": workload_config(world, games, sims),
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    bench.emit(line)
