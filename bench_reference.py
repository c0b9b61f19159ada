#!/usr/bin/env python
"""bench_reference.py -- times the UNMODIFIED reference on the host cores (CPU arm of bench.py).

Runs in its own process with only baseline/_ref/training on sys.path (the mirror of the reference's training/
directory that oracle/Makefile target `baseline` creates; the product package has modules of the same names and must
not be importable here).  Nothing of this repo's engine is on the measured path.

  selfplay   parallel_selfplay.parallel_self_play(model, config) in CPU mode (parallel_selfplay.py:264-388,
             _run_cpu_mode -> _cpu_worker_entry -> _play_one_game -> mcts.MCTS.search -> model.predict on the
             Cython rules engine): XiangqiNet(128, 6) seeded random init, num_workers = cpu_count - 1 (the
             reference's default, :287-288), one game per worker, 800 simulations per move, Dirichlet root
             noise on (the reference always passes add_noise=True), games truncated to a fixed ply budget with
             config.max_game_length (:77-87).  sims/s = plies searched x num_simulations / wall.
(The movegen reference arm -- cy_generate_legal_moves + cy_is_in_check of the same engine, compiled as oracle/_ref --
lives in bench.py:cpu_movegen_rate.)

Prints one JSON object on stdout.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(ROOT, "baseline", "_ref", "training")


def available():
    import glob
    return os.path.exists(os.path.join(REF, "parallel_selfplay.py")) and bool(
        glob.glob(os.path.join(REF, "cython_engine", "game_core*.so")))


def _clean_path():
    pkg = os.path.join(ROOT, "xiangqi-alphazero_b200")
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") not in (pkg, os.path.join(ROOT, "oracle"))]
    sys.path.insert(0, REF)


class RefConfig:
    """The reference's standard_train preset (train.py:677-689 over the TrainingConfig defaults :64-78) for the
    keys parallel_self_play reads (:184-187, :285), at BASELINE configs[2]'s 800 simulations per move."""

    def __init__(self, sims, games, ply_budget, opening):
        self.num_simulations = sims
        self.c_puct = 1.5
        self.temperature_threshold = 20
        self.max_game_length = ply_budget
        self.random_opening_moves = opening
        self.enable_resign = True
        self.resign_threshold = -0.9
        self.resign_check_steps = 5
        self.num_games_per_iter = games


def selfplay(args):
    _clean_path()
    import torch
    import game
    import parallel_selfplay as ps
    from model import XiangqiNet
    assert os.path.abspath(ps.__file__).startswith(REF), ps.__file__
    if not game._USE_CYTHON:
        raise RuntimeError("the reference's Cython engine did not load")
    cores = os.cpu_count() or 1
    workers = args.workers or max(1, cores - 1)
    torch.manual_seed(20261018)
    model = XiangqiNet(num_channels=args.channels, num_res_blocks=args.blocks).eval()
    out = {"impl": "reference", "cores": cores, "workers": workers, "sims": args.sims, "plies": args.plies,
           "module": os.path.relpath(ps.__file__, ROOT), "cython": True, "startup_s": [], "steps": []}
    # start-up cost of one parallel_self_play call (spawned pool, torch import and model rebuild in every worker):
    # the same call with a zero-ply budget.  A real iteration amortises it over whole games; the bounded sample
    # cannot, so it is measured and taken off (in the reference's favour).
    for _ in range(max(1, args.warmup)):
        t0 = time.perf_counter()
        ps.parallel_self_play(model, RefConfig(args.sims, workers, 0, args.opening), num_workers=workers)
        out["startup_s"].append(time.perf_counter() - t0)
    startup = min(out["startup_s"])                  # the first call also pays the cold imports: take the fastest
    for _ in range(args.steps):
        t0 = time.perf_counter()
        data, stats = ps.parallel_self_play(model, RefConfig(args.sims, workers, args.plies, args.opening), num_workers=workers)
        wall = time.perf_counter() - t0
        plies = len(data) // 2                       # every searched ply yields one sample + its mirrored twin
        out["steps"].append({"wall_s": wall, "plies": plies, "sims": plies * args.sims, "games": stats["games"],
                             "stats_total_time": stats["total_time"]})
    sims = sum(s["sims"] for s in out["steps"])
    wall = sum(s["wall_s"] for s in out["steps"])
    net = sum(s["wall_s"] - startup for s in out["steps"])
    if net < 0.25 * wall:                            # search time within the noise of the start-up: do not subtract
        net, out["startup_not_subtracted"] = wall, True
    out.update(startup_median_s=startup, sims_total=sims, wall_total_s=wall,
               sims_per_s_raw=sims / wall if wall else 0.0, sims_per_s=sims / net if net else 0.0)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["selfplay"])
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--plies", type=int, default=1)
    ap.add_argument("--opening", type=int, default=0)
    ap.add_argument("--workers", type=int, default=0)
    ap.add_argument("--channels", type=int, default=128)
    ap.add_argument("--blocks", type=int, default=6)
    a = ap.parse_args()
    if not available():
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/training is missing (make -C oracle baseline)"}))
        sys.exit(0)
    selfplay(a)
