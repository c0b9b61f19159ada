"""Drop-in replacement for the reference's training/parallel_selfplay.py: same entry point

    parallel_self_play(model, config, num_workers=None, use_gpu_server=False, gpu_device='cuda')
        -> (all_training_data, stats)

(parallel_selfplay.py:264-334) so train.py:313-321 runs unchanged.  Instead of spawning CPU worker
processes that each play whole games serially (and, in "GPU server" mode, pickle one position per
request over a Unix socket), every game lives on the GPU: the device-resident loop of
selfplay_engine.SelfPlayEngine plays `config.num_games_per_iter` games in lockstep slots, batching
all leaves of a step into one tcgen05 forward.  `num_workers`, `use_gpu_server` and `gpu_device` are
accepted for signature compatibility: num_workers caps the number of concurrent game slots'
multiplier only through XQ_SELFPLAY_SLOTS; inference always runs on the engine's GPU.

Multi-GPU: launched under torchrun (one process per GPU), each rank plays its share of the games
(game g -> rank g mod world) with NO collective during self-play; samples are gathered to every
rank with one all_gather_object at the end (the reference's fan-in of worker results, :373-386).
"""
import logging
import os
import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from game import engine
from selfplay_engine import SelfPlayEngine, decode_samples, samples_to_reference_tuples

logger = logging.getLogger(__name__)

_ENGINES = {}


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def shard_games(num_games: int, rank: int, world: int) -> int:
    """Games of this rank when game g goes to rank g mod world."""
    return num_games // world + (1 if rank < num_games % world else 0)


def _play_local(model, config, my_games: int, local_device: int, augment: bool = True, detail: bool = False):
    """This rank's share of the games on its GPU -> (samples, wins, total_plies, valid_games)
    (+ per-game winner and ply arrays when `detail`)."""
    data, wins, total_steps, valid = [], {1: 0, -1: 0, 0: 0}, 0, 0
    if my_games <= 0:
        return data, wins, total_steps, valid
    eng = engine(local_device)
    slots = min(my_games, int(os.environ.get("XQ_SELFPLAY_SLOTS", "4096")))
    sims = int(getattr(config, "num_simulations", 200))
    key = (id(eng), slots, model.num_channels, model.num_res_blocks)
    sp = _ENGINES.get(key)
    if sp is None or sp.max_games < my_games or sp.max_simulations < sims:
        if sp is not None:
            _ENGINES.clear()
            del sp
        sp = SelfPlayEngine(eng, model, n_slots=slots, max_games=my_games, max_simulations=sims)
        _ENGINES.clear()
        _ENGINES[key] = sp
    else:
        sp.set_model(model)                      # fresh weights every iteration
    sp.reset()
    seed = int.from_bytes(os.urandom(8), 'big')  # the reference seeds workers from os.urandom (:167-170)
    cfg = SelfPlayEngine.make_config(config, my_games, seed=seed, add_noise=True)
    c = sp.play_games(cfg)
    raw, winner, plies = sp.fetch(0, c["samples"])
    dec = decode_samples(raw)
    data = samples_to_reference_tuples(dec, winner, augment=augment)
    for g in range(my_games):
        if winner[g] != 2:
            valid += 1
            wins[int(winner[g])] += 1
            total_steps += int(plies[g])
    if detail:
        return data, wins, total_steps, valid, winner[:my_games].copy(), plies[:my_games].copy()
    return data, wins, total_steps, valid


def _play_one_game(model_or_client, config, device='cpu') -> Tuple[List, int, int]:
    """parallel_selfplay.py:42-134 for callers that play single games: one game on the device loop (one slot),
    returning (training_data without the mirrored twins, winner, plies).  `model_or_client` is a XiangqiNet or an
    inference_server.InferenceClient (its server's model is used); `device` is accepted and ignored -- the game
    runs on the engine's GPU."""
    model = model_or_client
    if not hasattr(model, "num_res_blocks"):
        import inference_server
        model = inference_server._MODELS.get(getattr(model_or_client, "socket_path", None))
        if model is None:
            raise RuntimeError("inference server is not running")
    local_device = int(os.environ.get("LOCAL_RANK", "0")) if _dist() else 0
    data, _, _, valid, winner, plies = _play_local(model, config, 1, local_device, augment=False, detail=True)
    if not valid:
        raise RuntimeError("self-play game did not finish")
    return data, int(winner[0]), int(plies[0])


def parallel_self_play(model, config, num_workers: Optional[int] = None, use_gpu_server: bool = False,
                       gpu_device: str = 'cuda') -> Tuple[List[Tuple[np.ndarray, np.ndarray, float]], Dict[str, Any]]:
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    num_games = int(config.num_games_per_iter)
    my_games = shard_games(num_games, rank, world)
    start = time.time()
    local_device = int(os.environ.get("LOCAL_RANK", "0")) if dist else 0
    if dist and dist.get_backend() == "nccl":
        import torch
        torch.cuda.set_device(local_device)          # object collectives stage through the current CUDA device
    data, wins, total_steps, valid = _play_local(model, config, my_games, local_device)
    if dist and world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (data, wins, total_steps, valid))
        data, wins, total_steps, valid = [], {1: 0, -1: 0, 0: 0}, 0, 0
        for d, w, t, v in parts:
            data.extend(d)
            for k in w:
                wins[k] += w[k]
            total_steps += t
            valid += v
    elapsed = time.time() - start
    stats = {                                        # keys of parallel_selfplay.py:316-326
        'games': valid,
        'red_wins': wins.get(1, 0),
        'black_wins': wins.get(-1, 0),
        'draws': wins.get(0, 0),
        'avg_steps': total_steps / max(valid, 1),
        'new_samples': len(data),
        'total_time': elapsed,
        'num_workers': world,
        'mode': 'gpu',
    }
    logger.info("self-play done: %d games, %.1fs, red %d black %d draw %d, avg %.0f plies, %d samples",
                valid, elapsed, stats['red_wins'], stats['black_wins'], stats['draws'], stats['avg_steps'], len(data))
    return data, stats


def _augment_data(data):
    """Column-mirror augmentation of dense samples (parallel_selfplay.py:137-151), as one index
    permutation instead of a Python loop over 8100 actions."""
    from selfplay_engine import MIRROR
    out = []
    for state, probs, value in data:
        out.append((state, probs, value))
        fp = np.zeros_like(probs)
        nz = np.nonzero(probs > 0)[0]
        fp[MIRROR[nz]] = probs[nz]
        out.append((np.flip(state, axis=2).copy(), fp, value))
    return out
