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
(game g -> rank g mod world) with NO collective during self-play; the 896-byte sparse sample records are gathered to
every rank with padded all_gathers of raw bytes at the end (the reference's fan-in of worker results, :373-386) and the
returned list densifies them on access (SampleList).
"""
import logging
import os
import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from game import engine
from selfplay_engine import SelfPlayEngine

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


class SampleList:
    """The reference's `all_training_data` -- a list of (planes float32[15,10,9], policy float64[8100], z) with every
    sample followed by its column-mirrored twin (parallel_selfplay.py:124-151) -- held as the 896-byte sparse records the
    device wrote and densified on access.  A dense pair is 2 x 86 KB; at the benchmark's scale (about a million samples
    per GPU) the dense list cannot exist on any host, while len(), indexing, slicing and iteration (what train.py:322
    `replay_buffer.extend(new_data)` and SelfPlayDataset do) work on this."""

    def __init__(self, records: np.ndarray, z: np.ndarray, augment: bool = True):
        from selfplay_engine import decode_samples
        self.records = np.ascontiguousarray(records, np.uint8).reshape(-1, 896)
        self.z = np.asarray(z, np.float32)
        self.augment = bool(augment)
        self._dec = decode_samples(self.records) if len(self.records) else None

    def __len__(self):
        return len(self.records) * (2 if self.augment else 1)

    def _one(self, i):
        from selfplay_engine import MIRROR, planes_from_board
        d = self._dec
        r, twin = (i >> 1, i & 1) if self.augment else (i, 0)
        n = int(d["n"][r])
        acts = d["actions"][r, :n].astype(np.int64)
        planes = planes_from_board(d["board"][r], int(d["side"][r]))
        pol = np.zeros(8100, np.float64)
        if twin:
            pol[MIRROR[acts]] = d["probs"][r, :n]
            return np.flip(planes, axis=2).copy(), pol, float(self.z[r])
        pol[acts] = d["probs"][r, :n]
        return planes, pol, float(self.z[r])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(j) for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self._one(i)

    def __iter__(self):
        for i in range(len(self)):
            yield self._one(i)

    def __eq__(self, other):
        return list(self) == list(other)


def records_with_labels(raw, winner):
    """Finished games' records and their value labels z (parallel_selfplay.py:124-132: 0 draw, +1 when the sample's side
    to move won, -1 otherwise), in record order."""
    raw = np.ascontiguousarray(raw, np.uint8).reshape(-1, 896)
    if len(raw) == 0:
        return raw, np.zeros(0, np.float32)
    uid = raw[:, 92:96].copy().view(np.int32)[:, 0]
    side = raw[:, 90].view(np.int8).astype(np.int32)
    w = np.asarray(winner, np.int8)[uid].astype(np.int32)
    keep = w != 2
    z = np.where(w == 0, 0.0, np.where(w == side, 1.0, -1.0)).astype(np.float32)
    return raw[keep], z[keep]


def _play_local(model, config, my_games: int, local_device: int, detail: bool = False):
    """This rank's share of the games on its GPU -> (records uint8 [n,896], z float32 [n], wins, total_plies, valid_games)
    (+ per-game winner and ply arrays when `detail`)."""
    wins, total_steps, valid = {1: 0, -1: 0, 0: 0}, 0, 0
    if my_games <= 0:
        out = (np.zeros((0, 896), np.uint8), np.zeros(0, np.float32), wins, total_steps, valid)
        return out + (np.zeros(0, np.int8), np.zeros(0, np.int16)) if detail else out
    t0 = time.perf_counter()
    eng = engine(local_device)
    slots = min(my_games, int(os.environ.get("XQ_SELFPLAY_SLOTS", "4096")))
    sims = int(getattr(config, "num_simulations", 200))
    kl = max(1, int(getattr(config, "selfplay_leaves_per_game", 1)))
    key = (id(eng), slots, model.num_channels, model.num_res_blocks, kl)
    sp = _ENGINES.get(key)
    if (sp is None or sp.max_games < my_games or sp.max_simulations < sims
            or getattr(eng, "_selfplay_owner", None) is not sp):     # another SelfPlayEngine took over this context since
        if sp is not None:
            _ENGINES.clear()
            del sp
        sp = SelfPlayEngine(eng, model, n_slots=slots, max_games=my_games, max_simulations=sims, leaves_per_game=kl)
        _ENGINES.clear()
        _ENGINES[key] = sp
    else:
        sp.set_model(model)                      # fresh weights every iteration
    sp.reset()
    seed = int.from_bytes(os.urandom(8), 'big')  # the reference seeds workers from os.urandom (:167-170)
    cfg = SelfPlayEngine.make_config(config, my_games, seed=seed, add_noise=True, leaves_per_game=kl)
    t1 = time.perf_counter()
    c = sp.play_games(cfg)
    t2 = time.perf_counter()
    raw, winner, plies = sp.fetch(0, c["samples"])
    rec, z = records_with_labels(raw, winner)
    LAST_TIMING.clear()
    LAST_TIMING.update(setup_s=t1 - t0, play_s=t2 - t1, fetch_label_s=time.perf_counter() - t2, sims=c["sims"])
    for g in range(my_games):
        if winner[g] != 2:
            valid += 1
            wins[int(winner[g])] += 1
            total_steps += int(plies[g])
    if detail:
        return rec, z, wins, total_steps, valid, winner[:my_games].copy(), plies[:my_games].copy()
    return rec, z, wins, total_steps, valid


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
    rec, z, _, _, valid, winner, plies = _play_local(model, config, 1, local_device, detail=True)
    if not valid:
        raise RuntimeError("self-play game did not finish")
    return list(SampleList(rec, z, augment=False)), int(winner[0]), int(plies[0])


LAST_TIMING = {}    # phases of the most recent local self-play call: engine / weight set-up, device play loop, record download + labels
LAST_FANIN = {}     # diagnostics of the most recent multi-rank fan-in: samples and bytes each rank contributed / received


def _gather_records(dist, rec: np.ndarray, z: np.ndarray, counters: List[int], device):
    """All ranks' (records, z) concatenated in rank order + summed counters: two padded all_gathers of raw bytes and one
    all_reduce (NCCL on the GPU, gloo on the CPU) -- 900 bytes per sample on the wire, where pickling the dense tuples
    (the reference's worker fan-in, :373-386) would move 86 KB per sample."""
    import torch
    world = dist.get_world_size()
    cnt = torch.tensor([len(rec)] + list(counters), dtype=torch.int64, device=device)
    all_cnt = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(all_cnt, cnt)
    ns = [int(c[0]) for c in all_cnt]
    tot = torch.stack(all_cnt).sum(0)[1:].tolist()
    m = max(max(ns), 1)
    pad_r = torch.zeros((m, 896), dtype=torch.uint8, device=device)
    pad_z = torch.zeros(m, dtype=torch.float32, device=device)
    if len(rec):
        pad_r[:len(rec)] = torch.from_numpy(rec).to(device)
        pad_z[:len(rec)] = torch.from_numpy(z).to(device)
    all_r = [torch.empty_like(pad_r) for _ in range(world)]
    all_z = [torch.empty_like(pad_z) for _ in range(world)]
    dist.all_gather(all_r, pad_r)
    dist.all_gather(all_z, pad_z)
    rec_all = torch.cat([r[:n] for r, n in zip(all_r, ns)]).cpu().numpy()
    z_all = torch.cat([v[:n] for v, n in zip(all_z, ns)]).cpu().numpy()
    LAST_FANIN.clear()
    LAST_FANIN.update(samples_per_rank=ns, wire_bytes_per_rank=m * (896 + 4) + cnt.numel() * 8,
                      bytes_per_sample=(m * (896 + 4)) / max(m, 1))
    return rec_all, z_all, [int(x) for x in tot]


def parallel_self_play(model, config, num_workers: Optional[int] = None, use_gpu_server: bool = False,
                       gpu_device: str = 'cuda') -> Tuple[List[Tuple[np.ndarray, np.ndarray, float]], Dict[str, Any]]:
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    num_games = int(config.num_games_per_iter)
    my_games = shard_games(num_games, rank, world)
    start = time.time()
    local_device = int(os.environ.get("LOCAL_RANK", "0")) if dist else 0
    device = "cpu"
    if dist and dist.get_backend() == "nccl":
        import torch
        torch.cuda.set_device(local_device)
        device = torch.device("cuda", local_device)
    rec, z, wins, total_steps, valid = _play_local(model, config, my_games, local_device)
    if dist and world > 1:
        rec, z, (r, b, d, total_steps, valid) = _gather_records(dist, rec, z, [wins[1], wins[-1], wins[0], total_steps, valid], device)
        wins = {1: r, -1: b, 0: d}
    data = SampleList(rec, z, augment=True)
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
