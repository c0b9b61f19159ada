"""N>1 path on CPU (gloo, world_size 2): games shard across ranks with no data-path collective,
the only exchange is the final fan-in of samples and counters (parallel_selfplay.py drop-in)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_games_partitions_exactly():
    sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
    import parallel_selfplay as ps
    for n in (0, 1, 7, 8, 9, 100, 32768):
        for w in (1, 2, 3, 8):
            parts = [ps.shard_games(n, r, w) for r in range(w)]
            assert sum(parts) == n and max(parts) - min(parts) <= 1
            # rank r owns games r, r+w, r+2w, ...
            assert parts == [len(range(r, n, w)) for r in range(w)]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import parallel_selfplay as ps

    def fake_local(model, config, my_games, local_device):
        # deterministic stand-in for the GPU loop: game ids of this rank, 3 sparse records per game (each stands for
        # a sample and its mirrored twin); the record's first board byte carries the game id
        recs, zs, wins = [], [], {1: 0, -1: 0, 0: 0}
        for j in range(my_games):
            gid = rank + j * world
            w = (1, -1, 0)[gid % 3]
            wins[w] += 1
            for k in range(3):
                r = np.zeros(896, np.uint8)
                r[0] = gid
                r[90] = 1                                  # side to move: red
                r[91] = 1                                  # one legal move ...
                r[128:130] = np.array([gid], np.int16).view(np.uint8)   # ... action id = game id
                r[384:388] = np.array([1.0], np.float32).view(np.uint8)
                recs.append(r)
                zs.append(float(w))
        return np.array(recs, np.uint8).reshape(-1, 896), np.array(zs, np.float32), wins, 10 * my_games, my_games

    ps._play_local = fake_local

    class Cfg:
        num_games_per_iter = 7
    data, stats = ps.parallel_self_play(object(), Cfg())
    if rank == 0:
        ids = sorted({int(np.argmax(s[1])) for s in data[0::2]})        # plain samples: policy mass on action id = game id
        first = data[0]
        ok_shapes = first[0].shape == (15, 10, 9) and first[1].shape == (8100,) and isinstance(first[2], float)
        out.put((ids, len(data), stats, dict(ps.LAST_FANIN), ok_shapes))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_fan_in():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ids, n, stats, fanin, ok_shapes = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ids == list(range(7)) and n == 7 * 6                 # every game exactly once, on some rank
    assert stats["games"] == 7 and stats["num_workers"] == 2 and stats["new_samples"] == 42
    assert stats["red_wins"] + stats["black_wins"] + stats["draws"] == 7 and stats["avg_steps"] == 10
    assert stats["mode"] == "gpu" and ok_shapes
    # the fan-in moves the sparse records, not pickled dense tuples: 900 bytes per sample on the wire
    assert sorted(fanin["samples_per_rank"]) == [9, 12] and fanin["bytes_per_sample"] == 900
    assert fanin["wire_bytes_per_rank"] < 12 * 1000
