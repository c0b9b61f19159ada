"""Launched by tests/test_multigpu_gpu.py under torchrun (one process per GPU, NCCL): the drop-in
parallel_self_play shards games over ranks, plays them with no collective, and every rank returns the
union of the samples."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))

import torch
import torch.distributed as dist


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import model as M
    import parallel_selfplay as ps

    class Cfg:
        num_simulations, c_puct, temperature_threshold, max_game_length = 8, 1.5, 20, 60
        random_opening_moves, enable_resign, resign_threshold, resign_check_steps = 4, True, -0.9, 5
        num_games_per_iter = 7
    torch.manual_seed(0)
    net = M.XiangqiNet(128, 1)
    # rank 0's weights are THE weights (best-model broadcast, train.py:187/528-533)
    for t in net.state_dict().values():
        tt = t.cuda()
        dist.broadcast(tt, src=0)
        t.copy_(tt.cpu())
    data, stats = ps.parallel_self_play(net, Cfg())
    assert stats["games"] == 7 and stats["num_workers"] == dist.get_world_size(), stats
    assert stats["new_samples"] == len(data) and len(data) % 2 == 0
    n = torch.tensor([len(data)], device="cuda")
    lo, hi = n.clone(), n.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert int(lo) == int(hi) == len(data)              # every rank holds the same union
    dist.barrier()
    if dist.get_rank() == 0:
        print(f"MGPU_OK ranks={dist.get_world_size()} games={stats['games']} samples={len(data)}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
