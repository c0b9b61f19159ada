"""torch.profiler over the data-parallel training step (run under torchrun): where do the milliseconds of a sharded
256-sample step go?  Prints rank 0's tables (CPU self time, CUDA time) and the step time with / without the profiler."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
import torch
import torch.distributed as dist


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bench_train
    import train as T
    from replay import policy_value_loss
    cfg = T.TrainingConfig()
    cfg.checkpoint_dir = "/tmp/xq_prof_train"
    cfg.dp_mode = os.environ.get("XQ_BENCH_DP_MODE", "shard")
    torch.manual_seed(1)
    tr = T.AlphaZeroTrainer(cfg)
    eng = tr.eng
    rec, z = bench_train.synthetic_records(eng, 20000, 1)
    tr.replay_buffer.append_raw(rec, z)
    n = len(tr.replay_buffer)
    tr.current_model.train()
    gen = torch.Generator().manual_seed(1)
    B = 256

    def step():
        gidx = torch.randint(0, n, (B,), generator=gen)
        mine = gidx if cfg.dp_mode == "replicate" else tr._shard(gidx)
        states, target, zz = tr.replay_buffer.batch(mine)
        logits, values = tr.current_model(states)
        pl, vl = policy_value_loss(eng, logits, values, target, zz, global_batch=B)
        tr.optimizer.zero_grad()
        (pl + vl).backward()
        tr.optimizer.step()

    def timed(k):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(k):
            step()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / k * 1e3
    for _ in range(10):
        step()
    ms = timed(50)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(10):
            step()
        torch.cuda.synchronize()
    if local == 0:
        print(f"== world {world} mode {cfg.dp_mode}: {ms:.2f} ms per step (50 steps, wall clock)")
        print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=22, max_name_column_width=60))
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
