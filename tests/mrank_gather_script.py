"""2-rank gloo run of the data-parallel host plumbing (no GPU): record fan-in and the shared epoch permutation."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
import train as T   # noqa: E402

dist.init_process_group("gloo")
rank = dist.get_rank()
n = 5 if rank == 0 else 9                                   # ragged contributions
rec = torch.full((n, 896), rank, dtype=torch.uint8)
z = torch.full((n,), float(rank))
r, zz = T.gather_records(rec, z, dist)
torch.manual_seed(100 + rank)                               # different local RNG states: rank 0's draw must win
perm = T.epoch_permutation(50, dist, "cpu")
torch.save({"rec": r, "z": zz, "perm": perm}, os.path.join(sys.argv[1], f"rank{rank}.pt"))
dist.destroy_process_group()
