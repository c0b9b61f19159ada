"""Per-step loss trajectories of the training step on the same weights, samples and minibatches:
hand-written tf32 step (tnet.HandStep) vs torch fp32 (cuDNN / cuBLAS, TF32 off) vs torch with TF32 on.
Run on the GPU box:  python profiles/tools/train_trajectory.py [channels blocks steps]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
sys.path.insert(0, ROOT)
import train as T                                   # noqa: E402
from replay import policy_value_loss                # noqa: E402
import bench_train                                  # noqa: E402

ch, blocks, steps = (int(a) for a in (sys.argv[1:4] + ["128", "2", "40"][len(sys.argv) - 1:]))
BATCHES = [64, 64, 64, 64, 44]


def run(mode):
    torch.backends.cudnn.allow_tf32 = mode == "torch_tf32"
    torch.backends.cuda.matmul.allow_tf32 = mode == "torch_tf32"
    cfg = T.TrainingConfig()
    cfg.num_channels, cfg.num_res_blocks, cfg.batch_size = ch, blocks, 64
    cfg.checkpoint_dir = "/tmp/xq_traj"
    cfg.hand_step = mode == "hand"
    torch.manual_seed(7)
    tr = T.AlphaZeroTrainer(cfg)
    rec, z = bench_train.synthetic_records(tr.eng, 400, 7)
    tr.replay_buffer.append_raw(rec[:150], z[:150])
    n = len(tr.replay_buffer)
    tr.current_model.train()
    gen = torch.Generator().manual_seed(3)
    out = []
    for s in range(steps):
        B = BATCHES[s % len(BATCHES)]
        idx = torch.randint(0, n, (B,), generator=gen)
        hb = tr._hand.buffers(B) if tr._hand is not None else None
        states, target, zz = tr.replay_buffer.batch(idx, out=(hb.states, hb.act, hb.prob, hb.n, hb.z) if hb else None)
        if tr._hand is not None:
            pl, vl = tr._hand.step(states, target[0], target[1], target[2], zz, 1.0 / B)
        else:
            logits, values = tr.current_model(states)
            pl, vl = policy_value_loss(tr.eng, logits, values, target, zz, global_batch=B)
            tr.optimizer.zero_grad()
            (pl + vl).backward()
        out.append((float(pl), float(vl)))
        tr.optimizer.step()
    return np.array(out)


res = {m: run(m) for m in ("torch_fp32", "torch_tf32", "hand")}
print("step  policy: fp32 / tf32 / hand          value: fp32 / tf32 / hand")
for s in range(steps):
    a, b, c = res["torch_fp32"][s], res["torch_tf32"][s], res["hand"][s]
    print(f"{s:3d}   {a[0]:.4f} {b[0]:.4f} {c[0]:.4f}     {a[1]:.4f} {b[1]:.4f} {c[1]:.4f}")
for m in ("torch_tf32", "hand"):
    d = np.abs(res[m] / res["torch_fp32"] - 1)
    print(f"{m}: largest relative deviation from torch fp32: policy {d[:, 0].max():.2e}, value {d[:, 1].max():.2e}; "
          f"mean over the run: policy {d[:, 0].mean():.2e}, value {d[:, 1].mean():.2e}")
