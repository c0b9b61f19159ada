"""Drop-in replacement for the reference's training/train.py: same TrainingConfig fields, same
AlphaZeroTrainer methods and return values, same checkpoint dictionaries -- with the iteration
living on the GPU(s):

  self_play()      device-resident lockstep games (selfplay_engine.SelfPlayEngine); the sparse sample records
                   go straight from the self-play buffer into the HBM replay ring (replay.DeviceReplayBuffer),
                   nothing is densified or pickled (train.py:313-327 + parallel_selfplay.py:264-334)
  train_network()  train.py:376-447: 5 epochs of shuffled minibatches, soft-target CE + MSE, Adam(lr, L2),
                   clip_grad_norm_(1.0), MultiStepLR per iteration.  The forward/backward stays the torch module;
                   minibatch building, loss + its gradient, gradient norm, clip + Adam are the kernels of
                   csrc/xq_train.cu working on flat parameter / gradient buffers
  evaluate()       train.py:453-535 as one batched arena (arena.py), same promotion rule (>= eval_win_rate, draws = 1/2)
                   and the same revert-on-failure
  save/load_checkpoint  train.py:537-580 key for key

Data parallel (launch with torchrun, one process per GPU, NCCL): games and evaluation pairs are sharded with
no collective inside; the new sample records are all-gathered so every rank holds the same replay ring; every
global minibatch of `batch_size` is split across the ranks, BatchNorm statistics are synchronised
(DPBatchNorm2d: the batch statistics are those of the reference's single 256-sample batch), the flat gradient is
summed with ONE all-reduce per step, and the promoted weights are broadcast from rank 0.
"""
import argparse
import json
import logging
import os
import random
import sys
import time
from typing import List, Tuple

import numpy as np
import torch
import torch.nn.functional as F
import torch.optim as optim

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.insert(0, _HERE)

from game import XiangqiGame, ACTION_SPACE, decode_action, engine   # noqa: E402
from model import XiangqiNet, count_parameters                       # noqa: E402
from mcts import MCTS                                                # noqa: E402
from selfplay_engine import SelfPlayEngine                           # noqa: E402
from replay import DeviceReplayBuffer, policy_value_loss, SAMPLE_BYTES   # noqa: E402
from parallel_selfplay import shard_games                            # noqa: E402
import arena as _arena                                               # noqa: E402

logging.basicConfig(level=logging.INFO, format='%(asctime)s [%(levelname)s] %(message)s')
logger = logging.getLogger(__name__)


class TrainingConfig:
    """Field for field the reference's TrainingConfig (train.py:55-111)."""

    def __init__(self):
        self.num_channels = 128
        self.num_res_blocks = 6
        self.num_simulations = 200
        self.c_puct = 1.5
        self.temperature_threshold = 20
        self.num_games_per_iter = 20
        self.max_game_length = 300
        self.resign_threshold = -0.9
        self.resign_check_steps = 5
        self.enable_resign = True
        self.random_opening_moves = 4
        self.num_workers = None
        self.parallel = True
        self.num_iterations = 100
        self.batch_size = 256
        self.num_epochs = 5
        self.learning_rate = 0.002
        self.weight_decay = 1e-4
        self.lr_milestones = [50, 80]
        self.lr_gamma = 0.1
        self.max_buffer_size = 50000
        self.min_buffer_size = 500
        self.eval_games = 10
        self.eval_win_rate = 0.55
        self.eval_simulations = 100
        self.checkpoint_dir = '../models'
        self.save_interval = 5
        self.use_gpu_server = False
        self.gpu_device = 'cuda'
        self.device = 'cuda' if torch.cuda.is_available() else 'cpu'
        # additions of this build
        self.selfplay_slots = 4096      # concurrent games per GPU
        self.sync_batchnorm = True      # data parallel: BatchNorm statistics over the GLOBAL minibatch (reference semantics)
        # opt-in multi-leaf search (virtual loss): descents per game and lockstep step.  1 = the reference's search
        # (mcts.py:126-153, one simulation at a time).  > 1 trades the exact visit sequence for a fuller tensor-core
        # batch when few games are in flight (evaluation's eval_games, small num_games_per_iter).
        self.selfplay_leaves_per_game = 1
        self.eval_leaves_per_game = 1
        # multi-GPU training of the reference's 256-sample minibatch: "shard" = data parallel (each rank holds
        # batch_size / world samples, global BatchNorm statistics, NCCL gradient all-reduce: the north star's mode); "replicate" = every
        # rank runs the full minibatch (identical replicas, no collective inside the step; the replicas are re-synchronised by
        # the weight broadcast after evaluation).  A 256-sample step is launch-latency bound on a B200 (4.7 ms, 60 TFLOP/s),
        # so sharding it 8 ways cannot shorten it and adds 30 small BatchNorm-statistics all-reduces per step.
        # "auto" (default): "replicate" when the hand-written step can run (it takes 1.9 ms per 256-sample step on one B200, less
        # than a sharded step spends in its collectives), else "shard".
        self.dp_mode = "auto"
        self.hand_batchnorm = True      # training-mode BatchNorm on csrc/xq_bn.cu (statistics over NVLink peer memory when sharded)
        # forward + loss + backward of the step on the hand-written tf32 tcgen05 kernels (csrc/xq_tnet.cu, tnet.HandStep): used
        # when a rank runs whole minibatches (one GPU or dp_mode "replicate") with a tower width that is a multiple of 32
        self.hand_step = True


class SelfPlayDataset(torch.utils.data.Dataset):
    """train.py:114-129: dense (planes, policy[8100], z) tuples as float tensors.  The trainer below does not use it
    (minibatches are built on the device from the sparse replay ring); it stays for scripts that feed their own
    DataLoader."""

    def __init__(self, data: List[Tuple[np.ndarray, np.ndarray, float]]):
        self.data = data

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        state, policy, value = self.data[idx]
        return (torch.as_tensor(np.ascontiguousarray(state), dtype=torch.float32),
                torch.as_tensor(np.ascontiguousarray(policy), dtype=torch.float32),
                torch.tensor([float(value)], dtype=torch.float32))


def augment_data(state: np.ndarray, policy: np.ndarray, value: float) -> List[Tuple]:
    """train.py:132-151: the sample and its column mirror; the policy is moved by the action permutation
    (fr,fc,tr,tc) -> (fr,8-fc,tr,8-tc) in one scatter instead of a loop over the 8100 actions."""
    from selfplay_engine import MIRROR
    flipped_policy = np.zeros_like(policy)
    nz = np.nonzero(policy > 0)[0]
    flipped_policy[MIRROR[nz]] = policy[nz]
    return [(state, policy, value), (np.flip(state, axis=2).copy(), flipped_policy, value)]


def make_random_opening(game: XiangqiGame, num_moves: int) -> XiangqiGame:
    """train.py:154-165: up to `num_moves` uniformly random legal moves, stopping at a finished game."""
    for _ in range(num_moves):
        moves = game.get_legal_moves()
        if not moves:
            break
        game.make_move(*random.choice(moves))
        if game.is_game_over()[0]:
            break
    return game


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def gather_records(rec, z, dist=None):
    """All ranks' new (records uint8 [n_r, 896], z float32 [n_r]) concatenated in rank order, so that every rank
    appends the same records to its replay ring (the fan-in of parallel_selfplay.py:373-386 without pickling)."""
    if dist is None or dist.get_world_size() <= 1:
        return rec, z
    world = dist.get_world_size()
    n = torch.tensor([rec.shape[0]], dtype=torch.int64, device=rec.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    pad_r = torch.zeros((m, SAMPLE_BYTES), dtype=torch.uint8, device=rec.device)
    pad_z = torch.zeros(m, dtype=torch.float32, device=rec.device)
    pad_r[:rec.shape[0]] = rec
    pad_z[:rec.shape[0]] = z
    all_r = [torch.empty_like(pad_r) for _ in range(world)]
    all_z = [torch.empty_like(pad_z) for _ in range(world)]
    dist.all_gather(all_r, pad_r)
    dist.all_gather(all_z, pad_z)
    return (torch.cat([r[:c] for r, c in zip(all_r, counts)]), torch.cat([v[:c] for v, c in zip(all_z, counts)]))


def epoch_permutation(n: int, dist=None, device="cpu") -> torch.Tensor:
    """The index order DataLoader(shuffle=True, num_workers=0) produces from the same global torch RNG state
    (train.py:384-391: the iterator's _base_seed draw, then RandomSampler's seed draw and randperm of a private
    generator); rank 0's draw is shared so that all ranks agree."""
    torch.empty((), dtype=torch.int64).random_()
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    if dist is not None and dist.get_world_size() > 1:
        s = torch.tensor([seed], dtype=torch.int64, device=device)
        dist.broadcast(s, 0)
        seed = int(s.item())
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)


def shard_batch(idx: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's part of one global minibatch.  A batch with fewer samples than ranks is processed whole by every
    rank (identical gradients; the loss scale divides the sum by the world size)."""
    if world <= 1 or idx.numel() < world:
        return idx
    return torch.tensor_split(idx, world)[rank]


class _DPBatchNormFn(torch.autograd.Function):
    """BatchNorm2d over the GLOBAL minibatch of a data-parallel step: one all-reduce of the per-channel sums in forward
    and one in backward, nothing that waits for the GPU on the host.  (torch.nn.SyncBatchNorm's forward masks the
    gathered counts with a boolean index -- a device-to-host synchronisation in every one of the 15 layers, which made
    the sharded step 2.3x slower than the single-GPU step.)  Statistics accumulate in float64."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, dist):
        C = x.shape[1]
        xd = x.double()
        stats = torch.cat([xd.sum((0, 2, 3)), (xd * xd).sum((0, 2, 3)),
                           torch.full((1,), x.numel() // C, dtype=torch.float64, device=x.device)])
        dist.all_reduce(stats)
        n = stats[-1]
        mean = stats[:C] / n
        var = (stats[C:2 * C] / n - mean * mean).clamp_(min=0.0)          # biased: what normalises the batch
        invstd = torch.rsqrt(var + eps)
        with torch.no_grad():                                               # F.batch_norm's running update (unbiased variance)
            running_mean.mul_(1.0 - momentum).add_(mean.to(running_mean.dtype), alpha=momentum)
            running_var.mul_(1.0 - momentum).add_((var * (n / (n - 1.0))).to(running_var.dtype), alpha=momentum)
        xhat = (x - mean.float().view(1, C, 1, 1)) * invstd.float().view(1, C, 1, 1)
        ctx.save_for_backward(xhat, weight, invstd.float(), n)
        ctx.dist = dist
        return xhat * weight.view(1, C, 1, 1) + bias.view(1, C, 1, 1)

    @staticmethod
    def backward(ctx, dy):
        xhat, weight, invstd, n = ctx.saved_tensors
        C = dy.shape[1]
        dyd = dy.double()
        local = torch.cat([dyd.sum((0, 2, 3)), (dyd * xhat.double()).sum((0, 2, 3))])
        red = local.clone()
        ctx.dist.all_reduce(red)                                            # global sums for dx; the parameter gradients stay
        mean_dy = (red[:C] / n).float().view(1, C, 1, 1)                    # local (the gradient all-reduce sums them)
        mean_dy_xhat = (red[C:] / n).float().view(1, C, 1, 1)
        dx = (weight * invstd).view(1, C, 1, 1) * (dy - mean_dy - xhat * mean_dy_xhat)
        return dx, local[C:].to(weight.dtype), local[:C].to(weight.dtype), None, None, None, None, None


class _BnKernelFn(torch.autograd.Function):
    """The same layer on the hand-written kernels of csrc/xq_bn.cu: per direction one reduce kernel that stores its partial
    sums into every rank's exchange buffer over NVLink and one apply kernel that waits for all ranks and normalises --
    two launches, no collective call, nothing on the host."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, eng):
        x = x.contiguous()
        N, C = x.shape[0], x.shape[1]
        HW = x.numel() // (N * C)
        y = torch.empty_like(x)
        save = torch.empty((2, C), dtype=torch.float32, device=x.device)
        eng._check(eng.L.xq_bn_forward(eng.h, x.data_ptr(), y.data_ptr(), weight.data_ptr(), bias.data_ptr(),
                                       running_mean.data_ptr(), running_var.data_ptr(), save[0].data_ptr(), save[1].data_ptr(),
                                       N, C, HW, float(eps), float(momentum), eng._stream()))
        ctx.save_for_backward(x, weight, save)
        ctx.eng = eng
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, save = ctx.saved_tensors
        eng = ctx.eng
        dy = dy.contiguous()
        N, C = x.shape[0], x.shape[1]
        HW = x.numel() // (N * C)
        dx = torch.empty_like(x)
        dwb = torch.empty((2, C), dtype=torch.float32, device=x.device)
        eng._check(eng.L.xq_bn_backward(eng.h, x.data_ptr(), dy.data_ptr(), weight.data_ptr(), save[0].data_ptr(), save[1].data_ptr(),
                                        dx.data_ptr(), dwb[0].data_ptr(), dwb[1].data_ptr(), N, C, HW, eng._stream()))
        return dx, dwb[0], dwb[1], None, None, None, None, None


def setup_peer_group(eng, dist=None) -> bool:
    """Peer-map the BatchNorm exchange buffers of all ranks (CUDA IPC over NVLink; one process per GPU on one box).
    Returns False -- and the layers fall back to all-reducing the sums with NCCL -- when the handles cannot be opened."""
    import ctypes as C
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    handle = (C.c_ubyte * 64)()
    ok = 1
    try:
        eng._check(eng.L.xq_peer_create(eng.h, rank, world, handle))
    except Exception as ex:
        logger.warning("xq_peer_create failed: %s", ex)
        ok = 0
    if world <= 1:
        return bool(ok)
    dev = eng.dev if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=dev)
    allh = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    if ok:
        buf = (C.c_ubyte * (64 * world))(*[int(v) for t in allh for v in t.cpu().tolist()])
        try:
            eng._check(eng.L.xq_peer_connect(eng.h, buf))
        except Exception as ex:
            logger.warning("xq_peer_connect failed (%s): BatchNorm statistics go through NCCL instead", ex)
            ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)           # all ranks or none
    return bool(int(flag.item()))


class DPBatchNorm2d(torch.nn.BatchNorm2d):
    """nn.BatchNorm2d (same parameters, buffers and state_dict keys) whose training-mode statistics are those of the whole
    data-parallel minibatch, i.e. of the reference's single-process 256-sample batch (train.py:397-423).  `eng` set: the
    hand-written peer-memory kernels (csrc/xq_bn.cu); else, with `dist` set: torch ops + one NCCL all-reduce each way."""
    dist = None
    eng = None

    def forward(self, x):
        if not self.training:
            return super().forward(x)
        if self.eng is not None and x.is_cuda and x.dtype == torch.float32 and self.momentum is not None:
            if self.num_batches_tracked is not None:
                self.num_batches_tracked.add_(1)
            return _BnKernelFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, self.momentum, self.eng)
        d = self.dist
        if d is None or d.get_world_size() <= 1:
            return super().forward(x)
        if self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
        return _DPBatchNormFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var, self.eps, self.momentum, d)


def convert_dp_batchnorm(module, dist, eng=None):
    """BatchNorm2d -> DPBatchNorm2d in place (parameters and buffers are kept, not copied)."""
    for name, child in list(module.named_children()):
        if isinstance(child, torch.nn.BatchNorm2d) and not isinstance(child, DPBatchNorm2d):
            new = DPBatchNorm2d(child.num_features, eps=child.eps, momentum=child.momentum, device=child.weight.device)
            new.weight, new.bias = child.weight, child.bias
            new.running_mean, new.running_var, new.num_batches_tracked = child.running_mean, child.running_var, child.num_batches_tracked
            new.dist = dist
            new.eng = eng
            new.train(child.training)
            setattr(module, name, new)
        else:
            convert_dp_batchnorm(child, dist, eng)
    return module


class FlatAdam(optim.Adam):
    """torch.optim.Adam whose parameters, gradients and moments are views of four flat float32 buffers, stepped by
    xq_grad_sumsq + xq_adam_step (clip_grad_norm_ + Adam in two kernels).  state_dict()/load_state_dict() are Adam's
    own, so checkpoints stay interchangeable with the reference's (train.py:543-544, 573-575)."""

    def __init__(self, eng, model, lr, weight_decay, max_grad_norm=1.0, dist=None):
        params = [p for p in model.parameters()]
        super().__init__(params, lr=lr, weight_decay=weight_decay)
        self.e = eng
        self.dist = dist
        self.max_grad_norm = float(max_grad_norm)
        self.n = sum(p.numel() for p in params)
        dev = params[0].device
        pad = (self.n + 3) // 4 * 4
        self.flat_p = torch.zeros(pad, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(pad, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(pad, dtype=torch.float32, device=dev)
        self.flat_v = torch.zeros(pad, dtype=torch.float32, device=dev)
        self.partial = torch.zeros(eng.torch.cuda.get_device_properties(dev).multi_processor_count * 4, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.steps = 0
        off = 0
        for p in params:
            k = p.numel()
            self.flat_p[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + k].view_as(p)
            p.grad = self.flat_g[off:off + k].view_as(p)
            self.state[p] = {'step': torch.tensor(0.0), 'exp_avg': self.flat_m[off:off + k].view_as(p),
                             'exp_avg_sq': self.flat_v[off:off + k].view_as(p)}
            off += k
        self._params = params
        self._model = model
        # Data parallel: the gradient all-reduce is split at the policy FC weight (23.3 M of the 25.2 M parameters, 93 MB
        # of the 100.7 MB).  Its gradient is the FIRST one backward produces (the FC is the last layer), so its all-reduce
        # is issued from a post-accumulate hook and runs under the whole convolutional backward; only the remaining
        # 7 MB are reduced after backward (C1).
        self._early = None
        self._big = None
        if dist is not None and dist.get_world_size() > 1:
            off = 0
            for p in params:
                if self._big is None or p.numel() > self._big[2]:
                    self._big = (p, off, p.numel())
                off += p.numel()
            big_p, big_off, big_n = self._big

            def _reduce_early(_param, self=self, lo=big_off, n=big_n):
                self._early = self.dist.all_reduce(self.flat_g[lo:lo + n], async_op=True)
            big_p.register_post_accumulate_grad_hook(_reduce_early)

    def zero_grad(self, set_to_none: bool = False):
        self.flat_g.zero_()

    def _rebind(self):
        """After load_state_dict (which re-creates the state tensors): copy the moments back into the flat buffers."""
        off = 0
        for p in self._params:
            k = p.numel()
            st = self.state[p]
            if 'exp_avg' not in st:
                # a checkpoint written before the first optimizer.step (train.py:543: the reference saves whatever
                # Adam holds, which is nothing while the buffer is below min_buffer_size): start from zero moments
                self.flat_m[off:off + k].zero_()
                self.flat_v[off:off + k].zero_()
                st['step'] = torch.tensor(0.0)
            else:
                self.flat_m[off:off + k].copy_(st['exp_avg'].reshape(-1))
                self.flat_v[off:off + k].copy_(st['exp_avg_sq'].reshape(-1))
            st['exp_avg'] = self.flat_m[off:off + k].view_as(p)
            st['exp_avg_sq'] = self.flat_v[off:off + k].view_as(p)
            self.steps = int(float(st['step']))
            if p.grad is None or p.grad.data_ptr() != self.flat_g[off:off + k].data_ptr():
                p.grad = self.flat_g[off:off + k].view_as(p)
            off += k

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._rebind()

    @torch.no_grad()
    def step(self, closure=None):
        e = self.e
        if self.dist is not None and self.dist.get_world_size() > 1:
            # C1: NCCL all-reduce of the flat gradient (sum of per-rank shards).  The policy FC weight's slice is already
            # in flight since the start of backward (see __init__); the two ranges around it follow here.
            if self._early is not None:
                _, lo, n = self._big
                works = [self._early]
                if lo > 0:
                    works.append(self.dist.all_reduce(self.flat_g[:lo], async_op=True))
                if lo + n < self.flat_g.numel():
                    works.append(self.dist.all_reduce(self.flat_g[lo + n:], async_op=True))
                for w in works:
                    w.wait()
                self._early = None
            else:
                self.dist.all_reduce(self.flat_g)
        g = self.param_groups[0]
        self.steps += 1
        b1, b2 = g['betas']
        e._check(e.L.xq_grad_sumsq(e.h, self.flat_g.data_ptr(), self.n, self.partial.data_ptr(), int(self.partial.numel()),
                                   self.sumsq.data_ptr(), e._stream()))
        e._check(e.L.xq_adam_step(e.h, self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.flat_m.data_ptr(),
                                  self.flat_v.data_ptr(), self.n, float(g['lr']), float(b1), float(b2), float(g['eps']),
                                  float(g['weight_decay']), self.steps, self.sumsq.data_ptr(), self.max_grad_norm, 1.0,
                                  e._stream()))
        for p in self._params:
            self.state[p]['step'] = torch.tensor(float(self.steps))
        # the kernels wrote the parameters through raw pointers: torch's version counters did not move, so tell the
        # model that its kernel-side (folded bf16) weight copy is stale
        if hasattr(self._model, "invalidate_b200"):
            self._model.invalidate_b200()
        elif hasattr(self._model, "_weights_generation"):
            self._model._weights_generation += 1
        return None


class AlphaZeroTrainer:
    """train.py:168-640 with the same public methods; see the module docstring for what runs where."""

    def __init__(self, config: TrainingConfig):
        self.config = config
        self.dist = _dist()
        self.rank, self.world = (self.dist.get_rank(), self.dist.get_world_size()) if self.dist else (0, 1)
        self.local_device = int(os.environ.get("LOCAL_RANK", "0")) if self.dist else 0
        torch.cuda.set_device(self.local_device)
        self.eng = engine(self.local_device)
        self.device = self.eng.dev

        self.current_model = XiangqiNet(config.num_channels, config.num_res_blocks).to(self.device)
        self.best_model = XiangqiNet(config.num_channels, config.num_res_blocks).to(self.device)
        if self.world > 1:
            self._broadcast_model(self.current_model)           # C2: every rank starts from rank 0's weights
        hand_ok = getattr(config, "hand_step", True) and config.num_channels % 32 == 0
        self.dp_mode = getattr(config, "dp_mode", "auto")
        if self.dp_mode == "auto":
            self.dp_mode = "replicate" if hand_ok else "shard"
        shard = self.world > 1 and getattr(config, "sync_batchnorm", True) and self.dp_mode == "shard"
        if getattr(config, "hand_batchnorm", True):
            # BatchNorm forward / backward on the kernels of csrc/xq_bn.cu; sharded minibatch: the statistics are exchanged
            # through NVLink peer memory inside those kernels (falls back to NCCL all-reduces if IPC mapping is refused)
            peers = setup_peer_group(self.eng, self.dist if shard else None)
            convert_dp_batchnorm(self.current_model, self.dist if shard else None, self.eng if peers else None)
        elif shard:
            convert_dp_batchnorm(self.current_model, self.dist)           # global-minibatch statistics through NCCL, no host sync
        self.best_model.load_state_dict(self.current_model.state_dict())

        self.optimizer = FlatAdam(self.eng, self.current_model, lr=config.learning_rate, weight_decay=config.weight_decay,
                                  max_grad_norm=1.0, dist=self.dist if self.dp_mode == "shard" else None)
        self.scheduler = optim.lr_scheduler.MultiStepLR(self.optimizer, milestones=config.lr_milestones, gamma=config.lr_gamma)
        # f1: forward + loss + backward of the step on the hand-written kernels (tnet.HandStep, csrc/xq_tnet.cu) whenever every
        # rank runs whole minibatches (one GPU, or dp_mode "replicate") and the tower is a multiple of 32 channels wide (all three
        # presets: 64, 128, 256); otherwise (sharded minibatch, other widths) the torch modules run the step as before.
        self._hand = None
        whole = self.world == 1 or self.dp_mode == "replicate"
        if hand_ok and whole:
            from tnet import HandStep
            self._hand = HandStep(self.eng, self.current_model)
        self.replay_buffer = DeviceReplayBuffer(self.eng, config.max_buffer_size)

        self.iteration = 0
        self.total_games = 0
        self.training_stats = []
        self._sp = None
        self._arena_eng = None          # second context on the same GPU: the arena's slots do not disturb the self-play state
        if self.rank == 0:
            os.makedirs(config.checkpoint_dir, exist_ok=True)
        self.num_workers = self.world
        logger.info("device %s, %d rank(s), parameters %s", self.device, self.world, f"{count_parameters(self.current_model):,}")

    # ---- collectives ----------------------------------------------------------------------------------
    def _broadcast_model(self, model, src=0):
        """Parameters and BatchNorm buffers (running_mean / running_var / num_batches_tracked) from `src` (train.py:187, 528-533)."""
        if self.world <= 1:
            return
        for t in list(model.parameters()) + list(model.buffers()):
            self.dist.broadcast(t.data, src)

    def _gather_new_records(self, rec, z):
        return gather_records(rec, z, self.dist if self.world > 1 else None)

    # ---- single-game path of the reference (train.py:227-301, 329-374): kept for debugging ------------------
    def self_play_game(self) -> Tuple[List, int, int]:
        """One game of the reference's SERIAL loop (its own temperature ramp 1.0 -> 0.1 over 10 plies after the
        threshold and its resign test from ply 41 on, train.py:245-283) on the drop-in MCTS / XiangqiGame, i.e. every
        search, rule query and evaluation runs on the GPU kernels, one position at a time."""
        cfg = self.config
        game = XiangqiGame()
        if cfg.random_opening_moves > 0:
            game = make_random_opening(game, random.randint(0, cfg.random_opening_moves))
        mcts = MCTS(self.best_model, num_simulations=cfg.num_simulations, c_puct=cfg.c_puct, device=self.device)
        records, step, low_values, done, winner = [], 0, 0, False, None
        while step < cfg.max_game_length:
            over = step - cfg.temperature_threshold
            temperature = 1.0 if over < 0 else (1.0 - 0.9 * over / 10 if over < 10 else 0.1)
            probs = mcts.search(game, temperature=temperature, add_noise=True)
            records.append((game.get_state_for_nn(), probs, game.current_player))
            action = int(np.random.choice(len(probs), p=probs)) if temperature > 0.05 else int(np.argmax(probs))
            game.make_move(*decode_action(action))
            step += 1
            done, winner = game.is_game_over()
            if done:
                break
            if cfg.enable_resign and step > 40:
                _, value = self.best_model.predict(game.get_state_for_nn(), self.device)
                low_values = low_values + 1 if value < cfg.resign_threshold else 0
                if low_values >= cfg.resign_check_steps:
                    done, winner = True, -game.current_player
                    break
        if not done:
            done, winner = game.is_game_over()
        winner = 0 if winner is None else winner
        data = [(st, pr, 0.0 if winner == 0 else (1.0 if winner == side else -1.0)) for st, pr, side in records]
        return data, winner, step

    def _serial_self_play(self) -> dict:
        """train.py:329-374: `num_games_per_iter` single games, each sample stored with its mirror."""
        cfg = self.config
        results, total_steps, new = {1: 0, -1: 0, 0: 0}, 0, 0
        for _ in range(cfg.num_games_per_iter):
            data, winner, steps = self.self_play_game()
            pairs = []
            for st, pr, z in data:
                pairs.extend(augment_data(st, pr, z))
            self.replay_buffer.extend(pairs)
            new += len(pairs)
            results[winner] = results.get(winner, 0) + 1
            total_steps += steps
            self.total_games += 1
        return {'games': cfg.num_games_per_iter, 'red_wins': results[1], 'black_wins': results[-1], 'draws': results[0],
                'avg_steps': total_steps / max(cfg.num_games_per_iter, 1), 'buffer_size': len(self.replay_buffer),
                'new_samples': new}

    def _parallel_self_play(self) -> dict:
        """train.py:313-327.  The parallel path IS the device-resident loop of self_play()."""
        return self.self_play()

    def _serial_evaluate(self) -> dict:
        """train.py:453-535.  evaluate() plays the same games (same colours, greedy moves, same promotion rule)
        as one batched arena; tests/test_train_gpu.py checks it move for move against the serial loop."""
        return self.evaluate()

    # ---- self-play -----------------------------------------------------------------------------------
    def self_play(self) -> dict:
        cfg = self.config
        start = time.time()
        my_games = shard_games(int(cfg.num_games_per_iter), self.rank, self.world)
        local = DeviceReplayBuffer(self.eng, 2 * max(1, my_games) * 202)
        wins = torch.zeros(5, dtype=torch.int64, device=self.device)      # red, black, draw, plies, games
        if my_games > 0:
            slots = min(my_games, int(getattr(cfg, "selfplay_slots", 4096)))
            sims = int(cfg.num_simulations)
            kl = max(1, int(getattr(cfg, "selfplay_leaves_per_game", 1)))
            sp = self._sp
            if (sp is None or getattr(self.eng, "_selfplay_owner", None) is not sp or sp.n_slots != slots
                    or sp.max_games < my_games or sp.max_simulations < sims or sp.leaves_per_game != kl):
                self._sp = None
                sp = SelfPlayEngine(self.eng, self.best_model, n_slots=slots, max_games=my_games, max_simulations=sims,
                                    leaves_per_game=kl)
                self._sp = sp
            else:
                sp.set_model(self.best_model)
            sp.reset()
            spcfg = SelfPlayEngine.make_config(cfg, my_games, seed=int.from_bytes(os.urandom(8), 'big'), add_noise=True,
                                               leaves_per_game=kl)
            c = sp.play_games(spcfg)
            local.append_from_selfplay(sp, c["samples"])
            wins += torch.tensor([c["red_wins"], c["black_wins"], c["draws"], c["plies_finished"], c["finished"]],
                                 dtype=torch.int64, device=self.device)
        rec, z = local.records_in_order()
        rec, z = self._gather_new_records(rec, z)
        self.replay_buffer.append_raw(rec, z)
        if self.world > 1:
            self.dist.all_reduce(wins)
        r, b, d, plies, games = (int(x) for x in wins.tolist())
        self.total_games += games
        stats = {'games': games, 'red_wins': r, 'black_wins': b, 'draws': d, 'avg_steps': plies / max(games, 1),
                 'new_samples': 2 * int(rec.shape[0]), 'total_time': time.time() - start, 'num_workers': self.world,
                 'mode': 'gpu', 'buffer_size': len(self.replay_buffer)}
        logger.info("self-play: %d games in %.1fs, red %d black %d draw %d, avg %.0f plies, %d new samples, buffer %d",
                    games, stats['total_time'], r, b, d, stats['avg_steps'], stats['new_samples'], stats['buffer_size'])
        return stats

    # ---- training --------------------------------------------------------------------------------------
    def _epoch_permutation(self, n: int) -> torch.Tensor:
        return epoch_permutation(n, self.dist if self.world > 1 else None, self.device)

    def _shard(self, idx: torch.Tensor) -> torch.Tensor:
        return shard_batch(idx, self.rank, self.world)

    def train_network(self) -> dict:
        cfg = self.config
        n = len(self.replay_buffer)
        if n < cfg.min_buffer_size:
            logger.info("buffer too small (%d/%d), skipping training", n, cfg.min_buffer_size)
            return {}
        logger.info("training: %d epochs over %d samples", cfg.num_epochs, n)
        self.current_model.train()
        sums = torch.zeros(2, dtype=torch.float64, device=self.device)     # policy, value loss summed over batches
        num_batches = 0
        for epoch in range(cfg.num_epochs):
            perm = self._epoch_permutation(n)
            ep = torch.zeros(2, dtype=torch.float64, device=self.device)
            ep_batches = 0
            for lo in range(0, n, cfg.batch_size):
                gidx = perm[lo:lo + cfg.batch_size]
                if self.dp_mode == "replicate":
                    mine, denom = gidx, gidx.numel()                           # every rank runs the whole minibatch
                else:
                    mine = self._shard(gidx)
                    replicated = self.world > 1 and gidx.numel() < self.world
                    denom = gidx.numel() * (self.world if replicated else 1)
                hb = self._hand.buffers(int(mine.numel())) if self._hand is not None else None
                states, target, z = self.replay_buffer.batch(mine, out=(hb.states, hb.act, hb.prob, hb.n, hb.z) if hb else None)
                if self._hand is not None:
                    # every layer, the loss and every gradient on the kernels of csrc/xq_tnet.cu (gradients are assigned)
                    p_loss, v_loss = self._hand.step(states, target[0], target[1], target[2], z, 1.0 / float(denom))
                else:
                    logits, values = self.current_model(states)
                    p_loss, v_loss = policy_value_loss(self.eng, logits, values, target, z, global_batch=denom)
                    self.optimizer.zero_grad()
                    (p_loss + v_loss).backward()
                self.optimizer.step()                                       # all-reduce + clip + Adam
                ep += torch.stack([p_loss.detach(), v_loss.detach()]).double()
                ep_batches += 1
            if self.world > 1 and self.dp_mode == "shard":
                self.dist.all_reduce(ep)                                    # per-rank partial means -> full-batch means
            sums += ep
            num_batches += ep_batches
            e = (ep / max(ep_batches, 1)).tolist()
            logger.info("  epoch %d: policy_loss=%.4f, value_loss=%.4f", epoch + 1, e[0], e[1])
        self.scheduler.step()
        if self._hand is not None:
            self._hand.sync_counters()
        p, v = (sums / max(num_batches, 1)).tolist()
        stats = {'policy_loss': p, 'value_loss': v, 'total_loss': p + v, 'learning_rate': self.optimizer.param_groups[0]['lr']}
        logger.info("training done: policy_loss=%.4f, value_loss=%.4f, lr=%.6f", p, v, stats['learning_rate'])
        return stats

    # ---- evaluation ------------------------------------------------------------------------------------
    def evaluate(self) -> dict:
        cfg = self.config
        if self._arena_eng is None:
            import xq_native
            self._arena_eng = xq_native.Engine(self.local_device)
        r = _arena.evaluate_models(self._arena_eng, self.current_model, self.best_model, int(cfg.eval_games), int(cfg.eval_simulations),
                                   float(cfg.c_puct), int(cfg.max_game_length), dist=self.dist if self.world > 1 else None,
                                   leaves_per_game=max(1, int(getattr(cfg, "eval_leaves_per_game", 1))))
        stats = {'new_wins': r['new_wins'], 'old_wins': r['old_wins'], 'draws': r['draws'], 'win_rate': r['win_rate'],
                 'model_updated': r['win_rate'] >= cfg.eval_win_rate}
        logger.info("evaluation: new %d old %d draw %d, win rate %.2f%%", r['new_wins'], r['old_wins'], r['draws'], 100 * r['win_rate'])
        if stats['model_updated']:
            self.best_model.load_state_dict(self.current_model.state_dict())
            logger.info(">>> best model updated <<<")
        else:
            self.current_model.load_state_dict(self.best_model.state_dict())
            logger.info("new model rejected, reverting to the best model")
        self._broadcast_model(self.current_model)        # C2: the promoted / reverted weights, from rank 0
        self._broadcast_model(self.best_model)
        return stats

    # ---- checkpoints (train.py:537-580, same keys) ---------------------------------------------------------
    def save_checkpoint(self, iteration: int, is_best: bool = False):
        if self.rank != 0:
            return
        cfgd = {'num_channels': self.config.num_channels, 'num_res_blocks': self.config.num_res_blocks}
        checkpoint = {
            'iteration': iteration,
            'model_state_dict': self.current_model.state_dict(),
            'best_model_state_dict': self.best_model.state_dict(),
            'optimizer_state_dict': self.optimizer.state_dict(),
            'scheduler_state_dict': self.scheduler.state_dict(),
            'config': cfgd,
            'total_games': self.total_games,
        }
        path = os.path.join(self.config.checkpoint_dir, f'checkpoint_iter{iteration}.pt')
        torch.save(checkpoint, path)
        logger.info("checkpoint saved: %s", path)
        if is_best:
            best_path = os.path.join(self.config.checkpoint_dir, 'best_model.pt')
            torch.save({'model_state_dict': self.best_model.state_dict(), 'config': cfgd, 'iteration': iteration,
                        'total_games': self.total_games}, best_path)
            logger.info("best model saved: %s", best_path)

    def load_checkpoint(self, path: str):
        checkpoint = torch.load(path, map_location=self.device)
        self.current_model.load_state_dict(checkpoint['model_state_dict'])
        self.best_model.load_state_dict(checkpoint['best_model_state_dict'])
        self.optimizer.load_state_dict(checkpoint['optimizer_state_dict'])
        if 'scheduler_state_dict' in checkpoint:
            self.scheduler.load_state_dict(checkpoint['scheduler_state_dict'])
        self.iteration = checkpoint['iteration']
        self.total_games = checkpoint.get('total_games', 0)
        logger.info("checkpoint loaded: %s, iteration=%d", path, self.iteration)

    # ---- main loop (train.py:582-640) ---------------------------------------------------------------------
    def train(self):
        cfg = self.config
        for iteration in range(self.iteration + 1, cfg.num_iterations + 1):
            self.iteration = iteration
            t0 = time.time()
            logger.info("iteration %d/%d", iteration, cfg.num_iterations)
            sp_stats = self.self_play()
            train_stats = self.train_network()
            eval_stats = {}
            if iteration % 2 == 0 and len(self.replay_buffer) >= cfg.min_buffer_size:
                eval_stats = self.evaluate()
            if iteration % cfg.save_interval == 0:
                self.save_checkpoint(iteration, is_best=True)
            stats = {'iteration': iteration, 'time': time.time() - t0, 'self_play': sp_stats, 'training': train_stats,
                     'evaluation': eval_stats}
            self.training_stats.append(stats)
            if self.rank == 0:
                with open(os.path.join(cfg.checkpoint_dir, 'training_stats.json'), 'w') as f:
                    json.dump(self.training_stats, f, indent=2, default=str)
        self.save_checkpoint(self.iteration, is_best=True)
        logger.info("training finished")


# ---- presets (train.py:647-699, same values) -------------------------------------------------------------------
def quick_train():
    c = TrainingConfig()
    c.num_channels, c.num_res_blocks, c.num_simulations, c.num_games_per_iter = 64, 3, 80, 6
    c.num_iterations, c.batch_size, c.num_epochs, c.min_buffer_size = 10, 64, 5, 100
    c.eval_games, c.eval_simulations, c.save_interval, c.temperature_threshold = 4, 40, 2, 15
    c.max_game_length, c.learning_rate, c.random_opening_moves = 200, 0.002, 4
    c.enable_resign, c.resign_threshold, c.resign_check_steps, c.parallel = True, -0.85, 3, True
    return c


def standard_train():
    c = TrainingConfig()
    c.num_channels, c.num_res_blocks, c.num_simulations, c.num_games_per_iter = 128, 6, 200, 20
    c.num_iterations, c.max_game_length, c.random_opening_moves, c.enable_resign, c.parallel = 50, 300, 6, True, True
    return c


def full_train():
    c = TrainingConfig()
    c.num_channels, c.num_res_blocks, c.num_simulations, c.num_games_per_iter = 256, 10, 400, 50
    c.num_iterations, c.max_game_length, c.random_opening_moves, c.enable_resign, c.parallel = 200, 400, 8, True, True
    return c


def main():
    ap = argparse.ArgumentParser(description='xiangqi AlphaZero training on B200')
    ap.add_argument('--mode', default='quick', choices=['quick', 'standard', 'full'])
    ap.add_argument('--iterations', type=int, default=None)
    ap.add_argument('--games-per-iter', type=int, default=None)
    ap.add_argument('--simulations', type=int, default=None)
    ap.add_argument('--channels', type=int, default=None)
    ap.add_argument('--res-blocks', type=int, default=None)
    ap.add_argument('--resume', type=str, default=None)
    ap.add_argument('--device', type=str, default=None)
    ap.add_argument('--workers', type=int, default=None)
    ap.add_argument('--no-parallel', action='store_true')
    ap.add_argument('--gpu-server', action='store_true')
    ap.add_argument('--gpu-device', type=str, default=None)
    a = ap.parse_args()
    cfg = {'quick': quick_train, 'standard': standard_train, 'full': full_train}[a.mode]()
    for key, val in (('num_iterations', a.iterations), ('num_games_per_iter', a.games_per_iter), ('num_simulations', a.simulations),
                     ('num_channels', a.channels), ('num_res_blocks', a.res_blocks), ('device', a.device), ('num_workers', a.workers),
                     ('gpu_device', a.gpu_device)):
        if val:
            setattr(cfg, key, val)
    if a.no_parallel:
        cfg.parallel = False
    if a.gpu_server:
        cfg.use_gpu_server = True
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    trainer = AlphaZeroTrainer(cfg)
    if a.resume:
        trainer.load_checkpoint(a.resume)
    trainer.train()


if __name__ == '__main__':
    main()
