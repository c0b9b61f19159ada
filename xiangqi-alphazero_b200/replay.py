"""Device-resident replay buffer, minibatch builder and fused loss for the training step.

The reference keeps `deque(maxlen=max_buffer_size)` of dense tuples (planes float32[15,10,9], policy
float64[8100], z) on the host (train.py:203) and feeds them through `SelfPlayDataset` + a Python
`DataLoader` (train.py:114-129, 384-391): 64.8 KB per sample, every epoch.  Here the buffer is a ring of
the 896-byte sparse self-play records in HBM; a record stands for the pair (sample, column-mirrored
sample) that `_augment_data` (parallel_selfplay.py:137-151) appends, so logical element 2j is record j
and 2j+1 its mirrored twin.  Kernels: csrc/xq_train.cu (xq_replay_append, xq_train_batch,
xq_policy_value_loss).  There is no host fallback: without the CUDA library every call raises.
"""
import ctypes as C

import numpy as np
import torch

import xq_native

SAMPLE_BYTES = 896
MAX_MOVES = 128
ACTION_SPACE = 8100


class _DevArray:
    """Zero-copy torch view of a raw device pointer (torch.as_tensor understands __cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def device_view(ptr, shape, typestr, device):
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


def selfplay_buffers(sp):
    """(records uint8 [cap,896], winner int8 [max_games], plies int16 [max_games]) of a SelfPlayEngine, on its GPU."""
    e = sp.e
    a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
    e._check(e.L.xq_selfplay_device_buffers(e.h, C.byref(a), C.byref(b), C.byref(c)))
    return (device_view(a.value, (sp.sample_capacity, SAMPLE_BYTES), "|u1", e.dev),
            device_view(b.value, (sp.max_games,), "|i1", e.dev), device_view(c.value, (sp.max_games,), "<i2", e.dev))


def dense_to_records(data):
    """Host helper for `extend()`: dense reference tuples (originals only) -> (records uint8 [n,896], z float32 [n]).
    Inverse of get_state_for_nn (game.py:618-640) and of the dense policy vector."""
    n = len(data)
    rec = np.zeros((n, SAMPLE_BYTES), np.uint8)
    z = np.zeros(n, np.float32)
    for i, (state, policy, value) in enumerate(data):
        state = np.asarray(state)
        side = 1 if state[14].all() else -1
        own = np.zeros((10, 9), np.int32)
        for k in range(1, 8):
            own += k * (state[k - 1] > 0.5) - k * (state[6 + k] > 0.5)
        board = (own * side).astype(np.int8).reshape(90)
        acts = np.nonzero(np.asarray(policy) > 0)[0]
        if len(acts) > MAX_MOVES:
            raise ValueError("policy with more than 128 non-zero actions")
        rec[i, :90] = board.view(np.uint8)
        rec[i, 90] = np.int8(side).view(np.uint8)
        rec[i, 91] = len(acts)
        rec[i, 92:96] = np.frombuffer(np.int32(-1).tobytes(), np.uint8)
        a16 = np.full(MAX_MOVES, -1, np.int16)
        a16[:len(acts)] = acts
        p32 = np.zeros(MAX_MOVES, np.float32)
        p32[:len(acts)] = np.asarray(policy)[acts]
        rec[i, 128:384] = a16.view(np.uint8)
        rec[i, 384:896] = p32.view(np.uint8)
        z[i] = value
    return rec, z


class DeviceReplayBuffer:
    """Ring of sparse sample records in HBM with the deque(maxlen) behaviour of train.py:203.

    len() counts logical samples (2 per record: the sample and its mirrored twin), like the reference's
    deque after `_augment_data`.  max_buffer_size tuples = max_buffer_size // 2 records."""

    def __init__(self, eng: "xq_native.Engine", max_buffer_size: int):
        self.e = eng
        self.capacity = max(1, int(max_buffer_size) // 2)
        self.ring = torch.zeros((self.capacity, SAMPLE_BYTES), dtype=torch.uint8, device=eng.dev)
        self.z = torch.zeros(self.capacity, dtype=torch.float32, device=eng.dev)
        self.start = 0          # ring slot of the oldest record
        self.count = 0          # records held

    def __len__(self):
        return 2 * self.count

    def clear(self):
        self.start = self.count = 0

    # -- appends -----------------------------------------------------------------------------------
    def _advance(self, n):
        total = self.count + n
        new_count = min(self.capacity, total)
        self.start = (self.start + total - new_count) % self.capacity
        self.count = new_count

    def append_records(self, records: torch.Tensor, index: torch.Tensor, winner: torch.Tensor) -> int:
        """records[index[j]] (device uint8 [*,896]) -> ring, z from winner[game uid] (xq_replay_append)."""
        n = int(index.numel())
        if n == 0:
            return 0
        if n > self.capacity:                      # only the newest `capacity` records can survive
            index = index[n - self.capacity:]
            n = self.capacity
        index = index.to(torch.int64).contiguous()
        head = (self.start + self.count) % self.capacity
        e = self.e
        e._check(e.L.xq_replay_append(e.h, records.data_ptr(), index.data_ptr(), n, winner.data_ptr(), int(winner.numel()),
                                      self.ring.data_ptr(), self.z.data_ptr(), self.capacity, head, e._stream()))
        self._advance(n)
        return n

    def append_from_selfplay(self, sp, n_samples: int) -> int:
        """All samples of FINISHED games among the first n_samples records of a SelfPlayEngine, GAME-MAJOR (game uid, then
        ply): the reference extends its deque game by game (parallel_selfplay.py:373-386), so when an iteration yields
        more samples than the buffer holds, whole recent games survive -- not the last plies of every game, which is
        what the device's ply-interleaved append order would keep."""
        if n_samples <= 0:
            return 0
        rec, winner, _ = selfplay_buffers(sp)
        uid = rec[:n_samples, 92:96].contiguous().view(torch.int32).reshape(-1).long()
        ply = rec[:n_samples, 96:100].contiguous().view(torch.int32).reshape(-1).long()
        ok = (uid >= 0) & (uid < winner.numel())
        done = torch.zeros_like(ok)
        done[ok] = winner[uid[ok]] != 2
        index = torch.nonzero(done).reshape(-1)
        order = torch.argsort(uid[index] * 1024 + ply[index], stable=True)
        return self.append_records(rec, index[order], winner)

    def append_raw(self, records: torch.Tensor, z: torch.Tensor) -> int:
        """Already-labelled records (device or host tensors): uid is ignored, z is taken as given."""
        records = records.to(self.e.dev).reshape(-1, SAMPLE_BYTES).contiguous()
        z = z.to(self.e.dev, torch.float32).reshape(-1)
        n = records.shape[0]
        if n == 0:
            return 0
        if n > self.capacity:
            records, z, n = records[n - self.capacity:], z[n - self.capacity:], self.capacity
        head = (self.start + self.count) % self.capacity
        first = min(n, self.capacity - head)
        self.ring[head:head + first] = records[:first]
        self.z[head:head + first] = z[:first]
        if first < n:
            self.ring[:n - first] = records[first:]
            self.z[:n - first] = z[first:]
        self._advance(n)
        return n

    def extend(self, data):
        """deque.extend of the reference's dense tuples as produced by parallel_self_play: (sample, mirrored sample)
        pairs.  Only the originals are stored; their twins are regenerated by the batch kernel."""
        data = list(data)
        if len(data) % 2:
            raise ValueError("extend() expects (sample, mirrored sample) pairs as _augment_data produces them")
        rec, z = dense_to_records(data[0::2])
        self.append_raw(torch.from_numpy(rec), torch.from_numpy(z))

    def records_in_order(self):
        """(records, z) oldest first, as device tensors (checkpointing, all-gather across ranks)."""
        idx = (torch.arange(self.count, device=self.e.dev) + self.start) % self.capacity
        return self.ring[idx], self.z[idx]

    # -- minibatch ---------------------------------------------------------------------------------
    def batch(self, logical_index: torch.Tensor, out=None):
        """Logical indices (0 <= L < len(self)) -> planes float32 [B,15,10,9], (actions int16 [B,128], probs float32
        [B,128], n int32 [B]), z float32 [B]  (xq_train_batch).  out = (planes, act, prob, n, z): write into these tensors
        (the static inputs of a captured training step) instead of new ones."""
        idx = logical_index.to(self.e.dev, torch.int64).contiguous()
        B = int(idx.numel())
        dev = self.e.dev
        if out is not None:
            planes, act, prob, n, z = out
        else:
            planes = torch.empty((B, 15, 10, 9), dtype=torch.float32, device=dev)
            act = torch.empty((B, MAX_MOVES), dtype=torch.int16, device=dev)
            prob = torch.empty((B, MAX_MOVES), dtype=torch.float32, device=dev)
            n = torch.empty((B,), dtype=torch.int32, device=dev)
            z = torch.empty((B,), dtype=torch.float32, device=dev)
        e = self.e
        e._check(e.L.xq_train_batch(e.h, self.ring.data_ptr(), self.z.data_ptr(), self.capacity, self.start, idx.data_ptr(), B,
                                    planes.data_ptr(), act.data_ptr(), prob.data_ptr(), n.data_ptr(), z.data_ptr(), e._stream()))
        return planes, (act, prob, n), z


class _PolicyValueLoss(torch.autograd.Function):
    """policy_loss = -mean(sum(pi * log_softmax(logits))), value_loss = mse(value, z) (train.py:408-414) with sparse pi.
    Forward computes both losses and the gradients in one kernel pass; backward only scales them."""

    @staticmethod
    def forward(ctx, logits, value, act, prob, n, z, inv_batch, eng):
        B = logits.shape[0]
        logits = logits.float().contiguous()
        v = value.float().reshape(-1).contiguous()
        g_logits = torch.empty_like(logits)
        g_value = torch.empty_like(v)
        prow = torch.empty_like(v)
        vrow = torch.empty_like(v)
        eng._check(eng.L.xq_policy_value_loss(eng.h, logits.data_ptr(), logits.stride(0), v.data_ptr(), act.data_ptr(),
                                              prob.data_ptr(), n.data_ptr(), z.data_ptr(), B, C.c_float(inv_batch),
                                              g_logits.data_ptr(), g_logits.stride(0), g_value.data_ptr(), prow.data_ptr(),
                                              vrow.data_ptr(), eng._stream()))
        ctx.save_for_backward(g_logits, g_value)
        ctx.value_shape = value.shape
        return prow.sum() * inv_batch, vrow.sum() * inv_batch

    @staticmethod
    def backward(ctx, gp, gv):
        g_logits, g_value = ctx.saved_tensors
        return g_logits * gp, (g_value * gv).reshape(ctx.value_shape), None, None, None, None, None, None


def policy_value_loss(eng, logits, value, target, z, global_batch=None):
    """(policy_loss, value_loss) as in train.py:408-413.  target = (actions, probs, n) from DeviceReplayBuffer.batch();
    global_batch: the mean's denominator (the full batch size when this rank holds a shard of it)."""
    act, prob, n = target
    B = logits.shape[0]
    inv = 1.0 / float(global_batch if global_batch else B)
    return _PolicyValueLoss.apply(logits, value, act, prob, n, z, inv, eng)
