"""Device-resident self-play: host-side driver of xq_selfplay_* (libxq_b200.so).

One `SelfPlayEngine` per GPU/process.  It owns the kernel-side network copy (model.B200Net), the
game slots and the sample buffer; `play()` advances every slot by whole plies (root evaluation +
num_simulations lockstep simulations + move selection) without any host round trip, and
`collect()` turns the sparse device records into the reference's sample tuples
(parallel_selfplay.py:124-151: (planes float32[15,10,9], policy[8100], z) + the mirrored copy).
"""
import ctypes as C
import time

import numpy as np
import torch

import xq_native
from model import B200Net, LOGIT_STRIDE, ROW0, _GemmDesc

SAMPLE_BYTES = 896


class _SpConfig(C.Structure):
    _fields_ = [("num_simulations", C.c_int32), ("c_puct", C.c_float), ("temperature_threshold", C.c_int32),
                ("max_game_length", C.c_int32), ("random_opening_moves", C.c_int32), ("enable_resign", C.c_int32),
                ("resign_threshold", C.c_float), ("resign_check_steps", C.c_int32), ("add_noise", C.c_int32),
                ("dirichlet_alpha", C.c_float), ("seed", C.c_uint64), ("target_games", C.c_int32),
                ("leaves_per_game", C.c_int32)]


class _NetPlan(C.Structure):
    _fields_ = [("layers", C.POINTER(_GemmDesc)), ("n_layers", C.c_int32), ("batch", C.c_int32),
                ("vfeats", C.c_void_p), ("w1t", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p),
                ("b2", C.c_float), ("value", C.c_void_p), ("x_planes", C.c_void_p), ("x_rows", C.c_int64),
                ("x_row0", C.c_int64), ("logits", C.c_void_p), ("logit_stride", C.c_int64), ("logits_kind", C.c_int32)]


# mirror permutation of the 8100 actions: (fr,fc,tr,tc) -> (fr,8-fc,tr,8-tc)  (parallel_selfplay.py:146-148)
def _mirror_table():
    a = np.arange(8100)
    f, t = a // 90, a % 90
    fm = (f // 9) * 9 + (8 - f % 9)
    tm = (t // 9) * 9 + (8 - t % 9)
    return (fm * 90 + tm).astype(np.int64)


MIRROR = _mirror_table()


def planes_from_board(board, side):
    """get_state_for_nn (game.py:618-640) for a recorded sample."""
    own = np.asarray(board, np.int8).reshape(10, 9).astype(np.int32) * int(side)
    f = np.zeros((15, 10, 9), np.float32)
    for k in range(1, 8):
        f[k - 1] = own == k
        f[6 + k] = own == -k
    if side == 1:
        f[14] = 1.0
    return f


def decode_samples(raw: np.ndarray):
    """raw uint8 [N,896] -> dict of arrays (layout: include/xq_b200.h XQ_SAMPLE_BYTES)."""
    raw = np.ascontiguousarray(raw).reshape(-1, SAMPLE_BYTES)
    return dict(
        board=raw[:, :90].view(np.int8),
        side=raw[:, 90].view(np.int8),
        n=raw[:, 91].copy(),
        uid=raw[:, 92:96].copy().view(np.int32)[:, 0],
        ply=raw[:, 96:100].copy().view(np.int32)[:, 0],
        played=raw[:, 100:102].copy().view(np.int16)[:, 0],
        actions=raw[:, 128:384].copy().view(np.int16),
        probs=raw[:, 384:896].copy().view(np.float32),
    )


class SelfPlayEngine:
    def __init__(self, eng: "xq_native.Engine", model, n_slots: int, max_games: int, sample_capacity=None,
                 node_capacity: int = 0, max_simulations: int = 800, leaves_per_game: int = 1):
        """leaves_per_game = K: 1 runs the reference's search (one simulation per game and step, visit counts bit-exact);
        K > 1 is the opt-in virtual-loss mode (K descents per game and step share one forward: the batch is n_slots * K
        leaves, which is what fills the tensor cores when only a few games are in flight)."""
        self.e = eng
        self.leaves_per_game = max(1, int(leaves_per_game))
        self.n_slots = int(n_slots)
        self.max_games = int(max_games)
        self.max_simulations = int(max_simulations)
        if sample_capacity is None:
            sample_capacity = self.max_games * 201 + self.n_slots
        self.sample_capacity = int(sample_capacity)
        if not node_capacity:
            # one expansion per simulation (+ the root), at most 128 children each, 35 on average: 64 per
            # expansion is a safe pool size and the kernels flag an overflow instead of writing past it
            node_capacity = self.n_slots * (self.max_simulations + 2) * 64 + self.n_slots
        self.node_capacity = int(node_capacity)
        eng._check(eng.L.xq_selfplay_create(eng.h, self.n_slots, self.max_games, self.sample_capacity, self.node_capacity))
        eng._selfplay_owner = self      # a context holds ONE self-play state: a later SelfPlayEngine on it supersedes this one
        self.net = None
        self.set_model(model)
        self.fetched = 0

    def set_model(self, model):
        """(Re)build the kernel-side weights: the 'hot update' of inference_server.py:479-496."""
        self.net = B200Net(self.e, model, max_batch=self.n_slots * self.leaves_per_game)
        n = self.net
        self.plan = _NetPlan(layers=n.desc_array, n_layers=n.n_layers, batch=n.max_batch, vfeats=n.vfeat.data_ptr(),
                             w1t=n.w1t.data_ptr(), b1=n.b1.data_ptr(), w2=n.w2.data_ptr(), b2=n.b2,
                             value=n.value.data_ptr(), x_planes=n.x0.data_ptr(), x_rows=n.rows, x_row0=ROW0,
                             logits=n.logits.data_ptr(), logit_stride=LOGIT_STRIDE, logits_kind=1)

    @staticmethod
    def make_config(config, target_games, seed=0, add_noise=True, leaves_per_game=1):
        """TrainingConfig / worker dict (parallel_selfplay.py:184-187) -> xq_selfplay_config."""
        g = (lambda k, d=None: config.get(k, d)) if isinstance(config, dict) else (lambda k, d=None: getattr(config, k, d))
        return _SpConfig(num_simulations=int(g("num_simulations", 200)), c_puct=float(g("c_puct", 1.5)),
                         temperature_threshold=int(g("temperature_threshold", 20)),
                         max_game_length=int(g("max_game_length", 300)),
                         random_opening_moves=int(g("random_opening_moves", 4)),
                         enable_resign=int(bool(g("enable_resign", True))),
                         resign_threshold=float(g("resign_threshold", -0.9)),
                         resign_check_steps=int(g("resign_check_steps", 5)), add_noise=int(bool(add_noise)),
                         dirichlet_alpha=0.3, seed=int(seed) & 0xFFFFFFFFFFFFFFFF, target_games=int(target_games),
                         leaves_per_game=int(leaves_per_game))

    def _own(self):
        if getattr(self.e, "_selfplay_owner", None) is not self:
            raise xq_native.XqError("this SelfPlayEngine was superseded by a newer one on the same engine context "
                                    "(one context holds one self-play state); create it on its own xq_native.Engine")

    def reset(self):
        self._own()
        self.e._check(self.e.L.xq_selfplay_reset(self.e.h, self.e._stream()))
        self.fetched = 0

    def play(self, cfg: _SpConfig, n_plies: int):
        self._own()
        if cfg.leaves_per_game > self.leaves_per_game:
            raise xq_native.XqError(f"leaves_per_game {cfg.leaves_per_game} exceeds the {self.leaves_per_game} this SelfPlayEngine "
                                    "(its network batch) was built for")
        if cfg.num_simulations > self.max_simulations:
            raise xq_native.XqError(f"num_simulations {cfg.num_simulations} exceeds the node pool sized for "
                                    f"{self.max_simulations}: build the SelfPlayEngine with max_simulations >= it")
        self.e._check(self.e.L.xq_selfplay_play(self.e.h, C.byref(cfg), C.byref(self.plan), int(n_plies), self.e._stream()))

    def counters(self):
        self._own()
        buf = (C.c_longlong * 14)()
        self.e._check(self.e.L.xq_selfplay_counters(self.e.h, buf))
        k = ["started", "finished", "samples", "red_wins", "black_wins", "draws", "plies_finished", "dropped",
             "sims", "terminal_sims", "max_depth", "evals", "nodes", "error"]
        return dict(zip(k, [int(x) for x in buf]))

    def fetch(self, first, count, out=None):
        """Sample records [first, first+count) and all per-game results -> host numpy."""
        raw = out if out is not None else np.empty((count, SAMPLE_BYTES), np.uint8)
        win = np.empty(self.max_games, np.int8)
        plies = np.empty(self.max_games, np.int16)
        self.e._check(self.e.L.xq_selfplay_fetch(self.e.h, int(first), int(count), C.c_void_p(raw.ctypes.data),
                                                 C.c_void_p(win.ctypes.data), C.c_void_p(plies.ctypes.data), self.max_games))
        return raw, win, plies

    def slots(self):
        G = self.n_slots
        boards = np.empty((G, 90), np.int8)
        meta = np.empty((G, 4), np.int32)
        status = np.empty(G, np.int32)
        uid = np.empty(G, np.int32)
        self.e._check(self.e.L.xq_selfplay_slots(self.e.h, C.c_void_p(boards.ctypes.data), C.c_void_p(meta.ctypes.data),
                                                 C.c_void_p(status.ctypes.data), C.c_void_p(uid.ctypes.data)))
        return boards, meta, status, uid

    def play_games(self, cfg: _SpConfig, max_plies: int = 100000, chunk: int = 8):
        """Play until cfg.target_games games have finished (or max_plies plies)."""
        played = 0
        last_finished, last_progress = -1, 0
        tail = False
        self.e._check(self.e.L.xq_selfplay_set_live_bound(self.e.h, 0))     # a bound left by an earlier call does not hold for this one
        while played < max_plies:
            # once every game has been started the loop is in its tail: ask after every ply, so that no ply of
            # num_simulations lockstep steps is enqueued for slots that have all finished
            step = min(1 if tail else chunk, max_plies - played)
            self.play(cfg, step)
            played += step
            c = self.counters()
            if c["error"]:
                raise xq_native.XqError(f"self-play device error bits {c['error']}")
            if c["finished"] >= min(cfg.target_games, self.max_games):
                break
            tail = c["started"] >= min(cfg.target_games, self.max_games)
            if tail:
                # no game will be started any more: the games still alive are at most the ones that have not finished; the
                # next ply's forwards are launched for that many boards (small-batch layer variants late in an iteration)
                self.e._check(self.e.L.xq_selfplay_set_live_bound(self.e.h, max(1, min(cfg.target_games, self.max_games) - c["finished"])))
            if c["finished"] != last_finished:
                last_finished, last_progress = c["finished"], played
            elif played - last_progress > 2 * 201 + 64:      # no game can last this long (game.py:595: 200 plies)
                raise xq_native.XqError(f"self-play made no progress for {played - last_progress} plies: {c}")
        return self.counters()


def samples_to_reference_tuples(dec, winner, augment=True, finished_only=True, dense_dtype=np.float64):
    """Sparse records -> the reference's list of (planes, policy[8100], z) (+ mirrored copies,
    parallel_selfplay.py:137-151).  z: +1 if the game's winner is the sample's side to move, -1 if
    the other side, 0 for a draw (parallel_selfplay.py:124-132)."""
    out = []
    for i in range(len(dec["side"])):
        w = int(winner[dec["uid"][i]])
        if w == 2:
            if finished_only:
                continue
            w = 0
        side = int(dec["side"][i])
        z = 0.0 if w == 0 else (1.0 if w == side else -1.0)
        n = int(dec["n"][i])
        acts = dec["actions"][i, :n].astype(np.int64)
        pol = np.zeros(8100, dense_dtype)
        pol[acts] = dec["probs"][i, :n]
        planes = planes_from_board(dec["board"][i], side)
        out.append((planes, pol, z))
        if augment:
            fp = np.zeros(8100, dense_dtype)
            fp[MIRROR[acts]] = dec["probs"][i, :n]
            out.append((np.flip(planes, axis=2).copy(), fp, z))
    return out
