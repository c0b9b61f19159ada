"""Deterministic evaluators for MCTS parity tests.

NumPy mirrors of the C evaluators in oracle/xq_oracle.c (xqo_eval_uniform / _hash / _ratio),
exposing the reference's `predict(state[, device]) -> (float32[8100], float)` contract
(model.py:109-124) so that the reference's own mcts.py can be driven with them when the
golden fixtures are generated (tests/golden/make_golden.py).  They recover (board, side)
from the 15 feature planes, so they work with any engine that emits reference planes.
"""
import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def planes_to_board(state):
    """Invert game.py:618-640."""
    state = np.asarray(state)
    red_to_move = bool(state[14, 0, 0] > 0.5)
    player = 1 if red_to_move else -1
    board = np.zeros((10, 9), np.int8)
    for k in range(7):
        board[state[k] > 0.5] = (k + 1) * player
        board[state[7 + k] > 0.5] = -(k + 1) * player
    return board, player


def board_hash(board, player):
    h = 2166136261
    for v in np.asarray(board, np.int8).reshape(90).view(np.uint8):
        h = ((h ^ int(v)) * 16777619) & 0xFFFFFFFF
    h = ((h ^ (player & 0xFF)) * 16777619) & 0xFFFFFFFF
    return h


def _mix32(x):
    x = x.astype(np.uint64) & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & _M32
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & _M32
    x ^= x >> np.uint64(16)
    return x


_A1 = np.arange(1, 8101, dtype=np.uint64)
_VAL = np.array([0, 0, 20, 20, 40, 90, 45, 10], np.int64)


def _material_diff(board, player):
    b = np.asarray(board, np.int64).reshape(90)
    red = _VAL[b[b > 0]].sum()
    black = _VAL[-b[b < 0]].sum()
    d = int(red - black) * (1 if player == 1 else -1)
    return max(-64, min(64, d))


def _weights(h):
    x = (np.uint64(h) + (np.uint64(0x9E3779B9) * _A1 & _M32)) & _M32
    return (_mix32(x) & np.uint64(255)) + np.uint64(1)


def uniform_predict(board, player):
    return np.full(8100, np.float32(1.0 / 8100.0), np.float32), 0.0


def hash_predict(board, player):
    h = board_hash(board, player)
    probs = _weights(h).astype(np.float32) * np.float32(1.0 / 1048576.0)
    jitter = int(_mix32(np.array([h ^ 0xA5A5A5A5], np.uint64))[0] & np.uint64(15)) - 8
    v = _material_diff(board, player) / 64.0 * 0.75 + jitter / 1024.0
    return probs, float(np.float32(v))


def ratio_predict(board, player):
    h = board_hash(board, player) ^ 0x5BD1E995
    w = _weights(h)
    w = w * w
    probs = w.astype(np.float32) / np.float32(int(w.sum()))
    jitter = int(_mix32(np.array([h ^ 0xA5A5A5A5], np.uint64))[0] & np.uint64(31)) - 16
    v = np.float32(_material_diff(board, player) + jitter) / np.float32(97.0)
    return probs.astype(np.float32), float(v)


PREDICT = {"uniform": uniform_predict, "hash": hash_predict, "ratio": ratio_predict}


class PlaneEvaluator:
    """`model` object for the reference MCTS: predict(state, device) on feature planes."""

    def __init__(self, name):
        self.fn = PREDICT[name]
        self.calls = 0

    def predict(self, state, device="cpu"):
        self.calls += 1
        board, player = planes_to_board(state)
        return self.fn(board, player)
