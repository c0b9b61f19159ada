"""ctypes binding of oracle/xq_oracle.c -- CPU ORACLE, test infrastructure only.

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference
legs of bench.py may import this module.  The product package
(xiangqi-alphazero_b200/) never does.

`ref_engine()` additionally loads oracle/_ref/game_core*.so -- the reference's
own Cython engine (training/cython_engine/game_core.pyx) compiled unmodified --
when it has been built (oracle/Makefile, target `ref`).
"""
from __future__ import annotations

import ctypes as C
import glob
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

EVAL_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_int8), C.c_int, C.POINTER(C.c_float), C.c_void_p)


def build(force: bool = False) -> str:
    """Compile libxq_oracle.so (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(_HERE, "libxq_oracle.so")
    src = os.path.join(_HERE, "xq_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libxq_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def build_ref() -> None:
    subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        i8p, i16p, u8p, f32p = (C.POINTER(C.c_int8), C.POINTER(C.c_int16), C.POINTER(C.c_uint8),
                                C.POINTER(C.c_float))
        L.xqo_find_king.argtypes = [i8p, C.c_int]
        L.xqo_is_attacked.argtypes = [i8p, C.c_int, C.c_int, C.c_int]
        L.xqo_in_check.argtypes = [i8p, C.c_int]
        L.xqo_move_is_legal.argtypes = [i8p, C.c_int, C.c_int, C.c_int]
        L.xqo_generate_moves.argtypes = [i8p, C.c_int, i16p]
        L.xqo_planes.argtypes = [i8p, C.c_int, f32p]
        L.xqo_movegen_batch.argtypes = [i8p, i8p, C.c_int, i16p, u8p, u8p, f32p]
        L.xqo_is_attacked_batch.argtypes = [i8p, u8p, i8p, C.c_int, u8p]
        L.xqo_game_init.argtypes = [C.c_void_p]
        L.xqo_game_move.argtypes = [C.c_void_p, C.c_int]
        L.xqo_material.argtypes = [i8p, C.c_int]
        L.xqo_game_over.argtypes = [C.c_void_p, C.POINTER(C.c_int), i16p, C.POINTER(C.c_int)]
        L.xqo_random_playout_positions.argtypes = [C.c_uint64, C.c_int, i8p, i8p]
        L.xqo_perft.argtypes = [i8p, C.c_int, C.c_int]
        L.xqo_perft.restype = C.c_uint64
        L.xqo_mcts_search.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_double), i16p, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.xqo_board_hash.argtypes = [i8p, C.c_int]
        L.xqo_board_hash.restype = C.c_uint32
        L.xqo_eval_uniform.restype = C.c_double
        L.xqo_eval_hash.restype = C.c_double
        L.xqo_eval_ratio.restype = C.c_double
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _board(b) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(b, dtype=np.int8).reshape(90))


def find_king(board, player):
    k = lib().xqo_find_king(_p(_board(board), C.c_int8), int(player))
    return None if k < 0 else (k // 9, k % 9)


def is_attacked(board, kr, kc, by):
    return bool(lib().xqo_is_attacked(_p(_board(board), C.c_int8), int(kr), int(kc), int(by)))


def in_check(board, player):
    return bool(lib().xqo_in_check(_p(_board(board), C.c_int8), int(player)))


def legal_actions(board, player) -> np.ndarray:
    out = np.empty(200, np.int16)
    n = lib().xqo_generate_moves(_p(_board(board), C.c_int8), int(player), _p(out, C.c_int16))
    return out[:n].copy()


def legal_moves(board, player):
    return [(a // 90 // 9, a // 90 % 9, a % 90 // 9, a % 90 % 9) for a in map(int, legal_actions(board, player))]


def planes(board, player) -> np.ndarray:
    out = np.empty((15, 10, 9), np.float32)
    lib().xqo_planes(_p(_board(board), C.c_int8), int(player), _p(out, C.c_float))
    return out


def movegen_batch(boards, sides, want_planes=False, allow_overflow=False):
    boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 90)
    sides = np.ascontiguousarray(sides, np.int8)
    B = boards.shape[0]
    acts = np.empty((B, 128), np.int16)
    n = np.empty(B, np.uint8)
    chk = np.empty(B, np.uint8)
    pl = np.empty((B, 15, 10, 9), np.float32) if want_planes else None
    ovf = lib().xqo_movegen_batch(_p(boards, C.c_int8), _p(sides, C.c_int8), B, _p(acts, C.c_int16),
                                  _p(n, C.c_uint8), _p(chk, C.c_uint8),
                                  _p(pl, C.c_float) if want_planes else None)
    assert ovf == 0 or allow_overflow, "a position had more than 128 legal moves"
    return acts, n, chk, pl


def is_attacked_batch(boards, sq, by):
    boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 90)
    sq = np.ascontiguousarray(sq, np.uint8)
    by = np.ascontiguousarray(by, np.int8)
    out = np.empty(boards.shape[0], np.uint8)
    lib().xqo_is_attacked_batch(_p(boards, C.c_int8), _p(sq, C.c_uint8), _p(by, C.c_int8),
                                boards.shape[0], _p(out, C.c_uint8))
    return out


def random_playout_positions(seed: int, count: int):
    boards = np.empty((count, 90), np.int8)
    sides = np.empty(count, np.int8)
    w = lib().xqo_random_playout_positions(C.c_uint64(seed), count, _p(boards, C.c_int8), _p(sides, C.c_int8))
    assert w == count
    return boards, sides


def perft(board, player, depth):
    return int(lib().xqo_perft(_p(_board(board), C.c_int8), int(player), int(depth)))


class _GameStruct(C.Structure):
    _fields_ = [("board", C.c_int8 * 90), ("ring", (C.c_int8 * 90) * 12), ("player", C.c_int32),
                ("move_count", C.c_int32), ("no_capture", C.c_int32)]


class OracleGame:
    """State object with the reference's XiangqiGame semantics (game.py:124-170, 528-616)."""

    def __init__(self):
        self.s = _GameStruct()
        lib().xqo_game_init(C.byref(self.s))

    def clone(self):
        g = OracleGame.__new__(OracleGame)
        g.s = _GameStruct.from_buffer_copy(self.s)
        return g

    @property
    def board(self):
        return np.frombuffer(self.s.board, dtype=np.int8).reshape(10, 9)

    @property
    def ring(self):
        return np.frombuffer(self.s.ring, dtype=np.int8).reshape(12, 90)

    @property
    def current_player(self):
        return self.s.player

    @current_player.setter
    def current_player(self, v):
        self.s.player = int(v)

    @property
    def move_count(self):
        return self.s.move_count

    @property
    def no_capture_count(self):
        return self.s.no_capture

    def make_action(self, a):
        lib().xqo_game_move(C.byref(self.s), int(a))

    def get_legal_actions(self):
        return legal_actions(self.board, self.s.player)

    def is_game_over(self):
        w = C.c_int(0)
        done = lib().xqo_game_over(C.byref(self.s), C.byref(w), None, None)
        return (True, w.value) if done else (False, None)

    def get_state_for_nn(self):
        return planes(self.board, self.s.player)

    def material(self, player):
        return lib().xqo_material(_p(_board(self.board), C.c_int8), int(player))


def c_evaluator(name: str):
    """Address of a built-in C evaluator ('uniform' | 'hash' | 'ratio')."""
    fn = {"uniform": lib().xqo_eval_uniform, "hash": lib().xqo_eval_hash,
          "ratio": lib().xqo_eval_ratio}[name]
    return C.cast(fn, C.c_void_p)


def py_evaluator(predict):
    """Wrap predict(board int8[10,9], player) -> (float32[8100], float) as a C callback."""

    def _cb(bp, player, probs_p, _user):
        board = np.ctypeslib.as_array(bp, shape=(90,)).reshape(10, 9)
        probs, v = predict(board, player)
        np.ctypeslib.as_array(probs_p, shape=(8100,))[:] = probs
        return float(v)

    return EVAL_FN(_cb)


def mcts_search(game: OracleGame, num_sims: int, c_puct: float = 1.5, evaluator="uniform",
                root_noise=None):
    """mcts.py:94-155 on the oracle.  Returns (actions int16[n], visits int32[n], W float64[n], stats)."""
    keep = None
    if isinstance(evaluator, str):
        fn = c_evaluator(evaluator)
    else:
        keep = evaluator if isinstance(evaluator, EVAL_FN) else py_evaluator(evaluator)
        fn = C.cast(keep, C.c_void_p)
    acts = np.empty(200, np.int16)
    vis = np.empty(200, np.int32)
    tot = np.empty(200, np.float64)
    stats = np.zeros(3, np.int64)
    noise_p = None
    if root_noise is not None:
        root_noise = np.ascontiguousarray(root_noise, np.float64)
        noise_p = _p(root_noise, C.c_double)
    n = lib().xqo_mcts_search(C.byref(game.s), int(num_sims), float(c_puct), fn, None, noise_p,
                              _p(acts, C.c_int16), _p(vis, C.c_int32), _p(tot, C.c_double),
                              _p(stats, C.c_int64))
    del keep
    return acts[:n].copy(), vis[:n].copy(), tot[:n].copy(), {"nodes": int(stats[0]),
                                                             "terminal_sims": int(stats[1]),
                                                             "max_depth": int(stats[2])}


def board_hash(board, player) -> int:
    return int(lib().xqo_board_hash(_p(_board(board), C.c_int8), int(player)))


def ref_engine():
    """The reference's Cython engine compiled as-is (oracle/_ref), or None if not built."""
    hits = glob.glob(os.path.join(_HERE, "_ref", "game_core*.so"))
    if not hits:
        return None
    spec = importlib.util.spec_from_file_location("game_core", hits[0])
    mod = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(mod)
    except ImportError:
        return None
    return mod
