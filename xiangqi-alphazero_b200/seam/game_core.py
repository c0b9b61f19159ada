"""game_core -- the reference's native seam answered by libxq_b200.so (Level 2).

The reference's game.py (:30-47) puts training/cython_engine on sys.path and does

    from game_core import cy_generate_legal_moves, cy_is_in_check, cy_find_king, cy_is_attacked, cy_has_legal_moves

This file is a module of that name with those five functions (signatures and return types of
cython_engine/game_core.pyx:493-569), implemented as a ctypes binding of the host-pointer entry points of
include/xq_b200.h: put this directory on sys.path BEFORE training/cython_engine (or copy the file there in place of
the compiled module) and the reference's game.py, mcts.py, parallel_selfplay.py run UNMODIFIED on the B200 rules
kernels.  Pure ctypes + numpy: no torch, nothing else of this repository.

Contract kept: the caller owns the numpy board, it is copied, never mutated or retained; a missing king
gives None / an empty move list / in-check True; moves come in the Cython engine's generation order; a dtype other than
int8 raises ValueError like the typed Cython signature does.  One call = one kernel launch on one position: this is the
compatibility path -- the throughput interface is the batched one (xq_movegen_batch on device pointers).
"""
import ctypes as _C
import os as _os

import numpy as _np

_LIB = _os.environ.get("XQ_B200_LIB") or _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "libxq_b200.so")
if not _os.path.exists(_LIB):
    raise ImportError(f"{_LIB} is missing: build it with `python __graft_entry__.py` (there is no CPU fallback)")
_L = _C.CDLL(_LIB)
_vp = _C.c_void_p
_L.xq_create.argtypes = [_C.c_int, _C.POINTER(_vp)]
_L.xq_last_error.restype = _C.c_char_p
_L.xq_last_error.argtypes = [_vp]
for _name, _n in (("xq_movegen_batch_host", 8), ("xq_is_attacked_batch_host", 6), ("xq_find_king_batch_host", 5),
                  ("xq_has_legal_moves_batch_host", 5)):
    getattr(_L, _name).restype = _C.c_int
_L.xq_movegen_batch_host.argtypes = [_vp, _vp, _vp, _C.c_int, _vp, _vp, _vp, _vp]
_L.xq_is_attacked_batch_host.argtypes = [_vp, _vp, _vp, _vp, _C.c_int, _vp]
_L.xq_find_king_batch_host.argtypes = [_vp, _vp, _vp, _C.c_int, _vp]
_L.xq_has_legal_moves_batch_host.argtypes = [_vp, _vp, _vp, _C.c_int, _vp]

_ctx = _vp()
if _L.xq_create(int(_os.environ.get("XQ_B200_DEVICE", "0")), _C.byref(_ctx)) != 0:
    raise ImportError("xq_create failed: " + _L.xq_last_error(None).decode())


def _board(board):
    if not isinstance(board, _np.ndarray) or board.dtype != _np.int8:
        raise ValueError("Buffer dtype mismatch, expected 'signed char'")      # what the typed Cython argument raises
    return _np.ascontiguousarray(board).reshape(1, 90)


def _p(a):
    return _vp(a.ctypes.data)


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"xq error {rc}: {_L.xq_last_error(_ctx).decode()}")


def _movegen(board, player):
    b = _board(board)
    s = _np.array([player], _np.int8)
    acts = _np.empty((1, 128), _np.int16)
    n = _np.empty(1, _np.uint8)
    chk = _np.empty(1, _np.uint8)
    _check(_L.xq_movegen_batch_host(_ctx, _p(b), _p(s), 1, _p(acts), _p(n), _p(chk), None))
    return acts[0, :int(n[0])], bool(chk[0])


def cy_generate_legal_moves(board, player):              # game_core.pyx:521-540
    acts, _ = _movegen(board, int(player))
    return [(a // 810, a // 90 % 9, a % 90 // 9, a % 9) for a in acts.tolist()]


def cy_is_in_check(board, player):                       # game_core.pyx:543-555
    return _movegen(board, int(player))[1]


def cy_find_king(board, player):                         # game_core.pyx:493-505
    b = _board(board)
    out = _np.empty(1, _np.int8)
    _check(_L.xq_find_king_batch_host(_ctx, _p(b), _p(_np.array([player], _np.int8)), 1, _p(out)))
    sq = int(out[0])
    return None if sq < 0 else (sq // 9, sq % 9)


def cy_is_attacked(board, kr, kc, by_player):            # game_core.pyx:508-518
    b = _board(board)
    out = _np.empty(1, _np.uint8)
    _check(_L.xq_is_attacked_batch_host(_ctx, _p(b), _p(_np.array([kr * 9 + kc], _np.uint8)),
                                        _p(_np.array([by_player], _np.int8)), 1, _p(out)))
    return bool(out[0])


def cy_has_legal_moves(board, player):                   # game_core.pyx:558-569
    b = _board(board)
    out = _np.empty(1, _np.uint8)
    _check(_L.xq_has_legal_moves_batch_host(_ctx, _p(b), _p(_np.array([player], _np.int8)), 1, _p(out)))
    return bool(out[0])
