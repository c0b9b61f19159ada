"""ctypes binding of libxq_b200.so (include/xq_b200.h) -- the only way the Python host
code reaches the GPU kernels.  torch is used for device memory and streams only.

There is no CPU path: importing is harmless, but `Engine()` raises if the shared library
has not been built (python __graft_entry__.py / make -C csrc) or no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libxq_b200.so")

MAX_MOVES = 128
MAX_PLIES = 201
ACTION_SPACE = 8100

_lib = None


class XqError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libxq_b200.so (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise XqError("building libxq_b200.so failed")
    return LIB_PATH


def _sig(L, name, restype, *argtypes):
    fn = getattr(L, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


def lib():
    """Load libxq_b200.so and declare every symbol of include/xq_b200.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise XqError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                      "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    _sig(L, "xq_create", i32, i32, C.POINTER(vp))
    _sig(L, "xq_destroy", None, vp)
    _sig(L, "xq_last_error", C.c_char_p, vp)
    _sig(L, "xq_version", i32)
    _sig(L, "xq_launch_count", C.c_longlong, vp, i32)
    _sig(L, "xq_set_timing", i32, vp, i32)
    _sig(L, "xq_last_kernel_ms", C.c_float, vp)
    _sig(L, "xq_movegen_batch", i32, vp, vp, vp, i32, vp, vp, vp, vp, vp)
    _sig(L, "xq_movegen_batch_host", i32, vp, vp, vp, i32, vp, vp, vp, vp)
    _sig(L, "xq_is_attacked_batch", i32, vp, vp, vp, vp, i32, vp, vp)
    _sig(L, "xq_is_attacked_batch_host", i32, vp, vp, vp, vp, i32, vp)
    _sig(L, "xq_movegen_batch_host_packed", i32, vp, vp, vp, i32, vp, vp, vp, vp)
    _sig(L, "xq_planes_bits", i32, vp, vp, vp, i32, vp, vp)
    _sig(L, "xq_find_king_batch", i32, vp, vp, vp, i32, vp, vp)
    _sig(L, "xq_find_king_batch_host", i32, vp, vp, vp, i32, vp)
    _sig(L, "xq_has_legal_moves_batch", i32, vp, vp, vp, i32, vp, vp)
    _sig(L, "xq_has_legal_moves_batch_host", i32, vp, vp, vp, i32, vp)
    _sig(L, "xq_move_is_legal_batch", i32, vp, vp, vp, vp, vp, i32, vp, vp)
    _sig(L, "xq_move_is_legal_batch_host", i32, vp, vp, vp, vp, vp, i32, vp)
    _sig(L, "xq_overflow_count", i32, vp, i32)
    _sig(L, "xq_set_movegen_impl", i32, vp, i32)
    _sig(L, "xq_random_playouts", i32, vp, u64, i32, vp, vp, vp, vp, vp)
    i64, dbl = C.c_longlong, C.c_double
    _sig(L, "xq_mcts_create", i32, vp, i32, i64)
    _sig(L, "xq_mcts_set_games", i32, vp, i32, vp, vp, vp, vp, vp, vp, vp)
    _sig(L, "xq_mcts_root_begin", i32, vp, vp, vp, i64, i64, vp, vp, vp)
    _sig(L, "xq_mcts_root_expand", i32, vp, vp, i32, i64, vp, i32, u64, dbl, vp)
    _sig(L, "xq_mcts_select", i32, vp, dbl, vp, vp, i64, i64, vp, vp, vp)
    _sig(L, "xq_mcts_expand_backup", i32, vp, vp, i32, i64, vp, vp)
    _sig(L, "xq_mcts_leaf_info", i32, vp, vp, vp, vp, vp)
    _sig(L, "xq_mcts_root_visits", i32, vp, vp, vp, vp, vp, vp)
    _sig(L, "xq_mcts_root_priors", i32, vp, vp, vp, vp)
    _sig(L, "xq_mcts_stats", i32, vp, C.POINTER(i64), i32)
    _sig(L, "xq_net_gemm", i32, vp, vp, vp)
    _sig(L, "xq_net_value_head", i32, vp, vp, vp, vp, vp, C.c_float, vp, i32, vp)
    _sig(L, "xq_net_run", i32, vp, vp, i32, vp, vp, vp, vp, C.c_float, vp, i32, vp)
    _sig(L, "xq_net_run_counted", i32, vp, vp, i32, vp, vp, vp, vp, C.c_float, vp, vp, i32, vp)
    _sig(L, "xq_selfplay_create", i32, vp, i32, i32, i64, i64)
    _sig(L, "xq_selfplay_reset", i32, vp, vp)
    _sig(L, "xq_selfplay_play", i32, vp, vp, vp, i32, vp)
    _sig(L, "xq_selfplay_set_live_bound", i32, vp, i32)
    _sig(L, "xq_selfplay_counters", i32, vp, C.POINTER(i64))
    _sig(L, "xq_selfplay_fetch", i32, vp, i64, i64, vp, vp, vp, i32)
    _sig(L, "xq_selfplay_slots", i32, vp, vp, vp, vp, vp)
    _sig(L, "xq_selfplay_device_buffers", i32, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp))
    _sig(L, "xq_arena_play", i32, vp, vp, vp, vp, i32, vp, vp)
    f32 = C.c_float
    _sig(L, "xq_replay_append", i32, vp, vp, vp, i32, vp, i32, vp, vp, i64, i64, vp)
    _sig(L, "xq_train_batch", i32, vp, vp, vp, i64, i64, vp, i32, vp, vp, vp, vp, vp, vp)
    _sig(L, "xq_policy_value_loss", i32, vp, vp, i64, vp, vp, vp, vp, vp, i32, f32, vp, i64, vp, vp, vp, vp)
    _sig(L, "xq_grad_sumsq", i32, vp, vp, i64, vp, i32, vp, vp)
    _sig(L, "xq_adam_step", i32, vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i64, vp, f32, f32, vp)
    _sig(L, "xq_peer_create", i32, vp, i32, i32, vp)
    _sig(L, "xq_peer_connect", i32, vp, vp)
    _sig(L, "xq_bn_forward", i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, f32, vp)
    _sig(L, "xq_bn_backward", i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, vp)
    _sig(L, "xq_tgemm", i32, vp, vp, vp)
    _sig(L, "xq_twgrad", i32, vp, vp, vp)
    _sig(L, "xq_tn_input", i32, vp, vp, i32, i32, i32, vp, vp, i64, vp)
    _sig(L, "xq_tn_wimage", i32, vp, vp, i32, i32, i32, vp, i32, i32, i32, i32, vp)
    _sig(L, "xq_tn_wimage_batch", i32, vp, vp, i32, vp)
    _sig(L, "xq_tn_wimage_dense2", i32, vp, vp, i32, i32, vp, i32, vp, i32, vp)
    _sig(L, "xq_tn_bn_forward", i32, vp, vp, vp)
    _sig(L, "xq_tn_bn_backward", i32, vp, vp, vp)
    _sig(L, "xq_tn_wgrad_reduce", i32, vp, vp, i32, i64, i32, i32, i32, i32, i32, i32, vp, i32, i32, i32, vp)
    _sig(L, "xq_tn_flatten", i32, vp, vp, i64, i32, i32, vp, vp, i64, vp)
    _sig(L, "xq_tn_unflatten", i32, vp, vp, i64, i32, i32, vp, i64, i32, i64, vp)
    _sig(L, "xq_tn_rows_layouts", i32, vp, vp, i64, i32, i32, vp, vp, i64, vp)
    _sig(L, "xq_tn_colsum", i32, vp, vp, i64, i32, i32, vp, vp)
    _sig(L, "xq_tn_value_forward", i32, vp, vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp)
    _sig(L, "xq_tn_value_backward", i32, vp, vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp)
    _lib = L
    return L


EXPORTS = ["xq_create", "xq_destroy", "xq_last_error", "xq_version", "xq_launch_count", "xq_set_timing",
           "xq_last_kernel_ms", "xq_movegen_batch", "xq_movegen_batch_host", "xq_is_attacked_batch",
           "xq_is_attacked_batch_host", "xq_movegen_batch_host_packed", "xq_planes_bits", "xq_find_king_batch",
           "xq_find_king_batch_host", "xq_has_legal_moves_batch", "xq_has_legal_moves_batch_host", "xq_move_is_legal_batch",
           "xq_move_is_legal_batch_host", "xq_overflow_count", "xq_set_movegen_impl", "xq_random_playouts",
           "xq_mcts_create", "xq_mcts_set_games", "xq_mcts_root_begin", "xq_mcts_root_expand", "xq_mcts_select",
           "xq_mcts_expand_backup", "xq_mcts_leaf_info", "xq_mcts_root_visits", "xq_mcts_root_priors", "xq_mcts_stats",
           "xq_net_gemm", "xq_net_value_head", "xq_net_run", "xq_net_run_counted", "xq_selfplay_create", "xq_selfplay_reset",
           "xq_selfplay_play", "xq_selfplay_set_live_bound", "xq_selfplay_counters", "xq_selfplay_fetch", "xq_selfplay_slots",
           "xq_selfplay_device_buffers", "xq_arena_play", "xq_replay_append", "xq_train_batch", "xq_policy_value_loss",
           "xq_grad_sumsq", "xq_adam_step", "xq_peer_create", "xq_peer_connect", "xq_bn_forward", "xq_bn_backward",
           "xq_tgemm", "xq_twgrad", "xq_tn_input", "xq_tn_wimage", "xq_tn_wimage_batch", "xq_tn_wimage_dense2", "xq_tn_bn_forward", "xq_tn_bn_backward", "xq_tn_wgrad_reduce",
           "xq_tn_flatten", "xq_tn_unflatten", "xq_tn_rows_layouts", "xq_tn_colsum", "xq_tn_value_forward", "xq_tn_value_backward"]


class TGemmDesc(C.Structure):
    """xq_tgemm_desc (include/xq_b200.h)"""
    _fields_ = [("a", C.c_void_p), ("a_rows", C.c_int64), ("a_row0", C.c_int64), ("w", C.c_void_p),
                ("kblocks", C.c_int32), ("ntaps", C.c_int32), ("img_kb", C.c_int32), ("b_mn", C.c_int32),
                ("shift_sign", C.c_int32), ("m_pairs", C.c_int32), ("n_tiles", C.c_int32), ("out_chunks", C.c_int32),
                ("m_rows", C.c_int64), ("out", C.c_void_p), ("out_rows", C.c_int64), ("out_row0", C.c_int64),
                ("residual", C.c_void_p), ("out_rm", C.c_void_p), ("out_stride", C.c_int64), ("bias", C.c_void_p),
                ("n_cols", C.c_int32), ("k_splits", C.c_int32), ("out_split_stride", C.c_int64)]


class TnWimageItem(C.Structure):
    """xq_tn_wimage_item (include/xq_b200.h)"""
    _fields_ = [("w", C.c_void_p), ("img", C.c_void_p), ("co", C.c_int32), ("ci", C.c_int32), ("taps", C.c_int32),
                ("img_kb", C.c_int32), ("n0", C.c_int32), ("k0", C.c_int32), ("transposed", C.c_int32), ("pad_", C.c_int32)]


class TnBnDesc(C.Structure):
    """xq_tn_bn_desc (include/xq_b200.h)"""
    _fields_ = [("y", C.c_void_p), ("res", C.c_void_p), ("out", C.c_void_p), ("out_g", C.c_void_p), ("rows", C.c_int64),
                ("n_boards", C.c_int32), ("chunk0", C.c_int32), ("n_channels", C.c_int32), ("relu", C.c_int32),
                ("partial", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("running_mean", C.c_void_p),
                ("running_var", C.c_void_p), ("save", C.c_void_p), ("eps", C.c_float), ("momentum", C.c_float)]


class TnBnBwdDesc(C.Structure):
    """xq_tn_bn_bwd_desc (include/xq_b200.h)"""
    _fields_ = [("dout", C.c_void_p), ("act", C.c_void_p), ("y", C.c_void_p), ("rows", C.c_int64),
                ("n_boards", C.c_int32), ("chunk0", C.c_int32), ("n_channels", C.c_int32), ("relu", C.c_int32),
                ("save", C.c_void_p), ("partial", C.c_void_p), ("gamma", C.c_void_p), ("dgamma", C.c_void_p),
                ("dbeta", C.c_void_p), ("dy", C.c_void_p), ("dy_g", C.c_void_p), ("dskip", C.c_void_p)]


class TWgradDesc(C.Structure):
    """xq_twgrad_desc (include/xq_b200.h)"""
    _fields_ = [("a", C.c_void_p), ("b", C.c_void_p), ("a_rows", C.c_int64), ("a_row0", C.c_int64), ("b_rows", C.c_int64),
                ("b_row0", C.c_int64), ("a_group0", C.c_int32), ("b_group0", C.c_int32), ("nbg", C.c_int32), ("kr", C.c_int32),
                ("stages_per_item", C.c_int32), ("n_slabs", C.c_int32), ("n_groups", C.c_int32), ("n_mtiles", C.c_int32),
                ("taps_per_group", C.c_int32), ("b_rows_stage", C.c_int32), ("b_groups_stage", C.c_int32),
                ("b_group_step", C.c_int32), ("b_row_lo", C.c_int32 * 4), ("tap_off", C.c_int32 * 16), ("out", C.c_void_p),
                ("mt_stride", C.c_int64), ("slab_stride", C.c_int64), ("g_stride", C.c_int64), ("tap_stride", C.c_int64),
                ("ldo", C.c_int64), ("m_limit", C.c_int32), ("n_limit", C.c_int32), ("g_cols", C.c_int32), ("t_cols", C.c_int32)]


def _np_ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


PLANE_WORDS = 44


def unpack_planes(bits: np.ndarray) -> np.ndarray:
    """uint32 [B,44] plane bits (xq_movegen_batch_host_packed / xq_planes_bits) -> float32 [B,15,10,9], the
    get_state_for_nn planes of game.py:618-640 (bit plane*90+square at word bit>>5, bit bit&31)."""
    bits = np.ascontiguousarray(bits, np.uint32).reshape(-1, PLANE_WORDS)
    b = np.unpackbits(bits.view(np.uint8), axis=1, bitorder="little")[:, :1350]
    return b.reshape(-1, 15, 10, 9).astype(np.float32)


class Engine:
    """One context per GPU/process (xq_create).  Methods mirror the C entry points."""

    def __init__(self, device: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise XqError("no CUDA device: the xq_b200 engine has no CPU path")
        self.torch = torch
        self.L = lib()
        self.device = int(device)
        self.dev = torch.device("cuda", self.device)
        h = C.c_void_p()
        rc = self.L.xq_create(self.device, C.byref(h))
        if rc != 0:
            raise XqError(self.L.xq_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.xq_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise XqError(f"xq error {rc}: {self.L.xq_last_error(self.h).decode()}")

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    # ---- diagnostics ----------------------------------------------------------------------
    def launch_count(self, reset=False) -> int:
        return int(self.L.xq_launch_count(self.h, int(reset)))

    def set_timing(self, on: bool):
        self._check(self.L.xq_set_timing(self.h, int(on)))

    def last_kernel_ms(self) -> float:
        return float(self.L.xq_last_kernel_ms(self.h))

    def set_movegen_impl(self, impl) -> str:
        """'warp' (one warp per board) or 'thread' (one thread per board); returns the previous setting."""
        prev = self.L.xq_set_movegen_impl(self.h, {"warp": 0, "thread": 1}[impl] if isinstance(impl, str) else int(impl))
        if prev < 0:
            raise XqError(self.L.xq_last_error(self.h).decode())
        return ("warp", "thread")[prev]

    @property
    def movegen_impl(self) -> str:
        cur = self.set_movegen_impl(0)
        self.set_movegen_impl(cur)
        return cur

    def overflow_count(self, reset=False) -> int:
        return int(self.L.xq_overflow_count(self.h, int(reset)))

    # ---- K1 ---------------------------------------------------------------------------------
    def movegen(self, boards, sides, planes: bool = False, out=None):
        """Device tensors in, device tensors out (async on the current torch stream).

        boards int8 [B,90] (or [B,10,9]), sides int8 [B] ->
        actions int16 [B,128], n_moves uint8 [B], in_check uint8 [B], planes float32 [B,15,10,9] | None
        """
        t = self.torch
        boards = boards.reshape(-1, 90)
        assert boards.dtype == t.int8 and sides.dtype == t.int8 and boards.is_cuda and sides.is_cuda
        boards = boards.contiguous()
        sides = sides.contiguous()
        B = boards.shape[0]
        if out is None:
            actions = t.empty((B, MAX_MOVES), dtype=t.int16, device=self.dev)
            n = t.empty((B,), dtype=t.uint8, device=self.dev)
            chk = t.empty((B,), dtype=t.uint8, device=self.dev)
            pl = t.empty((B, 15, 10, 9), dtype=t.float32, device=self.dev) if planes else None
        else:
            actions, n, chk, pl = out
        self._check(self.L.xq_movegen_batch(self.h, boards.data_ptr(), sides.data_ptr(), B, actions.data_ptr(),
                                            n.data_ptr(), chk.data_ptr(), pl.data_ptr() if pl is not None else None,
                                            self._stream()))
        return actions, n, chk, pl

    def movegen_host(self, boards: np.ndarray, sides: np.ndarray, planes=False, out=None):
        """Host arrays in, host arrays out (the reference-facing form: numpy boards, like
        cy_generate_legal_moves).  Copies run inside the call; pass pinned arrays for full speed.
        planes: False, True (float32 [B,15,10,9], 5400 bytes per position over PCIe) or "packed" (uint32 [B,44]
        plane bits, 176 bytes per position; unpack_planes() expands them on the host)."""
        boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 90)
        sides = np.ascontiguousarray(sides, np.int8)
        B = boards.shape[0]
        packed = isinstance(planes, str) and planes == "packed"
        if out is None:
            actions = np.empty((B, MAX_MOVES), np.int16)
            n = np.empty(B, np.uint8)
            chk = np.empty(B, np.uint8)
            pl = np.empty((B, PLANE_WORDS), np.uint32) if packed else (np.empty((B, 15, 10, 9), np.float32) if planes else None)
        else:
            actions, n, chk, pl = out
        fn = self.L.xq_movegen_batch_host_packed if packed else self.L.xq_movegen_batch_host
        rc = fn(self.h, _np_ptr(boards), _np_ptr(sides), B, _np_ptr(actions), _np_ptr(n), _np_ptr(chk),
                _np_ptr(pl) if pl is not None else None)
        self._check(rc)
        return actions, n, chk, pl

    def planes_bits(self, boards, sides):
        """Device tensors: uint32-packed get_state_for_nn planes, int32 tensor [B,44] (bit pattern of the uint32 words)."""
        t = self.torch
        boards = boards.reshape(-1, 90).contiguous()
        out = t.empty((boards.shape[0], PLANE_WORDS), dtype=t.int32, device=self.dev)
        self._check(self.L.xq_planes_bits(self.h, boards.data_ptr(), sides.contiguous().data_ptr(), boards.shape[0],
                                          out.data_ptr(), self._stream()))
        return out

    def find_king(self, boards, sides):
        """cy_find_king batched: int8 [B] square (row*9+col) of side's king inside its own palace, -1 if absent."""
        t = self.torch
        boards = boards.reshape(-1, 90).contiguous()
        out = t.empty((boards.shape[0],), dtype=t.int8, device=self.dev)
        self._check(self.L.xq_find_king_batch(self.h, boards.data_ptr(), sides.contiguous().data_ptr(), boards.shape[0],
                                              out.data_ptr(), self._stream()))
        return out

    def has_legal_moves(self, boards, sides):
        """cy_has_legal_moves batched: uint8 [B]."""
        t = self.torch
        boards = boards.reshape(-1, 90).contiguous()
        out = t.empty((boards.shape[0],), dtype=t.uint8, device=self.dev)
        self._check(self.L.xq_has_legal_moves_batch(self.h, boards.data_ptr(), sides.contiguous().data_ptr(),
                                                    boards.shape[0], out.data_ptr(), self._stream()))
        return out

    def move_is_legal(self, boards, frm, to, sides):
        """_is_move_legal batched (device tensors): uint8 [B]; frm / to uint8 squares, any move."""
        t = self.torch
        boards = boards.reshape(-1, 90).contiguous()
        out = t.empty((boards.shape[0],), dtype=t.uint8, device=self.dev)
        self._check(self.L.xq_move_is_legal_batch(self.h, boards.data_ptr(), frm.contiguous().data_ptr(), to.contiguous().data_ptr(),
                                                  sides.contiguous().data_ptr(), boards.shape[0], out.data_ptr(), self._stream()))
        return out

    def move_is_legal_host(self, boards: np.ndarray, frm: np.ndarray, to: np.ndarray, sides: np.ndarray) -> np.ndarray:
        boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 90)
        frm, to = np.ascontiguousarray(frm, np.uint8), np.ascontiguousarray(to, np.uint8)
        sides = np.ascontiguousarray(sides, np.int8)
        out = np.empty(boards.shape[0], np.uint8)
        self._check(self.L.xq_move_is_legal_batch_host(self.h, _np_ptr(boards), _np_ptr(frm), _np_ptr(to), _np_ptr(sides),
                                                       boards.shape[0], _np_ptr(out)))
        return out

    def find_king_host(self, boards: np.ndarray, sides: np.ndarray) -> np.ndarray:
        boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 90)
        sides = np.ascontiguousarray(sides, np.int8)
        out = np.empty(boards.shape[0], np.int8)
        self._check(self.L.xq_find_king_batch_host(self.h, _np_ptr(boards), _np_ptr(sides), boards.shape[0], _np_ptr(out)))
        return out

    def has_legal_moves_host(self, boards: np.ndarray, sides: np.ndarray) -> np.ndarray:
        boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 90)
        sides = np.ascontiguousarray(sides, np.int8)
        out = np.empty(boards.shape[0], np.uint8)
        self._check(self.L.xq_has_legal_moves_batch_host(self.h, _np_ptr(boards), _np_ptr(sides), boards.shape[0], _np_ptr(out)))
        return out

    def is_attacked(self, boards, sq, by):
        t = self.torch
        boards = boards.reshape(-1, 90).contiguous()
        B = boards.shape[0]
        out = t.empty((B,), dtype=t.uint8, device=self.dev)
        self._check(self.L.xq_is_attacked_batch(self.h, boards.data_ptr(), sq.contiguous().data_ptr(),
                                                by.contiguous().data_ptr(), B, out.data_ptr(), self._stream()))
        return out

    def is_attacked_host(self, boards: np.ndarray, sq: np.ndarray, by: np.ndarray) -> np.ndarray:
        boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 90)
        sq = np.ascontiguousarray(sq, np.uint8)
        by = np.ascontiguousarray(by, np.int8)
        out = np.empty(boards.shape[0], np.uint8)
        self._check(self.L.xq_is_attacked_batch_host(self.h, _np_ptr(boards), _np_ptr(sq), _np_ptr(by),
                                                     boards.shape[0], _np_ptr(out)))
        return out

    def random_playouts(self, seed: int, n_games: int, compact: bool = True):
        """Uniform-random legal playouts on the device.  Returns (boards int8 [N,90], sides int8 [N],
        n_positions int32 [G], winner int8 [G]) as device tensors; with compact=False boards/sides keep
        the [G*201] slot layout (side 0 marks unused slots)."""
        t = self.torch
        boards = t.empty((n_games * MAX_PLIES, 90), dtype=t.int8, device=self.dev)
        sides = t.empty((n_games * MAX_PLIES,), dtype=t.int8, device=self.dev)
        npos = t.empty((n_games,), dtype=t.int32, device=self.dev)
        win = t.empty((n_games,), dtype=t.int8, device=self.dev)
        self._check(self.L.xq_random_playouts(self.h, C.c_uint64(seed), n_games, boards.data_ptr(), sides.data_ptr(),
                                              npos.data_ptr(), win.data_ptr(), self._stream()))
        if compact:
            keep = sides != 0
            boards, sides = boards[keep].contiguous(), sides[keep].contiguous()
        return boards, sides, npos, win


POLICY_PROBS, POLICY_LOGITS_BF16, POLICY_LOGITS_F32 = 0, 1, 2


class MctsBatch:
    """Lockstep search of G games on the device (xq_mcts_*).  The evaluator is a callable
    between `select` and `expand_backup`; `search()` runs the whole loop of mcts.py:94-155."""

    def __init__(self, eng: Engine, max_games: int, node_capacity: int = 0):
        self.e = eng
        self.t = eng.torch
        self.G = int(max_games)
        eng._check(eng.L.xq_mcts_create(eng.h, self.G, int(node_capacity)))
        t, dev = self.t, eng.dev
        self.planes = t.zeros((self.G, 15, 10, 9), dtype=t.float32, device=dev)
        self.boards = t.zeros((self.G, 90), dtype=t.int8, device=dev)
        self.sides = t.zeros((self.G,), dtype=t.int8, device=dev)
        self.n = 0

    def set_games(self, boards, sides, move_count=None, no_capture=None, ring=None, active=None):
        t, dev = self.t, self.e.dev

        def d(a, dt):
            if a is None:
                return None
            if not t.is_tensor(a):
                a = t.from_numpy(np.ascontiguousarray(a))
            return a.to(device=dev, dtype=dt).contiguous()
        b = d(boards, t.int8).reshape(-1, 90)
        self.n = b.shape[0]
        keep = [b, d(sides, t.int8), d(move_count, t.int32), d(no_capture, t.int32), d(ring, t.int8), d(active, t.uint8)]
        ptr = [None if x is None else x.data_ptr() for x in keep]
        self.e._check(self.e.L.xq_mcts_set_games(self.e.h, self.n, *ptr, self.e._stream()))
        self._keep = keep

    def _xargs(self, x_planes, x_row0):
        if x_planes is None:
            return None, 0, 0
        return x_planes.data_ptr(), x_planes.shape[1], x_row0

    def root_begin(self, want_planes=True, x_planes=None, x_row0=16):
        xp, xr, x0 = self._xargs(x_planes, x_row0)
        self.e._check(self.e.L.xq_mcts_root_begin(self.e.h, self.planes.data_ptr() if want_planes else None, xp, xr, x0,
                                                  self.boards.data_ptr(), self.sides.data_ptr(), self.e._stream()))

    def root_expand(self, policy, kind=POLICY_PROBS, noise=None, add_noise=False, seed=0, alpha=0.3):
        self.e._check(self.e.L.xq_mcts_root_expand(self.e.h, policy.data_ptr(), kind, policy.stride(0),
                                                   None if noise is None else noise.data_ptr(), int(add_noise),
                                                   C.c_uint64(seed), float(alpha), self.e._stream()))

    def select(self, c_puct=1.5, want_planes=True, x_planes=None, x_row0=16):
        xp, xr, x0 = self._xargs(x_planes, x_row0)
        self.e._check(self.e.L.xq_mcts_select(self.e.h, float(c_puct), self.planes.data_ptr() if want_planes else None,
                                              xp, xr, x0, self.boards.data_ptr(), self.sides.data_ptr(), self.e._stream()))

    def expand_backup(self, policy, value, kind=POLICY_PROBS):
        self.e._check(self.e.L.xq_mcts_expand_backup(self.e.h, policy.data_ptr(), kind, policy.stride(0),
                                                     value.data_ptr(), self.e._stream()))

    def leaf_info(self):
        t, dev = self.t, self.e.dev
        st = t.empty((self.n,), dtype=t.int32, device=dev)
        n = t.empty((self.n,), dtype=t.int32, device=dev)
        acts = t.empty((self.n, MAX_MOVES), dtype=t.int16, device=dev)
        self.e._check(self.e.L.xq_mcts_leaf_info(self.e.h, st.data_ptr(), n.data_ptr(), acts.data_ptr(), self.e._stream()))
        return st, n, acts

    def root_visits(self, want_w=False):
        t, dev = self.t, self.e.dev
        acts = t.empty((self.n, MAX_MOVES), dtype=t.int16, device=dev)
        vis = t.empty((self.n, MAX_MOVES), dtype=t.int32, device=dev)
        n = t.empty((self.n,), dtype=t.int32, device=dev)
        w = t.empty((self.n, MAX_MOVES), dtype=t.float64, device=dev) if want_w else None
        self.e._check(self.e.L.xq_mcts_root_visits(self.e.h, acts.data_ptr(), vis.data_ptr(), n.data_ptr(),
                                                   None if w is None else w.data_ptr(), self.e._stream()))
        return acts, vis, n, w

    def root_priors(self):
        """float64 [n,128] priors of the root children (after root_expand: with the Dirichlet mix when add_noise), n [n]."""
        t, dev = self.t, self.e.dev
        pri = t.empty((self.n, MAX_MOVES), dtype=t.float64, device=dev)
        n = t.empty((self.n,), dtype=t.int32, device=dev)
        self.e._check(self.e.L.xq_mcts_root_priors(self.e.h, pri.data_ptr(), n.data_ptr(), self.e._stream()))
        return pri, n

    def stats(self, reset=False):
        buf = (C.c_longlong * 6)()
        self.e._check(self.e.L.xq_mcts_stats(self.e.h, buf, int(reset)))
        return dict(sims=buf[0], terminal_sims=buf[1], max_depth=buf[2], evals=buf[3], nodes=buf[4], error=buf[5])

    def search(self, evaluator, num_sims, c_puct=1.5, noise=None, add_noise=False, seed=0, kind=POLICY_PROBS):
        """evaluator(self) -> (policy tensor [n, >=8100], value float32 tensor [n]) for the current
        planes/boards/sides.  Returns (actions, visits, n) device tensors."""
        self.root_begin()
        policy, _ = evaluator(self)
        self.root_expand(policy, kind, noise=noise, add_noise=add_noise, seed=seed)
        for _ in range(num_sims):
            self.select(c_puct)
            policy, value = evaluator(self)
            self.expand_backup(policy, value, kind)
        st = self.stats()
        if st["error"]:
            raise XqError(f"MCTS device error bits {st['error']} (1 = node pool overflow, 2 = >128 legal moves)")
        return self.root_visits()[:3]
