"""Drop-in replacement for the reference's training/game.py (same names, same semantics), with
the rules engine on the GPU.

`XiangqiGame` keeps the reference's attribute and method surface (game.py:124-686) so that
train.py, benchmark.py, the demos and the reference tests run unchanged; every rule question
(legal moves, in-check, attacked squares) is answered by the CUDA kernels of libxq_b200.so
through xq_native -- there is no Python or Cython rules code here and no CPU fallback.  For
throughput use the batched API (xq_native.Engine.movegen / the device-resident self-play
loop); this per-object interface exists for API compatibility and pays one kernel launch per
uncached query.
"""
from typing import List, Optional, Tuple

import numpy as np

import xq_native

# piece codes and constants: game.py:50-80
EMPTY = 0
R_KING, R_ADVISOR, R_BISHOP, R_KNIGHT, R_ROOK, R_CANNON, R_PAWN = 1, 2, 3, 4, 5, 6, 7
B_KING, B_ADVISOR, B_BISHOP, B_KNIGHT, B_ROOK, B_CANNON, B_PAWN = -1, -2, -3, -4, -5, -6, -7
PIECE_NAMES = {
    0: '．', 1: '帅', 2: '仕', 3: '相', 4: '马', 5: '车', 6: '炮', 7: '兵',
    -1: '将', -2: '士', -3: '象', -4: '马', -5: '车', -6: '炮', -7: '卒'
}
PIECE_VALUES = np.array([0, 0, 20, 20, 40, 90, 45, 10], dtype=np.int32)
ROWS = 10
COLS = 9
ACTION_SPACE = 90 * 90

# the reference exposes this flag (game.py:31); here it means "native engine in use", always True
_USE_CYTHON = True

_engine = None


def engine(device: int = 0) -> "xq_native.Engine":
    """Process-wide engine context (one per GPU/process, like the C ABI requires)."""
    global _engine
    if _engine is None:
        _engine = xq_native.Engine(device)
    return _engine


def encode_action(from_row: int, from_col: int, to_row: int, to_col: int) -> int:
    return (from_row * COLS + from_col) * 90 + (to_row * COLS + to_col)      # game.py:112-114


def decode_action(action: int) -> Tuple[int, int, int, int]:
    f, t = divmod(int(action), 90)                                            # game.py:117-121
    return f // COLS, f % COLS, t // COLS, t % COLS


_START = np.zeros((ROWS, COLS), np.int8)
_START[0] = [5, 4, 3, 2, 1, 2, 3, 4, 5]
_START[9] = [-5, -4, -3, -2, -1, -2, -3, -4, -5]
_START[2, [1, 7]] = 6
_START[7, [1, 7]] = -6
_START[3, ::2] = 7
_START[6, ::2] = -7


class XiangqiGame:
    """State object of the reference (game.py:124); rule queries run on the GPU."""

    __slots__ = ['board', 'current_player', 'move_count', 'history', 'no_capture_count',
                 '_legal_moves_cache', '_check_cache']

    def __init__(self):
        self.board = np.zeros((ROWS, COLS), dtype=np.int8)
        self._init_board()
        self.current_player = 1
        self.move_count = 0
        self.history = []
        self.no_capture_count = 0
        self._legal_moves_cache = None
        self._check_cache = None

    def _init_board(self):
        """Start position (game.py:139-159): red on rows 0-4, black on rows 5-9."""
        self.board[:] = _START

    def clone(self) -> 'XiangqiGame':
        g = XiangqiGame.__new__(XiangqiGame)
        g.board = self.board.copy()
        g.current_player = self.current_player
        g.move_count = self.move_count
        g.history = self.history.copy()
        g.no_capture_count = self.no_capture_count
        g._legal_moves_cache = None
        g._check_cache = None
        return g

    # ---- rule queries (GPU) -------------------------------------------------------------
    def _query(self):
        b = np.ascontiguousarray(self.board, np.int8).reshape(1, 90)
        acts, n, chk, _ = engine().movegen_host(b, np.array([self.current_player], np.int8))
        k = int(n[0])
        self._legal_moves_cache = [decode_action(a) for a in acts[0, :k].tolist()]
        self._check_cache = bool(chk[0])

    def get_legal_moves(self) -> List[Tuple[int, int, int, int]]:
        if self._legal_moves_cache is None:
            self._query()
        return self._legal_moves_cache

    def get_legal_actions(self) -> List[int]:
        return [encode_action(*m) for m in self.get_legal_moves()]

    @staticmethod
    def _is_attacked(board: np.ndarray, kr: int, kc: int, by_player: int) -> bool:
        out = engine().is_attacked_host(np.ascontiguousarray(board, np.int8).reshape(1, 90),
                                        np.array([kr * COLS + kc], np.uint8), np.array([by_player], np.int8))
        return bool(out[0])

    def _find_king_pos(self, player: int, board: np.ndarray) -> Optional[Tuple[int, int]]:
        rows = range(0, 3) if player == 1 else range(7, 10)        # palace-only, game.py:426-439
        for r in rows:
            for c in range(3, 6):
                if board[r, c] == player:
                    return (r, c)
        return None

    def _find_king(self, player: int) -> Optional[Tuple[int, int]]:
        return self._find_king_pos(player, self.board)

    def _is_in_check(self, player: int, board: Optional[np.ndarray] = None) -> bool:
        if board is None:
            board = self.board
        k = self._find_king_pos(player, board)
        if k is None:
            return True                                            # game_core.pyx:552-554
        return self._is_attacked(board, k[0], k[1], -player)

    def _is_move_legal(self, fr: int, fc: int, tr: int, tc: int, player: int) -> bool:
        """game.py:441-490: after the move `player`'s king must stand in its palace, must not face the other
        king on an open file and must not be attacked.  One GPU query on the moved board: the attack test
        counts the enemy king as a rook (game.py:176-200), which is the facing-kings test."""
        out = engine().move_is_legal_host(self.board, np.array([fr * 9 + fc], np.uint8), np.array([tr * 9 + tc], np.uint8),
                                          np.array([player], np.int8))
        return bool(out[0])

    @staticmethod
    def _kings_facing_fast(board: np.ndarray) -> bool:
        r_pos = np.argwhere(board == R_KING)
        b_pos = np.argwhere(board == B_KING)
        if len(r_pos) == 0 or len(b_pos) == 0:
            return False
        (rr, rc), (br, bc) = r_pos[0], b_pos[0]
        if rc != bc:
            return False
        lo, hi = min(rr, br) + 1, max(rr, br)
        return bool(np.all(board[lo:hi, rc] == EMPTY))

    def _kings_facing(self, board: np.ndarray) -> bool:
        return self._kings_facing_fast(board)

    # ---- state update (game.py:528-550) ---------------------------------------------------
    def make_move(self, from_row: int, from_col: int, to_row: int, to_col: int) -> bool:
        captured = self.board[to_row, to_col]
        self.history.append(self.board.tobytes())
        self.board[to_row, to_col] = self.board[from_row, from_col]
        self.board[from_row, from_col] = EMPTY
        self.no_capture_count = 0 if captured != EMPTY else self.no_capture_count + 1
        self.current_player = -self.current_player
        self.move_count += 1
        self._legal_moves_cache = None
        self._check_cache = None
        return True

    def make_action(self, action: int) -> bool:
        return self.make_move(*decode_action(action))

    def get_material_score(self, player: int) -> int:
        b = self.board
        pieces = b[b > 0] if player == 1 else -b[b < 0]
        return int(PIECE_VALUES[pieces].sum())

    def is_game_over(self) -> Tuple[bool, Optional[int]]:
        """game.py:565-616, same rule order."""
        if self._find_king_pos(1, self.board) is None:
            return True, -1
        if self._find_king_pos(-1, self.board) is None:
            return True, 1
        if len(self.get_legal_moves()) == 0:
            return True, -self.current_player
        if self.no_capture_count >= 120:
            return True, 0
        if self.move_count >= 200:
            diff = self.get_material_score(1) - self.get_material_score(-1)
            return True, (1 if diff > 30 else (-1 if diff < -30 else 0))
        if len(self.history) >= 6:
            cur = self.board.tobytes()
            if sum(1 for h in self.history[-12:] if h == cur) >= 3:
                return True, 0
        return False, None

    def get_state_for_nn(self) -> np.ndarray:
        """15 planes (game.py:618-640): own K,A,B,N,R,C,P / other side / red-to-move."""
        own = self.board * self.current_player
        f = np.zeros((15, ROWS, COLS), dtype=np.float32)
        for k in range(1, 8):
            f[k - 1] = own == k
            f[6 + k] = own == -k
        if self.current_player == 1:
            f[14] = 1.0
        return f

    def get_canonical_board(self) -> np.ndarray:
        if self.current_player == 1:
            return self.board.copy()
        return -np.flip(self.board, axis=0).copy()

    def display(self):
        print("\n  ０ １ ２ ３ ４ ５ ６ ７ ８")
        for r in range(ROWS - 1, -1, -1):
            print(f"{r} " + " ".join(PIECE_NAMES[int(p)] for p in self.board[r]))
            if r == 5:
                print("  ＝＝＝＝＝楚河汉界＝＝＝＝＝")
        print(f"  当前: {'红方' if self.current_player == 1 else '黑方'}  步数: {self.move_count}")
