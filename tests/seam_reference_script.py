"""Runs the REFERENCE's own game.py and its own tests (test_v3.py, test_cython.py) on top of the C-ABI seam:
`game_core` = xiangqi-alphazero_b200/seam/game_core.py (ctypes over libxq_b200.so) instead of the Cython module.
Needs baseline/_ref/training (the unmodified mirror, `make -C oracle baseline`) and a GPU.  Prints one JSON line."""
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEAM = os.path.join(ROOT, "xiangqi-alphazero_b200", "seam")
REF = os.path.join(ROOT, "baseline", "_ref", "training")

sys.path[:] = [p for p in sys.path if "xiangqi-alphazero_b200" not in p and os.path.abspath(p or ".") != os.path.join(ROOT, "tests")]
sys.path.insert(0, SEAM)
import game_core                                    # the shim; game.py's "from game_core import ..." now resolves to it
assert os.path.abspath(game_core.__file__).startswith(SEAM), game_core.__file__
pkg = types.ModuleType("cython_engine")
pkg.__path__ = []
pkg.game_core = game_core
sys.modules["cython_engine"] = pkg
sys.modules["cython_engine.game_core"] = game_core  # test_cython.py imports the engine under this name
sys.path.insert(0, REF)
import random
import numpy as np
import game                                         # the reference's game.py, unmodified

out = {"game_file": os.path.relpath(game.__file__, ROOT), "use_cython": bool(game._USE_CYTHON),
       "bound_to_shim": game.cy_generate_legal_moves is game_core.cy_generate_legal_moves}
g = game.XiangqiGame()
out["initial_moves"] = len(g.get_legal_moves())                      # test_v3.py:115-120, test_cython.py:46-57: 44
for mv in [(2, 1, 4, 2), (7, 1, 5, 2), (0, 1, 2, 2), (9, 1, 7, 2), (3, 0, 4, 0), (6, 0, 5, 0)]:   # test_cython.py:62-84
    g.make_move(*mv)
out["six_ply_moves"] = len(g.get_legal_moves())                      # 42
g = game.XiangqiGame()
out["kings"] = [game_core.cy_find_king(g.board, 1), game_core.cy_find_king(g.board, -1)]     # test_cython.py:127-138
out["initial_in_check"] = [game_core.cy_is_in_check(g.board, 1), game_core.cy_is_in_check(g.board, -1)]
try:
    game_core.cy_generate_legal_moves(g.board.astype(np.int32), 1)
    out["dtype_error"] = False
except ValueError:
    out["dtype_error"] = True

random.seed(7)
buf = io.StringIO()
with redirect_stdout(buf):
    import test_v3
    out["test_v3_specific_positions"] = bool(test_v3.test_specific_positions())   # test_v3.py:106-203
    out["test_v3_correctness"] = bool(test_v3.test_correctness())                 # test_v3.py:16-103 (50 random games)
    # the reference's Python-vs-Cython differential test with the roles "reference pure-Python engine" vs "this seam"
    game._USE_CYTHON = False
    import test_cython
    out["test_cython_correctness_python_engine_vs_seam"] = bool(test_cython.test_correctness())   # test_cython.py:35-141
out["log_tail"] = buf.getvalue()[-400:]
print(json.dumps(out))
