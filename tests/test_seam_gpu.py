"""The native seam of the reference (game.py:30-47: the five cy_* functions of cython_engine/game_core.pyx:493-569)
answered by the C ABI: batched entry points against the oracle, and the reference's OWN game.py + tests running on
the ctypes shim xiangqi-alphazero_b200/seam/game_core.py."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def eng():
    import game
    return game.engine(0)


def _odd_boards(rng, n):
    """Boards no game reaches: kings missing, outside the palace, doubled; random piece soup."""
    b = np.zeros((n, 90), np.int8)
    for i in range(n):
        k = rng.integers(4, 30)
        sq = rng.choice(90, size=k, replace=False)
        b[i, sq] = rng.integers(-7, 8, size=k)
    return b


def test_find_king_and_has_legal_moves_match_the_oracle(eng, oracle):
    import torch
    boards, sides = oracle.random_playout_positions(11, 60_000)
    rng = np.random.default_rng(5)
    odd = _odd_boards(rng, 6000)
    boards = np.concatenate([boards, odd])
    sides = np.concatenate([sides, rng.choice(np.array([1, -1], np.int8), size=len(odd))]).astype(np.int8)
    db, ds = torch.from_numpy(boards).to(eng.dev), torch.from_numpy(sides).to(eng.dev)
    for side_sel in (ds, -ds):                                   # the side to move and the other side (cy_find_king(board, -player))
        got = eng.find_king(db, side_sel).cpu().numpy()
        s = side_sel.cpu().numpy()
        L = oracle.lib()
        want = np.array([L.xqo_find_king(oracle._p(boards[i], oracle.C.c_int8), int(s[i])) for i in range(len(s))], np.int64)
        assert np.array_equal(got.astype(np.int64), want)
    has = eng.has_legal_moves(db, ds).cpu().numpy()
    _, n, _, _ = oracle.movegen_batch(boards, sides)
    assert np.array_equal(has != 0, n > 0)
    assert (has == 0).sum() > 0                                   # the sample does contain mates / stalemates / kingless boards
    # host-pointer forms
    assert np.array_equal(eng.find_king_host(boards[:999], sides[:999]), eng.find_king(db[:999], ds[:999]).cpu().numpy())
    assert np.array_equal(eng.has_legal_moves_host(boards[-999:], sides[-999:]), has[-999:])


def test_move_is_legal_matches_the_oracle_on_any_move(eng, oracle):
    """_is_move_legal (game_core.pyx:209-252) called directly through the C ABI: the legal moves of a position pass, and so
    does exactly what the oracle passes among arbitrary (from, to) pairs -- the reference makes no pseudo-legality check."""
    import torch
    boards, sides = oracle.random_playout_positions(17, 4000)
    rng = np.random.default_rng(3)
    L = oracle.lib()
    # (a) arbitrary moves of own pieces, empty squares and enemy pieces alike
    frm = rng.integers(0, 90, size=len(sides)).astype(np.uint8)
    to = rng.integers(0, 90, size=len(sides)).astype(np.uint8)
    got = eng.move_is_legal(torch.from_numpy(boards).to(eng.dev), torch.from_numpy(frm).to(eng.dev), torch.from_numpy(to).to(eng.dev),
                            torch.from_numpy(sides).to(eng.dev)).cpu().numpy()
    want = np.array([L.xqo_move_is_legal(oracle._p(boards[i], oracle.C.c_int8), int(frm[i]), int(to[i]), int(sides[i]))
                     for i in range(len(sides))], np.uint8)
    assert np.array_equal(got, want) and 0 < want.sum() < len(want)
    # (b) every generated legal move is legal, through the host-pointer form
    acts, n, _, _ = oracle.movegen_batch(boards[:300], sides[:300])
    bb, ff, tt, ss = [], [], [], []
    for i in range(300):
        for a in acts[i, :n[i]]:
            bb.append(boards[i]); ff.append(a // 90); tt.append(a % 90); ss.append(sides[i])
    ok = eng.move_is_legal_host(np.array(bb, np.int8), np.array(ff, np.uint8), np.array(tt, np.uint8), np.array(ss, np.int8))
    assert ok.all()
    # (c) the drop-in XiangqiGame._is_move_legal answers through it
    import game
    g = game.XiangqiGame()
    assert g._is_move_legal(0, 0, 1, 0, 1) and not g._is_move_legal(0, 4, 5, 4, 1)     # rook step; king leaving its palace


def test_packed_planes_are_the_float_planes(eng, oracle):
    import torch
    import xq_native
    boards, sides = oracle.random_playout_positions(13, 20_001)
    _, _, _, planes = oracle.movegen_batch(boards, sides, want_planes=True)
    bits = eng.planes_bits(torch.from_numpy(boards).to(eng.dev), torch.from_numpy(sides).to(eng.dev)).cpu().numpy().view(np.uint32)
    assert np.array_equal(xq_native.unpack_planes(bits), planes)
    a, n, chk, pk = eng.movegen_host(boards, sides, planes="packed")
    a2, n2, chk2, pf = eng.movegen_host(boards, sides, planes=True)
    assert np.array_equal(a, a2) and np.array_equal(n, n2) and np.array_equal(chk, chk2)
    assert np.array_equal(xq_native.unpack_planes(pk), pf) and np.array_equal(pf, planes)
    assert pk.nbytes * 30 < pf.nbytes                             # 176 vs 5400 bytes per position across PCIe


def test_the_references_own_game_py_and_tests_run_on_the_seam():
    ref = os.path.join(ROOT, "baseline", "_ref", "training", "game.py")
    if not os.path.exists(ref):
        pytest.skip("baseline/_ref/training not present (built where the reference tree exists: make -C oracle baseline)")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "seam_reference_script.py")], capture_output=True, text=True,
                       timeout=900, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.startswith("{")][-1])
    assert d["game_file"] == "baseline/_ref/training/game.py" and d["use_cython"] and d["bound_to_shim"]
    assert d["initial_moves"] == 44 and d["six_ply_moves"] == 42
    assert d["kings"] == [[0, 4], [9, 4]] and d["initial_in_check"] == [False, False] and d["dtype_error"]
    assert d["test_v3_specific_positions"] and d["test_v3_correctness"], d["log_tail"]
    assert d["test_cython_correctness_python_engine_vs_seam"], d["log_tail"]
