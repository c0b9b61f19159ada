"""The scalar rules of the thread-per-board movegen kernel (csrc/xq_rules_tpb.h) are plain C++: compiled here with g++
and compared, bit for bit, with the oracle on the reference goldens, random-playout positions and piece-soup boards.
(The kernel itself is checked on the GPU by tests/test_movegen_gpu.py; this test pins its rules logic on the CPU box.)"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "native", "tpb_host.cpp")
HDR = os.path.join(ROOT, "xiangqi-alphazero_b200", "csrc", "xq_rules_tpb.h")


@pytest.fixture(scope="module")
def tpb(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("tpb") / "libtpb_host.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", out, SRC], check=True)
    L = ctypes.CDLL(out)
    L.xqt_host_movegen_batch.restype = ctypes.c_int
    L.xqt_host_movegen_batch.argtypes = [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 3

    def run(boards, sides):
        boards = np.ascontiguousarray(boards, np.int8).reshape(-1, 90)
        sides = np.ascontiguousarray(sides, np.int8)
        B = len(sides)
        a = np.empty((B, 128), np.int16)
        n = np.empty(B, np.uint8)
        c = np.empty(B, np.uint8)
        rc = L.xqt_host_movegen_batch(boards.ctypes.data, sides.ctypes.data, B, a.ctypes.data, n.ctypes.data, c.ctypes.data)
        assert rc >= 0, f"board {-1 - rc} was not restored"
        return a, n, c, rc
    return run


def test_reference_goldens(tpb, rules_golden, attacked_golden):
    g = rules_golden
    a, n, c, ov = tpb(g["board"], g["side"])
    assert ov == 0 and np.array_equal(n, g["n"]) and np.array_equal(a, g["actions"]) and np.array_equal(c, g["in_check"])
    s = attacked_golden
    a, n, c, ov = tpb(s["syn_board"], s["syn_side"])
    assert np.array_equal(n, s["syn_n"]) and np.array_equal(a, s["syn_actions"]) and np.array_equal(c, s["syn_in_check"])


def test_random_playout_positions(tpb, oracle):
    boards, sides = oracle.random_playout_positions(77, 300_000)
    ea, en, ec, _ = oracle.movegen_batch(boards, sides)
    a, n, c, ov = tpb(boards, sides)
    assert ov == 0 and np.array_equal(n, en) and np.array_equal(c, ec) and np.array_equal(a, ea)
    assert int(en.max()) >= 60 and int(ec.sum()) > 1000          # the sample holds busy and in-check positions


def soup(rs, count, lo, hi, kings=True):
    boards = np.zeros((count, 90), np.int8)
    for i in range(count):
        k = rs.randint(lo, hi)
        sq = rs.choice(90, k, replace=False)
        boards[i, sq] = rs.choice([-7, -6, -5, -4, -3, -2, -1, 1, 2, 3, 4, 5, 6, 7], k)
        if kings and i % 3:
            boards[i, rs.choice([3, 4, 5, 12, 13, 14, 21, 22, 23])] = 1
            boards[i, rs.choice([66, 67, 68, 75, 76, 77, 84, 85, 86])] = -1
    return boards, rs.choice(np.array([1, -1], np.int8), count)


def test_piece_soup_boards(tpb, oracle):
    rs = np.random.RandomState(5)
    for lo, hi in ((2, 28), (20, 60), (1, 6)):
        boards, sides = soup(rs, 20_000, lo, hi)
        ea, en, ec, _ = oracle.movegen_batch(boards, sides)
        a, n, c, _ = tpb(boards, sides)
        assert np.array_equal(n, en) and np.array_equal(c, ec) and np.array_equal(a, ea)


def test_many_knights_many_kings_and_long_lists(tpb, oracle):
    """More than two enemy knights (the generic 8-origin path), several kings per palace, and rook/cannon crowds whose
    pseudo-legal list exceeds one chunk of the per-board scratch (generation resumes after a legality pass)."""
    rs = np.random.RandomState(9)
    boards = np.zeros((6000, 90), np.int8)
    sides = rs.choice(np.array([1, -1], np.int8), len(boards))
    for i in range(len(boards)):
        s = int(sides[i])
        pal = [3, 4, 5, 12, 13, 14, 21, 22, 23] if s == 1 else [66, 67, 68, 75, 76, 77, 84, 85, 86]
        epal = [66, 67, 68, 75, 76, 77, 84, 85, 86] if s == 1 else [3, 4, 5, 12, 13, 14, 21, 22, 23]
        mode = i % 3
        if mode == 0:      # knights everywhere
            sq = rs.choice(90, 14, replace=False)
            boards[i, sq] = rs.choice([-4 * s, -4 * s, 7 * s, -7 * s, 5 * s], 14)
        elif mode == 1:    # several kings
            boards[i, rs.choice(pal, 3, replace=False)] = s
            boards[i, rs.choice(epal, 2, replace=False)] = -s
            sq = rs.choice(90, 8, replace=False)
            boards[i, sq] = rs.choice([5 * s, -5 * s, 6 * s, -6 * s, 4 * s, -4 * s], 8)
        else:              # 8-12 own rooks / cannons on an open board: > 131 pseudo-legal moves
            sq = rs.choice(90, rs.randint(8, 13), replace=False)
            boards[i, sq] = rs.choice([5 * s, 6 * s], len(sq))
        if mode != 1:
            boards[i, rs.choice(pal)] = s
            boards[i, rs.choice(epal)] = -s
    ea, en, ec, _ = oracle.movegen_batch(boards, sides, allow_overflow=True)
    a, n, c, ov = tpb(boards, sides)
    assert np.array_equal(n, en) and np.array_equal(c, ec) and np.array_equal(a, ea)
    assert int(en.max()) == 128 and ov > 0                      # the capped / overflow case is in the sample
