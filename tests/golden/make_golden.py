"""Generate tests/golden/*.npz|json by RUNNING THE REFERENCE (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/training/{game,mcts}.py unmodified, with the reference's Cython
engine (game_core.pyx compiled as-is by oracle/Makefile into oracle/_ref/) on sys.path so
that game._USE_CYTHON is True -- i.e. the goldens are outputs of the Cython engine, the
oracle BASELINE.json names.  The fixtures are committed; this script is not run on the GPU
box (there is no /root/reference there).

Fixtures
  rules_golden.npz     random legal playouts: per ply board, side, ordered legal actions,
                       in-check, packed feature planes, is_game_over() result, move played
  attacked_golden.npz  is_attacked on all 90 squares x both sides for sampled positions
                       (Python static method and cy_is_attacked agree) + hand-built
                       known-answer positions of test_v3.py:123-197
  mcts_golden.json     MCTS.search visit counts under deterministic evaluators
                       (tests/evaluators.py), with and without injected root noise
"""
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/training"
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))  # game_core (Cython, built from the reference .pyx)

import game as refgame  # noqa: E402
import mcts as refmcts  # noqa: E402
from game import XiangqiGame  # noqa: E402
import game_core  # noqa: E402
import evaluators  # noqa: E402

assert refgame._USE_CYTHON, "reference Cython engine not importable: run `make -C oracle ref`"
assert refgame.__file__.startswith(REF) and refmcts.__file__.startswith(REF)


def rules_fixture(num_games=40, seed=20261018):
    random.seed(seed)
    rec = {k: [] for k in ("board", "side", "n", "actions", "in_check", "planes", "done", "winner",
                           "game", "ply", "played", "no_capture", "move_count")}
    for gi in range(num_games):
        g = XiangqiGame()
        ply = 0
        while True:
            done, winner = g.is_game_over()
            acts = g.get_legal_actions()
            padded = np.full(128, -1, np.int16)
            padded[:len(acts)] = acts
            rec["board"].append(g.board.reshape(90).copy())
            rec["side"].append(g.current_player)
            rec["n"].append(len(acts))
            rec["actions"].append(padded)
            rec["in_check"].append(bool(g._is_in_check(g.current_player)))
            rec["planes"].append(np.packbits(g.get_state_for_nn().reshape(-1) > 0.5))
            rec["done"].append(bool(done))
            rec["winner"].append(2 if winner is None else int(winner))
            rec["game"].append(gi)
            rec["ply"].append(ply)
            rec["no_capture"].append(g.no_capture_count)
            rec["move_count"].append(g.move_count)
            if done:
                rec["played"].append(-1)
                break
            a = random.choice(acts)
            rec["played"].append(a)
            g.make_action(a)
            ply += 1
    out = dict(
        board=np.array(rec["board"], np.int8), side=np.array(rec["side"], np.int8),
        n=np.array(rec["n"], np.uint8), actions=np.array(rec["actions"], np.int16),
        in_check=np.array(rec["in_check"], np.uint8), planes=np.array(rec["planes"], np.uint8),
        done=np.array(rec["done"], np.uint8), winner=np.array(rec["winner"], np.int8),
        game=np.array(rec["game"], np.int32), ply=np.array(rec["ply"], np.int32),
        played=np.array(rec["played"], np.int16), no_capture=np.array(rec["no_capture"], np.int32),
        move_count=np.array(rec["move_count"], np.int32),
    )
    np.savez_compressed(os.path.join(HERE, "rules_golden.npz"), **out)
    print("rules_golden:", len(out["side"]), "positions,", num_games, "games, mean moves",
          out["n"].mean(), "max", out["n"].max(), "in-check", out["in_check"].mean(),
          "winners", {w: int((out["winner"][out["done"] > 0] == w).sum()) for w in (1, -1, 0)})
    return out


def shuffle_fixture(seed=7):
    """Short games with repetition: both sides shuffle a rook back and forth (repetition rule,
    game.py:606-614) and a long no-capture shuffle is cut by the 120-ply rule (game.py:591)."""
    seqs = []
    # rook a1-a2-a1 / a10-a9-a10 ...
    r_fwd, r_back = (0 * 9 + 0) * 90 + (1 * 9 + 0), (1 * 9 + 0) * 90 + (0 * 9 + 0)
    b_fwd, b_back = (9 * 9 + 0) * 90 + (8 * 9 + 0), (8 * 9 + 0) * 90 + (9 * 9 + 0)
    seqs.append([r_fwd, b_fwd, r_back, b_back] * 6)
    # knights out and back
    n1, n1b = (0 * 9 + 1) * 90 + (2 * 9 + 2), (2 * 9 + 2) * 90 + (0 * 9 + 1)
    m1, m1b = (9 * 9 + 1) * 90 + (7 * 9 + 2), (7 * 9 + 2) * 90 + (9 * 9 + 1)
    seqs.append([n1, m1, n1b, m1b, r_fwd, b_fwd, r_back, b_back, n1, m1, n1b, m1b] * 3)
    rows = []
    for si, seq in enumerate(seqs):
        g = XiangqiGame()
        for ply, a in enumerate(seq + [-1]):
            done, winner = g.is_game_over()
            rows.append((si, ply, int(done), 2 if winner is None else int(winner), a))
            if done or a < 0:
                break
            g.make_action(a)
    return np.array(rows, np.int32)


def attacked_fixture(rules, stride=97):
    idx = np.arange(0, len(rules["side"]), stride)
    boards, out = [], []
    for i in idx:
        b = rules["board"][i].reshape(10, 9).copy()
        row = np.zeros((2, 90), np.uint8)
        for si, by in enumerate((1, -1)):
            for sq in range(90):
                py = XiangqiGame._is_attacked(b, sq // 9, sq % 9, by)
                cy = game_core.cy_is_attacked(b, sq // 9, sq % 9, by)
                assert bool(py) == bool(cy)
                row[si, sq] = bool(cy)
        boards.append(b.reshape(90))
        out.append(row)
    # hand-built positions of the reference's own tests (test_v3.py:123-197)
    kb = []
    b = np.zeros((10, 9), np.int8); b[0, 4] = 1; b[9, 4] = -1; b[5, 4] = -5
    kb.append((b, 0, 4, -1, True))            # test_v3.py:139-151
    b = np.zeros((10, 9), np.int8); b[0, 4] = 1; b[2, 3] = -4
    kb.append((b, 0, 4, -1, True))            # test_v3.py:154-166
    b = b.copy(); b[1, 3] = 7
    kb.append((b, 0, 4, -1, False))           # test_v3.py:169-181 (horse leg)
    b = np.zeros((10, 9), np.int8); b[0, 4] = 1; b[9, 4] = -1; b[5, 4] = 7; b[8, 4] = -6
    kb.append((b, 0, 4, -1, True))            # test_v3.py:184-197 (cannon + screen)
    for b, kr, kc, by, want in kb:
        assert bool(game_core.cy_is_attacked(b, kr, kc, by)) == want
        assert bool(XiangqiGame._is_attacked(b, kr, kc, by)) == want
    # synthetic unreachable boards: the Cython engine is the oracle where Python differs
    # 
    syn = []
    b = np.zeros((10, 9), np.int8); b[0, 3] = 1; b[0, 4] = 2; b[9, 5] = -1
    syn.append((b, 1))
    b = np.zeros((10, 9), np.int8); b[0, 4] = 1; b[9, 4] = -1      # bare kings facing
    syn.append((b, 1)); syn.append((b, -1))
    b = np.zeros((10, 9), np.int8); b[9, 4] = -1; b[4, 4] = 7       # red king missing
    syn.append((b, 1)); syn.append((b, -1))
    b = np.zeros((10, 9), np.int8); b[1, 4] = 1; b[8, 4] = -1; b[4, 4] = 6; b[4, 0] = -5; b[6, 4] = -7
    syn.append((b, 1)); syn.append((b, -1))
    syn_boards, syn_sides, syn_n, syn_acts, syn_chk = [], [], [], [], []
    for b, side in syn:
        mv = game_core.cy_generate_legal_moves(b, side)
        acts = np.full(128, -1, np.int16)
        acts[:len(mv)] = [(fr * 9 + fc) * 90 + tr * 9 + tc for fr, fc, tr, tc in mv]
        syn_boards.append(b.reshape(90)); syn_sides.append(side); syn_n.append(len(mv))
        syn_acts.append(acts); syn_chk.append(bool(game_core.cy_is_in_check(b, side)))
    np.savez_compressed(
        os.path.join(HERE, "attacked_golden.npz"),
        board=np.array(boards, np.int8), attacked=np.array(out, np.uint8),
        known_board=np.array([k[0].reshape(90) for k in kb], np.int8),
        known_query=np.array([[k[1], k[2], k[3], int(k[4])] for k in kb], np.int32),
        syn_board=np.array(syn_boards, np.int8), syn_side=np.array(syn_sides, np.int8),
        syn_n=np.array(syn_n, np.uint8), syn_actions=np.array(syn_acts, np.int16),
        syn_in_check=np.array(syn_chk, np.uint8), shuffle=shuffle_fixture(),
    )
    print("attacked_golden:", len(idx), "positions x 180 queries, attacked fraction",
          float(np.mean(out)), "| synthetic boards", len(syn), "n=", syn_n)


def mcts_fixture(rules):
    """Visit counts of the reference MCTS.search (mcts.py:94-155) under deterministic evaluators."""
    cases = []
    # positions = prefixes of the golden playouts (so every engine can rebuild the history)
    def prefix(gi, ply):
        sel = (rules["game"] == gi) & (rules["ply"] < ply)
        return [int(a) for a in rules["played"][sel]]

    last_ply = {gi: int(rules["ply"][rules["game"] == gi].max()) for gi in range(6)}
    plan = [
        ("uniform", 0, 0, 200, False), ("uniform", 0, 0, 800, False),
        ("hash", 0, 0, 800, False), ("ratio", 0, 0, 800, False),
        ("hash", 1, 12, 400, False), ("ratio", 1, 37, 400, False),
        ("hash", 2, 80, 800, False), ("ratio", 3, 120, 300, False),
        ("hash", 0, 0, 400, True), ("ratio", 2, 45, 400, True), ("uniform", 4, 9, 128, True),
    ]
    # late positions: in-tree terminals from the 200-ply rule / mates (mcts.py:136-140)
    for gi in range(6):
        lp = last_ply[gi]
        if lp >= 6:
            plan.append(("hash", gi, max(0, lp - 3), 300, False))
            plan.append(("ratio", gi, max(0, lp - 2), 200, gi % 2 == 0))
    rs = np.random.RandomState(1234)
    for ev_name, gi, ply, sims, noisy in plan:
        moves = prefix(gi, ply)
        g = XiangqiGame()
        for a in moves:
            g.make_action(a)
        done, _ = g.is_game_over()
        if done:
            continue
        n_legal = len(g.get_legal_actions())
        noise = None
        if noisy:
            noise = rs.dirichlet([0.3] * n_legal)
            orig = np.random.dirichlet
            np.random.dirichlet = lambda alpha, _n=noise: _n.copy()
        try:
            ev = evaluators.PlaneEvaluator(ev_name)
            m = refmcts.MCTS(ev, num_simulations=sims, c_puct=1.5)
            # temperature 1.0 => probs = visits / sum(visits)
            probs = m.search(g, temperature=1.0, add_noise=noisy)
        finally:
            if noisy:
                np.random.dirichlet = orig
        acts = g.get_legal_actions()
        visits = [int(round(probs[a] * sims)) for a in acts]
        assert sum(visits) == sims, (sum(visits), sims)
        t0 = m.search(g, temperature=0, add_noise=False) if not noisy else None
        cases.append(dict(evaluator=ev_name, game=gi, ply=ply, moves=moves, sims=sims,
                          noise=None if noise is None else [float(x).hex() for x in noise],
                          actions=[int(a) for a in acts], visits=visits, evals=ev.calls,
                          argmax_t0=None if t0 is None else int(np.argmax(t0))))
        print(f"mcts {ev_name:8s} game {gi} ply {ply:3d} sims {sims:4d} noise {noisy!s:5s} "
              f"legal {n_legal:3d} max-visits {max(visits)} evals {ev.calls}")
    with open(os.path.join(HERE, "mcts_golden.json"), "w") as f:
        json.dump(dict(numpy=np.__version__, c_puct=1.5, cases=cases), f)


if __name__ == "__main__":
    rules = rules_fixture()
    attacked_fixture(rules)
    mcts_fixture(rules)
