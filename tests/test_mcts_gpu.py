"""K2 parity: GPU-resident MCTS visit counts, bit-exact against the reference goldens
(tests/golden/mcts_golden.json, produced by the reference's mcts.py) and the CPU oracle,
under deterministic evaluators with matched tie-breaking."""
import numpy as np
import pytest

import evaluators

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import xq_native
    e = xq_native.Engine(0)
    yield e
    e.close()


def host_evaluator(name, calls=None):
    """predict() contract on the host: reads the leaf boards back, fills float32 probs + value."""
    fn = evaluators.PREDICT[name]

    def ev(mb):
        t = mb.t
        n = mb.n
        boards = mb.boards[:n].cpu().numpy().reshape(n, 10, 9)
        sides = mb.sides[:n].cpu().numpy()
        probs = np.empty((n, 8100), np.float32)
        vals = np.empty(n, np.float32)
        for i in range(n):
            probs[i], vals[i] = fn(boards[i], int(sides[i]))
        if calls is not None:
            calls.append(n)
        return t.from_numpy(probs).to(mb.e.dev), t.from_numpy(vals).to(mb.e.dev)
    return ev


def game_arrays(games):
    return (np.stack([g.board.reshape(90) for g in games]), np.array([g.current_player for g in games], np.int8),
            np.array([g.move_count for g in games], np.int32), np.array([g.no_capture_count for g in games], np.int32),
            np.stack([g.ring.copy() for g in games]))


def test_reference_golden_visit_counts(eng, oracle, mcts_golden):
    import xq_native
    mb = xq_native.MctsBatch(eng, 1)
    for case in mcts_golden["cases"]:
        og = oracle.OracleGame()
        for a in case["moves"]:
            og.make_action(a)
        mb.set_games(*game_arrays([og]))
        noise = None
        if case["noise"] is not None:
            nz = np.zeros((1, 128), np.float64)
            vals = [float.fromhex(h) for h in case["noise"]]
            nz[0, :len(vals)] = vals
            noise = eng.torch.from_numpy(nz).to(eng.dev)
        acts, vis, n = mb.search(host_evaluator(case["evaluator"]), case["sims"], mcts_golden["c_puct"],
                                 noise=noise, add_noise=noise is not None)
        k = int(n[0])
        assert acts[0, :k].cpu().tolist() == case["actions"]
        assert vis[0, :k].cpu().tolist() == case["visits"], (case["evaluator"], case["game"], case["ply"], case["sims"])
        assert int(vis[0].sum()) == case["sims"]


@pytest.mark.parametrize("name,sims", [("ratio", 160), ("hash", 96)])
def test_batched_search_equals_oracle(eng, oracle, name, sims):
    """64 games at different plies searched in lockstep == 64 independent oracle searches
    (visit counts and float64 total values), including games close to the 200-ply rule."""
    import xq_native
    rs = np.random.RandomState(5)
    games = []
    for i in range(64):
        g = oracle.OracleGame()
        target = [0, 3, 11, 40, 90, 150, 196, 198][i % 8] + (i // 8)
        for _ in range(target):
            done, _ = g.is_game_over()
            if done:
                break
            acts = g.get_legal_actions()
            g.make_action(int(acts[rs.randint(len(acts))]))
        if g.is_game_over()[0]:
            g = oracle.OracleGame()
        games.append(g)
    mb = xq_native.MctsBatch(eng, 64)
    mb.set_games(*game_arrays(games))
    calls = []
    acts, vis, n = mb.search(host_evaluator(name, calls), sims, 1.5)
    _, _, _, w = mb.root_visits(want_w=True)
    st = mb.stats(reset=True)
    assert st["error"] == 0 and st["sims"] == 64 * sims
    terminal = 0
    for i, g in enumerate(games):
        ea, ev, ew, es = oracle.mcts_search(g, sims, 1.5, name)
        k = int(n[i])
        assert acts[i, :k].cpu().tolist() == ea.tolist()
        assert vis[i, :k].cpu().tolist() == ev.tolist(), i
        assert np.array_equal(w[i, :k].cpu().numpy(), ew), i
        terminal += es["terminal_sims"]
    assert st["terminal_sims"] == terminal
    assert len(calls) == sims + 1                      # one evaluator batch per step + the root batch


def test_logits_mode_matches_probs_mode(eng, oracle):
    """policy_kind=2 (float32 logits, softmax over legal entries on the device) gives the same
    search as feeding softmax probabilities when the two are numerically identical by construction."""
    import xq_native
    t = eng.torch
    g = oracle.OracleGame()
    mb = xq_native.MctsBatch(eng, 1)
    mb.set_games(*game_arrays([g]))

    def ev_logits(mb):
        return t.zeros((1, 8100), device=eng.dev), t.zeros((1,), device=eng.dev)      # uniform
    a1, v1, n1 = mb.search(ev_logits, 200, 1.5, kind=xq_native.POLICY_LOGITS_F32)
    assert int(v1.sum()) == 200 and int(n1[0]) == 44 and int(v1.max()) <= 6


def test_device_dirichlet_noise_and_idle_games(eng, oracle):
    import xq_native
    g = oracle.OracleGame()
    games = [g, g.clone(), g.clone(), g.clone()]
    mb = xq_native.MctsBatch(eng, 4)
    b, s, mc, nc, ring = game_arrays(games)
    mb.set_games(b, s, mc, nc, ring, active=np.array([1, 1, 0, 1], np.uint8))
    a, v, n = mb.search(host_evaluator("uniform"), 64, 1.5, add_noise=True, seed=11)
    v = v.cpu().numpy()
    assert v[0].sum() == 64 and v[1].sum() == 64 and v[3].sum() == 64
    assert v[2].sum() == 0 and int(n[2]) == 0          # inactive game idles
    assert not np.array_equal(v[0], v[1])              # per-game noise streams differ
    a2, v2, _ = mb.search(host_evaluator("uniform"), 64, 1.5, add_noise=True, seed=11)
    assert np.array_equal(v2.cpu().numpy(), v)         # same seed -> same search


def test_large_lockstep_batch_equals_oracle(eng, oracle):
    """512 games in one lockstep batch (128 CTAs of the tree kernels, one shared node pool with atomic
    bump allocation): every game's search equals its own independent oracle search."""
    import xq_native
    G = 512
    boards, sides = oracle.random_playout_positions(77, 40000)
    rs = np.random.RandomState(1)
    games = []
    while len(games) < G:
        g = oracle.OracleGame()
        for _ in range(int(rs.randint(0, 120))):
            if g.is_game_over()[0]:
                break
            a = g.get_legal_actions()
            g.make_action(int(a[rs.randint(len(a))]))
        if not g.is_game_over()[0]:
            games.append(g)
    mb = xq_native.MctsBatch(eng, G)
    mb.set_games(*game_arrays(games))
    sims = 24
    acts, vis, n = mb.search(host_evaluator("hash"), sims, 1.5)
    _, _, _, w = mb.root_visits(want_w=True)
    st = mb.stats(reset=True)
    assert st["error"] == 0 and st["sims"] == G * sims
    acts, vis, n, w = acts.cpu().numpy(), vis.cpu().numpy(), n.cpu().numpy(), w.cpu().numpy()
    for i, g in enumerate(games):
        ea, ev, ew, _ = oracle.mcts_search(g, sims, 1.5, "hash")
        k = int(n[i])
        assert acts[i, :k].tolist() == ea.tolist() and vis[i, :k].tolist() == ev.tolist(), i
        assert np.array_equal(w[i, :k], ew), i
