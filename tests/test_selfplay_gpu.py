"""Device-resident self-play loop (xq_selfplay_*): games are legal by the oracle's rules, end
exactly where the reference's is_game_over / resign rule ends them, samples follow the
reference's record semantics (parallel_selfplay.py:42-151)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class Cfg:
    num_simulations = 24
    c_puct = 1.5
    temperature_threshold = 20
    max_game_length = 300
    random_opening_moves = 0
    enable_resign = False
    resign_threshold = -0.9
    resign_check_steps = 5
    num_games_per_iter = 24


@pytest.fixture(scope="module")
def eng():
    import game
    return game.engine(0)


@pytest.fixture(scope="module")
def net_model():
    import torch
    import model as M
    torch.manual_seed(3)
    return M.XiangqiNet(128, 1).eval()


def play(eng, model, cfg, slots, games, seed=1):
    from selfplay_engine import SelfPlayEngine, decode_samples
    sp = SelfPlayEngine(eng, model, n_slots=slots, max_games=games)
    sp.reset()
    c = sp.play_games(SelfPlayEngine.make_config(cfg, games, seed=seed), chunk=16)
    raw, winner, plies = sp.fetch(0, c["samples"])
    return sp, c, decode_samples(raw), winner, plies


def test_games_are_legal_and_end_by_the_rules(eng, oracle, net_model):
    cfg = Cfg()
    sp, c, dec, winner, plies = play(eng, net_model, cfg, slots=16, games=24)
    assert c["error"] == 0 and c["dropped"] == 0
    assert c["started"] == 24 and c["finished"] == 24
    assert c["red_wins"] + c["black_wins"] + c["draws"] == 24
    assert c["sims"] == c["samples"] * cfg.num_simulations     # every recorded ply ran all its simulations
    total_plies = 0
    for uid in range(24):
        idx = np.nonzero(dec["uid"] == uid)[0]
        idx = idx[np.argsort(dec["ply"][idx])]
        assert len(idx) == plies[uid] and list(dec["ply"][idx]) == list(range(len(idx)))
        og = oracle.OracleGame()                                 # random_opening_moves = 0: full history known
        for i in idx:
            assert not og.is_game_over()[0]
            assert np.array_equal(dec["board"][i], og.board.reshape(90)) and dec["side"][i] == og.current_player
            legal = og.get_legal_actions()
            n = int(dec["n"][i])
            assert dec["actions"][i, :n].tolist() == legal.tolist()          # root children = ordered legal moves
            p = dec["probs"][i, :n]
            assert abs(float(p.sum()) - 1.0) < 1e-5 and (p >= 0).all()
            a = int(dec["played"][i])
            assert a in legal.tolist() and p[legal.tolist().index(a)] > 0    # only visited moves are played
            og.make_action(a)
        done, w = og.is_game_over()
        assert done and w == winner[uid], (uid, done, w, winner[uid])        # same terminal rule, same winner
        total_plies += len(idx)
    assert c["plies_finished"] == total_plies == c["samples"]


def test_temperature_schedule_in_records(eng, net_model):
    """T = 1 below temperature_threshold plies: probs = N/S (multiples of 1/S); T = 0.3 after:
    probs are N^(1/0.3) normalised, i.e. sharper (parallel_selfplay.py:92, mcts.py:201-203)."""
    cfg = Cfg()
    cfg.temperature_threshold = 4
    sp, c, dec, winner, plies = play(eng, net_model, cfg, slots=8, games=8, seed=5)
    S = cfg.num_simulations
    early = dec["probs"][dec["ply"] < 4]
    assert np.allclose(early * S, np.round(early * S), atol=1e-4)
    late = dec["probs"][dec["ply"] >= 4]
    cnt_like = late ** 0.3
    cnt_like = cnt_like / cnt_like.sum(axis=1, keepdims=True) * S
    assert np.allclose(cnt_like, np.round(cnt_like), atol=2e-3)


def test_resign_rule(eng, net_model):
    """enable_resign with a threshold every value is below: the probe starts once 11 samples exist
    and the game is resigned after resign_check_steps consecutive probes, by the side to move
    (parallel_selfplay.py:110-121)."""
    cfg = Cfg()
    cfg.enable_resign = True
    cfg.resign_threshold = 2.0
    cfg.resign_check_steps = 5
    sp, c, dec, winner, plies = play(eng, net_model, cfg, slots=8, games=12, seed=2)
    assert c["finished"] == 12
    assert (plies[:12] == 15).all()
    for uid in range(12):
        last = np.nonzero((dec["uid"] == uid) & (dec["ply"] == 14))[0]
        assert len(last) == 1
        # after ply 14 the side to move is the opponent of sample 14's side; that side resigns
        assert winner[uid] == dec["side"][last[0]]


def test_random_opening_and_slot_refill(eng, oracle, net_model):
    cfg = Cfg()
    cfg.random_opening_moves = 6
    cfg.num_simulations = 8
    sp, c, dec, winner, plies = play(eng, net_model, cfg, slots=4, games=10, seed=9)
    assert c["finished"] == 10 and c["started"] == 10
    first = [int(dec["ply"][dec["uid"] == u].min()) for u in range(10)]
    assert min(first) >= 0 and max(first) <= 6 and len(set(first)) > 1       # k ~ U{0..6} opening plies
    # every recorded position is reachable: legal list matches the oracle
    ea, en, _, _ = oracle.movegen_batch(dec["board"], dec["side"])
    assert np.array_equal(en, dec["n"]) and np.array_equal(ea, dec["actions"])


def test_parallel_self_play_dropin(eng, net_model):
    """The reference entry point and its return contract (parallel_selfplay.py:264-334)."""
    import parallel_selfplay as ps
    cfg = Cfg()
    cfg.num_games_per_iter = 6
    cfg.num_simulations = 8
    data, stats = ps.parallel_self_play(net_model, cfg, num_workers=3, use_gpu_server=True, gpu_device='cuda')
    assert set(stats) >= {'games', 'red_wins', 'black_wins', 'draws', 'avg_steps', 'new_samples', 'total_time',
                          'num_workers', 'mode'}
    assert stats['games'] == 6 and stats['new_samples'] == len(data) and len(data) % 2 == 0
    s, p, z = data[0]
    sm, pm, zm = data[1]
    assert s.shape == (15, 10, 9) and s.dtype == np.float32 and p.shape == (8100,) and z in (-1.0, 0.0, 1.0)
    assert abs(p.sum() - 1.0) < 1e-5 and abs(pm.sum() - 1.0) < 1e-5 and z == zm
    assert np.array_equal(sm, np.flip(s, axis=2))                            # mirrored copy follows the original
    # mirrored policy: (fr,fc,tr,tc) -> (fr,8-fc,tr,8-tc)
    from game import decode_action, encode_action
    for a in np.nonzero(p)[0]:
        fr, fc, tr, tc = decode_action(int(a))
        assert pm[encode_action(fr, 8 - fc, tr, 8 - tc)] == p[a]
    assert ps._augment_data([(s, p, z)])[1][1].tolist() == pm.tolist()


def test_same_seed_same_games(eng, net_model):
    """Counter-based RNG + deterministic kernels: two runs with one seed produce the same records
    (the append order of records and the node numbering may differ, the content may not)."""
    cfg = Cfg()
    cfg.random_opening_moves = 4
    cfg.num_simulations = 16
    runs = []
    for _ in range(2):
        sp, c, dec, winner, plies = play(eng, net_model, cfg, slots=8, games=12, seed=77)
        order = np.lexsort((dec["ply"], dec["uid"]))
        runs.append((c["samples"], winner[:12].copy(), plies[:12].copy(),
                     {k: dec[k][order] for k in ("board", "side", "n", "actions", "probs", "played", "uid", "ply")}))
    a, b = runs
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for k in a[3]:
        assert np.array_equal(a[3][k], b[3][k]), k
    sp2, c2, dec2, w2, p2 = play(eng, net_model, cfg, slots=8, games=12, seed=78)
    assert not np.array_equal(np.sort(dec2["played"]), np.sort(a[3]["played"])) or c2["samples"] != a[0]


def test_live_bound_changes_the_kernels_not_the_games(eng, net_model):
    """xq_selfplay_set_live_bound: with a correct upper bound on the live games the forwards are launched for that many
    boards (small-batch layer variants: 64-channel single-CTA tower items, 64-column FC tiles) and the games are the same
    records bit for bit; a bound that is too small is reported through the error counter, never silently."""
    import xq_native
    from selfplay_engine import SelfPlayEngine, decode_samples
    cfg = Cfg()
    cfg.num_simulations = 12
    cfg.max_game_length = 24
    cfg.random_opening_moves = 2
    slots, games = 320, 40                     # a plan of 320 boards (CTA-pair tower, 224-column FC); only 40 games ever play

    def run(bound):
        sp = SelfPlayEngine(eng, net_model, n_slots=slots, max_games=games)
        sp.reset()
        scfg = SelfPlayEngine.make_config(cfg, games, seed=5)
        if bound:
            eng._check(eng.L.xq_selfplay_set_live_bound(eng.h, bound))
        sp.play(scfg, 6)
        c = sp.counters()
        raw, winner, plies = sp.fetch(0, c["samples"])
        dec = decode_samples(raw)
        order = np.lexsort((dec["ply"], dec["uid"]))
        return c, {k: dec[k][order] for k in ("board", "side", "n", "actions", "probs", "played", "uid", "ply")}

    c0, d0 = run(0)
    c1, d1 = run(games)                         # 40 games: every layer on its small-batch variant
    assert c0["error"] == 0 and c1["error"] == 0 and c0["samples"] == c1["samples"] > 0
    for k in d0:
        assert np.array_equal(d0[k], d1[k]), k
    c2, _ = run(games // 4)                     # wrong: more games are alive than the bound says
    assert c2["error"] & 4
    sp = SelfPlayEngine(eng, net_model, n_slots=slots, max_games=games)
    sp.reset()                                  # reset clears the hint and the error bits
    assert sp.counters()["error"] == 0


def test_graph_replay_dependent_launches_and_small_variants_do_not_change_the_games():
    """The lockstep step replayed from a CUDA graph, programmatic dependent launches and the small-batch layer variants are
    scheduling choices: a seeded self-play run + arena produce the same records and moves with each of them switched off."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    digests = {}
    for name, env in (("default", {}), ("no_graph", {"XQ_SP_GRAPH": "0"}), ("no_pdl", {"XQ_NET_PDL": "0"}),
                      ("no_small", {"XQ_NET_SMALL": "0"})):
        e = dict(os.environ)
        e.update(env)
        out = subprocess.run([sys.executable, os.path.join(here, "sp_hash_script.py")], capture_output=True, text=True, timeout=600, env=e)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
        line = [l for l in out.stdout.splitlines() if l.startswith("SP_DIGEST")][-1].split()
        assert int(line[3]) == 0 and int(line[1]) > 0
        digests[name] = line[1:]
    assert digests["default"] == digests["no_graph"] == digests["no_pdl"] == digests["no_small"], digests
