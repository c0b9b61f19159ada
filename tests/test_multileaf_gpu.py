"""The self-play search path: leaf compaction (K = 1, must stay the reference's search exactly) and the opt-in
multi-leaf / virtual-loss mode (K > 1: visit sums, legality, determinism), mcts.py:94-155."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class Cfg:
    num_simulations = 24
    c_puct = 1.5
    temperature_threshold = 400          # T = 1 throughout: recorded probabilities are N / S
    max_game_length = 300
    random_opening_moves = 4
    enable_resign = False
    resign_threshold = -0.9
    resign_check_steps = 5


@pytest.fixture(scope="module")
def eng():
    import game
    return game.engine(0)


@pytest.fixture(scope="module")
def net_model():
    import torch
    import model as M
    torch.manual_seed(4)
    return M.XiangqiNet(128, 1).eval()


def run(eng, model, cfg, slots, games, plies, K=1, seed=1, add_noise=False):
    from selfplay_engine import SelfPlayEngine, decode_samples
    sp = SelfPlayEngine(eng, model, n_slots=slots, max_games=games, leaves_per_game=K)
    sp.reset()
    sp.play(SelfPlayEngine.make_config(cfg, games, seed=seed, add_noise=add_noise, leaves_per_game=K), plies)
    c = sp.counters()
    raw, winner, pl = sp.fetch(0, c["samples"])
    return sp, c, decode_samples(raw)


def test_compacted_k1_search_is_the_standalone_search(eng, net_model):
    """K = 1 through the self-play kernels (compacted evaluator rows, batch sized on the device) gives the visit
    counts of the standalone lockstep search (tests/test_mcts_gpu.py pins that one to the reference's mcts.py
    goldens) run on the recorded positions with the same network."""
    import torch
    import xq_native
    import model as M
    cfg = Cfg()
    sp, c, dec = run(eng, net_model, cfg, slots=12, games=12, plies=3)
    assert c["error"] == 0 and c["samples"] == 36 and c["sims"] == 36 * cfg.num_simulations
    S = cfg.num_simulations
    n = len(dec["side"])
    net = M.B200Net(eng, net_model, max_batch=n)
    e2 = xq_native.Engine(0)                      # own context: the standalone search must not disturb the self-play state
    mb = xq_native.MctsBatch(e2, n)
    # the positions were reached by play, so their history (repetition ring) is not known here: none of these short
    # searches can complete a threefold repetition, and move_count only matters at 200
    mb.set_games(dec["board"], dec["side"], move_count=dec["ply"].astype(np.int32))

    def ev(m):
        net.load_planes(m.planes[:m.n])
        net.run(m.n)
        return net.logits, net.value
    acts, vis, nn = mb.search(ev, S, cfg.c_puct, add_noise=False, kind=xq_native.POLICY_LOGITS_BF16)
    torch.cuda.synchronize()
    acts, vis, nn = acts.cpu().numpy(), vis.cpu().numpy(), nn.cpu().numpy()
    for i in range(n):
        k = int(dec["n"][i])
        assert k == nn[i] and dec["actions"][i, :k].tolist() == acts[i, :k].tolist()
        got = np.round(dec["probs"][i, :k] * S).astype(np.int64)
        assert abs(dec["probs"][i, :k] * S - got).max() < 1e-4
        assert got.tolist() == vis[i, :k].tolist(), (i, got.tolist(), vis[i, :k].tolist())


@pytest.mark.parametrize("K,S", [(4, 24), (8, 30), (3, 7)])
def test_multi_leaf_search_visit_sums_legality_determinism(eng, oracle, net_model, K, S):
    cfg = Cfg()
    cfg.num_simulations = S
    runs = []
    for _ in range(2):
        sp, c, dec = run(eng, net_model, cfg, slots=10, games=10, plies=4, K=K, seed=21, add_noise=True)
        assert c["error"] == 0 and c["dropped"] == 0 and c["samples"] == 40
        assert c["sims"] == c["samples"] * S                    # a search makes exactly num_simulations descents, whatever K
        order = np.lexsort((dec["ply"], dec["uid"]))
        runs.append({k: dec[k][order] for k in ("board", "side", "n", "actions", "probs", "played", "uid", "ply")})
    a, b = runs
    for k in a:
        assert np.array_equal(a[k], b[k]), k                    # same seed, same records
    ea, en, _, _ = oracle.movegen_batch(a["board"], a["side"])
    assert np.array_equal(en, a["n"]) and np.array_equal(ea, a["actions"])     # root children = ordered legal moves
    for i in range(len(a["n"])):
        k = int(a["n"][i])
        cnt = a["probs"][i, :k] * S
        assert abs(cnt - np.round(cnt)).max() < 1e-4 and int(np.round(cnt).sum()) == S     # every descent is one root-child visit
        assert a["probs"][i, k:].sum() == 0
        j = a["actions"][i, :k].tolist().index(int(a["played"][i]))
        assert cnt[j] > 0
    if K >= 4 and S >= 24:
        # virtual loss spreads the first descents of a step: more root children visited than sequential search does at K = 1
        _, _, d1 = run(eng, net_model, cfg, slots=10, games=10, plies=1, K=1, seed=21, add_noise=True)
        _, _, dk = run(eng, net_model, cfg, slots=10, games=10, plies=1, K=K, seed=21, add_noise=True)
        assert (dk["probs"] > 0).sum() >= (d1["probs"] > 0).sum()


def test_multi_leaf_needs_a_network_batch_of_slots_times_k(eng, net_model):
    import xq_native
    from selfplay_engine import SelfPlayEngine
    sp = SelfPlayEngine(eng, net_model, n_slots=4, max_games=4, leaves_per_game=2)
    sp.reset()
    with pytest.raises(xq_native.XqError):
        sp.play(SelfPlayEngine.make_config(Cfg(), 4, leaves_per_game=4), 1)


def test_arena_with_multi_leaf_search_plays_legal_games(eng, oracle, net_model):
    import torch
    import xq_native
    import arena
    import model as M
    torch.manual_seed(9)
    other = M.XiangqiNet(128, 1).eval()
    e2 = xq_native.Engine(0)
    r = arena.Arena(e2, net_model, other, 6, 16, leaves_per_game=4).play(16, 1.5, 40)
    assert r["new_wins"] + r["old_wins"] + r["draws"] == 6
    for g in range(6):
        og = oracle.OracleGame()
        for ply in range(int(r["plies"][g])):
            a = int(r["moves"][g, ply])
            assert a in og.get_legal_actions().tolist()
            og.make_action(a)
        done, w = og.is_game_over()
        assert (done and w == r["winners"][g]) or (not done and r["plies"][g] >= 40 and r["winners"][g] == 0)
