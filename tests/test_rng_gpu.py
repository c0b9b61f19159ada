"""Distribution tests of the device RNG against the reference's numpy draws: root Dirichlet(0.3) noise
(np.random.dirichlet, mcts.py:117-121) and the temperature sampling of the move (np.random.choice(p=probs),
parallel_selfplay.py:105).  The streams cannot be bit-equal (different generators); the distributions must be."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import game
    return game.engine(0)


def test_device_dirichlet_noise_has_numpy_dirichlet_marginals(eng, oracle):
    import torch
    from scipy import stats
    import xq_native
    G, alpha = 8192, 0.3
    g0 = oracle.OracleGame()
    n = len(g0.get_legal_actions())                                  # 44 root moves
    mb = xq_native.MctsBatch(xq_native.Engine(0), G)
    mb.set_games(np.stack([g0.board.reshape(90)] * G), np.full(G, 1, np.int8))
    mb.root_begin(want_planes=False)
    uniform = torch.full((G, 8100), 1.0 / 8100, dtype=torch.float32, device=eng.dev)
    mb.root_expand(uniform, xq_native.POLICY_PROBS, noise=None, add_noise=True, seed=20261019, alpha=alpha)
    pri, nn = mb.root_priors()
    torch.cuda.synchronize()
    pri = pri.cpu().numpy()[:, :n]
    assert (nn.cpu().numpy() == n).all()
    eta = (pri - 0.75 * np.float64(np.float32(1.0) / np.float32(n))) / 0.25     # 0.75 * P + 0.25 * eta, P = 1/44 in float32
    assert np.abs(eta.sum(1) - 1.0).max() < 1e-5 and eta.min() > -1e-6
    # (1) marginal of a symmetric Dirichlet: eta_i ~ Beta(alpha, (n-1) alpha); one component per game = independent sample
    for comp in (0, 17, n - 1):
        ks = stats.kstest(np.clip(eta[:, comp], 0, 1), stats.beta(alpha, (n - 1) * alpha).cdf)
        assert ks.pvalue > 1e-3, (comp, ks)
    # (2) two-sample test against numpy's own generator
    ref = np.random.default_rng(7).dirichlet([alpha] * n, size=G)
    assert stats.ks_2samp(eta[:, 5], ref[:, 5]).pvalue > 1e-3
    # (3) moments over all components: mean 1/n, variance (n-1) / (n^2 (n alpha + 1))
    assert abs(eta.mean() - 1.0 / n) < 1e-6
    var = (n - 1) / (n * n * (n * alpha + 1))
    assert abs(eta.var() / var - 1.0) < 0.03
    # (4) games draw different noise, components are not sorted or otherwise ordered
    assert np.abs(np.corrcoef(eta[:, 0], eta[:, 1])[0, 1] + 1.0 / (n - 1)) < 0.05


def test_temperature_sampling_follows_the_recorded_distribution(eng):
    """Every ply samples its move from the visit distribution it records (T = 1: N / S).  Over many independent plies the
    rank of the played move (0 = most visited) must be distributed as the recorded probabilities say: chi-square of the
    observed rank counts against the sum of the per-ply probabilities."""
    import torch
    from scipy import stats
    import model as M
    from selfplay_engine import SelfPlayEngine, decode_samples

    class Cfg:
        num_simulations, c_puct, temperature_threshold, max_game_length = 16, 1.5, 400, 300
        random_opening_moves, enable_resign, resign_threshold, resign_check_steps = 6, False, -0.9, 5
    torch.manual_seed(12)
    net = M.XiangqiNet(128, 1).eval()
    sp = SelfPlayEngine(eng, net, n_slots=2048, max_games=2048)
    sp.reset()
    sp.play(SelfPlayEngine.make_config(Cfg(), 2048, seed=99, add_noise=True), 3)
    c = sp.counters()
    raw, _, _ = sp.fetch(0, c["samples"])
    d = decode_samples(raw)
    assert c["samples"] == 3 * 2048
    R = 6
    exp = np.zeros(R + 1)
    obs = np.zeros(R + 1)
    for i in range(len(d["n"])):
        k = int(d["n"][i])
        p = d["probs"][i, :k].astype(np.float64)
        order = np.argsort(-p, kind="stable")
        ps = p[order]
        m = min(R, k)                          # a position in check may have fewer than R legal moves
        exp[:m] += ps[:m]
        exp[R] += ps[R:].sum()
        j = d["actions"][i, :k].tolist().index(int(d["played"][i]))
        r = int(np.nonzero(order == j)[0][0])
        obs[min(r, R)] += 1
    assert abs(exp.sum() - len(d["n"])) < 1e-3
    chi2 = ((obs - exp) ** 2 / exp).sum()
    assert stats.chi2.sf(chi2, R) > 1e-3, (chi2, obs, exp)
    assert obs[1:].sum() > 0.2 * obs.sum()            # it does sample: not the argmax every time
