"""The drop-in Python modules (game / mcts / model / parallel_selfplay / inference_server) keep the
reference's contracts: these are the acceptance tests a reference user would run (test_v3.py,
test_cython.py, test_gpu_train.py shapes), with the oracle as the checker."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_xiangqigame_random_games_follow_the_rules(oracle):
    """test_cython.py:87-123 / test_v3.py:16-103 recipe on the drop-in XiangqiGame."""
    import game as G
    rnd = random.Random(11)
    for _ in range(6):
        g, og = G.XiangqiGame(), oracle.OracleGame()
        for _ply in range(300):
            done, winner = g.is_game_over()
            odone, owinner = og.is_game_over()
            assert (done, winner) == (odone, owinner)
            if done:
                break
            moves = g.get_legal_moves()
            acts = g.get_legal_actions()
            assert acts == og.get_legal_actions().tolist()               # ordered, like the Cython engine
            assert len(acts) == len(moves) and g.get_legal_moves() is moves   # cache coherence (test_v3.py:40-47)
            assert all(G.encode_action(*G.decode_action(a)) == a for a in acts[:5])
            assert g._is_in_check(g.current_player) == oracle.in_check(g.board, g.current_player)
            assert np.array_equal(g.get_state_for_nn(), og.get_state_for_nn())
            assert g.get_material_score(1) == og.material(1) and g.get_material_score(-1) == og.material(-1)
            m = rnd.choice(moves)
            assert g.board[m[0], m[1]] != 0
            g.make_move(*m)
            og.make_action(G.encode_action(*m))
            assert g._legal_moves_cache is None
        c = g.clone()
        assert np.array_equal(c.board, g.board) and c.history == g.history and c.move_count == g.move_count


def test_known_positions_of_the_reference_tests():
    import game as G
    g = G.XiangqiGame()
    assert len(g.get_legal_moves()) == 44                            # test_v3.py:115-120
    g3 = G.XiangqiGame(); g3.board[:] = 0
    g3.board[0, 4], g3.board[9, 4], g3.board[5, 4] = 1, -1, -5
    g3._legal_moves_cache = None
    assert g3._is_in_check(1) is True                                # test_v3.py:139-151
    g4 = G.XiangqiGame(); g4.board[:] = 0
    g4.board[0, 4], g4.board[2, 3] = 1, -4
    assert G.XiangqiGame._is_attacked(g4.board, 0, 4, -1) is True    # test_v3.py:154-166
    g4.board[1, 3] = 7
    assert G.XiangqiGame._is_attacked(g4.board, 0, 4, -1) is False   # horse leg, test_v3.py:169-181
    g6 = G.XiangqiGame(); g6.board[:] = 0
    g6.board[0, 4], g6.board[9, 4], g6.board[5, 4], g6.board[8, 4] = 1, -1, 7, -6
    assert G.XiangqiGame._is_attacked(g6.board, 0, 4, -1) is True    # cannon + screen, test_v3.py:184-197


def test_mcts_dropin_equals_oracle_search_with_the_same_network(oracle):
    """MCTS(model).search on the GPU tree == the reference algorithm (oracle) driven by the very same
    predict(): visit counts must be identical because both consume identical float32 probabilities."""
    import torch
    import game as G
    import mcts as MC
    import model as M
    torch.manual_seed(4)
    net = M.XiangqiNet(128, 1).eval()
    g, og = G.XiangqiGame(), oracle.OracleGame()
    for a in (1792, 6337):
        g.make_action(a)
        og.make_action(a)

    def predict(board, player):
        return net.predict(oracle.planes(board, player))
    search = MC.MCTS(net, num_simulations=48, c_puct=1.5)
    probs = search.search(g, temperature=1.0, add_noise=False)
    acts, vis, _, _ = oracle.mcts_search(og, 48, 1.5, predict)
    want = np.zeros(8100)
    want[acts.astype(np.int64)] = vis / 48.0
    assert np.array_equal(probs, want)
    assert search.get_action(g, temperature=0.0) == int(acts[int(np.argmax(vis))])   # first max wins
    # with noise: a distribution over legal moves only
    p2 = search.search(g, temperature=0.3, add_noise=True)
    assert abs(p2.sum() - 1.0) < 1e-9 and set(np.nonzero(p2)[0]) <= set(acts.tolist())


def test_one_training_iteration_dropin():
    """test_gpu_train.py shape: self-play -> samples -> one optimiser step on the torch module ->
    self-play again with the UPDATED weights (the kernel-side copy must follow the parameters)."""
    import torch
    import torch.nn.functional as F
    import model as M
    import parallel_selfplay as ps

    class Cfg:
        num_simulations, c_puct, temperature_threshold, max_game_length = 8, 1.5, 20, 60
        random_opening_moves, enable_resign, resign_threshold, resign_check_steps = 4, True, -0.9, 5
        num_games_per_iter = 4
    torch.manual_seed(0)
    net = M.XiangqiNet(128, 1)
    data, stats = ps.parallel_self_play(net, Cfg(), num_workers=2, use_gpu_server=True, gpu_device='cpu')
    assert stats['games'] == 4 and len(data) > 0
    dev = torch.device('cuda')
    net.to(dev).train()
    opt = torch.optim.Adam(net.parameters(), lr=2e-3, weight_decay=1e-4)
    s = torch.from_numpy(np.stack([d[0] for d in data[:64]])).to(dev)
    p = torch.from_numpy(np.stack([d[1] for d in data[:64]])).float().to(dev)
    z = torch.tensor([d[2] for d in data[:64]], dtype=torch.float32, device=dev)
    logits, v = net(s)
    loss = -(p * F.log_softmax(logits, dim=1)).sum(1).mean() + F.mse_loss(v.squeeze(1), z)   # train.py:410-414
    loss.backward()
    torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
    opt.step()
    net.eval()
    state = data[0][0]
    before = net._b200_version
    probs, value = net.predict(state)                                 # refolds: parameters changed
    assert net._b200_version != before
    with torch.no_grad():
        lr, vr = net(torch.from_numpy(state)[None].to(dev))
    pr = torch.softmax(lr, 1)[0].cpu().numpy()
    assert np.abs(probs - pr).max() / pr.max() < 1e-2 and abs(value - float(vr)) < 1e-2
    data2, stats2 = ps.parallel_self_play(net, Cfg())
    assert stats2['games'] == 4 and stats2['mode'] == 'gpu'


def test_inference_server_shim():
    import torch
    import model as M
    import inference_server as IS
    torch.manual_seed(1)
    net = M.XiangqiNet(128, 1)
    srv = IS.InferenceServer(M.XiangqiNet, {'num_channels': 128, 'num_res_blocks': 1},
                             {k: v.clone() for k, v in net.state_dict().items()}, device='cuda', num_workers=2)
    path = srv.start()
    cl = IS.InferenceClient(0, path)
    state = np.zeros((15, 10, 9), np.float32); state[14] = 1; state[0, 0, 4] = 1; state[7, 9, 4] = 1
    probs, value = cl.predict(state)
    p2, v2 = net.eval().predict(state)
    assert probs.shape == (8100,) and np.array_equal(probs, p2) and value == v2
    srv.stop()
    with pytest.raises(RuntimeError):
        cl.predict(state)
    cl.close()


def test_hand_rolled_search_loop_like_benchmark_py(oracle):
    """benchmark.py:18-153 unrolls MCTS.search around MCTSNode (root.expand / select_child / backup) with
    game.clone / make_action / is_game_over and model.predict.  The same loop on the drop-ins must walk
    the same tree as the device search (and the oracle)."""
    import torch
    import game as G
    import mcts as MC
    import model as M
    torch.manual_seed(7)
    net = M.XiangqiNet(128, 1).eval()
    g = G.XiangqiGame()
    root = MC.MCTSNode()
    probs, _ = net.predict(g.get_state_for_nn(), 'cpu')
    priors = MC.MCTS._mask_and_normalize(probs, g.get_legal_actions())
    root.expand(priors)
    sims = 24
    for _ in range(sims):
        node, sim = root, g.clone()
        while not node.is_leaf():
            action, node = node.select_child(1.5)
            sim.make_action(action)
        done, winner = sim.is_game_over()
        if done:
            value = 0.0 if winner == 0 else 1.0                       # mcts.py:140
        else:
            p, value = net.predict(sim.get_state_for_nn(), 'cpu')
            node.expand(MC.MCTS._mask_and_normalize(p, sim.get_legal_actions()))
            value = -value
        node.backup(value)
    host_counts = [c.visit_count for c in root.children.values()]
    dev = MC.MCTS(net, num_simulations=sims, c_puct=1.5).search(g, temperature=1.0, add_noise=False)
    assert [int(round(dev[a] * sims)) for a in root.children] == host_counts


def test_is_move_legal_agrees_with_the_move_list():
    """game.py:441-490: every listed move passes, moves that leave the king attacked or facing fail."""
    import game as G
    g = G.XiangqiGame()
    g._init_board()
    assert len(g.get_legal_moves()) == 44
    rnd = random.Random(5)
    for _ in range(40):
        moves = g.get_legal_moves()
        if not moves:
            break
        legal = set(moves)
        side = g.current_player
        for m in rnd.sample(moves, min(4, len(moves))):
            assert g._is_move_legal(*m, side) is True
        before = g.board.copy()
        own = [(r, c) for r in range(10) for c in range(9) if g.board[r, c] * side > 0]
        for (r, c) in own[:6]:
            for (tr, tc) in ((r + 1, c), (r, c + 1)):
                if 0 <= tr < 10 and 0 <= tc < 9 and g.board[tr, tc] * side <= 0 and (r, c, tr, tc) in legal:
                    assert g._is_move_legal(r, c, tr, tc, side)
        assert np.array_equal(before, g.board)                        # the board is left untouched
        g.make_move(*rnd.choice(moves))
    # facing kings: the red king may not step onto the open file of the black king
    f = G.XiangqiGame(); f.board[:] = 0
    f.board[0, 3], f.board[9, 4] = 1, -1
    assert f._is_move_legal(0, 3, 0, 4, 1) is False and f._is_move_legal(0, 3, 1, 3, 1) is True
    assert (0, 3, 0, 4) not in f.get_legal_moves() and (0, 3, 1, 3) in f.get_legal_moves()


def test_single_game_entry_points_of_the_reference(tmp_path):
    """parallel_selfplay._play_one_game (:42-134) and AlphaZeroTrainer.self_play_game / _serial_self_play
    (train.py:227-301, 329-374): return shapes, labels and buffer accounting."""
    import torch
    import model as M
    import parallel_selfplay as ps
    import train as T

    class Cfg:
        num_simulations, c_puct, temperature_threshold, max_game_length = 8, 1.5, 4, 12
        random_opening_moves, enable_resign, resign_threshold, resign_check_steps = 2, True, -0.9, 5
        num_games_per_iter = 1
    torch.manual_seed(3)
    net = M.XiangqiNet(128, 1).eval()
    data, winner, plies = ps._play_one_game(net, Cfg(), 'cpu')
    assert winner in (1, -1, 0) and 0 < len(data) <= plies <= 12 + 2
    for st, pol, z in data:
        assert st.shape == (15, 10, 9) and pol.shape == (8100,) and abs(pol.sum() - 1) < 1e-6
        assert z == (0.0 if winner == 0 else (1.0 if (st[14, 0, 0] == 1) == (winner == 1) else -1.0))

    cfg = T.TrainingConfig()
    cfg.num_channels, cfg.num_res_blocks, cfg.num_simulations, cfg.num_games_per_iter = 128, 1, 6, 2
    cfg.max_game_length, cfg.temperature_threshold, cfg.random_opening_moves = 5, 2, 2
    cfg.checkpoint_dir = str(tmp_path)
    tr = T.AlphaZeroTrainer(cfg)
    random.seed(1); np.random.seed(1)
    data, winner, steps = tr.self_play_game()
    assert steps == len(data) == 5 and winner in (1, -1, 0)
    assert all(abs(p.sum() - 1) < 1e-9 and z in (-1.0, 0.0, 1.0) for _, p, z in data)
    st = tr._serial_self_play()
    assert st['games'] == 2 and st['new_samples'] == 20 == st['buffer_size'] == len(tr.replay_buffer)
    assert st['red_wins'] + st['black_wins'] + st['draws'] == 2
    pair = T.augment_data(*data[0])
    assert np.array_equal(pair[1][0], np.flip(data[0][0], axis=2)) and abs(pair[1][1].sum() - 1) < 1e-9
    g = T.make_random_opening(T.XiangqiGame(), 3)
    assert g.move_count == 3 and len(g.history) == 3
