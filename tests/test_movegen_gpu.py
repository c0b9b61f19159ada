"""K1 parity: the CUDA rules engine (through the C ABI) against the reference goldens and the
CPU oracle.  Bit-exact: ordered move lists, counts, in-check flags, feature planes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["warp", "thread"])
def eng(request):
    """Both K1 kernel generations (one warp per board / one thread per board) answer every test of this file."""
    import xq_native
    e = xq_native.Engine(0)
    e.set_movegen_impl(request.param)
    assert e.set_movegen_impl(request.param) == request.param
    yield e
    e.close()


def dev(eng, a):
    return eng.torch.from_numpy(np.ascontiguousarray(a)).to(eng.dev)


def run(eng, boards, sides, planes=False):
    a, n, c, p = eng.movegen(dev(eng, boards), dev(eng, sides), planes=planes)
    eng.torch.cuda.synchronize()
    return a.cpu().numpy(), n.cpu().numpy(), c.cpu().numpy(), None if p is None else p.cpu().numpy()


def test_reference_goldens(eng, rules_golden):
    g = rules_golden
    a, n, c, p = run(eng, g["board"], g["side"], planes=True)
    assert np.array_equal(n, g["n"])
    assert np.array_equal(a, g["actions"])
    assert np.array_equal(c, g["in_check"])
    assert np.array_equal(np.packbits(p.reshape(len(n), -1) > 0.5, axis=1), g["planes"])
    assert set(np.unique(p)) <= {0.0, 1.0}
    assert eng.overflow_count() == 0


def test_initial_position_44_and_six_ply_42(eng, oracle):
    g = oracle.OracleGame()
    a, n, c, _ = run(eng, g.board.reshape(1, 90), np.array([1], np.int8))
    assert n[0] == 44 and c[0] == 0                       # test_v3.py:115-120
    for fr, fc, tr, tc in [(2, 1, 4, 2), (7, 1, 5, 2), (0, 1, 2, 2), (9, 1, 7, 2), (3, 0, 4, 0), (6, 0, 5, 0)]:
        g.make_action((fr * 9 + fc) * 90 + tr * 9 + tc)
    a, n, c, _ = run(eng, g.board.reshape(1, 90), np.array([g.current_player], np.int8))
    assert n[0] == 42                                     # test_cython.py:62-84


def test_synthetic_unorthodox_boards(eng, attacked_golden):
    a, n, c, _ = run(eng, attacked_golden["syn_board"], attacked_golden["syn_side"])
    assert np.array_equal(n, attacked_golden["syn_n"])
    assert np.array_equal(a, attacked_golden["syn_actions"])
    assert np.array_equal(c, attacked_golden["syn_in_check"])


def test_unorthodox_random_boards_follow_oracle(eng, oracle):
    """Boards no game can reach (several kings, kings off-palace, piece soup): the general
    find-king path must agree with the Cython semantics restated in the oracle."""
    rs = np.random.RandomState(3)
    boards = np.zeros((4000, 90), np.int8)
    for i in range(len(boards)):
        k = rs.randint(2, 28)
        sq = rs.choice(90, k, replace=False)
        boards[i, sq] = rs.choice([-7, -6, -5, -4, -3, -2, -1, 1, 2, 3, 4, 5, 6, 7], k)
        if i % 3:  # usually put kings in palaces so that moves exist
            boards[i, rs.choice([3, 4, 5, 12, 13, 14, 21, 22, 23])] = 1
            boards[i, rs.choice([66, 67, 68, 75, 76, 77, 84, 85, 86])] = -1
    sides = rs.choice(np.array([1, -1], np.int8), len(boards))
    ea, en, ec, ep = oracle.movegen_batch(boards, sides, want_planes=True)
    a, n, c, p = run(eng, boards, sides, planes=True)
    assert np.array_equal(n, en) and np.array_equal(a, ea) and np.array_equal(c, ec)
    assert np.array_equal(p, ep)
    eng.overflow_count(reset=True)


def test_overflow_is_counted_and_lists_are_capped(eng, oracle):
    """Rook / cannon crowds with more than 128 legal moves (no game reaches them): the first 128 moves, n = 128 and the
    overflow counter -- never a silent truncation; the host call reports XQ_ERR_OVERFLOW."""
    import xq_native
    rs = np.random.RandomState(9)
    boards = np.zeros((600, 90), np.int8)
    sides = rs.choice(np.array([1, -1], np.int8), len(boards))
    for i in range(len(boards)):
        s = int(sides[i])
        sq = rs.choice(90, rs.randint(8, 13), replace=False)
        boards[i, sq] = rs.choice([5 * s, 6 * s], len(sq))
        boards[i, rs.choice([3, 4, 5, 12, 13, 14, 21, 22, 23] if s == 1 else [66, 67, 68, 75, 76, 77, 84, 85, 86])] = s
    ea, en, ec, _ = oracle.movegen_batch(boards, sides, allow_overflow=True)
    n_over = int((en == 128).sum())            # capped rows (a count of exactly 128 is not an overflow, but none occurs here)
    eng.overflow_count(reset=True)
    a, n, c, _ = run(eng, boards, sides)
    assert np.array_equal(n, en) and np.array_equal(a, ea) and np.array_equal(c, ec)
    assert n_over > 0 and 0 < eng.overflow_count(reset=True) <= n_over
    with pytest.raises(xq_native.XqError):
        eng.movegen_host(boards, sides)
    eng.overflow_count(reset=True)


@pytest.mark.parametrize("B", [0, 1, 31, 63, 64, 65, 127, 129, 1000])
def test_ragged_batches_and_unaligned_buffers(eng, oracle, B):
    boards, sides = oracle.random_playout_positions(11, max(B, 1))
    boards, sides = boards[:B], sides[:B]
    ea, en, ec, ep = oracle.movegen_batch(boards, sides, want_planes=True)
    a, n, c, p = run(eng, boards, sides, planes=True)
    assert np.array_equal(a, ea) and np.array_equal(n, en) and np.array_equal(c, ec) and np.array_equal(p, ep)
    if B:
        # caller buffer not 16-byte aligned -> the kernel must take its plain-copy path
        t = eng.torch
        raw_b = t.zeros(B * 90 + 1, dtype=t.int8, device=eng.dev)
        raw_s = t.zeros(B + 1, dtype=t.int8, device=eng.dev)
        raw_b[1:] = dev(eng, boards).reshape(-1)
        raw_s[1:] = dev(eng, sides)
        a2, n2, c2, _ = eng.movegen(raw_b[1:].reshape(B, 90), raw_s[1:])
        t.cuda.synchronize()
        assert np.array_equal(a2.cpu().numpy(), ea) and np.array_equal(n2.cpu().numpy(), en)
        assert np.array_equal(c2.cpu().numpy(), ec)


def test_host_api_equals_device_api(eng, oracle):
    boards, sides = oracle.random_playout_positions(21, 150001)       # > 2 pipeline chunks, ragged
    ea, en, ec, _ = oracle.movegen_batch(boards, sides)
    a, n, c, p = eng.movegen_host(boards, sides, planes=False)
    assert np.array_equal(a, ea) and np.array_equal(n, en) and np.array_equal(c, ec)
    sub = slice(0, 20000)
    a, n, c, p = eng.movegen_host(boards[sub], sides[sub], planes=True)
    _, _, _, ep = oracle.movegen_batch(boards[sub], sides[sub], want_planes=True)
    assert np.array_equal(p, ep)


def test_is_attacked_goldens_and_oracle(eng, oracle, attacked_golden):
    boards = attacked_golden["board"]
    want = attacked_golden["attacked"]
    P = len(boards)
    sq = np.tile(np.arange(90, dtype=np.uint8), 2 * P)
    by = np.tile(np.repeat(np.array([1, -1], np.int8), 90), P)
    got = eng.is_attacked_host(np.repeat(boards, 180, axis=0), sq, by)
    assert np.array_equal(got.reshape(P, 2, 90), want)
    for b, (kr, kc, byp, w) in zip(attacked_golden["known_board"], attacked_golden["known_query"]):
        got = eng.is_attacked_host(b.reshape(1, 90), np.array([kr * 9 + kc], np.uint8), np.array([byp], np.int8))
        assert bool(got[0]) == bool(w)                                # test_v3.py:139-197
    boards, sides = oracle.random_playout_positions(5, 50000)
    rs = np.random.RandomState(0)
    sq = rs.randint(0, 90, len(sides)).astype(np.uint8)
    by = rs.choice(np.array([1, -1], np.int8), len(sides))
    got = eng.is_attacked(dev(eng, boards), dev(eng, sq), dev(eng, by)).cpu().numpy()
    assert np.array_equal(got, oracle.is_attacked_batch(boards, sq, by))


def test_device_playouts_are_legal_games(eng, oracle):
    """The device random-playout generator (bench input) plays by the reference rules: every
    transition is a legal move of the oracle, games stop exactly when is_game_over fires
    (mate/stalemate, 120 quiet plies, 200 plies + material, repetition) with the same winner."""
    G = 96
    boards, sides, npos, win = eng.random_playouts(1234, G, compact=False)
    eng.torch.cuda.synchronize()
    boards = boards.cpu().numpy().reshape(G, 201, 90)
    sides = sides.cpu().numpy().reshape(G, 201)
    npos = npos.cpu().numpy()
    win = win.cpu().numpy()
    ended = {1: 0, -1: 0, 0: 0}
    for g in range(G):
        og = oracle.OracleGame()
        for ply in range(npos[g]):
            assert np.array_equal(og.board.reshape(90), boards[g, ply]) and sides[g, ply] == og.current_player
            done, w = og.is_game_over()
            if ply == npos[g] - 1:
                assert done and w == win[g], (g, ply, done, w, win[g])
                ended[int(w)] += 1
                break
            assert not done
            diff = np.nonzero(boards[g, ply] != boards[g, ply + 1])[0]
            assert len(diff) == 2
            frm = diff[0] if boards[g, ply + 1][diff[0]] == 0 else diff[1]
            to = diff[1] if frm == diff[0] else diff[0]
            a = int(frm) * 90 + int(to)
            assert a in og.get_legal_actions().tolist()
            og.make_action(a)
        assert (sides[g, npos[g]:] == 0).all()
    assert sum(ended.values()) == G


def test_full_size_million_positions(eng, oracle):
    """BASELINE config 1: 1M random-playout positions, bit-exact against the oracle, plus
    size-independent plane properties over the whole set."""
    boards_t, sides_t, npos, _ = eng.random_playouts(20261018, 5600)
    N = 1_000_000
    assert boards_t.shape[0] >= N
    boards_t, sides_t = boards_t[:N].contiguous(), sides_t[:N].contiguous()
    a, n, c, p = eng.movegen(boards_t, sides_t, planes=True)
    t = eng.torch
    t.cuda.synchronize()
    assert eng.overflow_count() == 0
    boards, sides = boards_t.cpu().numpy(), sides_t.cpu().numpy()
    ea, en, ec, _ = oracle.movegen_batch(boards, sides)
    assert np.array_equal(n.cpu().numpy(), en)
    assert np.array_equal(c.cpu().numpy(), ec)
    assert np.array_equal(a.cpu().numpy(), ea)
    # planes: one-hot per occupied square, turn plane == red to move; exact on a slice
    occ = (boards_t != 0).sum(dim=1).to(t.float32)
    assert t.equal(p[:, :14].sum(dim=(1, 2, 3)), occ)
    assert t.equal(p[:, 14].sum(dim=(1, 2)), (sides_t == 1).to(t.float32) * 90)
    _, _, _, ep = oracle.movegen_batch(boards[:50000], sides[:50000], want_planes=True)
    assert np.array_equal(p[:50000].cpu().numpy(), ep)
    # idempotence: a second launch over the same inputs gives identical bytes
    a2, n2, c2, _ = eng.movegen(boards_t, sides_t)
    assert t.equal(a2, a) and t.equal(n2, n) and t.equal(c2, c)


def test_output_buffers_with_minimal_alignment_and_impl_switch(eng, oracle):
    """The C ABI asks for 8-byte aligned move / plane buffers.  A planes pointer that is 8- but not 16-byte aligned (the
    thread-per-board kernel stores float4) must still give the same bytes, and a bad implementation id is an error."""
    import xq_native
    t = eng.torch
    B = 333
    boards, sides = oracle.random_playout_positions(31, B)
    ea, en, ec, ep = oracle.movegen_batch(boards, sides, want_planes=True)
    raw_p = t.zeros(B * 1350 + 2, dtype=t.float32, device=eng.dev)
    raw_a = t.zeros(B * 128 + 4, dtype=t.int16, device=eng.dev)
    pl = raw_p[2:].reshape(B, 15, 10, 9)
    ac = raw_a[4:].reshape(B, 128)
    assert pl.data_ptr() % 16 == 8 and ac.data_ptr() % 16 == 8
    out = (ac, t.empty(B, dtype=t.uint8, device=eng.dev), t.empty(B, dtype=t.uint8, device=eng.dev), pl)
    a, n, c, p = eng.movegen(dev(eng, boards), dev(eng, sides), planes=True, out=out)
    t.cuda.synchronize()
    assert np.array_equal(a.cpu().numpy(), ea) and np.array_equal(n.cpu().numpy(), en) and np.array_equal(c.cpu().numpy(), ec)
    assert np.array_equal(p.cpu().numpy(), ep)
    assert float(raw_p[:2].abs().sum()) == 0 and int(raw_a[:4].abs().sum()) == 0      # nothing written in front of the buffers
    cur = eng.movegen_impl
    with pytest.raises(xq_native.XqError):
        eng.set_movegen_impl(7)
    assert eng.movegen_impl == cur
