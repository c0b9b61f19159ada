"""CPU ORACLE of the training-step pieces -- numpy restatement, test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product package
(xiangqi-alphazero_b200/) never does.  Each function restates the reference lines it cites
(/root/reference/training/...).

Parity status: PINNED.  tests/test_train_cpu.py checks `augment`, `sample_tuple` and `policy_value_loss` against
fixtures produced by RUNNING THE REFERENCE (tests/golden/make_train_golden.py: `_augment_data` outputs, and the
loss statistics `AlphaZeroTrainer.train_network` reports for one full-buffer batch with the learning rate at 0);
`clip_and_adam` restates torch.nn.utils.clip_grad_norm_ + torch.optim.Adam (third-party: torch, unpinned by the
reference's requirements.txt `torch>=2.0.0`; 2.11.0 here) and is checked against torch itself in the same test.
"""
import numpy as np

ACTION_SPACE = 8100


def planes(board, side):
    """XiangqiGame.get_state_for_nn, game.py:618-640: 7 planes of the side to move, 7 of the opponent, plane 14 = 1 iff red moves."""
    own = np.asarray(board, np.int8).reshape(10, 9).astype(np.int32) * int(side)
    f = np.zeros((15, 10, 9), np.float32)
    for k in range(1, 8):
        f[k - 1] = own == k
        f[6 + k] = own == -k
    if side == 1:
        f[14] = 1.0
    return f


def mirror_action(a):
    """(fr,fc,tr,tc) -> (fr,8-fc,tr,8-tc) on the from*90+to encoding, parallel_selfplay.py:143-148 / train.py:141-147."""
    a = np.asarray(a, np.int64)
    f, t = a // 90, a % 90
    fm = (f // 9) * 9 + (8 - f % 9)
    tm = (t // 9) * 9 + (8 - t % 9)
    return fm * 90 + tm


def sample_tuple(board, side, actions, probs, n, z, mirrored=False):
    """One training tuple (planes float32[15,10,9], policy float32[8100], z) from a sparse record; `mirrored` gives the
    twin `_augment_data` appends (parallel_selfplay.py:137-151: np.flip(state, axis=2) + the action permutation)."""
    st = planes(board, side)
    a = np.asarray(actions[:n], np.int64)
    pol = np.zeros(ACTION_SPACE, np.float32)
    if mirrored:
        st = np.flip(st, axis=2).copy()
        a = mirror_action(a)
    pol[a] = np.asarray(probs[:n], np.float32)
    return st, pol, np.float32(z)


def augment(data):
    """_augment_data, parallel_selfplay.py:137-151, on dense tuples: [(s, p, v)] -> [(s, p, v), (flip(s), permuted p, v), ...]."""
    out = []
    for st, pol, v in data:
        out.append((st, pol, v))
        fp = np.zeros_like(pol)
        nz = np.nonzero(pol > 0)[0]
        fp[mirror_action(nz)] = pol[nz]
        out.append((np.flip(st, axis=2).copy(), fp, v))
    return out


def z_label(winner, side):
    """parallel_selfplay.py:124-132: 0 for a draw, +1 when the side to move at the sample won, -1 otherwise."""
    return 0.0 if winner == 0 else (1.0 if winner == side else -1.0)


def policy_value_loss(logits, value, target_policy, z):
    """train.py:408-413 in float64: policy_loss = -mean(sum(pi * log_softmax(logits), 1)), value_loss = mse(value, z).
    Also returns d(policy_loss + value_loss)/dlogits and /dvalue."""
    lg = np.asarray(logits, np.float64)
    pi = np.asarray(target_policy, np.float64)
    v = np.asarray(value, np.float64).reshape(-1)
    zz = np.asarray(z, np.float64).reshape(-1)
    B = lg.shape[0]
    m = lg.max(axis=1, keepdims=True)
    lse = m + np.log(np.exp(lg - m).sum(axis=1, keepdims=True))
    logp = lg - lse
    p_loss = -(pi * logp).sum(axis=1).mean()
    v_loss = ((v - zz) ** 2).mean()
    g_logits = (np.exp(logp) * pi.sum(axis=1, keepdims=True) - pi) / B
    g_value = 2.0 * (v - zz) / B
    return p_loss, v_loss, g_logits, g_value


def clip_and_adam(p, g, m, v, step, lr=0.002, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_(params, max_norm) then torch.optim.Adam.step (train.py:190-194, 418-419) on flat float64
    arrays; returns the new (p, m, v).  step counts from 1."""
    p, g, m, v = (np.asarray(x, np.float64).copy() for x in (p, g, m, v))
    if max_norm and max_norm > 0:
        g = g * min(1.0, max_norm / (np.sqrt((g * g).sum()) + 1e-6))
    g = g + weight_decay * p
    b1, b2 = betas
    m = m + (1 - b1) * (g - m)
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p = p - (lr / bc1) * m / (np.sqrt(v) / np.sqrt(bc2) + eps)
    return p, m, v
