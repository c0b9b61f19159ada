"""Host logic of the training path that needs no GPU: shuffles identical to the reference's DataLoader, minibatch
sharding, evaluation-pair sharding, and the 2-rank gloo fan-in of new sample records."""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_epoch_permutation_is_the_dataloaders():
    from torch.utils.data import DataLoader, TensorDataset
    import train as T
    for n, bs in ((1000, 256), (300, 64), (7, 3)):
        torch.manual_seed(123 + n)
        dl = DataLoader(TensorDataset(torch.arange(n)), batch_size=bs, shuffle=True, num_workers=0, drop_last=False)   # train.py:384-391
        want = [torch.cat([b[0] for b in dl]) for _ in range(3)]
        torch.manual_seed(123 + n)
        got = [T.epoch_permutation(n) for _ in range(3)]
        assert all(torch.equal(a, b) for a, b in zip(want, got))


def test_shard_batch_partitions_every_global_batch():
    import train as T
    for n in (256, 255, 44, 9, 8):
        idx = torch.randperm(1000)[:n]
        for world in (1, 2, 4, 8):
            parts = [T.shard_batch(idx, r, world) for r in range(world)]
            assert torch.equal(torch.cat(parts), idx) and min(p.numel() for p in parts) >= 1
    idx = torch.arange(5)
    assert all(torch.equal(T.shard_batch(idx, r, 8), idx) for r in range(8))      # fewer samples than ranks: replicated


def test_shard_pairs_keeps_colour_parity():
    import arena
    for games in range(0, 40):
        for world in (1, 2, 4, 8):
            mine = [arena.shard_pairs(games, r, world) for r in range(world)]
            assert sum(mine) == games
            # a rank's local game i has the new model as red iff i is even: every rank must hold whole pairs, except that
            # the rank with the last (unpaired, red) game holds it last
            assert sum(m % 2 for m in mine) == games % 2
            reds = sum((m + 1) // 2 for m in mine)
            assert reds == (games + 1) // 2                                       # train.py:474 new_is_red = (game_idx % 2 == 0)


def test_dense_tuples_round_trip_through_records():
    import xq_oracle
    from replay import dense_to_records
    from selfplay_engine import decode_samples, samples_to_reference_tuples
    boards, sides = xq_oracle.random_playout_positions(4, 60)
    acts, cnt, _, _ = xq_oracle.movegen_batch(boards, sides)
    rs = np.random.RandomState(0)
    data = []
    for i in range(len(sides)):
        if cnt[i] == 0:
            continue
        pol = np.zeros(8100)
        p = rs.dirichlet([1.0] * int(cnt[i])).astype(np.float32)
        pol[acts[i, :cnt[i]].astype(np.int64)] = p
        data.append((xq_oracle.planes(boards[i], int(sides[i])), pol, float(rs.choice([-1, 0, 1]))))
    rec, z = dense_to_records(data)
    dec = decode_samples(rec)
    for i, (st, pol, val) in enumerate(data):
        assert np.array_equal(xq_oracle.planes(dec["board"][i], int(dec["side"][i])), st)
        got = np.zeros(8100)
        got[dec["actions"][i, :dec["n"][i]].astype(np.int64)] = dec["probs"][i, :dec["n"][i]]
        assert np.array_equal(got.astype(np.float32), pol.astype(np.float32)) and z[i] == val


def test_two_rank_gloo_record_fan_in(tmp_path):
    script = os.path.join(ROOT, "tests", "mrank_gather_script.py")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29653", script, str(tmp_path)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    a = torch.load(os.path.join(str(tmp_path), "rank0.pt"))
    b = torch.load(os.path.join(str(tmp_path), "rank1.pt"))
    assert torch.equal(a["rec"], b["rec"]) and torch.equal(a["z"], b["z"]) and torch.equal(a["perm"], b["perm"])
    assert a["rec"].shape[0] == 5 + 9 and a["rec"][:5, 0].eq(0).all() and a["rec"][5:, 0].eq(1).all()


# ---- the training oracle (oracle/xq_train_oracle.py) pinned against outputs of the reference itself ------------------
GOLDEN = os.path.join(ROOT, "tests", "golden", "train_golden.npz")


def test_oracle_augment_equals_the_references_augment_data():
    import xq_train_oracle as O
    g = dict(np.load(GOLDEN))
    for i in range(len(g["mirror_planes"])):
        n = int(g["n"][i])
        st, pol, z = O.sample_tuple(g["board"][i], g["side"][i], g["actions"][i], g["probs"][i], n, g["z"][i], mirrored=True)
        assert np.array_equal(np.packbits(st.reshape(-1) > 0.5), g["mirror_planes"][i])          # parallel_selfplay.py:142
        nz = np.nonzero(pol > 0)[0]
        k = int((g["mirror_index"][i] >= 0).sum())
        assert np.array_equal(nz, g["mirror_index"][i, :k]) and np.array_equal(pol[nz], g["mirror_value"][i, :k])
        # and the dense-tuple form of the same function
        s0, p0, _ = O.sample_tuple(g["board"][i], g["side"][i], g["actions"][i], g["probs"][i], n, g["z"][i])
        pair = O.augment([(s0, p0, z)])
        assert np.array_equal(pair[1][0], st) and np.array_equal(pair[1][1], pol)


def test_oracle_loss_equals_the_references_full_batch_loss():
    """The reference's train_network on one batch holding the whole buffer with lr = 0 reports the loss of the initial
    weights; the oracle's tuples + loss on the same network must give the same numbers (train.py:408-413)."""
    import xq_train_oracle as O
    from model import XiangqiNet
    g = dict(np.load(GOLDEN))
    ch, blocks, records, batch, epochs, seed = (int(x) for x in g["meta"])
    torch.manual_seed(seed)
    net = XiangqiNet(ch, blocks).train()
    tuples = []
    for i in range(records):
        for m in (False, True):
            tuples.append(O.sample_tuple(g["board"][i], g["side"][i], g["actions"][i], g["probs"][i], int(g["n"][i]), g["z"][i], m))
    x = torch.from_numpy(np.stack([t[0] for t in tuples]))
    with torch.no_grad():
        logits, value = net(x)
    pl, vl, _, _ = O.policy_value_loss(logits.numpy(), value.numpy(), np.stack([t[1] for t in tuples]), np.array([t[2] for t in tuples]))
    assert abs(pl - g["full_batch"][0]) < 2e-4 * abs(g["full_batch"][0]) and abs(vl - g["full_batch"][1]) < 2e-4


def test_oracle_clip_and_adam_equals_torch():
    import xq_train_oracle as O
    torch.manual_seed(3)
    p = torch.nn.Parameter(torch.randn(1000, dtype=torch.float64))
    opt = torch.optim.Adam([p], lr=0.002, weight_decay=1e-4)
    pn, m, v = p.detach().numpy().copy(), np.zeros(1000), np.zeros(1000)
    for step in range(1, 8):
        grad = torch.randn(1000, dtype=torch.float64) * (5.0 if step % 2 else 0.01)
        p.grad = grad.clone()
        torch.nn.utils.clip_grad_norm_([p], 1.0)
        opt.step()
        pn, m, v = O.clip_and_adam(pn, grad.numpy(), m, v, step)
        assert np.allclose(pn, p.detach().numpy(), rtol=1e-10, atol=1e-12)


def test_train_module_augment_data_and_dataset_follow_the_reference():
    """train.augment_data (train.py:132-151) and SelfPlayDataset (train.py:114-129) against the committed outputs of the
    reference's own augmentation (tests/golden/make_train_golden.py)."""
    import ast
    src = open(os.path.join(ROOT, "xiangqi-alphazero_b200", "train.py")).read()
    # the module imports the CUDA library's bindings; the two helpers are pure numpy/torch and are checked on their own
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in ("augment_data", "SelfPlayDataset")]
    ns = {"np": np, "torch": torch, "List": list, "Tuple": tuple}
    sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
    exec(compile(ast.Module(body=keep, type_ignores=[]), "train.py", "exec"), ns)
    import xq_train_oracle as O
    g = dict(np.load(GOLDEN))
    for i in range(len(g["mirror_planes"])):
        n = int(g["n"][i])
        s0, p0, z = O.sample_tuple(g["board"][i], g["side"][i], g["actions"][i], g["probs"][i], n, g["z"][i])
        pair = ns["augment_data"](s0, p0, z)
        assert pair[0][0] is s0 and pair[0][1] is p0 and len(pair) == 2
        assert np.array_equal(np.packbits(pair[1][0].reshape(-1) > 0.5), g["mirror_planes"][i])
        nz = np.nonzero(pair[1][1] > 0)[0]
        k = int((g["mirror_index"][i] >= 0).sum())
        assert np.array_equal(nz, g["mirror_index"][i, :k]) and np.array_equal(pair[1][1][nz], g["mirror_value"][i, :k])
        ds = ns["SelfPlayDataset"](pair)
        st, pol, val = ds[1]
        assert len(ds) == 2 and st.dtype == torch.float32 and st.shape == (15, 10, 9) and pol.shape == (8100,) and val.shape == (1,)
        assert float(val) == float(z)


def test_dp_batchnorm_equals_batchnorm_on_the_global_minibatch():
    """DPBatchNorm2d (train.py: one all-reduce of the per-channel sums each way, no host synchronisation) = nn.BatchNorm2d
    on the concatenated minibatch: outputs, input gradients, parameter gradients (summed over the ranks, as the gradient
    all-reduce does) and running statistics.  Two ranks are emulated by feeding each half the other half's partial sums."""
    import torch
    import train as T
    torch.manual_seed(0)
    C = 6
    x = torch.randn(8, C, 10, 9)
    ref = torch.nn.BatchNorm2d(C)
    ref.weight.data.uniform_(0.5, 1.5)
    ref.bias.data.normal_()
    xa = x.clone().requires_grad_(True)
    y = ref(xa)
    gy = y.detach().sin()
    (y * gy).sum().backward()
    mean = x.double().mean((0, 2, 3))
    inv = torch.rsqrt(x.double().var((0, 2, 3), unbiased=False) + ref.eps)

    class OtherRank:
        def __init__(self):
            self.other = []

        def get_world_size(self):
            return 2

        def all_reduce(self, t):
            t.add_(self.other.pop(0))
    halves = [x[:4], x[4:]]
    out = []
    for h in range(2):
        o = halves[1 - h].double()
        fake = OtherRank()
        fake.other.append(torch.cat([o.sum((0, 2, 3)), (o * o).sum((0, 2, 3)), torch.tensor([o.numel() // C], dtype=torch.float64)]))
        dyo = gy[(1 - h) * 4:(1 - h) * 4 + 4].double()
        xhat_o = (o - mean.view(1, C, 1, 1)) * inv.view(1, C, 1, 1)
        fake.other.append(torch.cat([dyo.sum((0, 2, 3)), (dyo * xhat_o).sum((0, 2, 3))]))
        m = T.DPBatchNorm2d(C)
        m.weight.data.copy_(ref.weight.data)
        m.bias.data.copy_(ref.bias.data)
        m.dist = fake
        xh = halves[h].clone().requires_grad_(True)
        yh = m(xh)
        (yh * gy[h * 4:h * 4 + 4]).sum().backward()
        out.append((yh.detach(), xh.grad, m.weight.grad, m.bias.grad, m.running_mean, m.running_var, int(m.num_batches_tracked)))
    assert torch.allclose(torch.cat([out[0][0], out[1][0]]), y.detach(), atol=2e-6)
    assert torch.allclose(torch.cat([out[0][1], out[1][1]]), xa.grad, atol=2e-6)
    assert torch.allclose(out[0][2] + out[1][2], ref.weight.grad, atol=1e-4)
    assert torch.allclose(out[0][3] + out[1][3], ref.bias.grad, atol=1e-5)
    for h in range(2):
        assert torch.allclose(out[h][4], ref.running_mean, atol=1e-7) and torch.allclose(out[h][5], ref.running_var, atol=1e-7)
        assert out[h][6] == int(ref.num_batches_tracked) == 1
    # conversion keeps parameters, buffers and state_dict keys
    import model as M
    net = M.XiangqiNet(16, 1)
    keys = list(net.state_dict().keys())
    w = net.input_conv[1].weight
    T.convert_dp_batchnorm(net, fake)
    assert list(net.state_dict().keys()) == keys and net.input_conv[1].weight is w and isinstance(net.res_blocks[0].bn1, T.DPBatchNorm2d)


def test_training_plane_geometry_covers_every_read_of_the_kernels():
    """tnet.plane_rows / dense_rows / conv_wgrad_geometry (host side of csrc/xq_tnet.cu): for every minibatch size the plane
    tensors hold the rows the contraction kernels read -- the 11-row halo of the last 256-row work item (xq_tgemm) and the
    last slab of the weight-gradient kernel with its 8 + 12 rows of tap overhang (xq_twgrad) -- and work items tile the rows."""
    import tnet as T
    for B in list(range(1, 300)) + [511, 512, 513, 1024, 4096]:
        n_rows = B * T.BOARD_ROWS
        rows = T.plane_rows(B)
        pairs = (n_rows + T.PAIR - 1) // T.PAIR
        assert rows >= T.ROW0 + pairs * T.PAIR + 11                      # A block of the last item: [m0 - 11, m0 + 267)
        slabs, spi = T.conv_wgrad_geometry(n_rows)
        assert slabs * spi * 64 >= n_rows and slabs * 3 <= 148           # every row in a slab; one wave of (slab, tap group) items
        assert (slabs - 1) * spi * 64 < n_rows                           # no empty slab
        assert rows >= T.ROW0 + slabs * spi * 64 + 8 + 12                # B stage of the last slab: 72 rows from row k0 + 8 (dy = +1)
        assert T.ROW0 - 12 >= 0                                          # ... and from row k0 - 12 (dy = -1) of the first
        drows = T.dense_rows(B)
        dpairs = (B + T.PAIR - 1) // T.PAIR
        assert drows >= T.ROW0 + dpairs * T.PAIR + 11 and drows >= T.ROW0 + (B + 31) // 32 * 32
