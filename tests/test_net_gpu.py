"""K3 parity: the bf16 tcgen05/TMA forward against the fp32 PyTorch model (model.py:87-107).
Tolerance (BASELINE north star): policy and value within 1e-2 relative error of the fp32 model.
Layer-by-layer checks localise a failure to one kernel launch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import game
    return game.engine(0)


def make_model(C, R, seed=0):
    import torch
    import model as M
    torch.manual_seed(seed)
    m = M.XiangqiNet(C, R)
    # non-trivial BatchNorm statistics so that folding is really exercised
    g = torch.Generator().manual_seed(seed + 1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.2)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 1.5 + 0.5)
            mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) * 1.0 + 0.5)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    m.eval()
    return m


def positions(oracle, n, seed=3):
    boards, sides = oracle.random_playout_positions(seed, n)
    _, _, _, planes = oracle.movegen_batch(boards, sides, want_planes=True)
    return planes


def rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("n", [1, 7, 300])
def test_layer_by_layer(eng, oracle, n):
    import torch
    import model as M
    m = make_model(128, 2, seed=5)
    net = M.B200Net(eng, m, max_batch=n)
    x = torch.from_numpy(positions(oracle, n)).to(eng.dev)
    mg = m.to(eng.dev)
    net.load_planes(x)
    with torch.no_grad():
        ref = mg.input_conv(x)
        net.run_layer(0)
        torch.cuda.synchronize()
        got = net.planes_to_nchw(net.act[0], n)
        assert rel(got, ref) < 5e-3, ("input conv", rel(got, ref))
        # rows past the last board are never written (no halo in the plane layout: off-board taps are masked in the MMAs)
        assert float(net.act[0][:, 16 + n * 90:].float().abs().max()) == 0.0
        cur = 0
        li = 1
        for blk in mg.res_blocks:
            t, o = (cur + 1) % 3, (cur + 2) % 3
            # feed the reference the SAME (bf16-rounded) input the kernel sees, layer by layer
            xin = net.planes_to_nchw(net.act[cur], n)
            r1 = torch.relu(blk.bn1(blk.conv1(xin)))
            net.run_layer(li)
            torch.cuda.synchronize()
            g1 = net.planes_to_nchw(net.act[t], n)
            assert rel(g1, r1) < 5e-3, ("conv1", li, rel(g1, r1))
            r2 = torch.relu(blk.bn2(blk.conv2(g1)) + xin)
            net.run_layer(li + 1)
            torch.cuda.synchronize()
            g2 = net.planes_to_nchw(net.act[o], n)
            assert rel(g2, r2) < 5e-3, ("conv2+res", li + 1, rel(g2, r2))
            cur, li = o, li + 2
        trunk = net.planes_to_nchw(net.act[cur], n)
        # heads
        pf = mg.policy_head[2](mg.policy_head[1](mg.policy_head[0](trunk)))          # [n,32,10,9]
        vf = mg.value_head[2](mg.value_head[1](mg.value_head[0](trunk)))             # [n,4,10,9]
        net.run_layer(li)
        torch.cuda.synchronize()
        got_pf = net.fc_in[:, :n].permute(1, 0, 2).reshape(n, 90, 32).permute(0, 2, 1).reshape(n, 32, 10, 9).float()
        assert rel(got_pf, pf) < 5e-3, ("policy head conv", rel(got_pf, pf))
        got_vf = net.vfeat[:n].permute(0, 2, 1).reshape(n, 4, 10, 9)
        assert rel(got_vf, vf) < 5e-3, ("value head conv", rel(got_vf, vf))
        logits_ref = mg.policy_head[4](got_pf.flatten(1))
        net.run_layer(li + 1)
        torch.cuda.synchronize()
        assert rel(net.logits[:n, :8100].float(), logits_ref) < 5e-3, ("policy fc", rel(net.logits[:n, :8100].float(), logits_ref))
        v_ref = mg.value_head[7](mg.value_head[6](torch.relu(mg.value_head[4](got_vf.flatten(1)))))[:, 0]
        net.run()
        torch.cuda.synchronize()
        assert float((net.value[:n] - v_ref).abs().max()) < 2e-3, ("value mlp", float((net.value[:n] - v_ref).abs().max()))


@pytest.mark.parametrize("C,R,n", [(128, 6, 513), (256, 3, 130), (256, 2, 24), (128, 3, 40), (64, 3, 70), (192, 2, 40)])   # 64: the reference's quick preset (zero-padded to 128)
def test_end_to_end_vs_fp32(eng, oracle, C, R, n):
    import torch
    import model as M
    m = make_model(C, R, seed=11)
    net = M.B200Net(eng, m, max_batch=n)
    x = torch.from_numpy(positions(oracle, n, seed=9)).to(eng.dev)
    logits, value = net.forward_planes(x)
    torch.cuda.synchronize()
    mg = m.to(eng.dev)
    with torch.no_grad():
        lr, vr = mg(x)
    p = torch.softmax(logits[:, :8100].float(), dim=1)
    pr = torch.softmax(lr, dim=1)
    # 1e-2 relative error on policy (max-norm relative to the largest probability, and L2) and value
    assert float((p - pr).abs().max() / pr.max()) < 1e-2
    assert rel(p, pr) < 1e-2
    assert rel(logits[:, :8100].float(), lr) < 1e-2
    assert float((value - vr[:, 0]).abs().max()) < 1e-2
    # determinism: same input, same bits
    l1, v1 = logits.clone(), value.clone()
    l2, v2 = net.forward_planes(x)
    torch.cuda.synchronize()
    assert torch.equal(l2, l1) and torch.equal(v2, v1)


def test_predict_contract(eng, oracle):
    """XiangqiNet.predict keeps the reference contract (model.py:109-124)."""
    import torch
    m = make_model(128, 1, seed=2)
    state = positions(oracle, 1)[0]
    probs, value = m.predict(state)
    assert probs.shape == (8100,) and probs.dtype == np.float32 and isinstance(value, float)
    assert abs(float(probs.sum()) - 1.0) < 1e-3
    with torch.no_grad():
        lr, vr = m(torch.from_numpy(state)[None])
    pr = torch.softmax(lr, dim=1)[0].numpy()
    assert np.abs(probs - pr).max() / pr.max() < 1e-2 and abs(value - float(vr)) < 1e-2
