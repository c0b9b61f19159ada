"""f1: the hand-written training-step kernels (csrc/xq_tnet.cu) against fp32 torch on the same inputs.

Tolerance: tf32 products (10-bit mantissa, the hardware truncates the fp32 operands) with fp32 accumulation: every
result within 4e-3 of the fp32 reference, relative to the largest magnitude of the tensor (train.py:397-423 runs the
same contractions in fp32 on the reference's CPU; the whole-step tests in test_train_gpu.py hold the reference-run tolerances)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "xiangqi-alphazero_b200"))

pytestmark = pytest.mark.gpu
TOL = 4e-3


@pytest.fixture(scope="module")
def eng():
    import game
    return game.engine(0)


def _err(got, want):
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-20))


def _noise(*shape, dev):
    import torch
    return torch.randn(*shape, device=dev)


@pytest.mark.parametrize("B", [3, 37])
def test_conv_fprop_dgrad_wgrad_match_torch(eng, B):
    import torch
    import torch.nn.functional as F
    import tnet as T
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(B)
    dev = eng.dev
    x = _noise(B, 128, 10, 9, dev=dev)
    w = _noise(128, 128, 3, 3, dev=dev) * 0.05
    dy = _noise(B, 128, 10, 9, dev=dev)
    rows = T.plane_rows(B)
    pairs = (B * T.BOARD_ROWS + 255) // 256
    xp, dyp, img = T.to_planes(x, rows), T.to_planes(dy, rows), T.weight_image(w)
    assert torch.equal(T.from_planes(xp, B, 128), x)
    # fprop
    yp = torch.zeros_like(xp)
    T.tgemm(eng, xp, rows, 4, img, 9, 4, False, pairs, 1, B * T.BOARD_ROWS, out=yp, out_rows=rows, out_chunks=32)
    want = F.conv2d(x, w, padding=1)
    e1 = _err(T.from_planes(yp, B, 128), want)
    # fprop + residual
    yr = torch.zeros_like(xp)
    T.tgemm(eng, xp, rows, 4, img, 9, 4, False, pairs, 1, B * T.BOARD_ROWS, out=yr, out_rows=rows, out_chunks=32, residual=dyp)
    e1r = _err(T.from_planes(yr, B, 128), want + dy)
    # dgrad: fprop over dY with the image of the transposed weights and mirrored taps
    dxp = torch.zeros_like(xp)
    T.tgemm(eng, dyp, rows, 4, T.weight_image(w.transpose(0, 1)), 9, 4, True, pairs, 1, B * T.BOARD_ROWS, out=dxp, out_rows=rows, out_chunks=32)
    want_dx = torch.nn.grad.conv2d_input(x.shape, w, dy, padding=1)
    e2 = _err(T.from_planes(dxp, B, 128), want_dx)
    # wgrad: partial sums per slab, summed here
    slabs, _ = T.conv_wgrad_geometry(B * T.BOARD_ROWS)
    ws = torch.zeros((slabs, 9, 128, 128), device=dev)
    xg, dyg = T.to_glayout(x, rows), T.to_glayout(dy, rows)
    assert torch.equal(T.from_glayout(xg, B, 128), x)
    T.conv_wgrad(eng, dyg, xg, rows, B * T.BOARD_ROWS, 4, 9, ws)
    want_dw = torch.nn.grad.conv2d_weight(x, w.shape, dy, padding=1)
    got_dw = ws.sum(0).reshape(3, 3, 128, 128).permute(2, 3, 0, 1)
    e3 = _err(got_dw, want_dw)
    torch.cuda.synchronize()
    print(f"B={B}: fprop {e1:.2e} (+res {e1r:.2e}) dgrad {e2:.2e} wgrad {e3:.2e}")
    assert e1 < TOL and e1r < TOL and e2 < TOL and e3 < TOL


def test_narrow_and_1x1_layers(eng):
    """Input conv (15 planes in 8 chunks, wgrad with N = 32), 1x1 heads conv (36 outputs kept) with its dgrad and wgrad."""
    import torch
    import torch.nn.functional as F
    import tnet as T
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(5)
    dev, B = eng.dev, 9
    rows = T.plane_rows(B)
    pairs = (B * T.BOARD_ROWS + 255) // 256
    n_rows = B * T.BOARD_ROWS
    x0 = (_noise(B, 15, 10, 9, dev=dev) > 0.5).float()
    w0 = _noise(128, 15, 3, 3, dev=dev) * 0.1
    dy = _noise(B, 128, 10, 9, dev=dev)
    x0p, dyp = T.to_planes(x0, rows, chunks=8), T.to_planes(dy, rows)
    yp = torch.zeros((32, rows, 4), device=dev)
    T.tgemm(eng, x0p, rows, 1, T.weight_image(w0), 9, 1, False, pairs, 1, n_rows, out=yp, out_rows=rows, out_chunks=32)
    e1 = _err(T.from_planes(yp, B, 128), F.conv2d(x0, w0, padding=1))
    slabs, _ = T.conv_wgrad_geometry(n_rows)
    ws = torch.zeros((slabs, 9, 128, 32), device=dev)
    T.conv_wgrad(eng, T.to_glayout(dy, rows), T.to_glayout(x0, rows, groups=1), rows, n_rows, 1, 9, ws)
    got = ws.sum(0).reshape(3, 3, 128, 32).permute(2, 3, 0, 1)[:, :15]
    e2 = _err(got, torch.nn.grad.conv2d_weight(x0, w0.shape, dy, padding=1))
    # heads: 128 -> 36 channels, 1x1
    x = _noise(B, 128, 10, 9, dev=dev)
    wh = _noise(36, 128, 1, 1, dev=dev) * 0.1
    dyh = _noise(B, 36, 10, 9, dev=dev)
    xp = T.to_planes(x, rows)
    imgh = T.weight_image(wh)
    yh = torch.zeros((16, rows, 4), device=dev)
    T.tgemm(eng, xp, rows, 4, imgh, 1, 4, False, pairs, 1, n_rows, out=yh, out_rows=rows, out_chunks=9)
    e3 = _err(T.from_planes(yh, B, 36), F.conv2d(x, wh))
    assert float(yh[9:].abs().max()) == 0.0
    dyhp = T.to_planes(dyh, rows, chunks=16)
    dxp = torch.zeros_like(xp)
    T.tgemm(eng, dyhp, rows, 2, T.weight_image(wh.transpose(0, 1), img_kb=2), 1, 2, True, pairs, 1, n_rows, out=dxp, out_rows=rows, out_chunks=32)
    e4 = _err(T.from_planes(dxp, B, 128), torch.nn.grad.conv2d_input(x.shape, wh, dyh))
    ws = torch.zeros((slabs, 1, 128, 64), device=dev)
    T.conv_wgrad(eng, T.to_glayout(x, rows), T.to_glayout(dyh, rows, groups=2), rows, n_rows, 2, 1, ws)   # M = input channel, N = head channel
    e5 = _err(ws.sum(0)[0].t()[:36], torch.nn.grad.conv2d_weight(x, wh.shape, dyh).reshape(36, 128))
    torch.cuda.synchronize()
    print(f"input fprop {e1:.2e} wgrad {e2:.2e}; heads fprop {e3:.2e} dgrad {e4:.2e} wgrad {e5:.2e}")
    assert max(e1, e2, e3, e4, e5) < TOL


@pytest.mark.parametrize("B", [32, 256])
def test_dense_layer_forward_dgrad_wgrad(eng, B):
    """Policy FC 2880 -> 8100 (model.py:64-70): logits row-major with bias, input gradient as planes (transposed image),
    weight gradient straight into the [8100][2880] parameter layout."""
    import torch
    import tnet as T
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(B)
    dev = eng.dev
    K, N = 2880, 8100
    f = _noise(B, K, dev=dev)
    w = _noise(N, K, dev=dev) * 0.02
    bias = _noise(N, dev=dev)
    dl = _noise(B, N, dev=dev)
    img = T.weight_image(w, img_nt=64, img_kb=92)
    rows = T.ROW0 + (B + 255) // 256 * 256 + 192
    fp = T.rows_to_planes(f, rows, chunks=768)
    dlp = T.rows_to_planes(dl, rows, chunks=2048)
    pairs = (B + 255) // 256
    logits = torch.zeros((B, 8320), device=dev)
    T.tgemm(eng, fp, rows, 90, img, 1, 92, False, pairs, 64, B, out_rm=logits, out_stride=8320, bias=bias, n_cols=N)
    e1 = _err(logits[:, :N], f @ w.t() + bias)
    assert float(logits[:, N:].abs().max()) == 0.0
    dfp = torch.zeros((736, rows, 4), device=dev)
    img_t = T.weight_image(w.t().contiguous(), img_nt=23, img_kb=254)
    T.tgemm(eng, dlp, rows, 254, img_t, 1, 254, True, pairs, 23, B, out=dfp, out_rows=rows, out_chunks=720)
    e2 = _err(T.planes_to_rows(dfp, B, K), dl @ w)
    dw = torch.zeros((N, K), device=dev)
    T.dense_wgrad(eng, T.rows_to_glayout(dl, rows, 256), T.rows_to_glayout(f, rows, 96), rows, B, N, K, dw, K)
    e3 = _err(dw, dl.t() @ f)
    torch.cuda.synchronize()
    print(f"B={B}: dense fwd {e1:.2e} dgrad {e2:.2e} wgrad {e3:.2e}")
    assert max(e1, e2, e3) < TOL
