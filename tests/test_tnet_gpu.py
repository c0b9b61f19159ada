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
    # contraction splits: two-way sum into the row-major logits, six partial plane tensors for the input gradient
    logits2 = torch.full((B, 8320), 5.0, device=dev)
    T.tgemm(eng, fp, rows, 90, img, 1, 92, False, pairs, 64, B, out_rm=logits2, out_stride=8320, bias=bias, n_cols=N, k_splits=2)
    e1 = max(e1, _err(logits2[:B, :N], f @ w.t() + bias))
    parts = torch.zeros((6, 736, rows, 4), device=dev)
    T.tgemm(eng, dlp, rows, 254, img_t, 1, 254, True, pairs, 23, B, out=parts, out_rows=rows, out_chunks=720, k_splits=6,
            out_split_stride=736 * rows * 4)
    e2 = max(e2, _err(T.planes_to_rows(parts.sum(0), B, K), dl @ w))
    dw = torch.zeros((N, K), device=dev)
    T.dense_wgrad(eng, T.rows_to_glayout(dl, rows, 256), T.rows_to_glayout(f, rows, 96), rows, B, N, K, dw, K)
    e3 = _err(dw, dl.t() @ f)
    torch.cuda.synchronize()
    print(f"B={B}: dense fwd {e1:.2e} dgrad {e2:.2e} wgrad {e3:.2e}")
    assert max(e1, e2, e3) < TOL


# =================================================================================================
# The layers between the contractions (csrc/xq_tnet_ops.cuh) and the whole hand-written step
# =================================================================================================
def test_layout_writers_and_weight_images_are_exact(eng):
    """xq_tn_input / xq_tn_wimage / flatten / unflatten / rows_layouts / colsum move data only: bit-exact against the
    torch restatements of the layouts in tnet.py."""
    import torch
    import tnet as T
    torch.manual_seed(3)
    dev, B, L = eng.dev, 5, eng.L
    rows, drows = T.plane_rows(B), T.dense_rows(B)
    s = eng._stream
    x = _noise(B, 15, 10, 9, dev=dev)
    p, g = torch.zeros((8, rows, 4), device=dev), torch.zeros((1, rows, 32), device=dev)
    eng._check(L.xq_tn_input(eng.h, x.data_ptr(), B, 15, 4, p.data_ptr(), g.data_ptr(), rows, s()))
    assert torch.equal(p, T.to_planes(x, rows, chunks=8)) and torch.equal(g, T.to_glayout(x, rows, groups=1))
    # weight images: 3x3 conv both ways, merged heads, dense layer both ways
    w = _noise(128, 128, 3, 3, dev=dev)
    for tr in (0, 1):
        img = torch.zeros((1, 9, 4, 8, 128, 4), device=dev)
        eng._check(L.xq_tn_wimage(eng.h, w.data_ptr(), 128, 128, 9, img.data_ptr(), 4, 0, 0, tr, s()))
        assert torch.equal(img, T.weight_image(w.transpose(0, 1).contiguous() if tr else w))
    w0 = _noise(128, 15, 3, 3, dev=dev)
    img = torch.zeros((1, 9, 1, 8, 128, 4), device=dev)
    eng._check(L.xq_tn_wimage(eng.h, w0.data_ptr(), 128, 15, 9, img.data_ptr(), 1, 0, 0, 0, s()))
    assert torch.equal(img, T.weight_image(w0))
    wp, wv = _noise(32, 128, 1, 1, dev=dev), _noise(4, 128, 1, 1, dev=dev)
    wh = torch.cat([wp, wv])
    img, img_t = torch.zeros((1, 1, 4, 8, 128, 4), device=dev), torch.zeros((1, 1, 2, 8, 128, 4), device=dev)
    for wt, off in ((wp, 0), (wv, 32)):
        eng._check(L.xq_tn_wimage(eng.h, wt.data_ptr(), wt.shape[0], 128, 1, img.data_ptr(), 4, off, 0, 0, s()))
        eng._check(L.xq_tn_wimage(eng.h, wt.data_ptr(), wt.shape[0], 128, 1, img_t.data_ptr(), 2, 0, off, 1, s()))
    assert torch.equal(img, T.weight_image(wh)) and torch.equal(img_t, T.weight_image(wh.transpose(0, 1).contiguous(), img_kb=2))
    wd = _noise(8100, 2880, dev=dev)
    img, img_t = torch.zeros((64, 1, 90, 8, 128, 4), device=dev), torch.zeros((23, 1, 254, 8, 128, 4), device=dev)
    eng._check(L.xq_tn_wimage(eng.h, wd.data_ptr(), 8100, 2880, 1, img.data_ptr(), 90, 0, 0, 0, s()))
    eng._check(L.xq_tn_wimage(eng.h, wd.data_ptr(), 8100, 2880, 1, img_t.data_ptr(), 254, 0, 0, 1, s()))
    assert torch.equal(img, T.weight_image(wd, img_nt=64, img_kb=90))
    assert torch.equal(img_t, T.weight_image(wd.t().contiguous(), img_nt=23, img_kb=254))
    img2, img2_t = torch.zeros_like(img), torch.zeros_like(img_t)          # both images from one pass over the weight
    eng._check(L.xq_tn_wimage_dense2(eng.h, wd.data_ptr(), 8100, 2880, img2.data_ptr(), 90, img2_t.data_ptr(), 254, s()))
    assert torch.equal(img2, img) and torch.equal(img2_t, img_t)
    # flatten / unflatten
    a = _noise(B, 36, 10, 9, dev=dev)
    ap = T.to_planes(a, rows, chunks=16)
    fp, fg = torch.zeros((720, drows, 4), device=dev), torch.zeros((96, drows, 32), device=dev)
    eng._check(L.xq_tn_flatten(eng.h, ap.data_ptr(), rows, B, 32, fp.data_ptr(), fg.data_ptr(), drows, s()))
    flat = a[:, :32].reshape(B, 2880)
    assert torch.equal(fp, T.rows_to_planes(flat, drows, chunks=720)) and torch.equal(fg, T.rows_to_glayout(flat, drows, 96))
    back = torch.zeros((16, rows, 4), device=dev)
    eng._check(L.xq_tn_unflatten(eng.h, fp.data_ptr(), drows, B, 32, back.data_ptr(), rows, 1, 0, s()))
    assert torch.equal(T.from_planes(back, B, 32), a[:, :32]) and float(back[8:].abs().max()) == 0.0
    parts = torch.stack([fp, 2 * fp, -fp])                                   # three partial tensors, summed in order
    eng._check(L.xq_tn_unflatten(eng.h, parts.data_ptr(), drows, B, 32, back.data_ptr(), rows, 3, fp.numel(), s()))
    assert torch.equal(T.from_planes(back, B, 32), (a[:, :32] + 2 * a[:, :32]) - a[:, :32])
    # the batched image builder: the same images from one launch
    from xq_native import TnWimageItem
    ia, ib = torch.zeros((1, 9, 4, 8, 128, 4), device=dev), torch.zeros((1, 9, 4, 8, 128, 4), device=dev)
    ih, iht = torch.zeros((1, 1, 4, 8, 128, 4), device=dev), torch.zeros((1, 1, 2, 8, 128, 4), device=dev)
    i0 = torch.zeros((1, 9, 1, 8, 128, 4), device=dev)
    items = [TnWimageItem(w=w.data_ptr(), img=ia.data_ptr(), co=128, ci=128, taps=9, img_kb=4, n0=0, k0=0, transposed=0, pad_=0),
             TnWimageItem(w=w.data_ptr(), img=ib.data_ptr(), co=128, ci=128, taps=9, img_kb=4, n0=0, k0=0, transposed=1, pad_=0),
             TnWimageItem(w=w0.data_ptr(), img=i0.data_ptr(), co=128, ci=15, taps=9, img_kb=1, n0=0, k0=0, transposed=0, pad_=0)]
    for wt, off in ((wp, 0), (wv, 32)):
        items.append(TnWimageItem(w=wt.data_ptr(), img=ih.data_ptr(), co=wt.shape[0], ci=128, taps=1, img_kb=4, n0=off, k0=0, transposed=0, pad_=0))
        items.append(TnWimageItem(w=wt.data_ptr(), img=iht.data_ptr(), co=wt.shape[0], ci=128, taps=1, img_kb=2, n0=0, k0=off, transposed=1, pad_=0))
    arr = (TnWimageItem * len(items))(*items)
    eng._check(L.xq_tn_wimage_batch(eng.h, arr, len(items), s()))
    assert torch.equal(ia, T.weight_image(w)) and torch.equal(ib, T.weight_image(w.transpose(0, 1).contiguous()))
    assert torch.equal(i0, T.weight_image(w0)) and torch.equal(ih, T.weight_image(wh))
    assert torch.equal(iht, T.weight_image(wh.transpose(0, 1).contiguous(), img_kb=2))
    # row-major -> dense layouts, column sums
    m = _noise(B, 8100, dev=dev)
    dp, dg = torch.zeros((2048, drows, 4), device=dev), torch.zeros((256, drows, 32), device=dev)
    eng._check(L.xq_tn_rows_layouts(eng.h, m.data_ptr(), 8100, B, 8100, dp.data_ptr(), dg.data_ptr(), drows, s()))
    assert torch.equal(dp, T.rows_to_planes(m, drows, chunks=2048)) and torch.equal(dg, T.rows_to_glayout(m, drows, 256))
    cs = torch.zeros(8100, device=dev)
    eng._check(L.xq_tn_colsum(eng.h, m.data_ptr(), 8100, B, 8100, cs.data_ptr(), s()))
    assert torch.allclose(cs, m.sum(0), rtol=1e-5, atol=1e-5)
    # slab reduction
    ws = _noise(7, 9, 128, 128, dev=dev)
    dw = torch.zeros((128, 128, 3, 3), device=dev)
    eng._check(L.xq_tn_wgrad_reduce(eng.h, ws.data_ptr(), 7, 9 * 128 * 128, 9, 128, 128, 128, 0, 0, dw.data_ptr(), 128, 0, 0, s()))
    assert torch.allclose(dw, ws.sum(0).reshape(3, 3, 128, 128).permute(2, 3, 0, 1), rtol=1e-5, atol=1e-5)
    wst = _noise(7, 1, 128, 64, dev=dev)
    dwp, dwv = torch.zeros((32, 128), device=dev), torch.zeros((4, 128), device=dev)
    eng._check(L.xq_tn_wgrad_reduce(eng.h, wst.data_ptr(), 7, 128 * 64, 1, 64, 128, 32, 0, 1, dwp.data_ptr(), 128, 0, 0, s()))
    eng._check(L.xq_tn_wgrad_reduce(eng.h, wst.data_ptr(), 7, 128 * 64, 1, 64, 128, 4, 32, 1, dwv.data_ptr(), 128, 0, 0, s()))
    tot = wst.sum(0)[0].t()
    assert torch.allclose(dwp, tot[:32], rtol=1e-5, atol=1e-5) and torch.allclose(dwv, tot[32:36], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,with_res", [(3, False), (19, True)])
def test_plane_batchnorm_forward_backward_match_torch(eng, B, with_res):
    """xq_tn_bn_forward / xq_tn_bn_backward == relu(BatchNorm2d(train)(y) [+ res]) and its autograd gradients."""
    import ctypes as C
    import torch
    import tnet as T
    from xq_native import TnBnDesc, TnBnBwdDesc
    torch.manual_seed(B)
    dev, Cn = eng.dev, 128
    rows = T.plane_rows(B)
    y = (_noise(B, Cn, 10, 9, dev=dev) * 2 + 0.5).requires_grad_(True)
    res = _noise(B, Cn, 10, 9, dev=dev).requires_grad_(True) if with_res else None
    bn = torch.nn.BatchNorm2d(Cn).to(dev)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(Cn, device=dev) + 0.5)
        bn.bias.copy_(_noise(Cn, dev=dev) * 0.3)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    out = torch.relu(bn(y) + res) if with_res else torch.relu(bn(y))
    dout = _noise(B, Cn, 10, 9, dev=dev)
    out.backward(dout)
    # ours: garbage in the pad cells of y and dout must not matter
    yp = T.to_planes(y.detach(), rows)
    mask_pad = (T.to_planes(torch.ones_like(dout), rows) == 0)
    mask_pad[:, :T.ROW0] = False
    mask_pad[:, T.ROW0 + B * T.BOARD_ROWS:] = False
    yp[mask_pad] = 7.0
    doutp = T.to_planes(dout, rows)
    doutp[mask_pad] = -3.0
    resp = T.to_planes(res.detach(), rows) if with_res else None
    outp, outg = torch.full((Cn // 4, rows, 4), 9.0, device=dev), torch.zeros((Cn // 32, rows, 32), device=dev)
    partial = torch.zeros((Cn // 4 + 2, 16, 8), dtype=torch.float64, device=dev)
    save = torch.zeros((2, Cn), device=dev)
    rm, rv = rm0.clone(), rv0.clone()
    d = TnBnDesc(y=yp.data_ptr(), res=None if resp is None else resp.data_ptr(), out=outp.data_ptr(), out_g=outg.data_ptr(), rows=rows,
                 n_boards=B, chunk0=0, n_channels=Cn, relu=1, partial=partial.data_ptr(), gamma=bn.weight.data_ptr(),
                 beta=bn.bias.data_ptr(), running_mean=rm.data_ptr(), running_var=rv.data_ptr(), save=save.data_ptr(), eps=bn.eps,
                 momentum=bn.momentum)
    eng._check(eng.L.xq_tn_bn_forward(eng.h, C.byref(d), eng._stream()))
    got = T.from_planes(outp, B, Cn)
    assert torch.allclose(got, out.detach(), atol=2e-5, rtol=1e-5)
    assert torch.equal(T.from_glayout(outg, B, Cn), got)
    inner = outp[:, T.ROW0:T.ROW0 + B * T.BOARD_ROWS]
    assert float(inner[mask_pad[:, T.ROW0:T.ROW0 + B * T.BOARD_ROWS]].abs().max()) == 0.0     # pad cells are written as zeros
    assert torch.allclose(rm, bn.running_mean, atol=1e-6) and torch.allclose(rv, bn.running_var, atol=1e-5)
    dyp, dyg, dskip = torch.zeros_like(outp), torch.zeros_like(outg), torch.zeros_like(outp)
    dgamma, dbeta = torch.zeros(Cn, device=dev), torch.zeros(Cn, device=dev)
    d2 = TnBnBwdDesc(dout=doutp.data_ptr(), act=outp.data_ptr(), y=yp.data_ptr(), rows=rows, n_boards=B, chunk0=0, n_channels=Cn, relu=1,
                     save=save.data_ptr(), partial=partial.data_ptr(), gamma=bn.weight.data_ptr(), dgamma=dgamma.data_ptr(),
                     dbeta=dbeta.data_ptr(), dy=dyp.data_ptr(), dy_g=dyg.data_ptr(), dskip=dskip.data_ptr() if with_res else None)
    eng._check(eng.L.xq_tn_bn_backward(eng.h, C.byref(d2), eng._stream()))
    got_dy = T.from_planes(dyp, B, Cn)
    assert torch.allclose(got_dy, y.grad, atol=2e-5, rtol=1e-4)
    assert torch.equal(T.from_glayout(dyg, B, Cn), got_dy) and float(dyp[mask_pad].abs().max()) == 0.0
    assert torch.allclose(dgamma, bn.weight.grad, atol=2e-3, rtol=1e-4) and torch.allclose(dbeta, bn.bias.grad, atol=2e-3, rtol=1e-4)
    if with_res:
        assert torch.allclose(T.from_planes(dskip, B, Cn), res.grad, atol=1e-6)


def test_value_head_dense_layers_match_torch(eng):
    import torch
    import tnet as T
    torch.manual_seed(11)
    dev, B, L = eng.dev, 13, eng.L
    rows = T.plane_rows(B)
    a = torch.relu(_noise(B, 36, 10, 9, dev=dev))
    f = a[:, 32:36].reshape(B, 360).clone().requires_grad_(True)
    l1, l2 = torch.nn.Linear(360, 128).to(dev), torch.nn.Linear(128, 1).to(dev)
    v = torch.tanh(l2(torch.relu(l1(f)))).reshape(-1)
    gv = _noise(B, dev=dev)
    v.backward(gv)
    ap = T.to_planes(a, rows, chunks=16)
    h, vo = torch.zeros((B, 128), device=dev), torch.zeros(B, device=dev)
    s = eng._stream
    eng._check(L.xq_tn_value_forward(eng.h, ap.data_ptr(), rows, 8, B, l1.weight.data_ptr(), l1.bias.data_ptr(), l2.weight.data_ptr(),
                                     l2.bias.data_ptr(), h.data_ptr(), vo.data_ptr(), s()))
    assert torch.allclose(vo, v.detach(), atol=1e-5)
    dh, dpre, dact = torch.zeros((B, 128), device=dev), torch.zeros(B, device=dev), torch.zeros((16, rows, 4), device=dev)
    g = [torch.zeros_like(t) for t in (l1.weight, l1.bias, l2.weight, l2.bias)]
    eng._check(L.xq_tn_value_backward(eng.h, ap.data_ptr(), rows, 8, B, l1.weight.data_ptr(), l2.weight.data_ptr(), h.data_ptr(), vo.data_ptr(),
                                      gv.data_ptr(), dh.data_ptr(), dpre.data_ptr(), dact.data_ptr(), g[0].data_ptr(), g[1].data_ptr(),
                                      g[2].data_ptr(), g[3].data_ptr(), s()))
    for got, want in zip(g, (l1.weight.grad, l1.bias.grad, l2.weight.grad, l2.bias.grad)):
        assert torch.allclose(got.reshape(-1), want.reshape(-1), atol=1e-5, rtol=1e-4)
    assert torch.allclose(T.from_planes(dact, B, 36)[:, 32:36].reshape(B, 360), f.grad, atol=1e-6, rtol=1e-4)


@pytest.mark.parametrize("B,blocks,width", [(6, 1, 128), (64, 2, 128), (256, 6, 128), (10, 1, 256), (12, 1, 64), (7, 1, 96)])
def test_hand_step_matches_torch_autograd(eng, B, blocks, width):
    """HandStep (forward + loss + backward of XiangqiNet on the kernels) against the torch module on the same weights and
    minibatch (train.py:397-413): logits / value / losses and every parameter gradient."""
    import torch
    import model as M
    import replay
    import tnet as T
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(100 + B)
    dev = eng.dev
    net = M.XiangqiNet(width, blocks).to(dev).train()
    ref = M.XiangqiNet(width, blocks).to(dev).train()
    ref.load_state_dict(net.state_dict())
    for p in net.parameters():
        p.grad = torch.full_like(p, 123.0)                  # a step must ASSIGN every gradient
    states = (_noise(B, 15, 10, 9, dev=dev) > 0.8).float()
    act = torch.zeros((B, 128), dtype=torch.int16, device=dev)
    prob = torch.zeros((B, 128), device=dev)
    n = torch.randint(1, 40, (B,), dtype=torch.int32, device=dev)
    for i in range(B):
        k = int(n[i])
        act[i, :k] = torch.randperm(8100, device=dev)[:k].to(torch.int16)
        pr = torch.rand(k, device=dev)
        prob[i, :k] = pr / pr.sum()
    z = torch.randint(-1, 2, (B,), device=dev).float()
    hs = T.HandStep(eng, net)
    pl, vl = hs.step(states, act, prob, n, z, 1.0 / B)
    b = hs.buffers(B)
    # Reference: the torch module on the same weights, with the ReLU decisions of OUR forward (mask = our activation > 0).
    # tf32 forward errors (~1e-3) flip the few ReLU units whose pre-activation is that close to zero; a flipped fraction f of
    # the terms behind a random-sign sum moves it by ~sqrt(f) of its size, several per cent after 13 layers, which says
    # nothing about the backward kernels.  With the masks pinned the comparison is tight; the unpinned module bounds the rest.
    C = width
    masks = [(T.from_planes(b.Ap[i], B, C) > 0).float() for i in range(hs.L)]
    ah = T.from_planes(b.Ah, B, 36)
    mask_p, mask_v, mask_h = (ah[:, :32] > 0).float(), (ah[:, 32:36] > 0).float(), (b.h > 0).float()

    def pinned(m):
        x = m.input_conv[1](m.input_conv[0](states)) * masks[0]
        for i, blk in enumerate(m.res_blocks):
            h = blk.bn1(blk.conv1(x)) * masks[2 * i + 1]
            x = (blk.bn2(blk.conv2(h)) + x) * masks[2 * i + 2]
        p = m.policy_head[1](m.policy_head[0](x)) * mask_p
        v = m.value_head[1](m.value_head[0](x)) * mask_v
        hid = m.value_head[4](v.flatten(1)) * mask_h
        return m.policy_head[4](p.flatten(1)), torch.tanh(m.value_head[6](hid))

    logits, values = pinned(ref)
    rpl, rvl = replay.policy_value_loss(eng, logits, values, (act, prob, n), z)
    (rpl + rvl).backward()
    torch.cuda.synchronize()
    e_l, e_v = _err(b.logits, logits.detach()), float((b.v - values.detach().reshape(-1)).abs().max())
    print(f"B={B}: logits {e_l:.2e} value {e_v:.2e} losses {float(pl):.5f}/{float(rpl.detach()):.5f} {float(vl):.5f}/{float(rvl.detach()):.5f}")
    assert e_l < 1e-2 and e_v < 1e-2
    rpl_f, rvl_f = float(rpl.detach()), float(rvl.detach())
    assert abs(float(pl) - rpl_f) < 2e-3 * abs(rpl_f) and abs(float(vl) - rvl_f) < 5e-3 * abs(rvl_f) + 1e-4
    worst = worst2 = 0.0
    report = []
    for (name, p), q in zip(net.named_parameters(), ref.parameters()):
        err = _err(p.grad, q.grad)
        err2 = float((p.grad - q.grad).norm() / q.grad.norm().clamp_min(1e-20))
        report.append(f"{name}: max {err:.2e} l2 {err2:.2e}")
        worst, worst2 = max(worst, err), max(worst2, err2)
    print("\n".join(report))
    print(f"worst gradient error (ReLU decisions pinned): max {worst:.2e}, l2 {worst2:.2e}")
    assert worst2 < 1e-2 and worst < 2e-2
    for (name, p), (_, q) in zip(net.named_buffers(), ref.named_buffers()):
        if "num_batches" not in name:
            assert torch.allclose(p, q, atol=1e-4, rtol=1e-3), name
    hs.sync_counters()
    assert int(net.input_conv[1].num_batches_tracked) == 1
    # the same step again (now replayed from the CUDA graph captured after the first one): bit-identical losses and gradients
    first = [p.grad.clone() for p in net.parameters()]
    l_first = (float(pl), float(vl))
    for p in net.parameters():
        p.grad.fill_(-7.0)
    pl2, vl2 = hs.step(states, act, prob, n, z, 1.0 / B)
    torch.cuda.synchronize()
    assert hs.buffers(B).graph is not None
    assert (float(pl2), float(vl2)) == l_first
    for (name, p), q in zip(net.named_parameters(), first):
        assert torch.equal(p.grad, q), name
    # the unpinned torch module: same outputs, gradients within the ReLU-flip noise
    free = M.XiangqiNet(width, blocks).to(dev).train()
    free.load_state_dict(ref.state_dict())
    logits, values = free(states)
    fpl, fvl = replay.policy_value_loss(eng, logits, values, (act, prob, n), z)
    (fpl + fvl).backward()
    assert _err(b.logits, logits.detach()) < 1e-2
    cos = []
    for p, q in zip(net.parameters(), free.parameters()):
        cos.append(float((p.grad * q.grad).sum() / (p.grad.norm() * q.grad.norm()).clamp_min(1e-30)))
    print(f"unpinned module: smallest gradient cosine {min(cos):.5f}")
    assert min(cos) > 0.98


def test_hand_step_buffer_sets_are_bounded(eng):
    """The last minibatch of an epoch changes size from iteration to iteration: HandStep keeps three buffer sets (and their
    captured graphs) and a size that comes back after its set was dropped gives the same gradients again."""
    import torch
    import model as M
    import tnet as T
    torch.manual_seed(9)
    dev = eng.dev
    net = M.XiangqiNet(64, 1).to(dev).train()
    for p in net.parameters():
        p.grad = torch.zeros_like(p)
    hs = T.HandStep(eng, net)

    def batch(B):
        g = torch.Generator(device=dev).manual_seed(B)
        st = (torch.rand(B, 15, 10, 9, device=dev, generator=g) > 0.8).float()
        act = torch.zeros((B, 128), dtype=torch.int16, device=dev)
        prob = torch.zeros((B, 128), device=dev)
        act[:, :2] = torch.tensor([17, 4000], dtype=torch.int16, device=dev)
        prob[:, :2] = torch.tensor([0.75, 0.25], device=dev)
        n = torch.full((B,), 2, dtype=torch.int32, device=dev)
        z = torch.ones(B, device=dev)
        return st, act, prob, n, z

    first = None
    for B in (4, 5, 4, 6, 7, 8, 4):
        hs.step(*batch(B), 1.0 / B)
        torch.cuda.synchronize()
        assert len(hs._bufs) <= T.HandStep.MAX_BUFFER_SETS
        if B == 4:
            g = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).clone()
            assert first is None or torch.equal(g, first)       # eager, replayed, and rebuilt after eviction: the same gradients
            first = g
    assert list(hs._bufs) == [7, 8, 4]
