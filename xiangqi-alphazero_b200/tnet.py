"""Host side of the hand-written training step (csrc/xq_tnet.cu, f1 of SURVEY 8(f)): `HandStep` runs forward + loss +
backward of XiangqiNet (train.py:397-423 over model.py:39-107) on the kernels of libxq_b200.so and replays the ~140 launches
of a step from a CUDA graph.  This module holds the training plane layout, the weight images, the descriptors of the two
tcgen05 tf32 kernels (xq_tgemm: fprop / dgrad, xq_twgrad: weight gradients), torch restatements of the layouts for the tests,
and the orchestration of the step.

Training plane layout (include/xq_b200.h): float32 X[C/4][rows][4]; board b, cell (r, c) at row ROW0 + b*110 + r*10 + c.
Every board row carries a zero pad column (c = 9) and every board a zero pad row (r = 10): a 3x3 tap (dy, dx) is a row shift
of dy*10+dx and the pad cells are conv2d's zero padding.  G layout: float32 G[C/32][rows][32], the 32-byte units of row r
XOR-swizzled by r & 3 -- the form in which the tf32 tensor core takes the MN-major operands of a weight gradient."""
import ctypes as C
import os

import torch

import xq_native
from xq_native import TGemmDesc, TWgradDesc

ROW0 = 16
BOARD_ROWS = 110
PAIR = 256


def plane_rows(n_boards: int) -> int:
    """Rows of a plane tensor for n_boards boards: front padding, whole 256-row work items, tail padding (the wgrad slabs and
    the 11-row halo of the last tile read past the last board; those rows stay zero)."""
    n_rows = n_boards * BOARD_ROWS
    slabs, spi = conv_wgrad_geometry(n_rows)
    return ROW0 + max((n_rows + PAIR - 1) // PAIR * PAIR + 192, slabs * spi * 64 + 32)


def dense_rows(n_boards: int) -> int:
    """Rows of a dense tensor (one row per board)."""
    return ROW0 + (n_boards + PAIR - 1) // PAIR * PAIR + 192


def to_planes(x: torch.Tensor, rows: int = None, chunks: int = None) -> torch.Tensor:
    """[B, C, 10, 9] -> planes [C/4][rows][4] (zero pad cells)."""
    B, Cn = x.shape[0], x.shape[1]
    ch = (Cn + 3) // 4 if chunks is None else chunks
    rows = plane_rows(B) if rows is None else rows
    p = torch.zeros((ch, rows, 4), dtype=torch.float32, device=x.device)
    t = torch.zeros((B, 11, 10, ch * 4), dtype=torch.float32, device=x.device)
    t[:, :10, :9, :Cn] = x.permute(0, 2, 3, 1)
    p[:, ROW0:ROW0 + B * BOARD_ROWS] = t.reshape(B * BOARD_ROWS, ch, 4).permute(1, 0, 2)
    return p


def from_planes(p: torch.Tensor, B: int, Cn: int) -> torch.Tensor:
    """planes -> [B, C, 10, 9] (pad cells dropped)."""
    ch = p.shape[0]
    t = p[:, ROW0:ROW0 + B * BOARD_ROWS].permute(1, 0, 2).reshape(B, 11, 10, ch * 4)
    return t[:, :10, :9, :Cn].permute(0, 3, 1, 2).contiguous()


def rows_to_planes(x: torch.Tensor, rows: int = None, chunks: int = None) -> torch.Tensor:
    """[B, K] (one row per board, dense layers) -> planes [K/4][rows][4]."""
    B, K = x.shape
    ch = (K + 3) // 4 if chunks is None else chunks
    rows = ROW0 + (B + PAIR - 1) // PAIR * PAIR + 192 if rows is None else rows
    p = torch.zeros((ch, rows, 4), dtype=torch.float32, device=x.device)
    t = torch.zeros((B, ch * 4), dtype=torch.float32, device=x.device)
    t[:, :K] = x
    p[:, ROW0:ROW0 + B] = t.reshape(B, ch, 4).permute(1, 0, 2)
    return p


def planes_to_rows(p: torch.Tensor, B: int, K: int) -> torch.Tensor:
    return p[:, ROW0:ROW0 + B].permute(1, 0, 2).reshape(B, -1)[:, :K].contiguous()


def weight_image(w: torch.Tensor, img_nt: int = None, img_kb: int = None) -> torch.Tensor:
    """[Co][Ci][kh][kw] (or [N][K] for a dense layer) -> image [img_nt][taps][img_kb][8 chunks][128 n][4 k] float32, tap = kh*3+kw."""
    if w.dim() == 2:
        w = w[:, :, None, None]
    co, ci, kh, kw = w.shape
    nt = (co + 127) // 128 if img_nt is None else img_nt
    kb = (ci + 31) // 32 if img_kb is None else img_kb
    wp = torch.zeros((nt * 128, kb * 32, kh * kw), dtype=torch.float32, device=w.device)
    wp[:co, :ci] = w.reshape(co, ci, kh * kw)
    x = wp.reshape(nt, 128, kb, 8, 4, kh * kw).permute(0, 5, 2, 3, 1, 4).contiguous()      # [nt][tap][kb][chunk][n][4]
    return x


def _swizzle_g(t: torch.Tensor) -> torch.Tensor:
    """[groups][rows][4 units][8] -> the same with unit u of row r stored at position u ^ (r & 3)."""
    rows = t.shape[1]
    r = torch.arange(rows, device=t.device) & 3
    u = torch.arange(4, device=t.device)
    src = (u[None, :] ^ r[:, None])                                       # stored position p holds unit p ^ (r & 3)
    return torch.gather(t, 2, src[None, :, :, None].expand(t.shape[0], rows, 4, 8))


def to_glayout(x: torch.Tensor, rows: int = None, groups: int = None) -> torch.Tensor:
    """[B, C, 10, 9] -> G layout [C/32][rows][32] (the wgrad operand form, csrc/xq_tmma.cuh): zero pad cells, 32-byte units of
    row r XORed with r & 3."""
    B, Cn = x.shape[0], x.shape[1]
    g = (Cn + 31) // 32 if groups is None else groups
    rows = plane_rows(B) if rows is None else rows
    t = torch.zeros((B, 11, 10, g * 32), dtype=torch.float32, device=x.device)
    t[:, :10, :9, :Cn] = x.permute(0, 2, 3, 1)
    p = torch.zeros((g, rows, 4, 8), dtype=torch.float32, device=x.device)
    p[:, ROW0:ROW0 + B * BOARD_ROWS] = t.reshape(B * BOARD_ROWS, g, 4, 8).permute(1, 0, 2, 3)
    return _swizzle_g(p).reshape(g, rows, 32).contiguous()


def rows_to_glayout(x: torch.Tensor, rows: int, groups: int) -> torch.Tensor:
    """[B, K] (one row per board) -> G layout [groups][rows][32]."""
    B, K = x.shape
    t = torch.zeros((B, groups * 32), dtype=torch.float32, device=x.device)
    t[:, :K] = x
    p = torch.zeros((groups, rows, 4, 8), dtype=torch.float32, device=x.device)
    p[:, ROW0:ROW0 + B] = t.reshape(B, groups, 4, 8).permute(1, 0, 2, 3)
    return _swizzle_g(p).reshape(groups, rows, 32).contiguous()


def from_glayout(p: torch.Tensor, B: int, Cn: int) -> torch.Tensor:
    """G layout -> [B, C, 10, 9] (the swizzle is an involution)."""
    g, rows = p.shape[0], p.shape[1]
    t = _swizzle_g(p.reshape(g, rows, 4, 8))[:, ROW0:ROW0 + B * BOARD_ROWS].permute(1, 0, 2, 3).reshape(B, 11, 10, g * 32)
    return t[:, :10, :9, :Cn].permute(0, 3, 1, 2).contiguous()


def tgemm(eng, a, a_rows, kblocks, w_img, ntaps, img_kb, dgrad, m_pairs, n_tiles, m_rows, out=None, out_rows=0, out_chunks=0,
          residual=None, out_rm=None, out_stride=0, bias=None, n_cols=0, k_splits=1, out_split_stride=0):
    """fprop (w_img = image of the weights) or dgrad (w_img = image of the TRANSPOSED weights, taps mirrored).  k_splits > 1
    deals the contraction blocks to several work items per output tile (partial planes out_split_stride floats apart, or a
    two-way sum into the row-major output)."""
    d = TGemmDesc(a=a.data_ptr(), a_rows=a_rows, a_row0=ROW0, w=w_img.data_ptr(), kblocks=kblocks, ntaps=ntaps, img_kb=img_kb,
                  b_mn=0, shift_sign=-1 if dgrad else 1, m_pairs=m_pairs, n_tiles=n_tiles, out_chunks=out_chunks,
                  m_rows=m_rows, out=None if out is None else out.data_ptr(), out_rows=out_rows, out_row0=ROW0,
                  residual=None if residual is None else residual.data_ptr(), out_rm=None if out_rm is None else out_rm.data_ptr(),
                  out_stride=out_stride, bias=None if bias is None else bias.data_ptr(), n_cols=n_cols, k_splits=k_splits,
                  out_split_stride=out_split_stride)
    eng._check(eng.L.xq_tgemm(eng.h, C.byref(d), eng._stream()))


def conv_wgrad_geometry(n_rows: int, sm_count: int = 148):
    """(slabs, stages per item) of the conv weight-gradient launch: 64-row stages, 3 tap groups per slab, about one item per SM."""
    kr = 64
    slabs = max(1, min(sm_count // 3, (n_rows + kr - 1) // kr))
    spi = (n_rows + slabs * kr - 1) // (slabs * kr)
    slabs = (n_rows + spi * kr - 1) // (spi * kr)
    return slabs, spi


def conv_wgrad(eng, dy_g, x_g, rows, n_rows, nbg, ntaps, ws, a_group0=0, b_group0=0):
    """ws[slab][tap][128][32*nbg] = partial sums of dY[row][m] * X[row + shift(tap)][n] over the slab's rows (G-layout operands);
    m = channels [32*a_group0, +128) of dy_g, n = channels [32*b_group0, +32*nbg) of x_g."""
    slabs, spi = conv_wgrad_geometry(n_rows)
    N = 32 * nbg
    kr = 64
    if ntaps == 9:
        groups, tpg, brs = 3, 3, kr + 8
        lo = [-12, -4, 8, 0]                                   # floor4(dy*10 - 1)
        off = [0] * 16
        for g in range(3):
            for t in range(3):
                off[g * 4 + t] = ((g - 1) * 10 + (t - 1) - lo[g]) * 128
    else:
        groups, tpg, brs, lo, off = 1, 1, kr, [0, 0, 0, 0], [0] * 16
    d = TWgradDesc(a=dy_g.data_ptr(), b=x_g.data_ptr(), a_rows=rows, a_row0=ROW0, b_rows=rows, b_row0=ROW0, a_group0=a_group0,
                   b_group0=b_group0, nbg=nbg, kr=kr, stages_per_item=spi, n_slabs=slabs, n_groups=groups, n_mtiles=1, taps_per_group=tpg,
                   b_rows_stage=brs, b_groups_stage=nbg, b_group_step=0, b_row_lo=(C.c_int32 * 4)(*lo),
                   tap_off=(C.c_int32 * 16)(*off), out=ws.data_ptr(), mt_stride=0, slab_stride=ntaps * 128 * N,
                   g_stride=tpg * 128 * N, tap_stride=128 * N, ldo=N, m_limit=128, n_limit=1 << 30, g_cols=0, t_cols=0)
    eng._check(eng.L.xq_twgrad(eng.h, C.byref(d), eng._stream()))
    return slabs


def dense_wgrad(eng, dl_g, f_g, rows, n_boards, m_total, n_total, out, ldo):
    """out[m][n] = sum_board dL[board][m] * F[board][n] written straight into a row-major [m_total][ldo] matrix (no slabs);
    dl_g: G layout with >= ceil(m_total/128)*4 groups, f_g: G layout with >= ceil(n_total/384)*12 groups."""
    kr = 32
    spi = (n_boards + kr - 1) // kr
    n_groups = (n_total + 383) // 384
    off = [0] * 16
    for g in range(4):
        for t in range(3):
            off[g * 4 + t] = t * 4 * kr * 128
    d = TWgradDesc(a=dl_g.data_ptr(), b=f_g.data_ptr(), a_rows=rows, a_row0=ROW0, b_rows=rows, b_row0=ROW0, a_group0=0, b_group0=0,
                   nbg=4, kr=kr, stages_per_item=spi, n_slabs=1, n_groups=n_groups, n_mtiles=(m_total + 127) // 128, taps_per_group=3,
                   b_rows_stage=kr, b_groups_stage=12, b_group_step=12, b_row_lo=(C.c_int32 * 4)(0, 0, 0, 0),
                   tap_off=(C.c_int32 * 16)(*off), out=out.data_ptr(), mt_stride=128 * ldo, slab_stride=0, g_stride=384,
                   tap_stride=128, ldo=ldo, m_limit=m_total, n_limit=n_total, g_cols=384, t_cols=128)
    eng._check(eng.L.xq_twgrad(eng.h, C.byref(d), eng._stream()))


# =================================================================================================
# The whole training step of XiangqiNet on the hand-written kernels
# =================================================================================================
from xq_native import TnBnDesc, TnBnBwdDesc, TnWimageItem   # noqa: E402

ACTIONS = 8100
FC_IN = 2880            # 32 policy channels x 90 cells (model.py:64-70)
FC_NT, FC_KB = 64, 90   # policy FC image: 64 tiles of 128 logits, 90 blocks of 32 features
FCT_NT, FCT_KB = 23, 254  # transposed: 23 tiles of 128 features, 254 blocks of 32 logits
FC_SPLITS, FCT_SPLITS = 2, 6   # contraction splits of the policy FC: 64 x 2 and 23 x 6 work items for the 148 SMs


def _ptr(t):
    return None if t is None else t.data_ptr()


class _StepBuffers:
    """Activations, gradients and scratch of one minibatch size (allocated once, zero pad cells / pad rows never written)."""

    def __init__(self, dev, B, C, layers):
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        self.B = B
        R, RD = plane_rows(B), dense_rows(B)
        self.R, self.RD = R, RD
        self.n_rows = B * BOARD_ROWS
        self.pairs = (self.n_rows + PAIR - 1) // PAIR
        self.dpairs = (B + PAIR - 1) // PAIR
        cc, cg = C // 4, (C // 32 + 3) // 4 * 4      # G tensors in whole M tiles of 128 channels (the wgrad kernel reads 4 groups; extra ones stay zero)
        self.x0p, self.x0g = z(8, R, 4), z(1, R, 32)
        self.Y = [z(cc, R, 4) for _ in range(layers)]
        self.Ap = [z(cc, R, 4) for _ in range(layers)]
        self.Ag = [z(cg, R, 32) for _ in range(layers)]
        self.save = [z(2, C) for _ in range(layers)]
        self.Yh, self.Ah = z(16, R, 4), z(16, R, 4)
        self.save_p, self.save_v = z(2, 32), z(2, 4)
        self.Fp, self.Fg = z(FC_KB * 8, RD, 4), z(96, RD, 32)
        self.logits = z(B, ACTIONS)
        self.h, self.v = z(B, 128), z(B)
        self.g_logits, self.g_value, self.prow, self.vrow = z(B, ACTIONS), z(B), z(B), z(B)
        self.dlp, self.dlg = z(2048, RD, 4), z(256, RD, 32)
        self.dFp = z(FCT_SPLITS, FCT_NT * 32, RD, 4)       # partial sums of the FC's input gradient, one per contraction split
        self.dAh, self.dYh, self.dYhg = z(16, R, 4), z(16, R, 4), z(2, R, 32)
        self.dh, self.dpre = z(B, 128), z(B)
        self.dA, self.dB, self.dskip = z(cc, R, 4), z(cc, R, 4), z(cc, R, 4)
        self.dYp, self.dYg = z(cc, R, 4), z(cg, R, 32)
        slabs, _ = conv_wgrad_geometry(self.n_rows)
        self.slabs = slabs
        self.ws = z(slabs, 9, 128, 128)
        self.partial = torch.zeros((max(cc, 16) + 2, 16, 8), dtype=torch.float64, device=dev)
        self.states = z(B, 15, 10, 9)          # static inputs of a captured step
        self.act = torch.zeros((B, xq_native.MAX_MOVES), dtype=torch.int16, device=dev)
        self.prob = z(B, xq_native.MAX_MOVES)
        self.n = torch.zeros((B,), dtype=torch.int32, device=dev)
        self.z = z(B)
        self.losses = z(2)
        self.graph, self.graph_inv = None, None
        self.launches = 0


class HandStep:
    """Forward + loss + backward of XiangqiNet (model.py:39-107) for train.py:397-423, every layer on the kernels of
    csrc/xq_tnet.cu: tf32 tcgen05 contractions (xq_tgemm / xq_twgrad) and the plane-layout layers in between.  Parameters,
    BatchNorm buffers and gradients are the torch module's own tensors (read and written through raw pointers; every
    parameter's .grad must exist and is ASSIGNED by a step), so the optimiser, checkpoints and predict() are unchanged.
    No torch op, cuDNN or cuBLAS call runs inside a step."""

    def __init__(self, eng, model):
        self.eng, self.model = eng, model
        self.C, self.nb = model.num_channels, model.num_res_blocks
        if self.C % 32:
            raise ValueError("HandStep needs a channel count that is a multiple of 32")
        dev = next(model.parameters()).device
        self.dev = dev
        C = self.C
        self.convs = [(model.input_conv[0], model.input_conv[1])]
        for blk in model.res_blocks:
            self.convs += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2)]
        self.L = len(self.convs)
        z = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)
        nt, kb = (C + 127) // 128, C // 32
        self.img_f = [z(nt, 9, 1, 8, 128, 4)] + [z(nt, 9, kb, 8, 128, 4) for _ in range(self.L - 1)]
        self.img_t = [None] + [z(nt, 9, kb, 8, 128, 4) for _ in range(self.L - 1)]
        self.img_hf = z(1, 1, kb, 8, 128, 4)            # heads 1x1: n = 32 policy + 4 value channels
        self.img_ht = z(nt, 1, 2, 8, 128, 4)            # transposed: n = tower channel, k = head channel (64 = 2 blocks)
        self.img_fc = z(FC_NT, 1, FC_KB, 8, 128, 4)
        self.img_fct = z(FCT_NT, 1, FCT_KB, 8, 128, 4)
        self.pconv, self.pbn, self.fc = model.policy_head[0], model.policy_head[1], model.policy_head[4]
        self.vconv, self.vbn, self.v1, self.v2 = model.value_head[0], model.value_head[1], model.value_head[4], model.value_head[6]
        self._bufs = {}
        self._items = None
        self.use_graph = os.environ.get("XQ_TRAIN_GRAPH", "1") != "0"      # 0: issue every launch from the host (A/B, profiling)
        self.bn_steps = 0        # num_batches_tracked is brought up to date by sync_counters()

    # ---- pieces ----------------------------------------------------------------------------------------------
    MAX_BUFFER_SETS = 3      # the full minibatch, the shorter last minibatch of an epoch, one spare (a set is ~0.9 GB at 256 x 128 channels)

    def buffers(self, B):
        """The buffer set (and captured graph) of minibatch size B, most recently used last.  The last minibatch of an epoch
        changes size as the replay buffer grows: only MAX_BUFFER_SETS sets are kept, the least recently used one is dropped."""
        b = self._bufs.pop(B, None)
        if b is None:
            while len(self._bufs) >= self.MAX_BUFFER_SETS:
                old = self._bufs.pop(next(iter(self._bufs)))
                if old.graph is not None:
                    torch.cuda.synchronize()         # no replay of it may be in flight (happens once per epoch at most)
                old.graph = None                     # the graph references the set's tensors: drop it first
                del old
            b = _StepBuffers(self.dev, B, self.C, self.L)
        self._bufs[B] = b
        return b

    def _image_items(self):
        C_, kb = self.C, self.C // 32
        items = []
        for i, (conv, _) in enumerate(self.convs):
            w = conv.weight
            ci = w.shape[1]
            items.append(TnWimageItem(w=w.data_ptr(), img=self.img_f[i].data_ptr(), co=C_, ci=ci, taps=9, img_kb=1 if i == 0 else kb,
                                      n0=0, k0=0, transposed=0, pad_=0))
            if i:
                items.append(TnWimageItem(w=w.data_ptr(), img=self.img_t[i].data_ptr(), co=C_, ci=ci, taps=9, img_kb=kb, n0=0, k0=0,
                                          transposed=1, pad_=0))
        for conv, off in ((self.pconv, 0), (self.vconv, 32)):
            co = conv.weight.shape[0]
            items.append(TnWimageItem(w=conv.weight.data_ptr(), img=self.img_hf.data_ptr(), co=co, ci=C_, taps=1, img_kb=kb, n0=off, k0=0,
                                      transposed=0, pad_=0))
            items.append(TnWimageItem(w=conv.weight.data_ptr(), img=self.img_ht.data_ptr(), co=co, ci=C_, taps=1, img_kb=2, n0=0, k0=off,
                                      transposed=1, pad_=0))
        return (TnWimageItem * len(items))(*items), len(items)

    def build_images(self):
        """Weight images of every contraction from the current parameters (after each optimiser step): the convolutions in
        one launch per 32 tensors, the two images of the 93 MB policy FC weight in one pass over it."""
        e, L = self.eng, self.eng.L
        s = e._stream()
        if self._items is None:
            self._items = self._image_items()
        e._check(L.xq_tn_wimage_batch(e.h, self._items[0], self._items[1], s))
        e._check(L.xq_tn_wimage_dense2(e.h, self.fc.weight.data_ptr(), ACTIONS, FC_IN, self.img_fc.data_ptr(), FC_KB,
                                       self.img_fct.data_ptr(), FCT_KB, s))

    def _bn_fwd(self, b, bn, y, out, out_g, chunk0, nch, res=None, save=None):
        e = self.eng
        d = TnBnDesc(y=_ptr(y), res=_ptr(res), out=_ptr(out), out_g=_ptr(out_g), rows=b.R, n_boards=b.B, chunk0=chunk0, n_channels=nch,
                     relu=1, partial=b.partial.data_ptr(), gamma=bn.weight.data_ptr(), beta=bn.bias.data_ptr(),
                     running_mean=bn.running_mean.data_ptr(), running_var=bn.running_var.data_ptr(), save=save.data_ptr(),
                     eps=float(bn.eps), momentum=float(bn.momentum))
        e._check(e.L.xq_tn_bn_forward(e.h, C.byref(d), e._stream()))

    def _bn_bwd(self, b, bn, dout, act, y, save, dy, dy_g, chunk0, nch, dskip=None):
        e = self.eng
        d = TnBnBwdDesc(dout=_ptr(dout), act=_ptr(act), y=_ptr(y), rows=b.R, n_boards=b.B, chunk0=chunk0, n_channels=nch, relu=1,
                        save=save.data_ptr(), partial=b.partial.data_ptr(), gamma=bn.weight.data_ptr(), dgamma=bn.weight.grad.data_ptr(),
                        dbeta=bn.bias.grad.data_ptr(), dy=_ptr(dy), dy_g=_ptr(dy_g), dskip=_ptr(dskip))
        e._check(e.L.xq_tn_bn_backward(e.h, C.byref(d), e._stream()))

    def _conv_wgrad(self, b, dy_g, x_g, conv, ci):
        """conv.weight.grad[co][ci][3][3] from the G-layout gradient and input."""
        e, Cn = self.eng, self.C
        nbg = min(4, (ci + 31) // 32)
        for mt in range((Cn + 127) // 128):
            for nt in range((ci + 127) // 128):
                conv_wgrad(e, dy_g, x_g, b.R, b.n_rows, nbg, 9, b.ws, a_group0=mt * 4, b_group0=nt * 4)
                m_cnt, n_cnt = min(128, Cn - mt * 128), min(128, ci - nt * 128)
                e._check(e.L.xq_tn_wgrad_reduce(e.h, b.ws.data_ptr(), b.slabs, 9 * 128 * 32 * nbg, 9, 32 * nbg, m_cnt, n_cnt, 0, 0,
                                                conv.weight.grad.data_ptr(), ci, mt * 128, nt * 128, e._stream()))

    # ---- the step --------------------------------------------------------------------------------------------
    def forward(self, b, states):
        e, L, Cn = self.eng, self.eng.L, self.C
        s = e._stream
        nt, kb, cc = (Cn + 127) // 128, Cn // 32, Cn // 4
        e._check(L.xq_tn_input(e.h, states.data_ptr(), b.B, 15, 4, b.x0p.data_ptr(), b.x0g.data_ptr(), b.R, s()))
        tgemm(e, b.x0p, b.R, 1, self.img_f[0], 9, 1, False, b.pairs, nt, b.n_rows, out=b.Y[0], out_rows=b.R, out_chunks=cc)
        self._bn_fwd(b, self.convs[0][1], b.Y[0], b.Ap[0], b.Ag[0], 0, Cn, save=b.save[0])
        for i in range(1, self.L):
            tgemm(e, b.Ap[i - 1], b.R, kb, self.img_f[i], 9, kb, False, b.pairs, nt, b.n_rows, out=b.Y[i], out_rows=b.R, out_chunks=cc)
            res = b.Ap[i - 2] if i % 2 == 0 else None                      # second conv of a block: + the block's input
            self._bn_fwd(b, self.convs[i][1], b.Y[i], b.Ap[i], b.Ag[i], 0, Cn, res=res, save=b.save[i])
        top = b.Ap[self.L - 1]
        tgemm(e, top, b.R, kb, self.img_hf, 1, kb, False, b.pairs, 1, b.n_rows, out=b.Yh, out_rows=b.R, out_chunks=9)
        self._bn_fwd(b, self.pbn, b.Yh, b.Ah, None, 0, 32, save=b.save_p)
        self._bn_fwd(b, self.vbn, b.Yh, b.Ah, None, 8, 4, save=b.save_v)
        e._check(L.xq_tn_flatten(e.h, b.Ah.data_ptr(), b.R, b.B, 32, b.Fp.data_ptr(), b.Fg.data_ptr(), b.RD, s()))
        tgemm(e, b.Fp, b.RD, FC_KB, self.img_fc, 1, FC_KB, False, b.dpairs, FC_NT, b.B, out_rm=b.logits, out_stride=ACTIONS,
              bias=self.fc.bias, n_cols=ACTIONS, k_splits=FC_SPLITS)
        e._check(L.xq_tn_value_forward(e.h, b.Ah.data_ptr(), b.R, 8, b.B, self.v1.weight.data_ptr(), self.v1.bias.data_ptr(),
                                       self.v2.weight.data_ptr(), self.v2.bias.data_ptr(), b.h.data_ptr(), b.v.data_ptr(), s()))
        return b.logits, b.v

    def loss(self, b, act, prob, n, z, inv_batch):
        e = self.eng
        e._check(e.L.xq_policy_value_loss(e.h, b.logits.data_ptr(), ACTIONS, b.v.data_ptr(), act.data_ptr(), prob.data_ptr(), n.data_ptr(),
                                          z.data_ptr(), b.B, C.c_float(inv_batch), b.g_logits.data_ptr(), ACTIONS, b.g_value.data_ptr(),
                                          b.prow.data_ptr(), b.vrow.data_ptr(), e._stream()))

    def backward(self, b):
        e, L, Cn = self.eng, self.eng.L, self.C
        s = e._stream
        nt, kb, cc = (Cn + 127) // 128, Cn // 32, Cn // 4
        fc = self.fc
        # policy FC
        e._check(L.xq_tn_colsum(e.h, b.g_logits.data_ptr(), ACTIONS, b.B, ACTIONS, fc.bias.grad.data_ptr(), s()))
        e._check(L.xq_tn_rows_layouts(e.h, b.g_logits.data_ptr(), ACTIONS, b.B, ACTIONS, b.dlp.data_ptr(), b.dlg.data_ptr(), b.RD, s()))
        dense_wgrad(e, b.dlg, b.Fg, b.RD, b.B, ACTIONS, FC_IN, fc.weight.grad, FC_IN)
        tgemm(e, b.dlp, b.RD, FCT_KB, self.img_fct, 1, FCT_KB, True, b.dpairs, FCT_NT, b.B, out=b.dFp, out_rows=b.RD, out_chunks=FC_IN // 4,
              k_splits=FCT_SPLITS, out_split_stride=FCT_NT * 32 * b.RD * 4)
        e._check(L.xq_tn_unflatten(e.h, b.dFp.data_ptr(), b.RD, b.B, 32, b.dAh.data_ptr(), b.R, FCT_SPLITS, FCT_NT * 32 * b.RD * 4, s()))
        # value head
        v1, v2 = self.v1, self.v2
        e._check(L.xq_tn_value_backward(e.h, b.Ah.data_ptr(), b.R, 8, b.B, v1.weight.data_ptr(), v2.weight.data_ptr(), b.h.data_ptr(),
                                        b.v.data_ptr(), b.g_value.data_ptr(), b.dh.data_ptr(), b.dpre.data_ptr(), b.dAh.data_ptr(),
                                        v1.weight.grad.data_ptr(), v1.bias.grad.data_ptr(), v2.weight.grad.data_ptr(),
                                        v2.bias.grad.data_ptr(), s()))
        # heads: BatchNorm + 1x1 convs (one 36-channel contraction each way)
        self._bn_bwd(b, self.pbn, b.dAh, b.Ah, b.Yh, b.save_p, b.dYh, b.dYhg, 0, 32)
        self._bn_bwd(b, self.vbn, b.dAh, b.Ah, b.Yh, b.save_v, b.dYh, b.dYhg, 8, 4)
        top_g = b.Ag[self.L - 1]
        for mt in range(nt):
            conv_wgrad(e, top_g, b.dYhg, b.R, b.n_rows, 2, 1, b.ws, a_group0=mt * 4)
            for conv, n0, cnt in ((self.pconv, 0, 32), (self.vconv, 32, 4)):
                e._check(L.xq_tn_wgrad_reduce(e.h, b.ws.data_ptr(), b.slabs, 128 * 64, 1, 64, min(128, Cn - mt * 128), cnt, n0, 1,
                                              conv.weight.grad.data_ptr(), Cn, 0, mt * 128, s()))
        tgemm(e, b.dYh, b.R, 2, self.img_ht, 1, 2, True, b.pairs, nt, b.n_rows, out=b.dA, out_rows=b.R, out_chunks=cc)
        # tower, last layer first
        for i in range(self.L - 1, 0, -1):
            conv, bn = self.convs[i]
            second = i % 2 == 0
            dout = b.dA if second else b.dB
            self._bn_bwd(b, bn, dout, b.Ap[i], b.Y[i], b.save[i], b.dYp, b.dYg, 0, Cn, dskip=b.dskip if second else None)
            self._conv_wgrad(b, b.dYg, b.Ag[i - 1], conv, Cn)
            if second:
                tgemm(e, b.dYp, b.R, kb, self.img_t[i], 9, kb, True, b.pairs, nt, b.n_rows, out=b.dB, out_rows=b.R, out_chunks=cc)
            else:       # first conv of the block: + the gradient that bypassed the block
                tgemm(e, b.dYp, b.R, kb, self.img_t[i], 9, kb, True, b.pairs, nt, b.n_rows, out=b.dA, out_rows=b.R, out_chunks=cc,
                      residual=b.dskip)
        conv, bn = self.convs[0]
        self._bn_bwd(b, bn, b.dA, b.Ap[0], b.Y[0], b.save[0], b.dYp, b.dYg, 0, Cn)
        self._conv_wgrad(b, b.dYg, b.x0g, conv, 15)

    def _run(self, b, inv_batch):
        self.build_images()
        self.forward(b, b.states)
        self.loss(b, b.act, b.prob, b.n, b.z, inv_batch)
        self.backward(b)
        b.losses.copy_(torch.stack([b.prow.sum(), b.vrow.sum()]) * inv_batch)

    def step(self, states, act, prob, n, z, inv_batch):
        """One forward + loss + backward; returns (policy_loss, value_loss) as 0-d device tensors (views of a static buffer:
        read them before the next step).  Gradients are in .grad.  The ~165 launches of a step are replayed from a CUDA
        graph (captured after the first, eagerly run, step of a minibatch size): the step is shorter than the host time
        to issue it."""
        B = int(states.shape[0])
        b = self.buffers(B)
        for dst, src in ((b.states, states), (b.act, act), (b.prob, prob), (b.n, n), (b.z, z)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src)
        if not self.use_graph:
            self._run(b, inv_batch)
        elif b.graph is not None and b.graph_inv == inv_batch:
            b.graph.replay()
        else:
            n0 = self.eng.launch_count()
            self._run(b, inv_batch)                  # eager: also sets the kernels' attributes before any capture
            b.launches = self.eng.launch_count() - n0        # kernels of csrc/xq_tnet.cu + the loss kernel in one step (a replay launches the same)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # capture only (nothing executes): replayed from the next step on.  thread_local: CUDA calls of other threads
            # (the NCCL watchdog of a multi-rank run) must not invalidate the capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._run(b, inv_batch)
            b.graph, b.graph_inv = g, inv_batch
        self.bn_steps += 1
        return b.losses[0], b.losses[1]

    def sync_counters(self):
        """nn.BatchNorm2d.num_batches_tracked of the 15 layers (one add per layer per call instead of one per step)."""
        if self.bn_steps:
            for m in self.model.modules():
                if isinstance(m, torch.nn.BatchNorm2d) and m.num_batches_tracked is not None:
                    m.num_batches_tracked.add_(self.bn_steps)
            self.bn_steps = 0
