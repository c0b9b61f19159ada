"""Host side of the hand-written training step (csrc/xq_tnet.cu): the training plane layout, weight images and the
descriptors of the two tcgen05 tf32 kernels (xq_tgemm: fprop / dgrad, xq_twgrad: weight gradients).

Training plane layout (include/xq_b200.h): float32 X[C/4][rows][4]; board b, cell (r, c) at row ROW0 + b*110 + r*10 + c.
Every board row carries a zero pad column (c = 9) and every board a zero pad row (r = 10): a 3x3 tap (dy, dx) is a row shift
of dy*10+dx and the pad cells are conv2d's zero padding (train.py:397-423 over model.py:39-107)."""
import ctypes as C

import torch

import xq_native
from xq_native import TGemmDesc, TWgradDesc

ROW0 = 16
BOARD_ROWS = 110
PAIR = 256


def plane_rows(n_boards: int) -> int:
    """Rows of a plane tensor for n_boards boards: front padding, whole 256-row work items, tail padding (the wgrad slabs and
    the 11-row halo of the last tile read past the last board; those rows stay zero)."""
    return ROW0 + (n_boards * BOARD_ROWS + PAIR - 1) // PAIR * PAIR + 192


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
          residual=None, out_rm=None, out_stride=0, bias=None, n_cols=0):
    """fprop (w_img = image of the weights) or dgrad (w_img = image of the TRANSPOSED weights, taps mirrored)."""
    d = TGemmDesc(a=a.data_ptr(), a_rows=a_rows, a_row0=ROW0, w=w_img.data_ptr(), kblocks=kblocks, ntaps=ntaps, img_kb=img_kb,
                  b_mn=0, shift_sign=-1 if dgrad else 1, m_pairs=m_pairs, n_tiles=n_tiles, out_chunks=out_chunks,
                  m_rows=m_rows, out=None if out is None else out.data_ptr(), out_rows=out_rows, out_row0=ROW0,
                  residual=None if residual is None else residual.data_ptr(), out_rm=None if out_rm is None else out_rm.data_ptr(),
                  out_stride=out_stride, bias=None if bias is None else bias.data_ptr(), n_cols=n_cols, pad_=0)
    eng._check(eng.L.xq_tgemm(eng.h, C.byref(d), eng._stream()))


def conv_wgrad_geometry(n_rows: int, sm_count: int = 148):
    """(slabs, stages per item) of the conv weight-gradient launch: 64-row stages, 3 tap groups per slab, about one item per SM."""
    kr = 64
    slabs = max(1, min(sm_count // 3, (n_rows + kr - 1) // kr))
    spi = (n_rows + slabs * kr - 1) // (slabs * kr)
    slabs = (n_rows + spi * kr - 1) // (spi * kr)
    return slabs, spi


def conv_wgrad(eng, dy_g, x_g, rows, n_rows, nbg, ntaps, ws):
    """ws[slab][tap][128][32*nbg] = partial sums of dY[row][m] * X[row + shift(tap)][n] over the slab's rows (G-layout operands)."""
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
    d = TWgradDesc(a=dy_g.data_ptr(), b=x_g.data_ptr(), a_rows=rows, a_row0=ROW0, b_rows=rows, b_row0=ROW0, a_group0=0, b_group0=0,
                   nbg=nbg, kr=kr, stages_per_item=spi, n_slabs=slabs, n_groups=groups, n_mtiles=1, taps_per_group=tpg,
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
