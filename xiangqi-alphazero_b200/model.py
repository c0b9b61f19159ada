"""Drop-in replacement for the reference's training/model.py.

`XiangqiNet` keeps the reference architecture and `state_dict` keys (model.py:39-107:
input_conv.{0,1}, res_blocks.{i}.{conv1,bn1,conv2,bn2}, policy_head.{0,1,4}, value_head.{0,1,4,6})
so checkpoints, train.py's optimiser and export_model.py keep working -- training still runs
the torch module.  What changes is inference: `predict()` and the batched `B200Net` run the
forward as hand-written bf16 tcgen05/TMA implicit-GEMM kernels (libxq_b200.so, csrc/xq_net.cu)
with BatchNorm folded into the weights.  There is no torch/cuDNN fallback on that path.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

import xq_native
from game import ACTION_SPACE, COLS, ROWS


class ResBlock(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(channels)
        self.conv2 = nn.Conv2d(channels, channels, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(channels)

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + x)


class XiangqiNet(nn.Module):
    """Policy-value ResNet, same modules and parameter names as the reference."""

    def __init__(self, num_channels: int = 128, num_res_blocks: int = 6):
        super().__init__()
        self.num_channels = num_channels
        self.num_res_blocks = num_res_blocks
        self.input_conv = nn.Sequential(nn.Conv2d(15, num_channels, 3, padding=1, bias=False),
                                        nn.BatchNorm2d(num_channels), nn.ReLU())
        self.res_blocks = nn.ModuleList([ResBlock(num_channels) for _ in range(num_res_blocks)])
        self.policy_head = nn.Sequential(nn.Conv2d(num_channels, 32, 1, bias=False), nn.BatchNorm2d(32), nn.ReLU(),
                                         nn.Flatten(), nn.Linear(32 * ROWS * COLS, ACTION_SPACE))
        self.value_head = nn.Sequential(nn.Conv2d(num_channels, 4, 1, bias=False), nn.BatchNorm2d(4), nn.ReLU(),
                                        nn.Flatten(), nn.Linear(4 * ROWS * COLS, 128), nn.ReLU(), nn.Linear(128, 1),
                                        nn.Tanh())
        self._b200 = None
        self._b200_version = None
        self._weights_generation = 0      # bumped by anything that rewrites the parameters through raw pointers (FlatAdam.step)

    def forward(self, x):
        out = self.input_conv(x)
        for block in self.res_blocks:
            out = block(out)
        return self.policy_head(out), self.value_head(out)

    # ---- inference on the B200 kernels --------------------------------------------------------
    def _state_version(self):
        """Changes whenever the weights may have changed: torch's in-place version counters (optimizer.step of a torch
        optimiser, load_state_dict, BatchNorm statistics updates) plus the explicit generation counter for writers
        torch cannot see (FlatAdam steps the flat parameter buffer through xq_adam_step's raw pointers)."""
        return (self._weights_generation,) + tuple(int(p._version) for p in self.state_dict().values())

    def invalidate_b200(self):
        """Drop the kernel-side weight copy: the next predict()/b200() folds the current parameters again."""
        self._weights_generation += 1
        self._b200 = None

    def b200(self, max_batch: int = 1, device: int = 0) -> "B200Net":
        """Kernel-side copy of the current weights (rebuilt when the parameters change)."""
        ver = self._state_version()
        if self._b200 is None or self._b200_version != ver or self._b200.max_batch < max_batch:
            from game import engine
            self._b200 = B200Net(engine(device), self, max_batch=max_batch)
            self._b200_version = ver
        return self._b200

    def predict(self, state: np.ndarray, device: str = 'cpu') -> tuple:
        """model.py:109-124 contract: (15,10,9) planes -> (float32[8100] softmax probabilities, float).
        `device` is accepted for signature compatibility; inference always runs on the engine's GPU."""
        net = self.b200(1)
        x = torch.from_numpy(np.ascontiguousarray(state, np.float32)).reshape(1, 15, 10, 9)
        logits, value = net.forward_planes(x)
        probs = torch.softmax(logits[:, :ACTION_SPACE].float(), dim=1)[0].cpu().numpy()
        return probs, float(value[0].item())


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)


# =================================================================================================
# Kernel-side network: folded weights as shared-memory images + layer descriptors
# =================================================================================================
class _GemmDesc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("m_tiles", C.c_int32), ("n_tiles", C.c_int32), ("nt", C.c_int32),
                ("kchunks", C.c_int32), ("kch_iter", C.c_int32), ("relu", C.c_int32), ("n_boards", C.c_int32),
                ("a_rows", C.c_int64), ("a_row0", C.c_int64), ("out_rows", C.c_int64), ("out_row0", C.c_int64),
                ("out_stride", C.c_int64), ("a", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p),
                ("residual", C.c_void_p), ("out", C.c_void_p), ("out2", C.c_void_p), ("w_half", C.c_void_p)]


ROW0 = 16            # plane row of logical row 0 (front padding, >= 10 rows: the A block of the first tile starts 10 rows early)
BOARD_ROWS = 90      # plane rows per board: cell (r, c) of board b at row ROW0 + b*90 + r*9 + c, no halo (csrc/xq_net.cu)
TAP_ORDER = [4, 0, 1, 2, 3, 5, 6, 7, 8]   # weight-image tap k -> 3x3 cell kh*3+kw: the (unmasked) centre tap is issued first
FC_NT = 224          # output columns per FC work item (one N = 224 MMA per tile and K step, csrc/xq_net.cu fc_kernel)
FC_TILES = 37        # 37 x 224 = 8288 >= 8100
FC_NT_SMALL = 64     # plans of at most FC_SMALL_BOARDS boards: 127 column tiles of 64, one CTA each (the layer is weight-bound there)
FC_SMALL_BOARDS = 256
LOGIT_STRIDE = 8320  # row stride of the logits (>= FC_TILES * FC_NT, multiple of 64 elements)


def fold_bn(conv_w, bn):
    """Conv(no bias) + BatchNorm(eval) -> weight, bias (fp32)."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    w = conv_w.detach().float() * scale.view(-1, 1, 1, 1)
    b = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return w, b


def conv_image(w, nt, kch_iter):
    """[Co][Ci][kh][kw] fp32 -> bf16 image [n_tile][tap][k_block][chunk][nt][8] (csrc/xq_net.cu)."""
    co, ci, kh, kw = w.shape
    ci_pad = (ci + 7) // 8 * 8
    co_pad = (co + nt - 1) // nt * nt
    wp = torch.zeros((co_pad, ci_pad, kh, kw), dtype=torch.float32, device=w.device)
    wp[:co, :ci] = w
    taps = kh * kw
    x = wp.permute(2, 3, 0, 1).reshape(taps, co_pad, ci_pad)                 # [tap][co][ci]
    if taps == 9:
        x = x[TAP_ORDER]
    x = x.reshape(taps, co_pad // nt, nt, ci_pad // (8 * kch_iter), kch_iter, 8)  # [tap][ntile][n][kb][chunk][8]
    x = x.permute(1, 0, 3, 4, 2, 5).contiguous()                             # [ntile][tap][kb][chunk][n][8]
    return x.to(torch.bfloat16)


class B200Net:
    """Batched forward of XiangqiNet on the tcgen05 kernels.  Buffers are torch tensors (device
    memory only); every layer is one xq_net_gemm launch, replayed from C by xq_net_run."""

    def __init__(self, eng: "xq_native.Engine", model: XiangqiNet, max_batch: int):
        self.e = eng
        self.max_batch = int(max_batch)
        dev = eng.dev
        Cm, R = model.num_channels, model.num_res_blocks
        # The conv kernels tile output channels by 128, or by 64 for towers of at most 64 channels (the reference's quick
        # preset, train.py:661).  Other widths run zero-padded to the next multiple: padded channels have zero weights and
        # zero bias, stay exactly 0 through ReLU / residual adds and contribute nothing downstream.
        NTc = 64 if Cm <= 64 else 128
        Cc = (Cm + NTc - 1) // NTc * NTc
        self.C, self.R, self.C_model = Cc, R, Cm
        B = self.max_batch
        self.m_tiles = (B * BOARD_ROWS + 127) // 128
        pairs = (self.m_tiles + 1) // 2                            # conv kernels work on pairs of 128-row tiles,
        self.rows = ROW0 + (pairs + 1) // 2 * 2 * 256 + 16         # whole CTA-pair items (2 tile pairs) + the 10 rows read past the end
        self.b_tiles = (B + 127) // 128
        self.fc_rows = ((self.b_tiles + 1) // 2) * 256               # the FC kernel works on pairs of 128-board tiles
        bf = torch.bfloat16
        z = lambda *shape, dt=bf: torch.zeros(shape, dtype=dt, device=dev)
        self.x0 = z(2, self.rows, 8)                     # input planes (15 + 1 pad channels)
        self.act = [z(Cc // 8, self.rows, 8) for _ in range(3)]
        self.fc_in = z(360, self.fc_rows, 8)             # policy features, k = pos*32 + ch
        self.vfeat = z(B, 90, 4, dt=torch.float32)
        self.logits = z(self.fc_rows, LOGIT_STRIDE)
        self.value = z(B, dt=torch.float32)
        self.keep = []                                   # weight images / biases (owned here)
        self.layers = []
        # Fold and re-tile ON THE DEVICE from a float32 copy of the state dict (one 100 MB upload, then device ops): the caller's
        # module (device, mode) is left untouched.  _SD gives attribute / index access by the state_dict key names.
        sd = {k: v.detach().to(device=dev, dtype=torch.float32 if v.is_floating_point() else v.dtype) for k, v in model.state_dict().items()}
        eps = {name: mod.eps for name, mod in model.named_modules() if isinstance(mod, nn.modules.batchnorm._BatchNorm)}

        class _SD:
            def __init__(self, prefix):
                self._p = prefix

            def _sub(self, name):
                full = f"{self._p}.{name}" if self._p else str(name)
                return sd[full] if full in sd else _SD(full)

            __getattr__ = lambda self, name: self._sub(name) if not name.startswith("_") else object.__getattribute__(self, name)
            __getitem__ = lambda self, i: self._sub(i)

            @property
            def eps(self):
                return eps[self._p]
        m = _SD("")
        m_res_blocks = [_SD(f"res_blocks.{i}") for i in range(R)]

        def dev_t(t, dt=None):
            t = t.to(dev) if dt is None else t.to(device=dev, dtype=dt)
            t = t.contiguous()
            self.keep.append(t)
            return t

        def pad_channels(w, b, co, ci):
            wp = torch.zeros((co, ci) + tuple(w.shape[2:]), dtype=torch.float32, device=w.device)
            wp[:w.shape[0], :w.shape[1]] = w
            bp = torch.zeros(co, dtype=torch.float32, device=w.device)
            bp[:b.shape[0]] = b
            return wp, bp

        def conv_layer(w, b, a_buf, out_buf, residual, relu, kch_iter):
            w, b = pad_channels(w, b, Cc, w.shape[1] if w.shape[1] == 15 else Cc)
            img = dev_t(conv_image(w, NTc, kch_iter))
            img64 = dev_t(conv_image(w, 64, kch_iter)) if (kch_iter == 8 and NTc == 128) else None   # halves of N for the CTA-pair kernel
            bias = dev_t(b, torch.float32)
            d = _GemmDesc(mode=0, m_tiles=self.m_tiles, n_tiles=Cc // NTc, nt=NTc, kchunks=a_buf.shape[0],
                          kch_iter=kch_iter, relu=int(relu), n_boards=B, a_rows=self.rows, a_row0=ROW0,
                          out_rows=self.rows, out_row0=ROW0, out_stride=0, a=a_buf.data_ptr(), w=img.data_ptr(),
                          bias=bias.data_ptr(), residual=None if residual is None else residual.data_ptr(),
                          out=out_buf.data_ptr(), out2=None, w_half=None if img64 is None else img64.data_ptr())
            self.layers.append(d)

        with torch.no_grad():
            w, b = fold_bn(m.input_conv[0].weight, m.input_conv[1])
            conv_layer(w, b, self.x0, self.act[0], None, True, 2)
            cur = 0
            for blk in m_res_blocks:
                t, o = (cur + 1) % 3, (cur + 2) % 3
                w, b = fold_bn(blk.conv1.weight, blk.bn1)
                conv_layer(w, b, self.act[cur], self.act[t], None, True, 8)
                w, b = fold_bn(blk.conv2.weight, blk.bn2)
                conv_layer(w, b, self.act[t], self.act[o], self.act[cur], True, 8)
                cur = o
            self.trunk_out = cur
            # heads: policy conv1x1 (32) and value conv1x1 (4) share one GEMM with N = 48
            wp, bp = fold_bn(m.policy_head[0].weight, m.policy_head[1])
            wv, bv = fold_bn(m.value_head[0].weight, m.value_head[1])
            wh = torch.zeros((48, Cc, 1, 1), device=dev)
            bh = torch.zeros(48, device=dev)
            wh[:32, :Cm], wh[32:36, :Cm] = wp, wv
            bh[:32], bh[32:36] = bp, bv
            img = dev_t(conv_image(wh, 48, 8))
            bias = dev_t(bh, torch.float32)
            self.layers.append(_GemmDesc(mode=1, m_tiles=self.m_tiles, n_tiles=1, nt=48, kchunks=Cc // 8, kch_iter=8,
                                         relu=1, n_boards=B, a_rows=self.rows, a_row0=ROW0, out_rows=self.fc_rows,
                                         out_row0=0, out_stride=0, a=self.act[cur].data_ptr(), w=img.data_ptr(),
                                         bias=bias.data_ptr(), residual=None, out=self.fc_in.data_ptr(),
                                         out2=self.vfeat.data_ptr()))
            # policy FC 2880 -> 8100: torch flatten index ch*90+pos  ->  kernel index pos*32+ch
            wf = m.policy_head[4].weight.detach().float().reshape(ACTION_SPACE, 32, 90).permute(0, 2, 1)
            fc_nt = FC_NT_SMALL if B <= FC_SMALL_BOARDS else FC_NT
            fc_tiles = (ACTION_SPACE + fc_nt - 1) // fc_nt
            assert fc_tiles * fc_nt <= LOGIT_STRIDE
            wfp = torch.zeros((fc_tiles * fc_nt, 2880), device=dev)
            wfp[:ACTION_SPACE] = wf.reshape(ACTION_SPACE, 2880)
            bfp = torch.zeros(fc_tiles * fc_nt, device=dev)
            bfp[:ACTION_SPACE] = m.policy_head[4].bias.detach().float()
            img = dev_t(conv_image(wfp.reshape(fc_tiles * fc_nt, 2880, 1, 1), fc_nt, 8))
            bias = dev_t(bfp, torch.float32)
            # big plans also carry the image tiled by 64 columns: launches bounded to a few live boards (the tail of an
            # iteration, xq_selfplay_set_live_bound) run the weight-bound layer on 127 CTAs instead of 37
            img_small = None
            if fc_nt != FC_NT_SMALL:
                ts = (ACTION_SPACE + FC_NT_SMALL - 1) // FC_NT_SMALL
                ws = torch.zeros((ts * FC_NT_SMALL, 2880), device=dev)
                ws[:ACTION_SPACE] = wf.reshape(ACTION_SPACE, 2880)
                img_small = dev_t(conv_image(ws.reshape(ts * FC_NT_SMALL, 2880, 1, 1), FC_NT_SMALL, 8))
            self.layers.append(_GemmDesc(mode=2, m_tiles=self.b_tiles, n_tiles=fc_tiles, nt=fc_nt,
                                         kchunks=360, kch_iter=8, relu=0, n_boards=B, a_rows=self.fc_rows, a_row0=0,
                                         out_rows=0, out_row0=0, out_stride=LOGIT_STRIDE, a=self.fc_in.data_ptr(),
                                         w=img.data_ptr(), bias=bias.data_ptr(), residual=None,
                                         out=self.logits.data_ptr(), out2=None,
                                         w_half=None if img_small is None else img_small.data_ptr()))
            # value MLP: k = ch*90+pos -> pos*4+ch, stored transposed [360][128]
            w1 = m.value_head[4].weight.detach().float().reshape(128, 4, 90).permute(2, 1, 0).reshape(360, 128)
            self.w1t = dev_t(w1, torch.float32)
            self.b1 = dev_t(m.value_head[4].bias.detach().float(), torch.float32)
            self.w2 = dev_t(m.value_head[6].weight.detach().float().reshape(128), torch.float32)
            self.b2 = float(m.value_head[6].bias.detach().float().item())
        self.n_layers = len(self.layers)
        self.desc_array = (_GemmDesc * self.n_layers)(*self.layers)

    # -- input helpers ---------------------------------------------------------------------------
    def load_planes(self, planes: torch.Tensor):
        """float planes [n,15,10,9] -> the kernel's input plane tensor (test / predict() path; the
        self-play loop writes x0 directly from the MCTS kernels)."""
        n = planes.shape[0]
        assert n <= self.max_batch
        p = planes.to(self.e.dev, torch.float32)
        x = torch.zeros((n, 10, 9, 16), dtype=torch.bfloat16, device=self.e.dev)
        x[..., :15] = p.permute(0, 2, 3, 1).to(torch.bfloat16)
        self.x0[:, ROW0:ROW0 + n * BOARD_ROWS] = x.reshape(n * BOARD_ROWS, 2, 8).permute(1, 0, 2)

    def run(self, n_boards=None):
        """Enqueue the whole forward of the first n_boards boards (default: max_batch) on the current torch stream
        (xq_net_run sizes every launch to that count)."""
        L = self.e.L
        rc = L.xq_net_run(self.e.h, self.desc_array, self.n_layers, self.vfeat.data_ptr(), self.w1t.data_ptr(),
                          self.b1.data_ptr(), self.w2.data_ptr(), C.c_float(self.b2), self.value.data_ptr(),
                          self.max_batch if n_boards is None else int(n_boards), self.e._stream())
        self.e._check(rc)

    def run_layer(self, i):
        self.e._check(self.e.L.xq_net_gemm(self.e.h, C.byref(self.layers[i]), self.e._stream()))

    def forward_planes(self, planes: torch.Tensor):
        n = planes.shape[0]
        self.load_planes(planes)
        self.run(n)
        return self.logits[:n], self.value[:n]

    # -- debugging / tests: read a plane tensor back as [n,C,10,9] float --------------------------
    def planes_to_nchw(self, buf: torch.Tensor, n: int):
        ch = buf.shape[0] * 8
        x = buf[:, ROW0:ROW0 + n * BOARD_ROWS].permute(1, 0, 2).reshape(n, 10, 9, ch)
        return x.permute(0, 3, 1, 2).float()

    def flops_per_board(self):
        Cc, R = self.C_model, self.R                      # algorithmic FLOPs of the model, not of the padded tiles
        conv = 2 * 90 * 9
        return conv * 15 * Cc + 2 * R * conv * Cc * Cc + 2 * 90 * Cc * 36 + 2 * 2880 * ACTION_SPACE + 2 * (360 * 128 + 128)
