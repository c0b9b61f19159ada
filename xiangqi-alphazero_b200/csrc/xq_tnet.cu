// xq_tnet.cu -- f1: the training step of AlphaZeroTrainer.train_network (training/train.py:397-423)
// on hand-written kernels end to end: tf32 tcgen05 convolutions / dense layers (xq_tmma.cuh), BatchNorm, ReLU, residual
// adds, heads, loss and their backward passes in the training plane layout (xq_tnet_ops.cuh), driven from C++ (one host
// call per forward+backward, ~130 launches, no torch op and no cuDNN / cuBLAS call inside).
#include "xq_tmma.cuh"

#include <cstdlib>
#include <cstring>

using namespace xq;
using namespace xq::tn;

namespace {

struct TnState {
    uint32_t attr_done = 0;
};

TnState* tn_state(xq_ctx* c)
{
    if (!c->tnet) {
        c->tnet = new TnState();
    }
    return reinterpret_cast<TnState*>(c->tnet);
}

template <class K>
int ensure_attr(xq_ctx* c, K kern, int bit, int bytes)
{
    TnState* T = tn_state(c);
    if (!(T->attr_done & (1u << bit))) {
        XQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        T->attr_done |= 1u << bit;
    }
    return XQ_OK;
}

int launch_tg(xq_ctx* c, const xq_tgemm_desc* d, cudaStream_t s)
{
    if (!d || !d->a || !d->w || d->kblocks <= 0 || (d->ntaps != 1 && d->ntaps != 9) || d->m_pairs <= 0 || d->n_tiles <= 0 ||
        (!d->out && !d->out_rm) || d->a_row0 < kTHalo)
        return xq_fail(c, XQ_ERR_ARG, "xq_tgemm: bad descriptor");
    if (d->out_rm && ((d->out_stride & 3) || (d->n_cols & 3))) return xq_fail(c, XQ_ERR_ARG, "xq_tgemm: row-major output needs stride and n_cols multiples of 4");
    TgArgs a;
    a.a = (const uint8_t*)d->a;
    a.a_rows = d->a_rows;
    a.a_row0 = d->a_row0;
    a.w = (const uint8_t*)d->w;
    a.kblocks = d->kblocks;
    a.ntaps = d->ntaps;
    a.img_kb = d->img_kb;
    a.shift_sign = d->shift_sign;
    a.m_pairs = d->m_pairs;
    a.n_tiles = d->n_tiles;
    a.out_chunks = d->out_chunks;
    a.n_cols = d->n_cols;
    a.m_rows = d->m_rows;
    a.out = (uint8_t*)d->out;
    a.out_rows = d->out_rows;
    a.out_row0 = d->out_row0;
    a.residual = (const uint8_t*)d->residual;
    a.out_rm = d->out_rm;
    a.out_stride = d->out_stride;
    a.bias = d->bias;
    const int total = d->m_pairs * d->n_tiles;
    const int grid = c->sm_count < total ? c->sm_count : total;
    if (int rc = ensure_attr(c, tg_kernel, 0, kTgSmem)) return rc;
    tg_kernel<<<grid, kTgThreads, kTgSmem, s>>>(a);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

int launch_twg(xq_ctx* c, const xq_twgrad_desc* d, cudaStream_t s)
{
    if (!d || !d->a || !d->b || !d->out || d->nbg < 1 || d->nbg > 4 || d->kr < 8 || (d->kr & 7) || d->stages_per_item <= 0 ||
        d->n_slabs <= 0 || d->n_groups <= 0 || d->n_groups > 4 * 64 || d->n_mtiles <= 0 || d->taps_per_group < 1 || d->taps_per_group > 4 ||
        d->b_rows_stage < d->kr || (d->b_rows_stage & 3) || d->b_groups_stage < d->nbg || d->b_groups_stage > 28 || (d->a_row0 & 3) || (d->b_row0 & 3))
        return xq_fail(c, XQ_ERR_ARG, "xq_twgrad: bad descriptor");
    const int N = 32 * d->nbg, tcol = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    if (tcol * d->taps_per_group > 512) return xq_fail(c, XQ_ERR_ARG, "xq_twgrad: %d taps of %d columns exceed TMEM", d->taps_per_group, N);
    TwgArgs a;
    a.a = (const uint8_t*)d->a;
    a.b = (const uint8_t*)d->b;
    a.a_rows = d->a_rows;
    a.a_row0 = d->a_row0;
    a.b_rows = d->b_rows;
    a.b_row0 = d->b_row0;
    a.a_group0 = d->a_group0;
    a.b_group0 = d->b_group0;
    a.nbg = d->nbg;
    a.kr = d->kr;
    a.stages_per_item = d->stages_per_item;
    a.n_slabs = d->n_slabs;
    a.n_groups = d->n_groups;
    a.n_mtiles = d->n_mtiles;
    a.taps_per_group = d->taps_per_group;
    a.b_rows_stage = d->b_rows_stage;
    a.b_groups_stage = d->b_groups_stage;
    a.b_group_step = d->b_group_step;
    for (int i = 0; i < 4; ++i) {
        if (d->b_row_lo[i] & 3) return xq_fail(c, XQ_ERR_ARG, "xq_twgrad: b_row_lo must be multiples of 4 (swizzle phase)");
        a.b_row_lo[i] = d->b_row_lo[i];
    }
    for (int i = 0; i < 16; ++i) a.tap_off[i] = d->tap_off[i];
    a.out = d->out;
    a.mt_stride = d->mt_stride;
    a.slab_stride = d->slab_stride;
    a.g_stride = d->g_stride;
    a.tap_stride = d->tap_stride;
    a.ldo = d->ldo;
    a.m_limit = d->m_limit;
    a.n_limit = d->n_limit;
    a.g_cols = d->g_cols;
    a.t_cols = d->t_cols;
    a.a_stage_bytes = 4 * d->kr * 128;
    a.stage_bytes = a.a_stage_bytes + d->b_groups_stage * d->b_rows_stage * 128;      // multiples of 512: kr % 8 == 0, b_rows_stage % 4 == 0
    int ns = (224 * 1024) / a.stage_bytes;
    if (ns > kTwgMaxStages) ns = kTwgMaxStages;
    if (ns < 2) return xq_fail(c, XQ_ERR_ARG, "xq_twgrad: stage of %d bytes leaves fewer than 2 pipeline stages", a.stage_bytes);
    a.n_stages = ns;
    int smem_bytes = ns * a.stage_bytes + 512 + 256;
    if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;         // > half an SM: one CTA per SM (it owns all of TMEM)
    if (int rc = ensure_attr(c, twg_kernel, 2, 227 * 1024)) return rc;
    const int total = d->n_mtiles * d->n_slabs * d->n_groups;
    const int grid = c->sm_count < total ? c->sm_count : total;
    twg_kernel<<<grid, kTwgThreads, smem_bytes, s>>>(a);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

}  // namespace

extern "C" void xq_tnet_free_(xq_ctx* c)
{
    if (!c || !c->tnet) return;
    delete reinterpret_cast<TnState*>(c->tnet);
    c->tnet = nullptr;
}

extern "C" int xq_tgemm(xq_ctx* c, const xq_tgemm_desc* d, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_tgemm: ctx is NULL");
    XQ_CUDA(c, cudaSetDevice(c->device));
    XqTimer tm(c, (cudaStream_t)stream);
    return launch_tg(c, d, (cudaStream_t)stream);
}

extern "C" int xq_twgrad(xq_ctx* c, const xq_twgrad_desc* d, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_twgrad: ctx is NULL");
    XQ_CUDA(c, cudaSetDevice(c->device));
    XqTimer tm(c, (cudaStream_t)stream);
    return launch_twg(c, d, (cudaStream_t)stream);
}
