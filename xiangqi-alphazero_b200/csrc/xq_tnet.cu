// xq_tnet.cu -- f1: the training step of AlphaZeroTrainer.train_network (training/train.py:397-423)
// on hand-written kernels end to end: tf32 tcgen05 convolutions / dense layers (xq_tmma.cuh), BatchNorm, ReLU, residual
// adds, heads and their backward passes in the training plane layout (xq_tnet_ops.cuh).  The C ABI exposes the kernels one
// by one (include/xq_b200.h, xq_tgemm ... xq_tn_value_backward); tnet.HandStep (Python) strings the ~140 launches of a step
// together once and replays them from a CUDA graph.  No torch op and no cuDNN / cuBLAS call inside a step.
#include "xq_tmma.cuh"
#include "xq_tnet_ops.cuh"

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
    const int ks = d->k_splits > 1 ? d->k_splits : 1;
    if (d->out_rm && ks > 2) return xq_fail(c, XQ_ERR_ARG, "xq_tgemm: a row-major output takes at most 2 contraction splits (order-independent sum)");
    if (d->out && ks > 1 && (d->residual || d->out_split_stride <= 0)) return xq_fail(c, XQ_ERR_ARG, "xq_tgemm: split planes output needs out_split_stride and no residual");
    a.k_splits = ks;
    a.kb_per = (d->kblocks + ks - 1) / ks;
    a.out_split_bytes = d->out_split_stride * 4;
    if (d->out_rm && ks > 1) XQ_CUDA(c, cudaMemsetAsync(d->out_rm, 0, (size_t)d->m_rows * d->out_stride * sizeof(float), s));
    const int total = d->m_pairs * d->n_tiles * ks;
    const int grid = c->sm_count < total ? c->sm_count : total;
    if (int rc = ensure_attr(c, tg_kernel, 0, kTgSmem)) return rc;
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tg_kernel, dim3(grid), kTgThreads, (size_t)kTgSmem, s, (const TgArgs)a));
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
    constexpr int kTail = 128 + 4 * 32 * 33 * 4;                  // barriers + the epilogue's transpose tiles
    int ns = (227 * 1024 - 512 - kTail) / a.stage_bytes;
    if (ns > kTwgMaxStages) ns = kTwgMaxStages;
    if (ns < 2) return xq_fail(c, XQ_ERR_ARG, "xq_twgrad: stage of %d bytes leaves fewer than 2 pipeline stages", a.stage_bytes);
    a.n_stages = ns;
    int smem_bytes = ns * a.stage_bytes + 512 + kTail;
    if (smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;         // > half an SM: one CTA per SM (it owns all of TMEM)
    if (int rc = ensure_attr(c, twg_kernel, 2, 227 * 1024)) return rc;
    const int total = d->n_mtiles * d->n_slabs * d->n_groups;
    const int grid = c->sm_count < total ? c->sm_count : total;
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, twg_kernel, dim3(grid), kTwgThreads, (size_t)smem_bytes, s, (const TwgArgs)a));
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

// ---- the layers between the contractions (xq_tnet_ops.cuh) ---------------------------------------------------------
#define XQ_TN_ENTER(name)                                                          \
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, name ": ctx is NULL");            \
    XQ_CUDA(c, cudaSetDevice(c->device));                                          \
    cudaStream_t s = (cudaStream_t)stream;                                         \
    XqTimer tm(c, s)
#define XQ_TN_DONE()                 \
    XQ_CUDA(c, cudaGetLastError()); \
    return XQ_OK

static inline unsigned tn_blocks(long long total, int threads = 256) { return (unsigned)((total + threads - 1) / threads); }
static inline int tn_row_splits(int n_boards)
{
    const long long n = ((long long)n_boards * kTnBoard + 255) / 256;
    return (int)(n < 1 ? 1 : (n > 64 ? 64 : n));
}

extern "C" int xq_tn_input(xq_ctx* c, const float* x, int n_boards, int channels, int pairs, float* planes, float* g, int64_t rows, void* stream)
{
    XQ_TN_ENTER("xq_tn_input");
    if (!x || !planes || n_boards <= 0 || channels <= 0 || pairs <= 0 || channels > pairs * 8 || rows < kTnRow0 + (long long)n_boards * kTnBoard)
        return xq_fail(c, XQ_ERR_ARG, "xq_tn_input: bad arguments");
    tn_input_kernel<<<tn_blocks((long long)n_boards * 90 * pairs), 256, 0, s>>>(x, n_boards, channels, pairs, planes, g, rows);
    c->launches += 1;
    XQ_TN_DONE();
}

extern "C" int xq_tn_wimage(xq_ctx* c, const float* w, int co, int ci, int taps, float* img, int img_kb, int n0, int k0, int transposed,
                            void* stream)
{
    XQ_TN_ENTER("xq_tn_wimage");
    if (!w || !img || co <= 0 || ci <= 0 || (taps != 1 && taps != 9) || img_kb <= 0 || n0 < 0 || k0 < 0 || (k0 & 3) ||
        k0 + (transposed ? co : ci) > img_kb * 32)
        return xq_fail(c, XQ_ERR_ARG, "xq_tn_wimage: bad arguments");
    const int k_cnt = transposed ? co : ci, n_cnt = transposed ? ci : co;
    const int k4_cnt = (k_cnt + 3) / 4;
    const long long total = transposed ? (long long)taps * n_cnt * k4_cnt : (long long)taps * ((n_cnt + 3) / 4) * ((k4_cnt + 7) / 8) * 32;
    unsigned blocks = tn_blocks(total);
    if (blocks > (unsigned)c->sm_count * 32u) blocks = (unsigned)c->sm_count * 32u;
    tn_wimage_kernel<<<blocks, 256, 0, s>>>(w, co, ci, taps, img, img_kb, n0, k0, transposed);
    c->launches += 1;
    XQ_TN_DONE();
}

extern "C" int xq_tn_wimage_dense2(xq_ctx* c, const float* w, int co, int ci, float* img, int img_kb, float* img_t, int img_t_kb, void* stream)
{
    XQ_TN_ENTER("xq_tn_wimage_dense2");
    if (!w || !img || !img_t || co <= 0 || ci <= 0 || (ci & 31) || ci > img_kb * 32 || co > img_t_kb * 32)
        return xq_fail(c, XQ_ERR_ARG, "xq_tn_wimage_dense2: bad arguments");
    tn_wimage_dense2_kernel<<<dim3(ci / 32, (co + 31) / 32), 256, 0, s>>>(w, co, ci, img, img_kb, img_t, img_t_kb);
    c->launches += 1;
    XQ_TN_DONE();
}

extern "C" int xq_tn_wimage_batch(xq_ctx* c, const xq_tn_wimage_item* items, int n_items, void* stream)
{
    XQ_TN_ENTER("xq_tn_wimage_batch");
    if (!items || n_items <= 0) return xq_fail(c, XQ_ERR_ARG, "xq_tn_wimage_batch: bad arguments");
    for (int i0 = 0; i0 < n_items; i0 += 32) {
        TnWimageBatch bt;
        memset(&bt, 0, sizeof(bt));
        const int cnt = n_items - i0 < 32 ? n_items - i0 : 32;
        for (int i = 0; i < cnt; ++i) {
            const xq_tn_wimage_item& q = items[i0 + i];
            if (!q.w || !q.img || q.co <= 0 || q.ci <= 0 || (q.taps != 1 && q.taps != 9) || q.img_kb <= 0 || q.n0 < 0 || q.k0 < 0 || (q.k0 & 3) ||
                q.k0 + (q.transposed ? q.co : q.ci) > q.img_kb * 32)
                return xq_fail(c, XQ_ERR_ARG, "xq_tn_wimage_batch: bad item %d", i0 + i);
            bt.it[i].w = q.w; bt.it[i].img = q.img; bt.it[i].co = q.co; bt.it[i].ci = q.ci; bt.it[i].taps = q.taps; bt.it[i].img_kb = q.img_kb;
            bt.it[i].n0 = q.n0; bt.it[i].k0 = q.k0; bt.it[i].transposed = q.transposed;
        }
        tn_wimage_batch_kernel<<<dim3(16, cnt), 256, 0, s>>>(bt);
        c->launches += 1;
    }
    XQ_TN_DONE();
}

extern "C" int xq_tn_bn_forward(xq_ctx* c, const xq_tn_bn_desc* d, void* stream)
{
    XQ_TN_ENTER("xq_tn_bn_forward");
    if (!d || !d->y || !d->out || !d->partial || !d->gamma || !d->beta || !d->running_mean || !d->running_var || !d->save || d->n_boards <= 0 ||
        d->n_channels <= 0 || (d->chunk0 & 1) || d->rows < kTnRow0 + (long long)d->n_boards * kTnBoard)
        return xq_fail(c, XQ_ERR_ARG, "xq_tn_bn_forward: bad descriptor");
    const int pairs = (d->n_channels + 7) / 8;
    TnBnArgs a;
    a.y = d->y; a.res = d->res; a.out = d->out; a.out_g = d->out_g; a.rows = d->rows; a.n_boards = d->n_boards; a.chunk0 = d->chunk0;
    a.n_channels = d->n_channels; a.relu = d->relu; a.partial = d->partial; a.gamma = d->gamma; a.beta = d->beta;
    a.running_mean = d->running_mean; a.running_var = d->running_var; a.save = d->save; a.eps = d->eps; a.momentum = d->momentum;
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_bn_stat_kernel, dim3(2 * pairs, kTnSplit), 256, (size_t)0, s, d->y, (long long)d->rows, d->n_boards, d->chunk0,
                             d->partial));
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_bn_apply_kernel, dim3(pairs, tn_row_splits(d->n_boards)), 256, (size_t)0, s, (const TnBnArgs)a));
    c->launches += 2;
    XQ_TN_DONE();
}

extern "C" int xq_tn_bn_backward(xq_ctx* c, const xq_tn_bn_bwd_desc* d, void* stream)
{
    XQ_TN_ENTER("xq_tn_bn_backward");
    if (!d || !d->dout || !d->act || !d->y || !d->save || !d->partial || !d->gamma || !d->dgamma || !d->dbeta || !d->dy || d->n_boards <= 0 ||
        d->n_channels <= 0 || (d->chunk0 & 1) || d->rows < kTnRow0 + (long long)d->n_boards * kTnBoard)
        return xq_fail(c, XQ_ERR_ARG, "xq_tn_bn_backward: bad descriptor");
    const int pairs = (d->n_channels + 7) / 8;
    TnBnBwdArgs a;
    a.dout = d->dout; a.act = d->act; a.y = d->y; a.rows = d->rows; a.n_boards = d->n_boards; a.chunk0 = d->chunk0; a.n_channels = d->n_channels;
    a.relu = d->relu; a.save = d->save; a.partial = d->partial; a.gamma = d->gamma; a.dgamma = d->dgamma; a.dbeta = d->dbeta; a.dy = d->dy;
    a.dy_g = d->dy_g; a.dskip = d->dskip;
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_bn_bwd_stat_kernel, dim3(2 * pairs, kTnSplit), 256, (size_t)0, s, (const TnBnBwdArgs)a));
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_bn_bwd_apply_kernel, dim3(pairs, tn_row_splits(d->n_boards)), 256, (size_t)0, s, (const TnBnBwdArgs)a));
    c->launches += 2;
    XQ_TN_DONE();
}

extern "C" int xq_tn_wgrad_reduce(xq_ctx* c, const float* ws, int slabs, int64_t slab_stride, int taps, int ldn, int m_cnt, int n_cnt, int n_src0,
                                  int transposed, float* dw, int ci_total, int co0, int ci0, void* stream)
{
    XQ_TN_ENTER("xq_tn_wgrad_reduce");
    if (!ws || !dw || slabs <= 0 || taps <= 0 || m_cnt <= 0 || m_cnt > 128 || n_cnt <= 0 || n_src0 < 0 || n_src0 + n_cnt > ldn || ci_total <= 0)
        return xq_fail(c, XQ_ERR_ARG, "xq_tn_wgrad_reduce: bad arguments");
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_wgrad_reduce_kernel, dim3(tn_blocks((long long)taps * m_cnt * n_cnt)), 256, (size_t)0, s, ws, slabs, slab_stride, taps, ldn, m_cnt, n_cnt, n_src0, transposed,
                                                                                      dw, ci_total, co0, ci0));
    c->launches += 1;
    XQ_TN_DONE();
}

extern "C" int xq_tn_flatten(xq_ctx* c, const float* act, int64_t rows, int n_boards, int channels, float* dense, float* dense_g, int64_t drows,
                             void* stream)
{
    XQ_TN_ENTER("xq_tn_flatten");
    if (!act || !dense || n_boards <= 0 || channels <= 0 || (channels * 90) % 4 || drows < kTnRow0 + n_boards)
        return xq_fail(c, XQ_ERR_ARG, "xq_tn_flatten: bad arguments");
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_flatten_kernel, dim3(tn_blocks((long long)channels * 90 / 4 * n_boards)), 256, (size_t)0, s, act, rows, n_boards, channels, dense, dense_g, drows));
    c->launches += 1;
    XQ_TN_DONE();
}

extern "C" int xq_tn_unflatten(xq_ctx* c, const float* dense, int64_t drows, int n_boards, int channels, float* planes, int64_t rows,
                               int n_partials, int64_t part_stride, void* stream)
{
    XQ_TN_ENTER("xq_tn_unflatten");
    if (!dense || !planes || n_boards <= 0 || channels <= 0 || (channels & 3) || drows < kTnRow0 + n_boards || n_partials < 1)
        return xq_fail(c, XQ_ERR_ARG, "xq_tn_unflatten: bad arguments");
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_unflatten_kernel, dim3(tn_blocks((long long)n_boards * 90 * (channels / 4))), 256, (size_t)0, s, dense, drows, n_boards, channels, planes, rows,
                                                                                         n_partials, part_stride));
    c->launches += 1;
    XQ_TN_DONE();
}

extern "C" int xq_tn_rows_layouts(xq_ctx* c, const float* m, int64_t stride, int n_rows, int n_cols, float* dense, float* dense_g, int64_t drows,
                                  void* stream)
{
    XQ_TN_ENTER("xq_tn_rows_layouts");
    if (!m || (!dense && !dense_g) || n_rows <= 0 || n_cols <= 0 || (n_cols & 3) || (stride & 3) || drows < kTnRow0 + n_rows)
        return xq_fail(c, XQ_ERR_ARG, "xq_tn_rows_layouts: bad arguments");
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_rows_layouts_kernel, dim3(tn_blocks((long long)(n_cols / 4) * n_rows)), 256, (size_t)0, s, m, stride, n_rows, n_cols, dense, dense_g, drows));
    c->launches += 1;
    XQ_TN_DONE();
}

extern "C" int xq_tn_colsum(xq_ctx* c, const float* m, int64_t stride, int n_rows, int n_cols, float* out, void* stream)
{
    XQ_TN_ENTER("xq_tn_colsum");
    if (!m || !out || n_rows <= 0 || n_cols <= 0) return xq_fail(c, XQ_ERR_ARG, "xq_tn_colsum: bad arguments");
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_colsum_kernel, dim3(tn_blocks(n_cols, 32)), 256, (size_t)0, s, m, stride, n_rows, n_cols, out));
    c->launches += 1;
    XQ_TN_DONE();
}

extern "C" int xq_tn_value_forward(xq_ctx* c, const float* act, int64_t rows, int chunk, int n_boards, const float* w1, const float* b1,
                                   const float* w2, const float* b2, float* h, float* v, void* stream)
{
    XQ_TN_ENTER("xq_tn_value_forward");
    if (!act || !w1 || !b1 || !w2 || !b2 || !h || !v || n_boards <= 0 || chunk < 0) return xq_fail(c, XQ_ERR_ARG, "xq_tn_value_forward: bad arguments");
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_value_fwd_kernel, dim3(n_boards), 256, (size_t)0, s, act, rows, chunk, w1, b1, w2, b2, h, v));
    c->launches += 1;
    XQ_TN_DONE();
}

extern "C" int xq_tn_value_backward(xq_ctx* c, const float* act, int64_t rows, int chunk, int n_boards, const float* w1, const float* w2,
                                    const float* h, const float* v, const float* g_value, float* dh, float* dpre, float* dact, float* dw1,
                                    float* db1, float* dw2, float* db2, void* stream)
{
    XQ_TN_ENTER("xq_tn_value_backward");
    if (!act || !w1 || !w2 || !h || !v || !g_value || !dh || !dpre || !dact || !dw1 || !db1 || !dw2 || !db2 || n_boards <= 0 || chunk < 0)
        return xq_fail(c, XQ_ERR_ARG, "xq_tn_value_backward: bad arguments");
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_value_bwd_a_kernel, dim3(n_boards), 384, (size_t)0, s, w1, w2, h, v, g_value, dh, dpre, dact, rows, chunk));
    XQ_CUDA(c, xq_launch_pdl(c->train_pdl, tn_value_bwd_w_kernel, dim3(kTnVH + 1), 384, (size_t)0, s, act, rows, chunk, n_boards, h, dh, dpre, dw1, db1, dw2, db2));
    c->launches += 2;
    XQ_TN_DONE();
}
