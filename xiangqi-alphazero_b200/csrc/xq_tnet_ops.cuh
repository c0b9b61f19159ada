// xq_tnet_ops.cuh -- everything of the training step that is not a contraction (included by xq_tnet.cu only).
//
// AlphaZeroTrainer.train_network (training/train.py:397-423) runs XiangqiNet (model.py:39-107) forward and backward
// through torch; here the activations live in the training plane layout of xq_tmma.cuh and these kernels are the layers
// between the tensor-core contractions: layout writers, training-mode BatchNorm (+ residual + ReLU) forward and backward,
// the weight images, the slab reduction of the convolution weight gradients, flatten / unflatten around the policy FC
// and the value head's two small dense layers.  All reductions run in a fixed order (no floating-point atomics): two
// runs of a step give bit-identical gradients, which keeps replicated data-parallel ranks identical.
//
// Plane layout: P[chunk][rows][4] float32, board b cell (r, c) at row row0 + b*110 + r*10 + c; pad cells (c = 9, r = 10)
// hold zeros in every tensor a contraction reads.  G layout: G[group][rows][32], 32-byte unit u of row r at u ^ (r & 3).
#pragma once
#include "xq_ctx.h"

namespace xq {
namespace tn {

constexpr int kTnRow0 = XQ_T_ROW0;
constexpr int kTnBoard = XQ_T_BOARD_ROWS;        // 110
constexpr int kTnSplit = 16;                     // row splits of the statistics kernels (partial sums per chunk)

__device__ __forceinline__ bool tn_real(int rel, int n_boards)          // rel = row - row0, 0 <= rel (32-bit: a step holds < 2^31 rows)
{
    const unsigned b = (unsigned)rel / (unsigned)kTnBoard;
    const unsigned r = (unsigned)rel - b * (unsigned)kTnBoard;
    return b < (unsigned)n_boards && r < 100u && (r % 10u) != 9u;
}
// float offset of the 8 floats of chunk pair `cp` (channels 8*cp .. 8*cp+7) of row `row` in a G tensor of `rows` rows
__device__ __forceinline__ size_t tn_g_off(int cp, long long row, long long rows)
{
    const int group = cp >> 2, u = cp & 3;
    return ((size_t)group * rows + (size_t)row) * 32 + (size_t)((u ^ (int)(row & 3)) * 8);
}
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// sums 8 doubles per thread over the block (256 threads), result valid in threads 0..7 (value j in thread j)
__device__ __forceinline__ double tn_block_sum8(double (&v)[8], double (*sm)[8])
{
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int o = 16; o; o >>= 1) v[j] += __shfl_xor_sync(0xffffffffu, v[j], o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0)
        for (int j = 0; j < 8; ++j) sm[w][j] = v[j];
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 8)
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i][threadIdx.x];
    return t;
}

// ---- NCHW -> planes + G ------------------------------------------------------------------------------------
// x[b][c][10][9] -> planes chunks [0, 2*pairs) and G (optional); thread = (board, cell, chunk pair).  Pad cells are not
// written (they are zero since allocation).
__global__ void __launch_bounds__(256) tn_input_kernel(const float* __restrict__ x, int n_boards, int channels, int pairs,
                                                       float* __restrict__ planes, float* __restrict__ g, long long rows)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n_boards * 90 * pairs;
    if (idx >= total) return;
    const int cell = (int)(idx % 90);
    const long long rest = idx / 90;
    const int b = (int)(rest % n_boards), cp = (int)(rest / n_boards);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ch = cp * 8 + j;
        v[j] = ch < channels ? x[((size_t)b * channels + ch) * 90 + cell] : 0.0f;
    }
    const long long row = kTnRow0 + (long long)b * kTnBoard + (cell / 9) * 10 + cell % 9;
    st4(planes + ((size_t)(2 * cp) * rows + row) * 4, make_float4(v[0], v[1], v[2], v[3]));
    st4(planes + ((size_t)(2 * cp + 1) * rows + row) * 4, make_float4(v[4], v[5], v[6], v[7]));
    if (g) {
        float* d = g + tn_g_off(cp, row, rows);
        st4(d, make_float4(v[0], v[1], v[2], v[3]));
        st4(d + 4, make_float4(v[4], v[5], v[6], v[7]));
    }
}

// ---- weight images -----------------------------------------------------------------------------------------
// w[co][ci][taps] -> image [nt][taps][img_kb][8 chunks][128 n][4 k].  transposed = 0: n = n0 + co, k = k0 + ci (fprop);
// transposed = 1: n = n0 + ci, k = k0 + co (dgrad).  Only the positions of real weights are written.
__global__ void __launch_bounds__(256) tn_wimage_kernel(const float* __restrict__ w, int co, int ci, int taps, float* __restrict__ img,
                                                        int img_kb, int n0, int k0, int transposed)
{
    const int n_cnt = transposed ? ci : co, k_cnt = transposed ? co : ci;
    const int k4_cnt = (k_cnt + 3) >> 2;
    // transposed: a warp takes 32 consecutive n (= ci, contiguous in w for taps = 1) of one k4: 512 contiguous image bytes.
    // else: a warp takes 4 n x 8 k4: 128 contiguous bytes of w per n, 64 contiguous image bytes per k4.
    const int n4_cnt = (n_cnt + 3) >> 2, k32_cnt = (k4_cnt + 7) >> 3;
    const long long total = transposed ? (long long)taps * n_cnt * k4_cnt : (long long)taps * n4_cnt * k32_cnt * 32;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int n, k4, tap;
        if (transposed) {
            n = (int)(idx % n_cnt);
            const long long r = idx / n_cnt;
            k4 = (int)(r % k4_cnt);
            tap = (int)(r / k4_cnt);
        } else {
            const int l = (int)(idx & 31);
            long long r = idx >> 5;
            k4 = (int)(r % k32_cnt) * 8 + (l & 7);
            r /= k32_cnt;
            n = (int)(r % n4_cnt) * 4 + (l >> 3);
            tap = (int)(r / n4_cnt);
            if (n >= n_cnt || k4 >= k4_cnt) continue;
        }
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k4 * 4 + j;
            const int o = transposed ? k : n, i = transposed ? n : k;
            v[j] = k < k_cnt ? w[((size_t)o * ci + i) * taps + tap] : 0.0f;
        }
        const int N = n0 + n, K = k0 + k4 * 4;
        const int nt = N >> 7, nn = N & 127, kb = K >> 5, chunk = (K >> 2) & 7;
        st4(img + ((((size_t)(nt * taps + tap) * img_kb + kb) * 8 + chunk) * 128 + nn) * 4, make_float4(v[0], v[1], v[2], v[3]));
    }
}

// up to 32 small weight tensors in one launch (blockIdx.y = item): the 3x3 / 1x1 convolutions of the step
struct TnWimageItem {
    const float* w;
    float* img;
    int co, ci, taps, img_kb, n0, k0, transposed, pad;
};
struct TnWimageBatch {
    TnWimageItem it[32];
};
__global__ void __launch_bounds__(256) tn_wimage_batch_kernel(const TnWimageBatch bt)
{
    const TnWimageItem& q = bt.it[blockIdx.y];
    const int n_cnt = q.transposed ? q.ci : q.co, k_cnt = q.transposed ? q.co : q.ci;
    const int k4_cnt = (k_cnt + 3) >> 2;
    // a thread takes one (tap-independent) position (n, k4) and walks the taps: the 4 x taps weights it reads are contiguous
    // for transposed = 0 (w[n][k4*4 .. +3][tap])
    const int total = n_cnt * k4_cnt;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        int n, k4;
        if (q.transposed) {
            n = idx % n_cnt;
            k4 = idx / n_cnt;
        } else {
            k4 = idx % k4_cnt;
            n = idx / k4_cnt;
        }
        const int N = q.n0 + n, K = q.k0 + k4 * 4;
        const int nt = N >> 7, nn = N & 127, kb = K >> 5, chunk = (K >> 2) & 7;
        for (int tap = 0; tap < q.taps; ++tap) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k4 * 4 + j;
                const int o = q.transposed ? k : n, i = q.transposed ? n : k;
                v[j] = k < k_cnt ? q.w[((size_t)o * q.ci + i) * q.taps + tap] : 0.0f;
            }
            st4(q.img + ((((size_t)(nt * q.taps + tap) * q.img_kb + kb) * 8 + chunk) * 128 + nn) * 4, make_float4(v[0], v[1], v[2], v[3]));
        }
    }
}

// Both images of a dense layer's weight w[co][ci] (ci % 32 == 0) from ONE pass over it: 32 x 32 tiles through shared memory;
// img: n = co, k = ci (fprop), img_t: n = ci, k = co (dgrad).  The 93 MB policy FC weight is read once instead of twice.
__global__ void __launch_bounds__(256) tn_wimage_dense2_kernel(const float* __restrict__ w, int co, int ci, float* __restrict__ img, int img_kb,
                                                               float* __restrict__ img_t, int img_t_kb)
{
    __shared__ float tile[32][33];
    const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
    const int r = threadIdx.x >> 3, c4 = (threadIdx.x & 7) * 4;
    {
        const int o = co0 + r;
        float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (o < co) {
            v = ld4(w + (size_t)o * ci + ci0 + c4);
            const int K = ci0 + c4;
            st4(img + ((((size_t)(o >> 7) * img_kb + (K >> 5)) * 8 + ((K >> 2) & 7)) * 128 + (o & 127)) * 4, v);
        }
        tile[r][c4] = v.x; tile[r][c4 + 1] = v.y; tile[r][c4 + 2] = v.z; tile[r][c4 + 3] = v.w;
    }
    __syncthreads();
    {
        const int i = ci0 + r, K = co0 + c4;                    // r = input channel within the tile, c4 = first of 4 output channels
        if (K < co)                                             // (rows of the tile beyond co hold zeros: a partial group is zero padded)
            st4(img_t + ((((size_t)(i >> 7) * img_t_kb + (K >> 5)) * 8 + ((K >> 2) & 7)) * 128 + (i & 127)) * 4,
                make_float4(tile[c4][r], tile[c4 + 1][r], tile[c4 + 2][r], tile[c4 + 3][r]));
    }
}

// ---- BatchNorm forward -------------------------------------------------------------------------------------
// partial[chunk][split][8] = per-block sums of y (4 channels) and y^2 over the real cells of the block's rows
__global__ void __launch_bounds__(256) tn_bn_stat_kernel(const float* __restrict__ y, long long rows, int n_boards, int chunk0,
                                                         double* __restrict__ partial)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    __shared__ double sm[8][8];
    const int chunk = blockIdx.x, sp = blockIdx.y;
    const long long n_rows = (long long)n_boards * kTnBoard;
    const float* src = y + ((size_t)(chunk0 + chunk) * rows + kTnRow0) * 4;
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    const int stride = gridDim.y * 256;
    for (int rel0 = sp * 256 + threadIdx.x; rel0 < (int)n_rows; rel0 += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {                        // four independent loads in flight
            const int rel = rel0 + u * stride;
            v[u] = (rel < (int)n_rows && tn_real(rel, n_boards)) ? ld4(src + (size_t)rel * 4) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s[0] += v[u].x; s[1] += v[u].y; s[2] += v[u].z; s[3] += v[u].w;
            q[0] += v[u].x * v[u].x; q[1] += v[u].y * v[u].y; q[2] += v[u].z * v[u].z; q[3] += v[u].w * v[u].w;
        }
    }
    double d[8] = {s[0], s[1], s[2], s[3], q[0], q[1], q[2], q[3]};
    const double t = tn_block_sum8(d, sm);
    if (threadIdx.x < 8) partial[((size_t)chunk * gridDim.y + sp) * 8 + threadIdx.x] = t;
}

struct TnBnArgs {
    const float* y;          // conv output planes (pre-normalisation)
    const float* res;        // residual planes added before the ReLU (or null)
    float* out;              // activation planes
    float* out_g;            // activation in the G layout (or null)
    long long rows;
    int n_boards, chunk0, n_channels, relu;     // chunk0: first chunk of the tensors this call covers (even); channel c of the
                                                // call = chunk0*4 + c in the tensors, c in the parameter vectors
    const double* partial;
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    float* save;             // [2][n_channels]: batch mean, 1/sqrt(var + eps)
    float eps, momentum;
};

// grid (chunk pairs, row splits)
__global__ void __launch_bounds__(256) tn_bn_apply_kernel(const TnBnArgs p)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    __shared__ float sa[8], sb[8];
    const int cp = blockIdx.x;
    if (threadIdx.x < 8) {
        const int c = cp * 8 + threadIdx.x;                 // channel within the call
        const int chunk = c >> 2, j = c & 3;
        double s = 0.0, q = 0.0;
        for (int k = 0; k < kTnSplit; ++k) {
            s += p.partial[((size_t)chunk * kTnSplit + k) * 8 + j];
            q += p.partial[((size_t)chunk * kTnSplit + k) * 8 + 4 + j];
        }
        float a = 0.0f, b = 0.0f;
        if (c < p.n_channels) {
            const double n = (double)p.n_boards * 90.0;
            const double mean = s / n;
            double var = q / n - mean * mean;
            if (var < 0.0) var = 0.0;
            const double inv = 1.0 / sqrt(var + (double)p.eps);
            a = (float)((double)p.gamma[c] * inv);
            b = (float)((double)p.beta[c] - mean * (double)p.gamma[c] * inv);
            if (blockIdx.y == 0) {
                p.save[c] = (float)mean;
                p.save[p.n_channels + c] = (float)inv;
                const double m = (double)p.momentum;
                p.running_mean[c] = (float)((1.0 - m) * (double)p.running_mean[c] + m * mean);
                p.running_var[c] = (float)((1.0 - m) * (double)p.running_var[c] + m * var * (n / (n - 1.0)));
            }
        }
        sa[threadIdx.x] = a;
        sb[threadIdx.x] = b;
    }
    __syncthreads();
    const long long n_rows = (long long)p.n_boards * kTnBoard;
    const int c0 = p.chunk0 + 2 * cp;
    for (int rel = blockIdx.y * 256 + threadIdx.x; rel < (int)n_rows; rel += gridDim.y * 256) {
        const long long row = kTnRow0 + rel;
        float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (tn_real(rel, p.n_boards)) {
            const float4 y0 = ld4(p.y + ((size_t)c0 * p.rows + row) * 4), y1 = ld4(p.y + ((size_t)(c0 + 1) * p.rows + row) * 4);
            v[0] = y0.x; v[1] = y0.y; v[2] = y0.z; v[3] = y0.w; v[4] = y1.x; v[5] = y1.y; v[6] = y1.z; v[7] = y1.w;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = sa[j] * v[j] + sb[j];
            if (p.res) {
                const float4 r0 = ld4(p.res + ((size_t)c0 * p.rows + row) * 4), r1 = ld4(p.res + ((size_t)(c0 + 1) * p.rows + row) * 4);
                v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
            }
            if (p.relu)
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        st4(p.out + ((size_t)c0 * p.rows + row) * 4, make_float4(v[0], v[1], v[2], v[3]));
        st4(p.out + ((size_t)(c0 + 1) * p.rows + row) * 4, make_float4(v[4], v[5], v[6], v[7]));
        if (p.out_g) {
            float* d = p.out_g + tn_g_off(c0 >> 1, row, p.rows);
            st4(d, make_float4(v[0], v[1], v[2], v[3]));
            st4(d + 4, make_float4(v[4], v[5], v[6], v[7]));
        }
    }
}

// ---- BatchNorm backward ------------------------------------------------------------------------------------
struct TnBnBwdArgs {
    const float* dout;       // gradient w.r.t. the layer's activation (planes; pad cells may hold anything)
    const float* act;        // the activation itself (ReLU mask: act > 0)
    const float* y;          // conv output planes
    long long rows;
    int n_boards, chunk0, n_channels, relu;
    const float* save;       // [2][n_channels]
    double* partial;         // [chunks][kTnSplit][8]: sums of dz and dz * xhat
    const float* gamma;
    float* dgamma;
    float* dbeta;
    float* dy;               // gradient w.r.t. the conv output: planes ...
    float* dy_g;             // ... and G layout (pad cells zero)
    float* dskip;            // dz planes: the gradient of the residual input (or null)
};

__global__ void __launch_bounds__(256) tn_bn_bwd_stat_kernel(const TnBnBwdArgs p)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    __shared__ double sm[8][8];
    const int chunk = blockIdx.x, sp = blockIdx.y;
    const long long n_rows = (long long)p.n_boards * kTnBoard;
    const size_t base = ((size_t)(p.chunk0 + chunk) * p.rows + kTnRow0) * 4;
    float mean[4], inv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = chunk * 4 + j;
        mean[j] = c < p.n_channels ? p.save[c] : 0.0f;
        inv[j] = c < p.n_channels ? p.save[p.n_channels + c] : 0.0f;
    }
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    const int stride = gridDim.y * 256;
    const float4 zero4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int rel0 = sp * 256 + threadIdx.x; rel0 < (int)n_rows; rel0 += 2 * stride) {
        float4 g4[2], a4[2], y4[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {                        // six independent loads in flight
            const int rel = rel0 + u * stride;
            const bool ok = rel < (int)n_rows && tn_real(rel, p.n_boards);
            g4[u] = ok ? ld4(p.dout + base + (size_t)rel * 4) : zero4;    // a pad cell contributes dz = 0
            a4[u] = ok ? ld4(p.act + base + (size_t)rel * 4) : zero4;
            y4[u] = ok ? ld4(p.y + base + (size_t)rel * 4) : zero4;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const float g[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w}, a[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w},
                        yy[4] = {y4[u].x, y4[u].y, y4[u].z, y4[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float dz = (!p.relu || a[j] > 0.0f) ? g[j] : 0.0f;
                s[j] += dz;
                q[j] += dz * ((yy[j] - mean[j]) * inv[j]);
            }
        }
    }
    double d[8] = {s[0], s[1], s[2], s[3], q[0], q[1], q[2], q[3]};
    const double t = tn_block_sum8(d, sm);
    if (threadIdx.x < 8) p.partial[((size_t)chunk * gridDim.y + sp) * 8 + threadIdx.x] = t;
}

__global__ void __launch_bounds__(256) tn_bn_bwd_apply_kernel(const TnBnBwdArgs p)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    __shared__ float s_mean[8], s_inv[8], s_k[8], s_m1[8], s_m2[8];
    const int cp = blockIdx.x;
    if (threadIdx.x < 8) {
        const int c = cp * 8 + threadIdx.x;
        const int chunk = c >> 2, j = c & 3;
        double s = 0.0, q = 0.0;
        for (int k = 0; k < kTnSplit; ++k) {
            s += p.partial[((size_t)chunk * kTnSplit + k) * 8 + j];
            q += p.partial[((size_t)chunk * kTnSplit + k) * 8 + 4 + j];
        }
        const bool ok = c < p.n_channels;
        const double n = (double)p.n_boards * 90.0;
        s_mean[threadIdx.x] = ok ? p.save[c] : 0.0f;
        s_inv[threadIdx.x] = ok ? p.save[p.n_channels + c] : 0.0f;
        s_k[threadIdx.x] = ok ? p.gamma[c] * p.save[p.n_channels + c] : 0.0f;
        s_m1[threadIdx.x] = (float)(s / n);
        s_m2[threadIdx.x] = (float)(q / n);
        if (ok && blockIdx.y == 0) {
            p.dbeta[c] = (float)s;
            p.dgamma[c] = (float)q;
        }
    }
    __syncthreads();
    const long long n_rows = (long long)p.n_boards * kTnBoard;
    const int c0 = p.chunk0 + 2 * cp;
    for (int rel = blockIdx.y * 256 + threadIdx.x; rel < (int)n_rows; rel += gridDim.y * 256) {
        const long long row = kTnRow0 + rel;
        const size_t o0 = ((size_t)c0 * p.rows + row) * 4, o1 = ((size_t)(c0 + 1) * p.rows + row) * 4;
        float dz[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dy[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (tn_real(rel, p.n_boards)) {
            const float4 g0 = ld4(p.dout + o0), g1 = ld4(p.dout + o1), a0 = ld4(p.act + o0), a1 = ld4(p.act + o1), y0 = ld4(p.y + o0),
                         y1 = ld4(p.y + o1);
            const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w},
                        yy[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                dz[j] = (!p.relu || a[j] > 0.0f) ? g[j] : 0.0f;
                const float xhat = (yy[j] - s_mean[j]) * s_inv[j];
                dy[j] = s_k[j] * (dz[j] - s_m1[j] - xhat * s_m2[j]);
            }
        }
        st4(p.dy + o0, make_float4(dy[0], dy[1], dy[2], dy[3]));
        st4(p.dy + o1, make_float4(dy[4], dy[5], dy[6], dy[7]));
        if (p.dy_g) {
            float* d = p.dy_g + tn_g_off(c0 >> 1, row, p.rows);
            st4(d, make_float4(dy[0], dy[1], dy[2], dy[3]));
            st4(d + 4, make_float4(dy[4], dy[5], dy[6], dy[7]));
        }
        if (p.dskip) {
            st4(p.dskip + o0, make_float4(dz[0], dz[1], dz[2], dz[3]));
            st4(p.dskip + o1, make_float4(dz[4], dz[5], dz[6], dz[7]));
        }
    }
}

// ---- slab reduction of the weight gradients -------------------------------------------------------------------
// ws[slab][tap][128 m][ldn] -> dw[(co0 + co) * ci_total + ci0 + ci][tap]; transposed = 0: (co, ci) = (m, n - n_src0);
// transposed = 1: (ci, co) = (m, n - n_src0).  Slabs are added in order.
__global__ void __launch_bounds__(256) tn_wgrad_reduce_kernel(const float* __restrict__ ws, int slabs, long long slab_stride, int taps, int ldn,
                                                              int m_cnt, int n_cnt, int n_src0, int transposed, float* __restrict__ dw,
                                                              int ci_total, int co0, int ci0)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    const long long total = (long long)taps * m_cnt * n_cnt;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int n = (int)(idx % n_cnt);
    const long long r = idx / n_cnt;
    const int m = (int)(r % m_cnt), tap = (int)(r / m_cnt);
    const float* src = ws + ((size_t)tap * 128 + m) * ldn + n_src0 + n;
    float acc = 0.0f;
    for (int s = 0; s < slabs; ++s) acc += src[(size_t)s * slab_stride];
    const int co = co0 + (transposed ? n : m), ci = ci0 + (transposed ? m : n);
    dw[((size_t)co * ci_total + ci) * taps + tap] = acc;
}

// ---- flatten / unflatten around the policy FC ----------------------------------------------------------------------
// feature k = ch*90 + r*9 + c of board b (nn.Flatten of [B, channels, 10, 9]) -> dense planes F[k/4][drow0 + b][4] and G
__global__ void __launch_bounds__(256) tn_flatten_kernel(const float* __restrict__ act, long long rows, int n_boards, int channels,
                                                         float* __restrict__ dense, float* __restrict__ dense_g, long long drows)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    const int k4_cnt = channels * 90 / 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)k4_cnt * n_boards) return;
    const int b = (int)(idx % n_boards), k4 = (int)(idx / n_boards);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = k4 * 4 + j, ch = k / 90, cell = k - ch * 90;
        const long long row = kTnRow0 + (long long)b * kTnBoard + (cell / 9) * 10 + cell % 9;
        v[j] = act[((size_t)(ch >> 2) * rows + row) * 4 + (ch & 3)];
    }
    const long long drow = kTnRow0 + b;
    st4(dense + ((size_t)k4 * drows + drow) * 4, make_float4(v[0], v[1], v[2], v[3]));
    if (dense_g) st4(dense_g + tn_g_off(k4 >> 1, drow, drows) + (k4 & 1) * 4, make_float4(v[0], v[1], v[2], v[3]));
}

// dense planes dF[k/4][drow0 + b][4] -> board planes chunks [0, channels/4) (real cells only)
__global__ void __launch_bounds__(256) tn_unflatten_kernel(const float* __restrict__ dense, long long drows, int n_boards, int channels,
                                                           float* __restrict__ planes, long long rows, int n_partials, long long part_stride)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    const int chunks = channels / 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n_boards * 90 * chunks) return;
    const int cell = (int)(idx % 90);
    const long long r = idx / 90;
    const int b = (int)(r % n_boards), q = (int)(r / n_boards);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = (q * 4 + j) * 90 + cell;
        const float* src = dense + ((size_t)(k >> 2) * drows + kTnRow0 + b) * 4 + (k & 3);
        float acc = src[0];
        for (int s = 1; s < n_partials; ++s) acc += src[(size_t)s * part_stride];
        v[j] = acc;
    }
    const long long row = kTnRow0 + (long long)b * kTnBoard + (cell / 9) * 10 + cell % 9;
    st4(planes + ((size_t)q * rows + row) * 4, make_float4(v[0], v[1], v[2], v[3]));
}

// row-major m[b][n] (n_cols multiple of 4) -> dense planes + G (one row per board)
__global__ void __launch_bounds__(256) tn_rows_layouts_kernel(const float* __restrict__ m, long long stride, int n_rows, int n_cols,
                                                              float* __restrict__ dense, float* __restrict__ dense_g, long long drows)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    const int n4_cnt = n_cols / 4;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n4_cnt * n_rows) return;
    const int n4 = (int)(idx % n4_cnt), b = (int)(idx / n4_cnt);
    const float4 v = ld4(m + (size_t)b * stride + (size_t)n4 * 4);
    const long long drow = kTnRow0 + b;
    if (dense) st4(dense + ((size_t)n4 * drows + drow) * 4, v);
    if (dense_g) st4(dense_g + tn_g_off(n4 >> 1, drow, drows) + (n4 & 1) * 4, v);
}

// out[n] = sum_b m[b][n] (bias gradient): 32 columns x 8 row groups per block, fixed order
__global__ void __launch_bounds__(256) tn_colsum_kernel(const float* __restrict__ m, long long stride, int n_rows, int n_cols, float* __restrict__ out)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    __shared__ float sm[8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    float acc = 0.0f;
    if (n < n_cols)
#pragma unroll 4
        for (int b = ty; b < n_rows; b += 8) acc += m[(size_t)b * stride + n];
    sm[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && n < n_cols) {
        float t = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sm[i][tx];
        out[n] = t;
    }
}

// ---- value head: Linear(4*90 -> 128) + ReLU + Linear(128 -> 1) + tanh (model.py:72-83) ---------------------------------
constexpr int kTnVF = 360, kTnVH = 128;

__device__ __forceinline__ float tn_value_feature(const float* act, long long rows, int chunk, int b, int k)
{
    const int ch = k / 90, cell = k - ch * 90;
    const long long row = kTnRow0 + (long long)b * kTnBoard + (cell / 9) * 10 + cell % 9;
    return act[((size_t)chunk * rows + row) * 4 + ch];
}

// one block (256 threads) per board: h[b][128] (after the ReLU), v[b]; a warp takes two hidden units per iteration
__global__ void __launch_bounds__(256) tn_value_fwd_kernel(const float* __restrict__ act, long long rows, int chunk, const float* __restrict__ w1,
                                                           const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                                           float* __restrict__ h, float* __restrict__ v)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    __shared__ float f[kTnVF], hs[kTnVH], red[4];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = threadIdx.x; k < kTnVF; k += 256) f[k] = tn_value_feature(act, rows, chunk, b, k);
    __syncthreads();
    for (int j = warp * 2; j < kTnVH; j += 16) {
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll 4
        for (int k = lane; k < kTnVF; k += 32) {
            a0 += w1[(size_t)j * kTnVF + k] * f[k];
            a1 += w1[(size_t)(j + 1) * kTnVF + k] * f[k];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        }
        if (lane == 0) {
            hs[j] = fmaxf(a0 + b1[j], 0.0f);
            hs[j + 1] = fmaxf(a1 + b1[j + 1], 0.0f);
        }
    }
    __syncthreads();
    if (threadIdx.x < kTnVH) {
        h[(size_t)b * kTnVH + threadIdx.x] = hs[threadIdx.x];
        float acc = hs[threadIdx.x] * w2[threadIdx.x];
#pragma unroll
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) red[warp] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) v[b] = tanhf(red[0] + red[1] + red[2] + red[3] + b2[0]);
}

// one block (384 threads) per board: dh[b][128], dpre[b], and the gradient of the 360 features into chunk `chunk` of dact
__global__ void __launch_bounds__(384) tn_value_bwd_a_kernel(const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ h,
                                                             const float* __restrict__ v, const float* __restrict__ g_value, float* __restrict__ dh,
                                                             float* __restrict__ dpre, float* __restrict__ dact, long long rows, int chunk)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    __shared__ float s_dh[kTnVH];
    const int b = blockIdx.x;
    const float vv = v[b];
    const float dp = g_value[b] * (1.0f - vv * vv);
    if (threadIdx.x < kTnVH) {
        const float hv = h[(size_t)b * kTnVH + threadIdx.x];
        const float d = hv > 0.0f ? dp * w2[threadIdx.x] : 0.0f;
        s_dh[threadIdx.x] = d;
        dh[(size_t)b * kTnVH + threadIdx.x] = d;
    }
    if (threadIdx.x == 0) dpre[b] = dp;
    __syncthreads();
    const int k = threadIdx.x;
    if (k < kTnVF) {
        float acc = 0.0f;
        for (int j = 0; j < kTnVH; ++j) acc += s_dh[j] * w1[(size_t)j * kTnVF + k];
        const int ch = k / 90, cell = k - ch * 90;
        const long long row = kTnRow0 + (long long)b * kTnBoard + (cell / 9) * 10 + cell % 9;
        dact[((size_t)chunk * rows + row) * 4 + ch] = acc;
    }
}

// blocks 0..127: dW1[j][:] and db1[j]; block 128: dW2 and db2.  Boards are added in order.
__global__ void __launch_bounds__(384) tn_value_bwd_w_kernel(const float* __restrict__ act, long long rows, int chunk, int n_boards,
                                                             const float* __restrict__ h, const float* __restrict__ dh, const float* __restrict__ dpre,
                                                             float* __restrict__ dw1, float* __restrict__ db1, float* __restrict__ dw2,
                                                             float* __restrict__ db2)
{
    griddep_launch_dependents();
    griddep_wait();                                  // everything this kernel reads or overwrites belongs to earlier kernels of the step
    const int j = blockIdx.x, k = threadIdx.x;
    if (j < kTnVH) {
        float acc = 0.0f, bacc = 0.0f;
        // feature k of board b sits at a fixed offset plus b * 110 rows: the address arithmetic leaves the loop, whose 16-fold
        // unrolling keeps that many loads in flight (the sum itself stays one chain in board order)
        const int kk = k < kTnVF ? k : 0, ch = kk / 90, cell = kk - ch * 90;
        const float* fk = act + ((size_t)chunk * rows + kTnRow0 + (cell / 9) * 10 + cell % 9) * 4 + ch;
        const float* dj = dh + j;
#pragma unroll 16
        for (int b = 0; b < n_boards; ++b) {
            const float d = dj[(size_t)b * kTnVH];
            bacc += d;
            acc += d * fk[(size_t)b * (kTnBoard * 4)];
        }
        if (k < kTnVF) dw1[(size_t)j * kTnVF + k] = acc;
        if (k == 0) db1[j] = bacc;
    } else if (k < kTnVH) {
        float acc = 0.0f, bacc = 0.0f;
        for (int b = 0; b < n_boards; ++b) {
            const float d = dpre[b];
            bacc += d;
            acc += d * h[(size_t)b * kTnVH + k];
        }
        dw2[k] = acc;
        if (k == 0) db2[0] = bacc;
    }
}

}  // namespace tn
}  // namespace xq
