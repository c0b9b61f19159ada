// xq_train.cu -- the training-step side of the self-play pipeline.
//
// Replaces, around the torch forward/backward of XiangqiNet that train.py keeps:
//   - the host replay buffer of dense tuples (train.py:203 deque + SelfPlayDataset :114-129, 64.8 KB per
//     sample through a Python DataLoader) by a device-resident ring of the 896-byte sparse self-play records
//     (replay_append_kernel: z-labelling of parallel_selfplay.py:124-132 fused into the copy);
//   - _augment_data / augment_data (parallel_selfplay.py:137-151, train.py:132-151) + get_state_for_nn
//     (game.py:618-640) by train_batch_kernel: a minibatch is built from ring records by index, the mirrored
//     twin of a record is an index permutation (column flip of the board, (fr,fc,tr,tc) -> (fr,8-fc,tr,8-tc))
//     applied while the planes are written;
//   - the loss of train.py:408-414 (soft-target cross entropy + MSE) by pv_loss_kernel: one pass over the
//     [B, 8100] logits gives both losses AND d loss / d logits (softmax * sum(pi) - pi) / B, the sparse
//     targets are never densified;
//   - clip_grad_norm_(1.0) + Adam(weight_decay) (train.py:190-194, 418-419) by sumsq_* + adam_kernel over ONE flat
//     parameter / gradient buffer (the same buffer the NCCL all-reduce of the data-parallel step uses).
// All four are HBM-bound streaming kernels (no tensor cores): coalesced 16-byte accesses, grids sized from
// the SM count.
#include "xq_ctx.h"
#include "xq_rules.cuh"

namespace xq {

constexpr int kRecBytes = XQ_SAMPLE_BYTES;      // 896 = 56 x 16

// ---- replay ring ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
replay_append_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ idx, int n,
                     const int8_t* __restrict__ winner, int n_results, uint8_t* __restrict__ ring,
                     float* __restrict__ ring_z, long long cap, long long head)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n) return;
    const uint4* s = reinterpret_cast<const uint4*>(src + (size_t)idx[w] * kRecBytes);
    const long long slot = (head + w) % cap;
    uint4* d = reinterpret_cast<uint4*>(ring + (size_t)slot * kRecBytes);
    d[lane] = s[lane];
    if (lane < 56 - 32) d[32 + lane] = s[32 + lane];
    if (lane == 0) {
        const uint8_t* r = reinterpret_cast<const uint8_t*>(s);
        const int side = (int8_t)r[90];
        const int uid = *reinterpret_cast<const int32_t*>(r + 92);
        const int wv = (uid >= 0 && uid < n_results) ? winner[uid] : 0;
        // parallel_selfplay.py:124-132: 0 for a draw, +1 if the side to move at the sample won, else -1
        ring_z[slot] = (wv == 0 || wv == 2) ? 0.0f : (wv == side ? 1.0f : -1.0f);
    }
}

// ---- minibatch builder -------------------------------------------------------------------------------
constexpr int kTbWarps = 8;
constexpr int kPlaneWordsT = 44;

struct __align__(16) TrainBatchSmem {
    int8_t board[kTbWarps][96];
    uint32_t bits[kTbWarps][kPlaneWordsT];
    float4 nib_lut[16];
};

// logical index L: record (L >> 1) counted from the oldest record of the ring, mirrored twin if L & 1
// (the reference appends (sample, mirrored sample) pairs, parallel_selfplay.py:141-150).
__global__ void __launch_bounds__(kTbWarps * 32)
train_batch_kernel(const uint8_t* __restrict__ ring, const float* __restrict__ ring_z, long long cap, long long start,
                   const int64_t* __restrict__ logical, int B, float* __restrict__ planes, int16_t* __restrict__ act,
                   float* __restrict__ prob, int32_t* __restrict__ n_out, float* __restrict__ z_out)
{
    __shared__ TrainBatchSmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 16)
        sm.nib_lut[threadIdx.x] = make_float4((threadIdx.x & 1) ? 1.0f : 0.0f, (threadIdx.x & 2) ? 1.0f : 0.0f,
                                              (threadIdx.x & 4) ? 1.0f : 0.0f, (threadIdx.x & 8) ? 1.0f : 0.0f);
    __syncthreads();
    const int i = blockIdx.x * kTbWarps + warp;
    if (i >= B) return;
    const long long L = logical[i];
    const bool mirror = (L & 1) != 0;
    const long long slot = (start + (L >> 1)) % cap;
    const uint8_t* rec = ring + (size_t)slot * kRecBytes;
    int8_t* b = sm.board[warp];
    uint32_t* bits = sm.bits[warp];
    // board, column-flipped for the mirrored twin (np.flip(state, axis=2))
    for (int sq = lane; sq < 90; sq += 32) {
        const int r = sq / 9, c = sq - r * 9;
        b[sq] = (int8_t)rec[mirror ? r * 9 + 8 - c : sq];
    }
    bits[lane] = 0u;
    if (lane < kPlaneWordsT - 32) bits[32 + lane] = 0u;
    const int side = (int8_t)rec[90];
    const int n = rec[91];
    __syncwarp();
    // get_state_for_nn (game.py:618-640) as bits, then 4 bits -> one float4 (same scheme as movegen_kernel)
    for (int sq = lane; sq < 90; sq += 32) {
        const int v = b[sq] * side;
        if (v != 0) {
            const int e = (v > 0 ? v - 1 : 6 - v) * 90 + sq;
            atomicOr(&bits[e >> 5], 1u << (e & 31));
        }
    }
    if (side == 1 && lane < 4) atomicOr(&bits[39 + lane], lane == 0 ? 0xfffff000u : (lane == 3 ? 0x3fu : 0xffffffffu));
    __syncwarp();
    float* outp = planes + (size_t)i * 1350;
    const int head = (i & 1) * 2;
    if (lane == 0) {
        const uint32_t two = head ? bits[0] : bits[42] >> 4;
        float2 v;
        v.x = (two & 1u) ? 1.0f : 0.0f;
        v.y = (two & 2u) ? 1.0f : 0.0f;
        *reinterpret_cast<float2*>(head ? outp : outp + 1348) = v;
    }
    float4* out4 = reinterpret_cast<float4*>(outp + head);
    const int e0 = head + 4 * lane;
    const int sh = e0 & 31;
    const uint32_t* wp = bits + (e0 >> 5);
#pragma unroll
    for (int it = 0; it < 11; ++it) {
        const int k = it * 32 + lane;
        if (it < 10 || k < 337) out4[k] = sm.nib_lut[__funnelshift_r(wp[4 * it], wp[4 * it + 1], sh) & 15u];
    }
    // sparse policy target: action ids (mirrored: fc -> 8-fc, tc -> 8-tc) and visit probabilities
    const int16_t* ra = reinterpret_cast<const int16_t*>(rec + 128);
    const float* rp = reinterpret_cast<const float*>(rec + 384);
    for (int k = lane; k < XQ_MAX_MOVES; k += 32) {
        int a = ra[k];
        if (k < n && mirror) {
            const int f = a / 90, t = a - f * 90;
            const int fm = f + 8 - 2 * (f % 9), tm = t + 8 - 2 * (t % 9);
            a = fm * 90 + tm;
        }
        act[(size_t)i * XQ_MAX_MOVES + k] = k < n ? (int16_t)a : (int16_t)-1;
        prob[(size_t)i * XQ_MAX_MOVES + k] = k < n ? rp[k] : 0.0f;
    }
    if (lane == 0) {
        n_out[i] = n;
        z_out[i] = ring_z[slot];
    }
}

// ---- fused policy / value loss and its gradient ------------------------------------------------------
constexpr int kLossThreads = 256;
constexpr int kRow4 = XQ_ACTION_SPACE / 4;      // 2025 float4 per row
constexpr int kPerThread = 8;                   // 256 x 8 = 2048 >= 2025

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, t) : v + t;
    }
    __syncthreads();            // red may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < kLossThreads / 32; ++w) r = is_max ? fmaxf(r, red[w]) : r + red[w];
    return r;
}

// train.py:408-414 for one sample per CTA:
//   policy_loss_row = -sum_a pi_a log_softmax(logits)_a = sum(pi) * logsumexp(logits) - sum_a pi_a logits_a
//   value_loss_row  = (v - z)^2
//   d/dlogits = (softmax * sum(pi) - pi) * inv_batch,  d/dv = 2 (v - z) * inv_batch
// The row is read from HBM once (8 float4 per thread stay in registers between the reductions and the store).
__global__ void __launch_bounds__(kLossThreads)
pv_loss_kernel(const float* __restrict__ logits, long long stride, const float* __restrict__ value,
               const int16_t* __restrict__ act, const float* __restrict__ prob, const int32_t* __restrict__ n_moves,
               const float* __restrict__ z, float inv_batch, float* __restrict__ grad_logits, long long gstride,
               float* __restrict__ grad_value, float* __restrict__ ploss_rows, float* __restrict__ vloss_rows)
{
    __shared__ float red[kLossThreads / 32];
    const int i = blockIdx.x;
    const float4* row = reinterpret_cast<const float4*>(logits + (size_t)i * stride);
    float4 x[kPerThread];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
        const int j = threadIdx.x + k * kLossThreads;
        if (j < kRow4) {
            x[k] = __ldcs(row + j);
            mx = fmaxf(mx, fmaxf(fmaxf(x[k].x, x[k].y), fmaxf(x[k].z, x[k].w)));
        }
    }
    mx = block_reduce(mx, true, red);
    float se = 0.0f;
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
        const int j = threadIdx.x + k * kLossThreads;
        if (j < kRow4) se += expf(x[k].x - mx) + expf(x[k].y - mx) + expf(x[k].z - mx) + expf(x[k].w - mx);
    }
    se = block_reduce(se, false, red);
    const float lse = mx + logf(se);
    // sparse target: sum(pi) and pi . logits
    const int n = n_moves[i];
    float sp = 0.0f, dot = 0.0f;
    if ((int)threadIdx.x < n) {
        const float p = prob[(size_t)i * XQ_MAX_MOVES + threadIdx.x];
        const int a = act[(size_t)i * XQ_MAX_MOVES + threadIdx.x];
        sp = p;
        dot = p * logits[(size_t)i * stride + a];
    }
    sp = block_reduce(sp, false, red);
    dot = block_reduce(dot, false, red);
    float4* grow = reinterpret_cast<float4*>(grad_logits + (size_t)i * gstride);
    const float scale = sp * inv_batch;
#pragma unroll
    for (int k = 0; k < kPerThread; ++k) {
        const int j = threadIdx.x + k * kLossThreads;
        if (j < kRow4) {
            float4 g;
            g.x = expf(x[k].x - lse) * scale;
            g.y = expf(x[k].y - lse) * scale;
            g.z = expf(x[k].z - lse) * scale;
            g.w = expf(x[k].w - lse) * scale;
            grow[j] = g;
        }
    }
    __syncthreads();            // the dense row is written (block-visible) before the sparse correction
    if ((int)threadIdx.x < n) {
        const float p = prob[(size_t)i * XQ_MAX_MOVES + threadIdx.x];
        const int a = act[(size_t)i * XQ_MAX_MOVES + threadIdx.x];
        grad_logits[(size_t)i * gstride + a] -= p * inv_batch;     // legal actions of one position are distinct
    }
    if (threadIdx.x == 0) {
        const float d = value[i] - z[i];
        ploss_rows[i] = sp * lse - dot;
        vloss_rows[i] = d * d;
        grad_value[i] = 2.0f * d * inv_batch;
    }
}

// ---- gradient norm + fused clip / weight decay / Adam over the flat buffers -------------------------------
constexpr int kOptThreads = 256;

__global__ void __launch_bounds__(kOptThreads)
sumsq_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial)
{
    __shared__ float red[kOptThreads / 32];
    float s = 0.0f;
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long j = (long long)blockIdx.x * kOptThreads + threadIdx.x; j < n4; j += (long long)gridDim.x * kOptThreads) {
        const float4 v = g4[j];
        s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const float v = g[(n4 << 2) + threadIdx.x];
        s += v * v;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < kOptThreads / 32; ++w) t += red[w];
        partial[blockIdx.x] = t;
    }
}

// fixed summation order -> the norm (and with it the whole step) is reproducible run to run
__global__ void __launch_bounds__(kOptThreads)
sumsq_final_kernel(const float* __restrict__ partial, int m, float* __restrict__ out)
{
    __shared__ float red[kOptThreads];
    float s = 0.0f;
    for (int j = threadIdx.x; j < m; j += kOptThreads) s += partial[j];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = kOptThreads / 2; o; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

// torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.Adam.step (weight_decay = L2 added to the
// gradient, no amsgrad):  coef = min(1, max_norm / (||g|| + 1e-6));  g' = coef g + wd p;  m += (1-b1)(g' - m);
// v = b2 v + (1-b2) g'^2;  p -= (lr / (1-b1^t)) m / (sqrt(v) / sqrt(1-b2^t) + eps).   28 bytes of HBM traffic per parameter.
__global__ void __launch_bounds__(kOptThreads)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
            float step_size, float b1, float b2, float eps, float wd, float bc2_sqrt, const float* __restrict__ gradnorm2,
            float max_norm, float grad_scale)
{
    float coef = grad_scale;
    if (max_norm > 0.0f && gradnorm2) coef *= fminf(1.0f, max_norm / (sqrtf(*gradnorm2) * grad_scale + 1e-6f));
    const long long n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        const float gr = gg * coef + wd * pp;
        mm = mm + (1.0f - b1) * (gr - mm);
        vv = b2 * vv + (1.0f - b2) * gr * gr;
        pp -= step_size * mm / (sqrtf(vv) / bc2_sqrt + eps);
    };
    for (long long j = (long long)blockIdx.x * kOptThreads + threadIdx.x; j < n4; j += (long long)gridDim.x * kOptThreads) {
        float4 pp = p4[j], mm = m4[j], vv = v4[j];
        const float4 gg = g4[j];
        upd(pp.x, gg.x, mm.x, vv.x);
        upd(pp.y, gg.y, mm.y, vv.y);
        upd(pp.z, gg.z, mm.z, vv.z);
        upd(pp.w, gg.w, mm.w, vv.w);
        p4[j] = pp;
        m4[j] = mm;
        v4[j] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long j = (n4 << 2) + threadIdx.x;
        upd(p[j], g[j], m[j], v[j]);
    }
}

}  // namespace xq

using namespace xq;

extern "C" int xq_replay_append(xq_ctx* c, const void* d_src_records, const int64_t* d_src_index, int n,
                                const int8_t* d_winner, int n_results, void* d_ring, float* d_ring_z, long long capacity,
                                long long head, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_replay_append: ctx is NULL");
    if (n < 0 || capacity <= 0 || head < 0 || n > capacity ||
        (n > 0 && (!d_src_records || !d_src_index || !d_winner || !d_ring || !d_ring_z)))
        return xq_fail(c, XQ_ERR_ARG, "xq_replay_append: bad arguments (n=%d, capacity=%lld)", n, capacity);
    if (n == 0) return XQ_OK;
    cudaStream_t s = (cudaStream_t)stream;
    {
        XqTimer tm(c, s);
        replay_append_kernel<<<(n + 7) / 8, 256, 0, s>>>((const uint8_t*)d_src_records, d_src_index, n, d_winner, n_results,
                                                         (uint8_t*)d_ring, d_ring_z, capacity, head % capacity);
    }
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_train_batch(xq_ctx* c, const void* d_ring, const float* d_ring_z, long long capacity, long long start,
                              const int64_t* d_logical_index, int B, float* d_planes, int16_t* d_actions, float* d_probs,
                              int32_t* d_n_moves, float* d_z, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_train_batch: ctx is NULL");
    if (B < 0 || capacity <= 0 || start < 0 ||
        (B > 0 && (!d_ring || !d_ring_z || !d_logical_index || !d_planes || !d_actions || !d_probs || !d_n_moves || !d_z)))
        return xq_fail(c, XQ_ERR_ARG, "xq_train_batch: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    if ((uintptr_t)d_planes & 15) return xq_fail(c, XQ_ERR_ARG, "xq_train_batch: planes must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    {
        XqTimer tm(c, s);
        train_batch_kernel<<<(B + kTbWarps - 1) / kTbWarps, kTbWarps * 32, 0, s>>>(
            (const uint8_t*)d_ring, d_ring_z, capacity, start % capacity, d_logical_index, B, d_planes, d_actions, d_probs,
            d_n_moves, d_z);
    }
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_policy_value_loss(xq_ctx* c, const float* d_logits, long long logit_stride, const float* d_value,
                                    const int16_t* d_actions, const float* d_probs, const int32_t* d_n_moves, const float* d_z,
                                    int B, float inv_batch, float* d_grad_logits, long long grad_stride, float* d_grad_value,
                                    float* d_policy_loss_rows, float* d_value_loss_rows, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_policy_value_loss: ctx is NULL");
    if (B < 0 || (B > 0 && (!d_logits || !d_value || !d_actions || !d_probs || !d_n_moves || !d_z || !d_grad_logits ||
                            !d_grad_value || !d_policy_loss_rows || !d_value_loss_rows)))
        return xq_fail(c, XQ_ERR_ARG, "xq_policy_value_loss: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    if (logit_stride < XQ_ACTION_SPACE || grad_stride < XQ_ACTION_SPACE || (logit_stride & 3) || (grad_stride & 3) ||
        ((uintptr_t)d_logits & 15) || ((uintptr_t)d_grad_logits & 15))
        return xq_fail(c, XQ_ERR_ARG, "xq_policy_value_loss: rows must be 16-byte aligned, strides multiples of 4 and >= 8100");
    cudaStream_t s = (cudaStream_t)stream;
    {
        XqTimer tm(c, s);
        pv_loss_kernel<<<B, kLossThreads, 0, s>>>(d_logits, logit_stride, d_value, d_actions, d_probs, d_n_moves, d_z, inv_batch,
                                                  d_grad_logits, grad_stride, d_grad_value, d_policy_loss_rows, d_value_loss_rows);
    }
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_grad_sumsq(xq_ctx* c, const float* d_grad, long long n, float* d_partial, int n_partial, float* d_out,
                             void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_grad_sumsq: ctx is NULL");
    if (n < 0 || n_partial <= 0 || !d_grad || !d_partial || !d_out || ((uintptr_t)d_grad & 15))
        return xq_fail(c, XQ_ERR_ARG, "xq_grad_sumsq: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    int grid = c->sm_count * 4;
    if (grid > n_partial) grid = n_partial;
    {
        XqTimer tm(c, s);
        sumsq_partial_kernel<<<grid, kOptThreads, 0, s>>>(d_grad, n, d_partial);
    }
    sumsq_final_kernel<<<1, kOptThreads, 0, s>>>(d_partial, grid, d_out);
    c->launches += 2;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_adam_step(xq_ctx* c, float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, long long n,
                            float lr, float beta1, float beta2, float eps, float weight_decay, long long step,
                            const float* d_grad_sumsq, float max_norm, float grad_scale, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_adam_step: ctx is NULL");
    if (n < 0 || step < 1 || !d_param || !d_grad || !d_exp_avg || !d_exp_avg_sq ||
        (((uintptr_t)d_param | (uintptr_t)d_grad | (uintptr_t)d_exp_avg | (uintptr_t)d_exp_avg_sq) & 15))
        return xq_fail(c, XQ_ERR_ARG, "xq_adam_step: bad arguments (buffers must be 16-byte aligned, step >= 1)");
    if (n == 0) return XQ_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    const long long n4 = (n + 3) >> 2;
    long long want = (n4 + kOptThreads - 1) / kOptThreads;
    int grid = c->sm_count * 8;
    if (want < grid) grid = (int)want;
    {
        XqTimer tm(c, s);
        adam_kernel<<<grid, kOptThreads, 0, s>>>(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, (float)((double)lr / bc1), beta1, beta2,
                                                 eps, weight_decay, (float)sqrt(bc2), d_grad_sumsq, max_norm, grad_scale);
    }
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}
