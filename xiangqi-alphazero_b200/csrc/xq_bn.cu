// xq_bn.cu -- training-mode BatchNorm2d of the data-parallel step, statistics exchanged over NVLink peer memory.
//
// Replaces, inside AlphaZeroTrainer.train_network (training/train.py:397-423), the 15 BatchNorm2d layers of
// XiangqiNet (model.py:14-36, 49-83) in forward and backward when the 256-sample minibatch is split across the ranks
// of one NVSwitch box: the per-channel sums must be those of the WHOLE minibatch (the reference is one process).
// torch.nn.SyncBatchNorm does that with an NCCL collective per layer and direction plus a host synchronisation in
// every forward; a torch-op version without the synchronisation still issues ~50 tiny kernels per layer, and the
// 256-sample step is launch bound (profiles/r2_train_dp.md: 11.6 ms per step on 2 GPUs against 4.6 ms on one).
// Here a layer is TWO kernels per direction and no collective call at all:
//   reduce+push   one block per channel sums x and x^2 (dy and dy*xhat in backward) in float64 and STORES the two
//                 partial sums straight into every peer's exchange buffer (peer-mapped with CUDA IPC, plain stores
//                 over NVLink); the last block to finish publishes a sequence number in every peer's flag word;
//   wait+apply    every block spins (one thread, volatile loads) until all ranks' flags show this exchange, adds the
//                 partials in rank order -- every rank computes bit-identical statistics -- and normalises its share of
//                 the elements (forward: y, saved mean / invstd, running statistics; backward: dx).
// Exchange buffers are double buffered by the parity of the sequence number; a rank cannot run two exchanges ahead of a
// peer because its own wait+apply needs that peer's push.  With one rank the same kernels run on a local buffer.
#include "xq_ctx.h"

#include <cstring>

namespace xq {

constexpr int kBnMaxC = 512;
constexpr int kBnMaxWorld = 8;
constexpr int kBnMaxSplit = 16;              // blocks per channel
constexpr int kBnSlot = 2 * kBnMaxC + 8;     // doubles per (parity, source rank): sums, sums of products, count

struct BnXchg {
    double data[2][kBnMaxWorld][kBnSlot];
    unsigned long long flag[2][kBnMaxWorld];  // last sequence number (of that parity) whose data[parity][src] is complete
};

struct PeerState {
    int rank = 0, world = 1;
    BnXchg* mine = nullptr;                   // this rank's exchange buffer (cudaMalloc, exported with cudaIpcGetMemHandle)
    BnXchg* slots[kBnMaxWorld] = {nullptr};   // every rank's buffer as seen from here (slots[rank] == mine)
    bool opened[kBnMaxWorld] = {false};
    unsigned int* done = nullptr;             // [2 + kBnMaxC] finished-channel counters per parity, then finished-block counters per channel
    double* scratch = nullptr;                // [kBnMaxC][kBnMaxSplit][2] per-block partial sums of the reduce kernels
    unsigned long long seq = 0;               // exchanges issued so far (host side; every rank issues the same sequence)
};

struct PeerPtrs {
    BnXchg* slot[kBnMaxWorld];
};

__device__ __forceinline__ double block_sum(double v, double* sm)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sm[w] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sm[i];   // same order in every thread
    return t;
}

// BACKWARD = false: a = sum x, b = sum x^2.  BACKWARD = true: a = sum dy, b = sum dy * xhat (xhat from the saved statistics).
// grid (C, S): the samples of a channel are dealt to S blocks; the last of them to finish adds the S partial sums in
// block order (deterministic) and stores the channel's two sums into every rank's exchange buffer; the last channel to
// finish publishes the sequence number.
template <bool BACKWARD>
__global__ void __launch_bounds__(256)
bn_reduce_push_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ save_mean,
                      const float* __restrict__ save_invstd, float* __restrict__ dweight, float* __restrict__ dbias, int N, int C,
                      int HW, PeerPtrs peers, int rank, int world, unsigned long long seq, unsigned int* done, double* scratch)
{
    __shared__ double sm[8];
    __shared__ int last_flag;
    const int c = blockIdx.x, S = gridDim.y, sp = blockIdx.y;
    const int parity = (int)(seq & 1ull);
    double a = 0.0, b = 0.0;
    const float mean = BACKWARD ? save_mean[c] : 0.0f, inv = BACKWARD ? save_invstd[c] : 0.0f;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = sp * 8 + warp; n < N; n += S * 8) {               // a warp takes a sample's HW contiguous values
        const size_t base = ((size_t)n * C + c) * HW;
        for (int i = lane; i < HW; i += 32) {
            if (BACKWARD) {
                const float g = dy[base + i];
                a += (double)g;
                b += (double)g * (double)((x[base + i] - mean) * inv);
            } else {
                const double v = (double)x[base + i];
                a += v;
                b += v * v;
            }
        }
    }
    a = block_sum(a, sm);
    b = block_sum(b, sm);
    unsigned int* done_c = done + 2 + c;                           // [2] per-parity channel counters, then one counter per channel
    if (threadIdx.x == 0) {
        scratch[((size_t)c * S + sp) * 2 + 0] = a;
        scratch[((size_t)c * S + sp) * 2 + 1] = b;
        __threadfence();
        last_flag = atomicAdd(done_c, 1u) == (unsigned)(S - 1);
    }
    __syncthreads();
    if (!last_flag) return;
    if (threadIdx.x == 0) {
        __threadfence();
        *done_c = 0;
        double ta = 0.0, tb = 0.0;
        for (int k = 0; k < S; ++k) {
            ta += *(volatile double*)&scratch[((size_t)c * S + k) * 2 + 0];
            tb += *(volatile double*)&scratch[((size_t)c * S + k) * 2 + 1];
        }
        if (BACKWARD) {
            dbias[c] = (float)ta;         // the parameter gradients stay local: the gradient all-reduce sums them
            dweight[c] = (float)tb;
        }
        for (int p = 0; p < world; ++p) {
            volatile double* d = peers.slot[p]->data[parity][rank];
            d[c] = ta;
            d[C + c] = tb;
            if (c == 0) d[2 * C] = (double)N * (double)HW;
        }
        __threadfence_system();
        const unsigned int prev = atomicAdd(&done[parity], 1u);
        if (prev == gridDim.x - 1) {
            // every channel's sums are out (each fenced before its increment): publish the exchange everywhere
            done[parity] = 0;
            __threadfence_system();
            for (int p = 0; p < world; ++p) *(volatile unsigned long long*)&peers.slot[p]->flag[parity][rank] = seq;
        }
    }
}

__device__ __forceinline__ void bn_wait_all(const BnXchg* mine, int parity, int world, unsigned long long seq)
{
    if (threadIdx.x == 0) {
        for (int s = 0; s < world; ++s)
            while (*(volatile const unsigned long long*)&mine->flag[parity][s] < seq) __nanosleep(64);
        __threadfence_system();
    }
    __syncthreads();
}

// forward apply: y = (x - mean) * invstd * w + b over the block's share of channel c; block (c, 0) also writes the saved
// statistics and the running update of F.batch_norm (momentum, unbiased variance)
__global__ void __launch_bounds__(256)
bn_fwd_apply_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ weight, const float* __restrict__ bias,
                    float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ save_mean,
                    float* __restrict__ save_invstd, int N, int C, int HW, float eps, float momentum, const BnXchg* mine, int world,
                    unsigned long long seq)
{
    const int c = blockIdx.x;
    const int parity = (int)(seq & 1ull);
    bn_wait_all(mine, parity, world, seq);
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    for (int s = 0; s < world; ++s) {                          // rank order: identical result on every rank
        const volatile double* d = mine->data[parity][s];
        s1 += d[c];
        s2 += d[C + c];
        cnt += d[2 * C];
    }
    const double mean = s1 / cnt;
    double var = s2 / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float mf = (float)mean, w = weight[c], bb = bias[c];
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        save_mean[c] = mf;
        save_invstd[c] = invstd;
        const double unbiased = cnt > 1.0 ? var * (cnt / (cnt - 1.0)) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mf;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = blockIdx.y * 8 + warp; n < N; n += gridDim.y * 8) {
        const size_t base = ((size_t)n * C + c) * HW;
        for (int i = lane; i < HW; i += 32) y[base + i] = (x[base + i] - mf) * invstd * w + bb;
    }
}

// backward apply: dx = w * invstd * (dy - mean(dy) - xhat * mean(dy * xhat)), means over the whole minibatch
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ weight,
                    const float* __restrict__ save_mean, const float* __restrict__ save_invstd, float* __restrict__ dx, int N, int C,
                    int HW, const BnXchg* mine, int world, unsigned long long seq)
{
    const int c = blockIdx.x;
    const int parity = (int)(seq & 1ull);
    bn_wait_all(mine, parity, world, seq);
    double s1 = 0.0, s2 = 0.0, cnt = 0.0;
    for (int s = 0; s < world; ++s) {
        const volatile double* d = mine->data[parity][s];
        s1 += d[c];
        s2 += d[C + c];
        cnt += d[2 * C];
    }
    const float mean_dy = (float)(s1 / cnt), mean_dyx = (float)(s2 / cnt);
    const float mf = save_mean[c], inv = save_invstd[c], k = weight[c] * inv;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = blockIdx.y * 8 + warp; n < N; n += gridDim.y * 8) {
        const size_t base = ((size_t)n * C + c) * HW;
        for (int i = lane; i < HW; i += 32) {
            const float xhat = (x[base + i] - mf) * inv;
            dx[base + i] = k * (dy[base + i] - mean_dy - xhat * mean_dyx);
        }
    }
}

}  // namespace xq

using namespace xq;

static PeerState* PS_(xq_ctx* c) { return reinterpret_cast<PeerState*>(c->peer); }

extern "C" void xq_peer_free_(xq_ctx* c)
{
    PeerState* P = c ? PS_(c) : nullptr;
    if (!P) return;
    for (int r = 0; r < P->world; ++r)
        if (r != P->rank && P->opened[r] && P->slots[r]) cudaIpcCloseMemHandle(P->slots[r]);
    if (P->mine) cudaFree(P->mine);
    if (P->done) cudaFree(P->done);
    if (P->scratch) cudaFree(P->scratch);
    delete P;
    c->peer = nullptr;
}

// Allocates this rank's exchange buffer and returns its CUDA IPC handle (64 bytes) for the other ranks of the box.
extern "C" int xq_peer_create(xq_ctx* c, int rank, int world, unsigned char* handle_out64)
{
    if (!c || rank < 0 || world < 1 || world > kBnMaxWorld || rank >= world)
        return xq_fail(c, XQ_ERR_ARG, "xq_peer_create: rank %d of %d (at most %d ranks: one NVSwitch box)", rank, world, kBnMaxWorld);
    XQ_CUDA(c, cudaSetDevice(c->device));
    xq_peer_free_(c);
    PeerState* P = new PeerState();
    c->peer = P;
    P->rank = rank;
    P->world = world;
    XQ_CUDA(c, cudaMalloc(&P->mine, sizeof(BnXchg)));
    XQ_CUDA(c, cudaMemset(P->mine, 0, sizeof(BnXchg)));
    XQ_CUDA(c, cudaMalloc(&P->done, (2 + kBnMaxC) * sizeof(unsigned int)));
    XQ_CUDA(c, cudaMemset(P->done, 0, (2 + kBnMaxC) * sizeof(unsigned int)));
    XQ_CUDA(c, cudaMalloc(&P->scratch, (size_t)kBnMaxC * kBnMaxSplit * 2 * sizeof(double)));
    XQ_CUDA(c, cudaDeviceSynchronize());
    P->slots[rank] = P->mine;
    if (handle_out64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        cudaIpcMemHandle_t h;
        if (world > 1) XQ_CUDA(c, cudaIpcGetMemHandle(&h, P->mine));
        else memset(&h, 0, sizeof(h));
        memcpy(handle_out64, &h, 64);
    }
    return XQ_OK;
}

// handles = the world x 64 bytes gathered from every rank's xq_peer_create, in rank order
extern "C" int xq_peer_connect(xq_ctx* c, const unsigned char* handles)
{
    PeerState* P = c ? PS_(c) : nullptr;
    if (!P || !handles) return xq_fail(c, XQ_ERR_STATE, "xq_peer_connect: call xq_peer_create first");
    XQ_CUDA(c, cudaSetDevice(c->device));
    for (int r = 0; r < P->world; ++r) {
        if (r == P->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* ptr = nullptr;
        XQ_CUDA(c, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        P->slots[r] = reinterpret_cast<BnXchg*>(ptr);
        P->opened[r] = true;
    }
    return XQ_OK;
}

static int peer_ready(xq_ctx* c, PeerState** out, PeerPtrs* pp)
{
    if (!PS_(c)) {
        int rc = xq_peer_create(c, 0, 1, nullptr);       // single rank: the same kernels on a local buffer
        if (rc) return rc;
    }
    PeerState* P = PS_(c);
    for (int r = 0; r < kBnMaxWorld; ++r) pp->slot[r] = r < P->world ? P->slots[r] : nullptr;
    for (int r = 0; r < P->world; ++r)
        if (!pp->slot[r]) return xq_fail(c, XQ_ERR_STATE, "xq_bn_*: peer %d is not connected (xq_peer_connect)", r);
    *out = P;
    return XQ_OK;
}

// blocks per channel: a block's 8 warps take 8 samples per pass; about 32 samples (2 880 values) per block
static inline int bn_splits(int N, int HW)
{
    (void)HW;
    int s = (N + 31) / 32;
    return s < 1 ? 1 : (s > kBnMaxSplit ? kBnMaxSplit : s);
}

extern "C" int xq_bn_forward(xq_ctx* c, const float* d_x, float* d_y, const float* d_weight, const float* d_bias,
                             float* d_running_mean, float* d_running_var, float* d_save_mean, float* d_save_invstd, int N, int C,
                             int HW, float eps, float momentum, void* stream)
{
    if (!c || !d_x || !d_y || !d_weight || !d_bias || !d_running_mean || !d_running_var || !d_save_mean || !d_save_invstd ||
        N <= 0 || C <= 0 || C > kBnMaxC || HW <= 0)
        return xq_fail(c, XQ_ERR_ARG, "xq_bn_forward: bad arguments (N=%d C=%d HW=%d)", N, C, HW);
    PeerState* P = nullptr;
    PeerPtrs pp;
    if (int rc = peer_ready(c, &P, &pp)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned long long seq = ++P->seq;
    bn_reduce_push_kernel<false><<<dim3(C, bn_splits(N, HW)), 256, 0, s>>>(d_x, nullptr, nullptr, nullptr, nullptr, nullptr, N, C, HW, pp,
                                                                           P->rank, P->world, seq, P->done, P->scratch);
    bn_fwd_apply_kernel<<<dim3(C, bn_splits(N, HW)), 256, 0, s>>>(d_x, d_y, d_weight, d_bias, d_running_mean, d_running_var, d_save_mean,
                                                                  d_save_invstd, N, C, HW, eps, momentum, P->mine, P->world, seq);
    c->launches += 2;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_bn_backward(xq_ctx* c, const float* d_x, const float* d_dy, const float* d_weight, const float* d_save_mean,
                              const float* d_save_invstd, float* d_dx, float* d_dweight, float* d_dbias, int N, int C, int HW,
                              void* stream)
{
    if (!c || !d_x || !d_dy || !d_weight || !d_save_mean || !d_save_invstd || !d_dx || !d_dweight || !d_dbias || N <= 0 || C <= 0 ||
        C > kBnMaxC || HW <= 0)
        return xq_fail(c, XQ_ERR_ARG, "xq_bn_backward: bad arguments (N=%d C=%d HW=%d)", N, C, HW);
    PeerState* P = nullptr;
    PeerPtrs pp;
    if (int rc = peer_ready(c, &P, &pp)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned long long seq = ++P->seq;
    bn_reduce_push_kernel<true><<<dim3(C, bn_splits(N, HW)), 256, 0, s>>>(d_x, d_dy, d_save_mean, d_save_invstd, d_dweight, d_dbias, N, C, HW,
                                                                          pp, P->rank, P->world, seq, P->done, P->scratch);
    bn_bwd_apply_kernel<<<dim3(C, bn_splits(N, HW)), 256, 0, s>>>(d_x, d_dy, d_weight, d_save_mean, d_save_invstd, d_dx, N, C, HW, P->mine,
                                                                  P->world, seq);
    c->launches += 2;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}
