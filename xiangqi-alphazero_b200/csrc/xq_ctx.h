// xq_ctx.h -- internal context shared by the translation units of libxq_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include <string>

#include "../../include/xq_b200.h"

struct xq_ctx {
    int device = 0;
    int sm_count = 148;
    char err[512] = {0};
    long long launches = 0;
    bool timing = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool ev_valid = false;
    int* d_overflow = nullptr;            // movegen overflow counter
    // host-pipeline resources (lazy)
    cudaStream_t pipe[2] = {nullptr, nullptr};
    void* d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    void* d_scratch = nullptr;            // grow-only device scratch (xq_has_legal_moves_batch)
    size_t scratch_bytes = 0;
    // subsystem state owned by other translation units
    void* mcts = nullptr;
    void* net = nullptr;
    void* selfplay = nullptr;
    void* tnet = nullptr;                 // xq_tnet.cu: the hand-written training step
    void* peer = nullptr;                 // xq_bn.cu: NVLink peer exchange buffers of the data-parallel BatchNorm
    int movegen_impl = 1;                 // XQ_MOVEGEN_IMPL=thread|warp: K1 kernel generation (1 = one thread per board, the default: 2x the
                                          // positions/s of the one-warp-per-board kernel, same bytes out)
    bool net_fork = false;                // XQ_NET_FORK=1: value MLP on a side stream next to the policy FC (measured: 1.002 ms per forward against
                                          // 0.976 ms in sequence -- the co-running CTAs slow the FC more than the 19 us they hide; off)
    bool net_2cta = true;                 // XQ_NET_2CTA=0 (read once in xq_create): tower convs on the single-CTA kernel instead of CTA pairs
    bool sp_graph = true;                 // XQ_SP_GRAPH=0: the lockstep step of the self-play / arena loops is issued launch by launch (no CUDA graph)
    bool net_small = true;                // XQ_NET_SMALL=0: no 64-channel single-CTA items for tower layers whose work fits one wave (small batches)
    bool net_pdl = true;                  // XQ_NET_PDL=0 (read once in xq_create): no programmatic dependent launch between the tower's layers
    bool train_pdl = false;               // XQ_TRAIN_PDL=1: the same for the kernels of the training step (xq_tnet.cu).  Measured inside the
                                          // step's CUDA graph: 1.94 ms per step with it against 1.86 ms without -- those kernels fill the
                                          // GPU, a dependent CTA finds no free SM to set itself up on, and graph edges are cheap already
    bool tpb_attr_set = false;            // dynamic shared-memory limit of movegen_tpb_kernel raised on this context's device
};

extern char g_xq_last_error[512];

inline int xq_fail(xq_ctx* ctx, int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_xq_last_error, sizeof(g_xq_last_error), fmt, ap);
    va_end(ap);
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", g_xq_last_error);
    return code;
}

#define XQ_CUDA(ctx, expr)                                                                         \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return xq_fail(ctx, XQ_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,         \
                           cudaGetErrorString(e_));                                                \
    } while (0)

// brackets the main kernel of a call with events when timing is on
struct XqTimer {
    xq_ctx* c;
    cudaStream_t s;
    XqTimer(xq_ctx* ctx, cudaStream_t st) : c(ctx), s(st)
    {
        if (c->timing) cudaEventRecord(c->ev0, s);
    }
    ~XqTimer()
    {
        if (c->timing) {
            cudaEventRecord(c->ev1, s);
            c->ev_valid = true;
        }
    }
};

// ---- PTX helpers: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -----------------
namespace xq {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16 B aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Programmatic dependent launch (kernels launched with cudaLaunchAttributeProgrammaticStreamSerialization): a kernel's CTAs
// may start while the previous kernel of the stream is still running -- on SMs it does not occupy, or as its CTAs retire --
// and do everything that does not depend on it (barriers, TMEM, weight stages) before griddep_wait() returns, which is when
// the previous grid has completed and its stores are visible.  SASS: PREEXIT / ACQBULK.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
}  // namespace xq

// launch with the programmatic-stream-serialization attribute (see griddep_wait); pdl = false: a plain launch
template <class... KArgs, class... Args>
static inline cudaError_t xq_launch_pdl(bool pdl, void (*kern)(KArgs...), dim3 grid, int block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
