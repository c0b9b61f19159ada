// xq_net.cu -- K3: bf16 tcgen05/TMA implicit-GEMM kernels for the policy-value ResNet forward.
//
// Replaces the library calls of XiangqiNet.forward (training/model.py:87-107; conv3x3+BN+ReLU
// stack, residual blocks :30-36, 1x1 heads, 2880->8100 policy FC, value MLP + tanh) on the
// inference path of self-play (model.py:109-124 predict, inference_server.py:251-263).
// BatchNorm is folded into the conv weights/bias on the host (eval mode, model.py:118).
//
// Data layout ("channel-chunk planes").  An activation tensor is stored as
//     X[chunk = C/8][row][8 channels]   (bf16, 16 bytes per (chunk,row))
// where `row` walks the boards with a zero halo: board b, cell (r,c) lives at row
// b*110 + (r+1)*10 + c; row slots with c == 9 and the 10 slots before each board are zero
// and double as left/right/top/bottom padding for every neighbour, so a 3x3 tap (dy,dx) is
// the SAME matrix shifted by dy*10+dx rows.  Consequences:
//   * the A operand of a 128-row output tile for all 9 taps is one block of 150 rows per chunk,
//     fetched ONCE per tile with 1-D TMA bulk copies (2400 contiguous bytes per chunk);
//   * in shared memory the block is the canonical no-swizzle K-major UMMA layout with
//     SBO = 128 B, i.e. address = base + row*16 + chunk*LBO -- linear in the row, so each tap
//     only moves the descriptor start address by shift*16 bytes; no im2col, no re-load;
//   * the epilogue stores 16 B per (chunk,row) with consecutive lanes on consecutive rows:
//     fully coalesced, and the next layer's TMA reads exactly what was written.
// Weights are pre-arranged on the host as per-iteration shared-memory images
// [n_tile][tap][k_block][chunk][n][8] so a pipeline stage is one contiguous bulk copy.
//
// Kernel: persistent, warp-specialised -- warp 0 lane 0 = TMA producer, warp 1 = TMEM owner +
// single-thread tcgen05.mma issuer, warps 2-5 = epilogue (tcgen05.ld -> bias/residual/ReLU ->
// bf16 -> global).  mbarrier rings: w_full/w_empty per stage, a_full/a_empty for the resident A
// block, tmem_full/tmem_empty for the accumulator.  Two CTAs per SM (128 TMEM columns each)
// let one CTA's epilogue overlap the other's MMAs.
#include "xq_ctx.h"

#include <cuda_bf16.h>
#include <cstdlib>

namespace xq {

struct GemmArgs {
    int mode, m_tiles, n_tiles, kchunks, relu, n_boards;
    long long a_rows, a_row0, out_rows, out_row0, out_stride;
    const uint8_t* a;
    const uint8_t* w;
    const float* bias;
    const uint8_t* residual;
    uint8_t* out;
    float* out2;
    const uint8_t* w_half;   // weight image tiled by 64 output channels (cta_group::2 kernel: each CTA of a pair holds half of N)
    int dbg;   // timing experiments only (XQ_NET_DBG): 1 = no weight copies, 2 = quarter of the MMAs, 4 = empty epilogue
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives when every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// ---- programmatic dependent launch: a kernel launched with the PDL attribute may start while its
// predecessor is still draining; pdl_wait() blocks until the predecessor has completed and flushed, so
// everything before it (barrier init, TMEM allocation) overlaps the predecessor's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- thread-block cluster helpers (weight-stage multicast) ----
__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// global -> shared of EVERY CTA in cta_mask (same offset), completing tx bytes on each CTA's own barrier
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint16_t cta_mask)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
// arrive on the barrier at this offset in every CTA of cta_mask once the MMAs issued so far are done
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, no swizzle, K-major (cute/arch/mma_sm100_desc.hpp layout):
// [0,14) start>>4, [16,30) LBO>>4 (stride between the two 8-element K chunks of one MMA),
// [32,46) SBO>>4 (stride between 8-row groups), [46,48) version = 1, [61,64) layout = 0.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major A/B, N>>3 at 17, M>>4 at 24
__host__ __device__ constexpr uint32_t make_idesc(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

constexpr int kGemmThreads = 192;
constexpr int kARows = 150;                 // 128 + 11 halo rows either side
constexpr int kAPlane = kARows * 16;        // bytes per chunk of the resident A block
constexpr int kHalo = 11;                   // largest |dy*10+dx|

template <int MODE, int NT, int KCH>
struct GemmCfg {
    static constexpr int kStages = (MODE == 2) ? 3 : 4;
    static constexpr int kWStage = KCH * NT * 16;
    static constexpr int kAStage = (MODE == 2) ? KCH * 128 * 16 : 0;
    static constexpr int kStageBytes = kWStage + kAStage;
    static constexpr int kTmemCols = NT > 64 ? 128 : 64;
    static constexpr int kTaps = (MODE == 0) ? 9 : 1;
    static int smem_bytes(int kchunks)
    {
        int a = (MODE == 2) ? 0 : kchunks * kAPlane;
        a = (a + 127) & ~127;
        return a + kStages * kStageBytes + 256;
    }
};

// MODE 0: 3x3 conv (9 shifted taps), epilogue = bias (+residual) (+ReLU), halo rows zeroed, plane output
// MODE 1: 1x1 head conv (policy 32 ch + value 4 ch, NT = 48), epilogue scatters into the FC's A planes / value features
// MODE 2: dense FC (A streamed per k-block), epilogue = bias, row-major bf16 logits
template <int MODE, int NT, int KCH>
__global__ void __launch_bounds__(kGemmThreads) gemm_kernel(const GemmArgs p)
{
    using Cfg = GemmCfg<MODE, NT, KCH>;
    constexpr int S = Cfg::kStages;
    extern __shared__ __align__(128) uint8_t smem[];
    const int a_res = (MODE == 2) ? 0 : ((p.kchunks * kAPlane + 127) & ~127);
    uint8_t* sA = smem;
    uint8_t* sStage = smem + a_res;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + S * Cfg::kStageBytes);
    uint64_t* w_full = bars;
    uint64_t* w_empty = bars + S;
    uint64_t* a_full = bars + 2 * S;
    uint64_t* a_empty = bars + 2 * S + 1;
    uint64_t* t_full = bars + 2 * S + 2;
    uint64_t* t_empty = bars + 2 * S + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = p.kchunks / KCH;
    const int iters = Cfg::kTaps * kblocks;
    const int total_tiles = p.m_tiles * p.n_tiles;
    __shared__ __align__(16) float sBiasT[2][128];

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
        }
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        mbar_init(t_full, 1);
        mbar_init(t_empty, 4);      // one arrival per epilogue warp
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer (one thread) =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0, tile_ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
                const long long m0 = (long long)m_tile * 128;
                if (MODE != 2) {
                    mbar_wait(a_empty, tile_ph ^ 1);
                    mbar_expect_tx(a_full, (uint32_t)(p.kchunks * kAPlane));
                    for (int c = 0; c < p.kchunks; ++c)
                        bulk_g2s(sA + c * kAPlane, p.a + ((size_t)c * p.a_rows + (size_t)(p.a_row0 + m0 - kHalo)) * 16,
                                 kAPlane, a_full);
                }
                const uint8_t* wt = p.w + (size_t)n_tile * iters * Cfg::kWStage;
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&w_empty[s], ph ^ 1);
                    uint8_t* st = sStage + s * Cfg::kStageBytes;
                    mbar_expect_tx(&w_full[s], Cfg::kStageBytes);
                    bulk_g2s(st, wt + (size_t)it * Cfg::kWStage, Cfg::kWStage, &w_full[s]);
                    if (MODE == 2) {
#pragma unroll
                        for (int c = 0; c < KCH; ++c)
                            bulk_g2s(st + Cfg::kWStage + c * 2048,
                                     p.a + ((size_t)(it * KCH + c) * p.a_rows + (size_t)(p.a_row0 + m0)) * 16, 2048,
                                     &w_full[s]);
                    }
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                tile_ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (one thread) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, NT);
            int s = 0;
            uint32_t ph = 0, tile_ph = 0;
            const uint32_t sA_addr = smem_u32(sA);
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(t_empty, tile_ph ^ 1);     // epilogue has drained the accumulator
                tc_fence_after();
                if (MODE != 2) mbar_wait(a_full, tile_ph);
                for (int it = 0; it < iters; ++it) {
                    const int tap = (MODE == 0) ? it / kblocks : 0;
                    const int kb = it - tap * kblocks;
                    const int shift = (MODE == 0) ? ((tap / 3) - 1) * 10 + (tap % 3) - 1 : 0;
                    mbar_wait(&w_full[s], ph);
                    tc_fence_after();
                    const uint32_t st_addr = smem_u32(sStage + s * Cfg::kStageBytes);
                    constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);        // SBO = 128 B, version 1
                    // low descriptor words are additive in 16-byte units: build them once per stage
                    const uint32_t a_lo = (MODE == 2)
                        ? ((((st_addr + Cfg::kWStage) >> 4) & 0x3FFFu) | ((2048u >> 4) << 16))
                        : ((((sA_addr + (uint32_t)(kb * KCH * kAPlane + (kHalo + shift) * 16)) >> 4) & 0x3FFFu) | ((uint32_t)(kAPlane >> 4) << 16));
                    const uint32_t a_step = (MODE == 2) ? (2u * 2048u) >> 4 : (2u * kAPlane) >> 4;
                    const uint32_t b_lo = ((st_addr >> 4) & 0x3FFFu) | ((uint32_t)((NT * 16) >> 4) << 16);
#pragma unroll
                    for (int j = 0; j < KCH / 2; ++j) {
                        const uint64_t adesc = ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + (uint32_t)j * a_step);
                        const uint64_t bdesc = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + (uint32_t)(2 * j * NT));
                        umma_bf16(tmem_base, adesc, bdesc, idesc, (it | j) != 0 ? 1u : 0u);
                    }
                    umma_commit(&w_empty[s]);        // frees the stage when these MMAs have read it
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                umma_commit(t_full);
                if (MODE != 2) umma_commit(a_empty);
                tile_ph ^= 1;
            }
        }
    } else {
        // ===================== epilogue (4 warps, one TMEM lane = one output row each) =====================
        const int q = warp & 3;                      // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        uint32_t tile_ph = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int m_tile = tile / p.n_tiles, n_tile = tile - m_tile * p.n_tiles;
            const long long m = (long long)m_tile * 128 + row;
            if (MODE == 2) {
                // this tile's 128 bias values -> shared memory (double buffered by tile parity), one per epilogue thread
                sBiasT[tile_ph][row] = p.bias[n_tile * NT + row];
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            mbar_wait(t_full, tile_ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);

            if (MODE == 0) {
                const int rr = (int)(m % 110);
                const bool real = m < (long long)p.n_boards * 110 && rr >= 10 && (rr % 10) != 9;
#pragma unroll 1
                for (int c0 = 0; c0 < NT; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int n = n_tile * NT + c0 + g * 8;
                        const size_t off = ((size_t)(n >> 3) * p.out_rows + (size_t)(p.out_row0 + m)) * 16;
                        uint4 o = make_uint4(0, 0, 0, 0);
                        if (real) {
                            float f[8];
                            const float4 b0 = *reinterpret_cast<const float4*>(p.bias + n);
                            const float4 b1 = *reinterpret_cast<const float4*>(p.bias + n + 4);
                            f[0] = __uint_as_float(v[g * 8 + 0]) + b0.x;
                            f[1] = __uint_as_float(v[g * 8 + 1]) + b0.y;
                            f[2] = __uint_as_float(v[g * 8 + 2]) + b0.z;
                            f[3] = __uint_as_float(v[g * 8 + 3]) + b0.w;
                            f[4] = __uint_as_float(v[g * 8 + 4]) + b1.x;
                            f[5] = __uint_as_float(v[g * 8 + 5]) + b1.y;
                            f[6] = __uint_as_float(v[g * 8 + 6]) + b1.z;
                            f[7] = __uint_as_float(v[g * 8 + 7]) + b1.w;
                            if (p.residual) {
                                const uint4 r = *reinterpret_cast<const uint4*>(p.residual + off);
                                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    f[2 * k] += __uint_as_float(rw[k] << 16);
                                    f[2 * k + 1] += __uint_as_float(rw[k] & 0xffff0000u);
                                }
                            }
                            if (p.relu) {
#pragma unroll
                                for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.0f);
                            }
                            o = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                           pack_bf16(f[6], f[7]));
                        }
                        *reinterpret_cast<uint4*>(p.out + off) = o;
                    }
                }
            } else if (MODE == 1) {
                const int rr = (int)(m % 110);
                const bool real = m < (long long)p.n_boards * 110 && rr >= 10 && (rr % 10) != 9;
                const long long b = m / 110;
                const int pos = (rr / 10 - 1) * 9 + (rr % 10);
                uint32_t v[32], v2[16];
                tmem_ld32(taddr, v);
                tmem_ld16(taddr + 32, v2);
                tmem_ld_wait();
                if (real) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float f[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k) f[k] = fmaxf(__uint_as_float(v[g * 8 + k]) + p.bias[g * 8 + k], 0.0f);
                        const uint4 o = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                                   pack_bf16(f[6], f[7]));
                        // FC A plane (pos*4 + g), row b
                        *reinterpret_cast<uint4*>(p.out + ((size_t)(pos * 4 + g) * p.out_rows + (size_t)(p.out_row0 + b)) * 16) = o;
                    }
                    float4 vf;
                    vf.x = fmaxf(__uint_as_float(v2[0]) + p.bias[32], 0.0f);
                    vf.y = fmaxf(__uint_as_float(v2[1]) + p.bias[33], 0.0f);
                    vf.z = fmaxf(__uint_as_float(v2[2]) + p.bias[34], 0.0f);
                    vf.w = fmaxf(__uint_as_float(v2[3]) + p.bias[35], 0.0f);
                    *reinterpret_cast<float4*>(p.out2 + ((size_t)b * 90 + pos) * 4) = vf;
                }
            } else {
                const bool real = m < (long long)p.n_boards;
#pragma unroll 1
                for (int c0 = 0; c0 < NT; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    tmem_ld_wait();
                    if (real) {
                        const int n = n_tile * NT + c0;
                        uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)m * p.out_stride + n) * 2);
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const float4 b0 = *reinterpret_cast<const float4*>(&sBiasT[tile_ph][c0 + g * 8]);
                            const float4 b1 = *reinterpret_cast<const float4*>(&sBiasT[tile_ph][c0 + g * 8 + 4]);
                            dst[g] = make_uint4(pack_bf16(__uint_as_float(v[g * 8 + 0]) + b0.x, __uint_as_float(v[g * 8 + 1]) + b0.y),
                                                pack_bf16(__uint_as_float(v[g * 8 + 2]) + b0.z, __uint_as_float(v[g * 8 + 3]) + b0.w),
                                                pack_bf16(__uint_as_float(v[g * 8 + 4]) + b1.x, __uint_as_float(v[g * 8 + 5]) + b1.y),
                                                pack_bf16(__uint_as_float(v[g * 8 + 6]) + b1.z, __uint_as_float(v[g * 8 + 7]) + b1.w));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty);
            tile_ph ^= 1;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// =============================================================================================
// v2 conv kernel: TWO 128-row tiles per weight stage + double-buffered accumulators and A blocks
// =============================================================================================
// v1 (above) streams 288 KB of weights + 38 KB of A per 128-row tile and serialises A-load -> MMA ->
// epilogue inside a CTA (ncu: tensor pipe 30-45 % active, L2 25 %).  v2 keeps one CTA per SM and
//   * feeds each 16 KB weight stage to two row tiles (accumulators t=0,1): L2->SM bytes per row halve;
//   * double-buffers the accumulator pair in TMEM (2 x 2 x 128 = all 512 columns) so the epilogue of
//     pair i runs under the MMAs of pair i+1, and double-buffers the 278-row A block so its TMA load
//     is issued a whole pair ahead;
//   * prefetches the residual operand into registers before the accumulator is ready (v1 issued 16
//     dependent global loads per row after the MMAs: +60 us per layer).
constexpr int kPairRows = 256;
constexpr int kARows2 = kPairRows + 2 * kHalo;   // 278
constexpr int kAPlane2 = kARows2 * 16;           // 4448 B per chunk

template <int NT, int KCH, bool HEADS, int ABUFS>
struct Conv2Cfg {
    static constexpr int kStages = 4;
    static constexpr int kWStage = KCH * NT * 16;
    static constexpr int kTileCols = NT > 64 ? 128 : 64;       // TMEM columns per row tile
    static constexpr int kTmemCols = 4 * kTileCols;            // 2 accumulator stages x 2 tiles
    static constexpr int kTaps = HEADS ? 1 : 9;
    static int smem_bytes(int kchunks)
    {
        int a = ABUFS * ((kchunks * kAPlane2 + 127) & ~127);
        int total = a + kStages * kWStage + 256;
        return total < 120 * 1024 ? 120 * 1024 : total;        // > half an SM: exactly one CTA per SM (it owns all of TMEM)
    }
};

template <int NT, int KCH, bool HEADS, int ABUFS, int CL>
__global__ void __launch_bounds__(kGemmThreads, 1) conv2_kernel(const GemmArgs p)
{
    using Cfg = Conv2Cfg<NT, KCH, HEADS, ABUFS>;
    // CL > 1: the CTAs of a cluster work on CL consecutive row pairs with the SAME weights; each CTA fetches
    // 1/CL of every weight stage and multicasts it to all of them, so L2 -> SM weight traffic drops CL-fold.
    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
    const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    constexpr int S = Cfg::kStages;
    constexpr int TS = Cfg::kTileCols;
    extern __shared__ __align__(128) uint8_t smem[];
    const int a_buf_bytes = (p.kchunks * kAPlane2 + 127) & ~127;
    uint8_t* sA = smem;
    uint8_t* sStage = smem + ABUFS * a_buf_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + S * Cfg::kWStage);
    uint64_t* w_full = bars;
    uint64_t* w_empty = bars + S;
    uint64_t* a_full = bars + 2 * S;          // [ABUFS]
    uint64_t* a_empty = bars + 2 * S + 2;     // [ABUFS]
    uint64_t* t_full = bars + 2 * S + 4;      // [2]
    uint64_t* t_empty = bars + 2 * S + 6;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kblocks = p.kchunks / KCH;
    const int iters = Cfg::kTaps * kblocks;
    const int m_pairs = (p.m_tiles + 1) / 2;
    const int groups = (m_pairs + CL - 1) / CL;        // CL row pairs per work item (one per CTA of the cluster)
    const int total = groups * p.n_tiles;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], CL);                // every CTA of the cluster releases the stage
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], 4);
        }
        mbar_fence_init();
    }
    // bias is uniform across rows and tiles: stage it once (the first version re-read it from global memory
    // in every tile -- ncu showed the epilogue stalled on those dependent loads, and it paced the whole kernel)
    __shared__ __align__(16) float sBias[256];
    for (int i = threadIdx.x; i < p.n_tiles * NT && i < 256; i += kGemmThreads) sBias[i] = p.bias[i];
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                    // peers' barriers are initialised before anything lands on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int n = 0;
            for (int work = cluster_id; work < total; work += n_clusters, ++n) {
                const int grp = work / p.n_tiles, n_tile = work - grp * p.n_tiles;
                const int pair = grp * CL + (int)crank;
                const long long m0 = (long long)pair * kPairRows;
                const int ab = n % ABUFS;
                const uint32_t aph = (uint32_t)(n / ABUFS) & 1u;
                mbar_wait(&a_empty[ab], aph ^ 1);
                mbar_expect_tx(&a_full[ab], (uint32_t)(p.kchunks * kAPlane2));
                for (int c = 0; c < p.kchunks; ++c)
                    bulk_g2s(sA + ab * a_buf_bytes + c * kAPlane2,
                             p.a + ((size_t)c * p.a_rows + (size_t)(p.a_row0 + m0 - kHalo)) * 16, kAPlane2, &a_full[ab]);
                const uint8_t* wt = p.w + (size_t)n_tile * iters * Cfg::kWStage;
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&w_empty[s], ph ^ 1);
                    if (p.dbg & 1) {
                        mbar_arrive(&w_full[s]);
                        if (++s == S) { s = 0; ph ^= 1; }
                        continue;
                    }
                    mbar_expect_tx(&w_full[s], Cfg::kWStage);
                    if (CL == 1) {
                        bulk_g2s(sStage + s * Cfg::kWStage, wt + (size_t)it * Cfg::kWStage, Cfg::kWStage, &w_full[s]);
                    } else {
                        constexpr int kSlice = Cfg::kWStage / CL;
                        bulk_g2s_mc(sStage + s * Cfg::kWStage + crank * kSlice, wt + (size_t)it * Cfg::kWStage + crank * kSlice,
                                    kSlice, &w_full[s], kMask);
                    }
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, NT);
            int s = 0;
            uint32_t ph = 0;
            int n = 0;
            for (int work = cluster_id; work < total; work += n_clusters, ++n) {
                const int ab = n % ABUFS;
                const uint32_t aph = (uint32_t)(n / ABUFS) & 1u;
                const int acc = n & 1;
                const uint32_t tph = (uint32_t)(n >> 1) & 1u;
                mbar_wait(&t_empty[acc], tph ^ 1);
                tc_fence_after();
                mbar_wait(&a_full[ab], aph);
                // Descriptors are 16-byte-granular and additive in their low word, so the issue loop only adds
                // small constants (the first version rebuilt both 64-bit descriptors per MMA: ~25 dependent
                // instructions per MMA from ONE thread paced the tensor pipe at ~45 % -- see DESIGN.md).
                const uint32_t a_lo0 = (((smem_u32(sA + ab * a_buf_bytes)) >> 4) & 0x3FFFu) | ((uint32_t)(kAPlane2 >> 4) << 16);
                constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);            // SBO = 128 B, version 1
                const uint32_t d_addr = tmem_base + (uint32_t)(acc * 2 * TS);
                for (int it = 0; it < iters; ++it) {
                    const int tap = HEADS ? 0 : it / kblocks;
                    const int kb = it - tap * kblocks;
                    const int shift = HEADS ? 0 : ((tap / 3) - 1) * 10 + (tap % 3) - 1;
                    const uint32_t a_lo = a_lo0 + (uint32_t)(kb * KCH * (kAPlane2 >> 4) + kHalo + shift);
                    mbar_wait(&w_full[s], ph);
                    tc_fence_after();
                    const uint32_t b_lo = ((smem_u32(sStage + s * Cfg::kWStage) >> 4) & 0x3FFFu) | ((uint32_t)((NT * 16) >> 4) << 16);
                    const uint32_t first = it != 0 ? 1u : 0u;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
#pragma unroll
                        for (int j = 0; j < KCH / 2; ++j) {
                            if ((p.dbg & 2) && j > 0) continue;
                            const uint64_t adesc = ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + (uint32_t)(2 * j * (kAPlane2 >> 4) + t * 128));
                            const uint64_t bdesc = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + (uint32_t)(2 * j * NT));
                            umma_bf16(d_addr + (uint32_t)(t * TS), adesc, bdesc, idesc, j == 0 ? first : 1u);
                        }
                    }
                    if (CL == 1) umma_commit(&w_empty[s]);
                    else umma_commit_mc(&w_empty[s], kMask);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                umma_commit(&t_full[acc]);
                umma_commit(&a_empty[ab]);
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        int n = 0;
        uint4 res[NT / 8];
        for (int work = cluster_id; work < total; work += n_clusters, ++n) {
            const int grp = work / p.n_tiles, n_tile = work - grp * p.n_tiles;
            const int pair = grp * CL + (int)crank;
            const int acc = n & 1;
            const uint32_t tph = (uint32_t)(n >> 1) & 1u;
            const long long mrow[2] = {(long long)pair * kPairRows + row, (long long)pair * kPairRows + 128 + row};
            bool real[2];
            int rr[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                rr[t] = (int)(mrow[t] % 110);
                real[t] = mrow[t] < (long long)p.n_boards * 110 && rr[t] >= 10 && (rr[t] % 10) != 9;
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * TS);

            if (!HEADS) {
                // Residual operand: registers res[k] always hold the NEXT tile's values -- tile 0 of the first
                // pair is loaded before the loop, tile 1 is loaded while tile 0 is processed, and tile 0 of the
                // next pair while tile 1 is processed, so the loads are in flight for half a pair (~3 us).
                const bool has_res = p.residual != nullptr;
                if (n == 0) {
#pragma unroll
                    for (int k = 0; k < NT / 8; ++k) {
                        res[k] = make_uint4(0, 0, 0, 0);
                        if (has_res && real[0])
                            res[k] = __ldg(reinterpret_cast<const uint4*>(
                                p.residual + ((size_t)(n_tile * (NT / 8) + k) * p.out_rows + (size_t)(p.out_row0 + mrow[0])) * 16));
                    }
                }
                // row / validity of tile 0 of this CTA's next work item
                const int nwork = work + n_clusters;
                const int ngrp = nwork / p.n_tiles, nn_tile = nwork - ngrp * p.n_tiles;
                const long long nrow = (long long)(ngrp * CL + (int)crank) * kPairRows + row;
                const int nrr = (int)(nrow % 110);
                const bool nreal = nwork < total && nrow < (long long)p.n_boards * 110 && nrr >= 10 && (nrr % 10) != 9;
                mbar_wait(&t_full[acc], tph);
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (p.dbg & 4) continue;
#pragma unroll
                    for (int c0 = 0; c0 < NT; c0 += 32) {
                        uint32_t v[32];
                        tmem_ld32(taddr + (uint32_t)(t * TS + c0), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const int k = c0 / 8 + g;
                            const int nn = n_tile * NT + c0 + g * 8;
                            const size_t off = ((size_t)(nn >> 3) * p.out_rows + (size_t)(p.out_row0 + mrow[t])) * 16;
                            uint4 o = make_uint4(0, 0, 0, 0);
                            if (real[t]) {
                                float f[8];
                                const float4 b0 = *reinterpret_cast<const float4*>(&sBias[nn]);
                                const float4 b1 = *reinterpret_cast<const float4*>(&sBias[nn + 4]);
                                f[0] = __uint_as_float(v[g * 8 + 0]) + b0.x;
                                f[1] = __uint_as_float(v[g * 8 + 1]) + b0.y;
                                f[2] = __uint_as_float(v[g * 8 + 2]) + b0.z;
                                f[3] = __uint_as_float(v[g * 8 + 3]) + b0.w;
                                f[4] = __uint_as_float(v[g * 8 + 4]) + b1.x;
                                f[5] = __uint_as_float(v[g * 8 + 5]) + b1.y;
                                f[6] = __uint_as_float(v[g * 8 + 6]) + b1.z;
                                f[7] = __uint_as_float(v[g * 8 + 7]) + b1.w;
                                const uint32_t rw[4] = {res[k].x, res[k].y, res[k].z, res[k].w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    f[2 * e] += __uint_as_float(rw[e] << 16);
                                    f[2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
                                }
                                if (p.relu) {
#pragma unroll
                                    for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.0f);
                                }
                                o = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                               pack_bf16(f[6], f[7]));
                            }
                            *reinterpret_cast<uint4*>(p.out + off) = o;
                            // this register is free again: refill it for the tile after this one
                            res[k] = make_uint4(0, 0, 0, 0);
                            if (t == 0) {
                                if (has_res && real[1])
                                    res[k] = __ldg(reinterpret_cast<const uint4*>(
                                        p.residual + ((size_t)(nn >> 3) * p.out_rows + (size_t)(p.out_row0 + mrow[1])) * 16));
                            } else {
                                if (has_res && nreal)
                                    res[k] = __ldg(reinterpret_cast<const uint4*>(
                                        p.residual + ((size_t)(nn_tile * (NT / 8) + k) * p.out_rows + (size_t)(p.out_row0 + nrow)) * 16));
                            }
                        }
                    }
                }
            } else {
                mbar_wait(&t_full[acc], tph);
                tc_fence_after();
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const long long b = mrow[t] / 110;
                    const int pos = (rr[t] / 10 - 1) * 9 + (rr[t] % 10);
                    uint32_t v[32], v2[16];
                    tmem_ld32(taddr + (uint32_t)(t * TS), v);
                    tmem_ld16(taddr + (uint32_t)(t * TS + 32), v2);
                    tmem_ld_wait();
                    if (real[t]) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float f[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[g * 8 + e]) + sBias[g * 8 + e], 0.0f);
                            const uint4 o = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                                       pack_bf16(f[6], f[7]));
                            *reinterpret_cast<uint4*>(p.out + ((size_t)(pos * 4 + g) * p.out_rows + (size_t)(p.out_row0 + b)) * 16) = o;
                        }
                        float4 vf;
                        vf.x = fmaxf(__uint_as_float(v2[0]) + sBias[32], 0.0f);
                        vf.y = fmaxf(__uint_as_float(v2[1]) + sBias[33], 0.0f);
                        vf.z = fmaxf(__uint_as_float(v2[2]) + sBias[34], 0.0f);
                        vf.w = fmaxf(__uint_as_float(v2[3]) + sBias[35], 0.0f);
                        *reinterpret_cast<float4*>(p.out2 + ((size_t)b * 90 + pos) * 4) = vf;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                    // nobody leaves while a peer can still write to it
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int NT, int KCH, bool HEADS, int ABUFS, int CL>
static int launch_conv2(xq_ctx* c, const GemmArgs& a, cudaStream_t s)
{
    using Cfg = Conv2Cfg<NT, KCH, HEADS, ABUFS>;
    const int smem = Cfg::smem_bytes(a.kchunks);
    static bool configured = false;
    auto kern = conv2_kernel<NT, KCH, HEADS, ABUFS, CL>;
    if (!configured) {
        XQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));   // + 1 KB static (bias)
        configured = true;
    }
    if (smem > 225 * 1024) return xq_fail(c, XQ_ERR_ARG, "conv2 kernel needs %d bytes of shared memory", smem);
    const int groups = (((a.m_tiles + 1) / 2) + CL - 1) / CL;
    const int total = groups * a.n_tiles;
    int clusters = c->sm_count / CL;                           // persistent: one CTA per SM, whole clusters only
    if (clusters > total) clusters = total;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CL);
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CL > 1 ? 1 : 0;
    XQ_CUDA(c, cudaLaunchKernelEx(&cfg, kern, a));
    c->launches += 1;
    return XQ_OK;
}

// =============================================================================================
// v4 conv kernel: deep weight ring + rolling A segments
// =============================================================================================
// Measured on v2/v3 (XQ_NET_DBG experiments, profiles/r1_net_notes.md): a weight stage cannot be refilled
// before the MMAs that read it have completed, and that completion -> refill -> MMA-issue round trip is
// ~1.5-2.5 us.  A 4-stage ring holds only ~1.2 us of MMA work, so the tensor pipe idled ~55 % of the time no
// matter how cheap the copies or the epilogue were.  v4 keeps the tile-pair scheme and
//   * splits the resident A block into per-k-block SEGMENTS with their own full/empty barriers and walks
//     k-blocks in the OUTER loop, so a segment is released after its 9 taps and re-filled with the next
//     pair's rows while the other k-block computes -- A needs one buffer instead of two;
//   * spends the freed shared memory on a 9-deep weight ring (2.6 us of MMA work in flight);
//   * lets one stage carry several taps (TPS) when a tap is tiny (15-plane input conv: all 9 taps, 36 KB).
constexpr int kConv4Threads = 352;   // warp 0 TMA, warps 1 and 10 MMA issuers (row tile 0 / 1), warps 2-9 epilogue
constexpr int kMaxStages4 = 9;
constexpr int kMaxSeg4 = 4;

template <int NT, int KCH, bool HEADS, int TPS>
struct Conv4Cfg {
    static constexpr int kWTap = KCH * NT * 16;                // bytes of one (tap, k-block) weight slice
    static constexpr int kWStage = TPS * kWTap;
    static constexpr int kSeg = KCH * kAPlane2;                // bytes of one A segment (one k-block, 278 rows)
    static constexpr int kTileCols = NT > 64 ? 128 : 64;
    static constexpr int kTmemCols = 4 * kTileCols;
    static constexpr int kTaps = HEADS ? 1 : 9;
    static constexpr int kBudget = 225 * 1024 - 1024 - 512;    // dynamic smem minus static bias minus barriers
    static int stages(int kchunks)
    {
        int st = (kBudget - (kchunks / KCH) * kSeg) / kWStage;
        return st > kMaxStages4 ? kMaxStages4 : st;
    }
    static int smem_bytes(int kchunks)
    {
        int total = (kchunks / KCH) * kSeg + stages(kchunks) * kWStage + 512;
        return total < 120 * 1024 ? 120 * 1024 : total;        // one CTA per SM: it owns all of TMEM
    }
};

template <int NT, int KCH, bool HEADS, int TPS>
__global__ void __launch_bounds__(kConv4Threads, 1) conv4_kernel(const GemmArgs p, const int S)
{
    using Cfg = Conv4Cfg<NT, KCH, HEADS, TPS>;
    constexpr int TS = Cfg::kTileCols;
    extern __shared__ __align__(128) uint8_t smem[];
    const int kblocks = p.kchunks / KCH;
    uint8_t* sA = smem;                                         // [kblocks][KCH][278][16 B]
    uint8_t* sStage = smem + kblocks * Cfg::kSeg;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + S * Cfg::kWStage);
    uint64_t* w_full = bars;                                    // [kMaxStages4]
    uint64_t* w_empty = bars + kMaxStages4;
    uint64_t* a_full = bars + 2 * kMaxStages4;                  // [kMaxSeg4]
    uint64_t* a_empty = a_full + kMaxSeg4;
    uint64_t* t_full = a_empty + kMaxSeg4;                      // [2 accumulator stages][2 row tiles]
    uint64_t* t_empty = t_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tap_groups = Cfg::kTaps / TPS;
    const int m_pairs = (p.m_tiles + 1) / 2;
    const int total = m_pairs * p.n_tiles;

    __shared__ __align__(16) float sBias[256];
    for (int i = threadIdx.x; i < p.n_tiles * NT && i < 256; i += kConv4Threads) sBias[i] = p.bias[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 2);                 // both MMA issuers release a stage
        }
        for (int i = 0; i < kblocks; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 2);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&t_full[i], 1);                  // per (accumulator stage, row tile)
            mbar_init(&t_empty[i], 4);                 // the 4 epilogue warps (lane quarters) of that row tile
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_launch_dependents();                           // the next layer may start its prologue now
    pdl_wait();                                        // activations written by the previous kernel are visible after this

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int n = 0;
            for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
                const int pair = work / p.n_tiles, n_tile = work - pair * p.n_tiles;
                const long long m0 = (long long)pair * kPairRows;
                const uint8_t* wt = p.w + (size_t)n_tile * Cfg::kTaps * kblocks * Cfg::kWTap;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&a_empty[kb], (uint32_t)(n & 1) ^ 1u);
                    if (p.dbg & 8) {                          // timing experiment: no activation traffic
                        mbar_arrive(&a_full[kb]);
                    } else {
                        mbar_expect_tx(&a_full[kb], (uint32_t)Cfg::kSeg);
#pragma unroll
                        for (int c = 0; c < KCH; ++c)
                            bulk_g2s(sA + kb * Cfg::kSeg + c * kAPlane2,
                                     p.a + ((size_t)(kb * KCH + c) * p.a_rows + (size_t)(p.a_row0 + m0 - kHalo)) * 16, kAPlane2,
                                     &a_full[kb]);
                    }
                    for (int tg = 0; tg < tap_groups; ++tg) {
                        mbar_wait(&w_empty[s], ph ^ 1);
                        if (p.dbg & 1) {                      // timing experiment: no weight traffic
                            mbar_arrive(&w_full[s]);
                            if (++s == S) { s = 0; ph ^= 1; }
                            continue;
                        }
                        mbar_expect_tx(&w_full[s], Cfg::kWStage);
                        // host image order is [tap][k_block]: one copy per tap of the stage
#pragma unroll
                        for (int tp = 0; tp < TPS; ++tp)
                            bulk_g2s(sStage + s * Cfg::kWStage + tp * Cfg::kWTap,
                                     wt + (size_t)((tg * TPS + tp) * kblocks + kb) * Cfg::kWTap, Cfg::kWTap, &w_full[s]);
                        if (++s == S) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 10) {
        // ===================== MMA issuers: one thread per row tile =====================
        // A single issuing thread spends ~0.2 us per stage hand-off (wait, fence, commit) on top of ~66 cycles per
        // tcgen05.mma, which kept the tensor pipe at ~90 cycles per MMA; two issuers interleave their MMAs and
        // hide each other's hand-offs.
        const int t = warp == 1 ? 0 : 1;
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, NT);
            constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);            // SBO = 128 B, version 1
            constexpr uint32_t kLboA = (uint32_t)(kAPlane2 >> 4);
            int s = 0;
            uint32_t ph = 0;
            int n = 0;
            for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
                const int acc = n & 1;
                const uint32_t tph = (uint32_t)(n >> 1) & 1u;
                mbar_wait(&t_empty[acc * 2 + t], tph ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(acc * 2 * TS + t * TS);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&a_full[kb], (uint32_t)(n & 1));
                    const uint32_t a_seg = (((smem_u32(sA + kb * Cfg::kSeg)) >> 4) & 0x3FFFu) | (kLboA << 16);
                    for (int tg = 0; tg < tap_groups; ++tg) {
                        mbar_wait(&w_full[s], ph);
                        tc_fence_after();
                        const uint32_t b_st = ((smem_u32(sStage + s * Cfg::kWStage) >> 4) & 0x3FFFu) | ((uint32_t)((NT * 16) >> 4) << 16);
#pragma unroll
                        for (int tp = 0; tp < TPS; ++tp) {
                            const int tap = tg * TPS + tp;
                            const int shift = HEADS ? 0 : ((tap / 3) - 1) * 10 + (tap % 3) - 1;
                            const uint32_t a_lo = a_seg + (uint32_t)(kHalo + shift);
                            const uint32_t b_lo = b_st + (uint32_t)(tp * (Cfg::kWTap >> 4));
                            const uint32_t first = (kb | tap) != 0 ? 1u : 0u;
#pragma unroll
                            for (int j = 0; j < KCH / 2; ++j) {
                                if ((p.dbg & 2) && j > 0) continue;   // timing experiment: a quarter of the MMAs
                                const uint64_t adesc = ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + (uint32_t)(2 * j) * kLboA + (uint32_t)(t * 128));
                                const uint64_t bdesc = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + (uint32_t)(2 * j * NT));
                                umma_bf16(d_addr, adesc, bdesc, idesc, j == 0 ? first : 1u);
                            }
                        }
                        umma_commit(&w_empty[s]);
                        if (++s == S) { s = 0; ph ^= 1; }
                    }
                    umma_commit(&a_empty[kb]);                  // this k-block's rows may be replaced by the next pair's
                }
                umma_commit(&t_full[acc * 2 + t]);
            }
        }
    } else {
        // ===================== epilogue: 8 warps, warp -> (row tile t, TMEM lane quarter q) =====================
        const int q = warp & 3;                          // a warp may only read TMEM lanes 32*(warp%4) ..
        const int t = (warp - 2) >> 2;                   // warps 2-5: tile 0, warps 6-9: tile 1
        const int row = q * 32 + lane;
        int n = 0;
        uint4 res[NT / 8];
        for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
            const int pair = work / p.n_tiles, n_tile = work - pair * p.n_tiles;
            const int acc = n & 1;
            const uint32_t tph = (uint32_t)(n >> 1) & 1u;
            const long long mrow = (long long)pair * kPairRows + t * 128 + row;
            const int rr = (int)(mrow % 110);
            const bool real = mrow < (long long)p.n_boards * 110 && rr >= 10 && (rr % 10) != 9;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * TS + t * TS);

            if (!HEADS) {
                // Residual operand: res[k] always holds the values of this warp's NEXT tile; loaded before the
                // first wait, then refilled slab by slab for the next work item, i.e. a whole pair (~5 us) ahead.
                const bool has_res = p.residual != nullptr;
                if (n == 0) {
#pragma unroll
                    for (int k = 0; k < NT / 8; ++k) {
                        res[k] = make_uint4(0, 0, 0, 0);
                        if (has_res && real)
                            res[k] = __ldg(reinterpret_cast<const uint4*>(
                                p.residual + ((size_t)(n_tile * (NT / 8) + k) * p.out_rows + (size_t)(p.out_row0 + mrow)) * 16));
                    }
                }
                const int nwork = work + gridDim.x;
                const int npair = nwork / p.n_tiles, nn_tile = nwork - npair * p.n_tiles;
                const long long nrow = (long long)npair * kPairRows + t * 128 + row;
                const int nrr = (int)(nrow % 110);
                const bool nreal = nwork < total && nrow < (long long)p.n_boards * 110 && nrr >= 10 && (nrr % 10) != 9;
                mbar_wait(&t_full[acc * 2 + t], tph);
                tc_fence_after();
                // one 32-column slab: bias (+residual) (+ReLU), halo rows -> 0, bf16, 4 coalesced 16-byte stores
                auto emit = [&](const uint32_t* v, const int c0) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int k = c0 / 8 + g;
                        const int nn = n_tile * NT + c0 + g * 8;
                        const size_t off = ((size_t)(nn >> 3) * p.out_rows + (size_t)(p.out_row0 + mrow)) * 16;
                        uint4 o = make_uint4(0, 0, 0, 0);
                        if (real) {
                            float f[8];
                            const float4 b0 = *reinterpret_cast<const float4*>(&sBias[nn]);
                            const float4 b1 = *reinterpret_cast<const float4*>(&sBias[nn + 4]);
                            f[0] = __uint_as_float(v[g * 8 + 0]) + b0.x;
                            f[1] = __uint_as_float(v[g * 8 + 1]) + b0.y;
                            f[2] = __uint_as_float(v[g * 8 + 2]) + b0.z;
                            f[3] = __uint_as_float(v[g * 8 + 3]) + b0.w;
                            f[4] = __uint_as_float(v[g * 8 + 4]) + b1.x;
                            f[5] = __uint_as_float(v[g * 8 + 5]) + b1.y;
                            f[6] = __uint_as_float(v[g * 8 + 6]) + b1.z;
                            f[7] = __uint_as_float(v[g * 8 + 7]) + b1.w;
                            const uint32_t rw[4] = {res[k].x, res[k].y, res[k].z, res[k].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                f[2 * e] += __uint_as_float(rw[e] << 16);
                                f[2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
                            }
                            if (p.relu) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.0f);
                            }
                            o = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                        }
                        *reinterpret_cast<uint4*>(p.out + off) = o;
                        res[k] = make_uint4(0, 0, 0, 0);
                        if (has_res && nreal)
                            res[k] = __ldg(reinterpret_cast<const uint4*>(
                                p.residual + ((size_t)(nn_tile * (NT / 8) + k) * p.out_rows + (size_t)(p.out_row0 + nrow)) * 16));
                    }
                };
                if (!(p.dbg & 4)) {
                    // software pipeline over the NT/32 slabs: the TMEM load of slab i+1 is in flight while slab i is
                    // converted and stored (tcgen05.wait::ld waits for ALL outstanding loads, so it follows emit)
                    constexpr int kSlabs = NT / 32;
                    uint32_t va[32], vb[32];
                    tmem_ld32(taddr, va);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < kSlabs; ++i) {
                        if (i & 1) {
                            if (i + 1 < kSlabs) tmem_ld32(taddr + (uint32_t)((i + 1) * 32), va);
                            emit(vb, i * 32);
                        } else {
                            if (i + 1 < kSlabs) tmem_ld32(taddr + (uint32_t)((i + 1) * 32), vb);
                            emit(va, i * 32);
                        }
                        if (i + 1 < kSlabs) tmem_ld_wait();
                    }
                }
            } else {
                mbar_wait(&t_full[acc * 2 + t], tph);
                tc_fence_after();
                const long long b = mrow / 110;
                const int pos = (rr / 10 - 1) * 9 + (rr % 10);
                uint32_t v[32], v2[16];
                tmem_ld32(taddr, v);
                tmem_ld16(taddr + 32u, v2);
                tmem_ld_wait();
                if (real) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = fmaxf(__uint_as_float(v[g * 8 + e]) + sBias[g * 8 + e], 0.0f);
                        const uint4 o = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                                                   pack_bf16(f[6], f[7]));
                        *reinterpret_cast<uint4*>(p.out + ((size_t)(pos * 4 + g) * p.out_rows + (size_t)(p.out_row0 + b)) * 16) = o;
                    }
                    float4 vf;
                    vf.x = fmaxf(__uint_as_float(v2[0]) + sBias[32], 0.0f);
                    vf.y = fmaxf(__uint_as_float(v2[1]) + sBias[33], 0.0f);
                    vf.z = fmaxf(__uint_as_float(v2[2]) + sBias[34], 0.0f);
                    vf.w = fmaxf(__uint_as_float(v2[3]) + sBias[35], 0.0f);
                    *reinterpret_cast<float4*>(p.out2 + ((size_t)b * 90 + pos) * 4) = vf;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc * 2 + t]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

template <int NT, int KCH, bool HEADS, int TPS>
static int launch_conv4(xq_ctx* c, const GemmArgs& a, cudaStream_t s)
{
    using Cfg = Conv4Cfg<NT, KCH, HEADS, TPS>;
    const int kblocks = a.kchunks / KCH;
    const int S = Cfg::stages(a.kchunks);
    if (kblocks > kMaxSeg4 || S < 2) return xq_fail(c, XQ_ERR_ARG, "conv4 kernel: %d k-blocks, %d stages do not fit", kblocks, S);
    const int smem = Cfg::smem_bytes(a.kchunks);
    static bool configured = false;
    auto kern = conv4_kernel<NT, KCH, HEADS, TPS>;
    if (!configured) {
        XQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
        configured = true;
    }
    const int total = ((a.m_tiles + 1) / 2) * a.n_tiles;
    const int grid = c->sm_count < total ? c->sm_count : total;   // persistent, one CTA per SM
    if (getenv("XQ_DEBUG")) fprintf(stderr, "[xq] conv4<%d,%d,%d,%d> stages=%d smem=%d grid=%d\n", NT, KCH, (int)HEADS, TPS, S, smem, grid);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kConv4Threads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = c->net_pdl ? 1 : 0;
        XQ_CUDA(c, cudaLaunchKernelEx(&cfg, kern, a, S));
    }
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

// =============================================================================================
// v5 conv kernel: v4 on a CTA pair with cta_group::2 MMAs (M = 256 across two SMs)
// =============================================================================================
// v4 is bound by shared-memory bandwidth: every 128x128x16 MMA reads 4 KB of A and 4 KB of B, and the TMA
// writes share the same 128 B/clk.  With cta_group::2 the two SMs of a pair run ONE 256x128x16 MMA: each CTA
// supplies its own 128 rows of A and only HALF of B (64 of the 128 output channels, its own 8 KB weight
// stage); the tensor cores exchange the halves.  B reads and weight-stage writes per SM halve.
//   * cluster (2,1,1); CTA r of pair works on row pair 2*item + r; rank 0 is the MMA leader
//   * both CTAs run their own TMA producer (own A segments, own half of every weight stage); the follower's
//     warp 1 relays "my stage / my segment has landed" to the leader with remote mbarrier arrives
//   * the leader's commits are multicast to both CTAs (stage empty, segment empty, accumulator full);
//     both epilogues arrive on the leader's accumulator-empty barrier
constexpr int kMaxStages5 = 18;

__device__ __forceinline__ void mbar_arrive_remote(uint64_t* local_bar, uint32_t target_cta)
{
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(local_bar)),
        "r"(target_cta)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2()
{
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

template <int KCH>
struct Conv5Cfg {
    static constexpr int NT = 128;                             // output channels per MMA (64 per CTA)
    static constexpr int kWTap = KCH * 64 * 16;                // this CTA's half of one (tap, k-block) weight slice
    static constexpr int kWStage = kWTap;
    static constexpr int kSeg = KCH * kAPlane2;
    static constexpr int kTmemCols = 512;
    static constexpr int kBudget = 225 * 1024 - 1024 - 1024;
    static int stages(int kchunks)
    {
        int st = (kBudget - (kchunks / KCH) * kSeg) / kWStage;
        return st > kMaxStages5 ? kMaxStages5 : st;
    }
    static int smem_bytes(int kchunks)
    {
        int total = (kchunks / KCH) * kSeg + stages(kchunks) * kWStage + 1024;
        return total < 120 * 1024 ? 120 * 1024 : total;
    }
};

template <int KCH>
__global__ void __launch_bounds__(kConv4Threads, 1) conv5_kernel(const GemmArgs p, const int S)
{
    using Cfg = Conv5Cfg<KCH>;
    constexpr int NT = 128, TS = 128;
    extern __shared__ __align__(128) uint8_t smem[];
    const int kblocks = p.kchunks / KCH;
    uint8_t* sA = smem;
    uint8_t* sStage = smem + kblocks * Cfg::kSeg;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + S * Cfg::kWStage);
    uint64_t* w_full = bars;                                    // [18] own half of the stage has landed
    uint64_t* w_empty = bars + kMaxStages5;                     // [18] MMAs reading the stage are done (leader multicast)
    uint64_t* w_peer = bars + 2 * kMaxStages5;                  // [18] leader only: the follower's half has landed
    uint64_t* a_full = bars + 3 * kMaxStages5;                  // [4]
    uint64_t* a_empty = a_full + kMaxSeg4;
    uint64_t* a_peer = a_empty + kMaxSeg4;
    uint64_t* t_full = a_peer + kMaxSeg4;                       // [2]
    uint64_t* t_empty = t_full + 2;                             // [2] leader only: 16 epilogue warps of the pair
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader = crank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int m_pairs = (p.m_tiles + 1) / 2;
    const int items = (m_pairs + 1) / 2;
    const int total = items * p.n_tiles;

    __shared__ __align__(16) float sBias[256];
    for (int i = threadIdx.x; i < p.n_tiles * NT && i < 256; i += kConv4Threads) sBias[i] = p.bias[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
            mbar_init(&w_peer[i], 1);
        }
        for (int i = 0; i < kblocks; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&a_peer[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], 16);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc2(tmem_slot, Cfg::kTmemCols);
        tmem_relinquish2();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs): own rows, own half of the weights =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int n = 0;
            for (int work = cluster_id; work < total; work += n_clusters, ++n) {
                const int item = work / p.n_tiles, n_tile = work - item * p.n_tiles;
                const int pair = item * 2 + (int)crank;
                const long long m0 = (long long)pair * kPairRows;
                // image tiled by 64 channels: [n_tile64][tap][k_block][chunk][64][8], n_tile64 = 2 * n_tile + rank
                const uint8_t* wt = p.w_half + (size_t)(2 * n_tile + (int)crank) * 9 * kblocks * Cfg::kWTap;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&a_empty[kb], (uint32_t)(n & 1) ^ 1u);
                    mbar_expect_tx(&a_full[kb], (uint32_t)Cfg::kSeg);
#pragma unroll
                    for (int c = 0; c < KCH; ++c)
                        bulk_g2s(sA + kb * Cfg::kSeg + c * kAPlane2,
                                 p.a + ((size_t)(kb * KCH + c) * p.a_rows + (size_t)(p.a_row0 + m0 - kHalo)) * 16, kAPlane2,
                                 &a_full[kb]);
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&w_empty[s], ph ^ 1);
                        mbar_expect_tx(&w_full[s], Cfg::kWStage);
                        bulk_g2s(sStage + s * Cfg::kWStage, wt + (size_t)(tap * kblocks + kb) * Cfg::kWTap, Cfg::kWStage, &w_full[s]);
                        if (++s == S) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int n = 0;
            if (!leader) {
                // ===================== follower: relay "landed" events to the leader =====================
                for (int work = cluster_id; work < total; work += n_clusters, ++n) {
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&a_full[kb], (uint32_t)(n & 1));
                        mbar_arrive_remote(&a_peer[kb], 0);
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(&w_full[s], ph);
                            mbar_arrive_remote(&w_peer[s], 0);
                            if (++s == S) { s = 0; ph ^= 1; }
                        }
                    }
                }
            } else {
                // ===================== leader: MMA issuer for both SMs =====================
                constexpr uint32_t idesc = make_idesc(256, NT);
                constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
                constexpr uint32_t kLboA = (uint32_t)(kAPlane2 >> 4);
                for (int work = cluster_id; work < total; work += n_clusters, ++n) {
                    const int acc = n & 1;
                    const uint32_t tph = (uint32_t)(n >> 1) & 1u;
                    mbar_wait(&t_empty[acc], tph ^ 1);
                    tc_fence_after();
                    const uint32_t d_addr = tmem_base + (uint32_t)(acc * 2 * TS);
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&a_full[kb], (uint32_t)(n & 1));
                        mbar_wait(&a_peer[kb], (uint32_t)(n & 1));
                        const uint32_t a_seg = (((smem_u32(sA + kb * Cfg::kSeg)) >> 4) & 0x3FFFu) | (kLboA << 16);
                        for (int tap = 0; tap < 9; ++tap) {
                            mbar_wait(&w_full[s], ph);
                            mbar_wait(&w_peer[s], ph);
                            tc_fence_after();
                            const uint32_t b_lo = ((smem_u32(sStage + s * Cfg::kWStage) >> 4) & 0x3FFFu) | ((uint32_t)((64 * 16) >> 4) << 16);
                            const int shift = ((tap / 3) - 1) * 10 + (tap % 3) - 1;
                            const uint32_t a_lo = a_seg + (uint32_t)(kHalo + shift);
                            const uint32_t first = (kb | tap) != 0 ? 1u : 0u;
#pragma unroll
                            for (int t = 0; t < 2; ++t) {
#pragma unroll
                                for (int j = 0; j < KCH / 2; ++j) {
                                    const uint64_t adesc = ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + (uint32_t)(2 * j) * kLboA + (uint32_t)(t * 128));
                                    const uint64_t bdesc = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + (uint32_t)(2 * j * 64));
                                    umma2_bf16(d_addr + (uint32_t)(t * TS), adesc, bdesc, idesc, j == 0 ? first : 1u);
                                }
                            }
                            umma2_commit_mc(&w_empty[s], 3);
                            if (++s == S) { s = 0; ph ^= 1; }
                        }
                        umma2_commit_mc(&a_empty[kb], 3);
                    }
                    umma2_commit_mc(&t_full[acc], 3);
                }
            }
        }
    } else if (warp < 10) {
        // ===================== epilogue (both CTAs): as v4, accumulator release goes to the leader =====================
        const int q = warp & 3;
        const int t = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        int n = 0;
        uint4 res[NT / 8];
        for (int work = cluster_id; work < total; work += n_clusters, ++n) {
            const int item = work / p.n_tiles, n_tile = work - item * p.n_tiles;
            const int pair = item * 2 + (int)crank;
            const int acc = n & 1;
            const uint32_t tph = (uint32_t)(n >> 1) & 1u;
            const long long mrow = (long long)pair * kPairRows + t * 128 + row;
            const int rr = (int)(mrow % 110);
            const bool real = mrow < (long long)p.n_boards * 110 && rr >= 10 && (rr % 10) != 9;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * TS + t * TS);
            const bool has_res = p.residual != nullptr;
            if (n == 0) {
#pragma unroll
                for (int k = 0; k < NT / 8; ++k) {
                    res[k] = make_uint4(0, 0, 0, 0);
                    if (has_res && real)
                        res[k] = __ldg(reinterpret_cast<const uint4*>(
                            p.residual + ((size_t)(n_tile * (NT / 8) + k) * p.out_rows + (size_t)(p.out_row0 + mrow)) * 16));
                }
            }
            const int nwork = work + n_clusters;
            const int nitem = nwork / p.n_tiles, nn_tile = nwork - nitem * p.n_tiles;
            const long long nrow = (long long)(nitem * 2 + (int)crank) * kPairRows + t * 128 + row;
            const int nrr = (int)(nrow % 110);
            const bool nreal = nwork < total && nrow < (long long)p.n_boards * 110 && nrr >= 10 && (nrr % 10) != 9;
            mbar_wait(&t_full[acc], tph);
            tc_fence_after();
            auto emit = [&](const uint32_t* v, const int c0) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int k = c0 / 8 + g;
                    const int nn = n_tile * NT + c0 + g * 8;
                    const size_t off = ((size_t)(nn >> 3) * p.out_rows + (size_t)(p.out_row0 + mrow)) * 16;
                    uint4 o = make_uint4(0, 0, 0, 0);
                    if (real) {
                        float f[8];
                        const float4 b0 = *reinterpret_cast<const float4*>(&sBias[nn]);
                        const float4 b1 = *reinterpret_cast<const float4*>(&sBias[nn + 4]);
                        f[0] = __uint_as_float(v[g * 8 + 0]) + b0.x;
                        f[1] = __uint_as_float(v[g * 8 + 1]) + b0.y;
                        f[2] = __uint_as_float(v[g * 8 + 2]) + b0.z;
                        f[3] = __uint_as_float(v[g * 8 + 3]) + b0.w;
                        f[4] = __uint_as_float(v[g * 8 + 4]) + b1.x;
                        f[5] = __uint_as_float(v[g * 8 + 5]) + b1.y;
                        f[6] = __uint_as_float(v[g * 8 + 6]) + b1.z;
                        f[7] = __uint_as_float(v[g * 8 + 7]) + b1.w;
                        const uint32_t rw[4] = {res[k].x, res[k].y, res[k].z, res[k].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            f[2 * e] += __uint_as_float(rw[e] << 16);
                            f[2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
                        }
                        if (p.relu) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.0f);
                        }
                        o = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                    }
                    *reinterpret_cast<uint4*>(p.out + off) = o;
                    res[k] = make_uint4(0, 0, 0, 0);
                    if (has_res && nreal)
                        res[k] = __ldg(reinterpret_cast<const uint4*>(
                            p.residual + ((size_t)(nn_tile * (NT / 8) + k) * p.out_rows + (size_t)(p.out_row0 + nrow)) * 16));
                }
            };
            {
                constexpr int kSlabs = NT / 32;
                uint32_t va[32], vb[32];
                tmem_ld32(taddr, va);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < kSlabs; ++i) {
                    if (i & 1) {
                        if (i + 1 < kSlabs) tmem_ld32(taddr + (uint32_t)((i + 1) * 32), va);
                        emit(vb, i * 32);
                    } else {
                        if (i + 1 < kSlabs) tmem_ld32(taddr + (uint32_t)((i + 1) * 32), vb);
                        emit(va, i * 32);
                    }
                    if (i + 1 < kSlabs) tmem_ld_wait();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&t_empty[acc]);
                else mbar_arrive_remote(&t_empty[acc], 0);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                // both CTAs are done with TMEM and with each other's barriers
    if (warp == 1) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
}

template <int KCH>
static int launch_conv5(xq_ctx* c, const GemmArgs& a, cudaStream_t s)
{
    using Cfg = Conv5Cfg<KCH>;
    const int kblocks = a.kchunks / KCH;
    const int S = Cfg::stages(a.kchunks);
    if (kblocks > kMaxSeg4 || S < 2 || !a.w_half) return xq_fail(c, XQ_ERR_ARG, "conv5 kernel: %d k-blocks, %d stages, w_half %p", kblocks, S, (const void*)a.w_half);
    const int smem = Cfg::smem_bytes(a.kchunks);
    static bool configured = false;
    auto kern = conv5_kernel<KCH>;
    if (!configured) {
        XQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
        configured = true;
    }
    const int items = ((((a.m_tiles + 1) / 2) + 1) / 2) * a.n_tiles;
    int clusters = c->sm_count / 2;
    if (clusters > items) clusters = items;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * 2);
    cfg.blockDim = dim3(kConv4Threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (getenv("XQ_DEBUG")) fprintf(stderr, "[xq] conv5<%d> stages=%d smem=%d grid=%d\n", KCH, S, smem, clusters * 2);
    XQ_CUDA(c, cudaLaunchKernelEx(&cfg, kern, a, S));
    c->launches += 1;
    return XQ_OK;
}

// =============================================================================================
// fc4: dense layer (policy FC 2880 -> 8100) in the conv4 style
// =============================================================================================
// The first FC kernel (gemm_kernel<2>: 128x128 tiles, 3-stage ring, 2 CTAs/SM) spent more time handing
// stages over than multiplying (4 MMAs per hand-off).  fc4: one CTA per SM, a 256-board x 128-output work item
// (two M tiles share each weight stage, one MMA-issuing thread per tile), 48 KB stages (W 16 KB + A 2 x 16 KB,
// the two M tiles are adjacent rows of the A planes so a chunk is ONE 4 KB bulk copy), 4-deep ring,
// accumulators double buffered in TMEM (2 x 2 x 128 columns), epilogue on 8 warps.
template <int KCH_, int STAGES_>
struct Fc4CfgT {
    static constexpr int kKch = KCH_;                          // 8 input features per chunk
    static constexpr int kWBytes = kKch * 128 * 16;
    static constexpr int kABytes = kKch * 256 * 16;
    static constexpr int kStage = kWBytes + kABytes;
    static constexpr int kStages = STAGES_;
    static constexpr int kSmem = kStages * kStage + 512;
};

template <class Fc4Cfg>
__global__ void __launch_bounds__(kConv4Threads, 1) fc4_kernel(const GemmArgs p)
{
    constexpr int S = Fc4Cfg::kStages, TS = 128, NT = 128;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Fc4Cfg::kStage);
    uint64_t* w_full = bars;
    uint64_t* w_empty = bars + S;
    uint64_t* t_full = bars + 2 * S;          // [2 accumulator stages][2 tiles]
    uint64_t* t_empty = t_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 4);
    __shared__ __align__(16) float sBias[2][128];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int iters = p.kchunks / Fc4Cfg::kKch;
    const int m_pairs = (p.m_tiles + 1) / 2;
    const int total = m_pairs * p.n_tiles;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 2);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], 4);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int work = blockIdx.x; work < total; work += gridDim.x) {
                const int pair = work / p.n_tiles, n_tile = work - pair * p.n_tiles;
                const long long m0 = (long long)pair * 256;
                const uint8_t* wt = p.w + (size_t)n_tile * iters * Fc4Cfg::kWBytes;
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&w_empty[s], ph ^ 1);
                    uint8_t* st = smem + s * Fc4Cfg::kStage;
                    mbar_expect_tx(&w_full[s], Fc4Cfg::kStage);
                    bulk_g2s(st, wt + (size_t)it * Fc4Cfg::kWBytes, Fc4Cfg::kWBytes, &w_full[s]);
#pragma unroll
                    for (int c = 0; c < Fc4Cfg::kKch; ++c)
                        bulk_g2s(st + Fc4Cfg::kWBytes + c * 4096,
                                 p.a + ((size_t)(it * Fc4Cfg::kKch + c) * p.a_rows + (size_t)(p.a_row0 + m0)) * 16, 4096, &w_full[s]);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1 || warp == 10) {
        const int t = warp == 1 ? 0 : 1;
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(128, NT);
            constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
            int s = 0;
            uint32_t ph = 0;
            int n = 0;
            for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
                const int acc = n & 1;
                const uint32_t tph = (uint32_t)(n >> 1) & 1u;
                mbar_wait(&t_empty[acc * 2 + t], tph ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(acc * 2 * TS + t * TS);
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&w_full[s], ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * Fc4Cfg::kStage);
                    const uint32_t b_lo = ((st >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);
                    const uint32_t a_lo = (((st + Fc4Cfg::kWBytes + (uint32_t)t * 2048u) >> 4) & 0x3FFFu) | ((4096u >> 4) << 16);
#pragma unroll
                    for (int j = 0; j < Fc4Cfg::kKch / 2; ++j) {
                        const uint64_t adesc = ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + (uint32_t)(2 * j) * (4096u >> 4));
                        const uint64_t bdesc = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + (uint32_t)(2 * j * NT));
                        umma_bf16(d_addr, adesc, bdesc, idesc, (it | j) != 0 ? 1u : 0u);
                    }
                    umma_commit(&w_empty[s]);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                umma_commit(&t_full[acc * 2 + t]);
            }
        }
    } else {
        const int q = warp & 3;
        const int t = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;          // 0..255 among the epilogue threads
        int n = 0;
        for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
            const int pair = work / p.n_tiles, n_tile = work - pair * p.n_tiles;
            const int acc = n & 1;
            const uint32_t tph = (uint32_t)(n >> 1) & 1u;
            const long long m = (long long)pair * 256 + t * 128 + row;
            const bool real = m < (long long)p.n_boards;
            // this work item's 128 bias values -> shared memory (double buffered by item parity)
            if (et < 128) sBias[acc][et] = p.bias[n_tile * NT + et];
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * TS + t * TS);
            mbar_wait(&t_full[acc * 2 + t], tph);
            tc_fence_after();
            auto emit = [&](const uint32_t* v, const int c0) {
                if (!real) return;
                uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)m * p.out_stride + (size_t)(n_tile * NT + c0)) * 2);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 b0 = *reinterpret_cast<const float4*>(&sBias[acc][c0 + g * 8]);
                    const float4 b1 = *reinterpret_cast<const float4*>(&sBias[acc][c0 + g * 8 + 4]);
                    dst[g] = make_uint4(pack_bf16(__uint_as_float(v[g * 8 + 0]) + b0.x, __uint_as_float(v[g * 8 + 1]) + b0.y),
                                        pack_bf16(__uint_as_float(v[g * 8 + 2]) + b0.z, __uint_as_float(v[g * 8 + 3]) + b0.w),
                                        pack_bf16(__uint_as_float(v[g * 8 + 4]) + b1.x, __uint_as_float(v[g * 8 + 5]) + b1.y),
                                        pack_bf16(__uint_as_float(v[g * 8 + 6]) + b1.z, __uint_as_float(v[g * 8 + 7]) + b1.w));
                }
            };
            uint32_t va[32], vb[32];
            tmem_ld32(taddr, va);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i & 1) {
                    if (i + 1 < 4) tmem_ld32(taddr + (uint32_t)((i + 1) * 32), va);
                    emit(vb, i * 32);
                } else {
                    if (i + 1 < 4) tmem_ld32(taddr + (uint32_t)((i + 1) * 32), vb);
                    emit(va, i * 32);
                }
                if (i + 1 < 4) tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[acc * 2 + t]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <class Fc4Cfg>
static int launch_fc4(xq_ctx* c, const GemmArgs& a, cudaStream_t s)
{
    static bool configured = false;
    if (!configured) {
        XQ_CUDA(c, cudaFuncSetAttribute(fc4_kernel<Fc4Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
        configured = true;
    }
    if (a.kchunks % Fc4Cfg::kKch) return xq_fail(c, XQ_ERR_ARG, "fc4: K/8 = %d is not a multiple of %d", a.kchunks, Fc4Cfg::kKch);
    const int total = ((a.m_tiles + 1) / 2) * a.n_tiles;
    const int grid = c->sm_count < total ? c->sm_count : total;
    fc4_kernel<Fc4Cfg><<<grid, kConv4Threads, Fc4Cfg::kSmem, s>>>(a);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

// ---- value head: Linear(360,128)+ReLU -> Linear(128,1) -> tanh (model.py:74-83) ----------------
// feats [B][90][4] fp32 (already conv1x1+BN+ReLU), w1t [360][128] fp32 with k = pos*4+ch, 16 boards per CTA.
constexpr int kVhBoards = 16;
constexpr int kVhChunk = 40;   // rows of W1^T staged per step (40 x 128 floats = 20 KB)
__global__ void __launch_bounds__(128) value_head_kernel(const float* __restrict__ feats, const float* __restrict__ w1t,
                                                          const float* __restrict__ b1, const float* __restrict__ w2,
                                                          float b2, float* __restrict__ value, int B)
{
    __shared__ __align__(16) float f[kVhBoards][360];
    __shared__ __align__(16) float wsm[kVhChunk][128];
    __shared__ float red[kVhBoards][4];
    const int b0 = blockIdx.x * kVhBoards;
    const int nb = min(kVhBoards, B - b0);
    pdl_launch_dependents();
    pdl_wait();
    for (int i = threadIdx.x; i < kVhBoards * 360; i += 128) {
        const int bb = i / 360;
        f[bb][i - bb * 360] = bb < nb ? feats[(size_t)(b0 + bb) * 360 + (i - bb * 360)] : 0.0f;
    }
    const int j = threadIdx.x;
    float acc[kVhBoards];
#pragma unroll
    for (int bb = 0; bb < kVhBoards; ++bb) acc[bb] = b1[j];
    for (int k0 = 0; k0 < 360; k0 += kVhChunk) {
        __syncthreads();
        // coalesced float4 copy of 40 rows of W1^T (the first version read one dependent global word per k)
        for (int i = threadIdx.x; i < kVhChunk * 32; i += 128)
            reinterpret_cast<float4*>(&wsm[0][0])[i] = reinterpret_cast<const float4*>(w1t + (size_t)k0 * 128)[i];
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kVhChunk; ++k) {
            const float w = wsm[k][j];
#pragma unroll
            for (int bb = 0; bb < kVhBoards; ++bb) acc[bb] = fmaf(w, f[bb][k0 + k], acc[bb]);
        }
    }
    const float wj = w2[j];
#pragma unroll
    for (int bb = 0; bb < kVhBoards; ++bb) {
        float t = fmaxf(acc[bb], 0.0f) * wj;
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) red[bb][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x < nb) {
        const float s = red[threadIdx.x][0] + red[threadIdx.x][1] + red[threadIdx.x][2] + red[threadIdx.x][3] + b2;
        value[b0 + threadIdx.x] = tanhf(s);
    }
}

template <int MODE, int NT, int KCH>
static int launch_gemm(xq_ctx* c, const GemmArgs& a, cudaStream_t s)
{
    using Cfg = GemmCfg<MODE, NT, KCH>;
    const int smem = Cfg::smem_bytes(a.kchunks);
    static bool configured = false;
    static int per_sm = 1;
    if (!configured) {
        XQ_CUDA(c, cudaFuncSetAttribute(gemm_kernel<MODE, NT, KCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        // without this the driver keeps a smaller shared-memory carve-out and only ONE CTA fits per SM
        XQ_CUDA(c, cudaFuncSetAttribute(gemm_kernel<MODE, NT, KCH>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        (int)cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    // two CTAs per SM when shared memory (227 KB, 1 KB reserved per CTA) and TMEM (512 columns) allow it.
    // (cudaOccupancyMaxActiveBlocksPerMultiprocessor reported 1 here even with the max carve-out requested;
    // the persistent tile loop is correct for any grid size, so the launch does not depend on it.)
    per_sm = (227 * 1024) / (smem + 1024);
    if (per_sm < 1) return xq_fail(c, XQ_ERR_ARG, "gemm kernel does not fit: %d bytes of shared memory", smem);
    const int tmem_cap = 512 / Cfg::kTmemCols;
    if (per_sm > tmem_cap) per_sm = tmem_cap;
    if (per_sm > 2) per_sm = 2;
    const int total = a.m_tiles * a.n_tiles;
    int grid = c->sm_count * per_sm;
    if (grid > total) grid = total;
    if (getenv("XQ_DEBUG")) fprintf(stderr, "[xq] gemm<%d,%d,%d> per_sm=%d grid=%d smem=%d\n", MODE, NT, KCH, per_sm, grid, smem);
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(kGemmThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = c->net_pdl ? 1 : 0;
        XQ_CUDA(c, cudaLaunchKernelEx(&cfg, gemm_kernel<MODE, NT, KCH>, a));
    }
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

}  // namespace xq

using namespace xq;

extern "C" void xq_net_free_(xq_ctx*) {}

extern "C" int xq_net_gemm(xq_ctx* c, const xq_gemm_desc* d, void* stream)
{
    if (!c || !d) return xq_fail(c, XQ_ERR_ARG, "xq_net_gemm: NULL argument");
    if (!d->a || !d->w || !d->bias || !d->out || d->m_tiles <= 0 || d->n_tiles <= 0 || d->kchunks <= 0)
        return xq_fail(c, XQ_ERR_ARG, "xq_net_gemm: bad descriptor");
    GemmArgs a;
    a.mode = d->mode;
    a.m_tiles = d->m_tiles;
    a.n_tiles = d->n_tiles;
    a.kchunks = d->kchunks;
    a.relu = d->relu;
    a.n_boards = d->n_boards;
    a.a_rows = d->a_rows;
    a.a_row0 = d->a_row0;
    a.out_rows = d->out_rows;
    a.out_row0 = d->out_row0;
    a.out_stride = d->out_stride;
    a.a = (const uint8_t*)d->a;
    a.w = (const uint8_t*)d->w;
    a.bias = d->bias;
    a.residual = (const uint8_t*)d->residual;
    a.out = (uint8_t*)d->out;
    a.out2 = (float*)d->out2;
    a.w_half = (const uint8_t*)d->w_half;
    a.dbg = getenv("XQ_NET_DBG") ? atoi(getenv("XQ_NET_DBG")) : 0;
    cudaStream_t s = (cudaStream_t)stream;
    XqTimer tm(c, s);
    const bool v1 = c->net_v1;
    const int cl = c->net_cluster;
    if (!v1 && c->net_gen >= 5 && d->w_half && d->mode == 0 && d->nt == 128 && d->kch_iter == 8 && d->kchunks % 8 == 0 && d->kchunks <= 32)
        return launch_conv5<8>(c, a, s);
    if (!v1 && c->net_gen >= 4 && c->net_fc4 && d->mode == 2 && d->nt == 128 && d->kch_iter == 8 && d->kchunks % 8 == 0) {
        const int v = getenv("XQ_NET_FCK") ? atoi(getenv("XQ_NET_FCK")) : 12;
        if (v == 16 && d->kchunks % 16 == 0) return launch_fc4<Fc4CfgT<16, 2>>(c, a, s);
        if (v == 12 && d->kchunks % 12 == 0) return launch_fc4<Fc4CfgT<12, 3>>(c, a, s);
        return launch_fc4<Fc4CfgT<8, 4>>(c, a, s);
    }
    if (!v1 && c->net_gen >= 4) {
        // 128-channel layers: 3 taps per weight stage (48 KB x 3 stages): a stage hand-off costs the single MMA-issuing
        // thread ~0.2 us (measured with XQ_NET_DBG=15), so fewer, larger stages beat a finer ring
        if (d->mode == 0 && d->nt == 128 && d->kch_iter == 8 && d->kchunks == 16 && c->net_tps == 3) return launch_conv4<128, 8, false, 3>(c, a, s);
        if (d->mode == 0 && d->nt == 128 && d->kch_iter == 8 && d->kchunks % 8 == 0 && d->kchunks <= 32) return launch_conv4<128, 8, false, 1>(c, a, s);
        if (d->mode == 0 && d->nt == 128 && d->kch_iter == 2 && d->kchunks == 2) return launch_conv4<128, 2, false, 9>(c, a, s);
        if (d->mode == 1 && d->nt == 48 && d->kch_iter == 8 && d->kchunks % 8 == 0 && d->kchunks <= 32 && d->out2) return launch_conv4<48, 8, true, 1>(c, a, s);
    }
    if (!v1 && d->mode == 0 && d->nt == 128 && d->kch_iter == 8 && d->kchunks % 8 == 0 && d->kchunks <= 16) {
        if (cl == 4) return launch_conv2<128, 8, false, 2, 4>(c, a, s);
        if (cl == 2) return launch_conv2<128, 8, false, 2, 2>(c, a, s);
        return launch_conv2<128, 8, false, 2, 1>(c, a, s);
    }
    if (!v1 && d->mode == 0 && d->nt == 128 && d->kch_iter == 8 && d->kchunks % 8 == 0 && d->kchunks <= 32) {
        if (cl >= 2) return launch_conv2<128, 8, false, 1, 2>(c, a, s);
        return launch_conv2<128, 8, false, 1, 1>(c, a, s);
    }
    if (!v1 && d->mode == 0 && d->nt == 128 && d->kch_iter == 2 && d->kchunks == 2) return launch_conv2<128, 2, false, 2, 1>(c, a, s);
    if (!v1 && d->mode == 1 && d->nt == 48 && d->kch_iter == 8 && d->kchunks % 8 == 0 && d->kchunks <= 16 && d->out2) return launch_conv2<48, 8, true, 2, 1>(c, a, s);
    if (!v1 && d->mode == 1 && d->nt == 48 && d->kch_iter == 8 && d->kchunks % 8 == 0 && d->kchunks <= 32 && d->out2) return launch_conv2<48, 8, true, 1, 1>(c, a, s);
    if (d->mode == 0 && d->nt == 128 && d->kch_iter == 8 && d->kchunks % 8 == 0) return launch_gemm<0, 128, 8>(c, a, s);
    if (d->mode == 0 && d->nt == 128 && d->kch_iter == 2 && d->kchunks == 2) return launch_gemm<0, 128, 2>(c, a, s);
    if (d->mode == 1 && d->nt == 48 && d->kch_iter == 8 && d->kchunks % 8 == 0 && d->out2) return launch_gemm<1, 48, 8>(c, a, s);
    if (d->mode == 2 && d->nt == 128 && d->kch_iter == 8 && d->kchunks % 8 == 0) return launch_gemm<2, 128, 8>(c, a, s);
    return xq_fail(c, XQ_ERR_ARG, "xq_net_gemm: unsupported (mode=%d nt=%d kch_iter=%d kchunks=%d)", d->mode, d->nt,
                   d->kch_iter, d->kchunks);
}

extern "C" int xq_net_value_head(xq_ctx* c, const float* d_feats, const float* d_w1t, const float* d_b1,
                                 const float* d_w2, float b2, float* d_value, int B, void* stream)
{
    if (!c || !d_feats || !d_w1t || !d_b1 || !d_w2 || !d_value || B < 0)
        return xq_fail(c, XQ_ERR_ARG, "xq_net_value_head: bad arguments");
    if (B == 0) return XQ_OK;
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((B + kVhBoards - 1) / kVhBoards);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = c->net_pdl ? 1 : 0;
        XQ_CUDA(c, cudaLaunchKernelEx(&cfg, value_head_kernel, d_feats, d_w1t, d_b1, d_w2, b2, d_value, B));
    }
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

// Run a whole forward (a list of layer descriptors followed by the value head) from one call.
extern "C" int xq_net_run(xq_ctx* c, const xq_gemm_desc* layers, int n_layers, const float* d_vfeats,
                          const float* d_w1t, const float* d_b1, const float* d_w2, float b2, float* d_value, int B,
                          void* stream)
{
    for (int i = 0; i < n_layers; ++i) {
        int rc = xq_net_gemm(c, &layers[i], stream);
        if (rc) return rc;
    }
    return xq_net_value_head(c, d_vfeats, d_w1t, d_b1, d_w2, b2, d_value, B, stream);
}

// ---- micro-benchmark: issue rate of SS-mode tcgen05.mma from resident no-swizzle operands ----------
namespace xq {
template <int NT>
__global__ void __launch_bounds__(128) umma_rate_kernel(long long* out, int n_mma, int a_stride_rows)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = make_idesc(128, NT);
        const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 48 * 1024);
        long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const int j = i & 3, t = (i >> 2) & 1;
            const uint64_t ad = make_desc(a0 + (uint32_t)(2 * j * 4448 + (t * 128 + (i % 9)) * 16), 4448, 128);
            const uint64_t bd = make_desc(b0 + (uint32_t)(2 * j * NT * 16), NT * 16, 128);
            umma_bf16(tb + (uint32_t)(t * NT), ad, bd, idesc, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}
}  // namespace xq

extern "C" int xq_debug_umma_rate(xq_ctx* c, int nt, int n_mma, int grid, long long* h_cycles)
{
    long long* d = nullptr;
    XQ_CUDA(c, cudaMalloc(&d, sizeof(long long) * grid));
    XQ_CUDA(c, cudaMemset(d, 0, sizeof(long long) * grid));
    if (nt == 128) {
        XQ_CUDA(c, cudaFuncSetAttribute(xq::umma_rate_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        xq::umma_rate_kernel<128><<<grid, 128, 120 * 1024>>>(d, n_mma, 0);
    } else {
        XQ_CUDA(c, cudaFuncSetAttribute(xq::umma_rate_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        xq::umma_rate_kernel<256><<<grid, 128, 120 * 1024>>>(d, n_mma, 0);
    }
    XQ_CUDA(c, cudaDeviceSynchronize());
    XQ_CUDA(c, cudaMemcpy(h_cycles, d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return XQ_OK;
}
