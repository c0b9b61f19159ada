// xq_net.cu -- K3: bf16 tcgen05 implicit-GEMM kernels for the policy-value ResNet forward.
//
// Replaces the library calls of XiangqiNet.forward (training/model.py:87-107; conv3x3+BN+ReLU
// stack, residual blocks :30-36, 1x1 heads, 2880->8100 policy FC, value MLP + tanh) on the
// inference path of self-play (model.py:109-124 predict, inference_server.py:251-263).
// BatchNorm is folded into the conv weights/bias on the host (eval mode, model.py:118).
//
// Data layout ("channel-chunk planes", no padding).  An activation tensor is stored as
//     X[chunk = C/8][row][8 channels]   (bf16, 16 bytes per (chunk,row))
// with board b, cell (r,c) at row b*90 + r*9 + c: the 90 cells of a board are consecutive rows and
// boards follow each other without any halo.  A 3x3 tap (dy,dx) of a 128-row output tile is the SAME
// matrix shifted by dy*9+dx rows, so
//   * the A operand of a tile pair for all 9 taps is one block of 256+2*10 rows per chunk, fetched
//     ONCE with 1-D bulk copies (cp.async.bulk, the TMA engine without a tensor map: the source is a
//     contiguous byte range) into the canonical no-swizzle K-major UMMA layout with SBO = 128 B,
//     i.e. address = base + row*16 + chunk*LBO -- linear in the row, so a tap only moves the
//     descriptor's start address by shift*16 bytes; no im2col, no re-load;
//   * a shifted row that falls off the board (row 0 looking up, column 8 looking right, ...) would
//     read a cell of the neighbouring board row / board.  Those (output row, tap) pairs are switched
//     off with the disable-output-lane mask of tcgen05.mma: bit i of the 128-bit mask keeps TMEM lane
//     i (= output row i of the tile) from being updated by that MMA.  The masks depend only on
//     (tile start row mod 90, tap), i.e. on tile_index mod 45: a 45 x 9 table of 16-byte masks built
//     once per context into constant memory (a warp-uniform read lands in the uniform registers tcgen05.mma wants).  The centre tap is unmasked and is
//     issued first (it initialises the accumulator).  Result: every executed MMA row is a real cell
//     (the first generation carried a zero halo: 110 rows per board, 22 % of the MMAs wasted);
//   * the epilogue stores 16 B per (chunk,row) with consecutive lanes on consecutive rows: fully
//     coalesced, and the next layer's bulk copies read exactly what was written.
// Weights are pre-arranged on the host as per-stage shared-memory images
// [n_tile][tap][k_block][chunk][n][8] (tap 0 = centre, then row-major) so a pipeline stage is
// contiguous.
//
// conv_kernel: persistent, one CTA per SM (it owns all 512 TMEM columns), 11 warps, warp-specialised:
// warp 0 lane 0 = bulk-copy producer, warps 1 and 10 = one tcgen05.mma-issuing thread per row tile
// (warp 1 also owns the TMEM allocation), warps 2-9 = epilogue (4 TMEM lane quarters x 2 row tiles).
// A 256-row tile pair shares every weight stage (two accumulators); accumulator pairs are double
// buffered in TMEM so the epilogue of pair i runs under the MMAs of pair i+1.  The A block is split
// into per-k-block segments with their own full/empty barriers (k-blocks are the outer loop); layers
// with a small A block (input conv, heads) keep several pairs' blocks in flight (NA) and their whole
// weight image resident.  What bounds the kernel: shared-memory bandwidth -- an M128 N128 K16 MMA reads
// 8 KB of operands in 64 tensor cycles = the SM's 128 B/clk, and the bulk copies write another
// ~2.5 KB per MMA (profiles/r2_net_ncu.md).
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
    const uint8_t* w_half;   // 3x3 layers: the weight image tiled by 64 output channels (conv2_kernel: each CTA of a pair holds half of N)
    const int* n_dev;        // optional DEVICE board count (<= n_boards): the launch sizes itself to it (leaf compaction)
};

// boards / 128-row tiles this launch really has: the host's figures, cut down to the device-side count if there is one
__device__ __forceinline__ void live_size(const GemmArgs& p, int rows_per_board, int* n_boards, int* m_tiles)
{
    *n_boards = p.n_boards;
    *m_tiles = p.m_tiles;
    if (p.n_dev) {
        const int nd = *p.n_dev;
        if (nd < p.n_boards) {
            *n_boards = nd < 0 ? 0 : nd;
            *m_tiles = (int)(((long long)*n_boards * rows_per_board + 127) / 128);
        }
    }
}

// per-context state of this translation unit
struct NetState {
    bool tapmask_done = false;
    cudaStream_t side = nullptr;          // the value MLP runs here, next to the policy FC (both depend on the heads conv only)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // small batches: the value MLP of a forward runs on its own stream next to the policy FC; [0] for forwards on the caller's
    // stream, [1] for forwards that themselves run on `side` (the arena's second network)
    cudaStream_t vfork[2] = {nullptr, nullptr};
    cudaEvent_t vev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    uint32_t attr_done = 0;   // bit per kernel instantiation: dynamic shared-memory limit raised on THIS context's device
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
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, issued by ONE thread.  Bit i of the 128-bit
// mask (mk.x = lanes 0-31 ...) disables the update of TMEM lane i (output row i of the M = 128 tile).
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                          const uint4 mk)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(mk.x), "r"(mk.y), "r"(mk.z), "r"(mk.w)
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// instruction descriptor kind::f16: D=f32 (bit 4), A=B=bf16 (bits 7,10), K-major A/B, N>>3 at 17, M>>4 at 24.
// Shared-memory matrix descriptors (no swizzle, K-major; cute/arch/mma_sm100_desc.hpp): [0,14) start>>4,
// [16,30) LBO>>4 (stride between the two 8-element K chunks of one MMA), [32,46) SBO>>4 (stride between 8-row
// groups = 128 B here), [46,48) version = 1; they are assembled inline from a 32-bit low word per operand.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

constexpr int kBoardRows = 90;                     // plane rows per board (no halo)
constexpr int kHalo = 10;                          // largest |dy*9+dx|
constexpr int kPairRows = 256;                     // two 128-row tiles share every weight stage
constexpr int kARows = kPairRows + 2 * kHalo;      // 276 rows of the resident A block
constexpr int kAPlane = kARows * 16;               // 4416 B per 8-channel chunk
constexpr int kMaskPhases = 45;                    // lcm(128, 90) / 128 tile phases
constexpr int kMaskBytes = kMaskPhases * 9 * 16;   // 6480

// tap index of the weight image -> (dy, dx): tap 0 is the centre, 1..8 the others in row-major order
__host__ __device__ constexpr int tap_cell(int k) { return k == 0 ? 4 : (k <= 4 ? k - 1 : k); }
__host__ __device__ constexpr int tap_dy(int k) { return tap_cell(k) / 3 - 1; }
__host__ __device__ constexpr int tap_dx(int k) { return tap_cell(k) % 3 - 1; }

// disable-output-lane masks, [tile phase][tap] (filled once per context's device by ensure_tapmask).  In constant
// memory on purpose: tcgen05.mma takes its operands from UNIFORM registers, and a mask read with a warp-uniform index
// from the constant bank lands there directly; from shared memory it took four register->uniform moves per MMA.
__constant__ uint4 c_tapmask[kMaskPhases * 9];

// one lane of the (converged) warp; the other lanes run the same uniform control flow next to it
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

constexpr int kConvThreads = 352;   // warp 0 producer, warps 1 and 10 MMA issuers (row tile 0 / 1), warps 2-9 epilogue
constexpr int kMaxStages = 9;
constexpr int kMaxSeg = 4;          // k-blocks of a layer (256 channels = 4 x 64)
constexpr int kMaxABufs = 16;       // NA * k-blocks

// NT output channels per tile, KCH 8-channel chunks per k-block, HEADS = 1x1 head conv (one tap), TPS taps per weight
// stage, WRES = the layer's whole weight image stays resident (one stage per (k-block, tap group), loaded once)
template <int NT, int KCH, bool HEADS, int TPS, bool WRES>
struct ConvCfg {
    static constexpr int kWTap = KCH * NT * 16;                // bytes of one (tap, k-block) weight slice
    static constexpr int kWStage = TPS * kWTap;
    static constexpr int kSeg = KCH * kAPlane;                 // bytes of one A segment (one k-block, 276 rows)
    static constexpr int kTileCols = NT > 64 ? 128 : 64;
    static constexpr int kTmemCols = 4 * kTileCols;
    static constexpr int kTaps = HEADS ? 1 : 9;
    static constexpr int kFixed = 768;                                       // barriers
    static constexpr int kBudget = 227 * 1024 - 1024 - 256;                 // per-CTA limit minus static bias minus slack
    // stages and A buffers for a layer with `kblocks` k-blocks: (stages, na), 0 stages = does not fit
    static void plan(int kblocks, int* stages, int* na)
    {
        const int a1 = kblocks * kSeg;
        int st, n;
        if (WRES) {
            st = kblocks * (kTaps / TPS);
            n = (kBudget - kFixed - st * kWStage) / a1;
            if (n > kMaxABufs / kblocks) n = kMaxABufs / kblocks;
            if (n > 4) n = 4;
            if (st > kMaxStages) st = 0;
        } else {
            n = 1;
            st = (kBudget - kFixed - a1) / kWStage;
            if (st > kMaxStages) st = kMaxStages;
        }
        *stages = (n < 1) ? 0 : st;
        *na = n;
    }
    static int smem_bytes(int kblocks, int stages, int na)
    {
        const int total = na * kblocks * kSeg + stages * kWStage + kFixed;
        return total < 120 * 1024 ? 120 * 1024 : total;        // > half an SM: one CTA per SM (it owns all of TMEM)
    }
};

template <int NT, int KCH, bool HEADS, int TPS, bool WRES, bool RES>
__global__ void __launch_bounds__(kConvThreads, 1) conv_kernel(const GemmArgs p, const int S, const int NA)
{
    using Cfg = ConvCfg<NT, KCH, HEADS, TPS, WRES>;
    constexpr int TS = Cfg::kTileCols;
    extern __shared__ __align__(128) uint8_t smem[];
    griddep_launch_dependents();
    const int kblocks = p.kchunks / KCH;
    uint8_t* sA = smem;                                         // [NA][kblocks][KCH][276][16 B]
    uint8_t* sStage = smem + NA * kblocks * Cfg::kSeg;          // [S][kWStage]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + S * Cfg::kWStage);
    uint64_t* w_full = bars;                                    // [kMaxStages]
    uint64_t* w_empty = bars + kMaxStages;
    uint64_t* a_full = bars + 2 * kMaxStages;                   // [kMaxABufs]
    uint64_t* a_empty = a_full + kMaxABufs;
    uint64_t* t_full = a_empty + kMaxABufs;                     // [2 accumulator stages][2 row tiles]
    uint64_t* t_empty = t_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 4);

    // warp index through a shuffle: the compiler then knows it is warp-uniform (role tests, TMEM addresses and the
    // MMA operands derived from it stay in uniform registers)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    constexpr int tap_groups = Cfg::kTaps / TPS;
    int live_boards, live_tiles;
    live_size(p, kBoardRows, &live_boards, &live_tiles);
    const int m_pairs = (live_tiles + 1) / 2;
    const int total = m_pairs * p.n_tiles;
    const long long real_rows = (long long)live_boards * kBoardRows;

    __shared__ __align__(16) float sBias[256];
    for (int i = threadIdx.x; i < p.n_tiles * NT && i < 256; i += kConvThreads) sBias[i] = p.bias[i];
    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 2);                 // both MMA issuers release a stage
        }
        for (int i = 0; i < NA * kblocks; ++i) {
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

    if (warp == 0) {
        // ===================== bulk-copy producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int n = 0;
            if (WRES) {
                // the whole weight image of this layer (n_tiles == 1): stage index = kb * tap_groups + tg
                for (int kb = 0; kb < kblocks; ++kb)
                    for (int tg = 0; tg < tap_groups; ++tg) {
                        const int st = kb * tap_groups + tg;
                        mbar_expect_tx(&w_full[st], Cfg::kWStage);
#pragma unroll
                        for (int tp = 0; tp < TPS; ++tp)
                            bulk_g2s(sStage + st * Cfg::kWStage + tp * Cfg::kWTap,
                                     p.w + (size_t)((tg * TPS + tp) * kblocks + kb) * Cfg::kWTap, Cfg::kWTap, &w_full[st]);
                    }
            }
            // streamed weights: the first item's first stages (all empty at kernel start) go out before the dependency wait too
            int pre = 0;
            if (!WRES && (int)blockIdx.x < total) {
                const uint8_t* wt0 = p.w + (size_t)((int)blockIdx.x % p.n_tiles) * Cfg::kTaps * kblocks * Cfg::kWTap;
                const int per_item = kblocks * tap_groups;
                pre = per_item < S ? per_item : S;
                for (int q = 0; q < pre; ++q) {
                    const int kb = q / tap_groups, tg = q - kb * tap_groups;
                    mbar_expect_tx(&w_full[q], Cfg::kWStage);
#pragma unroll
                    for (int tp = 0; tp < TPS; ++tp)
                        bulk_g2s(sStage + q * Cfg::kWStage + tp * Cfg::kWTap, wt0 + (size_t)((tg * TPS + tp) * kblocks + kb) * Cfg::kWTap,
                                 Cfg::kWTap, &w_full[q]);
                }
            }
            griddep_wait();                                     // the weights are on their way; now the previous layer's output
            for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
                const int pair = work / p.n_tiles, n_tile = work - pair * p.n_tiles;
                const long long m0 = (long long)pair * kPairRows;
                const uint8_t* wt = p.w + (size_t)n_tile * Cfg::kTaps * kblocks * Cfg::kWTap;
                const int ab = (n % NA) * kblocks;
                const uint32_t aph = (uint32_t)(n / NA) & 1u;
                int q = 0;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&a_empty[ab + kb], aph ^ 1u);
                    mbar_expect_tx(&a_full[ab + kb], (uint32_t)Cfg::kSeg);
#pragma unroll
                    for (int c = 0; c < KCH; ++c)
                        bulk_g2s(sA + (ab + kb) * Cfg::kSeg + c * kAPlane,
                                 p.a + ((size_t)(kb * KCH + c) * p.a_rows + (size_t)(p.a_row0 + m0 - kHalo)) * 16, kAPlane,
                                 &a_full[ab + kb]);
                    if (WRES) continue;
                    for (int tg = 0; tg < tap_groups; ++tg, ++q) {
                        if (n != 0 || q >= pre) {               // (the first item's first `pre` stages are already in flight)
                            mbar_wait(&w_empty[s], ph ^ 1);
                            mbar_expect_tx(&w_full[s], Cfg::kWStage);
                            // host image order is [tap][k_block]: one copy per tap of the stage
#pragma unroll
                            for (int tp = 0; tp < TPS; ++tp)
                                bulk_g2s(sStage + s * Cfg::kWStage + tp * Cfg::kWTap,
                                         wt + (size_t)((tg * TPS + tp) * kblocks + kb) * Cfg::kWTap, Cfg::kWTap, &w_full[s]);
                        }
                        if (++s == S) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1 || warp == 10) {
        // ===================== MMA issuers: one warp per row tile =====================
        // The whole warp runs the (uniform) loop and ONE elected lane issues the MMAs and commits: every MMA operand
        // is then a warp-uniform value the compiler keeps in uniform registers.  (With the loop inside `if (lane == 0)`
        // ptxas wrapped every tcgen05.mma in an elect / 9 x R2UR.BROADCAST / branch sequence, ~15 instructions per
        // MMA on the critical issuing thread.)  Two issuers interleave their MMAs and hide each other's stage hand-offs
        // (~0.2 us each: wait, fence, commit).
        const int t = warp == 1 ? 0 : 1;
        const bool leader = elect_one();
        {
            constexpr uint32_t idesc = make_idesc(128, NT);
            constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);            // SBO = 128 B, version 1
            constexpr uint32_t kLboA = (uint32_t)(kAPlane >> 4);
            int s = 0;
            uint32_t ph = 0;
            int n = 0;
            for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
                const int pair = work / p.n_tiles;
                const int acc = n & 1;
                const uint32_t tph = (uint32_t)(n >> 1) & 1u;
                const int ab = (n % NA) * kblocks;
                const uint32_t aph = (uint32_t)(n / NA) & 1u;
                const int mphase = ((pair * 2 + t) % kMaskPhases) * 9;
                mbar_wait(&t_empty[acc * 2 + t], tph ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(acc * 2 * TS + t * TS);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&a_full[ab + kb], aph);
                    const uint32_t a_seg = (((smem_u32(sA + (ab + kb) * Cfg::kSeg)) >> 4) & 0x3FFFu) | (kLboA << 16);
                    for (int tg = 0; tg < tap_groups; ++tg) {
                        const int st = WRES ? kb * tap_groups + tg : s;
                        mbar_wait(&w_full[st], WRES ? 0u : ph);
                        tc_fence_after();
                        const uint32_t b_st = ((smem_u32(sStage + st * Cfg::kWStage) >> 4) & 0x3FFFu) | ((uint32_t)((NT * 16) >> 4) << 16);
#pragma unroll
                        for (int tp = 0; tp < TPS; ++tp) {
                            const int tap = tg * TPS + tp;
                            const int shift = HEADS ? 0 : tap_dy(tap) * 9 + tap_dx(tap);
                            const uint4 mk = HEADS ? make_uint4(0, 0, 0, 0) : c_tapmask[mphase + tap];
                            const uint32_t a_lo = a_seg + (uint32_t)(kHalo + shift);
                            const uint32_t b_lo = b_st + (uint32_t)(tp * (Cfg::kWTap >> 4));
                            const uint32_t first = (kb | tap) != 0 ? 1u : 0u;
#pragma unroll
                            for (int j = 0; j < KCH / 2; ++j) {
                                const uint64_t adesc = ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + (uint32_t)(2 * j) * kLboA + (uint32_t)(t * 128));
                                const uint64_t bdesc = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + (uint32_t)(2 * j * NT));
                                if (leader) umma_bf16(d_addr, adesc, bdesc, idesc, j == 0 ? first : 1u, mk);
                            }
                        }
                        if (!WRES) {
                            if (leader) umma_commit(&w_empty[s]);
                            if (++s == S) { s = 0; ph ^= 1; }
                        }
                        __syncwarp();
                    }
                    if (leader) umma_commit(&a_empty[ab + kb]);  // this k-block's rows may be replaced by a later pair's
                }
                if (leader) umma_commit(&t_full[acc * 2 + t]);
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: 8 warps, warp -> (row tile t, TMEM lane quarter q) =====================
        const int q = warp & 3;                          // a warp may only read TMEM lanes 32*(warp%4) ..
        const int t = (warp - 2) >> 2;                   // warps 2-5: tile 0, warps 6-9: tile 1
        const int row = q * 32 + lane;
        int n = 0;
        uint4 res[NT / 8];
        griddep_wait();                                  // residual reads and output stores come after the previous kernel
        for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
            const int pair = work / p.n_tiles, n_tile = work - pair * p.n_tiles;
            const int acc = n & 1;
            const uint32_t tph = (uint32_t)(n >> 1) & 1u;
            const long long mrow = (long long)pair * kPairRows + t * 128 + row;
            const bool real = mrow < real_rows;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 2 * TS + t * TS);

            if (!HEADS) {
                // Residual operand: res[k] always holds the values of this warp's NEXT tile; loaded before the
                // first wait, then refilled slab by slab for the next work item, i.e. a whole pair ahead.
                constexpr bool has_res = RES;              // compiled out for layers without a residual operand
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
                const bool nreal = nwork < total && nrow < real_rows;
                mbar_wait(&t_full[acc * 2 + t], tph);
                tc_fence_after();
                // one 32-column slab: bias (+residual) (+ReLU), bf16, 4 coalesced 16-byte stores
                auto emit = [&](const uint32_t* v, const int c0) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int k = c0 / 8 + g;
                        const int nn = n_tile * NT + c0 + g * 8;
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
                            if (RES) {
                                const uint32_t rw[4] = {res[k].x, res[k].y, res[k].z, res[k].w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    f[2 * e] += __uint_as_float(rw[e] << 16);
                                    f[2 * e + 1] += __uint_as_float(rw[e] & 0xffff0000u);
                                }
                            }
                            if (p.relu) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.0f);
                            }
                            const size_t off = ((size_t)(nn >> 3) * p.out_rows + (size_t)(p.out_row0 + mrow)) * 16;
                            *reinterpret_cast<uint4*>(p.out + off) =
                                make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                        }
                        if (RES) {
                            res[k] = make_uint4(0, 0, 0, 0);
                            if (nreal)
                                res[k] = __ldg(reinterpret_cast<const uint4*>(
                                    p.residual + ((size_t)(nn_tile * (NT / 8) + k) * p.out_rows + (size_t)(p.out_row0 + nrow)) * 16));
                        }
                    }
                };
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
            } else {
                mbar_wait(&t_full[acc * 2 + t], tph);
                tc_fence_after();
                const long long b = mrow / kBoardRows;
                const int pos = (int)(mrow - b * kBoardRows);
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

// =============================================================================================
// conv2_kernel: the 3x3 tower conv on a CTA PAIR with cta_group::2 MMAs (M = 256 across two SMs)
// =============================================================================================
// conv_kernel is bound by shared-memory bandwidth: every M128 N128 K16 MMA reads 4 KB of A and 4 KB of B against the
// SM's 128 B/clk, and the bulk copies write another ~2.5 KB per MMA into the same banks (87 cycles per MMA measured, 64
// ideal).  With cta_group::2 the two SMs of a cluster run ONE M256 N128 K16 MMA: each CTA supplies its own 128 rows of A
// and only HALF of B (64 of the 128 output channels: its own half of every weight stage); the tensor cores exchange the
// halves.  Operand reads per SM drop to 6 KB and weight-stage writes to half, one instruction covers what took two, and the
// freed shared memory holds six 24 KB stages instead of three 48 KB ones.
//   * cluster (2,1,1); a work item = two consecutive tile pairs (512 rows), CTA r takes pair 2*item + r; rank 0 issues;
//   * both CTAs run their own producer (own A segments, own half of every weight stage, image tiled by 64 channels);
//     the follower's warp 1 relays "my stage / my segment has landed" to the leader with remote mbarrier arrives;
//   * the leader's commits are multicast to both CTAs (stage empty, segment empty, accumulator full); both epilogues
//     arrive on the leader's accumulator-empty barrier;
//   * the disable-output-lane mask of a cta_group::2 MMA has 256 bits: lanes 0-127 = the leader's tile, 128-255 = the
//     follower's tile (each with its own tile phase).
constexpr int kMaxStages2 = 8;

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
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                           const uint4 m0, const uint4 m1)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(m0.x), "r"(m0.y), "r"(m0.z), "r"(m0.w), "r"(m1.x),
          "r"(m1.y), "r"(m1.z), "r"(m1.w)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}

template <int KCH, int TPS>
struct Conv2Cfg {
    static constexpr int NT = 128;                             // output channels per MMA (64 from each CTA's stage)
    static constexpr int kWTap = KCH * 64 * 16;                // this CTA's half of one (tap, k-block) weight slice
    static constexpr int kWStage = TPS * kWTap;
    static constexpr int kSeg = KCH * kAPlane;
    static constexpr int kTmemCols = 512;
    static constexpr int kFixed = 1024;
    static constexpr int kBudget = 227 * 1024 - 1024 - 256;
    static int stages(int kblocks)
    {
        int st = (kBudget - kFixed - kblocks * kSeg) / kWStage;
        return st > kMaxStages2 ? kMaxStages2 : st;
    }
    static int smem_bytes(int kblocks, int st)
    {
        const int total = kblocks * kSeg + st * kWStage + kFixed;
        return total < 120 * 1024 ? 120 * 1024 : total;
    }
};

template <int KCH, int TPS>
__global__ void __launch_bounds__(kConvThreads, 1) conv2_kernel(const GemmArgs p, const int S)
{
    using Cfg = Conv2Cfg<KCH, TPS>;
    constexpr int NT = 128, TS = 128;
    constexpr int tap_groups = 9 / TPS;
    extern __shared__ __align__(128) uint8_t smem[];
    const int kblocks = p.kchunks / KCH;
    uint8_t* sA = smem;
    uint8_t* sStage = smem + kblocks * Cfg::kSeg;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + S * Cfg::kWStage);
    uint64_t* w_full = bars;                                    // own half of the stage has landed
    uint64_t* w_empty = bars + kMaxStages2;                     // the MMAs reading the stage are done (leader multicast)
    uint64_t* w_peer = bars + 2 * kMaxStages2;                  // leader only: the follower's half has landed
    uint64_t* a_full = bars + 3 * kMaxStages2;                  // [kMaxSeg]
    uint64_t* a_empty = a_full + kMaxSeg;
    uint64_t* a_peer = a_empty + kMaxSeg;
    uint64_t* t_full = a_peer + kMaxSeg;                        // [2]
    uint64_t* t_empty = t_full + 2;                             // [2] leader only: the 16 epilogue warps of the pair of CTAs
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    griddep_launch_dependents();                                // the next layer may be scheduled as soon as every CTA of this one runs
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const bool leader_cta = crank == 0;
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    int live_boards, live_tiles;
    live_size(p, kBoardRows, &live_boards, &live_tiles);
    const int m_pairs = (live_tiles + 1) / 2;
    const int items = (m_pairs + 1) / 2;
    const int total = items * p.n_tiles;
    const long long real_rows = (long long)live_boards * kBoardRows;

    __shared__ __align__(16) float sBias[256];
    for (int i = threadIdx.x; i < p.n_tiles * NT && i < 256; i += kConvThreads) sBias[i] = p.bias[i];
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
        // ===================== producer (both CTAs): own rows, own half of the weights =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            int n = 0;
            // Weights do not depend on the previous layer: the first item's weight stages (all stages are empty at kernel
            // start) go out BEFORE the dependency wait, in the order the main loop would issue them.
            int pre = 0;
            if (cluster_id < total) {
                const int n_tile0 = cluster_id % p.n_tiles;
                const uint8_t* wt0 = p.w_half + (size_t)(2 * n_tile0 + (int)crank) * 9 * kblocks * Cfg::kWTap;
                const int per_item = kblocks * tap_groups;
                pre = per_item < S ? per_item : S;
                for (int q = 0; q < pre; ++q) {
                    const int kb = q / tap_groups, tg = q - kb * tap_groups;
                    mbar_expect_tx(&w_full[q], Cfg::kWStage);
#pragma unroll
                    for (int tp = 0; tp < TPS; ++tp)
                        bulk_g2s(sStage + q * Cfg::kWStage + tp * Cfg::kWTap, wt0 + (size_t)((tg * TPS + tp) * kblocks + kb) * Cfg::kWTap,
                                 Cfg::kWTap, &w_full[q]);
                }
            }
            griddep_wait();                                     // from here on the previous layer's output may be read
            for (int work = cluster_id; work < total; work += n_clusters, ++n) {
                const int item = work / p.n_tiles, n_tile = work - item * p.n_tiles;
                const int pair = item * 2 + (int)crank;
                const long long m0 = (long long)pair * kPairRows;
                // image tiled by 64 channels: [n_tile64][tap][k_block][chunk][64][8], n_tile64 = 2 * n_tile + rank
                const uint8_t* wt = p.w_half + (size_t)(2 * n_tile + (int)crank) * 9 * kblocks * Cfg::kWTap;
                int q = 0;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&a_empty[kb], (uint32_t)(n & 1) ^ 1u);
                    mbar_expect_tx(&a_full[kb], (uint32_t)Cfg::kSeg);
#pragma unroll
                    for (int c = 0; c < KCH; ++c)
                        bulk_g2s(sA + kb * Cfg::kSeg + c * kAPlane,
                                 p.a + ((size_t)(kb * KCH + c) * p.a_rows + (size_t)(p.a_row0 + m0 - kHalo)) * 16, kAPlane, &a_full[kb]);
                    for (int tg = 0; tg < tap_groups; ++tg, ++q) {
                        if (n != 0 || q >= pre) {               // (the first item's first `pre` stages are already in flight)
                            mbar_wait(&w_empty[s], ph ^ 1);
                            mbar_expect_tx(&w_full[s], Cfg::kWStage);
#pragma unroll
                            for (int tp = 0; tp < TPS; ++tp)
                                bulk_g2s(sStage + s * Cfg::kWStage + tp * Cfg::kWTap,
                                         wt + (size_t)((tg * TPS + tp) * kblocks + kb) * Cfg::kWTap, Cfg::kWTap, &w_full[s]);
                        }
                        if (++s == S) { s = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        int s = 0;
        uint32_t ph = 0;
        int n = 0;
        if (!leader_cta) {
            // ===================== follower: relay "landed" events to the leader =====================
            if (lane == 0) {
                for (int work = cluster_id; work < total; work += n_clusters, ++n) {
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(&a_full[kb], (uint32_t)(n & 1));
                        mbar_arrive_remote(&a_peer[kb], 0);
                        for (int tg = 0; tg < tap_groups; ++tg) {
                            mbar_wait(&w_full[s], ph);
                            mbar_arrive_remote(&w_peer[s], 0);
                            if (++s == S) { s = 0; ph ^= 1; }
                        }
                    }
                }
            }
        } else {
            // ===================== leader: one elected lane issues the MMAs of both SMs (uniform loop) =====================
            const bool leader = elect_one();
            constexpr uint32_t idesc = make_idesc(256, NT);
            constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
            constexpr uint32_t kLboA = (uint32_t)(kAPlane >> 4);
            for (int work = cluster_id; work < total; work += n_clusters, ++n) {
                const int item = work / p.n_tiles;
                const int acc = n & 1;
                const uint32_t tph = (uint32_t)(n >> 1) & 1u;
                // tile phases of the four tiles of the item: CTA r, row tile t -> tile index (item*2 + r)*2 + t
                const int ph00 = ((item * 4 + 0) % kMaskPhases) * 9, ph01 = ((item * 4 + 1) % kMaskPhases) * 9;
                const int ph10 = ((item * 4 + 2) % kMaskPhases) * 9, ph11 = ((item * 4 + 3) % kMaskPhases) * 9;
                mbar_wait(&t_empty[acc], tph ^ 1);
                tc_fence_after();
                const uint32_t d_addr = tmem_base + (uint32_t)(acc * 2 * TS);
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&a_full[kb], (uint32_t)(n & 1));
                    mbar_wait(&a_peer[kb], (uint32_t)(n & 1));
                    const uint32_t a_seg = (((smem_u32(sA + kb * Cfg::kSeg)) >> 4) & 0x3FFFu) | (kLboA << 16);
                    for (int tg = 0; tg < tap_groups; ++tg) {
                        mbar_wait(&w_full[s], ph);
                        mbar_wait(&w_peer[s], ph);
                        tc_fence_after();
                        const uint32_t b_st = ((smem_u32(sStage + s * Cfg::kWStage) >> 4) & 0x3FFFu) | ((uint32_t)((64 * 16) >> 4) << 16);
#pragma unroll
                        for (int tp = 0; tp < TPS; ++tp) {
                            const int tap = tg * TPS + tp;
                            const int shift = tap_dy(tap) * 9 + tap_dx(tap);
                            const uint32_t a_lo = a_seg + (uint32_t)(kHalo + shift);
                            const uint32_t b_lo = b_st + (uint32_t)(tp * (Cfg::kWTap >> 4));
                            const uint32_t first = (kb | tap) != 0 ? 1u : 0u;
#pragma unroll
                            for (int t = 0; t < 2; ++t) {
                                const uint4 m0 = c_tapmask[(t ? ph01 : ph00) + tap], m1 = c_tapmask[(t ? ph11 : ph10) + tap];
#pragma unroll
                                for (int j = 0; j < KCH / 2; ++j) {
                                    const uint64_t adesc = ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + (uint32_t)(2 * j) * kLboA + (uint32_t)(t * 128));
                                    const uint64_t bdesc = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + (uint32_t)(2 * j * 64));
                                    if (leader) umma2_bf16(d_addr + (uint32_t)(t * TS), adesc, bdesc, idesc, j == 0 ? first : 1u, m0, m1);
                                }
                            }
                        }
                        if (leader) umma2_commit_mc(&w_empty[s], 3);
                        if (++s == S) { s = 0; ph ^= 1; }
                        __syncwarp();
                    }
                    if (leader) umma2_commit_mc(&a_empty[kb], 3);
                }
                if (leader) umma2_commit_mc(&t_full[acc], 3);
                __syncwarp();
            }
        }
    } else if (warp < 10) {
        // ===================== epilogue (both CTAs): as conv_kernel, accumulator release goes to the leader =====================
        const int q = warp & 3;
        const int t = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        int n = 0;
        uint4 res[NT / 8];
        griddep_wait();                                         // the residual operand is the previous layers' output
        for (int work = cluster_id; work < total; work += n_clusters, ++n) {
            const int item = work / p.n_tiles, n_tile = work - item * p.n_tiles;
            const int pair = item * 2 + (int)crank;
            const int acc = n & 1;
            const uint32_t tph = (uint32_t)(n >> 1) & 1u;
            const long long mrow = (long long)pair * kPairRows + t * 128 + row;
            const bool real = mrow < real_rows;
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
            const bool nreal = nwork < total && nrow < real_rows;
            mbar_wait(&t_full[acc], tph);
            tc_fence_after();
            auto emit = [&](const uint32_t* v, const int c0) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int k = c0 / 8 + g;
                    const int nn = n_tile * NT + c0 + g * 8;
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
                        const size_t off = ((size_t)(nn >> 3) * p.out_rows + (size_t)(p.out_row0 + mrow)) * 16;
                        *reinterpret_cast<uint4*>(p.out + off) =
                            make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
                    }
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
                if (leader_cta) mbar_arrive(&t_empty[acc]);
                else mbar_arrive_remote(&t_empty[acc], 0);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                // both CTAs are done with TMEM and with each other's barriers
    if (warp == 1) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
}

static NetState* net_state(xq_ctx* c)
{
    if (!c->net) c->net = new NetState();
    return reinterpret_cast<NetState*>(c->net);
}

// dynamic shared-memory opt-in: a per-DEVICE attribute, so it is tracked per context (one bit per instantiation)
template <class K>
static int ensure_smem_attr(xq_ctx* c, K kern, int bit)
{
    NetState* N = net_state(c);
    if (!(N->attr_done & (1u << bit))) {
        cudaFuncAttributes fa;
        XQ_CUDA(c, cudaFuncGetAttributes(&fa, kern));
        // 227 KB per CTA in all: what the kernel's static __shared__ arrays do not take is opened for dynamic use
        XQ_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes));
        N->attr_done |= 1u << bit;
    }
    return XQ_OK;
}

template <class... KArgs, class... Args>
static cudaError_t launch_pdl(xq_ctx* c, void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args... args)
{
    return xq_launch_pdl(c->net_pdl, kern, dim3(grid), block, smem, s, args...);
}

static int ensure_tapmask(xq_ctx* c, cudaStream_t s)
{
    NetState* N = net_state(c);
    if (N->tapmask_done) return XQ_OK;
    static uint32_t h[kMaskPhases * 9 * 4];
    for (int ph = 0; ph < kMaskPhases; ++ph)
        for (int k = 0; k < 9; ++k) {
            uint32_t* w = &h[(ph * 9 + k) * 4];
            w[0] = w[1] = w[2] = w[3] = 0;
            for (int i = 0; i < 128; ++i) {
                const int pos = (ph * 128 + i) % kBoardRows, r = pos / 9 + tap_dy(k), cc = pos % 9 + tap_dx(k);
                if (r < 0 || r > 9 || cc < 0 || cc > 8) w[i >> 5] |= 1u << (i & 31);   // source cell is off the board
            }
        }
    // the __constant__ symbol has one instance per device: this writes the one of the context's (current) device
    XQ_CUDA(c, cudaMemcpyToSymbolAsync(c_tapmask, h, sizeof(h), 0, cudaMemcpyHostToDevice, s));
    XQ_CUDA(c, cudaStreamSynchronize(s));
    N->tapmask_done = true;
    return XQ_OK;
}

template <int NT, int KCH, bool HEADS, int TPS, bool WRES, bool RES>
static int launch_conv_r(xq_ctx* c, GemmArgs& a, cudaStream_t s, int bit)
{
    using Cfg = ConvCfg<NT, KCH, HEADS, TPS, WRES>;
    const int kblocks = a.kchunks / KCH;
    int S = 0, NA = 0;
    Cfg::plan(kblocks, &S, &NA);
    if (kblocks > kMaxSeg || S < (WRES ? 1 : 2) || (WRES && a.n_tiles != 1))
        return xq_fail(c, XQ_ERR_ARG, "conv kernel: %d k-blocks, %d stages, %d n-tiles do not fit", kblocks, S, a.n_tiles);
    auto kern = conv_kernel<NT, KCH, HEADS, TPS, WRES, RES>;
    if (int rc = ensure_smem_attr(c, kern, bit)) return rc;
    if (!HEADS)
        if (int rc = ensure_tapmask(c, s)) return rc;
    const int total = ((a.m_tiles + 1) / 2) * a.n_tiles;
    const int grid = c->sm_count < total ? c->sm_count : total;   // persistent, one CTA per SM
    XQ_CUDA(c, launch_pdl(c, kern, grid, kConvThreads, (size_t)Cfg::smem_bytes(kblocks, S, NA), s, (const GemmArgs)a, (const int)S, (const int)NA));
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

// the residual operand is a compile-time property of the epilogue (a layer without one runs a third fewer instructions there)
template <int NT, int KCH, bool HEADS, int TPS, bool WRES>
static int launch_conv(xq_ctx* c, GemmArgs& a, cudaStream_t s, int bit)
{
    if (!HEADS && a.residual) return launch_conv_r<NT, KCH, HEADS, TPS, WRES, true>(c, a, s, bit + 16);
    return launch_conv_r<NT, KCH, HEADS, TPS, WRES, false>(c, a, s, bit);
}

template <int KCH, int TPS>
static int launch_conv2(xq_ctx* c, GemmArgs& a, cudaStream_t s, int bit)
{
    using Cfg = Conv2Cfg<KCH, TPS>;
    const int kblocks = a.kchunks / KCH;
    const int S = Cfg::stages(kblocks);
    if (kblocks > kMaxSeg || S < 2 || !a.w_half) return xq_fail(c, XQ_ERR_ARG, "conv2 kernel: %d k-blocks, %d stages, w_half %p", kblocks, S, (const void*)a.w_half);
    auto kern = conv2_kernel<KCH, TPS>;
    if (int rc = ensure_smem_attr(c, kern, bit)) return rc;
    if (int rc = ensure_tapmask(c, s)) return rc;
    const int items = (((a.m_tiles + 1) / 2) + 1) / 2;
    const int total = items * a.n_tiles;
    int clusters = c->sm_count / 2;
    if (clusters > total) clusters = total;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kConvThreads);
    cfg.dynamicSmemBytes = Cfg::smem_bytes(kblocks, S);
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    // programmatic dependent launch: this layer may start while the previous kernel of the stream is draining (its CTAs call
    // griddepcontrol.launch_dependents first thing) and waits for it with griddepcontrol.wait before touching its output
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = c->net_pdl ? 2 : 1;
    XQ_CUDA(c, cudaLaunchKernelEx(&cfg, kern, (const GemmArgs)a, S));
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

// =============================================================================================
// fc_kernel: dense layer (policy FC 2880 -> 8100) in the conv kernel's style
// =============================================================================================
// Both operands stream, so what bounds the layer is shared-memory bandwidth (128 B/clk per SM): the operand reads
// of the MMAs plus the bulk copies' writes.  Work item = 256 boards x NT = 224 outputs: two M = 128 tiles share
// every stage, each with ONE N = 224 MMA per 16 input features (A 4 KB + B 7 KB read per MMA, 15 KB of stage
// written per K step: 165 B/clk at full tensor rate, against 219 B/clk for the first generation's 256 x 128 items
// of N = 128 MMAs).  37 x 224 = 8288 >= 8100 columns x 16 board pairs = 592 items = exactly 4 per SM at batch 4096.
// The two 224-column accumulators fill TMEM (no double buffering): the epilogue of an item (8 warps) is exposed,
// ~2 us of ~28, while the producer keeps prefetching the next item's stages.
template <int NT_, int KCH_, int STAGES_>
struct FcCfgT {
    static constexpr int kNT = NT_;
    static constexpr int kKch = KCH_;                          // 8 input features per chunk
    static constexpr int kWBytes = kKch * NT_ * 16;
    static constexpr int kABytes = kKch * 256 * 16;
    static constexpr int kStage = kWBytes + kABytes;
    static constexpr int kStages = STAGES_;
    static constexpr int kSmem = kStages * kStage + 512;
};

template <class FcCfg>
__global__ void __launch_bounds__(kConvThreads, 1) fc_kernel(const GemmArgs p)
{
    constexpr int S = FcCfg::kStages, NT = FcCfg::kNT;
    extern __shared__ __align__(128) uint8_t smem[];
    griddep_launch_dependents();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * FcCfg::kStage);
    uint64_t* w_full = bars;
    uint64_t* w_empty = bars + S;
    uint64_t* t_full = bars + 2 * S;          // [2 tiles]
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    __shared__ __align__(16) float sBias[2][256];

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform
    const int iters = p.kchunks / FcCfg::kKch;
    int live_boards, live_tiles;
    live_size(p, 1, &live_boards, &live_tiles);
    const int m_pairs = (live_tiles + 1) / 2;
    const int total = m_pairs * p.n_tiles;
    const uint4 nomask = make_uint4(0, 0, 0, 0);

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 2);
        }
        for (int i = 0; i < 2; ++i) {
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
            // the weight halves of the first item's first stages do not depend on the previous kernel: they go out before the wait
            int pre = 0;
            if ((int)blockIdx.x < total) {
                const uint8_t* wt0 = p.w + (size_t)((int)blockIdx.x % p.n_tiles) * iters * FcCfg::kWBytes;
                pre = iters < S ? iters : S;
                long long lr0 = (long long)live_boards - (long long)((int)blockIdx.x / p.n_tiles) * 256;
                lr0 = lr0 > 256 ? 256 : lr0;
                const uint32_t tx0 = (uint32_t)FcCfg::kWBytes + (uint32_t)FcCfg::kKch * ((uint32_t)((lr0 + 7) / 8 * 8) * 16u);
                for (int q = 0; q < pre; ++q) {
                    mbar_expect_tx(&w_full[q], tx0);
                    bulk_g2s(smem + q * FcCfg::kStage, wt0 + (size_t)q * FcCfg::kWBytes, FcCfg::kWBytes, &w_full[q]);
                }
            }
            griddep_wait();
            bool first = true;
            for (int work = blockIdx.x; work < total; work += gridDim.x, first = false) {
                const int pair = work / p.n_tiles, n_tile = work - pair * p.n_tiles;
                const long long m0 = (long long)pair * 256;
                const uint8_t* wt = p.w + (size_t)n_tile * iters * FcCfg::kWBytes;
                // only the live boards' rows of the A tile travel (a small batch is weight-bound: 46 MB of weights against a few
                // KB of features); the rows behind them keep stale shared memory and their outputs are dropped by the epilogue
                long long live_rows = (long long)live_boards - m0;
                live_rows = live_rows > 256 ? 256 : live_rows;
                const uint32_t a_bytes = (uint32_t)((live_rows + 7) / 8 * 8) * 16u;
                const uint32_t stage_tx = (uint32_t)FcCfg::kWBytes + (uint32_t)FcCfg::kKch * a_bytes;
                for (int it = 0; it < iters; ++it) {
                    uint8_t* st = smem + s * FcCfg::kStage;
                    if (!first || it >= pre) {
                        mbar_wait(&w_empty[s], ph ^ 1);
                        mbar_expect_tx(&w_full[s], stage_tx);
                        bulk_g2s(st, wt + (size_t)it * FcCfg::kWBytes, FcCfg::kWBytes, &w_full[s]);
                    }
#pragma unroll
                    for (int c = 0; c < FcCfg::kKch; ++c)
                        bulk_g2s(st + FcCfg::kWBytes + c * 4096,
                                 p.a + ((size_t)(it * FcCfg::kKch + c) * p.a_rows + (size_t)(p.a_row0 + m0)) * 16, a_bytes, &w_full[s]);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1 || warp == 10) {
        // the whole warp runs the loop, one elected lane issues (see conv_kernel): operands stay in uniform registers
        const int t = warp == 1 ? 0 : 1;
        const bool leader = elect_one();
        {
            constexpr uint32_t idesc = make_idesc(128, NT);
            constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);
            int s = 0;
            uint32_t ph = 0;
            int n = 0;
            const uint32_t d_addr = tmem_base + (uint32_t)(t * 256);
            for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
                mbar_wait(&t_empty[t], (uint32_t)(n & 1) ^ 1u);
                tc_fence_after();
                for (int it = 0; it < iters; ++it) {
                    mbar_wait(&w_full[s], ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * FcCfg::kStage);
                    const uint32_t b_lo = ((st >> 4) & 0x3FFFu) | ((uint32_t)((NT * 16) >> 4) << 16);
                    const uint32_t a_lo = (((st + FcCfg::kWBytes + (uint32_t)t * 2048u) >> 4) & 0x3FFFu) | ((4096u >> 4) << 16);
#pragma unroll
                    for (int j = 0; j < FcCfg::kKch / 2; ++j) {
                        const uint64_t adesc = ((uint64_t)kDescHi << 32) | (uint64_t)(a_lo + (uint32_t)(2 * j) * (4096u >> 4));
                        const uint64_t bdesc = ((uint64_t)kDescHi << 32) | (uint64_t)(b_lo + (uint32_t)(2 * j * NT));
                        if (leader) umma_bf16(d_addr, adesc, bdesc, idesc, (it | j) != 0 ? 1u : 0u, nomask);
                    }
                    if (leader) umma_commit(&w_empty[s]);
                    if (++s == S) { s = 0; ph ^= 1; }
                    __syncwarp();
                }
                if (leader) umma_commit(&t_full[t]);
                __syncwarp();
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
            const int bb = n & 1;
            const long long m = (long long)pair * 256 + t * 128 + row;
            const bool real = m < (long long)live_boards;
            // this work item's bias values -> shared memory (double buffered by item parity)
            if (et < NT) sBias[bb][et] = p.bias[n_tile * NT + et];
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * 256);
            mbar_wait(&t_full[t], (uint32_t)(n & 1));
            tc_fence_after();
            auto emit = [&](const uint32_t* v, const int c0) {
                if (!real) return;
                uint4* dst = reinterpret_cast<uint4*>(p.out + ((size_t)m * p.out_stride + (size_t)(n_tile * NT + c0)) * 2);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const float4 b0 = *reinterpret_cast<const float4*>(&sBias[bb][c0 + g * 8]);
                    const float4 b1 = *reinterpret_cast<const float4*>(&sBias[bb][c0 + g * 8 + 4]);
                    dst[g] = make_uint4(pack_bf16(__uint_as_float(v[g * 8 + 0]) + b0.x, __uint_as_float(v[g * 8 + 1]) + b0.y),
                                        pack_bf16(__uint_as_float(v[g * 8 + 2]) + b0.z, __uint_as_float(v[g * 8 + 3]) + b0.w),
                                        pack_bf16(__uint_as_float(v[g * 8 + 4]) + b1.x, __uint_as_float(v[g * 8 + 5]) + b1.y),
                                        pack_bf16(__uint_as_float(v[g * 8 + 6]) + b1.z, __uint_as_float(v[g * 8 + 7]) + b1.w));
                }
            };
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
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[t]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <class FcCfg>
static int launch_fc(xq_ctx* c, const GemmArgs& a, cudaStream_t s, int bit)
{
    if (int rc = ensure_smem_attr(c, fc_kernel<FcCfg>, bit)) return rc;
    if (a.kchunks % FcCfg::kKch) return xq_fail(c, XQ_ERR_ARG, "fc: K/8 = %d is not a multiple of %d", a.kchunks, FcCfg::kKch);
    if ((long long)a.n_tiles * FcCfg::kNT > a.out_stride) return xq_fail(c, XQ_ERR_ARG, "fc: %d tiles of %d columns exceed the output stride %lld", a.n_tiles, FcCfg::kNT, a.out_stride);
    const int total = ((a.m_tiles + 1) / 2) * a.n_tiles;
    const int grid = c->sm_count < total ? c->sm_count : total;
    XQ_CUDA(c, launch_pdl(c, fc_kernel<FcCfg>, grid, kConvThreads, (size_t)FcCfg::kSmem, s, (const GemmArgs)a));
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

// ---- value head: Linear(360,128)+ReLU -> Linear(128,1) -> tanh (model.py:74-83) ----------------
// feats [B][90][4] fp32 (already conv1x1+BN+ReLU) = [B][360] with k = pos*4+ch, w1t [360][128] fp32.
// 32 boards per CTA, 4 warps; a thread owns 4 hidden units x 8 boards (32 accumulators): per k one 16-byte
// read of W1^T (conflict-free) and two broadcast 16-byte reads of the features feed 32 FMAs.  The next 40-row chunk of
// W1^T and of the features travels global -> registers while the current one is multiplied (the first version loaded,
// synchronised and multiplied in turn and spent three quarters of its 38 us waiting for those loads).
constexpr int kVhBoards = 32;
constexpr int kVhChunk = 40;    // k rows staged per step
__global__ void __launch_bounds__(128) value_head_kernel(const float* __restrict__ feats, const float* __restrict__ w1t,
                                                          const float* __restrict__ b1, const float* __restrict__ w2,
                                                          float b2, float* __restrict__ value, int B, const int* __restrict__ n_dev)
{
    griddep_launch_dependents();
    griddep_wait();                                              // feats is the heads conv's output
    if (n_dev) B = min(B, *n_dev);
    if ((int)(blockIdx.x * kVhBoards) >= B) return;
    __shared__ __align__(16) float wsm[kVhChunk][128];
    __shared__ __align__(16) float fsm[kVhChunk][kVhBoards];     // transposed: [k][board]
    const int b0 = blockIdx.x * kVhBoards;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWPer = kVhChunk * 32 / 128;                   // 10 float4 of W1^T per thread and chunk
    constexpr int kFPer = (kVhBoards * (kVhChunk / 4) + 127) / 128;   // 3 (the last one partly used)
    float4 wreg[kWPer], freg[kFPer];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < kWPer; ++u)
            wreg[u] = __ldg(reinterpret_cast<const float4*>(w1t + (size_t)k0 * 128) + threadIdx.x + u * 128);
#pragma unroll
        for (int u = 0; u < kFPer; ++u) {
            const int i = threadIdx.x + u * 128, bb = i & 31, kq = i >> 5;     // consecutive lanes = consecutive boards
            freg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (kq < kVhChunk / 4 && b0 + bb < B)
                freg[u] = __ldg(reinterpret_cast<const float4*>(feats + (size_t)(b0 + bb) * 360 + k0 + 4 * kq));
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int u = 0; u < kWPer; ++u) reinterpret_cast<float4*>(&wsm[0][0])[threadIdx.x + u * 128] = wreg[u];
#pragma unroll
        for (int u = 0; u < kFPer; ++u) {
            const int i = threadIdx.x + u * 128, bb = i & 31, kq = i >> 5;
            if (kq < kVhChunk / 4) {
                fsm[4 * kq + 0][bb] = freg[u].x;
                fsm[4 * kq + 1][bb] = freg[u].y;
                fsm[4 * kq + 2][bb] = freg[u].z;
                fsm[4 * kq + 3][bb] = freg[u].w;
            }
        }
    };
    float acc[8][4];
    {
        const float4 bb = *reinterpret_cast<const float4*>(b1 + 4 * lane);
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i][0] = bb.x; acc[i][1] = bb.y; acc[i][2] = bb.z; acc[i][3] = bb.w; }
    }
    fetch(0);
    for (int k0 = 0; k0 < 360; k0 += kVhChunk) {
        __syncthreads();                                  // everyone is done with the previous chunk
        stash();
        __syncthreads();
        if (k0 + kVhChunk < 360) fetch(k0 + kVhChunk);    // in flight during the multiply below
#pragma unroll 4
        for (int k = 0; k < kVhChunk; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(&wsm[k][4 * lane]);
            const float4 f0 = *reinterpret_cast<const float4*>(&fsm[k][warp * 8]);
            const float4 f1 = *reinterpret_cast<const float4*>(&fsm[k][warp * 8 + 4]);
            const float f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                acc[i][0] = fmaf(w.x, f[i], acc[i][0]);
                acc[i][1] = fmaf(w.y, f[i], acc[i][1]);
                acc[i][2] = fmaf(w.z, f[i], acc[i][2]);
                acc[i][3] = fmaf(w.w, f[i], acc[i][3]);
            }
        }
    }
    const float4 wj = *reinterpret_cast<const float4*>(w2 + 4 * lane);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float t = fmaxf(acc[i][0], 0.0f) * wj.x;
        t = fmaf(fmaxf(acc[i][1], 0.0f), wj.y, t);
        t = fmaf(fmaxf(acc[i][2], 0.0f), wj.z, t);
        t = fmaf(fmaxf(acc[i][3], 0.0f), wj.w, t);
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        const int b = b0 + warp * 8 + i;
        if (lane == 0 && b < B) value[b] = tanhf(t + b2);
    }
}

}  // namespace xq

using namespace xq;

extern "C" void xq_net_free_(xq_ctx* c)
{
    if (!c || !c->net) return;
    NetState* N = reinterpret_cast<NetState*>(c->net);
    if (N->side) cudaStreamDestroy(N->side);
    for (int i = 0; i < 2; ++i) {
        if (N->vfork[i]) cudaStreamDestroy(N->vfork[i]);
        for (int j = 0; j < 2; ++j)
            if (N->vev[i][j]) cudaEventDestroy(N->vev[i][j]);
    }
    if (N->ev_fork) cudaEventDestroy(N->ev_fork);
    if (N->ev_join) cudaEventDestroy(N->ev_join);
    delete N;
    c->net = nullptr;
}

// n_boards_now > 0 overrides the descriptor's board count (and tile counts): the self-play loop sizes every
// launch to the leaves that are really waiting for an evaluation.
static int net_gemm(xq_ctx* c, const xq_gemm_desc* d, int n_boards_now, const int* n_dev, cudaStream_t s)
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
    if (n_boards_now > 0 && n_boards_now < d->n_boards) {
        a.n_boards = n_boards_now;
        a.m_tiles = d->mode == 2 ? (n_boards_now + 127) / 128 : (int)(((long long)n_boards_now * kBoardRows + 127) / 128);
    }
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
    a.n_dev = n_dev;
    XQ_CUDA(c, cudaSetDevice(c->device));
    XqTimer tm(c, s);
    if (d->mode == 2 && d->nt == 224 && d->kch_iter == 8 && d->kchunks % 8 == 0) {
        if (c->net_small && d->w_half && a.n_boards <= 256) {      // a launch bounded to a few boards: the 64-column tiling (below)
            GemmArgs a2 = a;
            a2.w = a.w_half;
            a2.n_tiles = (d->n_tiles * 224) / 64;                  // the bias vector and the logit rows are the 224-tiling's, 8288 >= 127 x 64 + ...
            if (a2.n_tiles > 127) a2.n_tiles = 127;                // 127 x 64 = 8128 >= 8100 logits
            return launch_fc<FcCfgT<64, 8, 5>>(c, a2, s, 10);
        }
        return launch_fc<FcCfgT<224, 8, 3>>(c, a, s, 0);
    }
    if (d->mode == 2 && d->nt == 128 && d->kch_iter == 8 && d->kchunks % 8 == 0) return launch_fc<FcCfgT<128, 8, 4>>(c, a, s, 1);
    // plans of at most 256 boards (one pair of row tiles): the layer is the 46 MB of weights, streamed by as many CTAs as there are
    // column tiles -- 127 tiles of 64 columns instead of 37 of 224 (29 -> ~10 us); the host tiles the image accordingly
    if (d->mode == 2 && d->nt == 64 && d->kch_iter == 8 && d->kchunks % 8 == 0) return launch_fc<FcCfgT<64, 8, 5>>(c, a, s, 10);
    if (d->mode == 0 && d->nt == 128 && d->kch_iter == 8 && d->a_row0 >= kHalo) {
        // 128-channel layers: 3 taps per weight stage (48 KB x 3 stages): a stage hand-off costs an MMA-issuing
        // thread ~0.2 us, so fewer, larger stages beat a finer ring; wider layers only have room for 16 KB stages
        // A handful of boards (evaluation games, the tail of self-play): what counts is the latency of a layer, and a 512-row
        // item of a CTA pair is 5.9 us of MMAs whatever the batch.  While every item fits one wave the layer runs as
        // 256-row x 64-channel items on single CTAs instead (the image tiled by 64 output channels is the CTA pairs' own):
        // 4x the items, a quarter of the MMA cycles each.
        if (c->net_small && d->w_half && (d->kchunks == 16 || d->kchunks == 32) && ((a.m_tiles + 1) / 2) * a.n_tiles * 2 <= c->sm_count) {
            GemmArgs a2 = a;
            a2.w = a.w_half;
            a2.n_tiles = a.n_tiles * 2;
            return launch_conv<64, 8, false, 3, false>(c, a2, s, 8);
        }
        if (c->net_2cta && d->w_half && d->kchunks % 8 == 0 && d->kchunks <= 32) {
            const int rc = launch_conv2<8, 3>(c, a, s, 7);                             // CTA pairs, cta_group::2
            if (rc != XQ_ERR_CUDA) return rc;
            // a device / partition that refuses the cluster launch: the single-CTA kernel computes the same layer
            (void)cudaGetLastError();
            c->net_2cta = false;
            fprintf(stderr, "[xq_b200] cluster launch of conv2_kernel failed (%s): using the single-CTA conv kernel\n", c->err);
        }
        if (d->kchunks == 16) return launch_conv<128, 8, false, 3, false>(c, a, s, 2);
        if (d->kchunks % 8 == 0 && d->kchunks <= 32) return launch_conv<128, 8, false, 1, false>(c, a, s, 3);
    }
    // towers of at most 64 channels (the reference's quick preset, train.py:661): N = 64 tiles, no zero-padded channels
    if (d->mode == 0 && d->nt == 64 && d->kch_iter == 8 && d->kchunks == 8 && d->n_tiles == 1 && d->a_row0 >= kHalo)
        return launch_conv<64, 8, false, 3, false>(c, a, s, 8);
    if (d->mode == 0 && d->nt == 64 && d->kch_iter == 2 && d->kchunks == 2 && d->n_tiles == 1 && d->a_row0 >= kHalo)
        return launch_conv<64, 2, false, 9, true>(c, a, s, 9);
    if (d->mode == 0 && d->nt == 128 && d->kch_iter == 2 && d->kchunks == 2 && d->n_tiles == 1 && d->a_row0 >= kHalo)
        return launch_conv<128, 2, false, 9, true>(c, a, s, 4);                      // 15-plane input conv, weights resident
    if (d->mode == 0 && d->nt == 128 && d->kch_iter == 2 && d->kchunks == 2 && d->a_row0 >= kHalo)
        return launch_conv<128, 2, false, 9, false>(c, a, s, 5);
    if (d->mode == 1 && d->nt == 48 && d->kch_iter == 8 && d->kchunks % 8 == 0 && d->kchunks <= 32 && d->out2 && d->n_tiles == 1)
        return launch_conv<48, 8, true, 1, true>(c, a, s, 6);                        // 1x1 heads, weights resident
    return xq_fail(c, XQ_ERR_ARG, "xq_net_gemm: unsupported (mode=%d nt=%d kch_iter=%d kchunks=%d n_tiles=%d a_row0=%lld)", d->mode,
                   d->nt, d->kch_iter, d->kchunks, d->n_tiles, (long long)d->a_row0);
}

extern "C" int xq_net_gemm(xq_ctx* c, const xq_gemm_desc* d, void* stream) { return net_gemm(c, d, 0, nullptr, (cudaStream_t)stream); }

static int net_value_head(xq_ctx* c, const float* d_feats, const float* d_w1t, const float* d_b1, const float* d_w2, float b2,
                          float* d_value, int B, const int* n_dev, cudaStream_t s)
{
    if (!c || !d_feats || !d_w1t || !d_b1 || !d_w2 || !d_value || B < 0)
        return xq_fail(c, XQ_ERR_ARG, "xq_net_value_head: bad arguments");
    if (B == 0) return XQ_OK;
    XQ_CUDA(c, cudaSetDevice(c->device));
    XQ_CUDA(c, launch_pdl(c, value_head_kernel, (B + kVhBoards - 1) / kVhBoards, 128, (size_t)0, s, d_feats, d_w1t, d_b1, d_w2, b2, d_value, B, n_dev));
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_net_value_head(xq_ctx* c, const float* d_feats, const float* d_w1t, const float* d_b1,
                                 const float* d_w2, float b2, float* d_value, int B, void* stream)
{
    return net_value_head(c, d_feats, d_w1t, d_b1, d_w2, b2, d_value, B, nullptr, (cudaStream_t)stream);
}

// side stream + fork / join events of the context (created on first use): the value-MLP fork below and the arena's second
// network (xq_mcts.cu) run there
extern "C" int xq_net_side_stream_(xq_ctx* c, cudaStream_t* side, cudaEvent_t* ev_fork, cudaEvent_t* ev_join)
{
    NetState* N = net_state(c);
    if (!N->side) {
        XQ_CUDA(c, cudaSetDevice(c->device));
        XQ_CUDA(c, cudaStreamCreateWithFlags(&N->side, cudaStreamNonBlocking));
        XQ_CUDA(c, cudaEventCreateWithFlags(&N->ev_fork, cudaEventDisableTiming));
        XQ_CUDA(c, cudaEventCreateWithFlags(&N->ev_join, cudaEventDisableTiming));
    }
    *side = N->side;
    *ev_fork = N->ev_fork;
    *ev_join = N->ev_join;
    return XQ_OK;
}

// All layers of one forward.  The value MLP depends on the heads conv only; with XQ_NET_FORK=1 it is forked onto a side
// stream right after that layer and runs next to the policy FC (their shared memory and registers fit one SM together).
// Measured on one box (profiles/r2_fwd_ab.txt): 1.002 ms per forward forked against 0.976 ms in sequence, so the default
// keeps the sequence.
static int net_run_impl(xq_ctx* c, const xq_gemm_desc* layers, int n_layers, const float* d_vfeats, const float* d_w1t,
                        const float* d_b1, const float* d_w2, float b2, float* d_value, int B, const int* n_dev, cudaStream_t s)
{
    NetState* N = net_state(c);
    // Up to kForkBoards boards the forward is a chain of kernels that each leave most SMs idle, and the value MLP (19 us on a
    // few CTAs) hides completely behind the policy FC (29 us on 37 CTAs): it is forked onto its own stream.  At batch 4096 the
    // co-running CTAs cost the FC more than they hide (profiles/r2_fwd_ab.txt), so there the sequence stays (XQ_NET_FORK=1 forces
    // the fork at every size).
    constexpr int kForkBoards = 1024;
    const bool want_fork = c->net_fork || B <= kForkBoards;
    cudaStream_t fs = nullptr;
    cudaEvent_t ef = nullptr, ej = nullptr;
    if (want_fork) {
        const int slot = (N->side && s == N->side) ? 1 : 0;
        if (!N->vfork[slot]) {
            XQ_CUDA(c, cudaSetDevice(c->device));
            XQ_CUDA(c, cudaStreamCreateWithFlags(&N->vfork[slot], cudaStreamNonBlocking));
            XQ_CUDA(c, cudaEventCreateWithFlags(&N->vev[slot][0], cudaEventDisableTiming));
            XQ_CUDA(c, cudaEventCreateWithFlags(&N->vev[slot][1], cudaEventDisableTiming));
        }
        fs = N->vfork[slot];
        ef = N->vev[slot][0];
        ej = N->vev[slot][1];
    }
    bool forked = false;
    for (int i = 0; i < n_layers; ++i) {
        int rc = net_gemm(c, &layers[i], B, n_dev, s);
        if (rc) return rc;
        if (want_fork && layers[i].mode == 1 && i + 1 < n_layers && !forked) {
            XQ_CUDA(c, cudaEventRecord(ef, s));
            XQ_CUDA(c, cudaStreamWaitEvent(fs, ef, 0));
            rc = net_value_head(c, d_vfeats, d_w1t, d_b1, d_w2, b2, d_value, B, n_dev, fs);
            if (rc) return rc;
            XQ_CUDA(c, cudaEventRecord(ej, fs));
            forked = true;
        }
    }
    if (forked) {
        XQ_CUDA(c, cudaStreamWaitEvent(s, ej, 0));
        return XQ_OK;
    }
    return net_value_head(c, d_vfeats, d_w1t, d_b1, d_w2, b2, d_value, B, n_dev, s);
}

// Run a whole forward (a list of layer descriptors followed by the value head) for the first B boards.
extern "C" int xq_net_run(xq_ctx* c, const xq_gemm_desc* layers, int n_layers, const float* d_vfeats,
                          const float* d_w1t, const float* d_b1, const float* d_w2, float b2, float* d_value, int B,
                          void* stream)
{
    if (B <= 0) return XQ_OK;
    return net_run_impl(c, layers, n_layers, d_vfeats, d_w1t, d_b1, d_w2, b2, d_value, B, nullptr, (cudaStream_t)stream);
}

// Same with the board count on the DEVICE (*d_n_boards, clamped to max_boards): the grids are sized for max_boards and
// every kernel cuts its tile loop to the live count, so a caller that compacts its leaves never synchronises.
extern "C" int xq_net_run_counted(xq_ctx* c, const xq_gemm_desc* layers, int n_layers, const float* d_vfeats,
                                  const float* d_w1t, const float* d_b1, const float* d_w2, float b2, float* d_value,
                                  const int* d_n_boards, int max_boards, void* stream)
{
    if (max_boards <= 0) return XQ_OK;
    if (!d_n_boards) return xq_fail(c, XQ_ERR_ARG, "xq_net_run_counted: d_n_boards is NULL");
    return net_run_impl(c, layers, n_layers, d_vfeats, d_w1t, d_b1, d_w2, b2, d_value, max_boards, d_n_boards, (cudaStream_t)stream);
}
