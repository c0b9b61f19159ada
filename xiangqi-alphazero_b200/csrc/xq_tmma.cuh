// xq_tmma.cuh -- the two tcgen05 kind::tf32 kernels of the training step (included by xq_tnet.cu only).
//
// Replaces cuDNN's convolution forward / backward-data / backward-filter and cuBLAS' sgemm inside the training step of
// AlphaZeroTrainer.train_network (training/train.py:397-423 over model.py:39-107).  fp32 storage, tf32 products, fp32
// accumulation in TMEM.
//
// Training plane layout (include/xq_b200.h): X[C/4][rows][4] float32, 110 rows per board (10 x (9 cells + 1 zero pad
// column) + one zero pad row).  A 3x3 tap is a row shift of dy*10+dx and the pad cells ARE conv2d's zero padding, so the
// three convolutions of a layer need no masks:
//   fprop  Y[row][co]  = sum_tap sum_ci X[row + shift][ci]  W[co][ci][tap]      tg_kernel
//   dgrad  dX[row][ci] = sum_tap sum_co dY[row - shift][co] W[co][ci][tap]      tg_kernel, transposed image, shift sign -1
//   wgrad  dW[co][ci][tap] = sum_row dY[row][co] X[row + shift][ci]             twg_kernel
// fprop and dgrad read the plane block [chunk][row][16 B] as a canonical no-swizzle K-major operand (rows = M, channels = K:
// 8 rows x 16 B core matrices, SBO = 128 B, LBO = chunk stride); dgrad is fprop with the transposed weight image and the
// taps mirrored.  wgrad contracts over ROWS, so both its operands are MN-major -- and for tf32 the tensor core accepts
// MN-major operands in ONE shared-memory layout only, SWIZZLE_128B_BASE32B (measured: no-swizzle MN-major tf32 MMAs return
// zeros; cutlass sm100_common.inl says the same): rows of 32 channels (128 B), the four 32-byte units of a row XORed with
// (row & 3).  The wgrad operands therefore exist a second time in global memory in exactly that form ("G layout":
// G[C/32][row][32 channels], unit u of row r stored at u ^ (r & 3), written by the same elementwise kernels that write the
// planes), so that a 1-D bulk copy of whole rows lands them in shared memory already swizzled.  No tensor maps anywhere.
#pragma once
#include "xq_ctx.h"

namespace xq {
namespace tn {

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 x tf32 -> fp32 (the low 13 mantissa bits of the fp32 operands are ignored)
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// instruction descriptor, kind::tf32: D = f32 (bit 4), A = B = tf32 (2 at bits 7 and 10), a_major bit 15, b_major bit 16
// (1 = MN-major), N >> 3 at 17, M >> 4 at 24 (cute/arch/mma_sm100_desc.hpp)
__host__ __device__ constexpr uint32_t tf32_idesc(int m, int n, int a_mn, int b_mn)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
// shared-memory matrix descriptor, no swizzle: [0,14) start >> 4, [16,30) LBO >> 4, [32,46) SBO >> 4, [46,48) version 1
// [61,64) layout type: 0 none, 1 SWIZZLE_128B_BASE32B
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 0)
{
    return (uint64_t)(((addr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16)) |
           ((uint64_t)((sbo_bytes >> 4) | (1u << 14) | (layout_type << 29)) << 32);
}

constexpr int kTB = XQ_T_BOARD_ROWS;         // 110 plane rows per board
constexpr int kTHalo = 11;                   // largest |dy*10+dx|
constexpr int kTPair = 256;                  // two 128-row tiles share every weight stage
constexpr int kTARows = kTPair + 2 * kTHalo; // 278
constexpr int kTAPlane = kTARows * 16;       // bytes of one chunk of the A block
constexpr int kTSeg = 8 * kTAPlane;          // one contraction block (32 channels) of the A block: 35 584 B
constexpr int kTStage = 8 * 128 * 16;        // one weight stage: 32 k x 128 n floats = 16 KB
constexpr int kTNSeg = 4, kTNStage = 5;
constexpr int kTgThreads = 352;              // warp 0 producer, warps 1 and 10 MMA issuers (row tile 0 / 1), warps 2-9 epilogue
constexpr int kTgSmem = kTNSeg * kTSeg + kTNStage * kTStage + 256;

struct TgArgs {
    const uint8_t* a;
    long long a_rows, a_row0;
    const uint8_t* w;
    int kblocks, ntaps, img_kb, shift_sign, m_pairs, n_tiles, out_chunks, n_cols;
    long long m_rows;
    uint8_t* out;
    long long out_rows, out_row0;
    const uint8_t* residual;
    float* out_rm;
    long long out_stride;
    const float* bias;
    int k_splits;                            // contraction split: item (pair, n_tile, ks) covers k-blocks [ks*kb_per, +kb_per)
    int kb_per;
    long long out_split_bytes;               // planes output of split ks goes to out + ks*out_split_bytes (partial sums, added by the reader)
};

// out[row][n] = sum_tap sum_k A[row + sign*shift(tap)][k] * image[n][k] (+ residual / + bias): fprop with the layer's weight
// image, dgrad with the image of the transposed weights and sign = -1.  B stage (n_tile, tap, k_block) = one 16 KB copy.
__global__ void __launch_bounds__(kTgThreads, 1) tg_kernel(const TgArgs p)
{
    extern __shared__ __align__(128) uint8_t smem[];
    griddep_launch_dependents();                 // the next kernel of the step may set itself up while this one runs
    uint8_t* sA = smem;
    uint8_t* sB = smem + kTNSeg * kTSeg;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + kTNStage * kTStage);
    uint64_t* a_full = bars;                 // [kTNSeg]
    uint64_t* a_empty = a_full + kTNSeg;
    uint64_t* w_full = a_empty + kTNSeg;     // [kTNStage]
    uint64_t* w_empty = w_full + kTNStage;
    uint64_t* t_full = w_empty + kTNStage;   // [2 accumulator stages][2 row tiles]
    uint64_t* t_empty = t_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 4);

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int total = p.m_pairs * p.n_tiles * p.k_splits;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kTNSeg; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 2);       // both MMA issuers release a segment
        }
        for (int i = 0; i < kTNStage; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 2);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&t_full[i], 1);
            mbar_init(&t_empty[i], 4);       // the 4 epilogue warps of that row tile
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
        // ===================== bulk-copy producer =====================
        int sg = 0, s = 0;
        uint32_t sgph = 0, ph = 0;
        // The weight image was written at the start of the step, long before the kernel in front of this one: the first
        // item's first weight stages (all stages are empty at kernel start) go out BEFORE the dependency wait, in the order
        // the loop below would issue them.
        int pre = 0;
        if ((int)blockIdx.x < total) {
            const int ks0 = (int)blockIdx.x % p.k_splits, wr0 = (int)blockIdx.x / p.k_splits;
            const int n_tile0 = wr0 % p.n_tiles;
            const int kb00 = ks0 * p.kb_per, kb01 = min(p.kblocks, kb00 + p.kb_per);
            const int n_first = (kb01 - kb00) * p.ntaps;
            pre = n_first < kTNStage ? n_first : kTNStage;
            if (lane == 0)
                for (int q = 0; q < pre; ++q) {
                    const int kb = kb00 + q / p.ntaps, tap = q % p.ntaps;
                    mbar_expect_tx(&w_full[q], (uint32_t)kTStage);
                    bulk_g2s(sB + q * kTStage, p.w + ((size_t)(n_tile0 * p.ntaps + tap) * p.img_kb + kb) * kTStage, kTStage, &w_full[q]);
                }
        }
        griddep_wait();                          // from here on the previous kernels' outputs (the A operand) may be read
        bool first = true;
        for (int work = blockIdx.x; work < total; work += gridDim.x, first = false) {
            const int ks = work % p.k_splits, wr = work / p.k_splits;
            const int pair = wr / p.n_tiles, n_tile = wr - pair * p.n_tiles;
            const long long m0 = (long long)pair * kTPair;
            const int kb0 = ks * p.kb_per, kb1 = min(p.kblocks, kb0 + p.kb_per);
            int q = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                if (lane == 0) {
                    mbar_wait(&a_empty[sg], sgph ^ 1u);
                    mbar_expect_tx(&a_full[sg], (uint32_t)kTSeg);
                    for (int c = 0; c < 8; ++c)
                        bulk_g2s(sA + sg * kTSeg + c * kTAPlane,
                                 p.a + ((size_t)(kb * 8 + c) * p.a_rows + (size_t)(p.a_row0 + m0 - kTHalo)) * 16, kTAPlane, &a_full[sg]);
                }
                if (++sg == kTNSeg) { sg = 0; sgph ^= 1u; }
                for (int tap = 0; tap < p.ntaps; ++tap, ++q) {
                    if (lane == 0 && (!first || q >= pre)) {
                        mbar_wait(&w_empty[s], ph ^ 1u);
                        mbar_expect_tx(&w_full[s], (uint32_t)kTStage);
                        bulk_g2s(sB + s * kTStage, p.w + ((size_t)(n_tile * p.ntaps + tap) * p.img_kb + kb) * kTStage, kTStage, &w_full[s]);
                    }
                    if (++s == kTNStage) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1 || warp == 10) {
        // ===================== MMA issuers: one warp per row tile, one elected lane issues =====================
        const int t = warp == 1 ? 0 : 1;
        const bool leader = elect_one();
        constexpr uint32_t idesc = tf32_idesc(128, 128, 0, 0);
        int sg = 0, s = 0, n = 0;
        uint32_t sgph = 0, ph = 0;
        for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
            const int acc = n & 1;
            const uint32_t tph = (uint32_t)(n >> 1) & 1u;
            mbar_wait(&t_empty[acc * 2 + t], tph ^ 1u);
            tc_fence_after();
            const uint32_t d_addr = tmem_base + (uint32_t)(acc * 256 + t * 128);
            const int ks = work % p.k_splits;
            const int kb0 = ks * p.kb_per, kb1 = min(p.kblocks, kb0 + p.kb_per);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&a_full[sg], sgph);
                const uint32_t a_seg = smem_u32(sA + sg * kTSeg);
                for (int tap = 0; tap < p.ntaps; ++tap) {
                    mbar_wait(&w_full[s], ph);
                    tc_fence_after();
                    const int shift = p.ntaps == 9 ? p.shift_sign * ((tap / 3 - 1) * 10 + (tap % 3 - 1)) : 0;
                    const uint32_t a_addr = a_seg + (uint32_t)((kTHalo + shift + t * 128) * 16);
                    const uint32_t b_addr = smem_u32(sB + s * kTStage);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint64_t adesc = smem_desc(a_addr + (uint32_t)(2 * j * kTAPlane), (uint32_t)kTAPlane, 128u);
                        const uint64_t bdesc = smem_desc(b_addr + (uint32_t)(2 * j * 2048), 2048u, 128u);
                        if (leader) umma_tf32(d_addr, adesc, bdesc, idesc, ((kb - kb0) | tap | j) != 0 ? 1u : 0u);
                    }
                    if (leader) umma_commit(&w_empty[s]);
                    if (++s == kTNStage) { s = 0; ph ^= 1u; }
                    __syncwarp();
                }
                if (leader) umma_commit(&a_empty[sg]);
                if (++sg == kTNSeg) { sg = 0; sgph ^= 1u; }
            }
            if (leader) umma_commit(&t_full[acc * 2 + t]);
            __syncwarp();
        }
    } else {
        // ===================== epilogue: 8 warps, warp -> (row tile t, TMEM lane quarter q) =====================
        const int q = warp & 3;
        const int t = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        int n = 0;
        griddep_wait();                          // residual reads and output stores come after the previous kernels
        for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
            const int ks = work % p.k_splits, wr = work / p.k_splits;
            const int pair = wr / p.n_tiles, n_tile = wr - pair * p.n_tiles;
            const int acc = n & 1;
            const uint32_t tph = (uint32_t)(n >> 1) & 1u;
            const long long mrow = (long long)pair * kTPair + t * 128 + row;
            uint8_t* const out_ks = p.out + (size_t)ks * p.out_split_bytes;
            const bool real = mrow < p.m_rows;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + t * 128);
            mbar_wait(&t_full[acc * 2 + t], tph);
            tc_fence_after();
#pragma unroll 1
            for (int sl = 0; sl < 4; ++sl) {
                uint32_t v[32];
                tmem_ld32(taddr + (uint32_t)(sl * 32), v);
                tmem_ld_wait();
                if (!real) continue;
                if (p.out_rm) {
                    float* dst = p.out_rm + (size_t)mrow * p.out_stride;
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int col = n_tile * 128 + sl * 32 + g * 4;
                        if (col < p.n_cols) {
                            float4 o = make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]),
                                                   __uint_as_float(v[g * 4 + 3]));
                            if (p.bias && ks == 0) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                            }
                            // two splits add into a zeroed matrix: 0 + a + b = 0 + b + a bit for bit, the order does not matter
                            if (p.k_splits > 1) atomicAdd(reinterpret_cast<float4*>(dst + col), o);
                            else *reinterpret_cast<float4*>(dst + col) = o;
                        }
                    }
                } else {
                    // the 8 residual loads of a slice are issued together, ahead of the stores (one round trip, not eight)
                    float4 r[8];
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int chunk = n_tile * 32 + sl * 8 + g;
                        const size_t off = ((size_t)chunk * p.out_rows + (size_t)(p.out_row0 + mrow)) * 16;
                        r[g] = (p.residual && chunk < p.out_chunks) ? __ldg(reinterpret_cast<const float4*>(p.residual + off))
                                                                    : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    }
#pragma unroll
                    for (int g = 0; g < 8; ++g) {
                        const int chunk = n_tile * 32 + sl * 8 + g;
                        if (chunk < p.out_chunks) {
                            const size_t off = ((size_t)chunk * p.out_rows + (size_t)(p.out_row0 + mrow)) * 16;
                            *reinterpret_cast<float4*>(out_ks + off) =
                                make_float4(__uint_as_float(v[g * 4]) + r[g].x, __uint_as_float(v[g * 4 + 1]) + r[g].y,
                                            __uint_as_float(v[g * 4 + 2]) + r[g].z, __uint_as_float(v[g * 4 + 3]) + r[g].w);
                        }
                    }
                }
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

// =============================================================================================================
// twg_kernel: weight gradients.  D_t[m][n] = sum_row A[row][m] * B[row + off_t][n], both operands MN-major in the
// SWIZZLE_128B_BASE32B form, read straight from the G layout (see the file header).
// =============================================================================================================
// A stage: [4 groups of 32 channels][kr rows][128 B]; B stage: [b_groups_stage][b_rows_stage][128 B]; every group of a stage
// is ONE bulk copy of whole rows.  Descriptors: LBO = group stride (next 32 channels), SBO = 512 B (next 4 rows), start =
// first row of the MMA's 8 rows -- the swizzle is a function of the absolute shared-memory address, so a tap may start
// at any row as long as every stage row r sits at an address whose bits 7-8 equal (global row & 3): stage bases are 512-byte
// aligned and copies start at global rows that are multiples of 4.
constexpr int kTwgThreads = 192;             // warp 0 producer (lanes copy groups), warp 1 MMA issuer, warps 2-5 epilogue
constexpr int kTwgMaxStages = 4;

struct TwgArgs {
    const uint8_t* a;
    const uint8_t* b;
    long long a_rows, a_row0, b_rows, b_row0;
    int a_group0, b_group0, nbg, kr, stages_per_item, n_slabs, n_groups, n_mtiles, taps_per_group;
    int b_rows_stage, b_groups_stage, b_group_step;
    int b_row_lo[4];                         // first stage row of tap group g relative to the stage's first A row (multiple of 4)
    int tap_off[16];                         // [group][tap]: byte offset of the tap's first row / first channel group in the B stage
    float* out;
    long long mt_stride, slab_stride, g_stride, tap_stride, ldo;
    int m_limit, n_limit, g_cols, t_cols;
    int n_stages, stage_bytes, a_stage_bytes; // pipeline geometry (host computed)
};

__global__ void __launch_bounds__(kTwgThreads, 1) twg_kernel(const TwgArgs p)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    griddep_launch_dependents();
    uint8_t* smem = smem_raw + ((512u - (smem_u32(smem_raw) & 511u)) & 511u);     // stage bases 512-byte aligned (swizzle phase)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * p.stage_bytes);
    uint64_t* full = bars;                   // [kTwgMaxStages]
    uint64_t* empty = bars + kTwgMaxStages;
    uint64_t* t_full = empty + kTwgMaxStages;
    uint64_t* t_empty = t_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 1);
    float* scratch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 128);    // [4 epilogue warps][32][33]: transpose tiles

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int total = p.n_mtiles * p.n_slabs * p.n_groups;
    const int N = 32 * p.nbg;
    const int tcol = N <= 32 ? 32 : (N <= 64 ? 64 : 128);      // TMEM columns per tap accumulator

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.n_stages; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(t_full, 1);
        mbar_init(t_empty, 4);
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
    const uint32_t a_bytes = (uint32_t)(p.kr * 128), b_bytes = (uint32_t)(p.b_rows_stage * 128);   // one channel group of a stage

    if (warp == 0) {
        int s = 0;
        uint32_t ph = 0;
        griddep_wait();                          // both operands are outputs of the kernels in front (barriers and TMEM are set up by now)
        for (int work = blockIdx.x; work < total; work += gridDim.x) {
            const int g = work % p.n_groups, rest = work / p.n_groups;
            const int slab = rest % p.n_slabs, mt = rest / p.n_slabs;
            for (int st = 0; st < p.stages_per_item; ++st) {
                const long long k0 = ((long long)slab * p.stages_per_item + st) * p.kr;
                if (lane == 0) {
                    mbar_wait(&empty[s], ph ^ 1u);
                    mbar_expect_tx(&full[s], 4u * a_bytes + (uint32_t)p.b_groups_stage * b_bytes);
                }
                __syncwarp();
                uint8_t* sa = smem + (size_t)s * p.stage_bytes;
                uint8_t* sb = sa + p.a_stage_bytes;
                if (lane < 4)
                    bulk_g2s(sa + lane * a_bytes, p.a + ((size_t)(p.a_group0 + mt * 4 + lane) * p.a_rows + (size_t)(p.a_row0 + k0)) * 128, a_bytes,
                             &full[s]);
                else if (lane - 4 < p.b_groups_stage)
                    bulk_g2s(sb + (lane - 4) * b_bytes,
                             p.b + ((size_t)(p.b_group0 + g * p.b_group_step + lane - 4) * p.b_rows + (size_t)(p.b_row0 + k0 + p.b_row_lo[p.n_groups <= 4 ? g : 0])) * 128,
                             b_bytes, &full[s]);
                if (++s == p.n_stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        const uint32_t idesc = tf32_idesc(128, N, 1, 1);
        int s = 0, n = 0;
        uint32_t ph = 0;
        for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
            const int gi = p.n_groups <= 4 ? work % p.n_groups : 0;      // more than 4 tap groups (dense layers): all use entry 0
            mbar_wait(t_empty, (uint32_t)(n & 1) ^ 1u);
            tc_fence_after();
            for (int st = 0; st < p.stages_per_item; ++st) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)s * p.stage_bytes);
                const uint32_t sb = sa + (uint32_t)p.a_stage_bytes;
                for (int ks = 0; ks < p.kr / 8; ++ks) {
                    const uint64_t adesc = smem_desc(sa + (uint32_t)(ks * 1024), a_bytes, 512u, 1u);
                    for (int t = 0; t < p.taps_per_group; ++t) {
                        const uint64_t bdesc = smem_desc(sb + (uint32_t)p.tap_off[gi * 4 + t] + (uint32_t)(ks * 1024), b_bytes, 512u, 1u);
                        if (leader) umma_tf32(tmem_base + (uint32_t)(t * tcol), adesc, bdesc, idesc, (st | ks) != 0 ? 1u : 0u);
                    }
                }
                if (leader) umma_commit(&empty[s]);
                if (++s == p.n_stages) { s = 0; ph ^= 1u; }
                __syncwarp();
            }
            if (leader) umma_commit(t_full);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;                                  // warps 2,3,4,5 -> TMEM lane quarters 2,3,0,1
        float* scr = scratch + (warp - 2) * (32 * 33);
        int n = 0;
        griddep_wait();                          // the output buffer may still be read by the kernel in front
        for (int work = blockIdx.x; work < total; work += gridDim.x, ++n) {
            const int g = work % p.n_groups, rest = work / p.n_groups;
            const int slab = rest % p.n_slabs, mt = rest / p.n_slabs;
            mbar_wait(t_full, (uint32_t)(n & 1));
            tc_fence_after();
            // A thread holds 32 consecutive columns of ONE output row; stored as they are, a warp's store instruction would
            // touch 32 rows x 16 B.  Each 32 x 32 tile goes through a padded shared-memory tile instead and leaves as 4 rows x
            // 128 contiguous bytes per instruction (ncu: the scattered form kept the tensor pipe at 32 % of the active cycles).
            float* base = p.out + (size_t)mt * p.mt_stride + (size_t)slab * p.slab_stride + (size_t)g * p.g_stride;
            const int r_sub = lane >> 3, c4 = (lane & 7) * 4;
            for (int t = 0; t < p.taps_per_group; ++t) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * tcol);
                float* dst = base + (size_t)t * p.tap_stride;
                const int col0 = g * p.g_cols + t * p.t_cols;
                for (int c0 = 0; c0 < N; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + (uint32_t)c0, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) scr[lane * 33 + j] = __uint_as_float(v[j]);
                    __syncwarp();
                    const bool n_ok = col0 + c0 + c4 < p.n_limit;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = i * 4 + r_sub, m = q * 32 + r;
                        const float* sp = scr + r * 33 + c4;
                        const float4 o = make_float4(sp[0], sp[1], sp[2], sp[3]);
                        if (n_ok && mt * 128 + m < p.m_limit) *reinterpret_cast<float4*>(dst + (size_t)m * p.ldo + c0 + c4) = o;
                    }
                    __syncwarp();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tn
}  // namespace xq
