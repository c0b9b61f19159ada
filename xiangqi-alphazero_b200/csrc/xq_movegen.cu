// xq_movegen.cu -- K1: batched rules engine (legal moves, in-check, feature planes),
// is_attacked queries and device-side random playouts, plus the context entry points.
//
// Replaces training/cython_engine/game_core.pyx (cy_generate_legal_moves :521-540,
// cy_is_in_check :543-555, cy_is_attacked :508-518) and game.py get_state_for_nn :618-640.
//
// Two kernel generations answer xq_movegen_batch with the same bytes (xq_set_movegen_impl):
//   movegen_tpb_kernel (default): one THREAD per board, a warp per 32 positions; rules in xq_rules_tpb.h.
//       915 M positions/s with planes = 78 % of the measured HBM bandwidth (profiles/r1_movegen_ncu.md).
//   movegen_kernel: one WARP per board (rules in xq_rules.cuh, the generator the MCTS kernels use for the one
//       game a warp owns).  419 M positions/s, bound by instruction issue at 17.5 of 32 active lanes.
// Shape of the first generation (HBM/instruction-bound integer work, no tensor cores):
//   - persistent CTAs (grid = SMs x resident CTAs), 8 warps, one warp per board;
//   - boards arrive in shared memory as 64-board tiles (5760 B + 64 B of sides) through the
//     TMA engine (cp.async.bulk + mbarrier, double buffered) so the next tile streams in
//     while the current one is being searched;
//   - move lists leave as one 8-byte store per lane (256 B per board, coalesced), planes as
//     coalesced float2 streaming stores, counts / in-check flags as one 64 B row per tile.
#include "xq_ctx.h"
#include "xq_rules.cuh"
#include "xq_rules_tpb.h"
#include <cstdlib>

char g_xq_last_error[512] = {0};

namespace xq {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kTile = 64;                       // boards per CTA tile
constexpr int kTileBytes = kTile * kSquares;    // 5760, multiple of 16
constexpr int kPlaneWords = 44;                 // 1350 bits = 43 words, + 1 so that the funnel shift may read word 43

struct __align__(16) MovegenSmem {
    int8_t boards[2][kTileBytes];
    int8_t sides[2][kTile];
    uint8_t n_out[2][kTile];
    uint8_t chk_out[2][kTile];
    uint64_t bar[2];
    WarpScratch ws[kWarps];
    uint32_t pbits[kWarps][2][kPlaneWords];     // the 1350 plane values of a position as bits (two buffers: the next
                                                // position's bits are cleared while slow lanes still read these)
    float4 nib_lut[16];                         // 4 bits -> 4 floats
};

template <bool PLANES>
__global__ void __launch_bounds__(kThreads, 4)   // 4 CTAs/SM (61 registers); 5 or 6 resident CTAs measured no faster
movegen_kernel(const int8_t* __restrict__ boards, const int8_t* __restrict__ sides, int B,
               int16_t* __restrict__ actions, uint8_t* __restrict__ n_moves,
               uint8_t* __restrict__ in_check, float* __restrict__ planes, int* __restrict__ overflow,
               int bulk_ok)
{
    __shared__ MovegenSmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (B + kTile - 1) / kTile;

    if (threadIdx.x == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        mbar_fence_init();
    }
    if (PLANES && threadIdx.x < 16)
        sm.nib_lut[threadIdx.x] = make_float4((threadIdx.x & 1) ? 1.0f : 0.0f, (threadIdx.x & 2) ? 1.0f : 0.0f,
                                              (threadIdx.x & 4) ? 1.0f : 0.0f, (threadIdx.x & 8) ? 1.0f : 0.0f);
    __syncthreads();

    auto tile_is_bulk = [&](int t) { return bulk_ok && (t + 1) * kTile <= B; };
    auto issue = [&](int t, int buf) {   // thread 0 only
        mbar_expect_tx(&sm.bar[buf], kTileBytes + kTile);
        bulk_g2s(sm.boards[buf], boards + (size_t)t * kTileBytes, kTileBytes, &sm.bar[buf]);
        bulk_g2s(sm.sides[buf], sides + (size_t)t * kTile, kTile, &sm.bar[buf]);
    };

    uint32_t phase[2] = {0, 0};
    int t = blockIdx.x;
    if (t < num_tiles && tile_is_bulk(t) && threadIdx.x == 0) issue(t, 0);

    for (int it = 0; t < num_tiles; ++it, t += gridDim.x) {
        const int buf = it & 1;
        const int tn = t + gridDim.x;
        // the other buffer was released by the __syncthreads that ended the previous iteration
        if (tn < num_tiles && tile_is_bulk(tn) && threadIdx.x == 0) issue(tn, buf ^ 1);

        const int nb = min(kTile, B - t * kTile);
        if (tile_is_bulk(t)) {
            mbar_wait(&sm.bar[buf], phase[buf]);
            phase[buf] ^= 1;
        } else {
            // ragged last tile / unaligned caller buffer: plain cooperative copy
            for (int i = threadIdx.x; i < nb * kSquares; i += kThreads)
                sm.boards[buf][i] = boards[(size_t)t * kTileBytes + i];
            for (int i = threadIdx.x; i < nb; i += kThreads) sm.sides[buf][i] = sides[(size_t)t * kTile + i];
            __syncthreads();
        }

        for (int j = warp, flip = 0; j < nb; j += kWarps, flip ^= 1) {
            const int8_t* b = &sm.boards[buf][j * kSquares];
            const int side = sm.sides[buf][j];
            const size_t gi = (size_t)t * kTile + j;
            WarpScratch& S = sm.ws[warp];

            uint32_t* bits = sm.pbits[warp][flip];
            if (PLANES) {         // cleared here: warp_movegen's warp-wide syncs order it before the atomics below
                bits[lane] = 0u;
                if (lane < kPlaneWords - 32) bits[32 + lane] = 0u;
            }
            MovegenResult r = warp_movegen(b, side, S);
            if (lane == 0) {
                sm.n_out[buf][j] = (uint8_t)min(r.n_legal, kMaxMoves);
                sm.chk_out[buf][j] = r.in_check ? 1 : 0;
                if (r.overflow) atomicAdd(overflow, 1);
            }
            // 128 int16 = 256 B: one 8-byte store per lane
            {
                const uint2 v = reinterpret_cast<const uint2*>(S.actions)[lane];
                __stcs(reinterpret_cast<uint2*>(actions + gi * kMaxMoves) + lane, v);
            }
            if (PLANES) {
                // game.py:618-640: plane = kind-1 for the side to move, 7+kind-1 for the other side, plane 14 = 1
                // iff red moves.  The 1350 values are first set as BITS (one shared-memory atomic per piece), then
                // every lane expands 4 bits at a time into one float4 through a 16-entry table: ~10 instructions
                // per 16-byte store instead of 4 compares + selects.
                for (int sq = lane; sq < kSquares; sq += 32) {
                    const int v = b[sq] * side;
                    if (v != 0) {
                        const int e = (v > 0 ? v - 1 : 6 - v) * kSquares + sq;
                        atomicOr(&bits[e >> 5], 1u << (e & 31));
                    }
                }
                // plane 14 = elements 1260..1349 = bits 12..31 of word 39, words 40-41, bits 0..5 of word 42
                if (side == 1 && lane < 4)
                    atomicOr(&bits[39 + lane], lane == 0 ? 0xfffff000u : (lane == 3 ? 0x3fu : 0xffffffffu));
                warp_sync();
                float* outp = planes + gi * (15 * kSquares);
                // 1350 floats per position; positions alternate between 16-byte aligned and 8-byte offset bases:
                // one float2 at the head (odd positions) or tail (even), 337 coalesced float4 in between
                // ... for a 16-byte aligned planes buffer; an 8-byte aligned one (the ABI's minimum) swaps the two cases
                const int head = (int)((gi + ((reinterpret_cast<uintptr_t>(planes) >> 3) & 1)) & 1) * 2;
                if (lane == 0) {
                    const uint32_t two = head ? bits[0] : bits[42] >> 4;
                    float2 v;
                    v.x = (two & 1u) ? 1.0f : 0.0f;
                    v.y = (two & 2u) ? 1.0f : 0.0f;
                    __stcs(reinterpret_cast<float2*>(head ? outp : outp + 1348), v);
                }
                float4* out4 = reinterpret_cast<float4*>(outp + head);
                const int e0 = head + 4 * lane;              // first element of this lane's float4 in iteration 0
                const int sh = e0 & 31;
                const uint32_t* wp = bits + (e0 >> 5);       // advances 4 words (128 elements) per iteration
#pragma unroll
                for (int it = 0; it < 11; ++it) {
                    const int k = it * 32 + lane;
                    if (it < 10 || k < 337) {
                        const uint32_t nib = __funnelshift_r(wp[4 * it], wp[4 * it + 1], sh) & 15u;
                        __stcs(out4 + k, sm.nib_lut[nib]);
                    }
                }
            }
            // no warp-wide sync here: the next position starts with warp_board_scan (a __syncwarp), which orders this
            // position's reads of S.actions before their re-initialisation; the plane bits are double buffered
        }
        __syncthreads();
        // per-tile rows of counts and flags
        if (nb == kTile && (((uintptr_t)n_moves | (uintptr_t)in_check) & 3) == 0) {
            if (threadIdx.x < 16)
                reinterpret_cast<uint32_t*>(n_moves + (size_t)t * kTile)[threadIdx.x] =
                    reinterpret_cast<const uint32_t*>(sm.n_out[buf])[threadIdx.x];
            else if (threadIdx.x < 32)
                reinterpret_cast<uint32_t*>(in_check + (size_t)t * kTile)[threadIdx.x - 16] =
                    reinterpret_cast<const uint32_t*>(sm.chk_out[buf])[threadIdx.x - 16];
        } else {
            for (int i = threadIdx.x; i < nb; i += kThreads) {
                n_moves[(size_t)t * kTile + i] = sm.n_out[buf][i];
                in_check[(size_t)t * kTile + i] = sm.chk_out[buf][i];
            }
        }
    }
}

// ---- K1, second generation: one THREAD per board (rules: xq_rules_tpb.h) -----------------------------------
// A warp takes 32 consecutive positions (tasks are handed out by an atomic counter): their 2 880 board bytes arrive
// with coalesced 16-byte loads into the warp's own shared-memory slice, every lane then runs the scalar generator on
// its board (pseudo-legal list built and compacted in a 98-entry per-lane shared-memory array), and the warp leaves
// the results cooperatively -- one 8-byte store per lane per move list (256 B rows), the planes of a PAIR of
// positions as 675 coalesced float4 expanded from a 2 700-bit array, counts and flags as one 32-byte row.  Warps never
// wait for each other: no __syncthreads after the table set-up.  9.3 KB of shared memory per warp and 80 registers
// give 6 CTAs x 4 warps per SM.
constexpr int kTpbWarps = 4;
constexpr int kTpbThreads = kTpbWarps * 32;
constexpr int kPairWords = 86;             // 2 x 1350 plane bits = 2700 bits = 84.4 words (+1: the last float4 group reads word 84)
constexpr int kTpbListStride = 98;         // uint16 per lane: 196 B = 49 words (odd: equal indices fall on different banks)
static_assert(kTpbListStride >= xqt::kListCap && (kTpbListStride & 1) == 0 && ((kTpbListStride / 2) & 1) == 1, "list stride");

struct __align__(16) TpbWarpSmem {
    uint16_t list[32 * kTpbListStride];    // per-lane pseudo-legal scratch (first: the generator may READ up to 20 bytes
                                           // before / after a board for off-board targets it then discards)
    int8_t boards[32 * kSquares];          // 2 880 B, lane l owns bytes [90 l, 90 l + 90)
    int8_t sides[32];
    uint8_t n_out[32];                     // legal-move count of each lane's position and how many ids it already
    uint8_t n_flushed[32];                 // wrote to its output row itself
};
static_assert(16 * kPairWords * 4 <= 32 * kTpbListStride * 2, "the 16 plane-bit buffers of a task reuse the list storage");
static_assert(sizeof(TpbWarpSmem) % 16 == 0 && (sizeof(uint16_t) * 32 * kTpbListStride) % 16 == 0, "16-byte aligned boards");

struct __align__(16) TpbSmem {
    TpbWarpSmem w[kTpbWarps];
    uint32_t slot_tab[xqt::kSlotTableSize];   // leaper table of xq_rules_tpb.h (256 B)
};

template <bool PLANES>
__global__ void __launch_bounds__(kTpbThreads, 6)
movegen_tpb_kernel(const int8_t* __restrict__ boards, const int8_t* __restrict__ sides, int B,
                   int16_t* __restrict__ actions, uint8_t* __restrict__ n_moves,
                   uint8_t* __restrict__ in_check, float* __restrict__ planes, int* __restrict__ overflow,
                   int* __restrict__ task_counter, int vec_ok)
{
    extern __shared__ __align__(16) unsigned char tpb_smem_raw[];
    TpbSmem& sm = *reinterpret_cast<TpbSmem*>(tpb_smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TpbWarpSmem& W = sm.w[warp];
    if (threadIdx.x < xqt::kSlotTableSize) sm.slot_tab[threadIdx.x] = xqt::slot_entry(threadIdx.x);
    __syncthreads();

    const int n_tasks = (B + 31) >> 5;
    // tasks are handed out dynamically (one atomic per 32 positions): positions differ in cost and a static split
    // leaves the last wave of warps unevenly loaded
    for (;;) {
        int t = 0;
        if (lane == 0) t = atomicAdd(task_counter, 1);
        t = warp_bcast(t, 0);
        if (t >= n_tasks) break;
        const int base = t << 5;
        const int nb = min(32, B - base);
        // the previous task's cooperative reads of W are complete for every lane before anything is overwritten
        warp_sync();
        if (vec_ok && nb == 32) {
            const uint4* src = reinterpret_cast<const uint4*>(boards + (size_t)base * kSquares);
            uint4* dst = reinterpret_cast<uint4*>(W.boards);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int k = i * 32 + lane;
                if (k < 180) dst[k] = __ldcs(src + k);
            }
        } else {
            // ragged last task / unaligned caller buffer: plain copy; lanes without a position get an empty board
            for (int i = lane; i < 32 * kSquares; i += 32)
                W.boards[i] = i < nb * kSquares ? boards[(size_t)base * kSquares + i] : (int8_t)0;
        }
        int side = 1;
        if (lane < nb) side = sides[base + lane];
        W.sides[lane] = (int8_t)side;
        warp_sync();

        uint16_t* list = W.list + lane * kTpbListStride;
        int chk = 0, flushed = 0;
        uint32_t occ[3];
        int16_t* out_row = actions + (size_t)(base + (lane < nb ? lane : 0)) * kMaxMoves;   // idle lanes find no move
        int n = xqt::movegen(W.boards + lane * kSquares, side, list, out_row, &flushed, &chk, sm.slot_tab, occ);   // warp-synchronous: all 32 lanes
        if (n > kMaxMoves) {
            atomicAdd(overflow, 1);
            n = kMaxMoves;
        }
        W.n_out[lane] = (uint8_t)n;
        W.n_flushed[lane] = (uint8_t)flushed;
        if (lane < nb) {
            n_moves[base + lane] = (uint8_t)n;
            in_check[base + lane] = (uint8_t)chk;
        }
        warp_sync();

        // move lists: 128 int16 per position leave as one coalesced 8-byte store per lane (entries 4 lane .. 4 lane + 3
        // from the position's staged ids, -1 behind the last move).  A position that flushed ids during generation
        // (piece crowds no game reaches) keeps them and gets the rest entry by entry.
        for (int j = 0; j < nb; ++j) {
            const int nj = W.n_out[j], fj = W.n_flushed[j];
            const uint16_t* lj = W.list + j * kTpbListStride;
            int16_t* row = actions + (size_t)(base + j) * kMaxMoves;
            if (fj == 0) {
                const uint32_t* lw = reinterpret_cast<const uint32_t*>(lj);
                const int k0 = 4 * lane;
                // words beyond the staged ids are never read as data: the selects below replace them
                uint32_t a = k0 < nj ? lw[2 * lane] : 0xffffffffu;
                uint32_t c = k0 + 2 < nj ? lw[2 * lane + 1] : 0xffffffffu;
                if (k0 + 1 >= nj) a |= 0xffff0000u;
                if (k0 + 3 >= nj) c |= 0xffff0000u;
                __stcs(reinterpret_cast<uint2*>(row) + lane, make_uint2(a, c));
            } else {
                for (int k = lane; k < kMaxMoves; k += 32)
                    if (k >= fj) row[k] = k < nj ? (int16_t)lj[k - fj] : (int16_t)-1;
            }
        }
        warp_sync();

        if (PLANES) {
            // game.py:618-640.  Two consecutive positions (even, odd) are 2 700 floats = 675 float4 starting on a 16-byte
            // boundary, so a PAIR is expanded at a time.  The 16 pairs of the task get 16 bit arrays of 2 700 bits in the
            // (now idle) list storage; every lane sets the bits of ITS position by walking the occupied squares it
            // found in the scan (one shared-memory atomic per piece, ~26 per board instead of a 90-cell sweep), and then
            // float4 k of a pair is nibble k of its bit array: one shared load, a shift, a 16-entry table and one
            // coalesced streaming store.
            uint32_t* pair_bits = reinterpret_cast<uint32_t*>(W.list);
            for (int i = lane; i < 16 * kPairWords; i += 32) pair_bits[i] = 0u;
            warp_sync();
            {
                uint32_t* bits = pair_bits + (lane >> 1) * kPairWords;
                const int ebase = (lane & 1) * (15 * kSquares);
                const int8_t* b = W.boards + lane * kSquares;
                uint32_t o0 = occ[0], o1 = occ[1], o2 = occ[2];
                // warp-uniform trip count and branch-free plane-14 words: a per-lane loop bound or a red/black branch here
                // left the warp split in two for the rest of the task (ncu: the stores below ran with 17.7 lanes)
                const int trips = XQT_WARP_MAX(__popc(o0) + __popc(o1) + __popc(o2));
                for (int t = 0; t < trips; ++t) {
                    if ((o0 | o1 | o2) != 0u) {
                        int sq;
                        if (o0) { sq = __ffs(o0) - 1; o0 &= o0 - 1u; }
                        else if (o1) { sq = 31 + __ffs(o1); o1 &= o1 - 1u; }
                        else { sq = 63 + __ffs(o2); o2 &= o2 - 1u; }
                        const int v = b[sq] * side;
                        const int e = ebase + (v > 0 ? v - 1 : 6 - v) * kSquares + sq;
                        atomicOr(&bits[e >> 5], 1u << (e & 31));
                    }
                    __syncwarp();
                }
                // plane 14 = ones iff red moves: elements 1260..1349 of the position = bits 1260..1349 (words 39..42) of an
                // even lane's half, bits 2610..2699 (words 81..84) of an odd lane's; the two middle words belong to it alone
                const uint32_t ones = (side == 1 && lane < nb) ? 0xffffffffu : 0u;
                const bool odd = (lane & 1) != 0;
                uint32_t* w14 = bits + (odd ? 81 : 39);
                atomicOr(w14, (odd ? 0xfffc0000u : 0xfffff000u) & ones);
                w14[1] = ones;
                w14[2] = ones;
                atomicOr(w14 + 3, (odd ? 0xfffu : 0x3fu) & ones);
            }
            warp_sync();
            const int npair = nb >> 1;
            const int sh = 4 * (lane & 7);
            for (int jp = 0; jp < npair; ++jp) {
                float4* out4 = reinterpret_cast<float4*>(planes + (size_t)(base + 2 * jp) * (15 * kSquares));
                const uint32_t* wp = pair_bits + jp * kPairWords + (lane >> 3);
#pragma unroll 2                 // not fully: the kernel is sensitive to its instruction footprint (ncu: no_instruction stalls)
                for (int it = 0; it < 22; ++it) {
                    const int k = it * 32 + lane;
                    if (it < 21 || k < 675) {
                        // 4 bits -> 4 floats without a table: 1.0f is 0x3f800000, so bit i of the nibble times
                        // (0x3f800000 >> i) is the float; a 16-entry float4 table in shared memory cost 4 wavefronts
                        // per load and a quarter of the kernel's shared-memory traffic (ncu)
                        const uint32_t nib = wp[4 * it] >> sh;
                        float4 v;
                        v.x = __uint_as_float((nib & 1u) * 0x3f800000u);
                        v.y = __uint_as_float((nib & 2u) * 0x1fc00000u);
                        v.z = __uint_as_float((nib & 4u) * 0x0fe00000u);
                        v.w = __uint_as_float((nib & 8u) * 0x07f00000u);
                        __stcs(out4 + k, v);
                    }
                }
            }
            if (nb & 1) {
                // the last position of the whole batch stands alone (once per launch): element by element
                const int j = nb - 1;
                const int8_t* b = W.boards + j * kSquares;
                const int sd = W.sides[j];
                float* outp = planes + (size_t)(base + j) * (15 * kSquares);
                for (int e = lane; e < 15 * kSquares; e += 32) {
                    const int pl = e / kSquares, sq = e - pl * kSquares;
                    const int v = b[sq] * sd;
                    const bool one = pl == 14 ? sd == 1 : (v != 0 && (v > 0 ? v - 1 : 6 - v) == pl);
                    outp[e] = one ? 1.0f : 0.0f;
                }
            }
        }
    }
}

// ---- is_attacked queries: one thread per query, boards staged per CTA ---------------------
constexpr int kAtkThreads = 128;

__global__ void __launch_bounds__(kAtkThreads)
is_attacked_kernel(const int8_t* __restrict__ boards, const uint8_t* __restrict__ sq,
                   const int8_t* __restrict__ by, int B, uint8_t* __restrict__ out)
{
    __shared__ int8_t sb[kAtkThreads * kSquares];
    const int base = blockIdx.x * kAtkThreads;
    const int nb = min(kAtkThreads, B - base);
    for (int i = threadIdx.x; i < nb * kSquares; i += kAtkThreads) sb[i] = boards[(size_t)base * kSquares + i];
    __syncthreads();
    if (threadIdx.x < nb) {
        const int s = sq[base + threadIdx.x];
        bool a = false;
        if (s < kSquares)
            a = attacked_sq(&sb[threadIdx.x * kSquares], s / 9, s % 9, by[base + threadIdx.x], -1, -1, 0);
        out[base + threadIdx.x] = a ? 1 : 0;
    }
}

// _is_move_legal (game_core.pyx:209-252; game.py:441-490): make the move from -> to on a copy (ANY move: the reference
// does not ask for pseudo-legality here), then `side`'s king must stand in its own palace, must not face the other king on
// an open file and must not be attacked.  The facing test needs no clause of its own: _is_attacked counts the enemy king as
// a rook on an open ray (pyx:117).
__global__ void __launch_bounds__(kAtkThreads)
move_is_legal_kernel(const int8_t* __restrict__ boards, const uint8_t* __restrict__ from, const uint8_t* __restrict__ to,
                     const int8_t* __restrict__ sides, int B, uint8_t* __restrict__ out)
{
    __shared__ int8_t sb[kAtkThreads * kSquares];
    const int base = blockIdx.x * kAtkThreads;
    const int nb = min(kAtkThreads, B - base);
    for (int i = threadIdx.x; i < nb * kSquares; i += kAtkThreads) sb[i] = boards[(size_t)base * kSquares + i];
    __syncthreads();
    if (threadIdx.x >= nb) return;
    int8_t* b = &sb[threadIdx.x * kSquares];
    const int f = from[base + threadIdx.x], t = to[base + threadIdx.x], side = sides[base + threadIdx.x];
    bool ok = false;
    if (f < kSquares && t < kSquares) {
        b[t] = b[f];
        b[f] = 0;
        const int target = side == 1 ? 1 : -1, r0 = side == 1 ? 0 : 7;
        int k = -1;
        for (int r = r0 + 2; r >= r0; --r)
            for (int c = 5; c >= 3; --c)
                if (b[r * 9 + c] == target) k = r * 9 + c;
        if (k >= 0) ok = !attacked_sq(b, k / 9, k % 9, -side, -1, -1, 0);
    }
    out[base + threadIdx.x] = ok ? 1 : 0;
}

// cy_find_king (game_core.pyx:78-101, 493-505): the side's king searched ONLY in its own 3x3 palace, row-major;
// out = row*9+col, or -1 when there is none (the reference returns None)
__global__ void __launch_bounds__(128)
find_king_kernel(const int8_t* __restrict__ boards, const int8_t* __restrict__ sides, int B, int8_t* __restrict__ out)
{
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= B) return;
    const int8_t* b = boards + (size_t)i * kSquares;
    const int side = sides[i];
    const int target = side == 1 ? 1 : -1, r0 = side == 1 ? 0 : 7;
    int found = -1;
    for (int r = r0 + 2; r >= r0; --r)
        for (int c = 5; c >= 3; --c)
            if (b[r * 9 + c] == target) found = r * 9 + c;      // descending scan: the first match in row-major order wins
    out[i] = (int8_t)found;
}

// cy_has_legal_moves (game_core.pyx:558-569) = _generate_moves(...) > 0, from the counts of a movegen launch
__global__ void __launch_bounds__(256) nonzero_kernel(const uint8_t* __restrict__ n, int B, uint8_t* __restrict__ out)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < B) out[i] = n[i] ? 1 : 0;
}

// get_state_for_nn (game.py:618-640) as bits: word w of a position holds plane-major bits 32w .. 32w+31, bit
// index = plane*90 + square (1350 bits in 43 words, padded to 44 = 176 bytes; 5400 bytes as float32)
__global__ void __launch_bounds__(256)
planes_bits_kernel(const int8_t* __restrict__ boards, const int8_t* __restrict__ sides, int B, uint32_t* __restrict__ out)
{
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    if (t >= (long long)B * kPlaneWords) return;
    const int i = (int)(t / kPlaneWords), w = (int)(t - (long long)i * kPlaneWords);
    const int8_t* b = boards + (size_t)i * kSquares;
    const int side = sides[i];
    uint32_t bits = 0;
    for (int k = 0; k < 32; ++k) {
        const int e = w * 32 + k;
        if (e >= 15 * kSquares) break;
        const int p = e / kSquares, sq = e - p * kSquares;
        bool on;
        if (p == 14) on = side == 1;
        else {
            const int v = b[sq] * side;
            on = (v > 0 ? v - 1 : (v < 0 ? 6 - v : -1)) == p;
        }
        bits |= (on ? 1u : 0u) << k;
    }
    out[t] = bits;
}

// ---- random playouts: one warp per game ----------------------------------------------------
constexpr int kPlayWarps = 4;

struct __align__(16) PlayoutSmem {
    int8_t board[kPlayWarps][kBoardPad];
    int8_t ring[kPlayWarps][kRing * kBoardPad];
    WarpScratch ws[kPlayWarps];
};

__device__ __forceinline__ void warp_init_board(int8_t* b)
{
    // game.py:139-159
    const int lane = lane_id();
    for (int sq = lane; sq < kBoardPad; sq += 32) {
        int r = sq / 9, c = sq % 9, v = 0;
        if (sq < kSquares) {
            const int back[9] = {5, 4, 3, 2, 1, 2, 3, 4, 5};
            if (r == 0) v = back[c];
            else if (r == 9) v = -back[c];
            else if (r == 2 && (c == 1 || c == 7)) v = 6;
            else if (r == 7 && (c == 1 || c == 7)) v = -6;
            else if (r == 3 && (c & 1) == 0) v = 7;
            else if (r == 6 && (c & 1) == 0) v = -7;
        }
        b[sq] = (int8_t)v;
    }
    warp_sync();
}

__global__ void __launch_bounds__(kPlayWarps * 32)
playout_kernel(uint64_t seed, int n_games, int8_t* __restrict__ boards, int8_t* __restrict__ sides,
               int32_t* __restrict__ n_positions, int8_t* __restrict__ winner, int* __restrict__ overflow)
{
    __shared__ PlayoutSmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kPlayWarps + warp;
    if (g >= n_games) return;
    int8_t* b = sm.board[warp];
    int8_t* ring = sm.ring[warp];
    WarpScratch& S = sm.ws[warp];
    warp_init_board(b);
    GameMeta gm{1, 0, 0};
    int ply = 0, w = 2;
    for (;; ++ply) {
        MovegenResult r = warp_movegen(b, gm.side, S);
        if (r.overflow && lane == 0) atomicAdd(overflow, 1);
        w = warp_game_over(b, ring, gm, r);
        const size_t slot = (size_t)g * XQ_MAX_PLIES + ply;
        for (int i = lane; i < kSquares; i += 32) boards[slot * kSquares + i] = b[i];
        if (lane == 0) sides[slot] = (int8_t)gm.side;
        if (w != 2 || ply + 1 >= XQ_MAX_PLIES) break;
        const uint64_t u = rng_u64(seed, (uint64_t)g, (uint64_t)ply, 0);
        const int pick = (int)(u % (uint64_t)min(r.n_legal, kMaxMoves));
        const int action = S.actions[pick];
        warp_sync();
        warp_make_move(b, ring, gm, action);
    }
    for (int p = ply + 1 + lane; p < XQ_MAX_PLIES; p += 32) sides[(size_t)g * XQ_MAX_PLIES + p] = 0;
    if (lane == 0) {
        n_positions[g] = ply + 1;
        winner[g] = (int8_t)w;
    }
}

}  // namespace xq

// =============================================================================================
// C ABI
// =============================================================================================
using namespace xq;

extern "C" int xq_version(void) { return 100; }

extern "C" const char* xq_last_error(const xq_ctx* ctx) { return ctx ? ctx->err : g_xq_last_error; }

extern "C" int xq_create(int device, xq_ctx** out)
{
    if (!out) return xq_fail(nullptr, XQ_ERR_ARG, "xq_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return xq_fail(nullptr, XQ_ERR_CUDA, "xq_create: no CUDA device (%s); this library has no CPU path",
                       cudaGetErrorString(e));
    if (device < 0 || device >= n) return xq_fail(nullptr, XQ_ERR_ARG, "xq_create: device %d out of range", device);
    xq_ctx* c = new xq_ctx();
    c->device = device;
    if (const char* e = getenv("XQ_MOVEGEN_IMPL")) c->movegen_impl = (e[0] == 'w' || e[0] == '0') ? 0 : 1;
    if (const char* e = getenv("XQ_NET_2CTA")) c->net_2cta = atoi(e) != 0;
    if (const char* e = getenv("XQ_NET_PDL")) c->net_pdl = atoi(e) != 0;
    if (const char* e = getenv("XQ_NET_SMALL")) c->net_small = atoi(e) != 0;
    if (const char* e = getenv("XQ_SP_GRAPH")) c->sp_graph = atoi(e) != 0;
    if (const char* e = getenv("XQ_TRAIN_PDL")) c->train_pdl = atoi(e) != 0;
    if (const char* e = getenv("XQ_NET_FORK")) c->net_fork = atoi(e) != 0;
    XQ_CUDA(c, cudaSetDevice(device));
    cudaDeviceProp prop;
    XQ_CUDA(c, cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    if (prop.major != 10)
        fprintf(stderr, "[xq_b200] warning: built for sm_100a, device is sm_%d%d\n", prop.major, prop.minor);
    XQ_CUDA(c, cudaMalloc(&c->d_overflow, 32 * sizeof(int)));     // [0] overflow counter, [1..16] task counters of the tpb kernel
    XQ_CUDA(c, cudaMemset(c->d_overflow, 0, 32 * sizeof(int)));
    XQ_CUDA(c, cudaEventCreate(&c->ev0));
    XQ_CUDA(c, cudaEventCreate(&c->ev1));
    *out = c;
    return XQ_OK;
}

extern "C" void xq_mcts_free_(xq_ctx*);
extern "C" void xq_net_free_(xq_ctx*);
extern "C" void xq_selfplay_free_(xq_ctx*);
extern "C" void xq_peer_free_(xq_ctx*);
extern "C" void xq_tnet_free_(xq_ctx*);

extern "C" void xq_destroy(xq_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    xq_selfplay_free_(c);
    xq_mcts_free_(c);
    xq_net_free_(c);
    xq_peer_free_(c);
    xq_tnet_free_(c);
    for (int i = 0; i < 2; ++i) {
        if (c->pipe[i]) cudaStreamDestroy(c->pipe[i]);
        if (c->d_stage[i]) cudaFree(c->d_stage[i]);
    }
    if (c->d_overflow) cudaFree(c->d_overflow);
    if (c->d_scratch) cudaFree(c->d_scratch);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    delete c;
}

extern "C" long long xq_launch_count(xq_ctx* c, int reset)
{
    long long v = c->launches;
    if (reset) c->launches = 0;
    return v;
}

extern "C" int xq_set_timing(xq_ctx* c, int enabled)
{
    c->timing = enabled != 0;
    c->ev_valid = false;
    return XQ_OK;
}

extern "C" float xq_last_kernel_ms(xq_ctx* c)
{
    if (!c->timing || !c->ev_valid) return -1.0f;
    float ms = -1.0f;
    if (cudaEventSynchronize(c->ev1) != cudaSuccess) return -1.0f;
    if (cudaEventElapsedTime(&ms, c->ev0, c->ev1) != cudaSuccess) return -1.0f;
    return ms;
}

static int movegen_grid(xq_ctx* c, bool with_planes, int B)
{
    int per_sm = 0;
    if (with_planes)
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, movegen_kernel<true>, kThreads, 0);
    else
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, movegen_kernel<false>, kThreads, 0);
    if (per_sm < 1) per_sm = 1;
    const int tiles = (B + kTile - 1) / kTile;
    int grid = c->sm_count * per_sm;       // persistent: a whole number of resident waves
    return grid < tiles ? grid : tiles;
}

extern "C" int xq_movegen_batch(xq_ctx* c, const int8_t* d_boards, const int8_t* d_sides, int B,
                                int16_t* d_actions, uint8_t* d_n_moves, uint8_t* d_in_check,
                                float* d_planes, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_movegen_batch: ctx is NULL");
    if (B < 0 || (B > 0 && (!d_boards || !d_sides || !d_actions || !d_n_moves || !d_in_check)))
        return xq_fail(c, XQ_ERR_ARG, "xq_movegen_batch: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    if (((uintptr_t)d_actions & 7) || (d_planes && ((uintptr_t)d_planes & 7)))
        return xq_fail(c, XQ_ERR_ARG, "xq_movegen_batch: actions/planes must be 8-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (c->movegen_impl == 1 && (d_planes == nullptr || ((uintptr_t)d_planes & 15) == 0)) {
        // second generation: one thread per board (xq_rules_tpb.h); a planes pointer that is not 16-byte aligned takes
        // the first-generation kernel (same outputs)
        const int smem = (int)sizeof(TpbSmem);
        if (!c->tpb_attr_set) {                  // per context: the attribute belongs to the context's device
            XQ_CUDA(c, cudaFuncSetAttribute(movegen_tpb_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            XQ_CUDA(c, cudaFuncSetAttribute(movegen_tpb_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            c->tpb_attr_set = true;
        }
        int per_sm = 0;
        if (d_planes)
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, movegen_tpb_kernel<true>, kTpbThreads, smem);
        else
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, movegen_tpb_kernel<false>, kTpbThreads, smem);
        if (per_sm < 1) per_sm = 1;
        const int tasks = (B + 31) / 32;
        const int ctas = (tasks + kTpbWarps - 1) / kTpbWarps;
        int grid = c->sm_count * per_sm;
        if (grid > ctas) grid = ctas;
        const int vec_ok = ((uintptr_t)d_boards & 15) == 0;
        // one task counter per launch in flight: a small ring, zeroed on the launch's own stream
        int* counter = c->d_overflow + 1 + (c->launches & 15);
        XQ_CUDA(c, cudaMemsetAsync(counter, 0, sizeof(int), s));
        {
            XqTimer tm(c, s);
            if (d_planes)
                movegen_tpb_kernel<true><<<grid, kTpbThreads, smem, s>>>(d_boards, d_sides, B, d_actions, d_n_moves,
                                                                         d_in_check, d_planes, c->d_overflow, counter, vec_ok);
            else
                movegen_tpb_kernel<false><<<grid, kTpbThreads, smem, s>>>(d_boards, d_sides, B, d_actions, d_n_moves,
                                                                          d_in_check, nullptr, c->d_overflow, counter, vec_ok);
        }
        c->launches += 1;
        XQ_CUDA(c, cudaGetLastError());
        return XQ_OK;
    }
    const int bulk_ok = (((uintptr_t)d_boards | (uintptr_t)d_sides) & 15) == 0;
    const int grid = movegen_grid(c, d_planes != nullptr, B);
    {
        XqTimer tm(c, s);
        if (d_planes)
            movegen_kernel<true><<<grid, kThreads, 0, s>>>(d_boards, d_sides, B, d_actions, d_n_moves,
                                                           d_in_check, d_planes, c->d_overflow, bulk_ok);
        else
            movegen_kernel<false><<<grid, kThreads, 0, s>>>(d_boards, d_sides, B, d_actions, d_n_moves,
                                                            d_in_check, nullptr, c->d_overflow, bulk_ok);
    }
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_set_movegen_impl(xq_ctx* c, int impl)
{
    if (!c || impl < 0 || impl > 1) return xq_fail(c, XQ_ERR_ARG, "xq_set_movegen_impl: impl must be 0 (warp) or 1 (thread)");
    const int prev = c->movegen_impl;
    c->movegen_impl = impl;
    return prev;
}

extern "C" int xq_overflow_count(xq_ctx* c, int reset)
{
    int v = 0;
    if (cudaMemcpy(&v, c->d_overflow, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    if (reset) cudaMemset(c->d_overflow, 0, sizeof(int));
    return v;
}

// ---- host-buffer pipeline -------------------------------------------------------------------
static int ensure_pipe(xq_ctx* c, size_t bytes)
{
    for (int i = 0; i < 2; ++i)
        if (!c->pipe[i]) XQ_CUDA(c, cudaStreamCreateWithFlags(&c->pipe[i], cudaStreamNonBlocking));
    if (c->stage_bytes < bytes) {
        for (int i = 0; i < 2; ++i) {
            if (c->d_stage[i]) XQ_CUDA(c, cudaFree(c->d_stage[i]));
            c->d_stage[i] = nullptr;
            XQ_CUDA(c, cudaMalloc(&c->d_stage[i], bytes));
        }
        c->stage_bytes = bytes;
    }
    return XQ_OK;
}

extern "C" int xq_planes_bits(xq_ctx* c, const int8_t* d_boards, const int8_t* d_sides, int B, uint32_t* d_bits, void* stream);

static int movegen_host_impl(xq_ctx* c, const int8_t* h_boards, const int8_t* h_sides, int B, int16_t* h_actions, uint8_t* h_n_moves,
                             uint8_t* h_in_check, float* h_planes, uint32_t* h_plane_bits);

extern "C" int xq_movegen_batch_host(xq_ctx* c, const int8_t* h_boards, const int8_t* h_sides, int B,
                                     int16_t* h_actions, uint8_t* h_n_moves, uint8_t* h_in_check,
                                     float* h_planes)
{
    return movegen_host_impl(c, h_boards, h_sides, B, h_actions, h_n_moves, h_in_check, h_planes, nullptr);
}

extern "C" int xq_movegen_batch_host_packed(xq_ctx* c, const int8_t* h_boards, const int8_t* h_sides, int B,
                                            int16_t* h_actions, uint8_t* h_n_moves, uint8_t* h_in_check,
                                            uint32_t* h_plane_bits)
{
    return movegen_host_impl(c, h_boards, h_sides, B, h_actions, h_n_moves, h_in_check, nullptr, h_plane_bits);
}

static int movegen_host_impl(xq_ctx* c, const int8_t* h_boards, const int8_t* h_sides, int B, int16_t* h_actions, uint8_t* h_n_moves,
                             uint8_t* h_in_check, float* h_planes, uint32_t* h_plane_bits)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_movegen_batch_host: ctx is NULL");
    if (B < 0 || (B > 0 && (!h_boards || !h_sides || !h_actions || !h_n_moves || !h_in_check)))
        return xq_fail(c, XQ_ERR_ARG, "xq_movegen_batch_host: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    XQ_CUDA(c, cudaSetDevice(c->device));
    // chunked so that H2D of chunk i+1, the kernel of chunk i and D2H of chunk i-1 overlap
    const int chunk = 65536;
    const size_t per_pos = 90 + 1 + 256 + 1 + 1 + (h_planes ? 5400 : 0);
    // staging layout per chunk (each region 256 B aligned)
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_boards = 0, o_sides = al(o_boards + (size_t)chunk * 90), o_act = al(o_sides + chunk),
                 o_n = al(o_act + (size_t)chunk * 256), o_chk = al(o_n + chunk), o_pl = al(o_chk + chunk),
                 total = o_pl + (h_planes ? (size_t)chunk * 5400 : (h_plane_bits ? (size_t)chunk * kPlaneWords * 4 : 0));
    (void)per_pos;
    int rc = ensure_pipe(c, total);
    if (rc) return rc;
    const int before = xq_overflow_count(c, 0);
    int k = 0;
    for (int off = 0; off < B; off += chunk, ++k) {
        const int n = (B - off < chunk) ? (B - off) : chunk;
        cudaStream_t s = c->pipe[k & 1];
        char* d = (char*)c->d_stage[k & 1];
        XQ_CUDA(c, cudaMemcpyAsync(d + o_boards, h_boards + (size_t)off * 90, (size_t)n * 90, cudaMemcpyHostToDevice, s));
        XQ_CUDA(c, cudaMemcpyAsync(d + o_sides, h_sides + off, (size_t)n, cudaMemcpyHostToDevice, s));
        rc = xq_movegen_batch(c, (const int8_t*)(d + o_boards), (const int8_t*)(d + o_sides), n,
                              (int16_t*)(d + o_act), (uint8_t*)(d + o_n), (uint8_t*)(d + o_chk),
                              h_planes ? (float*)(d + o_pl) : nullptr, s);
        if (rc) return rc;
        XQ_CUDA(c, cudaMemcpyAsync(h_actions + (size_t)off * 128, d + o_act, (size_t)n * 256, cudaMemcpyDeviceToHost, s));
        XQ_CUDA(c, cudaMemcpyAsync(h_n_moves + off, d + o_n, (size_t)n, cudaMemcpyDeviceToHost, s));
        XQ_CUDA(c, cudaMemcpyAsync(h_in_check + off, d + o_chk, (size_t)n, cudaMemcpyDeviceToHost, s));
        if (h_planes)
            XQ_CUDA(c, cudaMemcpyAsync(h_planes + (size_t)off * 1350, d + o_pl, (size_t)n * 5400, cudaMemcpyDeviceToHost, s));
        if (h_plane_bits) {
            rc = xq_planes_bits(c, (const int8_t*)(d + o_boards), (const int8_t*)(d + o_sides), n, (uint32_t*)(d + o_pl), s);
            if (rc) return rc;
            XQ_CUDA(c, cudaMemcpyAsync(h_plane_bits + (size_t)off * kPlaneWords, d + o_pl, (size_t)n * kPlaneWords * 4,
                                       cudaMemcpyDeviceToHost, s));
        }
    }
    XQ_CUDA(c, cudaStreamSynchronize(c->pipe[0]));
    XQ_CUDA(c, cudaStreamSynchronize(c->pipe[1]));
    const int after = xq_overflow_count(c, 0);
    if (after > before)
        return xq_fail(c, XQ_ERR_OVERFLOW, "xq_movegen_batch_host: %d position(s) exceeded %d legal moves",
                       after - before, XQ_MAX_MOVES);
    return XQ_OK;
}

extern "C" int xq_is_attacked_batch(xq_ctx* c, const int8_t* d_boards, const uint8_t* d_sq, const int8_t* d_by,
                                    int B, uint8_t* d_out, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_is_attacked_batch: ctx is NULL");
    if (B < 0 || (B > 0 && (!d_boards || !d_sq || !d_by || !d_out)))
        return xq_fail(c, XQ_ERR_ARG, "xq_is_attacked_batch: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    cudaStream_t s = (cudaStream_t)stream;
    {
        XqTimer tm(c, s);
        is_attacked_kernel<<<(B + kAtkThreads - 1) / kAtkThreads, kAtkThreads, 0, s>>>(d_boards, d_sq, d_by, B, d_out);
    }
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_is_attacked_batch_host(xq_ctx* c, const int8_t* h_boards, const uint8_t* h_sq,
                                         const int8_t* h_by, int B, uint8_t* h_out)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_is_attacked_batch_host: ctx is NULL");
    if (B < 0 || (B > 0 && (!h_boards || !h_sq || !h_by || !h_out)))
        return xq_fail(c, XQ_ERR_ARG, "xq_is_attacked_batch_host: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    XQ_CUDA(c, cudaSetDevice(c->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_b = 0, o_sq = al((size_t)B * 90), o_by = al(o_sq + B), o_out = al(o_by + B), total = o_out + B;
    int rc = ensure_pipe(c, total);
    if (rc) return rc;
    cudaStream_t s = c->pipe[0];
    char* d = (char*)c->d_stage[0];
    XQ_CUDA(c, cudaMemcpyAsync(d + o_b, h_boards, (size_t)B * 90, cudaMemcpyHostToDevice, s));
    XQ_CUDA(c, cudaMemcpyAsync(d + o_sq, h_sq, (size_t)B, cudaMemcpyHostToDevice, s));
    XQ_CUDA(c, cudaMemcpyAsync(d + o_by, h_by, (size_t)B, cudaMemcpyHostToDevice, s));
    rc = xq_is_attacked_batch(c, (const int8_t*)(d + o_b), (const uint8_t*)(d + o_sq), (const int8_t*)(d + o_by), B,
                              (uint8_t*)(d + o_out), s);
    if (rc) return rc;
    XQ_CUDA(c, cudaMemcpyAsync(h_out, d + o_out, (size_t)B, cudaMemcpyDeviceToHost, s));
    XQ_CUDA(c, cudaStreamSynchronize(s));
    return XQ_OK;
}

extern "C" int xq_move_is_legal_batch(xq_ctx* c, const int8_t* d_boards, const uint8_t* d_from, const uint8_t* d_to, const int8_t* d_sides,
                                      int B, uint8_t* d_out, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_move_is_legal_batch: ctx is NULL");
    if (B < 0 || (B > 0 && (!d_boards || !d_from || !d_to || !d_sides || !d_out)))
        return xq_fail(c, XQ_ERR_ARG, "xq_move_is_legal_batch: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    move_is_legal_kernel<<<(B + kAtkThreads - 1) / kAtkThreads, kAtkThreads, 0, (cudaStream_t)stream>>>(d_boards, d_from, d_to, d_sides, B, d_out);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_find_king_batch(xq_ctx* c, const int8_t* d_boards, const int8_t* d_sides, int B, int8_t* d_king_sq, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_find_king_batch: ctx is NULL");
    if (B < 0 || (B > 0 && (!d_boards || !d_sides || !d_king_sq))) return xq_fail(c, XQ_ERR_ARG, "xq_find_king_batch: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    find_king_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_boards, d_sides, B, d_king_sq);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

static int ensure_pipe(xq_ctx* c, size_t bytes);

// grow-only device scratch of the context (outputs a caller did not ask for)
static int ensure_scratch(xq_ctx* c, size_t bytes)
{
    if (c->scratch_bytes >= bytes) return XQ_OK;
    if (c->d_scratch) XQ_CUDA(c, cudaFree(c->d_scratch));
    c->d_scratch = nullptr;
    c->scratch_bytes = 0;
    XQ_CUDA(c, cudaMalloc(&c->d_scratch, bytes));
    c->scratch_bytes = bytes;
    return XQ_OK;
}

extern "C" int xq_has_legal_moves_batch(xq_ctx* c, const int8_t* d_boards, const int8_t* d_sides, int B, uint8_t* d_has, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_has_legal_moves_batch: ctx is NULL");
    if (B < 0 || (B > 0 && (!d_boards || !d_sides || !d_has))) return xq_fail(c, XQ_ERR_ARG, "xq_has_legal_moves_batch: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    XQ_CUDA(c, cudaSetDevice(c->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_n = al((size_t)B * 256), o_chk = al(o_n + B);
    int rc = ensure_scratch(c, o_chk + B);
    if (rc) return rc;
    char* d = (char*)c->d_scratch;
    rc = xq_movegen_batch(c, d_boards, d_sides, B, (int16_t*)d, (uint8_t*)(d + o_n), (uint8_t*)(d + o_chk), nullptr, stream);
    if (rc) return rc;
    nonzero_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)(d + o_n), B, d_has);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

// host-pointer forms of the two queries above (the per-position calls a game.py-level binding makes)
static int seam_query_host(xq_ctx* c, const int8_t* h_boards, const int8_t* h_sides, int B, void* h_out, bool king)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_*_batch_host: ctx is NULL");
    if (B < 0 || (B > 0 && (!h_boards || !h_sides || !h_out))) return xq_fail(c, XQ_ERR_ARG, "xq_*_batch_host: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    XQ_CUDA(c, cudaSetDevice(c->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_s = al((size_t)B * 90), o_out = al(o_s + B);
    int rc = ensure_pipe(c, o_out + B);
    if (rc) return rc;
    cudaStream_t s = c->pipe[0];
    char* d = (char*)c->d_stage[0];
    XQ_CUDA(c, cudaMemcpyAsync(d, h_boards, (size_t)B * 90, cudaMemcpyHostToDevice, s));
    XQ_CUDA(c, cudaMemcpyAsync(d + o_s, h_sides, (size_t)B, cudaMemcpyHostToDevice, s));
    rc = king ? xq_find_king_batch(c, (const int8_t*)d, (const int8_t*)(d + o_s), B, (int8_t*)(d + o_out), s)
              : xq_has_legal_moves_batch(c, (const int8_t*)d, (const int8_t*)(d + o_s), B, (uint8_t*)(d + o_out), s);
    if (rc) return rc;
    XQ_CUDA(c, cudaMemcpyAsync(h_out, d + o_out, (size_t)B, cudaMemcpyDeviceToHost, s));
    XQ_CUDA(c, cudaStreamSynchronize(s));
    return XQ_OK;
}
extern "C" int xq_move_is_legal_batch_host(xq_ctx* c, const int8_t* h_boards, const uint8_t* h_from, const uint8_t* h_to, const int8_t* h_sides,
                                           int B, uint8_t* h_out)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_move_is_legal_batch_host: ctx is NULL");
    if (B < 0 || (B > 0 && (!h_boards || !h_from || !h_to || !h_sides || !h_out)))
        return xq_fail(c, XQ_ERR_ARG, "xq_move_is_legal_batch_host: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    XQ_CUDA(c, cudaSetDevice(c->device));
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t o_f = al((size_t)B * 90), o_t = al(o_f + B), o_s = al(o_t + B), o_out = al(o_s + B);
    int rc = ensure_pipe(c, o_out + B);
    if (rc) return rc;
    cudaStream_t s = c->pipe[0];
    char* d = (char*)c->d_stage[0];
    XQ_CUDA(c, cudaMemcpyAsync(d, h_boards, (size_t)B * 90, cudaMemcpyHostToDevice, s));
    XQ_CUDA(c, cudaMemcpyAsync(d + o_f, h_from, (size_t)B, cudaMemcpyHostToDevice, s));
    XQ_CUDA(c, cudaMemcpyAsync(d + o_t, h_to, (size_t)B, cudaMemcpyHostToDevice, s));
    XQ_CUDA(c, cudaMemcpyAsync(d + o_s, h_sides, (size_t)B, cudaMemcpyHostToDevice, s));
    rc = xq_move_is_legal_batch(c, (const int8_t*)d, (const uint8_t*)(d + o_f), (const uint8_t*)(d + o_t), (const int8_t*)(d + o_s), B,
                                (uint8_t*)(d + o_out), s);
    if (rc) return rc;
    XQ_CUDA(c, cudaMemcpyAsync(h_out, d + o_out, (size_t)B, cudaMemcpyDeviceToHost, s));
    XQ_CUDA(c, cudaStreamSynchronize(s));
    return XQ_OK;
}

extern "C" int xq_find_king_batch_host(xq_ctx* c, const int8_t* h_boards, const int8_t* h_sides, int B, int8_t* h_king_sq)
{
    return seam_query_host(c, h_boards, h_sides, B, h_king_sq, true);
}
extern "C" int xq_has_legal_moves_batch_host(xq_ctx* c, const int8_t* h_boards, const int8_t* h_sides, int B, uint8_t* h_has)
{
    return seam_query_host(c, h_boards, h_sides, B, h_has, false);
}

extern "C" int xq_planes_bits(xq_ctx* c, const int8_t* d_boards, const int8_t* d_sides, int B, uint32_t* d_bits, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_planes_bits: ctx is NULL");
    if (B < 0 || (B > 0 && (!d_boards || !d_sides || !d_bits))) return xq_fail(c, XQ_ERR_ARG, "xq_planes_bits: bad arguments (B=%d)", B);
    if (B == 0) return XQ_OK;
    const long long n = (long long)B * kPlaneWords;
    planes_bits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_boards, d_sides, B, d_bits);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_random_playouts(xq_ctx* c, uint64_t seed, int n_games, int8_t* d_boards, int8_t* d_sides,
                                  int32_t* d_n_positions, int8_t* d_winner, void* stream)
{
    if (!c) return xq_fail(nullptr, XQ_ERR_ARG, "xq_random_playouts: ctx is NULL");
    if (n_games < 0 || (n_games > 0 && (!d_boards || !d_sides || !d_n_positions || !d_winner)))
        return xq_fail(c, XQ_ERR_ARG, "xq_random_playouts: bad arguments");
    if (n_games == 0) return XQ_OK;
    cudaStream_t s = (cudaStream_t)stream;
    {
        XqTimer tm(c, s);
        playout_kernel<<<(n_games + kPlayWarps - 1) / kPlayWarps, kPlayWarps * 32, 0, s>>>(
            seed, n_games, d_boards, d_sides, d_n_positions, d_winner, c->d_overflow);
    }
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}
