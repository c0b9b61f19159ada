// xq_rules.cuh -- warp-cooperative xiangqi rules for sm_100a.
//
// One warp owns one board (90 int8 cells staged in shared memory).  The functions here are
// the device-side replacement of the reference's only native component,
// training/cython_engine/game_core.pyx, plus the termination rules of training/game.py:
//
//   warp_movegen      <- _generate_moves  game_core.pyx:262-486 (ordered legal list)
//   legal_after_move  <- _is_move_legal   game_core.pyx:209-252
//   attacked_sq       <- _is_attacked     game_core.pyx:104-189
//   warp_find_kings   <- _find_king       game_core.pyx:78-101 (palace-only scan)
//   warp_game_over    <- is_game_over     game.py:565-616
//
// Ordering contract: moves come out in (from-square row-major, per-piece
// direction order, ray step) order, because the MCTS child order and every first-max
// tie-break downstream depend on it.  The generator is split in two warp-parallel phases:
//   A. pseudo-legal targets: one lane per (own piece, direction slot) task, two-pass
//      count/scan/write so the compact list keeps the reference order;
//   B. legality: one lane per pseudo-legal move (make-move overlay on the shared board, no
//      per-lane board copy), ballot-compaction keeps the order.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace xq {

constexpr int kSquares = 90;
constexpr int kMaxMoves = 128;    // output slots per position (max seen in play: 74)
constexpr int kMaxPseudo = 192;   // pseudo-legal scratch (orthodox upper bound: 119)
// ---- warp collectives ------------------------------------------------------------------------
// Every collective of this library goes through these NON-INLINED helpers.  Reason (measured on
// B200 with CUDA 12.9): when the collectives were inlined after the divergent move-generation
// switch, ptxas trusted the closing BSYNC.RECONVERGENT and emitted SHFL/VOTE with no WARPSYNC in
// front; some lanes (the late arrivals: cannon walks, far switch arms) were still running apart
// there -- __activemask() showed groups like {22},{28,30},{29},{31} -- and the prefix scan of
// warp_movegen combined stale values (moves written at wrong offsets; caught by the MCTS golden
// test).  A call boundary makes ptxas assume nothing about convergence, so each helper starts
// with a real WARPSYNC.ALL.
constexpr unsigned kFullMask = 0xffffffffu;
static __device__ __noinline__ void warp_sync() { __syncwarp(); }
static __device__ __noinline__ unsigned warp_ballot(bool p)
{
    __syncwarp();
    return __ballot_sync(kFullMask, p);
}
static __device__ __noinline__ bool warp_any(bool p)
{
    __syncwarp();
    return __any_sync(kFullMask, p) != 0;
}
static __device__ __noinline__ int warp_bcast(int v, int src)
{
    __syncwarp();
    return __shfl_sync(kFullMask, v, src);
}
static __device__ __noinline__ float warp_bcast_f(float v, int src)
{
    __syncwarp();
    return __shfl_sync(kFullMask, v, src);
}
static __device__ __noinline__ int warp_sum(int v)
{
    __syncwarp();
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}
static __device__ __noinline__ float warp_sum_f(float v)
{
    __syncwarp();
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}
static __device__ __noinline__ double warp_sum_d(double v)
{
    __syncwarp();
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}
static __device__ __noinline__ float warp_max_f(float v)
{
    __syncwarp();
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}
// inclusive prefix sum over lanes; *total receives the warp total
static __device__ __noinline__ int warp_incl_scan(int v, int* total)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(kFullMask, v, o);
        if (lane >= o) v += t;
    }
    *total = __shfl_sync(kFullMask, v, 31);
    return v;
}
// argmax with first-maximum tie-break (smaller index wins among equal scores)
static __device__ __noinline__ int warp_argmax_first(double best, int best_i)
{
    __syncwarp();
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ob = __shfl_xor_sync(kFullMask, best, o);
        const int oi = __shfl_xor_sync(kFullMask, best_i, o);
        if (ob > best || (ob == best && oi < best_i)) {
            best = ob;
            best_i = oi;
        }
    }
    return best_i;
}

// 16 bytes from lane `src` to every lane (non-inlined, __syncwarp first: see the note on warp collectives in DESIGN.md)
static __device__ __noinline__ uint4 warp_bcast16(uint4 v, int src)
{
    __syncwarp();
    v.x = __shfl_sync(kFullMask, v.x, src);
    v.y = __shfl_sync(kFullMask, v.y, src);
    v.z = __shfl_sync(kFullMask, v.z, src);
    v.w = __shfl_sync(kFullMask, v.w, src);
    return v;
}

struct alignas(16) WarpScratch {
    uint16_t pm[kMaxPseudo];      // pseudo-legal moves in reference order: from << 7 | to
    int16_t actions[kMaxMoves];   // compacted legal actions, unused slots = -1
    uint8_t own_sq[96];           // own-piece squares in row-major order (phase A task table)
    uint16_t rowocc[10];          // occupancy of each row (bit c) and
    uint16_t colocc[10];          // of each column (bit r) of the un-moved board: ray scans become bit scans
    uint8_t kingto[4];            // target square of the palace king's move in direction d (0xff: none), 4-byte aligned
    uint8_t pad_[4];
};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool is_own(int p, int side) { return side == 1 ? p > 0 : p < 0; }
__device__ __forceinline__ bool is_foe(int p, int side) { return side == 1 ? p < 0 : p > 0; }
__device__ __forceinline__ bool can_land(int p, int side) { return p == 0 || is_foe(p, side); }

// board cell as seen after the overlay move from->to (from < 0: no move)
__device__ __forceinline__ int cell_after(const int8_t* b, int sq, int from, int to, int mover)
{
    int v = b[sq];
    v = (sq == from) ? 0 : v;
    v = (sq == to) ? mover : v;
    return v;
}

// game_core.pyx:104-189 evaluated on the overlaid board.  The rook/king ray test and the
// cannon ray test of the reference walk the same four rays; one walk sees the first piece
// (rook or king attacks) and the second piece (cannon attacks).
__device__ __forceinline__ bool attacked_sq(const int8_t* b, int kr, int kc, int by, int from, int to,
                                            int mover)
{
    const int rook = 5 * by, cannon = 6 * by, horse = 4 * by, pawn = 7 * by, king = by;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const int dr = (d == 0) ? -1 : (d == 1 ? 1 : 0);
        const int dc = (d == 2) ? -1 : (d == 3 ? 1 : 0);
        int r = kr + dr, c = kc + dc;
        bool screen = false;
        while (r >= 0 && r < 10 && c >= 0 && c < 9) {
            int p = cell_after(b, r * 9 + c, from, to, mover);
            if (p != 0) {
                if (!screen) {
                    if (p == rook || p == king) return true;
                    screen = true;
                } else {
                    if (p == cannon) return true;
                    break;
                }
            }
            r += dr;
            c += dc;
        }
    }
    // knights: leg is adjacent to the knight, on the knight's long axis (pyx:156-169)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int jr = (i < 4) ? ((i < 2) ? -2 : 2) : ((i < 6) ? -1 : 1);
        const int jc = (i < 4) ? ((i & 1) ? 1 : -1) : ((i & 1) ? 2 : -2);
        int nr = kr + jr, nc = kc + jc;
        if (nr < 0 || nr >= 10 || nc < 0 || nc >= 9) continue;
        if (cell_after(b, nr * 9 + nc, from, to, mover) != horse) continue;
        // move from knight to target = (-jr, -jc); leg = knight + half of the long component
        int lr = nr, lc = nc;
        if (jr == 2 || jr == -2) lr = nr - jr / 2; else lc = nc - jc / 2;
        if (cell_after(b, lr * 9 + lc, from, to, mover) == 0) return true;
    }
    // pawns (pyx:172-187)
    if (by == 1) {
        if (kr - 1 >= 0 && cell_after(b, (kr - 1) * 9 + kc, from, to, mover) == pawn) return true;
        if (kr >= 5) {
            if (kc - 1 >= 0 && cell_after(b, kr * 9 + kc - 1, from, to, mover) == pawn) return true;
            if (kc + 1 < 9 && cell_after(b, kr * 9 + kc + 1, from, to, mover) == pawn) return true;
        }
    } else {
        if (kr + 1 < 10 && cell_after(b, (kr + 1) * 9 + kc, from, to, mover) == pawn) return true;
        if (kr <= 4) {
            if (kc - 1 >= 0 && cell_after(b, kr * 9 + kc - 1, from, to, mover) == pawn) return true;
            if (kc + 1 < 9 && cell_after(b, kr * 9 + kc + 1, from, to, mover) == pawn) return true;
        }
    }
    return false;
}

// Same predicate as attacked_sq, with the four ray walks replaced by bit scans of the king's row and
// column occupancy (overlay applied to the two masks): first blocker = rook/king test, second = cannon test.
__device__ __forceinline__ bool attacked_sq_occ(const int8_t* b, const uint16_t* rowocc, const uint16_t* colocc, int kr,
                                                int kc, int by, int from, int to, int mover)
{
    const int rook = 5 * by, cannon = 6 * by, horse = 4 * by, pawn = 7 * by, king = by;
    unsigned R = rowocc[kr], Cm = colocc[kc];
    if (from >= 0) {
        const int fr = from / 9, fc = from - fr * 9, tr = to / 9, tc = to - tr * 9;
        if (fr == kr) R &= ~(1u << fc);
        if (fc == kc) Cm &= ~(1u << fr);
        if (tr == kr) R |= 1u << tc;
        if (tc == kc) Cm |= 1u << tr;
    }
    // toward smaller index: highest set bit below the king; toward larger: lowest set bit above
    {
        unsigned m = R & ((1u << kc) - 1u);
        if (m) {
            int c1 = 31 - __clz(m);
            int p1 = cell_after(b, kr * 9 + c1, from, to, mover);
            if (p1 == rook || p1 == king) return true;
            m ^= 1u << c1;
            if (m && cell_after(b, kr * 9 + 31 - __clz(m), from, to, mover) == cannon) return true;
        }
        m = R >> (kc + 1);
        if (m) {
            int c1 = kc + __ffs(m);
            int p1 = cell_after(b, kr * 9 + c1, from, to, mover);
            if (p1 == rook || p1 == king) return true;
            m &= m - 1u;
            if (m && cell_after(b, kr * 9 + kc + __ffs(m), from, to, mover) == cannon) return true;
        }
        m = Cm & ((1u << kr) - 1u);
        if (m) {
            int r1 = 31 - __clz(m);
            int p1 = cell_after(b, r1 * 9 + kc, from, to, mover);
            if (p1 == rook || p1 == king) return true;
            m ^= 1u << r1;
            if (m && cell_after(b, (31 - __clz(m)) * 9 + kc, from, to, mover) == cannon) return true;
        }
        m = Cm >> (kr + 1);
        if (m) {
            int r1 = kr + __ffs(m);
            int p1 = cell_after(b, r1 * 9 + kc, from, to, mover);
            if (p1 == rook || p1 == king) return true;
            m &= m - 1u;
            if (m && cell_after(b, (kr + __ffs(m)) * 9 + kc, from, to, mover) == cannon) return true;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int jr = (i < 4) ? ((i < 2) ? -2 : 2) : ((i < 6) ? -1 : 1);
        const int jc = (i < 4) ? ((i & 1) ? 1 : -1) : ((i & 1) ? 2 : -2);
        int nr = kr + jr, nc = kc + jc;
        if (nr < 0 || nr >= 10 || nc < 0 || nc >= 9) continue;
        if (cell_after(b, nr * 9 + nc, from, to, mover) != horse) continue;
        int lr = nr, lc = nc;
        if (jr == 2 || jr == -2) lr = nr - jr / 2; else lc = nc - jc / 2;
        if (cell_after(b, lr * 9 + lc, from, to, mover) == 0) return true;
    }
    if (by == 1) {
        if (kr - 1 >= 0 && cell_after(b, (kr - 1) * 9 + kc, from, to, mover) == pawn) return true;
        if (kr >= 5) {
            if (kc - 1 >= 0 && cell_after(b, kr * 9 + kc - 1, from, to, mover) == pawn) return true;
            if (kc + 1 < 9 && cell_after(b, kr * 9 + kc + 1, from, to, mover) == pawn) return true;
        }
    } else {
        if (kr + 1 < 10 && cell_after(b, (kr + 1) * 9 + kc, from, to, mover) == pawn) return true;
        if (kr <= 4) {
            if (kc - 1 >= 0 && cell_after(b, kr * 9 + kc - 1, from, to, mover) == pawn) return true;
            if (kc + 1 < 9 && cell_after(b, kr * 9 + kc + 1, from, to, mover) == pawn) return true;
        }
    }
    return false;
}

// palace square index 0..8 -> board square (rows r0..r0+2, cols 3..5, row-major like pyx:93-98)
__device__ __forceinline__ int palace_sq(int side, int i) { return ((side == 1 ? 0 : 7) + i / 3) * 9 + 3 + i % 3; }

// first king of `side` in its own palace on the overlaid board, or -1
__device__ __forceinline__ int find_king_after(const int8_t* b, int side, int from, int to, int mover)
{
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        int sq = palace_sq(side, i);
        if (cell_after(b, sq, from, to, mover) == side) return sq;
    }
    return -1;
}

struct KingInfo {
    int own_sq, foe_sq;       // first king in each palace on the un-moved board (-1: none)
    int own_cnt, foe_cnt;     // kings standing in each palace (1/1 on any orthodox board)
};

// Warp-parallel palace scan: lanes 0-8 own palace, 9-17 enemy palace.
__device__ __forceinline__ KingInfo warp_find_kings(const int8_t* b, int side)
{
    const int lane = lane_id();
    bool hit = false;
    if (lane < 9) hit = b[palace_sq(side, lane)] == side;
    else if (lane < 18) hit = b[palace_sq(-side, lane - 9)] == -side;
    unsigned m = warp_ballot(hit);
    unsigned own = m & 0x1ffu, foe = (m >> 9) & 0x1ffu;
    KingInfo k;
    k.own_sq = own ? palace_sq(side, __ffs(own) - 1) : -1;
    k.foe_sq = foe ? palace_sq(-side, __ffs(foe) - 1) : -1;
    k.own_cnt = __popc(own);
    k.foe_cnt = __popc(foe);
    return k;
}

// game_core.pyx:209-252.  from < 0 probes the current board ("null move") without the
// flying-general clause, i.e. plain is_attacked(own king) -- used for the in-check flag.
__device__ __forceinline__ bool legal_after_move(const int8_t* b, const uint16_t* rowocc, const uint16_t* colocc, int side,
                                                 const KingInfo& ki, int from, int to)
{
    const int mover = b[from];
    int k, e;
    if (ki.own_cnt == 1 && ki.foe_cnt <= 1 && (mover != side || from == ki.own_sq)) {
        // orthodox fast path: only a king move relocates the own king, only a capture on the
        // enemy king square removes the enemy king
        k = (from == ki.own_sq) ? to : ki.own_sq;
        e = (to == ki.foe_sq) ? -1 : ki.foe_sq;
        if (from == ki.own_sq) {      // the king must still stand in its palace to be "found"
            int r = to / 9, c = to % 9;
            bool in_palace = c >= 3 && c <= 5 && (side == 1 ? r <= 2 : r >= 7);
            if (!in_palace) k = -1;
        }
    } else {
        k = find_king_after(b, side, from, to, mover);
        e = find_king_after(b, -side, from, to, mover);
    }
    if (k < 0) return false;
    const int kr = k / 9, kc = k % 9;
    if (e >= 0 && e % 9 == kc) {
        // flying general: no piece strictly between the kings on their common file (overlaid column mask)
        const int er = e / 9;
        const int lo = min(kr, er), hi = max(kr, er);
        unsigned Cm = colocc[kc];
        const int fr = from / 9, fc = from - fr * 9, tr = to / 9, tc = to - tr * 9;
        if (fc == kc) Cm &= ~(1u << fr);
        if (tc == kc) Cm |= 1u << tr;
        const unsigned between = ((1u << hi) - 1u) & ~((1u << (lo + 1)) - 1u);
        if ((Cm & between) == 0) return false;
    }
    return !attacked_sq_occ(b, rowocc, colocc, kr, kc, -side, from, to, mover);
}

// ---- phase A tables ------------------------------------------------------------------------------
// Leaper moves (king, advisor, elephant, knight, pawn) of one (kind, direction slot d) task, two entries
// per task (only the knight uses the second: slot d covers KNIGHT_MOVES[2d], [2d+1], pyx:31-39).
// Entry: byte0 flags, byte1 dr, byte2 dc, byte3 leg offset in squares (0 = no leg).
//   flags: 1 valid, 2 target must be in the mover's palace (king pyx:287-304, advisor pyx:307-326: the Cython
//   advisor only checks the palace box), 4 target must be on the mover's half (elephant pyx:329-349),
//   8 dr is multiplied by side (pawn forward, pyx:434-484), 16 only after crossing the river (pawn sideways).
// One table-driven path replaces a 5-way divergent switch: a warp iteration holds tasks of every kind.
#define XQ_LEAP(fl, dr, dc, leg) \
    ((uint32_t)(fl) | ((uint32_t)(uint8_t)(int8_t)(dr) << 8) | ((uint32_t)(uint8_t)(int8_t)(dc) << 16) | \
     ((uint32_t)(uint8_t)(int8_t)(leg) << 24))
__device__ const uint2 kLeapTable[32] = {
    // kind 0 (empty)
    {0, 0}, {0, 0}, {0, 0}, {0, 0},
    // kind 1 king: DIRECTIONS pyx:42-46 up, down, left, right
    {XQ_LEAP(3, -1, 0, 0), 0}, {XQ_LEAP(3, 1, 0, 0), 0}, {XQ_LEAP(3, 0, -1, 0), 0}, {XQ_LEAP(3, 0, 1, 0), 0},
    // kind 2 advisor: (dr, dc) in (-1,-1), (-1,1), (1,-1), (1,1)
    {XQ_LEAP(3, -1, -1, 0), 0}, {XQ_LEAP(3, -1, 1, 0), 0}, {XQ_LEAP(3, 1, -1, 0), 0}, {XQ_LEAP(3, 1, 1, 0), 0},
    // kind 3 elephant: two steps diagonally, eye = one step
    {XQ_LEAP(5, -2, -2, -10), 0}, {XQ_LEAP(5, -2, 2, -8), 0}, {XQ_LEAP(5, 2, -2, 8), 0}, {XQ_LEAP(5, 2, 2, 10), 0},
    // kind 4 knight: KNIGHT_MOVES[2d], [2d+1] with the leg next to the knight on its long axis
    {XQ_LEAP(1, -2, -1, -9), XQ_LEAP(1, -2, 1, -9)}, {XQ_LEAP(1, 2, -1, 9), XQ_LEAP(1, 2, 1, 9)},
    {XQ_LEAP(1, -1, -2, -1), XQ_LEAP(1, -1, 2, 1)}, {XQ_LEAP(1, 1, -2, -1), XQ_LEAP(1, 1, 2, 1)},
    // kinds 5, 6 rook / cannon: slider path
    {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {0, 0},
    // kind 7 pawn: forward, left, right
    {XQ_LEAP(9, 1, 0, 0), 0}, {XQ_LEAP(17, 0, -1, 0), 0}, {XQ_LEAP(17, 0, 1, 0), 0}, {0, 0},
};

// one leaper entry -> target square or -1
__device__ __forceinline__ int leap_target(const int8_t* b, int side, int sq, int r, int c, uint32_t e)
{
    if (!(e & 1u)) return -1;
    int dr = (int)(int8_t)(e >> 8), dc = (int)(int8_t)(e >> 16), leg = (int)(int8_t)(e >> 24);
    if (e & 8u) dr *= side;
    const int nr = r + dr, nc = c + dc;
    bool ok = (unsigned)nr < 10u && (unsigned)nc < 9u;
    if (e & 2u) ok = ok && nc >= 3 && nc <= 5 && (side == 1 ? nr <= 2 : nr >= 7);
    if (e & 4u) ok = ok && (side == 1 ? nr <= 4 : nr >= 5);
    if (e & 16u) ok = ok && (side == 1 ? r >= 5 : r <= 4);
    if (!ok) return -1;
    if (leg != 0 && b[sq + leg] != 0) return -1;
    const int t = nr * 9 + nc;
    return (b[t] * side <= 0) ? t : -1;       // _can_move_to: empty or enemy
}

// Pseudo-legal targets of one (piece, direction slot) task in reference order, as a run plus one extra target:
// squares t0, t0+step, ... (n_run of them) followed by x (if >= 0).  Leapers: run = first table entry (0/1
// squares), x = second entry.  Rook pyx:370-396 / cannon pyx:399-431: run = the empty squares up to the first
// blocker (a bit scan of the row / column occupancy), x = the capture (first blocker for the rook, second for
// the cannon) when it is an enemy piece.
__device__ __forceinline__ int gen_task_run(const int8_t* b, const uint16_t* rowocc, const uint16_t* colocc, int side,
                                            int sq, int kind, int d, int& t0, int& step, int& x)
{
    const int r = sq / 9, c = sq - r * 9;
    int n_run = 0;
    t0 = 0;
    step = 0;
    x = -1;
    if (kind == 5 || kind == 6) {
        const bool along_row = d >= 2;
        const unsigned line = along_row ? rowocc[r] : colocc[c];
        const int pos = along_row ? c : r;
        const int len = along_row ? 9 : 10;
        const int stride = along_row ? 1 : 9;
        const bool neg = (d & 1) == 0;                    // d = 0 up, 2 left: toward smaller index
        int first = -1, second = -1;
        if (neg) {
            unsigned m = line & ((1u << pos) - 1u);
            if (m) {
                first = 31 - __clz(m);
                m ^= 1u << first;
                if (m) second = 31 - __clz(m);
            }
        } else {
            unsigned m = line >> (pos + 1);
            if (m) {
                first = pos + __ffs(m);
                m &= m - 1u;
                if (m) second = pos + __ffs(m);
            }
        }
        const int end = first >= 0 ? first : (neg ? -1 : len);
        n_run = neg ? pos - 1 - end : end - pos - 1;
        step = neg ? -stride : stride;
        t0 = sq + step;
        const int cap = kind == 5 ? first : second;
        if (cap >= 0) {
            const int csq = sq + (cap - pos) * stride;
            if (is_foe(b[csq], side)) x = csq;
        }
    } else {
        const uint2 e = kLeapTable[(kind <= 7 ? kind : 0) * 4 + d];
        const int a = leap_target(b, side, sq, r, c, e.x);
        if (a >= 0) {
            t0 = a;
            n_run = 1;
        }
        if (e.y) x = leap_target(b, side, sq, r, c, e.y);
    }
    return n_run;
}

// ---- attack tests used by the legality phase ---------------------------------------------------------
// Both directions of one line (the king's row or column) on the overlaid board: first piece = rook/king test,
// second piece = cannon test (pyx:104-153).  line = occupancy with the overlay applied, pos = the king's index
// on it, sq0 = square of index 0, stride = squares per index.
__device__ __forceinline__ bool line_attacked(const int8_t* b, unsigned line, int pos, int sq0, int stride, int by,
                                              int from, int to, int mover)
{
    const int rook = 5 * by, cannon = 6 * by, king = by;
    unsigned m = line & ((1u << pos) - 1u);
    if (m) {
        const int i1 = 31 - __clz(m);
        const int p1 = cell_after(b, sq0 + i1 * stride, from, to, mover);
        if (p1 == rook || p1 == king) return true;
        m ^= 1u << i1;
        if (m && cell_after(b, sq0 + (31 - __clz(m)) * stride, from, to, mover) == cannon) return true;
    }
    m = line >> (pos + 1);
    if (m) {
        const int i1 = pos + __ffs(m);
        const int p1 = cell_after(b, sq0 + i1 * stride, from, to, mover);
        if (p1 == rook || p1 == king) return true;
        m &= m - 1u;
        if (m && cell_after(b, sq0 + (pos + __ffs(m)) * stride, from, to, mover) == cannon) return true;
    }
    return false;
}

// _is_attacked (pyx:104-189) spread over the 16 lanes of a half warp: sub-lanes 0-3 one ray each, 4-11 one knight
// origin each, 12-14 the three pawn origins.  Each half warp answers its own query (kr, kc, overlay from/to/mover;
// from < 0: no overlay); the returned ballot has the low / high 16 bits set where a half found an attacker.
__device__ __forceinline__ unsigned warp_attacked_pair(const int8_t* b, const uint16_t* rowocc, const uint16_t* colocc,
                                                       bool active, int kr, int kc, int by, int from, int to, int mover)
{
    const int l = lane_id() & 15;
    bool hit = false;
    if (active) {
        if (l < 4) {
            const bool col = l < 2;
            unsigned line = col ? colocc[kc] : rowocc[kr];
            const int pos = col ? kr : kc;
            const int sq0 = col ? kc : kr * 9;
            const int stride = col ? 9 : 1;
            if (from >= 0) {
                const int fr = from / 9, fc = from - fr * 9, tr = to / 9, tc = to - tr * 9;
                if (col) {
                    if (fc == kc) line &= ~(1u << fr);
                    if (tc == kc) line |= 1u << tr;
                } else {
                    if (fr == kr) line &= ~(1u << fc);
                    if (tr == kr) line |= 1u << tc;
                }
            }
            const int rook = 5 * by, cannon = 6 * by, king = by;
            unsigned m;
            int i1 = -1, i2 = -1;
            if ((l & 1) == 0) {
                m = line & ((1u << pos) - 1u);
                if (m) {
                    i1 = 31 - __clz(m);
                    m ^= 1u << i1;
                    if (m) i2 = 31 - __clz(m);
                }
            } else {
                m = line >> (pos + 1);
                if (m) {
                    i1 = pos + __ffs(m);
                    m &= m - 1u;
                    if (m) i2 = pos + __ffs(m);
                }
            }
            if (i1 >= 0) {
                const int p1 = cell_after(b, sq0 + i1 * stride, from, to, mover);
                hit = (p1 == rook || p1 == king);
                if (!hit && i2 >= 0) hit = cell_after(b, sq0 + i2 * stride, from, to, mover) == cannon;
            }
        } else if (l < 12) {
            const int i = l - 4;
            const int jr = (i < 4) ? ((i < 2) ? -2 : 2) : ((i < 6) ? -1 : 1);
            const int jc = (i < 4) ? ((i & 1) ? 1 : -1) : ((i & 1) ? 2 : -2);
            const int nr = kr + jr, nc = kc + jc;
            if ((unsigned)nr < 10u && (unsigned)nc < 9u && cell_after(b, nr * 9 + nc, from, to, mover) == 4 * by) {
                int lr = nr, lc = nc;
                if (i < 4) lr = nr - jr / 2; else lc = nc - jc / 2;
                hit = cell_after(b, lr * 9 + lc, from, to, mover) == 0;
            }
        } else if (l < 15) {
            const int j = l - 12;
            int pr = kr, pc = kc;
            bool ok;
            if (j == 0) {
                pr = kr - by;                               // the pawn stands one step behind its forward move
                ok = (unsigned)pr < 10u;
            } else {
                pc = kc + (j == 1 ? -1 : 1);
                ok = (by == 1 ? kr >= 5 : kr <= 4) && (unsigned)pc < 9u;
            }
            if (ok) hit = cell_after(b, pr * 9 + pc, from, to, mover) == 7 * by;
        }
    }
    return warp_ballot(hit);
}

struct MovegenResult {
    int n_legal;      // may exceed kMaxMoves only on unorthodox boards (then truncated + flagged)
    bool in_check;    // cy_is_in_check: own king attacked, or missing
    bool overflow;
    KingInfo kings;
};

// Everything the generator needs to know about the whole board, from ONE warp-synchronous pass: every lane reads
// its three squares (row-major index k*32+lane) and its three squares in column-major order, 18 votes follow.
struct BoardScan {
    uint32_t own[3];      // own pieces, bit = square
    uint32_t occ[3];      // occupied squares, bit = square
    uint32_t occT[3];     // occupied squares, bit = col*10 + row
    uint32_t kown[3];     // squares holding a king of `side`
    uint32_t kfoe[3];     // squares holding a king of `-side`
};
static __device__ __noinline__ BoardScan warp_board_scan(const int8_t* b, int side)
{
    __syncwarp();
    const int lane = threadIdx.x & 31;
    BoardScan s;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int i = k * 32 + lane;
        const int p = i < kSquares ? b[i] : 0;
        const int c = i / 10, r = i - c * 10;
        const int pT = i < kSquares ? b[r * 9 + c] : 0;
        s.own[k] = __ballot_sync(kFullMask, p * side > 0);
        s.occ[k] = __ballot_sync(kFullMask, p != 0);
        s.occT[k] = __ballot_sync(kFullMask, pT != 0);
        s.kown[k] = __ballot_sync(kFullMask, p == side);
        s.kfoe[k] = __ballot_sync(kFullMask, p == -side);
    }
    return s;
}

// `n` bits starting at bit `start` of the 96-bit value w2:w1:w0 (start + n <= 96)
__device__ __forceinline__ unsigned bits96(uint32_t w0, uint32_t w1, uint32_t w2, int start, int n)
{
    const int w = start >> 5;
    const uint32_t lo = w == 0 ? w0 : (w == 1 ? w1 : w2);
    const uint32_t hi = w == 0 ? w1 : (w == 1 ? w2 : 0u);
    return __funnelshift_r(lo, hi, start & 31) & ((1u << n) - 1u);
}

constexpr uint32_t kPalaceRedW0 = (7u << 3) | (7u << 12) | (7u << 21);     // squares 3-5, 12-14, 21-23
constexpr uint32_t kPalaceBlackW2 = (7u << 2) | (7u << 11) | (7u << 20);   // squares 66-68, 75-77, 84-86 (bit = sq-64)

// Full ordered legal-move generation for one board by one warp.
// b: 90 cells in shared memory; S.actions receives the compacted list (unused slots = -1).
__device__ __forceinline__ MovegenResult warp_movegen(const int8_t* b, int side, WarpScratch& S)
{
    const int lane = lane_id();
    MovegenResult res;
    if (lane == 0) *reinterpret_cast<uint32_t*>(S.kingto) = 0xffffffffu;
    const BoardScan bs = warp_board_scan(b, side);
    // kings: _find_king scans the own palace only, row-major (pyx:78-101) = lowest set bit
    bool odd;                     // a king-valued own piece outside its palace (unreachable boards)
    {
        const uint32_t ownp = side == 1 ? (bs.kown[0] & kPalaceRedW0) : (bs.kown[2] & kPalaceBlackW2);
        const uint32_t foep = side == 1 ? (bs.kfoe[2] & kPalaceBlackW2) : (bs.kfoe[0] & kPalaceRedW0);
        const int own_base = side == 1 ? 0 : 64, foe_base = side == 1 ? 64 : 0;
        res.kings.own_sq = ownp ? own_base + __ffs(ownp) - 1 : -1;
        res.kings.foe_sq = foep ? foe_base + __ffs(foep) - 1 : -1;
        res.kings.own_cnt = __popc(ownp);
        res.kings.foe_cnt = __popc(foep);
        odd = (side == 1 ? ((bs.kown[0] & ~kPalaceRedW0) | bs.kown[1] | bs.kown[2])
                         : (bs.kown[0] | bs.kown[1] | (bs.kown[2] & ~kPalaceBlackW2))) != 0u;
    }
    // occupancy masks: lanes 0-9 one row each, lanes 10-18 one column each
    if (lane < 10) S.rowocc[lane] = (uint16_t)bits96(bs.occ[0], bs.occ[1], bs.occ[2], lane * 9, 9);
    else if (lane < 19) S.colocc[lane - 10] = (uint16_t)bits96(bs.occT[0], bs.occT[1], bs.occT[2], (lane - 10) * 10, 10);
    // own pieces in square order: rank among own pieces = task group
    const int n0 = __popc(bs.own[0]), n1 = __popc(bs.own[1]);
    const int n_pieces = n0 + n1 + __popc(bs.own[2]);
    {
        const uint32_t lt = (1u << lane) - 1u;
        if (bs.own[0] >> lane & 1u) S.own_sq[__popc(bs.own[0] & lt)] = (uint8_t)lane;
        if (bs.own[1] >> lane & 1u) S.own_sq[n0 + __popc(bs.own[1] & lt)] = (uint8_t)(32 + lane);
        if (bs.own[2] >> lane & 1u) S.own_sq[n0 + n1 + __popc(bs.own[2] & lt)] = (uint8_t)(64 + lane);
    }
    // unused output slots read back as -1
    reinterpret_cast<uint2*>(S.actions)[lane] = make_uint2(0xffffffffu, 0xffffffffu);   // 128 x int16
    warp_sync();

    // ---- phase A: ordered pseudo-legal list --------------------------------------------
    const int K = res.kings.own_sq;
    int n_pseudo = 0;
    for (int t0 = 0; t0 < 4 * n_pieces; t0 += 32) {
        const int t = t0 + lane;
        int cnt = 0, first_t = 0, step = 0, x = -1, n_run = 0, sq = 0;
        if (t < 4 * n_pieces) {
            sq = S.own_sq[t >> 2];
            const int p = b[sq];
            const int kind = p < 0 ? -p : p;
            n_run = gen_task_run(b, S.rowocc, S.colocc, side, sq, kind, t & 3, first_t, step, x);
            cnt = n_run + (x >= 0 ? 1 : 0);
            if (sq == K && n_run) S.kingto[t & 3] = (uint8_t)first_t;    // the palace king's target in direction t&3
        }
        int total;
        const int incl = warp_incl_scan(cnt, &total);
        const int off = n_pseudo + incl - cnt;
        if (cnt > 0 && off + cnt <= kMaxPseudo) {
            const int hi = sq << 7;
            for (int i = 0; i < n_run; ++i) S.pm[off + i] = (uint16_t)(hi | (first_t + i * step));
            if (x >= 0) S.pm[off + n_run] = (uint16_t)(hi | x);
        }
        n_pseudo += total;
    }
    res.overflow = n_pseudo > kMaxPseudo;
    if (n_pseudo > kMaxPseudo) n_pseudo = kMaxPseudo;
    warp_sync();                  // pm / kingto / actions writes above are visible to every lane below

    // ---- attack queries on half warps: the in-check flag (cy_is_in_check pyx:543-555: missing king => True) and
    // the palace king's own moves (full _is_attacked at the new square, old square vacated) ----------------
    const int kr = K < 0 ? 0 : K / 9, kc = K < 0 ? 0 : K - kr * 9;
    unsigned king_ok = 0u;        // bit d: the king's move in direction d is legal
    res.in_check = true;
    if (K >= 0) {
        const uint32_t kt = *reinterpret_cast<const uint32_t*>(S.kingto);
        // bit d set where byte d of kt is a square (< 0x80) and not 0xff
        unsigned vm = (~kt & 0x80808080u) * 0x00204081u >> 28;
        bool first = true;
        do {
            // low half: the probe (first pass) or a king move; high half: the next king move
            int d0 = -1, d1 = -1;
            if (!first && vm) { d0 = __ffs(vm) - 1; vm &= vm - 1u; }
            if (vm) { d1 = __ffs(vm) - 1; vm &= vm - 1u; }
            const int d = lane < 16 ? d0 : d1;
            const bool probe = first && lane < 16;
            const int qto = d >= 0 ? (int)((kt >> (8 * d)) & 0xffu) : K;
            const int qr = qto / 9, qc = qto - qr * 9;
            const unsigned hits = warp_attacked_pair(b, S.rowocc, S.colocc, probe || d >= 0, qr, qc, -side,
                                                     probe ? -1 : K, qto, side);
            if (first) res.in_check = (hits & 0xffffu) != 0u;
            else if (d0 >= 0 && (hits & 0xffffu) == 0u) king_ok |= 1u << d0;
            if (d1 >= 0 && (hits >> 16) == 0u) king_ok |= 1u << d1;
            first = false;
        } while (vm);
    }

    // ---- phase B: legality + ordered compaction ---------------------------------------------------
    int n_legal = 0;
    const bool fast = !res.in_check && !odd && res.kings.own_cnt == 1 && res.kings.foe_cnt <= 1;
    if (fast) {
        // The king is not attacked now, so a move of another piece can only expose it by changing the king's own
        // row or column (a blocker leaves, a cannon screen arrives) or by vacating a knight leg next to the king;
        // only those lines / those two knight origins are re-examined on the overlaid board.  Enemy pawns and
        // every other knight see the same squares as before.  Facing kings (pyx:231-245) are covered by the
        // column test: the enemy king counts as a rook on open rays (pyx:117).  King moves were answered above.
        const int by = -side;
        for (int i0 = 0; i0 < n_pseudo; i0 += 32) {
            const int i = i0 + lane;
            bool ok = false;
            int from = 0, to = 0;
            if (i < n_pseudo) {
                const int e = S.pm[i];
                from = e >> 7;
                to = e & 127;
                const int mover = b[from];
                if (mover == side) {
                    const int dl = to - from;             // -9 up, 9 down, -1 left, 1 right = directions 0..3
                    const int d = dl == -9 ? 0 : (dl == 9 ? 1 : (dl == -1 ? 2 : 3));
                    ok = (king_ok >> d) & 1u;
                } else {
                    const int fr = from / 9, fc = from - fr * 9, tr = to / 9, tc = to - tr * 9;
                    bool bad = false;
                    if (fr == kr || tr == kr) {
                        unsigned R = S.rowocc[kr];
                        if (fr == kr) R &= ~(1u << fc);
                        if (tr == kr) R |= 1u << tc;
                        bad = line_attacked(b, R, kc, kr * 9, 1, by, from, to, mover);
                    }
                    if (!bad && (fc == kc || tc == kc)) {
                        unsigned Cm = S.colocc[kc];
                        if (fc == kc) Cm &= ~(1u << fr);
                        if (tc == kc) Cm |= 1u << tr;
                        bad = line_attacked(b, Cm, kr, kc, 9, by, from, to, mover);
                    }
                    const int ddr = fr - kr, ddc = fc - kc;
                    if (!bad && (ddr == 1 || ddr == -1) && (ddc == 1 || ddc == -1)) {
                        const int r1 = kr + 2 * ddr, c2 = kc + 2 * ddc;
                        if ((unsigned)r1 < 10u && cell_after(b, r1 * 9 + fc, from, to, mover) == 4 * by) bad = true;
                        if ((unsigned)c2 < 9u && cell_after(b, fr * 9 + c2, from, to, mover) == 4 * by) bad = true;
                    }
                    ok = !bad;
                }
            }
            const unsigned m = warp_ballot(ok);
            if (ok) {
                const int pos = n_legal + __popc(m & ((1u << lane) - 1u));
                if (pos < kMaxMoves) S.actions[pos] = (int16_t)(from * 90 + to);
            }
            n_legal += __popc(m);
        }
    } else {
        // in check, or a board no game can reach: every move through the general test (pyx:209-252)
        for (int i0 = 0; i0 < n_pseudo; i0 += 32) {
            const int i = i0 + lane;
            bool ok = false;
            int from = 0, to = 0;
            if (i < n_pseudo) {
                const int e = S.pm[i];
                from = e >> 7;
                to = e & 127;
                ok = legal_after_move(b, S.rowocc, S.colocc, side, res.kings, from, to);
            }
            const unsigned m = warp_ballot(ok);
            if (ok) {
                const int pos = n_legal + __popc(m & ((1u << lane) - 1u));
                if (pos < kMaxMoves) S.actions[pos] = (int16_t)(from * 90 + to);
            }
            n_legal += __popc(m);
        }
    }
    if (n_legal > kMaxMoves) res.overflow = true;
    res.n_legal = n_legal;
    warp_sync();
    return res;
}

// ------------------------------------------------------------------------------------------
// Game state and termination (game.py:528-616).  len(history) == move_count and only
// history[-12:] is ever compared, so a 12-slot ring of pre-move boards is the whole history.
// ------------------------------------------------------------------------------------------
constexpr int kRing = 12;
constexpr int kBoardPad = 96;     // 90 cells padded to 96 B so every board is 16 B aligned

struct GameMeta {
    int side;          // current_player: +1 red, -1 black
    int move_count;
    int no_capture;
};

__device__ __forceinline__ int piece_value(int kind)
{
    // PIECE_VALUES game.py:74: K0 A20 B20 N40 R90 C45 P10
    return kind == 2 || kind == 3 ? 20 : kind == 4 ? 40 : kind == 5 ? 90 : kind == 6 ? 45 : kind == 7 ? 10 : 0;
}

// red material minus black material, warp-reduced (game.py:552-563, 595-604)
__device__ __forceinline__ int warp_material_diff(const int8_t* b)
{
    const int lane = lane_id();
    int s = 0;
    for (int sq = lane; sq < kSquares; sq += 32) {
        int p = b[sq];
        s += p > 0 ? piece_value(p) : -piece_value(-p);
    }
    return warp_sum(s);
}

// make_move (game.py:528-545) on a warp-owned state: ring[move_count % 12] <- board before the move.
__device__ __forceinline__ void warp_make_move(int8_t* b, int8_t* ring, GameMeta& g, int action)
{
    const int lane = lane_id();
    const int from = action / 90, to = action % 90;
    int8_t* slot = ring + (g.move_count % kRing) * kBoardPad;
    for (int i = lane; i < kSquares; i += 32) slot[i] = b[i];
    warp_sync();
    int taken = b[to];
    warp_sync();
    if (lane == 0) {
        b[to] = b[from];
        b[from] = 0;
    }
    g.no_capture = taken != 0 ? 0 : g.no_capture + 1;
    g.side = -g.side;
    g.move_count += 1;
    warp_sync();
}

// is_game_over (game.py:565-616).  Returns winner in {1,-1,0} or 2 when the game goes on.
// mg must be the movegen result for (b, g.side).
__device__ __forceinline__ int warp_game_over(const int8_t* b, const int8_t* ring, const GameMeta& g,
                                              const MovegenResult& mg)
{
    const int lane = lane_id();
    // king presence: _find_king scans the own palace only
    int red_king = g.side == 1 ? mg.kings.own_sq : mg.kings.foe_sq;
    int black_king = g.side == 1 ? mg.kings.foe_sq : mg.kings.own_sq;
    if (red_king < 0) return -1;
    if (black_king < 0) return 1;
    if (mg.n_legal == 0) return -g.side;
    if (g.no_capture >= 120) return 0;
    if (g.move_count >= 200) {
        int diff = warp_material_diff(b);
        return diff > 30 ? 1 : (diff < -30 ? -1 : 0);
    }
    if (g.move_count >= 6) {
        const int depth = min(g.move_count, kRing);
        // lane k compares history slot k (one of the last `depth` pre-move boards)
        bool same = false;
        if (lane < depth) {
            const int8_t* h = ring + ((g.move_count - 1 - lane) % kRing) * kBoardPad;
            same = true;
            // 90 bytes; boards are 16 B aligned in shared memory
            const uint32_t* a = reinterpret_cast<const uint32_t*>(b);
            const uint32_t* c = reinterpret_cast<const uint32_t*>(h);
#pragma unroll
            for (int w = 0; w < 22; ++w) same = same && (a[w] == c[w]);
            same = same && (b[88] == h[88]) && (b[89] == h[89]);
        }
        if (__popc(warp_ballot(same)) >= 3) return 0;
    }
    return 2;
}

// counter-based RNG (stateless, reproducible per (seed, game, ply, stream))
__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t rng_u64(uint64_t seed, uint64_t a, uint64_t b, uint64_t c)
{
    return mix64(mix64(mix64(seed + 0x9E3779B97F4A7C15ull * (a + 1)) ^ (b * 0xD1B54A32D192ED03ull)) ^
                 (c * 0x8CB92BA72F3D8DD7ull));
}

}  // namespace xq
