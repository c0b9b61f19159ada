// xq_rules_tpb.h -- the xiangqi rules as ONE THREAD PER BOARD scalar code (K1, second generation).
//
// Same contract as the warp-per-board generator of xq_rules.cuh: the ordered legal move list of
// game_core.pyx:_generate_moves (:262-486) with _is_move_legal (:209-252) and _is_attacked (:104-189),
// plus the in-check flag of cy_is_in_check (:543-555), bit-exact on every int8[90] input with piece
// codes in -7..7 (orthodox or not: extra kings, kings outside their palace, any number of knights).
//
// Why a second shape: the warp-per-board kernel is bound by instruction issue (1 867 warp-instructions
// per position, 17.5 of 32 lanes active, profiles/r1_movegen_ncu.md) -- a 90-cell board does not have
// 32-wide parallelism in its legality tests.  Here a lane owns a whole board, so a warp advances 32
// positions per instruction stream and the only cost of divergence is the union of the code paths the
// 32 boards take, which the structure below keeps small (916 warp-instructions per position, 22 lanes):
//   1. scan_board: one pass over the 90 cells (a row per iteration) builds, in registers, the occupancy
//      sets (row-major and column-major 96-bit), the own-piece set, the occupancy of the three rows and
//      columns of the own palace, the own kings standing in the palace and the enemy knights;
//   2. gen_piece: pseudo-legal generation walks the own-piece set in square order.  The five leapers
//      (king, advisor, elephant, knight, pawn) share one table-driven loop (4 signed bytes per slot:
//      dr, dc, leg offset, square delta; a per-kind box bounds the target), rook and cannon share one
//      path in which a ray is a bit scan of the row / column occupancy (first blocker = rook capture,
//      second = cannon capture) and only the emission of the empty run is a loop;
//   3. a GENERAL pass, all lanes together: the in-check probe of the position and the king's own moves
//      (at most 4), with the full attack test `attacked` -- king square, lines, knights, pawns all depend
//      on the move; the verdict replaces the list entry;
//   4. the main loop tests every other pseudo-legal move with `attacked_fixed`: the king stays where it
//      is, so the row / column masks, the leg square of each enemy knight and the squares of attacking
//      pawns and the rook / king / cannon sets of the two lines are prepared once per board (KingCtx) and a
//      move costs four rays of bit arithmetic on the overlaid masks, two leg reads and five compares; the
//      board is not touched (only the general pass makes and unmakes its moves in place).  Facing kings need
//      no extra clause: _is_attacked counts the enemy king as a rook on an
//      open ray (pyx:117), which is the flying-general test of pyx:226-240.
// Both long loops run for the warp-wide maximum trip count with predicated bodies (XQT_WARP_MAX): with
// per-lane trip counts the lanes of a warp drifted apart for good (9.5 of 32 lanes active).  Legal moves
// are compacted in place into action ids, which the kernel copies out with coalesced stores.
//
// The file is plain C++ (the two warp collectives are macros that are the identity on the host):
// tests/test_tpb_cpu.py compiles it with g++ and checks it against the oracle over random-playout and
// piece-soup positions, so the rules logic the kernel runs is verified on the CPU build box too; the
// kernel wrapper is xq_movegen.cu:movegen_tpb_kernel.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define XQT_HD __host__ __device__ __forceinline__
#else
#define XQT_HD inline
#endif

// Warp-uniform loops.  A loop whose trip count differs between the lanes of a warp lets the lanes drift apart for
// good (the first version of this file ran with 9.5 of 32 lanes active: ncu, profiles/r1_movegen_ncu.md), so the two
// long loops of movegen() run for the warp-wide MAXIMUM trip count with a predicated body and a reconvergence point
// at the end of every iteration.  On the device all 32 lanes of a warp must therefore call movegen() together (lanes
// without a position pass an empty board); on the host the three macros are the identity.
#if defined(__CUDA_ARCH__)
namespace xqt {
static __device__ __noinline__ int warp_max_i(int v)
{
    __syncwarp();
    return __reduce_max_sync(0xffffffffu, v);
}
}  // namespace xqt
#define XQT_WARP_MAX(v) ::xqt::warp_max_i(v)
#define XQT_RECONVERGE() __syncwarp()
#else
#define XQT_WARP_MAX(v) (v)
#define XQT_RECONVERGE() ((void)0)
#endif

namespace xqt {

constexpr int kListCap = 98;    // pseudo-legal scratch entries per board and round (a piece adds at most 17; orthodox maximum seen: 75)
constexpr int kMaxOut = 128;

XQT_HD int div9(int x) { return (x * 57) >> 9; }   // x / 9 for 0 <= x < 128

XQT_HD int popc32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}

XQT_HD int ctz32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return v ? __builtin_ctz(v) : -1;
#endif
}
XQT_HD int top32(uint32_t v)   // index of the highest set bit; -1 for 0 (like ctz32), callers select a safe index then
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return v ? 31 - __builtin_clz(v) : -1;
#endif
}

// n <= 16 bits of a 96-bit set starting at bit `start`
XQT_HD uint32_t bits96(uint32_t w0, uint32_t w1, uint32_t w2, int start, int n)
{
    const int w = start >> 5, sh = start & 31;
    const uint32_t lo = w == 0 ? w0 : (w == 1 ? w1 : w2);
    const uint32_t hi = w == 0 ? w1 : (w == 1 ? w2 : 0u);
#if defined(__CUDA_ARCH__)
    const uint32_t v = __funnelshift_r(lo, hi, sh);
#else
    const uint32_t v = (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
    return v & ((1u << n) - 1u);
}

struct Scan {
    uint32_t occR[3];   // occupied squares, bit r*9+c
    uint32_t occC[3];   // occupied squares, bit c*10+r
    uint32_t own[3];    // own pieces, bit r*9+c
    uint32_t kmask;     // own kings standing in the own palace, bit (r-r0)*3+(c-3)  (find_king order, pyx:93-98)
    int n_ek;           // enemy knights on the board; the first two squares in ek0/ek1 (-1: none) with their
    int ek0, ek1;       // rows and columns (-100 when absent: no king square is a knight's move away)
    int ek0r, ek0c, ek1r, ek1c;
    int n_own_kings;    // own king pieces anywhere on the board (1 in play)
    int n_own_knights;  // own knights (each takes two generation steps of four slots)
    uint32_t prow[3];   // occupancy of the three rows (bit c) and
    uint32_t pcol[3];   // of the three columns (bit r) of the own palace: the only lines a king's attack test looks along
};

XQT_HD Scan scan_board(const int8_t* b, int side)
{
    Scan s;
    s.occR[0] = s.occR[1] = s.occR[2] = 0u;
    s.own[0] = s.own[1] = s.own[2] = 0u;
    s.prow[0] = s.prow[1] = s.prow[2] = 0u;
    s.kmask = 0u;
    s.n_ek = 0;
    s.n_own_kings = 0;
    s.n_own_knights = 0;
    s.ek0 = s.ek1 = -1;
    s.ek0r = s.ek0c = s.ek1r = s.ek1c = -100;
    const int r0 = side == 1 ? 0 : 7;
    uint32_t colm[9] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};   // column occupancy, bit r (compile-time column index)
    // one row per iteration, NOT unrolled across rows: the fully unrolled 90-cell scan was 1 400 instructions and the
    // kernel stalled on instruction fetch (ncu: 69 % of this section's samples)
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = 0; r < 10; ++r) {
        const int8_t* rowp = b + r * 9;
        uint32_t rm = 0u, om = 0u, kings = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int c = 0; c < 9; ++c) {
            const int p = rowp[c];
            const uint32_t nz = p != 0 ? 1u : 0u;
            rm |= nz << c;
            om |= ((p * side > 0) ? 1u : 0u) << c;
            colm[c] |= nz << r;
            s.n_own_kings += p == side ? 1 : 0;
            s.n_own_knights += p == 4 * side ? 1 : 0;
            if (c >= 3 && c <= 5) kings |= (p == side ? 1u : 0u) << (c - 3);
            if (p == -4 * side) {
                if (s.n_ek == 0) { s.ek0 = r * 9 + c; s.ek0r = r; s.ek0c = c; }
                else if (s.n_ek == 1) { s.ek1 = r * 9 + c; s.ek1r = r; s.ek1c = c; }
                ++s.n_ek;
            }
        }
        // insert the 9 row bits at bit r*9 of the 96-bit sets
        const int at = r * 9, w = at >> 5, sh = at & 31;
        const uint32_t rlo = rm << sh, rhi = sh > 23 ? rm >> (32 - sh) : 0u;
        const uint32_t olo = om << sh, ohi = sh > 23 ? om >> (32 - sh) : 0u;
        if (w == 0) { s.occR[0] |= rlo; s.occR[1] |= rhi; s.own[0] |= olo; s.own[1] |= ohi; }
        else if (w == 1) { s.occR[1] |= rlo; s.occR[2] |= rhi; s.own[1] |= olo; s.own[2] |= ohi; }
        else { s.occR[2] |= rlo; s.own[2] |= olo; }
        const int pr = r - r0;
        if (pr >= 0 && pr <= 2) {
            s.kmask |= kings << (pr * 3);
            if (pr == 0) s.prow[0] = rm; else if (pr == 1) s.prow[1] = rm; else s.prow[2] = rm;
        }
    }
    s.occC[0] = colm[0] | colm[1] << 10 | colm[2] << 20 | colm[3] << 30;
    s.occC[1] = colm[3] >> 2 | colm[4] << 8 | colm[5] << 18 | colm[6] << 28;
    s.occC[2] = colm[6] >> 4 | colm[7] << 6 | colm[8] << 16;
    s.pcol[0] = colm[3];
    s.pcol[1] = colm[4];
    s.pcol[2] = colm[5];
    return s;
}

// pyx:104-189 for a KING square, given as palace coordinates (pr, pc) of the attacked side's palace (row r0 + pr,
// column 3 + pc; a king is only ever "found" there, pyx:78-101) -- the cell holds a piece of the attacked side, which is what lets absent
// blockers read that cell as a harmless stand-in -- on the board as it stands in b[] (a move, if any, already made in
// place); the row / column occupancy of the un-moved board is corrected for the move (fr, fc) -> (tr, tc) (fr < 0: no
// move).  `cap` is the square a captured piece stood on (an enemy knight there no longer attacks), -1 for none.
// Straight-line code: a warp executes the union of its lanes' paths anyway, so early exits would only split it; the
// 13 cell reads are independent and overlap.
XQT_HD bool attacked(const int8_t* b, const Scan& s, int r0, int pr, int pc, int by, int fr, int fc, int tr, int tc, int cap)
{
    const int rook = 5 * by, cannon = 6 * by, pawn = 7 * by, king = by;
    const int kr = r0 + pr, kc = 3 + pc;
    uint32_t R = pr == 0 ? s.prow[0] : (pr == 1 ? s.prow[1] : s.prow[2]);
    uint32_t C = pc == 0 ? s.pcol[0] : (pc == 1 ? s.pcol[1] : s.pcol[2]);
    if (fr >= 0) {
        R &= ~((fr == kr ? 1u : 0u) << fc);
        C &= ~((fc == kc ? 1u : 0u) << fr);
        R |= (tr == kr ? 1u : 0u) << tc;
        C |= (tc == kc ? 1u : 0u) << tr;
    }
    const int ksq = kr * 9 + kc;
    const int8_t* row = b + kr * 9;
    const int8_t* col = b + kc;
    bool hit = false;
    {   // towards column 0: nearest piece = highest bit below kc, the one behind it = next highest
        const uint32_t m = R & ((1u << kc) - 1u);
        const int i1 = m ? top32(m) : kc;
        const int p1 = row[i1];
        const uint32_t m2 = m & ~(1u << i1);
        const int i2 = m2 ? top32(m2) : kc;
        hit |= (p1 == rook) | (p1 == king) | (row[i2] == cannon);
    }
    {   // towards column 8
        const uint32_t m = R >> (kc + 1);
        const int i1 = m ? kc + 1 + ctz32(m) : kc;
        const int p1 = row[i1];
        const uint32_t m2 = m & (m - 1u);
        const int i2 = m2 ? kc + 1 + ctz32(m2) : kc;
        hit |= (p1 == rook) | (p1 == king) | (row[i2] == cannon);
    }
    {   // towards row 0
        const uint32_t m = C & ((1u << kr) - 1u);
        const int i1 = m ? top32(m) : kr;
        const int p1 = col[i1 * 9];
        const uint32_t m2 = m & ~(1u << i1);
        const int i2 = m2 ? top32(m2) : kr;
        hit |= (p1 == rook) | (p1 == king) | (col[i2 * 9] == cannon);
    }
    {   // towards row 9
        const uint32_t m = C >> (kr + 1);
        const int i1 = m ? kr + 1 + ctz32(m) : kr;
        const int p1 = col[i1 * 9];
        const uint32_t m2 = m & (m - 1u);
        const int i2 = m2 ? kr + 1 + ctz32(m2) : kr;
        hit |= (p1 == rook) | (p1 == king) | (col[i2 * 9] == cannon);
    }
    // knights (pyx:156-169): the leg is the cell next to the knight on its long axis
    if (s.n_ek <= 2) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 2; ++j) {
            const int nsq = j == 0 ? s.ek0 : s.ek1;
            const int dr = kr - (j == 0 ? s.ek0r : s.ek1r), dc = kc - (j == 0 ? s.ek0c : s.ek1c);
            const int adr = dr < 0 ? -dr : dr, adc = dc < 0 ? -dc : dc;
            const bool geo = (adr * adc == 2) & (adr + adc == 3) & (nsq != cap);       // (1,2) or (2,1), still on the board
            const int leg = nsq + (adr == 2 ? (dr > 0 ? 9 : -9) : (dc > 0 ? 1 : -1));
            hit |= geo & (b[geo ? leg : ksq] == 0);
        }
    } else {
        const int horse = 4 * by;
        for (int i = 0; i < 8; ++i) {                          // boards no game reaches: look at the 8 origins
            const int jr = (i < 4) ? ((i < 2) ? -2 : 2) : ((i < 6) ? -1 : 1);
            const int jc = (i < 4) ? ((i & 1) ? 1 : -1) : ((i & 1) ? 2 : -2);
            const int nr = kr + jr, nc = kc + jc;
            if (nr < 0 || nr >= 10 || nc < 0 || nc >= 9) continue;
            if (b[nr * 9 + nc] != horse) continue;
            int lr = nr, lc = nc;
            if (jr == 2 || jr == -2) lr = nr - jr / 2; else lc = nc - jc / 2;
            if (b[lr * 9 + lc] == 0) hit = true;
        }
    }
    // pawns (pyx:172-187): from behind, and from the side once the target row is across the river for that colour
    {
        const bool back_ok = by == 1 ? kr >= 1 : kr <= 8;
        const bool side_ok = by == 1 ? kr >= 5 : kr <= 4;
        const bool l_ok = side_ok & (kc >= 1), r_ok = side_ok & (kc <= 7);
        hit |= (b[back_ok ? ksq - 9 * by : ksq] == pawn) | (b[l_ok ? ksq - 1 : ksq] == pawn) | (b[r_ok ? ksq + 1 : ksq] == pawn);
    }
    return hit;
}

// The attack test for the common case: the own king stays on its square K0 (every move of another piece).  Everything
// that depends only on K0 is prepared once per board: the occupancy of its row and column, for each of the (at most two)
// enemy knights the leg square through which it would attack K0 (-1: it is not a knight's move away), and the up to
// three squares from which an enemy pawn attacks K0 right now (-1: none there; pawn attacks cannot be blocked, only
// captured), and the sets of enemy rooks-or-kings and cannons standing in that row and column.  Per move that leaves
// four rays of pure bit arithmetic on the overlaid masks, two leg reads and five compares -- a third of attacked().
struct KingCtx {
    int kr, kc, ksq;
    uint32_t R, C;           // occupancy of K0's row (bit c) and column (bit r)
    uint32_t rk_r, cn_r;     // enemy rooks-or-kings / cannons standing in that row ...
    uint32_t rk_c, cn_c;     // ... and in that column
    int leg0, leg1;          // leg square of enemy knight 0 / 1 against K0, or -1
    int pw0, pw1, pw2;       // squares holding an enemy pawn that attacks K0 (behind, left, right), or -1
};

XQT_HD KingCtx king_context(const int8_t* b, const Scan& s, int r0, int pr, int pc, int by)
{
    KingCtx k;
    k.kr = r0 + pr;
    k.kc = 3 + pc;
    k.ksq = k.kr * 9 + k.kc;
    k.R = pr == 0 ? s.prow[0] : (pr == 1 ? s.prow[1] : s.prow[2]);
    k.C = pc == 0 ? s.pcol[0] : (pc == 1 ? s.pcol[1] : s.pcol[2]);
    const int rook = 5 * by, cannon = 6 * by, king = by;
    k.rk_r = k.cn_r = k.rk_c = k.cn_c = 0u;
    for (int c = 0; c < 9; ++c) {
        const int p = b[k.kr * 9 + c];
        k.rk_r |= ((p == rook) | (p == king) ? 1u : 0u) << c;
        k.cn_r |= (p == cannon ? 1u : 0u) << c;
    }
    for (int r = 0; r < 10; ++r) {
        const int p = b[r * 9 + k.kc];
        k.rk_c |= ((p == rook) | (p == king) ? 1u : 0u) << r;
        k.cn_c |= (p == cannon ? 1u : 0u) << r;
    }
    {
        const int dr = k.kr - s.ek0r, dc = k.kc - s.ek0c;
        const int adr = dr < 0 ? -dr : dr, adc = dc < 0 ? -dc : dc;
        const bool geo = (adr * adc == 2) & (adr + adc == 3);
        k.leg0 = geo ? s.ek0 + (adr == 2 ? (dr > 0 ? 9 : -9) : (dc > 0 ? 1 : -1)) : -1;
    }
    {
        const int dr = k.kr - s.ek1r, dc = k.kc - s.ek1c;
        const int adr = dr < 0 ? -dr : dr, adc = dc < 0 ? -dc : dc;
        const bool geo = (adr * adc == 2) & (adr + adc == 3);
        k.leg1 = geo ? s.ek1 + (adr == 2 ? (dr > 0 ? 9 : -9) : (dc > 0 ? 1 : -1)) : -1;
    }
    const int pawn = 7 * by;
    const bool back_ok = by == 1 ? k.kr >= 1 : k.kr <= 8;
    const bool side_ok = by == 1 ? k.kr >= 5 : k.kr <= 4;
    const int bsq = back_ok ? k.ksq - 9 * by : k.ksq, lsq = (side_ok & (k.kc >= 1)) ? k.ksq - 1 : k.ksq,
              rsq = (side_ok & (k.kc <= 7)) ? k.ksq + 1 : k.ksq;
    k.pw0 = b[bsq] == pawn ? bsq : -1;          // the king's own cell never holds an enemy pawn
    k.pw1 = b[lsq] == pawn ? lsq : -1;
    k.pw2 = b[rsq] == pawn ? rsq : -1;
    return k;
}

XQT_HD uint32_t hibit(uint32_t m) { return m ? 1u << top32(m) : 0u; }   // highest set bit as a mask

// b[] is the UN-moved board: the move from -> to is overlaid, nothing is written (only the two knight legs are read).
// A ray is pure bit arithmetic: the nearest occupied cell of the overlaid line attacks if it is in the rook-or-king
// set, the one behind it if it is in the cannon set; the piece captured on `to` leaves both sets, the moved piece is
// not in them.
XQT_HD bool attacked_fixed(const int8_t* b, const Scan& s, const KingCtx& k, int from, int fr, int fc, int to, int tr, int tc)
{
    const int kr = k.kr, kc = k.kc;
    const uint32_t tb_r = (tr == kr ? 1u : 0u) << tc, tb_c = (tc == kc ? 1u : 0u) << tr;
    const uint32_t R = (k.R & ~((fr == kr ? 1u : 0u) << fc)) | tb_r;
    const uint32_t C = (k.C & ~((fc == kc ? 1u : 0u) << fr)) | tb_c;
    const uint32_t rk_r = k.rk_r & ~tb_r, cn_r = k.cn_r & ~tb_r, rk_c = k.rk_c & ~tb_c, cn_c = k.cn_c & ~tb_c;
    uint32_t att = 0u;
    {
        const uint32_t lo = R & ((1u << kc) - 1u);            // towards column 0: nearest = highest bit
        const uint32_t f = hibit(lo);
        att |= (f & rk_r) | (hibit(lo ^ f) & cn_r);
        const uint32_t hi = R & ~((2u << kc) - 1u);           // towards column 8: nearest = lowest bit
        const uint32_t g = hi & (0u - hi);
        const uint32_t hi2 = hi ^ g;
        att |= (g & rk_r) | (hi2 & (0u - hi2) & cn_r);
    }
    {
        const uint32_t lo = C & ((1u << kr) - 1u);
        const uint32_t f = hibit(lo);
        att |= (f & rk_c) | (hibit(lo ^ f) & cn_c);
        const uint32_t hi = C & ~((2u << kr) - 1u);
        const uint32_t g = hi & (0u - hi);
        const uint32_t hi2 = hi ^ g;
        att |= (g & rk_c) | (hi2 & (0u - hi2) & cn_c);
    }
    bool hit = att != 0u;
    // knights: still there (not the captured piece) and the leg empty after the move (vacated by it, or empty before and
    // not landed on)
    const bool e0 = (k.leg0 == from) | ((k.leg0 != to) & (b[k.leg0 >= 0 ? k.leg0 : k.ksq] == 0));
    const bool e1 = (k.leg1 == from) | ((k.leg1 != to) & (b[k.leg1 >= 0 ? k.leg1 : k.ksq] == 0));
    hit |= ((k.leg0 >= 0) & (s.ek0 != to) & e0) | ((k.leg1 >= 0) & (s.ek1 != to) & e1);
    // pawns: still there
    hit |= ((k.pw0 >= 0) & (k.pw0 != to)) | ((k.pw1 >= 0) & (k.pw1 != to)) | ((k.pw2 >= 0) & (k.pw2 != to));
    return hit;
}

// Leaper table (pyx:287-367, 434-484 orders): 8 rows x 8 slots, row = piece kind (1 king, 2 advisor, 3 elephant,
// 4 knight, 7 red pawn; row 0 = black pawn), entry = four signed bytes {dr, dc, leg offset, square delta dr*9+dc};
// delta == 0 marks an unused slot, leg offset == 0 a move without a leg / eye.  256 bytes of shared memory in the
// kernel (filled with slot_entry()); the host build uses a plain array.
constexpr int kSlotTableSize = 64;
XQT_HD uint32_t slot_entry(int idx)
{
    const int row = idx >> 3, sl = idx & 7;
    int dr = 0, dc = 0, leg = 0;
    if (row == 1 && sl < 4) {                       // king: up, down, left, right
        dr = sl == 0 ? -1 : (sl == 1 ? 1 : 0);
        dc = sl == 2 ? -1 : (sl == 3 ? 1 : 0);
    } else if ((row == 2 || row == 3) && sl < 4) {  // advisor / elephant: (-,-) (-,+) (+,-) (+,+)
        const int k = row == 2 ? 1 : 2;
        dr = sl < 2 ? -k : k;
        dc = (sl & 1) ? k : -k;
        if (row == 3) leg = (dr / 2) * 9 + dc / 2;
    } else if (row == 4) {                          // knight (pyx:31-39)
        dr = sl < 4 ? (sl < 2 ? -2 : 2) : (sl < 6 ? -1 : 1);
        dc = sl < 4 ? ((sl & 1) ? 1 : -1) : ((sl & 1) ? 2 : -2);
        leg = sl < 4 ? (dr / 2) * 9 : dc / 2;
    } else if ((row == 7 || row == 0) && sl < 3) {  // pawn: forward, left, right
        dr = sl == 0 ? (row == 7 ? 1 : -1) : 0;
        dc = sl == 1 ? -1 : (sl == 2 ? 1 : 0);
    }
    return (uint32_t)(dr & 0xff) | (uint32_t)(dc & 0xff) << 8 | (uint32_t)(leg & 0xff) << 16 | (uint32_t)((dr * 9 + dc) & 0xff) << 24;
}

// Pseudo-legal targets of the piece on `from` appended to list[m...] as from << 7 | to; returns the new m.
// Two sections (sliders / leapers), each straight-line with predicated stores: the lanes of a warp that are in the same
// section stay together.
XQT_HD int gen_piece(const int8_t* b, const Scan& s, int side, int from, int slot0, uint16_t* list, int m, const uint32_t* tab)
{
    const int p = b[from];
    const int kind = p < 0 ? -p : p;
    const int r = div9(from), c = from - r * 9;
    if (kind == 5 || kind == 6) {
        const uint32_t Rm = bits96(s.occR[0], s.occR[1], s.occR[2], r * 9, 9);
        const uint32_t Cm = bits96(s.occC[0], s.occC[1], s.occC[2], c * 10, 10);
        // pyx:42-46 order: up (row - 1), down (row + 1), left (col - 1), right (col + 1)
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int d = 0; d < 4; ++d) {
            const bool vertical = d < 2, neg = (d & 1) == 0;
            const uint32_t line = vertical ? Cm : Rm;
            const int pos = vertical ? r : c;
            const int len = vertical ? 10 : 9;
            const int step = vertical ? 9 : 1;
            int n_empty, first, second;
            if (neg) {
                const uint32_t q = line & ((1u << pos) - 1u);
                first = top32(q);                               // -1: open to the edge
                n_empty = pos - 1 - first;
                const uint32_t q2 = q & ~(1u << (first & 31));
                second = q ? top32(q2) : -1;
            } else {
                const uint32_t q = line >> (pos + 1);
                const int z = ctz32(q);
                first = q ? pos + 1 + z : -1;
                n_empty = q ? z : len - 1 - pos;
                const uint32_t q2 = q & (q - 1u);
                second = q2 ? pos + 1 + ctz32(q2) : -1;
            }
            const int sgn = neg ? -step : step;
            int to = from;
            for (int i = 0; i < n_empty; ++i) {
                to += sgn;
                list[m++] = (uint16_t)(from << 7 | to);
            }
            const int hit = kind == 5 ? first : second;       // rook takes the first piece on the ray, cannon the second
            const int tsq = from + (hit - pos) * step;
            if (hit >= 0 && b[tsq] * side < 0) list[m++] = (uint16_t)(from << 7 | tsq);
        }
    } else if (kind == 1 || kind == 2 || kind == 3 || kind == 4 || kind == 7) {
        // target box: palace for king and advisor (the advisor test of pyx:307-326 is the box too), own half for the
        // elephant, the board for knight and pawn
        const bool palace = kind <= 2;
        const int rlo = palace ? (side == 1 ? 0 : 7) : (kind == 3 && side != 1 ? 5 : 0);
        const int rspan = palace ? 2 : (kind == 3 ? 4 : 9);          // target rows rlo .. rlo + rspan
        const int clo = palace ? 3 : 0, cspan = palace ? 2 : 8;
        const bool crossed = side == 1 ? r >= 5 : r <= 4;
        // four slots per call: a knight's eight come in two calls (slot0 = 0, then 4), so that the slot loop of a warp
        // never runs its second half for the knights alone
        const int nsl = kind == 7 ? (crossed ? 3 : 1) : 4;
        const uint32_t* row = tab + ((kind == 7 && side != 1) ? 0 : kind) * 8 + slot0;
        const int from7 = from << 7 | (kind == 1 ? 0x4000 : 0);       // bit 14 marks a king move (tested apart, see movegen)
        // b[] may be read up to 20 bytes outside the board for targets that are then discarded (the caller keeps that
        // much readable memory on both sides)
        for (int sl = 0; sl < nsl; ++sl) {
            const uint32_t e = row[sl];
            const int dr = (int)(int8_t)e, dc = (int)(int8_t)(e >> 8), legoff = (int)(int8_t)(e >> 16), delta = (int)e >> 24;
            const int to = from + delta;
            const int legv = b[from + legoff];                        // elephant eye / horse leg
            const int tgt = b[to];
            const bool ok = ((unsigned)(r + dr - rlo) <= (unsigned)rspan) & ((unsigned)(c + dc - clo) <= (unsigned)cspan) &
                            ((legoff == 0) | (legv == 0)) & !(tgt * side > 0);
            if (ok) list[m] = (uint16_t)(from7 | to);
            m += ok ? 1 : 0;
        }
    }
    return m;
}

// Ordered legal moves of `side` on b[] (mutated during the call, restored on return).  list = kListCap entries: the
// pseudo-legal moves of a round, compacted in place into the action ids from*90+to of the legal ones (a legal move is
// written at or before the entry it was read from).  A board with more than kListCap - 17 pseudo-legal moves (piece
// crowds no game reaches) takes further rounds; before each of them the ids staged so far are flushed to out[] (the
// position's row of the output array).  On return *flushed ids are in out[0 .. *flushed), the following
// min(n, 128) - *flushed ones in list[]; the caller copies those out.  Returns n, or 129 when the position has more
// than 128 legal moves (the caller counts it as an overflow).  *in_check = cy_is_in_check, occ[0..2] = the
// occupied-square set.
// Device: warp-synchronous -- every lane of the warp calls it (see XQT_WARP_MAX above).
XQT_HD int movegen(int8_t* b, int side, uint16_t* list, int16_t* out, int* flushed, int* in_check, const uint32_t* tab, uint32_t* occ)
{
    const Scan s = scan_board(b, side);
    occ[0] = s.occR[0];                            // occupied squares (bit r*9+c): the kernel walks them for the planes
    occ[1] = s.occR[1];
    occ[2] = s.occR[2];
    const int r0 = side == 1 ? 0 : 7;
    int n = 0;                                     // legal moves so far
    int nf = 0;                                    // ... of which already flushed to out[]
    const int ki0 = s.kmask ? ctz32(s.kmask) : 0;
    const int pr0 = (ki0 * 11) >> 5, pc0 = ki0 - pr0 * 3;          // the king every non-king move leaves where it is
    const KingCtx kc0 = king_context(b, s, r0, pr0, pc0, -side);
    const bool have_king0 = s.kmask != 0u;         // no own king in the palace: every move is illegal (pyx:219-224)
    // boards no game reaches -- several own kings, or more than two enemy knights -- get the general test for every move
    const bool general_all = s.n_own_kings > 1 || s.n_ek > 2;
    bool first_round = true;
    bool knight_half = false;                      // the knight on top of the set has had its first four slots
    *in_check = 1;
    uint32_t w0 = s.own[0], w1 = s.own[1], w2 = s.own[2];
    for (;;) {                                     // one round unless a board has more than kListCap - 17 pseudo-legal moves
        const int left = popc32(w0) + popc32(w1) + popc32(w2);
        const int ptrips = XQT_WARP_MAX(left != 0 ? left + s.n_own_knights : 0);   // generation steps: one per piece, two per knight
        if (ptrips == 0 && !first_round) break;
        if (left != 0 && n > nf) {                 // another round for this board: make room
            for (int k = nf; k < n && k < kMaxOut; ++k) out[k] = (int16_t)list[k - nf];
            nf = n < kMaxOut ? n : kMaxOut;
        }
        const int m0 = n - nf;                     // staged legal ids occupy list[0 .. m0)
        int m = m0;
        int kstart = 0, kcount = 0;                // list entries of the king's moves in this round
        for (int t = 0; t < ptrips; ++t) {
            if ((w0 | w1 | w2) != 0u && m + 17 <= kListCap) {
                const int from = w0 ? ctz32(w0) : (w1 ? 32 + ctz32(w1) : 64 + ctz32(w2));
                const bool knight = b[from] == 4 * side;
                const int slot0 = (knight && knight_half) ? 4 : 0;
                knight_half = knight && !knight_half;          // a knight stays in the set for its second step
                if (!knight_half) {
                    if (w0) w0 &= w0 - 1u;
                    else if (w1) w1 &= w1 - 1u;
                    else w2 &= w2 - 1u;
                }
                const int before = m;
                m = gen_piece(b, s, side, from, slot0, list, m, tab);
                if (b[from] == side) {
                    kstart = before;
                    kcount = m - before;
                }
            }
            XQT_RECONVERGE();
        }
        // General pass, all lanes together: the in-check probe of the position (first round, q = -1) and the king's
        // moves (every board has a king with at most 4 moves).  They need the full attack test -- the king's square,
        // lines, knights and pawns all change -- and would otherwise drag that path into nine out of ten iterations of
        // the main loop for one or two lanes.  The verdict replaces the list entry: 0x8000 | id (legal), 0xc000 (not).
        const int gstart = general_all ? m0 : kstart;
        const int gcount = general_all ? m - m0 : kcount;
        const int gtrips = XQT_WARP_MAX(gcount);
        for (int q = first_round ? -1 : 0; q < gtrips; ++q) {
            if (q < gcount) {
                const bool probe = q < 0;
                const int mv = probe ? 0 : (list[gstart + q] & 0x3fff);
                const int from = mv >> 7, to = mv & 127;
                const int fr = div9(from), fc = from - fr * 9, tr = div9(to), tc = to - tr * 9;
                const int8_t mover = b[from], taken = b[to];
                if (!probe) {
                    b[to] = mover;
                    b[from] = 0;
                }
                int pr = pr0, pc = pc0;
                bool have_king = have_king0;
                if (!probe && mover == side) {     // the king "found" afterwards is the first one in palace order (pyx:93-98)
                    uint32_t km = s.kmask;
                    const int fpr = fr - r0;
                    if (fpr >= 0 && fpr <= 2 && fc >= 3 && fc <= 5) km &= ~(1u << (fpr * 3 + fc - 3));
                    km |= 1u << ((tr - r0) * 3 + tc - 3);
                    const int ki = ctz32(km);
                    pr = (ki * 11) >> 5;
                    pc = ki - pr * 3;
                    have_king = true;
                }
                const bool att = attacked(b, s, r0, pr, pc, -side, probe ? -1 : fr, fc, tr, tc, (!probe && taken != 0) ? to : -1);
                if (probe) {
                    *in_check = (att || !have_king0) ? 1 : 0;      // pyx:552-554: no king in the palace counts as check
                } else {
                    b[from] = mover;
                    b[to] = taken;
                    list[gstart + q] = (!att && have_king) ? (uint16_t)(0x8000 | (from * 90 + to)) : (uint16_t)0xc000;
                }
            }
            XQT_RECONVERGE();
        }
        first_round = false;
        const int ltrips = XQT_WARP_MAX(m - m0);
        int i = m0;
        for (int t = 0; t < ltrips; ++t) {
            if (i < m) {
                const int e = list[i++];
                bool ok;
                int id;
                if (e & 0x8000) {                  // judged in the general pass
                    ok = (e & 0x4000) == 0;
                    id = e & 0x1fff;
                } else {
                    const int from = e >> 7, to = e & 127;
                    const int fr = div9(from), fc = from - fr * 9, tr = div9(to), tc = to - tr * 9;
                    ok = !attacked_fixed(b, s, kc0, from, fr, fc, to, tr, tc) && have_king0;    // overlay: the board is not touched
                    id = from * 90 + to;
                }
                if (ok) {
                    if (n < kMaxOut) list[n - nf] = (uint16_t)id;
                    if (++n > kMaxOut) {           // overflow: stop this board (the warp-uniform loops run on, idle)
                        i = m;
                        w0 = w1 = w2 = 0u;
                    }
                }
            }
            XQT_RECONVERGE();
        }
    }
    *flushed = nf;
    return n;
}

}  // namespace xqt
