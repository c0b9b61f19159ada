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
// 32 boards take, which the structure below keeps small:
//   1. one unrolled 90-cell scan builds, in registers, the occupancy bitboards (row-major and
//      column-major 96-bit sets), the own-piece set, the own kings standing in their palace and the
//      enemy knights;
//   2. pseudo-legal generation walks the own-piece set in square order; the five leapers (king,
//      advisor, elephant, knight, pawn) share one table-driven path (per-kind slot word: dr, dc; the
//      leg is (dr/2, dc/2); a per-kind box bounds the target), rook and cannon share one path in which
//      a ray is a bit scan of the row / column occupancy (first blocker = rook capture, second =
//      cannon capture) and only the emission of the empty run is a loop;
//   3. ONE uniform loop tests every pseudo-legal move: make the move in place, take the own king from
//      the palace set, run the attack test on the moved board (rays = bit scans of the king's row and
//      column with the move overlaid on the two masks; knights = the enemy-knight list; pawns = three
//      cells), unmake.  Facing kings need no extra clause: _is_attacked counts the enemy king as a
//      rook on an open ray (pyx:117), which is the flying-general test of pyx:226-240.
// The move list is compacted in place (legal count <= tested count), so the per-thread scratch is one
// array that ends up holding the action ids the kernel copies out with coalesced stores.
//
// The file is plain C++ (no warp intrinsics): tests/test_tpb_cpu.py compiles it with g++ and checks it
// against the oracle over random-playout and piece-soup positions, so the rules logic the kernel runs
// is verified on the CPU build box too; the kernel wrapper is xq_movegen.cu:movegen_tpb_kernel.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define XQT_HD __host__ __device__ __forceinline__
#else
#define XQT_HD inline
#endif

namespace xqt {

constexpr int kListCap = 148;   // per-board scratch entries: 128 outputs + room for one more piece (a rook or cannon: 17 targets) + 3
constexpr int kMaxOut = 128;

XQT_HD int ctz32(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
XQT_HD int top32(uint32_t v)   // index of the highest set bit (v != 0)
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
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
    int n_ek;           // enemy knights on the board; the first two squares in ek0/ek1
    int ek0, ek1;
};

XQT_HD Scan scan_board(const int8_t* b, int side)
{
    Scan s;
    s.occR[0] = s.occR[1] = s.occR[2] = 0u;
    s.occC[0] = s.occC[1] = s.occC[2] = 0u;
    s.own[0] = s.own[1] = s.own[2] = 0u;
    s.kmask = 0u;
    s.n_ek = 0;
    s.ek0 = s.ek1 = -1;
    const int r0 = side == 1 ? 0 : 7;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int sq = 0; sq < 90; ++sq) {
        const int r = sq / 9, c = sq % 9;          // compile-time after unrolling
        const int p = b[sq];
        const uint32_t nz = p != 0 ? 1u : 0u;
        const uint32_t mine = (p * side > 0) ? 1u : 0u;
        s.occR[sq >> 5] |= nz << (sq & 31);
        const int cq = c * 10 + r;
        s.occC[cq >> 5] |= nz << (cq & 31);
        s.own[sq >> 5] |= mine << (sq & 31);
        if (c >= 3 && c <= 5) {                    // palace columns; the row test depends on the side
            const int pr = r - r0;
            if (pr >= 0 && pr <= 2 && p == side) s.kmask |= 1u << (pr * 3 + (c - 3));
        }
        if (p == -4 * side) {
            if (s.n_ek == 0) s.ek0 = sq;
            else if (s.n_ek == 1) s.ek1 = sq;
            ++s.n_ek;
        }
    }
    return s;
}

// pyx:104-189 on the board as it stands in b[] (a move, if any, already made in place); the row / column occupancy
// of the un-moved board is corrected for the move (fr, fc) -> (tr, tc) (fr < 0: no move).  `cap` is the square a
// captured piece stood on (an enemy knight there no longer attacks), -1 for none.
XQT_HD bool attacked(const int8_t* b, const Scan& s, int kr, int kc, int by, int fr, int fc, int tr, int tc, int cap)
{
    const int rook = 5 * by, cannon = 6 * by, horse = 4 * by, pawn = 7 * by, king = by;
    uint32_t R = bits96(s.occR[0], s.occR[1], s.occR[2], kr * 9, 9);
    uint32_t C = bits96(s.occC[0], s.occC[1], s.occC[2], kc * 10, 10);
    if (fr >= 0) {
        if (fr == kr) R &= ~(1u << fc);
        if (fc == kc) C &= ~(1u << fr);
        if (tr == kr) R |= 1u << tc;
        if (tc == kc) C |= 1u << tr;
    }
    const int8_t* row = b + kr * 9;
    uint32_t m = R & ((1u << kc) - 1u);                       // towards column 0: nearest = highest bit
    if (m) {
        const int c1 = top32(m);
        const int p1 = row[c1];
        if (p1 == rook || p1 == king) return true;
        m ^= 1u << c1;
        if (m && row[top32(m)] == cannon) return true;
    }
    m = R >> (kc + 1);                                        // towards column 8: nearest = lowest bit
    if (m) {
        const int p1 = row[kc + 1 + ctz32(m)];
        if (p1 == rook || p1 == king) return true;
        m &= m - 1u;
        if (m && row[kc + 1 + ctz32(m)] == cannon) return true;
    }
    m = C & ((1u << kr) - 1u);                                // towards row 0
    if (m) {
        const int r1 = top32(m);
        const int p1 = b[r1 * 9 + kc];
        if (p1 == rook || p1 == king) return true;
        m ^= 1u << r1;
        if (m && b[top32(m) * 9 + kc] == cannon) return true;
    }
    m = C >> (kr + 1);                                        // towards row 9
    if (m) {
        const int p1 = b[(kr + 1 + ctz32(m)) * 9 + kc];
        if (p1 == rook || p1 == king) return true;
        m &= m - 1u;
        if (m && b[(kr + 1 + ctz32(m)) * 9 + kc] == cannon) return true;
    }
    // knights (pyx:156-169): the leg is the cell next to the knight on its long axis
    if (s.n_ek <= 2) {
        for (int j = 0; j < s.n_ek; ++j) {
            const int nsq = j == 0 ? s.ek0 : s.ek1;
            if (nsq == cap) continue;
            const int nr = nsq / 9, nc = nsq - nr * 9;
            const int dr = kr - nr, dc = kc - nc;
            const int adr = dr < 0 ? -dr : dr, adc = dc < 0 ? -dc : dc;
            int leg;
            if (adr == 2 && adc == 1) leg = nsq + (dr / 2) * 9;
            else if (adr == 1 && adc == 2) leg = nsq + dc / 2;
            else continue;
            if (b[leg] == 0) return true;
        }
    } else {
        for (int i = 0; i < 8; ++i) {                          // boards no game reaches: look at the 8 origins
            const int jr = (i < 4) ? ((i < 2) ? -2 : 2) : ((i < 6) ? -1 : 1);
            const int jc = (i < 4) ? ((i & 1) ? 1 : -1) : ((i & 1) ? 2 : -2);
            const int nr = kr + jr, nc = kc + jc;
            if (nr < 0 || nr >= 10 || nc < 0 || nc >= 9) continue;
            if (b[nr * 9 + nc] != horse) continue;
            int lr = nr, lc = nc;
            if (jr == 2 || jr == -2) lr = nr - jr / 2; else lc = nc - jc / 2;
            if (b[lr * 9 + lc] == 0) return true;
        }
    }
    // pawns (pyx:172-187): from behind, and from the side once the target row is across the river for that colour
    if (by == 1) {
        if (kr >= 1 && row[kc - 9] == pawn) return true;
        if (kr >= 5) {
            if (kc >= 1 && row[kc - 1] == pawn) return true;
            if (kc <= 7 && row[kc + 1] == pawn) return true;
        }
    } else {
        if (kr <= 8 && row[kc + 9] == pawn) return true;
        if (kr <= 4) {
            if (kc >= 1 && row[kc - 1] == pawn) return true;
            if (kc <= 7 && row[kc + 1] == pawn) return true;
        }
    }
    return false;
}

// leaper slot words: slot i in bits 8i..8i+7 = (dr+2) | (dc+2) << 3 | 0x40 (valid).  Orders: pyx:287-367, 434-484.
#define XQT_SLOT(dr, dc) ((uint64_t)(((dr) + 2) | (((dc) + 2) << 3) | 0x40))
#define XQT_SLOTS4(a, b, c, d) ((a) | ((b) << 8) | ((c) << 16) | ((d) << 24))
constexpr uint64_t kSlotsKing = XQT_SLOTS4(XQT_SLOT(-1, 0), XQT_SLOT(1, 0), XQT_SLOT(0, -1), XQT_SLOT(0, 1));
constexpr uint64_t kSlotsAdvisor = XQT_SLOTS4(XQT_SLOT(-1, -1), XQT_SLOT(-1, 1), XQT_SLOT(1, -1), XQT_SLOT(1, 1));
constexpr uint64_t kSlotsElephant = XQT_SLOTS4(XQT_SLOT(-2, -2), XQT_SLOT(-2, 2), XQT_SLOT(2, -2), XQT_SLOT(2, 2));
constexpr uint64_t kSlotsKnight = XQT_SLOTS4(XQT_SLOT(-2, -1), XQT_SLOT(-2, 1), XQT_SLOT(2, -1), XQT_SLOT(2, 1)) |
                                  (XQT_SLOTS4(XQT_SLOT(-1, -2), XQT_SLOT(-1, 2), XQT_SLOT(1, -2), XQT_SLOT(1, 2)) << 32);
constexpr uint64_t kSlotsPawnRed = XQT_SLOT(1, 0) | (XQT_SLOT(0, -1) << 8) | (XQT_SLOT(0, 1) << 16);
constexpr uint64_t kSlotsPawnBlack = XQT_SLOT(-1, 0) | (XQT_SLOT(0, -1) << 8) | (XQT_SLOT(0, 1) << 16);

// Pseudo-legal targets of the piece on `from` appended to list[m...] as from << 7 | to; returns the new m.
XQT_HD int gen_piece(const int8_t* b, const Scan& s, int side, int from, uint16_t* list, int m)
{
    const int p = b[from];
    const int kind = p < 0 ? -p : p;
    const int r = from / 9, c = from - r * 9;
    if (kind == 5 || kind == 6) {
        const uint32_t Rm = bits96(s.occR[0], s.occR[1], s.occR[2], r * 9, 9);
        const uint32_t Cm = bits96(s.occC[0], s.occC[1], s.occC[2], c * 10, 10);
        // pyx:42-46 order: up (row - 1), down (row + 1), left (col - 1), right (col + 1)
        for (int d = 0; d < 4; ++d) {
            const bool vertical = d < 2, neg = (d & 1) == 0;
            const uint32_t line = vertical ? Cm : Rm;
            const int pos = vertical ? r : c;
            const int len = vertical ? 10 : 9;
            const int step = vertical ? 9 : 1;
            int n_empty, first = -1, second = -1;
            if (neg) {
                uint32_t q = line & ((1u << pos) - 1u);
                n_empty = pos;
                if (q) {
                    first = top32(q);
                    n_empty = pos - 1 - first;
                    q ^= 1u << first;
                    if (q) second = top32(q);
                }
            } else {
                uint32_t q = line >> (pos + 1);
                n_empty = len - 1 - pos;
                if (q) {
                    first = pos + 1 + ctz32(q);
                    n_empty = first - pos - 1;
                    q &= q - 1u;
                    if (q) second = pos + 1 + ctz32(q);
                }
            }
            const int sgn = neg ? -step : step;
            int to = from;
            for (int i = 0; i < n_empty; ++i) {
                to += sgn;
                list[m++] = (uint16_t)(from << 7 | to);
            }
            const int hit = kind == 5 ? first : second;       // rook takes the first piece on the ray, cannon the second
            if (hit >= 0) {
                const int tsq = from + (hit - pos) * step;
                if (b[tsq] * side < 0) list[m++] = (uint16_t)(from << 7 | tsq);
            }
        }
        return m;
    }
    uint64_t slots;
    int rlo = 0, rhi = 9, clo = 0, chi = 8;
    if (kind == 1 || kind == 2) {                              // palace box (the advisor test of pyx:307-326 is the box too)
        slots = kind == 1 ? kSlotsKing : kSlotsAdvisor;
        rlo = side == 1 ? 0 : 7;
        rhi = rlo + 2;
        clo = 3;
        chi = 5;
    } else if (kind == 3) {                                    // own half of the board
        slots = kSlotsElephant;
        rlo = side == 1 ? 0 : 5;
        rhi = rlo + 4;
    } else if (kind == 4) {
        slots = kSlotsKnight;
    } else if (kind == 7) {
        slots = side == 1 ? kSlotsPawnRed : kSlotsPawnBlack;
        const bool crossed = side == 1 ? r >= 5 : r <= 4;
        if (!crossed) slots &= 0xffu;
    } else {
        return m;
    }
    for (; slots; slots >>= 8) {
        const int e = (int)(slots & 0xffu);
        const int dr = (e & 7) - 2, dc = ((e >> 3) & 7) - 2;
        const int nr = r + dr, nc = c + dc;
        if (nr < rlo || nr > rhi || nc < clo || nc > chi) continue;
        if (dr == 2 || dr == -2 || dc == 2 || dc == -2)
            if (b[from + (dr / 2) * 9 + dc / 2] != 0) continue;          // elephant eye / horse leg
        const int to = nr * 9 + nc;
        if (b[to] * side > 0) continue;
        list[m++] = (uint16_t)(from << 7 | to);
    }
    return m;
}

// Ordered legal moves of `side` on b[] (mutated during the call, restored on return).  list must hold kListCap
// entries; on return list[0 .. min(n, 128)) are action ids from*90+to.  Returns n, or 129 when the position has more
// than 128 legal moves (no game reaches that; the caller counts it as an overflow).  *in_check = cy_is_in_check.
XQT_HD int movegen(int8_t* b, int side, uint16_t* list, int* in_check)
{
    const Scan s = scan_board(b, side);
    const int r0 = side == 1 ? 0 : 7;
    if (s.kmask) {
        const int ki = ctz32(s.kmask);
        const int kdiv = (ki * 11) >> 5;           // ki / 3 for ki < 9
        *in_check = attacked(b, s, r0 + kdiv, 3 + ki - kdiv * 3, -side, -1, 0, 0, 0, -1) ? 1 : 0;
    } else {
        *in_check = 1;                             // pyx:552-554: no king in the palace counts as check
    }
    int n = 0;                                     // legal moves so far = list[0..n)
    uint32_t w0 = s.own[0], w1 = s.own[1], w2 = s.own[2];
    while ((w0 | w1 | w2) != 0u) {
        int m = n;
        while ((w0 | w1 | w2) != 0u && m + 17 <= kListCap) {
            int from;
            if (w0) { from = ctz32(w0); w0 &= w0 - 1u; }
            else if (w1) { from = 32 + ctz32(w1); w1 &= w1 - 1u; }
            else { from = 64 + ctz32(w2); w2 &= w2 - 1u; }
            m = gen_piece(b, s, side, from, list, m);
        }
        for (int i = n; i < m; ++i) {
            const int mv = list[i];
            const int from = mv >> 7, to = mv & 127;
            const int fr = from / 9, fc = from - fr * 9, tr = to / 9, tc = to - tr * 9;
            const int8_t mover = b[from], taken = b[to];
            b[to] = mover;
            b[from] = 0;
            uint32_t km = s.kmask;
            if (mover == side) {                   // a king move (targets are always inside the palace box)
                const int pr = fr - r0;
                if (pr >= 0 && pr <= 2 && fc >= 3 && fc <= 5) km &= ~(1u << (pr * 3 + fc - 3));
                km |= 1u << ((tr - r0) * 3 + tc - 3);
            }
            bool ok = false;
            if (km) {
                const int ki = ctz32(km);
                const int kdiv = (ki * 11) >> 5;
                ok = !attacked(b, s, r0 + kdiv, 3 + ki - kdiv * 3, -side, fr, fc, tr, tc, taken != 0 ? to : -1);
            }
            b[from] = mover;
            b[to] = taken;
            if (ok) {
                if (n < kMaxOut) list[n] = (uint16_t)(from * 90 + to);
                ++n;
                if (n > kMaxOut) return kMaxOut + 1;
            }
        }
    }
    return n;
}

}  // namespace xqt
