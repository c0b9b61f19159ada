/*
 * xq_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's self-play hot path, used only as
 * the checker for the CUDA kernels (tests/, __graft_entry__.smoke(), and the
 * cpu_baseline / --impl reference legs of bench.py).  Nothing under
 * xiangqi-alphazero_b200/ may link or call this file.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 *   - the reference's own known answers (test_v3.py:115-197, test_cython.py:46-138),
 *   - perft(1..3) = 44 / 1920 / 79666,
 *   - tests/golden/ fixtures produced by importing the reference (game.py with
 *     the rebuilt Cython engine, mcts.py) in the build container
 *     (tests/golden/make_golden.py),
 *   - oracle/_ref/game_core*.so (the reference .pyx compiled as-is) when present.
 *
 * What each function follows (paths relative to /root/reference/training):
 *   xqo_find_king        cython_engine/game_core.pyx:78-101   (palace-only scan)
 *   xqo_is_attacked      cython_engine/game_core.pyx:104-189
 *   xqo_move_is_legal    cython_engine/game_core.pyx:209-252
 *   xqo_generate_moves   cython_engine/game_core.pyx:262-486  (order = child order)
 *   xqo_in_check         cython_engine/game_core.pyx:543-555
 *   xqo_planes           game.py:618-640
 *   xqo_game_*           game.py:528-545 (make_move), 552-563, 565-616 (is_game_over)
 *   xqo_mcts_search      mcts.py:21-73, 94-155, 176-206 with the NumPy-2 scalar
 *                        promotion described in SURVEY.md appendix A.4
 *
 * Board: int8[90], index r*9+c, r=0 red back rank; red>0, black<0;
 * 1 K, 2 A, 3 B, 4 N, 5 R, 6 C, 7 P (game.py:50-65).
 * Action id = from*90 + to (game.py:112-114).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared  (no -ffast-math: the MCTS part
 * depends on IEEE float32/float64 evaluation order).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define NR 10
#define NC 9
#define NSQ 90
#define XQO_MAX_MOVES 200 /* game_core.pyx:50 */

#define AT(b, r, c) ((b)[(r) * NC + (c)])

static inline int on_board(int r, int c) { return r >= 0 && r < NR && c >= 0 && c < NC; }
static inline int own(int p, int pl) { return pl == 1 ? p > 0 : p < 0; }
static inline int foe(int p, int pl) { return pl == 1 ? p < 0 : p > 0; }
static inline int may_land(int p, int pl) { return p == 0 || foe(p, pl); }

/* orthogonal steps in the reference order up, down, left, right (pyx:42-46) */
static const int ORTHO[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
/* knight jump + leg, reference order (pyx:31-39) */
static const int HORSE[8][4] = {{-2, -1, -1, 0}, {-2, 1, -1, 0}, {2, -1, 1, 0}, {2, 1, 1, 0},
                                {-1, -2, 0, -1}, {-1, 2, 0, 1},  {1, -2, 0, -1}, {1, 2, 0, 1}};

/* pyx:78-101 -- returns square or -1; only the side's own 3x3 palace is searched */
int xqo_find_king(const int8_t *b, int player)
{
    int want = player == 1 ? 1 : -1;
    int r0 = player == 1 ? 0 : 7;
    for (int r = r0; r < r0 + 3; ++r)
        for (int c = 3; c <= 5; ++c)
            if (AT(b, r, c) == want) return r * NC + c;
    return -1;
}

/* pyx:104-189 */
int xqo_is_attacked(const int8_t *b, int kr, int kc, int by)
{
    const int rook = 5 * by, cannon = 6 * by, horse = 4 * by, pawn = 7 * by, king = 1 * by;
    /* rook or king on an open ray (pyx:121-133) */
    for (int d = 0; d < 4; ++d) {
        int r = kr + ORTHO[d][0], c = kc + ORTHO[d][1];
        while (on_board(r, c)) {
            int p = AT(b, r, c);
            if (p != 0) {
                if (p == rook || p == king) return 1;
                break;
            }
            r += ORTHO[d][0];
            c += ORTHO[d][1];
        }
    }
    /* cannon behind exactly one screen (pyx:136-153) */
    for (int d = 0; d < 4; ++d) {
        int r = kr + ORTHO[d][0], c = kc + ORTHO[d][1];
        int screen = 0;
        while (on_board(r, c)) {
            int p = AT(b, r, c);
            if (p != 0) {
                if (screen) {
                    if (p == cannon) return 1;
                    break;
                }
                screen = 1;
            }
            r += ORTHO[d][0];
            c += ORTHO[d][1];
        }
    }
    /* knights, leg measured from the knight's side (pyx:156-169) */
    for (int i = 0; i < 8; ++i) {
        int nr = kr + HORSE[i][0], nc = kc + HORSE[i][1];
        if (!on_board(nr, nc) || AT(b, nr, nc) != horse) continue;
        int mdr = kr - nr, mdc = kc - nc, lr, lc;
        if (mdr == 2 || mdr == -2) {
            lr = nr + mdr / 2;
            lc = nc;
        } else {
            lr = nr;
            lc = nc + mdc / 2;
        }
        if (AT(b, lr, lc) == 0) return 1;
    }
    /* pawns (pyx:172-187) */
    if (by == 1) {
        if (kr - 1 >= 0 && AT(b, kr - 1, kc) == pawn) return 1;
        if (kr >= 5) {
            if (kc - 1 >= 0 && AT(b, kr, kc - 1) == pawn) return 1;
            if (kc + 1 < NC && AT(b, kr, kc + 1) == pawn) return 1;
        }
    } else {
        if (kr + 1 < NR && AT(b, kr + 1, kc) == pawn) return 1;
        if (kr <= 4) {
            if (kc - 1 >= 0 && AT(b, kr, kc - 1) == pawn) return 1;
            if (kc + 1 < NC && AT(b, kr, kc + 1) == pawn) return 1;
        }
    }
    return 0;
}

/* pyx:543-555 -- a missing king counts as "in check" */
int xqo_in_check(const int8_t *b, int player)
{
    int k = xqo_find_king(b, player);
    if (k < 0) return 1;
    return xqo_is_attacked(b, k / NC, k % NC, -player);
}

/* pyx:209-252 -- board is mutated and restored */
static int move_is_legal(int8_t *b, int from, int to, int player)
{
    int8_t mover = b[from], taken = b[to];
    int ok = 1;
    b[to] = mover;
    b[from] = 0;
    int k = xqo_find_king(b, player);
    if (k < 0) {
        ok = 0;
    } else {
        int kr = k / NC, kc = k % NC;
        int e = xqo_find_king(b, -player);
        if (e >= 0 && e % NC == kc) {
            int er = e / NC;
            int lo = kr < er ? kr : er, hi = kr < er ? er : kr;
            int blocked = 0;
            for (int r = lo + 1; r < hi; ++r)
                if (AT(b, r, kc) != 0) {
                    blocked = 1;
                    break;
                }
            if (!blocked) ok = 0; /* flying general */
        }
        if (ok) ok = !xqo_is_attacked(b, kr, kc, -player);
    }
    b[from] = mover;
    b[to] = taken;
    return ok;
}

int xqo_move_is_legal(const int8_t *board, int from, int to, int player)
{
    int8_t b[NSQ];
    memcpy(b, board, NSQ);
    return move_is_legal(b, from, to, player);
}

#define TRY(tr, tc)                                                      \
    do {                                                                 \
        int to_ = (tr) * NC + (tc);                                      \
        if (move_is_legal(b, from, to_, player)) {                       \
            if (n < XQO_MAX_MOVES) out[n] = (int16_t)(from * NSQ + to_); \
            ++n;                                                         \
        }                                                                \
    } while (0)

/* pyx:262-486 -- emits action ids in the reference's generation order */
int xqo_generate_moves(const int8_t *board, int player, int16_t *out)
{
    int8_t b[NSQ];
    memcpy(b, board, NSQ);
    int n = 0;
    for (int r = 0; r < NR; ++r)
        for (int c = 0; c < NC; ++c) {
            int p = AT(b, r, c);
            if (p == 0 || !own(p, player)) continue;
            int from = r * NC + c;
            int kind = p > 0 ? p : -p;
            switch (kind) {
            case 1: { /* king: pyx:287-304 */
                int lo = player == 1 ? 0 : 7, hi = lo + 2;
                for (int d = 0; d < 4; ++d) {
                    int nr = r + ORTHO[d][0], nc = c + ORTHO[d][1];
                    if (nr < lo || nr > hi || nc < 3 || nc > 5) continue;
                    if (may_land(AT(b, nr, nc), player)) TRY(nr, nc);
                }
                break;
            }
            case 2: /* advisor: pyx:307-326 (palace box only, no advisor-point test) */
                for (int dr = -1; dr <= 1; dr += 2)
                    for (int dc = -1; dc <= 1; dc += 2) {
                        int nr = r + dr, nc = c + dc;
                        if (!on_board(nr, nc) || nc < 3 || nc > 5) continue;
                        if (player == 1 && nr > 2) continue;
                        if (player == -1 && nr < 7) continue;
                        if (may_land(AT(b, nr, nc), player)) TRY(nr, nc);
                    }
                break;
            case 3: /* elephant: pyx:329-349 */
                for (int dr = -2; dr <= 2; dr += 4)
                    for (int dc = -2; dc <= 2; dc += 4) {
                        int nr = r + dr, nc = c + dc;
                        if (!on_board(nr, nc)) continue;
                        if (player == 1 && nr > 4) continue;
                        if (player == -1 && nr < 5) continue;
                        if (AT(b, r + dr / 2, c + dc / 2) != 0) continue;
                        if (may_land(AT(b, nr, nc), player)) TRY(nr, nc);
                    }
                break;
            case 4: /* knight: pyx:352-367 */
                for (int i = 0; i < 8; ++i) {
                    int nr = r + HORSE[i][0], nc = c + HORSE[i][1];
                    if (!on_board(nr, nc)) continue;
                    if (AT(b, r + HORSE[i][2], c + HORSE[i][3]) != 0) continue;
                    if (may_land(AT(b, nr, nc), player)) TRY(nr, nc);
                }
                break;
            case 5: /* rook: pyx:370-396 */
                for (int d = 0; d < 4; ++d) {
                    int nr = r + ORTHO[d][0], nc = c + ORTHO[d][1];
                    while (on_board(nr, nc)) {
                        int q = AT(b, nr, nc);
                        if (q == 0) {
                            TRY(nr, nc);
                        } else {
                            if (foe(q, player)) TRY(nr, nc);
                            break;
                        }
                        nr += ORTHO[d][0];
                        nc += ORTHO[d][1];
                    }
                }
                break;
            case 6: /* cannon: pyx:399-431 */
                for (int d = 0; d < 4; ++d) {
                    int nr = r + ORTHO[d][0], nc = c + ORTHO[d][1];
                    while (on_board(nr, nc) && AT(b, nr, nc) == 0) {
                        TRY(nr, nc);
                        nr += ORTHO[d][0];
                        nc += ORTHO[d][1];
                    }
                    if (on_board(nr, nc)) { /* hop the screen */
                        nr += ORTHO[d][0];
                        nc += ORTHO[d][1];
                        while (on_board(nr, nc)) {
                            int q = AT(b, nr, nc);
                            if (q != 0) {
                                if (foe(q, player)) TRY(nr, nc);
                                break;
                            }
                            nr += ORTHO[d][0];
                            nc += ORTHO[d][1];
                        }
                    }
                }
                break;
            case 7: { /* pawn: pyx:434-484 */
                int fwd = player == 1 ? 1 : -1;
                int crossed = player == 1 ? r >= 5 : r <= 4;
                int nr = r + fwd;
                if (nr >= 0 && nr < NR && may_land(AT(b, nr, c), player)) TRY(nr, c);
                if (crossed) {
                    if (c - 1 >= 0 && may_land(AT(b, r, c - 1), player)) TRY(r, c - 1);
                    if (c + 1 < NC && may_land(AT(b, r, c + 1), player)) TRY(r, c + 1);
                }
                break;
            }
            default:
                break;
            }
        }
    return n;
}

/* game.py:618-640 -- 15 planes, own pieces 0..6, other side 7..13, plane 14 = red to move */
void xqo_planes(const int8_t *b, int player, float *out)
{
    memset(out, 0, sizeof(float) * 15 * NSQ);
    for (int s = 0; s < NSQ; ++s) {
        int p = b[s];
        if (p == 0) continue;
        int kind = (p > 0 ? p : -p) - 1;
        int mine = own(p, player);
        out[(mine ? kind : 7 + kind) * NSQ + s] = 1.0f;
    }
    if (player == 1)
        for (int s = 0; s < NSQ; ++s) out[14 * NSQ + s] = 1.0f;
}

/* Batched form with the product's output layout: actions[B][128] int16 (unused slots -1),
 * n_moves[B] uint8, in_check[B] uint8, planes[B][15][90] float (nullable).
 * Returns the number of positions whose move count exceeded 128 (must be 0). */
int xqo_movegen_batch(const int8_t *boards, const int8_t *sides, int B, int16_t *actions,
                      uint8_t *n_moves, uint8_t *in_check, float *planes)
{
    int overflow = 0;
    int16_t tmp[XQO_MAX_MOVES];
    for (int i = 0; i < B; ++i) {
        const int8_t *b = boards + (size_t)i * NSQ;
        int pl = sides[i];
        int n = xqo_generate_moves(b, pl, tmp);
        if (n > 128) {
            ++overflow;
            n = 128;
        }
        for (int k = 0; k < 128; ++k) actions[(size_t)i * 128 + k] = k < n ? tmp[k] : (int16_t)-1;
        n_moves[i] = (uint8_t)n;
        in_check[i] = (uint8_t)xqo_in_check(b, pl);
        if (planes) xqo_planes(b, pl, planes + (size_t)i * 15 * NSQ);
    }
    return overflow;
}

void xqo_is_attacked_batch(const int8_t *boards, const uint8_t *sq, const int8_t *by, int B,
                           uint8_t *out)
{
    for (int i = 0; i < B; ++i)
        out[i] = (uint8_t)xqo_is_attacked(boards + (size_t)i * NSQ, sq[i] / NC, sq[i] % NC, by[i]);
}

/* ------------------------------------------------------------------------- */
/* Game state (game.py XiangqiGame: board, current_player, move_count,        */
/* no_capture_count, history).  len(history) == move_count always, and only   */
/* history[-12:] is ever read (game.py:607-614), so a 12-deep ring suffices.  */
/* ------------------------------------------------------------------------- */
typedef struct {
    int8_t board[NSQ];
    int8_t ring[12][NSQ]; /* ring[i % 12] = board before move i */
    int32_t player;
    int32_t move_count;
    int32_t no_capture;
} xqo_game;

static const int8_t START_BACK[9] = {5, 4, 3, 2, 1, 2, 3, 4, 5};

/* game.py:139-159 */
void xqo_game_init(xqo_game *g)
{
    memset(g, 0, sizeof(*g));
    for (int c = 0; c < 9; ++c) {
        g->board[0 * NC + c] = START_BACK[c];
        g->board[9 * NC + c] = (int8_t)-START_BACK[c];
    }
    g->board[2 * NC + 1] = g->board[2 * NC + 7] = 6;
    g->board[7 * NC + 1] = g->board[7 * NC + 7] = -6;
    for (int c = 0; c < 9; c += 2) {
        g->board[3 * NC + c] = 7;
        g->board[6 * NC + c] = -7;
    }
    g->player = 1;
}

/* game.py:528-545 (no legality check, like the reference) */
void xqo_game_move(xqo_game *g, int action)
{
    int from = action / NSQ, to = action % NSQ;
    memcpy(g->ring[g->move_count % 12], g->board, NSQ);
    int taken = g->board[to];
    g->board[to] = g->board[from];
    g->board[from] = 0;
    g->no_capture = taken != 0 ? 0 : g->no_capture + 1;
    g->player = -g->player;
    g->move_count += 1;
}

/* game.py:552-563 with PIECE_VALUES game.py:74 */
static const int PIECE_VAL[8] = {0, 0, 20, 20, 40, 90, 45, 10};
int xqo_material(const int8_t *b, int player)
{
    int s = 0;
    for (int i = 0; i < NSQ; ++i) {
        int p = b[i];
        if (player == 1 && p > 0) s += PIECE_VAL[p];
        if (player == -1 && p < 0) s += PIECE_VAL[-p];
    }
    return s;
}

/* game.py:565-616.  Returns done; *winner in {1,-1,0} when done.
 * If moves_out != NULL the legal action list is written there and *n_out set
 * (the reference caches it the same way for the following expand). */
int xqo_game_over(const xqo_game *g, int *winner, int16_t *moves_out, int *n_out)
{
    int16_t local[XQO_MAX_MOVES];
    int16_t *mv = moves_out ? moves_out : local;
    if (n_out) *n_out = 0;
    if (xqo_find_king(g->board, 1) < 0) {
        *winner = -1;
        return 1;
    }
    if (xqo_find_king(g->board, -1) < 0) {
        *winner = 1;
        return 1;
    }
    int n = xqo_generate_moves(g->board, g->player, mv);
    if (n_out) *n_out = n;
    if (n == 0) {
        *winner = -g->player;
        return 1;
    }
    if (g->no_capture >= 120) {
        *winner = 0;
        return 1;
    }
    if (g->move_count >= 200) {
        int diff = xqo_material(g->board, 1) - xqo_material(g->board, -1);
        *winner = diff > 30 ? 1 : (diff < -30 ? -1 : 0);
        return 1;
    }
    if (g->move_count >= 6) {
        int depth = g->move_count < 12 ? g->move_count : 12;
        int rep = 0;
        for (int k = 1; k <= depth; ++k)
            if (memcmp(g->ring[(g->move_count - k) % 12], g->board, NSQ) == 0) ++rep;
        if (rep >= 3) {
            *winner = 0;
            return 1;
        }
    }
    *winner = 2; /* not over */
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Deterministic PRNG for synthetic playouts (test inputs only).              */
/* ------------------------------------------------------------------------- */
static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* Uniform-random legal playouts from the start position:
 * loop {is_game_over -> record (board, side) -> random legal move}; terminal positions
 * are recorded too.  Fills up to `cap` positions; returns how many were written. */
int xqo_random_playout_positions(uint64_t seed, int cap, int8_t *boards, int8_t *sides)
{
    uint64_t s = seed;
    int w = 0;
    int16_t mv[XQO_MAX_MOVES];
    while (w < cap) {
        xqo_game g;
        xqo_game_init(&g);
        for (;;) {
            int winner, n;
            int done = xqo_game_over(&g, &winner, mv, &n);
            memcpy(boards + (size_t)w * NSQ, g.board, NSQ);
            sides[w] = (int8_t)g.player;
            if (++w >= cap) break;
            if (done) break;
            xqo_game_move(&g, mv[splitmix64(&s) % (uint64_t)n]);
        }
    }
    return w;
}

/* perft over the full game-state rules is not what the reference tests; this is
 * plain move-tree counting on boards (44/1920/79666/3290240). */
uint64_t xqo_perft(const int8_t *board, int player, int depth)
{
    int16_t mv[XQO_MAX_MOVES];
    int n = xqo_generate_moves(board, player, mv);
    if (depth <= 1) return (uint64_t)n;
    uint64_t tot = 0;
    int8_t b[NSQ];
    for (int i = 0; i < n; ++i) {
        memcpy(b, board, NSQ);
        int from = mv[i] / NSQ, to = mv[i] % NSQ;
        b[to] = b[from];
        b[from] = 0;
        tot += xqo_perft(b, -player, depth - 1);
    }
    return tot;
}

/* ------------------------------------------------------------------------- */
/* MCTS (mcts.py).                                                            */
/*                                                                            */
/* Evaluator contract (model.predict, model.py:109-124): state -> (float32    */
/* probs[8100], python float value).  Here: callback fills probs and returns  */
/* the value as double (value.item() of a float32 is exactly that float).     */
/* ------------------------------------------------------------------------- */
typedef double (*xqo_eval_fn)(const int8_t *board, int player, float *probs8100, void *user);

typedef struct {
    int32_t parent;      /* -1 for root */
    int32_t first_child; /* -1 while leaf (mcts.py:40 is_leaf) */
    int32_t n_child;
    int32_t visits;      /* visit_count */
    double total;        /* total_value: python float */
    double prior64;      /* prior when it is a python float / np.float64 */
    float prior32;       /* prior when it is np.float32 */
    int16_t action;
    int8_t wide;         /* 1: prior is 64-bit => UCB evaluated in double (A.4) */
} xqo_node;

typedef struct {
    xqo_node *nodes;
    int32_t count, cap;
} xqo_tree;

static int32_t tree_alloc(xqo_tree *t, int n)
{
    if (t->count + n > t->cap) {
        int32_t nc = t->cap ? t->cap * 2 : 4096;
        while (nc < t->count + n) nc *= 2;
        t->nodes = (xqo_node *)realloc(t->nodes, sizeof(xqo_node) * (size_t)nc);
        t->cap = nc;
    }
    int32_t at = t->count;
    t->count += n;
    return at;
}

/* mcts.py:176-188 + 60-64: children created in legal-move order.
 * noise==NULL: priors stay np.float32 (or python float 1/n if the legal mass is <= 0).
 * noise!=NULL: 0.75*P32 + 0.25*noise -> float64 (mcts.py:117-121). */
static void expand(xqo_tree *t, int32_t node, const float *probs, const int16_t *acts, int n,
                   const double *noise)
{
    float sum = 0.0f; /* sum() over np.float32 scalars: sequential float32 adds */
    for (int i = 0; i < n; ++i) sum = sum + probs[acts[i]];
    int32_t base = tree_alloc(t, n);
    t->nodes[node].first_child = base;
    t->nodes[node].n_child = n;
    for (int i = 0; i < n; ++i) {
        xqo_node *c = &t->nodes[base + i];
        c->parent = node;
        c->first_child = -1;
        c->n_child = 0;
        c->visits = 0;
        c->total = 0.0;
        c->action = acts[i];
        if (sum > 0.0f) {
            float p = probs[acts[i]] / sum;
            if (noise) {
                float a = 0.75f * p; /* python float * np.float32 -> np.float32 */
                c->prior64 = (double)a + 0.25 * noise[i];
                c->prior32 = 0.0f;
                c->wide = 1;
            } else {
                c->prior32 = p;
                c->prior64 = 0.0;
                c->wide = 0;
            }
        } else {
            double u = 1.0 / (double)n; /* python float */
            c->prior64 = noise ? 0.75 * u + 0.25 * noise[i] : u;
            c->prior32 = 0.0f;
            c->wide = 1;
        }
    }
}

/* mcts.py:43-58 with NEP-50 promotion */
static int32_t select_child(const xqo_tree *t, int32_t node, double c_puct)
{
    const xqo_node *nd = &t->nodes[node];
    double sqrt_parent = sqrt((double)nd->visits);
    int32_t best = -1;
    double best_score = -INFINITY; /* float32 scores compare exactly after widening */
    for (int i = 0; i < nd->n_child; ++i) {
        const xqo_node *c = &t->nodes[nd->first_child + i];
        double q = c->visits == 0 ? 0.0 : c->total / (double)c->visits;
        double score;
        if (c->wide) {
            score = q + c_puct * c->prior64 * sqrt_parent / (double)(1 + c->visits);
        } else {
            float u = (float)c_puct * c->prior32;
            u = u * (float)sqrt_parent;
            u = u / (float)(1 + c->visits);
            score = (double)((float)q + u);
        }
        if (score > best_score) {
            best_score = score;
            best = nd->first_child + i;
        }
    }
    return best;
}

/* mcts.py:94-155.  Output: visit counts per root child in child order
 * (root_actions[n_root], root_visits[n_root]); returns n_root (0 if no legal move).
 * stats_out (nullable): [0]=nodes allocated, [1]=terminal-leaf sims, [2]=max depth. */
int xqo_mcts_search(const xqo_game *game, int num_sims, double c_puct, xqo_eval_fn eval, void *user,
                    const double *root_noise, int16_t *root_actions, int32_t *root_visits,
                    double *root_total_out, int64_t *stats_out)
{
    float *probs = (float *)malloc(sizeof(float) * 8100);
    xqo_tree t = {0, 0, 0};
    int16_t mv[XQO_MAX_MOVES];
    int64_t terminal_sims = 0, max_depth = 0;

    int32_t root = tree_alloc(&t, 1);
    memset(&t.nodes[root], 0, sizeof(xqo_node));
    t.nodes[root].parent = -1;
    t.nodes[root].first_child = -1;

    eval(game->board, game->player, probs, user); /* root value is discarded (mcts.py:108) */
    int n0 = xqo_generate_moves(game->board, game->player, mv);
    if (n0 == 0) {
        free(probs);
        free(t.nodes);
        return 0;
    }
    expand(&t, root, probs, mv, n0, root_noise);

    for (int s = 0; s < num_sims; ++s) {
        xqo_game g = *game; /* game.clone() */
        int32_t node = root;
        int64_t depth = 0;
        while (t.nodes[node].first_child >= 0) {
            node = select_child(&t, node, c_puct);
            xqo_game_move(&g, t.nodes[node].action);
            ++depth;
        }
        if (depth > max_depth) max_depth = depth;
        int winner, n;
        double value;
        if (xqo_game_over(&g, &winner, mv, &n)) {
            value = winner == 0 ? 0.0 : 1.0; /* mcts.py:140 -- no sign logic */
            ++terminal_sims;
        } else {
            value = eval(g.board, g.player, probs, user);
            if (n > 0) expand(&t, node, probs, mv, n, NULL);
            value = -value;
        }
        for (int32_t u = node; u >= 0; u = t.nodes[u].parent) { /* mcts.py:66-73 */
            t.nodes[u].visits += 1;
            t.nodes[u].total += value;
            value = -value;
        }
    }
    const xqo_node *rn = &t.nodes[root];
    for (int i = 0; i < rn->n_child; ++i) {
        root_actions[i] = t.nodes[rn->first_child + i].action;
        root_visits[i] = t.nodes[rn->first_child + i].visits;
        if (root_total_out) root_total_out[i] = t.nodes[rn->first_child + i].total;
    }
    if (stats_out) {
        stats_out[0] = t.count;
        stats_out[1] = terminal_sims;
        stats_out[2] = max_depth;
    }
    int n_root = rn->n_child;
    free(probs);
    free(t.nodes);
    return n_root;
}

/* ---- built-in deterministic evaluators (mirrored in tests/evaluators.py) ---- */

/* policy = 1/8100 everywhere, value 0 (SURVEY.md appendix B.5 "UniformEval") */
double xqo_eval_uniform(const int8_t *board, int player, float *probs, void *user)
{
    (void)board;
    (void)player;
    (void)user;
    const float u = (float)(1.0 / 8100.0);
    for (int i = 0; i < 8100; ++i) probs[i] = u;
    return 0.0;
}

/* Position hash shared by the test evaluators: FNV-style over (board bytes, side). */
uint32_t xqo_board_hash(const int8_t *board, int player)
{
    uint32_t h = 2166136261u;
    for (int i = 0; i < NSQ; ++i) h = (h ^ (uint8_t)board[i]) * 16777619u;
    h = (h ^ (uint8_t)(player & 0xff)) * 16777619u;
    return h;
}

static inline uint32_t mix32(uint32_t x)
{
    x ^= x >> 16;
    x *= 0x7feb352du;
    x ^= x >> 15;
    x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

/* "Hash evaluator": probs[a] = w(a)/2^20 with integer weights w in 1..256 (dyadic, so the
 * float32 partial sums of the legal mass are exact and order independent); value =
 * clamp(material(side) - material(other), -64, 64) / 64 + small dyadic hash term / 1024. */
double xqo_eval_hash(const int8_t *board, int player, float *probs, void *user)
{
    (void)user;
    uint32_t h = xqo_board_hash(board, player);
    for (int a = 0; a < 8100; ++a) {
        uint32_t w = (mix32(h + 0x9E3779B9u * (uint32_t)(a + 1)) & 255u) + 1u;
        probs[a] = (float)w * (1.0f / 1048576.0f);
    }
    int diff = xqo_material(board, player) - xqo_material(board, -player);
    if (diff > 64) diff = 64;
    if (diff < -64) diff = -64;
    int jitter = (int)(mix32(h ^ 0xA5A5A5A5u) & 15u) - 8; /* -8..7 */
    double v = (double)diff / 64.0 * 0.75 + (double)jitter / 1024.0;
    return (double)(float)v;
}

/* "Ratio evaluator": probs[a] = float32(w(a)) / float32(sum_a w(a)) -- NOT dyadic, so the
 * sequential float32 legal-mass sum (mcts.py:179) is order sensitive; value = one float32
 * division.  Used to pin summation order and the float32 UCB path. */
double xqo_eval_ratio(const int8_t *board, int player, float *probs, void *user)
{
    (void)user;
    uint32_t h = xqo_board_hash(board, player) ^ 0x5bd1e995u;
    uint32_t tot = 0;
    for (int a = 0; a < 8100; ++a) {
        uint32_t w = (mix32(h + 0x9E3779B9u * (uint32_t)(a + 1)) & 255u) + 1u;
        w = w * w; /* sharper policy: 1..65536 */
        probs[a] = (float)w;
        tot += w;
    }
    float ftot = (float)tot;
    for (int a = 0; a < 8100; ++a) probs[a] = probs[a] / ftot;
    int diff = xqo_material(board, player) - xqo_material(board, -player);
    if (diff > 64) diff = 64;
    if (diff < -64) diff = -64;
    int jitter = (int)(mix32(h ^ 0xA5A5A5A5u) & 31u) - 16;
    float v = (float)(diff + jitter) / 97.0f;
    return (double)v;
}

int xqo_sizeof_game(void) { return (int)sizeof(xqo_game); }
