/*
 * xq_b200.h -- C ABI of libxq_b200.so, the B200 (sm_100a) self-play engine.
 *
 * This is the drop-in boundary for the reference's self-play hot path
 * (wenjunyang/xiangqi-alphazero, training/): plain C, raw pointers and sizes, no torch or
 * Python types.  Each entry point names the reference interface it replaces
 * (paths relative to the reference's training/ directory).  INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - return value: 0 = ok, negative = error (xq_last_error() gives the text). No C++
 *     exceptions cross the ABI.
 *   - `d_` pointers are device pointers on the context's GPU, `h_` pointers are host
 *     pointers (pinned memory recommended).  Device-pointer calls are asynchronous on the
 *     `stream` argument (a cudaStream_t, 0 = default stream); host-pointer calls return when
 *     the outputs are complete.
 *   - the caller owns every buffer; the library borrows them for the duration of the
 *     enqueued work and never retains them.
 *   - one context per GPU and process; a context is not thread-safe (the reference's callers
 *     are single-threaded, game.py / mcts.py hold the GIL throughout).
 *
 * Board encoding (game.py:50-65): int8[90], index row*9+col, row 0 = red back rank; red > 0,
 * black < 0; 1 K, 2 A, 3 B, 4 N, 5 R, 6 C, 7 P.  side: +1 red, -1 black.
 * Action id (game.py:112-114): from_square*90 + to_square, 0 <= id < 8100.
 */
#ifndef XQ_B200_H
#define XQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XQ_SQUARES 90
#define XQ_MAX_MOVES 128      /* action slots per position (reference MAX_MOVES=200, game_core.pyx:50;
                                 max observed in legal play is 74; overflow is detected, never silent) */
#define XQ_ACTION_SPACE 8100
#define XQ_PLANES 15
#define XQ_MAX_PLIES 201      /* positions a game can show: move_count 0..200 (game.py:595) */

#define XQ_OK 0
#define XQ_ERR_ARG (-1)
#define XQ_ERR_CUDA (-2)
#define XQ_ERR_OVERFLOW (-3)  /* a position produced more than XQ_MAX_MOVES legal moves */
#define XQ_ERR_STATE (-4)

typedef struct xq_ctx xq_ctx;

/* ---- context ---------------------------------------------------------------------------- */
int xq_create(int device, xq_ctx** out);
void xq_destroy(xq_ctx* ctx);
const char* xq_last_error(const xq_ctx* ctx);   /* ctx may be NULL: last global error */
int xq_version(void);
/* number of kernels this library launched since the last call with reset != 0 */
long long xq_launch_count(xq_ctx* ctx, int reset);
/* CUDA-event timing of the most recent xq_* device call's main kernel, in milliseconds
 * (valid after the stream has been synchronised; <0 if timing is disabled) */
int xq_set_timing(xq_ctx* ctx, int enabled);
float xq_last_kernel_ms(xq_ctx* ctx);

/* ---- K1: batched rules engine -------------------------------------------------------------
 * Replaces cy_generate_legal_moves / cy_is_in_check (cython_engine/game_core.pyx:521-555) and
 * XiangqiGame.get_legal_actions / get_state_for_nn (game.py:523-526, 618-640).
 *
 *   d_boards   [B][90]  int8
 *   d_sides    [B]      int8  (+1 / -1)
 *   d_actions  [B][128] int16  ordered legal action ids (reference generation order), unused = -1
 *   d_n_moves  [B]      uint8  legal-move count (saturates at 128 and raises the overflow flag)
 *   d_in_check [B]      uint8  cy_is_in_check(board, side): own king attacked or missing
 *   d_planes   [B][15][10][9] float32, or NULL: get_state_for_nn planes (0.0 / 1.0)
 */
int xq_movegen_batch(xq_ctx* ctx, const int8_t* d_boards, const int8_t* d_sides, int B,
                     int16_t* d_actions, uint8_t* d_n_moves, uint8_t* d_in_check, float* d_planes,
                     void* stream);

/* Same, host buffers: chunked H2D -> kernel -> D2H pipeline on internal streams.  Returns
 * XQ_ERR_OVERFLOW if any position overflowed (outputs are still written, truncated). */
int xq_movegen_batch_host(xq_ctx* ctx, const int8_t* h_boards, const int8_t* h_sides, int B,
                          int16_t* h_actions, uint8_t* h_n_moves, uint8_t* h_in_check, float* h_planes);

/* Replaces cy_is_attacked (game_core.pyx:508-518) / XiangqiGame._is_attacked (game.py:176-265):
 *   out[i] = is square d_sq[i] (row*9+col) of board i attacked by side d_by[i]. */
int xq_is_attacked_batch(xq_ctx* ctx, const int8_t* d_boards, const uint8_t* d_sq, const int8_t* d_by,
                         int B, uint8_t* d_out, void* stream);
int xq_is_attacked_batch_host(xq_ctx* ctx, const int8_t* h_boards, const uint8_t* h_sq, const int8_t* h_by,
                              int B, uint8_t* h_out);

/* Same as xq_movegen_batch_host with the feature planes as BITS: h_plane_bits [B][44] uint32, bit plane*90+square of
 * a position's 1350 plane values at word (bit >> 5), bit (bit & 31) -- 176 bytes per position across PCIe instead of the
 * 5400 of the float32 planes (which are 0.0 / 1.0 only); xiangqi-alphazero_b200/xq_native.py:unpack_planes expands them. */
int xq_movegen_batch_host_packed(xq_ctx* ctx, const int8_t* h_boards, const int8_t* h_sides, int B,
                                 int16_t* h_actions, uint8_t* h_n_moves, uint8_t* h_in_check, uint32_t* h_plane_bits);
/* the packed planes alone, device buffers: d_bits [B][44] uint32 */
int xq_planes_bits(xq_ctx* ctx, const int8_t* d_boards, const int8_t* d_sides, int B, uint32_t* d_bits, void* stream);

/* Replaces cy_find_king (game_core.pyx:493-505 over _find_king :78-101): the king of side d_sides[i] searched only in
 * that side's own palace (rows 0-2 / 7-9, columns 3-5) in row-major order; d_king_sq[i] = row*9+col, -1 if absent
 * (the reference returns None). */
int xq_find_king_batch(xq_ctx* ctx, const int8_t* d_boards, const int8_t* d_sides, int B, int8_t* d_king_sq, void* stream);
int xq_find_king_batch_host(xq_ctx* ctx, const int8_t* h_boards, const int8_t* h_sides, int B, int8_t* h_king_sq);
/* Replaces cy_has_legal_moves (game_core.pyx:558-569): d_has[i] = 1 iff side d_sides[i] has at least one legal move
 * (the reference generates the full list and tests the count; so does this call, into the context's grow-only scratch). */
int xq_has_legal_moves_batch(xq_ctx* ctx, const int8_t* d_boards, const int8_t* d_sides, int B, uint8_t* d_has, void* stream);
int xq_has_legal_moves_batch_host(xq_ctx* ctx, const int8_t* h_boards, const int8_t* h_sides, int B, uint8_t* h_has);

/* Replaces _is_move_legal (game_core.pyx:209-252) / XiangqiGame._is_move_legal (game.py:441-490): out[i] = 1 iff after the
 * move d_from[i] -> d_to[i] (squares row*9+col; ANY move, not only pseudo-legal ones, as in the reference) the king of
 * d_sides[i] stands in its palace, does not face the other king on an open file and is not attacked. */
int xq_move_is_legal_batch(xq_ctx* ctx, const int8_t* d_boards, const uint8_t* d_from, const uint8_t* d_to, const int8_t* d_sides,
                           int B, uint8_t* d_out, void* stream);
int xq_move_is_legal_batch_host(xq_ctx* ctx, const int8_t* h_boards, const uint8_t* h_from, const uint8_t* h_to, const int8_t* h_sides,
                                int B, uint8_t* h_out);

/* Which K1 kernel xq_movegen_batch[_host] launches: 1 = one thread per board (csrc/xq_rules_tpb.h, the default),
 * 0 = one warp per board (first generation; also the generator inside the MCTS kernels, where a warp owns a game).
 * Same outputs bit for bit; returns the previous value (or < 0 on a bad argument).  The default can also be chosen
 * with the environment variable XQ_MOVEGEN_IMPL=warp|thread. */
int xq_set_movegen_impl(xq_ctx* ctx, int impl);

/* overflow positions seen by movegen calls since the last reset (device counter, synchronises) */
int xq_overflow_count(xq_ctx* ctx, int reset);

/* Device-side uniform-random legal playouts from the start position (the recipe of the
 * reference's own differential test, test_cython.py:87-123, and of BASELINE config 1):
 * game g writes its positions to slots [g*201 .. g*201+n_positions[g]) of d_boards/d_sides;
 * unused slots get side 0.  d_winner[g] in {1,-1,0} (is_game_over, game.py:565-616).
 *   d_boards [G*201][90] int8, d_sides [G*201] int8, d_n_positions [G] int32, d_winner [G] int8 */
int xq_random_playouts(xq_ctx* ctx, uint64_t seed, int n_games, int8_t* d_boards, int8_t* d_sides,
                       int32_t* d_n_positions, int8_t* d_winner, void* stream);

/* ---- K2: batched GPU-resident MCTS -----------------------------------------------------------
 * Replaces training/mcts.py (MCTSNode :21-73, MCTS.search :94-155) for n_games independent games
 * searched in lockstep: one warp per game, one simulation per game per step, so the leaves of a
 * step form ONE evaluator batch (what inference_server.py:145-279 assembled over sockets).
 * Arithmetic follows the reference bit for bit (float32 UCB, float64 noisy root, float64 W,
 * first-max tie-breaks); see csrc/xq_mcts.cu.
 *
 * One search =  set_games -> root_begin -> [evaluate roots] -> root_expand ->
 *               S x ( select -> [evaluate leaves] -> expand_backup ) -> root_visits.
 * The evaluator sits between the calls and may be anything that fills `policy`/`value`.
 *
 * policy_kind: 0 = float32 probabilities, the predict() contract (model.py:109-124);
 *              1 = bf16 logits, 2 = float32 logits (softmax over the legal entries is taken here).
 * row_stride: elements between consecutive games' policy rows (>= 8100).
 */
int xq_mcts_create(xq_ctx* ctx, int max_games, long long node_capacity /* 0 = max_games*801*64 */);

/* Copies game states in: d_boards [G][90], d_sides [G]; optional d_move_count / d_no_capture [G]
 * int32 (NULL = 0), d_ring [G][12][90] = board before move i at slot i%12, the last 12 entries of
 * XiangqiGame.history (NULL = zeros), d_active [G] uint8 (NULL = all active). */
int xq_mcts_set_games(xq_ctx* ctx, int n_games, const int8_t* d_boards, const int8_t* d_sides,
                      const int32_t* d_move_count, const int32_t* d_no_capture, const int8_t* d_ring,
                      const uint8_t* d_active, void* stream);

/* Evaluator inputs of the roots / of the current leaves; every output pointer is optional:
 *   d_planes [G][15][10][9] float32 (get_state_for_nn); d_x_planes = input planes of xq_net_gemm
 *   (bf16 [2][x_rows][8], cell (r,c) of game g at plane row x_row0 + g*90 + r*9 + c); d_boards_out [G][90] + d_sides_out [G]. */
int xq_mcts_root_begin(xq_ctx* ctx, float* d_planes, void* d_x_planes, long long x_rows, long long x_row0,
                       int8_t* d_boards_out, int8_t* d_sides_out, void* stream);
/* mcts.py:110-123: mask + normalise the root policy, optional Dirichlet mixing 0.75 P + 0.25 eta.
 * d_noise [G][128] float64 injects eta (tests); NULL draws Dirichlet(alpha) on the device. */
int xq_mcts_root_expand(xq_ctx* ctx, const void* d_policy, int policy_kind, long long row_stride,
                        const double* d_noise, int add_noise, uint64_t noise_seed, double alpha, void* stream);
/* mcts.py:126-140: descend by PUCT, replay the moves, is_game_over at the leaf; terminal leaves are
 * backed up here, the others wait for expand_backup. */
int xq_mcts_select(xq_ctx* ctx, double c_puct, float* d_planes, void* d_x_planes, long long x_rows, long long x_row0,
                   int8_t* d_boards_out, int8_t* d_sides_out, void* stream);
/* mcts.py:142-153: expand the leaf with masked priors, back up -value. d_value [G] float32. */
int xq_mcts_expand_backup(xq_ctx* ctx, const void* d_policy, int policy_kind, long long row_stride,
                          const float* d_value, void* stream);
/* leaf bookkeeping of the last root_begin/select: state (0 wants evaluation, 1 terminal, 2 idle),
 * legal count and ordered actions; any pointer may be NULL. */
int xq_mcts_leaf_info(xq_ctx* ctx, int32_t* d_state, int32_t* d_n, int16_t* d_actions, void* stream);
/* visit counts of the root children in child (= move generation) order: d_actions/d_visits [G][128],
 * d_n [G], optional d_w [G][128] float64 total values. */
int xq_mcts_root_visits(xq_ctx* ctx, int16_t* d_actions, int32_t* d_visits, int32_t* d_n, double* d_w,
                        void* stream);
/* priors of the root children in child order as the PUCT arithmetic uses them (float64 [G][128]: the float32 P, or the
 * noisy mix 0.75 P + 0.25 eta of mcts.py:117-121 after xq_mcts_root_expand with add_noise); d_n [G] optional. */
int xq_mcts_root_priors(xq_ctx* ctx, double* d_priors, int32_t* d_n, void* stream);
/* h_stats6: simulations, terminal-leaf simulations, max depth, evaluations consumed, nodes in the
 * current search, error bits (1 node pool overflow, 2 >128 legal moves). Synchronises. */
int xq_mcts_stats(xq_ctx* ctx, long long* h_stats6, int reset);

/* ---- K3: ResNet policy-value forward (bf16 tcgen05/TMA implicit GEMM) ---------------------------
 * Replaces XiangqiNet.forward on the inference path (model.py:87-124): every conv / linear layer
 * is one xq_net_gemm launch over "channel-chunk plane" tensors (layout: csrc/xq_net.cu; cell (r,c) of
 * board b at plane row row0 + b*90 + r*9 + c, no halo rows: off-board taps are masked inside the MMAs), BatchNorm
 * folded into weights + bias by the host (eval mode).  The host builds the descriptors once
 * (xiangqi-alphazero_b200/model.py, class B200Net) and replays them with xq_net_run.
 *
 *   mode 0  3x3 conv: out = act(conv(a) + bias [+ residual]); a/out/residual are plane tensors
 *           [chunk][rows][8] bf16 with logical row m at plane row row0+m (row0 >= 10; the buffer must
 *           extend 10 rows past the last 256-row tile pair)
 *   mode 1  1x1 policy/value head conv (nt = 48: 32 policy + 4 value channels + padding):
 *           out = FC input planes [360][out_rows][8] bf16, out2 = value features [B][90][4] fp32
 *   mode 2  dense layer: a = planes [K/8][a_rows][8], out = row-major bf16 [B][out_stride] logits
 * w = weight image [n_tile][tap][k_block][chunk][nt][8] bf16 (3x3: tap 0 = centre, then the other 8 taps
 * row-major), bias fp32 [n_tiles*nt].
 */
typedef struct xq_gemm_desc {
    int32_t mode;
    int32_t m_tiles;      /* 128-row output tiles */
    int32_t n_tiles;      /* tiles of nt output channels */
    int32_t nt;           /* 128, or 48 for the head conv */
    int32_t kchunks;      /* K / 8 */
    int32_t kch_iter;     /* chunks per pipeline stage: 8, or 2 for the 15(+1)-plane input conv */
    int32_t relu;
    int32_t n_boards;
    int64_t a_rows, a_row0;
    int64_t out_rows, out_row0;
    int64_t out_stride;
    const void* a;
    const void* w;
    const float* bias;
    const void* residual; /* mode 0 only, may be NULL */
    void* out;
    void* out2;           /* mode 1 only */
    const void* w_half;   /* mode 0 with 128 output channels per tile, optional: the same weights tiled by 64 output
                             channels [n_tile64][tap][k_block][chunk][64][8] for the CTA-pair kernel (cta_group::2 MMAs, each
                             CTA holds half of N); NULL: the single-CTA kernel runs the layer */
} xq_gemm_desc;

int xq_net_gemm(xq_ctx* ctx, const xq_gemm_desc* desc, void* stream);
/* value head tail, model.py:79-83: tanh(w2 . relu(W1 f + b1) + b2); d_feats [B][90][4] fp32,
 * d_w1t [360][128] fp32 with k = pos*4 + ch. */
int xq_net_value_head(xq_ctx* ctx, const float* d_feats, const float* d_w1t, const float* d_b1,
                      const float* d_w2, float b2, float* d_value, int B, void* stream);
/* all layers of one forward + the value head for the first B boards (B <= the descriptors' n_boards: every
 * launch is sized to B), one call */
int xq_net_run(xq_ctx* ctx, const xq_gemm_desc* layers, int n_layers, const float* d_vfeats,
               const float* d_w1t, const float* d_b1, const float* d_w2, float b2, float* d_value, int B,
               void* stream);

/* xq_net_run with the board count on the DEVICE (*d_n_boards, clamped to max_boards): grids are sized for max_boards and
 * every kernel cuts its tile loop to the live count -- the leaf-compacting self-play loop never synchronises. */
int xq_net_run_counted(xq_ctx* ctx, const xq_gemm_desc* layers, int n_layers, const float* d_vfeats,
                       const float* d_w1t, const float* d_b1, const float* d_w2, float b2, float* d_value,
                       const int* d_n_boards, int max_boards, void* stream);

/* ---- device-resident self-play ------------------------------------------------------------------
 * Replaces parallel_selfplay.py:_play_one_game (:42-134) + the worker fan-out (:337-474): n_slots
 * games advance in lockstep on one GPU, each ply = root evaluation -> num_simulations x (select,
 * network forward, expand+backup) -> temperature sampling; finished games are replaced in their
 * slot until target_games have been started.  The leaves that need the network are COMPACTED into the
 * first rows of the batch (what inference_server.py:163-249 did by batching only live requests): idle
 * slots, finished games and terminal leaves cost no forward rows.  Config fields are the reference's 8 worker keys
 * (parallel_selfplay.py:184-187).
 */
typedef struct xq_selfplay_config {
    int32_t num_simulations;
    float c_puct;
    int32_t temperature_threshold;
    int32_t max_game_length;
    int32_t random_opening_moves;
    int32_t enable_resign;
    float resign_threshold;
    int32_t resign_check_steps;
    int32_t add_noise;          /* root Dirichlet noise (mcts.py:117-121); the reference always uses 1 */
    float dirichlet_alpha;      /* 0.3 */
    uint64_t seed;
    int32_t target_games;       /* games to start in total (num_games_per_iter share of this GPU) */
    int32_t leaves_per_game;    /* K: descents per game and lockstep step.  0 or 1 = the reference's one simulation at a
                                   time (mcts.py:126-153, visit counts bit-exact).  K > 1 = opt-in multi-leaf mode: K
                                   descents per step with a VIRTUAL LOSS on the path of each (N += 1, W -= 1 until its
                                   backup), all K leaves evaluated in the same forward; a search still makes exactly
                                   num_simulations descents.  The network plan's batch must be >= n_slots * K. */
} xq_selfplay_config;

/* what the loop needs to run the evaluator: the layer list of xq_net_run plus its I/O buffers */
typedef struct xq_net_plan {
    const xq_gemm_desc* layers;
    int32_t n_layers;
    int32_t batch;              /* boards the descriptors were built for (>= n_slots) */
    const float* vfeats;
    const float* w1t;
    const float* b1;
    const float* w2;
    float b2;
    float* value;               /* [batch] */
    void* x_planes;             /* network input planes written by the search kernels */
    int64_t x_rows, x_row0;
    const void* logits;         /* [batch][logit_stride] */
    int64_t logit_stride;
    int32_t logits_kind;        /* 1 bf16, 2 float32 */
} xq_net_plan;

/* Optional hint for the self-play / arena loops: an upper bound on the games that are still alive (e.g. target_games -
 * finished once every game has been started).  The forwards of the following plies are launched for that many boards instead
 * of one per slot, which selects the small-batch variants of the layers when few games are left (the latency of a forward is
 * what a lockstep step costs then).  0 = no bound.  A leaf beyond the bound is not evaluated and sets error bit 4 in the
 * counters; xq_selfplay_reset clears the hint. */
int xq_selfplay_set_live_bound(xq_ctx* ctx, int32_t max_live_games);

/* sample record (896 bytes): board int8[90] @0, side int8 @90, n_moves uint8 @91, game uid int32 @92,
 * ply int32 @96, action played int16 @100, actions int16[128] @128, visit probabilities float32[128] @384 */
#define XQ_SAMPLE_BYTES 896

int xq_selfplay_create(xq_ctx* ctx, int n_slots, int max_games_total, long long sample_capacity,
                       long long node_capacity);
int xq_selfplay_reset(xq_ctx* ctx, void* stream);
int xq_selfplay_play(xq_ctx* ctx, const xq_selfplay_config* cfg, const xq_net_plan* net, int n_plies, void* stream);
/* h_out14: games started, finished, samples, red wins, black wins, draws, plies of finished games, dropped
 * samples, then the 6 values of xq_mcts_stats */
int xq_selfplay_counters(xq_ctx* ctx, long long* h_out14);
int xq_selfplay_fetch(xq_ctx* ctx, long long first, long long count, void* h_samples, int8_t* h_winner,
                      int16_t* h_plies, int n_results);
int xq_selfplay_slots(xq_ctx* ctx, int8_t* h_boards, int32_t* h_meta, int32_t* h_status, int32_t* h_uid);

/* Device pointers of the sample records [sample_capacity][896] and of the per-game results (winner in {1,-1,0}, 2 = not
 * finished; plies), for consumers that stay on the GPU (xq_replay_append). */
int xq_selfplay_device_buffers(xq_ctx* ctx, void** d_samples, int8_t** d_winner, int16_t** d_plies);

/* ---- evaluation arena -----------------------------------------------------------------------------
 * Replaces AlphaZeroTrainer._serial_evaluate (train.py:453-535): cfg->target_games evaluation games played in the
 * slots of xq_selfplay_create, game uid with the NEW model as red when uid is even (:474), every move =
 * get_action(game, temperature=0, add_noise=False) of the model to move (:481-483) with cfg->num_simulations
 * simulations, from the initial position, no resignation; a game still undecided after cfg->max_game_length plies is a
 * draw (:496-498).  A game's leaves go to the network of the player to move at the root only (two compacted batches);
 * results land in the per-game result arrays
 * (xq_selfplay_fetch / xq_selfplay_counters).  d_move_log [max_games_total][XQ_MAX_PLIES] int16 receives the actions played
 * (optional).  Call xq_selfplay_reset before a new match. */
int xq_arena_play(xq_ctx* ctx, const xq_selfplay_config* cfg, const xq_net_plan* net_new, const xq_net_plan* net_old,
                  int n_plies, int16_t* d_move_log, void* stream);

/* ---- training step around the torch forward/backward (train.py:376-447) ---------------------------------
 * Device-resident replay ring of the 896-byte sparse self-play records, replacing the deque of dense tuples
 * (train.py:203) and SelfPlayDataset (:114-129).
 *
 * xq_replay_append: records d_src_records[d_src_index[j]], j < n, go to ring slots (head + j) % capacity with their value
 * label z (parallel_selfplay.py:124-132: 0 draw, +1 if the side to move at the sample won, -1 otherwise) computed from
 * d_winner[game uid]. */
int xq_replay_append(xq_ctx* ctx, const void* d_src_records, const int64_t* d_src_index, int n, const int8_t* d_winner,
                     int n_results, void* d_ring, float* d_ring_z, long long capacity, long long head, void* stream);
/* xq_train_batch: minibatch from LOGICAL indices L: ring record (start + (L >> 1)) % capacity, its column-mirrored twin
 * when L is odd (_augment_data, parallel_selfplay.py:137-151, as an index permutation).  Outputs: planes float32
 * [B][15][10][9] (get_state_for_nn, game.py:618-640), sparse policy target (action ids int16 [B][128], visit
 * probabilities float32 [B][128], counts int32 [B]) and z float32 [B]. */
int xq_train_batch(xq_ctx* ctx, const void* d_ring, const float* d_ring_z, long long capacity, long long start,
                   const int64_t* d_logical_index, int B, float* d_planes, int16_t* d_actions, float* d_probs,
                   int32_t* d_n_moves, float* d_z, void* stream);
/* xq_policy_value_loss: train.py:408-414 and its gradient in one pass over the logits.  Per row i:
 *   policy_loss_rows[i] = -sum_a pi_a log_softmax(logits_i)_a,  value_loss_rows[i] = (value_i - z_i)^2,
 *   grad_logits_i = (softmax(logits_i) sum(pi) - pi) * inv_batch,  grad_value_i = 2 (value_i - z_i) * inv_batch
 * (inv_batch = 1 / GLOBAL batch size, so that the data-parallel sum of gradients is the full-batch gradient). */
int xq_policy_value_loss(xq_ctx* ctx, const float* d_logits, long long logit_stride, const float* d_value,
                         const int16_t* d_actions, const float* d_probs, const int32_t* d_n_moves, const float* d_z, int B,
                         float inv_batch, float* d_grad_logits, long long grad_stride, float* d_grad_value,
                         float* d_policy_loss_rows, float* d_value_loss_rows, void* stream);
/* sum of squares of a flat float32 buffer in a fixed order (d_partial: scratch of n_partial >= 1 floats) */
int xq_grad_sumsq(xq_ctx* ctx, const float* d_grad, long long n, float* d_partial, int n_partial, float* d_out, void* stream);
/* clip_grad_norm_(max_norm) + torch.optim.Adam.step (L2 weight decay, no amsgrad) over flat buffers (train.py:190-194,
 * 418-419).  d_grad_sumsq: device scalar from xq_grad_sumsq (NULL or max_norm <= 0: no clipping); grad_scale multiplies
 * the gradient first (1 / world size when the all-reduce summed per-rank means).  step = 1, 2, ... */
int xq_adam_step(xq_ctx* ctx, float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq, long long n,
                 float lr, float beta1, float beta2, float eps, float weight_decay, long long step,
                 const float* d_grad_sumsq, float max_norm, float grad_scale, void* stream);

/* ---- data-parallel BatchNorm of the training step (csrc/xq_bn.cu) -------------------------------------------
 * Training-mode BatchNorm2d (the 15 layers of XiangqiNet inside train.py:397-423) whose statistics are those of the
 * WHOLE minibatch when it is split across the ranks of one NVSwitch box: per layer and direction one reduce kernel
 * stores its per-channel partial sums straight into every peer's exchange buffer (CUDA IPC peer mapping, NVLink) and one
 * apply kernel waits for all ranks' flags, adds the partials in rank order and normalises -- no collective call.
 *   xq_peer_create: allocate this rank's exchange buffer, return its 64-byte IPC handle;
 *   xq_peer_connect: open the other ranks' buffers (handles = world x 64 bytes in rank order, e.g. from an all_gather).
 * Without these two calls a context is its own single-rank group.  Every rank must issue the same sequence of
 * xq_bn_forward / xq_bn_backward calls.  x, y, dy, dx: float32 [N][C][HW] contiguous (NCHW); C <= 512. */
int xq_peer_create(xq_ctx* ctx, int rank, int world, unsigned char* handle_out64);
int xq_peer_connect(xq_ctx* ctx, const unsigned char* handles);
/* y = (x - mean) * invstd * weight + bias with the global batch statistics; running_mean / running_var get F.batch_norm's
 * momentum update (unbiased variance); save_mean / save_invstd [C] are what xq_bn_backward needs. */
int xq_bn_forward(xq_ctx* ctx, const float* d_x, float* d_y, const float* d_weight, const float* d_bias,
                  float* d_running_mean, float* d_running_var, float* d_save_mean, float* d_save_invstd, int N, int C,
                  int HW, float eps, float momentum, void* stream);
/* dx over the global batch; dweight / dbias [C] are this rank's LOCAL sums (the gradient all-reduce adds them up). */
int xq_bn_backward(xq_ctx* ctx, const float* d_x, const float* d_dy, const float* d_weight, const float* d_save_mean,
                   const float* d_save_invstd, float* d_dx, float* d_dweight, float* d_dbias, int N, int C, int HW,
                   void* stream);


/* ---- f1: the training step's contractions on tcgen05 (csrc/xq_tnet.cu) ----------------------------------------
 * Replaces the library calls (cuDNN convolutions, cuBLAS linears) inside AlphaZeroTrainer.train_network's
 * forward/backward (train.py:397-423 over model.py:39-107) by hand-written kind::tf32 tcgen05 kernels, fp32 storage and
 * fp32 accumulation.  TRAINING PLANE LAYOUT: a tensor is float32 X[C/4][rows][4 channels] (16 bytes per (chunk,row));
 * board b, cell (r,c) lives at row row0 + b*110 + r*10 + c: every board row carries one zero pad column (c = 9) and every
 * board one zero pad row (r = 10), so a 3x3 tap (dy,dx) is the same matrix shifted by dy*10+dx rows and the pad cells
 * supply the zero padding of conv2d(padding=1) -- in the forward (fprop), the data gradient (dgrad) and the weight
 * gradient (wgrad) alike, without masks.  Dense layers use the same layout with one row per board.
 *
 * xq_tgemm: out[row][n] = sum_tap sum_k A[row + s*shift(tap)][k] * Wtap[n][k], s = shift_sign.
 * A: planes [kblocks*8 chunks][a_rows][4]; W: a weight image [img_nt][ntaps][img_kb][8 chunks][128 n][4 k] float32
 * (tap = kh*3+kw for ntaps = 9), one 16 KB stage per (n_tile, tap, k_block).  fprop: the image of the layer's weights,
 * s = +1.  dgrad: the image of the TRANSPOSED weights (n = input channel, k = output channel), s = -1.
 * Output: planes (out, optional residual planes added) or row-major with bias (out_rm). */
#define XQ_T_BOARD_ROWS 110
#define XQ_T_ROW0 16
typedef struct xq_tgemm_desc {
    const void* a;
    int64_t a_rows, a_row0;
    const void* w;
    int32_t kblocks;        /* contraction blocks of 32 (A chunks / 8) */
    int32_t ntaps;          /* 9 (3x3) or 1 */
    int32_t img_kb;         /* k-blocks of the weight image */
    int32_t b_mn;           /* reserved, 0 */
    int32_t shift_sign;     /* +1 / -1 */
    int32_t m_pairs;        /* 256-row work items along M */
    int32_t n_tiles;        /* 128-column output tiles */
    int32_t out_chunks;     /* planes output: chunks (of 4 channels) to store, the rest of the last tile is dropped */
    int64_t m_rows;         /* rows to store (rows >= m_rows of the last item are dropped) */
    void* out;              /* planes [out_chunks][out_rows][4], logical row 0 at out_row0; or NULL */
    int64_t out_rows, out_row0;
    const void* residual;   /* planes like out, added in the epilogue; or NULL */
    float* out_rm;          /* row-major [m_rows][out_stride] (+ bias[n]); or NULL */
    int64_t out_stride;
    const float* bias;
    int32_t n_cols;         /* row-major output: columns to store (multiple of 4) */
    int32_t k_splits;       /* 0/1: none.  > 1: the contraction blocks are dealt to k_splits work items per output tile (fills the
                               SMs when M x N alone gives few tiles).  Planes output: split s writes its partial sums to
                               out + s*out_split_stride floats, the reader adds them (xq_tn_unflatten); row-major output: at most 2,
                               added into the zeroed matrix (a two-term sum does not depend on the order) */
    int64_t out_split_stride;
} xq_tgemm_desc;
int xq_tgemm(xq_ctx* ctx, const xq_tgemm_desc* desc, void* stream);

/* xq_twgrad: D_t[m][n] = sum_row A[row][m] * B[row + off_t][n] over a range of rows (conv wgrad: A = dY, B = the layer
 * input, taps = row shifts; dense wgrad: taps = column groups).  Both operands are MN-major for the tensor core, which for
 * tf32 exists in one shared-memory form only (SWIZZLE_128B_BASE32B), so they are read from the G LAYOUT: float32
 * G[C/32][rows][32 channels], the four 32-byte units of row r stored at position u ^ (r & 3); rows as in the plane layout.
 * Work item = (m_tile of 128 A channels, slab of kr*stages_per_item rows, tap group g of taps_per_group taps).  B stage of
 * group g: channel groups [b_group0 + g*b_group_step, + b_groups_stage) x rows [k0 + b_row_lo[g], + b_rows_stage)
 * (b_row_lo and b_rows_stage multiples of 4); tap t of group g reads it at byte offset tap_off[g*4 + t] (with more than 4 groups every group uses entry 0 of both).  Output element
 * (m, n) of tap t of item (mt, slab, g) goes to out[mt*mt_stride + slab*slab_stride + g*g_stride + t*tap_stride + m*ldo + n],
 * stored when mt*128+m < m_limit and g*g_cols + t*t_cols + n < n_limit. */
typedef struct xq_twgrad_desc {
    const void* a;
    const void* b;
    int64_t a_rows, a_row0, b_rows, b_row0;
    int32_t a_group0, b_group0;
    int32_t nbg;              /* B channel groups per tap: N = 32*nbg (32 .. 128) */
    int32_t kr;               /* rows per pipeline stage (multiple of 8) */
    int32_t stages_per_item;
    int32_t n_slabs, n_groups, n_mtiles, taps_per_group;
    int32_t b_rows_stage, b_groups_stage, b_group_step;
    int32_t b_row_lo[4];
    int32_t tap_off[16];
    float* out;
    int64_t mt_stride, slab_stride, g_stride, tap_stride, ldo;
    int32_t m_limit, n_limit, g_cols, t_cols;
} xq_twgrad_desc;
int xq_twgrad(xq_ctx* ctx, const xq_twgrad_desc* desc, void* stream);

/* ---- the layers between the contractions (csrc/xq_tnet_ops.cuh) --------------------------------------------------------
 * Everything else train.py:397-423 runs through torch modules (model.py:14-107): layout writers, training-mode BatchNorm
 * (+ residual + ReLU) forward / backward, weight images, the slab reduction of the weight gradients, nn.Flatten around the
 * policy FC and the value head's two small dense layers.  All reductions run in a fixed order (bit-reproducible).
 * Tensors: planes float32 P[chunk][rows][4], G layout float32 G[group][rows][32] (see above); `rows` is the row count of
 * the tensor, board b cell (r, c) at row XQ_T_ROW0 + b*110 + r*10 + c; dense tensors have one row per board at
 * XQ_T_ROW0 + b. */

/* x[n_boards][channels][10][9] -> planes chunks [0, 2*pairs) (channels >= `channels` zero) and, if g != NULL, the G layout. */
int xq_tn_input(xq_ctx* ctx, const float* x, int32_t n_boards, int32_t channels, int32_t pairs, float* planes, float* g,
                int64_t rows, void* stream);
/* w[co][ci][taps] -> weight image [.][taps][img_kb][8][128][4] of xq_tgemm at (n0 + n, k0 + k): transposed = 0: n = co,
 * k = ci (fprop); 1: n = ci, k = co (dgrad).  Positions without a weight are left untouched (allocate the image zeroed). */
int xq_tn_wimage(xq_ctx* ctx, const float* w, int32_t co, int32_t ci, int32_t taps, float* img, int32_t img_kb, int32_t n0,
                 int32_t k0, int32_t transposed, void* stream);
/* Both images of a dense layer's weight w[co][ci] (ci a multiple of 32) in one pass over it: img (n = co, k = ci, fprop) and
 * img_t (n = ci, k = co, dgrad). */
int xq_tn_wimage_dense2(xq_ctx* ctx, const float* w, int32_t co, int32_t ci, float* img, int32_t img_kb, float* img_t,
                        int32_t img_t_kb, void* stream);
/* The same for many small weight tensors in one launch per 32 items (the 3x3 and 1x1 convolutions of a step). */
typedef struct xq_tn_wimage_item {
    const float* w;
    float* img;
    int32_t co, ci, taps, img_kb, n0, k0, transposed, pad_;
} xq_tn_wimage_item;
int xq_tn_wimage_batch(xq_ctx* ctx, const xq_tn_wimage_item* items, int32_t n_items, void* stream);

/* BatchNorm2d in training mode over chunks [chunk0, chunk0 + ceil(n_channels/8)*2) of y (nn.BatchNorm2d as used by
 * model.py:14-36, 49-83): batch statistics over the n_boards*90 real cells, running statistics updated with `momentum`
 * (unbiased variance), out = [relu](gamma*(y - mean)*invstd + beta [+ res]), pad cells zero.  partial: scratch,
 * 256 doubles per chunk.  save[2][n_channels] receives mean and invstd for the backward pass. */
typedef struct xq_tn_bn_desc {
    const float* y;
    const float* res;
    float* out;
    float* out_g;
    int64_t rows;
    int32_t n_boards, chunk0, n_channels, relu;
    double* partial;
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    float* save;
    float eps, momentum;
} xq_tn_bn_desc;
int xq_tn_bn_forward(xq_ctx* ctx, const xq_tn_bn_desc* desc, void* stream);

/* Backward of the above: dz = dout * (act > 0) (relu) ; dgamma = sum dz*xhat, dbeta = sum dz (assigned, not accumulated);
 * dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)) as planes and (dy_g != NULL) G layout, pad cells zero;
 * dskip (optional) receives dz, the gradient of the residual input. */
typedef struct xq_tn_bn_bwd_desc {
    const float* dout;
    const float* act;
    const float* y;
    int64_t rows;
    int32_t n_boards, chunk0, n_channels, relu;
    const float* save;
    double* partial;
    const float* gamma;
    float* dgamma;
    float* dbeta;
    float* dy;
    float* dy_g;
    float* dskip;
} xq_tn_bn_bwd_desc;
int xq_tn_bn_backward(xq_ctx* ctx, const xq_tn_bn_bwd_desc* desc, void* stream);

/* Sum of the xq_twgrad slabs into a weight gradient in the parameter layout: ws[slab][tap][128][ldn] ->
 * dw[(co0 + co)*ci_total + ci0 + ci][tap], (co, ci) = (m, n - n_src0), or (ci, co) = (m, n - n_src0) when transposed. */
int xq_tn_wgrad_reduce(xq_ctx* ctx, const float* ws, int32_t slabs, int64_t slab_stride, int32_t taps, int32_t ldn, int32_t m_cnt,
                       int32_t n_cnt, int32_t n_src0, int32_t transposed, float* dw, int32_t ci_total, int32_t co0, int32_t ci0,
                       void* stream);

/* nn.Flatten of the first `channels` channels: feature k = ch*90 + r*9 + c -> dense planes [k/4][drows][4] (+ G layout). */
int xq_tn_flatten(xq_ctx* ctx, const float* act, int64_t rows, int32_t n_boards, int32_t channels, float* dense, float* dense_g,
                  int64_t drows, void* stream);
/* ... and back: dense planes (the sum of n_partials tensors part_stride floats apart, added in order) -> chunks
 * [0, channels/4) of board planes (real cells). */
int xq_tn_unflatten(xq_ctx* ctx, const float* dense, int64_t drows, int32_t n_boards, int32_t channels, float* planes, int64_t rows,
                    int32_t n_partials, int64_t part_stride, void* stream);
/* Row-major m[n_rows][stride] (n_cols used, multiple of 4) -> dense planes and / or G layout, one row per board. */
int xq_tn_rows_layouts(xq_ctx* ctx, const float* m, int64_t stride, int32_t n_rows, int32_t n_cols, float* dense, float* dense_g,
                       int64_t drows, void* stream);
/* out[n] = sum over rows of m[row][n] (bias gradient of a dense layer). */
int xq_tn_colsum(xq_ctx* ctx, const float* m, int64_t stride, int32_t n_rows, int32_t n_cols, float* out, void* stream);

/* Value head after its BatchNorm (model.py:72-83): the 4 channels of plane chunk `chunk` of `act`, flattened (k = ch*90 + cell),
 * h = relu(W1 f + b1) [n_boards][128], v = tanh(w2 h + b2) [n_boards]. */
int xq_tn_value_forward(xq_ctx* ctx, const float* act, int64_t rows, int32_t chunk, int32_t n_boards, const float* w1, const float* b1,
                        const float* w2, const float* b2, float* h, float* v, void* stream);
/* Backward: g_value = dLoss/dv.  Writes the four parameter gradients and the gradient of the 4 channels into chunk
 * `chunk` of dact (real cells); dh [n_boards][128] and dpre [n_boards] are scratch. */
int xq_tn_value_backward(xq_ctx* ctx, const float* act, int64_t rows, int32_t chunk, int32_t n_boards, const float* w1, const float* w2,
                         const float* h, const float* v, const float* g_value, float* dh, float* dpre, float* dact, float* dw1,
                         float* db1, float* dw2, float* db2, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XQ_B200_H */
