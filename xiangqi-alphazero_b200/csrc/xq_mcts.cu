// xq_mcts.cu -- K2: GPU-resident batched MCTS (one flat tree per game, one warp per game).
//
// Replaces training/mcts.py: MCTSNode (:21-73: select_child :43-58, expand :60-64, backup
// :66-73) and the body of MCTS.search (:94-155) for B independent games advanced in lockstep,
// one simulation per game per step, so that the B leaves of a step form one evaluator batch
// (the in-device replacement of inference_server.py's socket batching).
//
// Bit-level contract (pinned by tests/golden/mcts_golden.json):
//   - children are created in legal-move generation order; argmax uses strict '>' so the
//     first maximum wins;
//   - priors are np.float32 (p[a] / sequential-float32-sum), UCB is evaluated in float32 with
//     the reference's operation order;  a root with Dirichlet noise has float64 priors
//     (0.75*P32 + 0.25*noise) and float64 UCB;  the all-zero-mass fallback is the python float 1/n;
//   - W accumulates in float64, q = W/N in float64 then (float32 path) rounded to float32;
//   - a decisive terminal leaf backs up +1.0 whoever won (mcts.py:140), a drawn one 0.0, and
//     stays a leaf; a non-terminal leaf is expanded and backs up -value.
//
// Memory: two 16-byte records per node -- Hot {W f64, N i32, P f32} read by the select
// argmax with one coalesced 16 B load per lane, and Link {first child, parent, action, child
// count, flags} read once per descent level.  Nodes of one expansion are contiguous.
#include "xq_ctx.h"
#include "xq_rules.cuh"

#include <cuda_bf16.h>

namespace xq {

struct __align__(16) NodeHot {
    double W;
    int32_t N;
    float P;
};
struct __align__(16) NodeLink {
    int32_t child0;   // -1 while leaf
    int32_t parent;   // -1 for a root
    int16_t action;
    uint8_t nchild;
    uint8_t flags;    // priors of this node's CHILDREN: 0 float32, 1 float64 uniform 1/n, 2 float64 in rootP64
    int32_t pad;
};

enum { kLeafEval = 0, kLeafTerminal = 1, kLeafIdle = 2, kLeafDup = 3 };   // kLeafDup: the leaf another descent of the same step already selected
constexpr uint8_t kPendingFlag = 0x80;   // NodeLink.flags of a leaf that waits for its evaluation (pad = evaluator row)

struct MctsState {
    int max_games = 0;
    long long cap_nodes = 0;
    int n_games = 0;
    // game state
    int8_t* board = nullptr;      // [G][96]
    int8_t* ring = nullptr;       // [G][12][96]
    int32_t* meta = nullptr;      // [G][4] side, move_count, no_capture, active
    // tree
    NodeHot* hot = nullptr;
    NodeLink* link = nullptr;
    double* rootP64 = nullptr;    // [G][128]
    int* alloc = nullptr;         // bump pointer into the node pool
    int* error = nullptr;         // bit0 pool overflow, bit1 move overflow
    // per-step leaf records
    int32_t* leaf_node = nullptr;   // [G]
    int32_t* leaf_state = nullptr;  // [G]
    int16_t* leaf_actions = nullptr;  // [G][128]
    int32_t* leaf_n = nullptr;        // [G]
    int64_t* stats = nullptr;         // [4] sims, terminal sims, max depth, evals
    int32_t* root_winner = nullptr;   // [G] is_game_over at the root: 1/-1/0, or 2 = game goes on
    // self-play / arena loop (leaf slots are [G][leaf_k], slot g*K+j; the standalone xq_mcts_* calls use slot g, K = 1)
    int leaf_k = 1;                   // leaf slots allocated per game
    int32_t* leaf_row = nullptr;      // [G*leaf_k] evaluator row of the slot (compacted), -1 none
    int32_t* sims_left = nullptr;     // [G] simulations still to run in the current search
    int* n_eval = nullptr;            // [2] rows handed out in the current step, per network
};

constexpr int kSelWarps = 4;

struct __align__(16) SelectSmem {
    int8_t board[kSelWarps][kBoardPad];
    int8_t ring[kSelWarps][kRing * kBoardPad];
    WarpScratch ws[kSelWarps];
};

__device__ __forceinline__ void warp_load_game(const MctsState& M, int g, int8_t* b, int8_t* ring, GameMeta& gm)
{
    const int lane = lane_id();
    // 96 B board = 6 x uint4, 1152 B ring = 72 x uint4
    const uint4* gb = reinterpret_cast<const uint4*>(M.board + (size_t)g * kBoardPad);
    const uint4* gr = reinterpret_cast<const uint4*>(M.ring + (size_t)g * kRing * kBoardPad);
    if (lane < 6) reinterpret_cast<uint4*>(b)[lane] = gb[lane];
    for (int i = lane; i < 72; i += 32) reinterpret_cast<uint4*>(ring)[i] = gr[i];
    gm.side = M.meta[g * 4 + 0];
    gm.move_count = M.meta[g * 4 + 1];
    gm.no_capture = M.meta[g * 4 + 2];
    warp_sync();
}

// get_state_for_nn planes (game.py:618-640) and/or the conv input tile of the network kernels
// (bf16, 16 channels per cell, rows padded as [board][11][10] with zero halo cells never written).
__device__ __forceinline__ void warp_emit_eval_inputs(const int8_t* b, int side, int g, float* planes_f32,
                                                      __nv_bfloat16* x_planes, long long x_rows, long long x_row0,
                                                      int8_t* boards_out, int8_t* sides_out)
{
    const int lane = lane_id();
    if (boards_out) {
        for (int i = lane; i < kSquares; i += 32) boards_out[(size_t)g * kSquares + i] = b[i];
        if (lane == 0) sides_out[g] = (int8_t)side;
    }
    if (planes_f32) {
        float2* out = reinterpret_cast<float2*>(planes_f32 + (size_t)g * 15 * kSquares);
        const float turn = side == 1 ? 1.0f : 0.0f;
        for (int e2 = lane; e2 < 15 * kSquares / 2; e2 += 32) {
            const int e = 2 * e2, p = e / kSquares, sq = e - p * kSquares;
            float2 v;
            if (p == 14) v.x = v.y = turn;
            else {
                int v0 = b[sq] * side, v1 = b[sq + 1] * side;
                int c0 = v0 > 0 ? v0 - 1 : (v0 < 0 ? 6 - v0 : -1);
                int c1 = v1 > 0 ? v1 - 1 : (v1 < 0 ? 6 - v1 : -1);
                v.x = c0 == p ? 1.0f : 0.0f;
                v.y = c1 == p ? 1.0f : 0.0f;
            }
            out[e2] = v;
        }
    }
    if (x_planes) {
        // network input planes [2 chunks][x_rows][8 ch] bf16: channels 0-7 in chunk 0, 8-14 (+pad) in chunk 1
        const uint32_t one = 0x3f80u;   // bf16 1.0
        for (int sq = lane; sq < kSquares; sq += 32) {
            int v = b[sq] * side;
            int ch = v > 0 ? v - 1 : (v < 0 ? 6 - v : -1);
            uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (ch >= 0) w[ch >> 1] = one << ((ch & 1) * 16);
            if (side == 1) w[7] |= one;   // channel 14 = low half of word 7; channel 15 is padding
            const size_t row = (size_t)(x_row0 + (long long)g * 90 + sq);
            uint4* p0 = reinterpret_cast<uint4*>(x_planes + row * 8);
            uint4* p1 = reinterpret_cast<uint4*>(x_planes + ((size_t)x_rows + row) * 8);
            *p0 = make_uint4(w[0], w[1], w[2], w[3]);
            *p1 = make_uint4(w[4], w[5], w[6], w[7]);
        }
    }
}

// lane 0 walks the parent chain (mcts.py:66-73)
__device__ __forceinline__ int backup_path(const MctsState& M, int node, double value)
{
    int depth = 0;
    while (node >= 0) {
        NodeHot h = M.hot[node];
        h.N += 1;
        h.W = __dadd_rn(h.W, value);
        M.hot[node] = h;
        value = -value;
        node = M.link[node].parent;
        ++depth;
    }
    return depth;
}

// Backup of a descent that carried a virtual loss: the selection already counted the visit (N += 1) and charged a loss
// (W -= 1) on every node of the path, so the real result replaces the loss and N stays.
__device__ __forceinline__ int backup_path_vl(const MctsState& M, int node, double value)
{
    int depth = 0;
    while (node >= 0) {
        NodeHot h = M.hot[node];
        h.W = __dadd_rn(__dadd_rn(h.W, 1.0), value);
        M.hot[node] = h;
        value = -value;
        node = M.link[node].parent;
        ++depth;
    }
    return depth;
}

// ---- root preparation: movegen on every root, emit evaluator inputs -------------------------
__global__ void __launch_bounds__(kSelWarps * 32)
mcts_root_begin_kernel(MctsState M, float* planes_f32, __nv_bfloat16* x_planes, long long x_rows, long long x_row0,
                       int8_t* boards_out, int8_t* sides_out)
{
    __shared__ SelectSmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kSelWarps + warp;
    if (g >= M.n_games) return;
    int8_t* b = sm.board[warp];
    GameMeta gm;
    warp_load_game(M, g, b, sm.ring[warp], gm);
    WarpScratch& S = sm.ws[warp];
    const bool active = M.meta[g * 4 + 3] != 0;
    MovegenResult r = warp_movegen(b, gm.side, S);
    if (r.overflow && lane == 0) atomicOr(M.error, 2);
    const int rw = warp_game_over(b, sm.ring[warp], gm, r);
    if (lane == 0) M.root_winner[g] = rw;
    const int n = active ? min(r.n_legal, kMaxMoves) : 0;
    reinterpret_cast<uint2*>(M.leaf_actions + (size_t)g * kMaxMoves)[lane] = reinterpret_cast<const uint2*>(S.actions)[lane];
    if (lane == 0) {
        M.leaf_n[g] = n;
        M.leaf_node[g] = g;
        M.leaf_state[g] = n > 0 ? kLeafEval : kLeafIdle;
        // fresh root (mcts.py:104)
        M.hot[g] = NodeHot{0.0, 0, 0.0f};
        M.link[g] = NodeLink{-1, -1, (int16_t)-1, 0, 0, 0};
    }
    warp_emit_eval_inputs(b, gm.side, g, planes_f32, x_planes, x_rows, x_row0, boards_out, sides_out);
}

// Marsaglia-Tsang gamma(alpha<1) via gamma(alpha+1) * U^(1/alpha); counter-based uniforms
__device__ __forceinline__ double u01(uint64_t x) { return ((x >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
__device__ double gamma_sample(double alpha, uint64_t seed, uint64_t a, uint64_t b)
{
    const double d = alpha + 1.0 - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
    uint64_t ctr = 0;
    for (int it = 0; it < 64; ++it) {
        double u1 = u01(rng_u64(seed, a, b, ctr++)), u2 = u01(rng_u64(seed, a, b, ctr++));
        double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        double v = 1.0 + c * z;
        if (v <= 0.0) continue;
        v = v * v * v;
        double u = u01(rng_u64(seed, a, b, ctr++));
        if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) {
            double uu = u01(rng_u64(seed, a, b, ctr++));
            return d * v * pow(uu, 1.0 / alpha);
        }
    }
    return alpha;
}

// priors of one expansion (mcts.py:176-188).  KIND 0: float32 probabilities [8100] per game
// (the predict() contract); KIND 1: bf16 logits, KIND 2: float32 logits -> softmax over the legal
// entries (== softmax over all 8100 then renormalised over the legal ones).
// Returns true when the legal mass is <= 0 (uniform python-float fallback).
template <int KIND>
__device__ __forceinline__ bool warp_priors(const void* policy, size_t row_stride, int g, const int16_t* acts, int n,
                                            float* pri /* smem [128] */)
{
    const int lane = lane_id();
    if (KIND == 0) {
        const float* p = reinterpret_cast<const float*>(policy) + (size_t)g * row_stride;
        for (int i = lane; i < n; i += 32) pri[i] = p[acts[i]];
        warp_sync();
        float sum = 0.0f;
        if (lane == 0)
            for (int i = 0; i < n; ++i) sum = __fadd_rn(sum, pri[i]);   // python sum(): sequential float32
        sum = warp_bcast_f(sum, 0);
        if (!(sum > 0.0f)) return true;
        for (int i = lane; i < n; i += 32) pri[i] = __fdiv_rn(pri[i], sum);
        warp_sync();
        return false;
    } else {
        float m = -INFINITY;
        for (int i = lane; i < n; i += 32) {
            float l = KIND == 1 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(policy)[(size_t)g * row_stride + acts[i]])
                                : reinterpret_cast<const float*>(policy)[(size_t)g * row_stride + acts[i]];
            pri[i] = l;
            m = fmaxf(m, l);
        }
        m = warp_max_f(m);
        float s = 0.0f;
        for (int i = lane; i < n; i += 32) {
            float e = __expf(pri[i] - m);
            pri[i] = e;
            s += e;
        }
        s = warp_sum_f(s);
        const float inv = 1.0f / s;
        for (int i = lane; i < n; i += 32) pri[i] *= inv;
        warp_sync();
        return false;
    }
}

// create the children of `node` (mcts.py:60-64); returns false on pool overflow
__device__ __forceinline__ bool warp_expand(const MctsState& M, int node, const int16_t* acts, int n, const float* pri,
                                            bool uniform, int root_game, const double* noise, bool have_noise)
{
    const int lane = lane_id();
    int base = 0;
    if (lane == 0) base = atomicAdd(M.alloc, n);
    base = warp_bcast(base, 0);
    if ((long long)base + n > M.cap_nodes) {
        if (lane == 0) atomicOr(M.error, 1);
        return false;
    }
    for (int i = lane; i < n; i += 32) {
        float p32 = uniform ? 0.0f : pri[i];
        if (have_noise) {
            // 0.75 * P + 0.25 * noise[i] -> np.float64 (mcts.py:117-121)
            double mixed;
            if (uniform) mixed = __dadd_rn(__dmul_rn(0.75, __ddiv_rn(1.0, (double)n)), __dmul_rn(0.25, noise[i]));
            else mixed = __dadd_rn((double)__fmul_rn(0.75f, p32), __dmul_rn(0.25, noise[i]));
            M.rootP64[(size_t)root_game * kMaxMoves + i] = mixed;
        }
        M.hot[base + i] = NodeHot{0.0, 0, p32};
        M.link[base + i] = NodeLink{-1, node, acts[i], 0, 0, 0};
    }
    if (lane == 0) {
        NodeLink l = M.link[node];
        l.child0 = base;
        l.nchild = (uint8_t)n;
        l.flags = have_noise ? 2 : (uniform ? 1 : 0);
        M.link[node] = l;
    }
    return true;
}

template <int KIND>
__global__ void __launch_bounds__(kSelWarps * 32)
mcts_root_expand_kernel(MctsState M, const void* policy, size_t row_stride, const double* noise_in, int add_noise,
                        uint64_t seed, double alpha)
{
    __shared__ float pri[kSelWarps][kMaxMoves];
    __shared__ double nz[kSelWarps][kMaxMoves];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kSelWarps + warp;
    if (g >= M.n_games) return;
    if (M.leaf_state[g] != kLeafEval) return;
    const int n = M.leaf_n[g];
    const int16_t* acts = M.leaf_actions + (size_t)g * kMaxMoves;
    bool uniform = warp_priors<KIND>(policy, row_stride, g, acts, n, pri[warp]);
    if (add_noise) {
        if (noise_in) {
            for (int i = lane; i < n; i += 32) nz[warp][i] = noise_in[(size_t)g * kMaxMoves + i];
        } else {
            // np.random.dirichlet([alpha]*n): normalised gamma(alpha) draws (statistical parity only)
            double s = 0.0;
            for (int i = lane; i < n; i += 32) {
                double x = gamma_sample(alpha, seed, (uint64_t)g, (uint64_t)i);
                nz[warp][i] = x;
                s += x;
            }
            s = warp_sum_d(s);
            for (int i = lane; i < n; i += 32) nz[warp][i] = nz[warp][i] / s;
        }
        warp_sync();
    }
    warp_expand(M, g, acts, n, pri[warp], uniform, g, nz[warp], add_noise != 0);
    if (lane == 0) atomicAdd((unsigned long long*)&M.stats[3], 1ull);
}

// ---- select: descend to a leaf, replay the moves, test termination, emit evaluator inputs ----
__global__ void __launch_bounds__(kSelWarps * 32)
mcts_select_kernel(MctsState M, double c_puct, float* planes_f32, __nv_bfloat16* x_planes, long long x_rows,
                   long long x_row0, int8_t* boards_out, int8_t* sides_out)
{
    __shared__ SelectSmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kSelWarps + warp;
    if (g >= M.n_games) return;
    // a game without a tree (finished, or no legal move at the root) idles
    if (M.link[g].child0 < 0) {
        if (lane == 0) M.leaf_state[g] = kLeafIdle;
        return;
    }
    int8_t* b = sm.board[warp];
    int8_t* ring = sm.ring[warp];
    GameMeta gm;
    warp_load_game(M, g, b, ring, gm);

    int node = g, depth = 0;
    NodeLink ln = M.link[node];
    const float c32 = (float)c_puct;
    while (ln.child0 >= 0) {
        const int nch = ln.nchild, c0 = ln.child0, mode = ln.flags;
        const double sqrt_parent = sqrt((double)M.hot[node].N);   // math.sqrt(self.visit_count)
        const float sp32 = (float)sqrt_parent;
        double best = -INFINITY;
        int best_i = 0x7fffffff;
        // The link record of every child travels WITH its statistics: the winner's link (its children, its move) is then in a
        // register of the lane that scored it, and a level of the descent is one memory round trip instead of two.
        uint4 best_ln = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
        for (int i = lane; i < nch; i += 32) {
            const NodeHot h = M.hot[c0 + i];
            const uint4 l4 = *reinterpret_cast<const uint4*>(&M.link[c0 + i]);
            const double q = h.N == 0 ? 0.0 : __ddiv_rn(h.W, (double)h.N);
            double score;
            if (mode == 0) {
                float u = __fmul_rn(c32, h.P);
                u = __fmul_rn(u, sp32);
                u = __fdiv_rn(u, (float)(1 + h.N));
                score = (double)__fadd_rn((float)q, u);
            } else {
                const double p = mode == 2 ? M.rootP64[(size_t)g * kMaxMoves + i] : __ddiv_rn(1.0, (double)nch);
                double u = __dmul_rn(c_puct, p);
                u = __dmul_rn(u, sqrt_parent);
                u = __ddiv_rn(u, (double)(1 + h.N));
                score = __dadd_rn(q, u);
            }
            if (score > best) {   // strict '>' : first maximum wins (i ascends within a lane)
                best = score;
                best_i = i;
                best_ln = l4;
            }
        }
        best_i = warp_argmax_first(best, best_i);
        if (best_i == 0x7fffffff) {   // all-NaN guard; cannot happen with finite inputs
            best_i = 0;
            best_ln = *reinterpret_cast<const uint4*>(&M.link[c0]);
        }
        node = c0 + best_i;
        {
            const uint4 w4 = warp_bcast16(best_ln, best_i & 31);       // child i was scored by lane i % 32
            ln = *reinterpret_cast<const NodeLink*>(&w4);
        }
        warp_make_move(b, ring, gm, ln.action);
        ++depth;
    }

    WarpScratch& S = sm.ws[warp];
    MovegenResult r = warp_movegen(b, gm.side, S);
    if (r.overflow && lane == 0) atomicOr(M.error, 2);
    const int w = warp_game_over(b, ring, gm, r);
    if (lane == 0) {
        atomicAdd((unsigned long long*)&M.stats[0], 1ull);
        atomicMax((unsigned long long*)&M.stats[2], (unsigned long long)depth);
    }
    if (w != 2) {
        // terminal leaf: value needs no evaluator; back up now (mcts.py:137-140,153)
        if (lane == 0) {
            M.leaf_state[g] = kLeafTerminal;
            M.leaf_node[g] = node;
            backup_path(M, node, w == 0 ? 0.0 : 1.0);
            atomicAdd((unsigned long long*)&M.stats[1], 1ull);
        }
        return;
    }
    reinterpret_cast<uint2*>(M.leaf_actions + (size_t)g * kMaxMoves)[lane] = reinterpret_cast<const uint2*>(S.actions)[lane];
    if (lane == 0) {
        M.leaf_state[g] = kLeafEval;
        M.leaf_node[g] = node;
        M.leaf_n[g] = min(r.n_legal, kMaxMoves);
    }
    warp_emit_eval_inputs(b, gm.side, g, planes_f32, x_planes, x_rows, x_row0, boards_out, sides_out);
}

// ---- expand + backup ---------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(kSelWarps * 32)
mcts_expand_backup_kernel(MctsState M, const void* policy, size_t row_stride, const float* value)
{
    __shared__ float pri[kSelWarps][kMaxMoves];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kSelWarps + warp;
    if (g >= M.n_games) return;
    if (M.leaf_state[g] != kLeafEval) return;
    const int node = M.leaf_node[g];
    const int n = M.leaf_n[g];
    const int16_t* acts = M.leaf_actions + (size_t)g * kMaxMoves;
    bool uniform = warp_priors<KIND>(policy, row_stride, g, acts, n, pri[warp]);
    warp_expand(M, node, acts, n, pri[warp], uniform, g, nullptr, false);
    warp_sync();
    if (lane == 0) {
        __threadfence();
        backup_path(M, node, -(double)value[g]);   // value = -value (mcts.py:150)
        atomicAdd((unsigned long long*)&M.stats[3], 1ull);
    }
}

// ---- results ---------------------------------------------------------------------------------
__global__ void mcts_root_visits_kernel(MctsState M, int16_t* actions, int32_t* visits, int32_t* n_out, double* w_out)
{
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= M.n_games) return;
    const NodeLink l = M.link[g];
    const int n = l.child0 >= 0 ? l.nchild : 0;
    for (int i = lane; i < kMaxMoves; i += 32) {
        const bool ok = i < n;
        actions[(size_t)g * kMaxMoves + i] = ok ? M.link[l.child0 + i].action : (int16_t)-1;
        visits[(size_t)g * kMaxMoves + i] = ok ? M.hot[l.child0 + i].N : 0;
        if (w_out) w_out[(size_t)g * kMaxMoves + i] = ok ? M.hot[l.child0 + i].W : 0.0;
    }
    if (lane == 0) n_out[g] = n;
}

// priors of the root children as the select arithmetic sees them: float32 P widened (flags 0), 1/n (flags 1: zero-mass
// fallback) or the float64 noisy mix 0.75 P + 0.25 eta (flags 2, mcts.py:117-121)
__global__ void mcts_root_priors_kernel(MctsState M, double* priors, int32_t* n_out)
{
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= M.n_games) return;
    const NodeLink l = M.link[g];
    const int n = l.child0 >= 0 ? l.nchild : 0;
    const int mode = l.flags & 3;
    for (int i = lane; i < kMaxMoves; i += 32) {
        double p = 0.0;
        if (i < n) p = mode == 2 ? M.rootP64[(size_t)g * kMaxMoves + i] : (mode == 1 ? __ddiv_rn(1.0, (double)n) : (double)M.hot[l.child0 + i].P);
        priors[(size_t)g * kMaxMoves + i] = p;
    }
    if (lane == 0 && n_out) n_out[g] = n;
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

__global__ void mcts_set_games_kernel(MctsState M, const int8_t* boards, const int8_t* sides, const int32_t* move_count,
                                      const int32_t* no_capture, const int8_t* ring, const uint8_t* active)
{
    const int g = blockIdx.x;
    if (g >= M.n_games) return;
    for (int i = threadIdx.x; i < kBoardPad; i += blockDim.x)
        M.board[(size_t)g * kBoardPad + i] = i < kSquares ? boards[(size_t)g * kSquares + i] : (int8_t)0;
    for (int i = threadIdx.x; i < kRing * kBoardPad; i += blockDim.x) {
        const int s = i / kBoardPad, c = i % kBoardPad;
        M.ring[(size_t)g * kRing * kBoardPad + i] = (ring && c < kSquares) ? ring[((size_t)g * kRing + s) * kSquares + c] : (int8_t)0;
    }
    if (threadIdx.x == 0) {
        M.meta[g * 4 + 0] = sides[g];
        M.meta[g * 4 + 1] = move_count ? move_count[g] : 0;
        M.meta[g * 4 + 2] = no_capture ? no_capture[g] : 0;
        M.meta[g * 4 + 3] = active ? active[g] : 1;
    }
}

}  // namespace xq

using namespace xq;

static MctsState* S_(xq_ctx* c) { return reinterpret_cast<MctsState*>(c->mcts); }

extern "C" void xq_mcts_free_(xq_ctx* c)
{
    MctsState* M = S_(c);
    if (!M) return;
    void* ptrs[] = {M->board, M->ring, M->meta, M->hot, M->link, M->rootP64, M->alloc, M->error,
                    M->leaf_node, M->leaf_state, M->leaf_actions, M->leaf_n, M->stats, M->root_winner,
                    M->leaf_row, M->sims_left, M->n_eval};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete M;
    c->mcts = nullptr;
}

extern "C" int xq_mcts_create(xq_ctx* c, int max_games, long long node_capacity)
{
    if (!c || max_games <= 0) return xq_fail(c, XQ_ERR_ARG, "xq_mcts_create: bad arguments");
    XQ_CUDA(c, cudaSetDevice(c->device));
    if (node_capacity <= 0) node_capacity = (long long)max_games * 801 * 64;
    if (node_capacity > 0x7fffff00ll) node_capacity = 0x7fffff00ll;
    // One search state per context, shared by xq_mcts_* and xq_selfplay_*: it only ever grows, so a small
    // request (a single-game MCTS object) never shrinks the arrays a self-play engine was built on.
    if (MctsState* old = S_(c)) {
        if (old->max_games >= max_games && old->cap_nodes >= node_capacity) {
            old->n_games = 0;
            return XQ_OK;
        }
        if (old->max_games > max_games) max_games = old->max_games;
        if (old->cap_nodes > node_capacity) node_capacity = old->cap_nodes;
    }
    xq_mcts_free_(c);
    MctsState* M = new MctsState();
    c->mcts = M;
    M->max_games = max_games;
    M->cap_nodes = node_capacity;
    const size_t G = (size_t)max_games;
    XQ_CUDA(c, cudaMalloc(&M->board, G * kBoardPad));
    XQ_CUDA(c, cudaMalloc(&M->ring, G * kRing * kBoardPad));
    XQ_CUDA(c, cudaMalloc(&M->meta, G * 4 * sizeof(int32_t)));
    XQ_CUDA(c, cudaMalloc(&M->hot, (size_t)node_capacity * sizeof(NodeHot)));
    XQ_CUDA(c, cudaMalloc(&M->link, (size_t)node_capacity * sizeof(NodeLink)));
    XQ_CUDA(c, cudaMalloc(&M->rootP64, G * kMaxMoves * sizeof(double)));
    XQ_CUDA(c, cudaMalloc(&M->alloc, sizeof(int)));
    XQ_CUDA(c, cudaMalloc(&M->error, sizeof(int)));
    XQ_CUDA(c, cudaMalloc(&M->leaf_node, G * sizeof(int32_t)));
    XQ_CUDA(c, cudaMalloc(&M->leaf_state, G * sizeof(int32_t)));
    XQ_CUDA(c, cudaMalloc(&M->leaf_actions, G * kMaxMoves * sizeof(int16_t)));
    XQ_CUDA(c, cudaMalloc(&M->leaf_n, G * sizeof(int32_t)));
    XQ_CUDA(c, cudaMalloc(&M->stats, 4 * sizeof(int64_t)));
    XQ_CUDA(c, cudaMalloc(&M->root_winner, G * sizeof(int32_t)));
    XQ_CUDA(c, cudaMalloc(&M->leaf_row, G * sizeof(int32_t)));
    XQ_CUDA(c, cudaMalloc(&M->sims_left, G * sizeof(int32_t)));
    XQ_CUDA(c, cudaMalloc(&M->n_eval, 2 * sizeof(int)));
    XQ_CUDA(c, cudaMemset(M->sims_left, 0, G * sizeof(int32_t)));
    XQ_CUDA(c, cudaMemset(M->n_eval, 0, 2 * sizeof(int)));
    M->leaf_k = 1;
    XQ_CUDA(c, cudaMemset(M->error, 0, sizeof(int)));
    XQ_CUDA(c, cudaMemset(M->stats, 0, 4 * sizeof(int64_t)));
    XQ_CUDA(c, cudaMemset(M->meta, 0, G * 4 * sizeof(int32_t)));
    return XQ_OK;
}

// leaf slots per game for the multi-leaf search of the self-play loop (grow-only)
static int mcts_reserve_leaves(xq_ctx* c, MctsState* M, int K)
{
    if (K <= M->leaf_k) return XQ_OK;
    XQ_CUDA(c, cudaDeviceSynchronize());
    const size_t n = (size_t)M->max_games * (size_t)K;
    void* old[] = {M->leaf_node, M->leaf_state, M->leaf_actions, M->leaf_n, M->leaf_row};
    for (void* q : old)
        if (q) cudaFree(q);
    M->leaf_node = M->leaf_state = M->leaf_n = M->leaf_row = nullptr;
    M->leaf_actions = nullptr;
    XQ_CUDA(c, cudaMalloc(&M->leaf_node, n * sizeof(int32_t)));
    XQ_CUDA(c, cudaMalloc(&M->leaf_state, n * sizeof(int32_t)));
    XQ_CUDA(c, cudaMalloc(&M->leaf_actions, n * kMaxMoves * sizeof(int16_t)));
    XQ_CUDA(c, cudaMalloc(&M->leaf_n, n * sizeof(int32_t)));
    XQ_CUDA(c, cudaMalloc(&M->leaf_row, n * sizeof(int32_t)));
    M->leaf_k = K;
    return XQ_OK;
}

#define NEED_MCTS(c)                                                                    \
    MctsState* Mp = (c) ? S_(c) : nullptr;                                              \
    if (!Mp) return xq_fail(c, XQ_ERR_STATE, "%s: call xq_mcts_create first", __func__); \
    MctsState& M = *Mp;                                                                 \
    cudaStream_t s = (cudaStream_t)stream;

static inline int blocks_for(int n) { return (n + kSelWarps - 1) / kSelWarps; }

extern "C" int xq_mcts_set_games(xq_ctx* c, int n_games, const int8_t* d_boards, const int8_t* d_sides,
                                 const int32_t* d_move_count, const int32_t* d_no_capture, const int8_t* d_ring,
                                 const uint8_t* d_active, void* stream)
{
    NEED_MCTS(c);
    if (n_games < 0 || n_games > M.max_games || (n_games && (!d_boards || !d_sides)))
        return xq_fail(c, XQ_ERR_ARG, "xq_mcts_set_games: bad arguments (n_games=%d, max %d)", n_games, M.max_games);
    M.n_games = n_games;
    if (n_games == 0) return XQ_OK;
    mcts_set_games_kernel<<<n_games, 128, 0, s>>>(M, d_boards, d_sides, d_move_count, d_no_capture, d_ring, d_active);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_mcts_root_begin(xq_ctx* c, float* d_planes, void* d_x_planes, long long x_rows, long long x_row0,
                                  int8_t* d_boards_out, int8_t* d_sides_out, void* stream)
{
    NEED_MCTS(c);
    if (M.n_games == 0) return XQ_OK;
    if (d_boards_out && !d_sides_out) return xq_fail(c, XQ_ERR_ARG, "xq_mcts_root_begin: sides_out missing");
    // a new search: the pool restarts after the reserved root slots (one tree per search, mcts.py:104)
    set_int_kernel<<<1, 1, 0, s>>>(M.alloc, M.max_games);
    mcts_root_begin_kernel<<<blocks_for(M.n_games), kSelWarps * 32, 0, s>>>(M, d_planes, (__nv_bfloat16*)d_x_planes, x_rows, x_row0,
                                                                            d_boards_out, d_sides_out);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_mcts_root_expand(xq_ctx* c, const void* d_policy, int policy_kind, long long row_stride,
                                   const double* d_noise, int add_noise, uint64_t noise_seed, double alpha, void* stream)
{
    NEED_MCTS(c);
    if (M.n_games == 0) return XQ_OK;
    if (!d_policy || policy_kind < 0 || policy_kind > 2) return xq_fail(c, XQ_ERR_ARG, "xq_mcts_root_expand: bad policy");
    const int nb = blocks_for(M.n_games), nt = kSelWarps * 32;
    if (policy_kind == 0)
        mcts_root_expand_kernel<0><<<nb, nt, 0, s>>>(M, d_policy, (size_t)row_stride, d_noise, add_noise, noise_seed, alpha);
    else if (policy_kind == 1)
        mcts_root_expand_kernel<1><<<nb, nt, 0, s>>>(M, d_policy, (size_t)row_stride, d_noise, add_noise, noise_seed, alpha);
    else
        mcts_root_expand_kernel<2><<<nb, nt, 0, s>>>(M, d_policy, (size_t)row_stride, d_noise, add_noise, noise_seed, alpha);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_mcts_select(xq_ctx* c, double c_puct, float* d_planes, void* d_x_planes, long long x_rows,
                              long long x_row0, int8_t* d_boards_out, int8_t* d_sides_out, void* stream)
{
    NEED_MCTS(c);
    if (M.n_games == 0) return XQ_OK;
    if (d_boards_out && !d_sides_out) return xq_fail(c, XQ_ERR_ARG, "xq_mcts_select: sides_out missing");
    {
        XqTimer tm(c, s);
        mcts_select_kernel<<<blocks_for(M.n_games), kSelWarps * 32, 0, s>>>(M, c_puct, d_planes, (__nv_bfloat16*)d_x_planes, x_rows,
                                                                            x_row0, d_boards_out, d_sides_out);
    }
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_mcts_expand_backup(xq_ctx* c, const void* d_policy, int policy_kind, long long row_stride,
                                     const float* d_value, void* stream)
{
    NEED_MCTS(c);
    if (M.n_games == 0) return XQ_OK;
    if (!d_policy || !d_value || policy_kind < 0 || policy_kind > 2)
        return xq_fail(c, XQ_ERR_ARG, "xq_mcts_expand_backup: bad arguments");
    const int nb = blocks_for(M.n_games), nt = kSelWarps * 32;
    if (policy_kind == 0)
        mcts_expand_backup_kernel<0><<<nb, nt, 0, s>>>(M, d_policy, (size_t)row_stride, d_value);
    else if (policy_kind == 1)
        mcts_expand_backup_kernel<1><<<nb, nt, 0, s>>>(M, d_policy, (size_t)row_stride, d_value);
    else
        mcts_expand_backup_kernel<2><<<nb, nt, 0, s>>>(M, d_policy, (size_t)row_stride, d_value);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_mcts_leaf_info(xq_ctx* c, int32_t* d_state, int32_t* d_n, int16_t* d_actions, void* stream)
{
    NEED_MCTS(c);
    const size_t G = (size_t)M.n_games;
    if (G == 0) return XQ_OK;
    if (d_state) XQ_CUDA(c, cudaMemcpyAsync(d_state, M.leaf_state, G * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (d_n) XQ_CUDA(c, cudaMemcpyAsync(d_n, M.leaf_n, G * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (d_actions)
        XQ_CUDA(c, cudaMemcpyAsync(d_actions, M.leaf_actions, G * kMaxMoves * sizeof(int16_t), cudaMemcpyDeviceToDevice, s));
    return XQ_OK;
}

extern "C" int xq_mcts_root_visits(xq_ctx* c, int16_t* d_actions, int32_t* d_visits, int32_t* d_n, double* d_w,
                                   void* stream)
{
    NEED_MCTS(c);
    if (M.n_games == 0) return XQ_OK;
    if (!d_actions || !d_visits || !d_n) return xq_fail(c, XQ_ERR_ARG, "xq_mcts_root_visits: bad arguments");
    mcts_root_visits_kernel<<<blocks_for(M.n_games), kSelWarps * 32, 0, s>>>(M, d_actions, d_visits, d_n, d_w);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

extern "C" int xq_mcts_root_priors(xq_ctx* c, double* d_priors, int32_t* d_n, void* stream)
{
    NEED_MCTS(c);
    if (M.n_games == 0) return XQ_OK;
    if (!d_priors) return xq_fail(c, XQ_ERR_ARG, "xq_mcts_root_priors: bad arguments");
    mcts_root_priors_kernel<<<blocks_for(M.n_games), kSelWarps * 32, 0, s>>>(M, d_priors, d_n);
    c->launches += 1;
    XQ_CUDA(c, cudaGetLastError());
    return XQ_OK;
}

// stats[0..3] = simulations, terminal-leaf simulations, max depth, evaluator calls consumed;
// stats[4] = nodes allocated in the current search; stats[5] = error bits (1 pool overflow, 2 move overflow)
extern "C" int xq_mcts_stats(xq_ctx* c, long long* h_stats6, int reset)
{
    MctsState* Mp = c ? S_(c) : nullptr;
    if (!Mp) return xq_fail(c, XQ_ERR_STATE, "xq_mcts_stats: call xq_mcts_create first");
    int64_t st[4];
    int alloc = 0, err = 0;
    XQ_CUDA(c, cudaMemcpy(st, Mp->stats, sizeof(st), cudaMemcpyDeviceToHost));
    XQ_CUDA(c, cudaMemcpy(&alloc, Mp->alloc, sizeof(int), cudaMemcpyDeviceToHost));
    XQ_CUDA(c, cudaMemcpy(&err, Mp->error, sizeof(int), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) h_stats6[i] = st[i];
    h_stats6[4] = alloc - Mp->max_games;
    h_stats6[5] = err;
    if (reset) {
        XQ_CUDA(c, cudaMemset(Mp->stats, 0, sizeof(st)));
        XQ_CUDA(c, cudaMemset(Mp->error, 0, sizeof(int)));
    }
    return XQ_OK;
}


// =============================================================================================
// Device-resident self-play loop
// =============================================================================================
// Replaces parallel_selfplay.py:_play_one_game (:42-134) for n_slots games advanced in lockstep:
// random opening, per-ply search with root Dirichlet noise, temperature schedule (1.0 below
// temperature_threshold plies, 0.3 after), visit-count sampling, resign rule, z-labelling inputs
// (winner per game), and the sample records _augment_data / train.py consume.  A finished game
// is replaced by a fresh one in the same slot while games remain to be started, so the evaluator
// batch stays full.  Nothing crosses PCIe inside the loop.
namespace xq {

struct SpConfig {
    int num_simulations;
    float c_puct;
    int temperature_threshold;
    int max_game_length;
    int random_opening_moves;
    int enable_resign;
    float resign_threshold;
    int resign_check_steps;
    int add_noise;
    float dirichlet_alpha;
    unsigned long long seed;
    int target_games;
    // evaluation arena (train.py:453-535): greedy move choice (temperature 0: first maximum of the visit counts in child
    // order, mcts.py:197-200), no sample records, a game that reaches max_game_length undecided is a draw (:496-498)
    int arena;
    int16_t* move_log;              // [max_games_total][XQ_MAX_PLIES] actions played, or nullptr
    int leaves_per_game;            // K: descents per game and step (1 = mcts.py's one simulation at a time; > 1: virtual loss)
};

constexpr int kSampleBytes = 896;   // board 90 | side 1 | n 1 | uid 4 | ply 4 | played action 2 | pad | actions @128 (256) | probs @384 (512)

struct SpState {
    int n_slots = 0;
    int max_games_total = 0;
    long long sample_cap = 0;
    int32_t* status = nullptr;      // [slots] 0 = empty (wants a new game), 1 = playing
    int32_t* n_samples = nullptr;   // [slots] samples recorded by the current game (len(training_data))
    int32_t* resign_run = nullptr;  // [slots] consecutive resign-probe values below the threshold
    int32_t* game_uid = nullptr;    // [slots]
    int32_t* counters = nullptr;    // [8] started, finished, samples, red wins, black wins, draws, plies of finished games, dropped samples
    int8_t* res_winner = nullptr;   // [max_games_total]
    int16_t* res_plies = nullptr;   // [max_games_total]
    uint8_t* samples = nullptr;     // [sample_cap][kSampleBytes]
    unsigned long long ply_counter = 0;   // host side: plies played since reset (RNG stream index)
    int live_bound = 0;                   // host side: upper bound on the games still alive (0 = unknown: the slot count), xq_selfplay_set_live_bound
    // host side: the lockstep step (select -> forward(s) -> expand/backup) as an instantiated CUDA graph, rebuilt when anything
    // that is baked into it changes (sp_play_loop)
    cudaStream_t cap_stream = nullptr;
    cudaGraphExec_t step_exec = nullptr;
    unsigned long long step_key = 0;
    long long step_launches = 0;
};

__device__ __forceinline__ void warp_store_game(const MctsState& M, int g, const int8_t* b, const int8_t* ring, const GameMeta& gm,
                                                int active)
{
    const int lane = lane_id();
    uint4* gb = reinterpret_cast<uint4*>(M.board + (size_t)g * kBoardPad);
    uint4* gr = reinterpret_cast<uint4*>(M.ring + (size_t)g * kRing * kBoardPad);
    if (lane < 6) gb[lane] = reinterpret_cast<const uint4*>(b)[lane];
    for (int i = lane; i < 72; i += 32) gr[i] = reinterpret_cast<const uint4*>(ring)[i];
    if (lane == 0) {
        M.meta[g * 4 + 0] = gm.side;
        M.meta[g * 4 + 1] = gm.move_count;
        M.meta[g * 4 + 2] = gm.no_capture;
        M.meta[g * 4 + 3] = active;
    }
}

__device__ __forceinline__ void warp_start_position(int8_t* b)
{
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

// fill empty slots with fresh games: start position + k ~ U{0..random_opening_moves} uniformly random
// legal plies; a game that ends inside its opening restarts from the start position (parallel_selfplay.py:60-72)
__global__ void __launch_bounds__(kSelWarps * 32) sp_new_games_kernel(MctsState M, SpState P, SpConfig cfg, unsigned long long ply_idx)
{
    __shared__ SelectSmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kSelWarps + warp;
    if (g >= M.n_games) return;
    if (P.status[g] != 0) return;
    int uid = -1;
    if (lane == 0) {
        uid = atomicAdd(&P.counters[0], 1);
        if (uid >= cfg.target_games) {
            atomicSub(&P.counters[0], 1);
            uid = -1;
        }
    }
    uid = warp_bcast(uid, 0);
    int8_t* b = sm.board[warp];
    int8_t* ring = sm.ring[warp];
    WarpScratch& S = sm.ws[warp];
    GameMeta gm{1, 0, 0};
    warp_start_position(b);
    for (int i = lane; i < kRing * kBoardPad; i += 32) ring[i] = 0;
    warp_sync();
    if (uid < 0) {   // nothing left to start: the slot idles
        warp_store_game(M, g, b, ring, gm, 0);
        return;
    }
    const int k = (int)(rng_u64(cfg.seed, 0x0A11CEull + (uint64_t)uid, ply_idx, 1) % (uint64_t)(cfg.random_opening_moves + 1));
    for (int i = 0; i < k; ++i) {
        MovegenResult r = warp_movegen(b, gm.side, S);
        if (r.n_legal == 0) break;
        const int pick = (int)(rng_u64(cfg.seed, 0x0A11CEull + (uint64_t)uid, ply_idx, 2 + i) % (uint64_t)min(r.n_legal, kMaxMoves));
        const int action = S.actions[pick];
        warp_sync();
        warp_make_move(b, ring, gm, action);
        MovegenResult r2 = warp_movegen(b, gm.side, S);
        if (warp_game_over(b, ring, gm, r2) != 2) {
            warp_start_position(b);
            for (int j = lane; j < kRing * kBoardPad; j += 32) ring[j] = 0;
            gm = GameMeta{1, 0, 0};
            warp_sync();
            break;
        }
    }
    warp_store_game(M, g, b, ring, gm, 1);
    if (lane == 0) {
        P.status[g] = 1;
        P.n_samples[g] = 0;
        P.resign_run[g] = 0;
        P.game_uid[g] = uid;
    }
}

__device__ __forceinline__ void sp_finish_game(const MctsState& M, const SpState& P, int g, int winner, int plies)
{
    // lane 0 only
    const int uid = P.game_uid[g];
    if (uid >= 0 && uid < P.max_games_total) {
        P.res_winner[uid] = (int8_t)winner;
        P.res_plies[uid] = (int16_t)plies;
    }
    atomicAdd(&P.counters[1], 1);
    atomicAdd(&P.counters[winner == 1 ? 3 : (winner == -1 ? 4 : 5)], 1);
    atomicAdd(&P.counters[6], plies);
    P.status[g] = 0;
    M.meta[g * 4 + 3] = 0;
}

// ---- evaluator ports: where a leaf's network input goes and where its results come from ------------------
// Rows are COMPACTED: a leaf that needs the network takes the next row of its port's batch (atomic counter
// n_eval[port]), so a forward is sized to the leaves that are really waiting (xq_net_run_counted), whatever the
// number of idle slots, finished games and terminal leaves.  The arena has two ports (new / old model): a game's
// search belongs to the player to move at the ROOT (train.py:481-483), so each leaf goes to ONE network.
struct EvalPort {
    __nv_bfloat16* x_planes;
    long long x_rows, x_row0;
    const void* logits;
    size_t row_stride;
    const float* value;
};
struct StepArgs {
    EvalPort port[2];
    int arena;
    int K;                  // leaf slots per game and step (1 = the reference's one simulation at a time)
    int bound;              // rows the forwards of this step are sized for (the caller's upper bound on live games x K); a row beyond it sets error bit 4
};

__device__ __forceinline__ int sp_port_of(const SpState& P, const StepArgs& A, int g, int root_side)
{
    if (!A.arena) return 0;
    const int uid = P.game_uid[g];
    return (((uid & 1) == 0) == (root_side == 1)) ? 0 : 1;     // the new model is red in even games (train.py:474)
}

// root preparation of the self-play loop: as mcts_root_begin_kernel, for the playing slots only, rows compacted
__global__ void __launch_bounds__(kSelWarps * 32) sp_root_begin_kernel(MctsState M, SpState P, StepArgs A)
{
    __shared__ SelectSmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kSelWarps + warp;
    if (g >= M.n_games) return;
    const int slot = g * A.K;
    const bool active = M.meta[g * 4 + 3] != 0 && P.status[g] == 1;
    if (!active) {
        if (lane == 0) {
            M.leaf_n[slot] = 0;
            M.leaf_node[slot] = g;
            M.leaf_state[slot] = kLeafIdle;
            M.leaf_row[slot] = -1;
            M.sims_left[g] = 0;
            M.root_winner[g] = 2;
            M.hot[g] = NodeHot{0.0, 0, 0.0f};
            M.link[g] = NodeLink{-1, -1, (int16_t)-1, 0, 0, 0};
        }
        return;
    }
    int8_t* b = sm.board[warp];
    GameMeta gm;
    warp_load_game(M, g, b, sm.ring[warp], gm);
    WarpScratch& S = sm.ws[warp];
    MovegenResult r = warp_movegen(b, gm.side, S);
    if (r.overflow && lane == 0) atomicOr(M.error, 2);
    const int rw = warp_game_over(b, sm.ring[warp], gm, r);
    const int n = min(r.n_legal, kMaxMoves);
    reinterpret_cast<uint2*>(M.leaf_actions + (size_t)slot * kMaxMoves)[lane] = reinterpret_cast<const uint2*>(S.actions)[lane];
    const int port = sp_port_of(P, A, g, gm.side);
    int row = 0;
    if (lane == 0) {
        row = atomicAdd(&M.n_eval[port], 1);       // every playing root is evaluated: the resign probe reads its value
        if (row >= A.bound) atomicOr(M.error, 4);   // the host's live-games bound was wrong: this leaf is not evaluated
        M.root_winner[g] = rw;
        M.leaf_n[slot] = n;
        M.leaf_node[slot] = g;
        M.leaf_state[slot] = n > 0 ? kLeafEval : kLeafIdle;
        M.leaf_row[slot] = row;
        M.sims_left[g] = 0;
        M.hot[g] = NodeHot{0.0, 0, 0.0f};          // fresh root (mcts.py:104)
        M.link[g] = NodeLink{-1, -1, (int16_t)-1, 0, 0, 0};
    }
    row = warp_bcast(row, 0);
    const EvalPort& e = A.port[port];
    warp_emit_eval_inputs(b, gm.side, row, nullptr, e.x_planes, e.x_rows, e.x_row0, nullptr, nullptr);
}

// After the root evaluation: resign rule, termination, length adjudication, else expand the root
// (parallel_selfplay.py:74-95,110-121 in the reference's order: resign probe first).
template <int KIND>
__global__ void __launch_bounds__(kSelWarps * 32)
sp_after_root_kernel(MctsState M, SpState P, SpConfig cfg, StepArgs A, unsigned long long ply_idx)
{
    __shared__ float pri[kSelWarps][kMaxMoves];
    __shared__ double nz[kSelWarps][kMaxMoves];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kSelWarps + warp;
    if (g >= M.n_games) return;
    if (P.status[g] != 1) return;
    const int slot = g * A.K;
    const int row = M.leaf_row[slot];
    if (row < 0) return;
    const int side = M.meta[g * 4 + 0], move_count = M.meta[g * 4 + 1];
    const EvalPort& e = A.port[sp_port_of(P, A, g, side)];
    // resign probe: the value of the position just reached, for the side to move
    if (cfg.enable_resign && P.n_samples[g] > 10) {
        // lane 0 alone reads, updates and stores the counter; the warp gets its value by broadcast (every lane
        // loading it while lane 0 stores would be a race, and a lane seeing the new value would leave the warp)
        int run = 0;
        if (lane == 0) {
            run = e.value[row] < cfg.resign_threshold ? P.resign_run[g] + 1 : 0;
            P.resign_run[g] = run;
        }
        run = warp_bcast(run, 0);
        if (run >= cfg.resign_check_steps) {
            if (lane == 0) sp_finish_game(M, P, g, -side, move_count);
            return;
        }
    }
    const int rw = M.root_winner[g];
    if (rw != 2) {
        if (lane == 0) sp_finish_game(M, P, g, rw, move_count);
        return;
    }
    if (move_count >= cfg.max_game_length && cfg.arena) {
        if (lane == 0) sp_finish_game(M, P, g, 0, move_count);
        return;
    }
    if (move_count >= cfg.max_game_length) {
        // material adjudication (parallel_selfplay.py:79-89); unreachable while max_game_length >= 200
        __shared__ int8_t tb[kSelWarps][kBoardPad];
        for (int i = lane; i < kBoardPad; i += 32) tb[warp][i] = M.board[(size_t)g * kBoardPad + i];
        warp_sync();
        const int diff = warp_material_diff(tb[warp]);
        if (lane == 0) sp_finish_game(M, P, g, diff > 30 ? 1 : (diff < -30 ? -1 : 0), move_count);
        return;
    }
    const int n = M.leaf_n[slot];
    const int16_t* acts = M.leaf_actions + (size_t)slot * kMaxMoves;
    bool uniform = warp_priors<KIND>(e.logits, e.row_stride, row, acts, n, pri[warp]);
    if (cfg.add_noise) {
        double s = 0.0;
        const uint64_t stream = ((uint64_t)P.game_uid[g] << 20) ^ ply_idx;
        for (int i = lane; i < n; i += 32) {
            double x = gamma_sample((double)cfg.dirichlet_alpha, cfg.seed ^ 0xD1B1C1E7ull, stream, (uint64_t)i);
            nz[warp][i] = x;
            s += x;
        }
        s = warp_sum_d(s);
        for (int i = lane; i < n; i += 32) nz[warp][i] = nz[warp][i] / s;
        warp_sync();
    }
    warp_expand(M, g, acts, n, pri[warp], uniform, g, nz[warp], cfg.add_noise != 0);
    if (lane == 0) {
        atomicAdd((unsigned long long*)&M.stats[3], 1ull);
        M.sims_left[g] = cfg.num_simulations;
    }
}

// ---- one lockstep step of the self-play search: up to K descents per game -----------------------------------
// K = 1 is mcts.py:126-153 exactly (the arithmetic of mcts_select_kernel / mcts_expand_backup_kernel).  K > 1 is the
// opt-in multi-leaf mode: a game runs K descents per step, each leaving a VIRTUAL LOSS on its path (N += 1, W -= 1 on
// every node, so the next descent of the step sees a worse q there and turns elsewhere); the K leaves are evaluated in
// the same forward and backed up in descent order, each replacing its virtual loss by the real value.  A descent that
// ends on a leaf already waiting for its evaluation in this step shares that evaluation (kLeafDup).  Small game counts
// (the arena's eval_games, the tail of an iteration) fill the tensor-core batch this way.
__global__ void __launch_bounds__(kSelWarps * 32) mcts_select_multi_kernel(MctsState M, SpState P, double c_puct, StepArgs A)
{
    __shared__ SelectSmem sm;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kSelWarps + warp;
    if (g >= M.n_games) return;
    const int K = A.K;
    const bool vl = K > 1;
    int left = M.link[g].child0 < 0 ? 0 : M.sims_left[g];       // no tree (finished, idle): nothing to do
    const int port = sp_port_of(P, A, g, M.meta[g * 4 + 0]);
    const EvalPort& e = A.port[port];
    int8_t* b = sm.board[warp];
    int8_t* ring = sm.ring[warp];
    const float c32 = (float)c_puct;
    for (int j = 0; j < K; ++j) {
        const int slot = g * K + j;
        if (left <= 0) {
            if (lane == 0) M.leaf_state[slot] = kLeafIdle;
            continue;
        }
        --left;
        GameMeta gm;
        warp_load_game(M, g, b, ring, gm);
        int node = g, depth = 0;
        NodeLink ln = M.link[node];
        if (vl) {
            if (lane == 0) {
                NodeHot h = M.hot[node];
                h.N += 1;
                h.W = __dadd_rn(h.W, -1.0);
                M.hot[node] = h;
            }
            warp_sync();
        }
        while (ln.child0 >= 0) {
            const int nch = ln.nchild, c0 = ln.child0, mode = ln.flags & 3;
            // with a virtual loss the parent's count includes this descent's own visit: sqrt(N - 1) keeps K = 1 arithmetic
            const double sqrt_parent = sqrt((double)(M.hot[node].N - (vl ? 1 : 0)));   // math.sqrt(self.visit_count)
            const float sp32 = (float)sqrt_parent;
            double best = -INFINITY;
            int best_i = 0x7fffffff;
            // the link record of every child travels with its statistics (see mcts_select_kernel): one round trip per level
            uint4 best_ln = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u);
            for (int i = lane; i < nch; i += 32) {
                const NodeHot h = M.hot[c0 + i];
                const uint4 l4 = *reinterpret_cast<const uint4*>(&M.link[c0 + i]);
                const double q = h.N == 0 ? 0.0 : __ddiv_rn(h.W, (double)h.N);
                double score;
                if (mode == 0) {
                    float u = __fmul_rn(c32, h.P);
                    u = __fmul_rn(u, sp32);
                    u = __fdiv_rn(u, (float)(1 + h.N));
                    score = (double)__fadd_rn((float)q, u);
                } else {
                    const double p = mode == 2 ? M.rootP64[(size_t)g * kMaxMoves + i] : __ddiv_rn(1.0, (double)nch);
                    double u = __dmul_rn(c_puct, p);
                    u = __dmul_rn(u, sqrt_parent);
                    u = __ddiv_rn(u, (double)(1 + h.N));
                    score = __dadd_rn(q, u);
                }
                if (score > best) {   // strict '>' : first maximum wins (i ascends within a lane)
                    best = score;
                    best_i = i;
                    best_ln = l4;
                }
            }
            best_i = warp_argmax_first(best, best_i);
            if (best_i == 0x7fffffff) {
                best_i = 0;
                best_ln = *reinterpret_cast<const uint4*>(&M.link[c0]);
            }
            node = c0 + best_i;
            {
                const uint4 w4 = warp_bcast16(best_ln, best_i & 31);   // child i was scored by lane i % 32
                ln = *reinterpret_cast<const NodeLink*>(&w4);
            }
            if (vl) {
                if (lane == 0) {
                    NodeHot h = M.hot[node];
                    h.N += 1;
                    h.W = __dadd_rn(h.W, -1.0);
                    M.hot[node] = h;
                }
                warp_sync();
            }
            warp_make_move(b, ring, gm, ln.action);
            ++depth;
        }
        if (lane == 0) {
            atomicAdd((unsigned long long*)&M.stats[0], 1ull);
            atomicMax((unsigned long long*)&M.stats[2], (unsigned long long)depth);
        }
        if (vl && (ln.flags & kPendingFlag)) {
            // an earlier descent of this step waits on this very leaf: share its evaluation
            if (lane == 0) {
                M.leaf_state[slot] = kLeafDup;
                M.leaf_node[slot] = node;
                M.leaf_row[slot] = ln.pad;
            }
            continue;
        }
        WarpScratch& S = sm.ws[warp];
        MovegenResult r = warp_movegen(b, gm.side, S);
        if (r.overflow && lane == 0) atomicOr(M.error, 2);
        const int w = warp_game_over(b, ring, gm, r);
        if (w != 2) {
            // terminal leaf: value needs no evaluator; back up now (mcts.py:137-140,153)
            if (lane == 0) {
                M.leaf_state[slot] = kLeafTerminal;
                M.leaf_node[slot] = node;
                if (vl) backup_path_vl(M, node, w == 0 ? 0.0 : 1.0);
                else backup_path(M, node, w == 0 ? 0.0 : 1.0);
                atomicAdd((unsigned long long*)&M.stats[1], 1ull);
            }
            warp_sync();
            continue;
        }
        reinterpret_cast<uint2*>(M.leaf_actions + (size_t)slot * kMaxMoves)[lane] = reinterpret_cast<const uint2*>(S.actions)[lane];
        int row = 0;
        if (lane == 0) {
            row = atomicAdd(&M.n_eval[port], 1);
            if (row >= A.bound) atomicOr(M.error, 4);   // the host's live-games bound was wrong: this leaf is not evaluated
            M.leaf_state[slot] = kLeafEval;
            M.leaf_node[slot] = node;
            M.leaf_n[slot] = min(r.n_legal, kMaxMoves);
            M.leaf_row[slot] = row;
            if (vl) {
                NodeLink l2 = M.link[node];
                l2.flags |= kPendingFlag;
                l2.pad = row;
                M.link[node] = l2;
            }
        }
        row = warp_bcast(row, 0);
        warp_emit_eval_inputs(b, gm.side, row, nullptr, e.x_planes, e.x_rows, e.x_row0, nullptr, nullptr);
        warp_sync();
    }
    if (lane == 0 && M.link[g].child0 >= 0) M.sims_left[g] = left;
}

template <int KIND>
__global__ void __launch_bounds__(kSelWarps * 32) mcts_expand_backup_multi_kernel(MctsState M, SpState P, StepArgs A)
{
    __shared__ float pri[kSelWarps][kMaxMoves];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kSelWarps + warp;
    if (g >= M.n_games) return;
    const int K = A.K;
    const bool vl = K > 1;
    const EvalPort& e = A.port[sp_port_of(P, A, g, M.meta[g * 4 + 0])];
    for (int j = 0; j < K; ++j) {
        const int slot = g * K + j;
        const int st = M.leaf_state[slot];
        if (st != kLeafEval && st != kLeafDup) continue;
        const int node = M.leaf_node[slot];
        const int row = M.leaf_row[slot];
        if (st == kLeafEval) {
            const int n = M.leaf_n[slot];
            const int16_t* acts = M.leaf_actions + (size_t)slot * kMaxMoves;
            bool uniform = warp_priors<KIND>(e.logits, e.row_stride, row, acts, n, pri[warp]);
            const bool ok = warp_expand(M, node, acts, n, pri[warp], uniform, g, nullptr, false);   // rewrites flags: pending bit gone
            if (!ok && vl && lane == 0) M.link[node].flags &= (uint8_t)~kPendingFlag;
            warp_sync();
        }
        if (lane == 0) {
            __threadfence();
            const double v = -(double)e.value[row];          // value = -value (mcts.py:150)
            if (vl) backup_path_vl(M, node, v);
            else backup_path(M, node, v);
            if (st == kLeafEval) atomicAdd((unsigned long long*)&M.stats[3], 1ull);
        }
        warp_sync();
    }
}

// End of a search: visit distribution -> sample record, temperature sampling, make the move
// (parallel_selfplay.py:91-107, mcts.py:190-206).
__global__ void __launch_bounds__(kSelWarps * 32)
sp_end_move_kernel(MctsState M, SpState P, SpConfig cfg, unsigned long long ply_idx)
{
    __shared__ SelectSmem sm;
    __shared__ double wts[kSelWarps][kMaxMoves];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = blockIdx.x * kSelWarps + warp;
    if (g >= M.n_games) return;
    if (P.status[g] != 1) return;
    const NodeLink root = M.link[g];
    if (root.child0 < 0) return;
    int8_t* b = sm.board[warp];
    int8_t* ring = sm.ring[warp];
    GameMeta gm;
    warp_load_game(M, g, b, ring, gm);
    const int n = root.nchild;
    const double inv_t = gm.move_count < cfg.temperature_threshold ? 1.0 : 1.0 / 0.3;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) {
        const int cnt = M.hot[root.child0 + i].N;
        const double w = cnt > 0 ? (inv_t == 1.0 ? (double)cnt : pow((double)cnt, inv_t)) : 0.0;
        wts[warp][i] = w;
        s += w;
    }
    s = warp_sum_d(s);
    warp_sync();
    // np.random.choice(8100, p=probs): inverse-CDF draw in child order (statistical parity only)
    int chosen = 0;
    if (lane == 0) {
        const double u = u01(rng_u64(cfg.seed ^ 0x5A3D1Eull, (uint64_t)P.game_uid[g], ply_idx, 7)) * s;
        double acc = 0.0;
        chosen = n - 1;
        for (int i = 0; i < n; ++i) {
            acc += wts[warp][i];
            if (u < acc) { chosen = i; break; }
        }
        while (chosen > 0 && wts[warp][chosen] == 0.0) --chosen;   // never pick an unvisited move by rounding
    }
    if (cfg.arena && lane == 0) {
        // temperature 0: max(children, key=visit_count) keeps the first maximum in child order (mcts.py:197-200)
        int best = -1;
        chosen = 0;
        for (int i = 0; i < n; ++i) {
            const int cnt = M.hot[root.child0 + i].N;
            if (cnt > best) { best = cnt; chosen = i; }
        }
    }
    chosen = warp_bcast(chosen, 0);
    const int action = M.link[root.child0 + chosen].action;
    if (cfg.move_log && lane == 0) {
        const int uid = P.game_uid[g];
        if (uid >= 0 && uid < P.max_games_total && gm.move_count < XQ_MAX_PLIES)
            cfg.move_log[(size_t)uid * XQ_MAX_PLIES + gm.move_count] = (int16_t)action;
    }
    // sample record
    int slot = -1;
    if (lane == 0 && !cfg.arena) {
        slot = atomicAdd(&P.counters[2], 1);
        if ((long long)slot >= P.sample_cap) {
            atomicSub(&P.counters[2], 1);
            atomicAdd(&P.counters[7], 1);
            slot = -1;
        }
    }
    slot = warp_bcast(slot, 0);
    if (slot >= 0) {
        uint8_t* rec = P.samples + (size_t)slot * kSampleBytes;
        for (int i = lane; i < kSquares; i += 32) rec[i] = (uint8_t)b[i];
        if (lane == 0) {
            rec[90] = (uint8_t)(int8_t)gm.side;
            rec[91] = (uint8_t)n;
            *reinterpret_cast<int32_t*>(rec + 92) = P.game_uid[g];
            *reinterpret_cast<int32_t*>(rec + 96) = gm.move_count;
            *reinterpret_cast<int16_t*>(rec + 100) = (int16_t)action;
        }
        int16_t* ra = reinterpret_cast<int16_t*>(rec + 128);
        float* rp = reinterpret_cast<float*>(rec + 384);
        for (int i = lane; i < kMaxMoves; i += 32) {
            ra[i] = i < n ? M.link[root.child0 + i].action : (int16_t)-1;
            rp[i] = i < n ? (float)(wts[warp][i] / s) : 0.0f;
        }
    }
    warp_sync();
    warp_make_move(b, ring, gm, action);
    warp_store_game(M, g, b, ring, gm, 1);
    if (lane == 0) P.n_samples[g] += 1;
}

}  // namespace xq

static SpState* SP_(xq_ctx* c) { return reinterpret_cast<SpState*>(c->selfplay); }

static void sp_drop_step_graph(SpState* P)
{
    if (P->step_exec) {
        cudaDeviceSynchronize();                // no replay may be in flight when the executable graph goes
        cudaGraphExecDestroy(P->step_exec);
        P->step_exec = nullptr;
    }
    P->step_key = 0;
}

extern "C" void xq_selfplay_free_(xq_ctx* c)
{
    SpState* P = SP_(c);
    if (!P) return;
    void* ptrs[] = {P->status, P->n_samples, P->resign_run, P->game_uid, P->counters, P->res_winner, P->res_plies, P->samples};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    sp_drop_step_graph(P);
    if (P->cap_stream) cudaStreamDestroy(P->cap_stream);
    delete P;
    c->selfplay = nullptr;
}

extern "C" int xq_selfplay_create(xq_ctx* c, int n_slots, int max_games_total, long long sample_capacity,
                                  long long node_capacity)
{
    if (!c || n_slots <= 0 || max_games_total <= 0 || sample_capacity <= 0)
        return xq_fail(c, XQ_ERR_ARG, "xq_selfplay_create: bad arguments");
    int rc = xq_mcts_create(c, n_slots, node_capacity);
    if (rc) return rc;
    xq_selfplay_free_(c);
    SpState* P = new SpState();
    c->selfplay = P;
    P->n_slots = n_slots;
    P->max_games_total = max_games_total;
    P->sample_cap = sample_capacity;
    const size_t G = (size_t)n_slots;
    XQ_CUDA(c, cudaMalloc(&P->status, G * 4));
    XQ_CUDA(c, cudaMalloc(&P->n_samples, G * 4));
    XQ_CUDA(c, cudaMalloc(&P->resign_run, G * 4));
    XQ_CUDA(c, cudaMalloc(&P->game_uid, G * 4));
    XQ_CUDA(c, cudaMalloc(&P->counters, 8 * 4));
    XQ_CUDA(c, cudaMalloc(&P->res_winner, (size_t)max_games_total));
    XQ_CUDA(c, cudaMalloc(&P->res_plies, (size_t)max_games_total * 2));
    XQ_CUDA(c, cudaMalloc(&P->samples, (size_t)sample_capacity * kSampleBytes));
    XQ_CUDA(c, cudaMemset(P->status, 0, G * 4));
    XQ_CUDA(c, cudaMemset(P->counters, 0, 8 * 4));
    XQ_CUDA(c, cudaMemset(P->res_winner, 2, (size_t)max_games_total));
    S_(c)->n_games = n_slots;
    return XQ_OK;
}

extern "C" int xq_selfplay_reset(xq_ctx* c, void* stream)
{
    SpState* P = c ? SP_(c) : nullptr;
    if (!P) return xq_fail(c, XQ_ERR_STATE, "xq_selfplay_reset: call xq_selfplay_create first");
    cudaStream_t s = (cudaStream_t)stream;
    XQ_CUDA(c, cudaMemsetAsync(P->status, 0, (size_t)P->n_slots * 4, s));
    XQ_CUDA(c, cudaMemsetAsync(P->counters, 0, 8 * 4, s));
    XQ_CUDA(c, cudaMemsetAsync(P->res_winner, 2, (size_t)P->max_games_total, s));
    XQ_CUDA(c, cudaMemsetAsync(S_(c)->meta, 0, (size_t)P->n_slots * 16, s));
    XQ_CUDA(c, cudaMemsetAsync(S_(c)->stats, 0, 4 * sizeof(int64_t), s));
    XQ_CUDA(c, cudaMemsetAsync(S_(c)->error, 0, sizeof(int), s));
    P->ply_counter = 0;
    P->live_bound = 0;
    return XQ_OK;
}

extern "C" int xq_net_run_counted(xq_ctx* c, const xq_gemm_desc* layers, int n_layers, const float* d_vfeats, const float* d_w1t,
                                  const float* d_b1, const float* d_w2, float b2, float* d_value, const int* d_n_boards,
                                  int max_boards, void* stream);

extern "C" int xq_net_side_stream_(xq_ctx* c, cudaStream_t* side, cudaEvent_t* ev_fork, cudaEvent_t* ev_join);

static EvalPort port_of(const xq_net_plan* n)
{
    EvalPort e;
    e.x_planes = (__nv_bfloat16*)n->x_planes;
    e.x_rows = n->x_rows;
    e.x_row0 = n->x_row0;
    e.logits = n->logits;
    e.row_stride = (size_t)n->logit_stride;
    e.value = n->value;
    return e;
}

// The ply loop of self-play (one network) and of the arena (two networks, each evaluating only its own games'
// leaves): per ply  new games -> roots -> forward -> resign/terminal/expand -> ceil(S / K) x (select K leaves per game,
// forward over the compacted leaves, expand + backup) -> move.  No host synchronisation anywhere: the forwards read
// their batch size from the device counter the select kernel just filled.
static int sp_play_loop(xq_ctx* c, MctsState& M, SpState& P, const SpConfig& k, const xq_net_plan* net0, const xq_net_plan* net1,
                        int n_plies, cudaStream_t s)
{
    const int K = k.leaves_per_game;
    if (int rc = mcts_reserve_leaves(c, &M, K)) return rc;
    StepArgs A;
    A.port[0] = port_of(net0);
    A.port[1] = port_of(net1 ? net1 : net0);
    A.arena = net1 ? 1 : 0;
    A.K = K;
    // Forward launches are sized (grid, kernel variant) for `rows` boards and cut to the live count on the device.  Late in an
    // iteration few games are alive and what counts is the latency of a forward: with the caller's bound on the live games
    // the small-batch variants of the layers are chosen (xq_net.cu) instead of the ones sized for every slot.
    long long rows_ll = (long long)(P.live_bound > 0 && P.live_bound < M.n_games ? P.live_bound : M.n_games) * K;
    const int rows0 = (int)(rows_ll < net0->batch ? rows_ll : net0->batch);
    const int rows1 = net1 ? (int)(rows_ll < net1->batch ? rows_ll : net1->batch) : 0;
    A.bound = rows0;
    const int kind = net0->logits_kind;
    const int nb = blocks_for(M.n_games), nt = kSelWarps * 32;
    const int steps = (k.num_simulations + K - 1) / K;
    // The arena's two networks evaluate disjoint leaves with their own buffers: the second forward runs on a side stream
    // next to the first (with a few evaluation games per GPU a forward is a chain of small kernels that leaves most SMs idle).
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    if (net1 && !c->net_fork)
        if (int rc = xq_net_side_stream_(c, &side, &ev_fork, &ev_join)) return rc;
    auto run_nets = [&](cudaStream_t q) -> int {
        if (net1 && side) {
            XQ_CUDA(c, cudaEventRecord(ev_fork, q));
            XQ_CUDA(c, cudaStreamWaitEvent(side, ev_fork, 0));
            int rc1 = xq_net_run_counted(c, net1->layers, net1->n_layers, net1->vfeats, net1->w1t, net1->b1, net1->w2, net1->b2,
                                         net1->value, M.n_eval + 1, rows1, (void*)side);
            if (rc1) return rc1;
            XQ_CUDA(c, cudaEventRecord(ev_join, side));
        }
        int rc = xq_net_run_counted(c, net0->layers, net0->n_layers, net0->vfeats, net0->w1t, net0->b1, net0->w2, net0->b2, net0->value,
                                    M.n_eval + 0, rows0, (void*)q);
        if (rc || !net1) return rc;
        if (side) {
            XQ_CUDA(c, cudaStreamWaitEvent(q, ev_join, 0));
            return XQ_OK;
        }
        return xq_net_run_counted(c, net1->layers, net1->n_layers, net1->vfeats, net1->w1t, net1->b1, net1->w2, net1->b2, net1->value,
                                  M.n_eval + 1, rows1, (void*)q);
    };
    // one lockstep step: every game selects a leaf (or K), the live leaves are evaluated, every game expands and backs up
    auto enqueue_step = [&](cudaStream_t q) -> int {
        XQ_CUDA(c, cudaMemsetAsync(M.n_eval, 0, 2 * sizeof(int), q));
        mcts_select_multi_kernel<<<nb, nt, 0, q>>>(M, P, (double)k.c_puct, A);
        int rc = run_nets(q);
        if (rc) return rc;
        if (kind == 1) mcts_expand_backup_multi_kernel<1><<<nb, nt, 0, q>>>(M, P, A);
        else mcts_expand_backup_multi_kernel<2><<<nb, nt, 0, q>>>(M, P, A);
        c->launches += 2;
        return XQ_OK;
    };
    // With up to ~1000 boards per forward a step is 25-45 launches of kernels that take a few microseconds each, and the host
    // needs as long to issue them as the device to run them (arena, 32 games: 162 us of enqueue against 174 us on the device).
    // The step is therefore captured once -- on a private stream: the caller's may be the legacy default stream, which cannot
    // capture -- and replayed num_simulations times per ply.  Everything baked into the graph is in the key; the capture
    // happens after the first eager forward of the call, which has done every one-time initialisation.
    const bool want_graph = c->sp_graph && !c->timing && rows0 <= 1024 && steps >= 8;
    unsigned long long key = 0;
    if (want_graph) {
        auto mix = [&](unsigned long long v) { key = (key ^ v) * 0x9E3779B97F4A7C15ull + 0x7F4A7C15ull; };
        mix((unsigned long long)rows0); mix((unsigned long long)rows1); mix((unsigned long long)K); mix((unsigned long long)kind);
        mix((unsigned long long)M.n_games); mix((unsigned long long)(uintptr_t)net0->layers); mix((unsigned long long)(uintptr_t)net0->value);
        mix((unsigned long long)(uintptr_t)(net1 ? net1->layers : nullptr)); mix((unsigned long long)(uintptr_t)(net1 ? net1->value : nullptr));
        mix((unsigned long long)(uintptr_t)net0->logits); mix((unsigned long long)net0->n_layers); mix((unsigned long long)net0->batch);
        mix((unsigned long long)(uintptr_t)M.leaf_state); mix((unsigned long long)(uintptr_t)M.leaf_row); mix((unsigned long long)(uintptr_t)M.n_eval);
        double cp = (double)k.c_puct;
        unsigned long long cpb;
        memcpy(&cpb, &cp, sizeof(cpb));
        mix(cpb); mix((unsigned long long)(uintptr_t)M.hot); mix((unsigned long long)(uintptr_t)P.status); mix((unsigned long long)(uintptr_t)s);
        key |= 1ull;
        if (P.step_key != key) sp_drop_step_graph(&P);
    } else {
        sp_drop_step_graph(&P);
    }
    auto capture_step = [&]() -> int {
        if (!P.cap_stream) XQ_CUDA(c, cudaStreamCreateWithFlags(&P.cap_stream, cudaStreamNonBlocking));
        const long long l0 = c->launches;
        XQ_CUDA(c, cudaStreamBeginCapture(P.cap_stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_step(P.cap_stream);
        cudaGraph_t g = nullptr;
        const cudaError_t e_end = cudaStreamEndCapture(P.cap_stream, &g);
        P.step_launches = c->launches - l0;
        c->launches = l0;                                       // nothing ran: replays add the count
        if (rc || e_end != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            (void)cudaGetLastError();
            return rc ? rc : xq_fail(c, XQ_ERR_CUDA, "capture of the lockstep step failed: %s", cudaGetErrorString(e_end));
        }
        const cudaError_t e_inst = cudaGraphInstantiate(&P.step_exec, g, 0);
        cudaGraphDestroy(g);
        if (e_inst != cudaSuccess) {
            P.step_exec = nullptr;
            return xq_fail(c, XQ_ERR_CUDA, "instantiation of the lockstep step graph failed: %s", cudaGetErrorString(e_inst));
        }
        P.step_key = key;
        return XQ_OK;
    };
    for (int ply = 0; ply < n_plies; ++ply) {
        const unsigned long long pi = P.ply_counter++;
        sp_new_games_kernel<<<nb, nt, 0, s>>>(M, P, k, pi);
        set_int_kernel<<<1, 1, 0, s>>>(M.alloc, M.max_games);
        XQ_CUDA(c, cudaMemsetAsync(M.n_eval, 0, 2 * sizeof(int), s));
        sp_root_begin_kernel<<<nb, nt, 0, s>>>(M, P, A);
        c->launches += 3;
        int rc = run_nets(s);
        if (rc) return rc;
        if (kind == 1) sp_after_root_kernel<1><<<nb, nt, 0, s>>>(M, P, k, A, pi);
        else sp_after_root_kernel<2><<<nb, nt, 0, s>>>(M, P, k, A, pi);
        c->launches += 1;
        if (want_graph && !P.step_exec) {
            rc = capture_step();
            if (rc) {                                           // a driver that cannot capture this step: issue it launch by launch
                fprintf(stderr, "[xq_b200] %s: the lockstep step runs without a CUDA graph\n", c->err);
                c->sp_graph = false;
                sp_drop_step_graph(&P);
            }
        }
        for (int step = 0; step < steps; ++step) {
            if (P.step_exec && c->sp_graph) {
                XQ_CUDA(c, cudaGraphLaunch(P.step_exec, s));
                c->launches += P.step_launches;
            } else {
                rc = enqueue_step(s);
                if (rc) return rc;
            }
        }
        sp_end_move_kernel<<<nb, nt, 0, s>>>(M, P, k, pi);
        c->launches += 1;
        XQ_CUDA(c, cudaGetLastError());
    }
    return XQ_OK;
}

extern "C" int xq_selfplay_set_live_bound(xq_ctx* c, int max_live_games)
{
    SpState* Pp = c ? SP_(c) : nullptr;
    if (!Pp) return xq_fail(c, XQ_ERR_STATE, "xq_selfplay_set_live_bound: call xq_selfplay_create first");
    Pp->live_bound = max_live_games > 0 ? max_live_games : 0;
    return XQ_OK;
}

extern "C" int xq_selfplay_play(xq_ctx* c, const xq_selfplay_config* cfg, const xq_net_plan* net, int n_plies, void* stream)
{
    SpState* Pp = c ? SP_(c) : nullptr;
    MctsState* Mp = c ? S_(c) : nullptr;
    if (!Pp || !Mp) return xq_fail(c, XQ_ERR_STATE, "xq_selfplay_play: call xq_selfplay_create first");
    if (!cfg || !net || n_plies < 0) return xq_fail(c, XQ_ERR_ARG, "xq_selfplay_play: bad arguments");
    if (net->batch < Pp->n_slots) return xq_fail(c, XQ_ERR_ARG, "xq_selfplay_play: network batch %d < slots %d", net->batch, Pp->n_slots);
    if (net->logits_kind != 1 && net->logits_kind != 2) return xq_fail(c, XQ_ERR_ARG, "xq_selfplay_play: logits_kind must be 1 (bf16) or 2 (f32)");
    MctsState& M = *Mp;
    SpState& P = *Pp;
    if (M.max_games < P.n_slots) return xq_fail(c, XQ_ERR_STATE, "xq_selfplay_play: search state smaller than the slot count");
    M.n_games = P.n_slots;     // a standalone xq_mcts_set_games call may have changed it since the last ply
    cudaStream_t s = (cudaStream_t)stream;
    SpConfig k;
    k.num_simulations = cfg->num_simulations;
    k.c_puct = cfg->c_puct;
    k.temperature_threshold = cfg->temperature_threshold;
    k.max_game_length = cfg->max_game_length;
    k.random_opening_moves = cfg->random_opening_moves;
    k.enable_resign = cfg->enable_resign;
    k.resign_threshold = cfg->resign_threshold;
    k.resign_check_steps = cfg->resign_check_steps;
    k.add_noise = cfg->add_noise;
    k.dirichlet_alpha = cfg->dirichlet_alpha;
    k.seed = cfg->seed;
    k.target_games = cfg->target_games < P.max_games_total ? cfg->target_games : P.max_games_total;
    k.arena = 0;
    k.move_log = nullptr;
    k.leaves_per_game = cfg->leaves_per_game > 1 ? cfg->leaves_per_game : 1;
    if ((long long)P.n_slots * k.leaves_per_game > net->batch)
        return xq_fail(c, XQ_ERR_ARG, "xq_selfplay_play: %d slots x %d leaves per game exceed the network batch %d", P.n_slots,
                       k.leaves_per_game, net->batch);
    return sp_play_loop(c, M, P, k, net, nullptr, n_plies, s);
}

// ---- evaluation arena: new model vs best model (train.py:453-535) -------------------------------------------
// Game uid plays with the NEW model as red when uid is even (train.py:474).  Every search belongs to the player to
// move at the ROOT, so a game's leaves go to that player's network only (sp_port_of): each forward is sized to its
// own share of the leaves.
extern "C" int xq_arena_play(xq_ctx* c, const xq_selfplay_config* cfg, const xq_net_plan* net_new, const xq_net_plan* net_old,
                             int n_plies, int16_t* d_move_log, void* stream)
{
    SpState* Pp = c ? SP_(c) : nullptr;
    MctsState* Mp = c ? S_(c) : nullptr;
    if (!Pp || !Mp) return xq_fail(c, XQ_ERR_STATE, "xq_arena_play: call xq_selfplay_create first");
    if (!cfg || !net_new || !net_old || n_plies < 0) return xq_fail(c, XQ_ERR_ARG, "xq_arena_play: bad arguments");
    if (net_new->batch < Pp->n_slots || net_old->batch < Pp->n_slots || net_new->x_rows != net_old->x_rows ||
        net_new->x_row0 != net_old->x_row0 || net_new->logit_stride != net_old->logit_stride ||
        net_new->logits_kind != net_old->logits_kind || (net_new->logits_kind != 1 && net_new->logits_kind != 2))
        return xq_fail(c, XQ_ERR_ARG, "xq_arena_play: the two network plans must have the same batch geometry");
    MctsState& M = *Mp;
    SpState& P = *Pp;
    if (M.max_games < P.n_slots) return xq_fail(c, XQ_ERR_STATE, "xq_arena_play: search state smaller than the slot count");
    M.n_games = P.n_slots;
    cudaStream_t s = (cudaStream_t)stream;
    SpConfig k;
    k.num_simulations = cfg->num_simulations;
    k.c_puct = cfg->c_puct;
    k.temperature_threshold = 0;
    k.max_game_length = cfg->max_game_length;
    k.random_opening_moves = 0;          // evaluation games start from the initial position (train.py:473)
    k.enable_resign = 0;
    k.resign_threshold = -2.0f;
    k.resign_check_steps = 1 << 30;
    k.add_noise = 0;                     // get_action(game, temperature=0, add_noise=False) (train.py:481-483)
    k.dirichlet_alpha = cfg->dirichlet_alpha;
    k.seed = cfg->seed;
    k.target_games = cfg->target_games < P.max_games_total ? cfg->target_games : P.max_games_total;
    k.arena = 1;
    k.move_log = d_move_log;
    k.leaves_per_game = cfg->leaves_per_game > 1 ? cfg->leaves_per_game : 1;
    if ((long long)P.n_slots * k.leaves_per_game > net_new->batch)
        return xq_fail(c, XQ_ERR_ARG, "xq_arena_play: %d slots x %d leaves per game exceed the network batch %d", P.n_slots,
                       k.leaves_per_game, net_new->batch);
    return sp_play_loop(c, M, P, k, net_new, net_old, n_plies, s);
}

// device pointers of the sample records and per-game results, for consumers that stay on the GPU (the replay ring)
extern "C" int xq_selfplay_device_buffers(xq_ctx* c, void** d_samples, int8_t** d_winner, int16_t** d_plies)
{
    SpState* P = c ? SP_(c) : nullptr;
    if (!P) return xq_fail(c, XQ_ERR_STATE, "xq_selfplay_device_buffers: call xq_selfplay_create first");
    if (d_samples) *d_samples = P->samples;
    if (d_winner) *d_winner = P->res_winner;
    if (d_plies) *d_plies = P->res_plies;
    return XQ_OK;
}

// counters: games started, games finished, samples, red wins, black wins, draws, plies of finished games,
// dropped samples; then the 6 values of xq_mcts_stats.  Synchronises.
extern "C" int xq_selfplay_counters(xq_ctx* c, long long* h_out14)
{
    SpState* P = c ? SP_(c) : nullptr;
    if (!P) return xq_fail(c, XQ_ERR_STATE, "xq_selfplay_counters: call xq_selfplay_create first");
    int32_t v[8];
    XQ_CUDA(c, cudaMemcpy(v, P->counters, sizeof(v), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; ++i) h_out14[i] = v[i];
    return xq_mcts_stats(c, h_out14 + 8, 0);
}

// Copies sample records [first, first+count) (896 B each, see kSampleBytes) and the per-game results to HOST buffers.
extern "C" int xq_selfplay_fetch(xq_ctx* c, long long first, long long count, void* h_samples, int8_t* h_winner,
                                 int16_t* h_plies, int n_results)
{
    SpState* P = c ? SP_(c) : nullptr;
    if (!P) return xq_fail(c, XQ_ERR_STATE, "xq_selfplay_fetch: call xq_selfplay_create first");
    if (first < 0 || count < 0 || first + count > P->sample_cap || n_results > P->max_games_total)
        return xq_fail(c, XQ_ERR_ARG, "xq_selfplay_fetch: range out of bounds");
    if (count && h_samples)
        XQ_CUDA(c, cudaMemcpy(h_samples, P->samples + (size_t)first * kSampleBytes, (size_t)count * kSampleBytes, cudaMemcpyDeviceToHost));
    if (n_results > 0 && h_winner) XQ_CUDA(c, cudaMemcpy(h_winner, P->res_winner, (size_t)n_results, cudaMemcpyDeviceToHost));
    if (n_results > 0 && h_plies) XQ_CUDA(c, cudaMemcpy(h_plies, P->res_plies, (size_t)n_results * 2, cudaMemcpyDeviceToHost));
    return XQ_OK;
}

// current game states of all slots (device -> host), for inspection and tests
extern "C" int xq_selfplay_slots(xq_ctx* c, int8_t* h_boards /*[slots][90]*/, int32_t* h_meta /*[slots][4]*/, int32_t* h_status,
                                 int32_t* h_uid)
{
    SpState* P = c ? SP_(c) : nullptr;
    MctsState* M = c ? S_(c) : nullptr;
    if (!P || !M) return xq_fail(c, XQ_ERR_STATE, "xq_selfplay_slots: call xq_selfplay_create first");
    const int G = P->n_slots;
    if (h_boards)
        XQ_CUDA(c, cudaMemcpy2D(h_boards, 90, M->board, kBoardPad, 90, (size_t)G, cudaMemcpyDeviceToHost));
    if (h_meta) XQ_CUDA(c, cudaMemcpy(h_meta, M->meta, (size_t)G * 16, cudaMemcpyDeviceToHost));
    if (h_status) XQ_CUDA(c, cudaMemcpy(h_status, P->status, (size_t)G * 4, cudaMemcpyDeviceToHost));
    if (h_uid) XQ_CUDA(c, cudaMemcpy(h_uid, P->game_uid, (size_t)G * 4, cudaMemcpyDeviceToHost));
    return XQ_OK;
}
