"""Drop-in replacement for the reference's training/mcts.py: same `MCTS(model, num_simulations,
c_puct, device)` / `.search(game, temperature, add_noise)` / `.get_action(...)` surface, with the
tree on the GPU (xq_mcts_* in libxq_b200.so).

`model` is anything with the reference's evaluator contract `predict(state[, device]) ->
(float32[8100] probabilities, float)` (model.py:109-124, inference_server.py:333-349).  This
per-game interface runs a batch of one and calls `predict` once per simulation, exactly like
the reference; the high-throughput path is the lockstep batched search
(xq_native.MctsBatch / selfplay_engine) where thousands of games share each evaluator batch.
"""
import math
from typing import Optional

import numpy as np

import xq_native
from game import ACTION_SPACE, XiangqiGame, engine


class MCTSNode:
    """Host-side node with the reference's fields and methods (mcts.py:21-73).

    `MCTS.search` below does NOT use it -- the live tree is a flat device array (csrc/xq_mcts.cu).
    The class exists for callers that hand-roll the search loop around it, as the reference's own
    benchmark.py:18-153 does (root.expand / node.select_child / node.backup), so that script keeps
    running with the rules engine and the network on the GPU."""

    __slots__ = ['parent', 'children', 'visit_count', 'total_value', 'prior']

    def __init__(self, parent: Optional['MCTSNode'] = None, prior: float = 0.0):
        self.parent = parent
        self.children = {}
        self.visit_count = 0
        self.total_value = 0.0
        self.prior = prior

    @property
    def q_value(self) -> float:
        return 0.0 if self.visit_count == 0 else self.total_value / self.visit_count

    def is_leaf(self) -> bool:
        return not self.children

    def select_child(self, c_puct: float = 1.5):
        """PUCT argmax, strict '>' so the first maximum in insertion order wins (mcts.py:43-58)."""
        root_n = math.sqrt(self.visit_count)
        pick, pick_score = (-1, None), -math.inf
        for action, child in self.children.items():
            score = child.q_value + c_puct * child.prior * root_n / (1 + child.visit_count)
            if score > pick_score:
                pick_score, pick = score, (action, child)
        return pick

    def expand(self, action_priors):
        for action, prior in action_priors.items():
            self.children.setdefault(action, MCTSNode(parent=self, prior=prior))

    def backup(self, value: float):
        node = self
        while node is not None:                      # leaf -> root, sign flips per ply (mcts.py:66-73)
            node.visit_count += 1
            node.total_value += value
            value, node = -value, node.parent


def game_to_arrays(game: XiangqiGame):
    """XiangqiGame -> (board, side, move_count, no_capture, ring[12][90]) for xq_mcts_set_games.
    ring slot i%12 holds the board before move i: the last 12 entries of game.history."""
    ring = np.zeros((12, 90), np.int8)
    h = game.history
    for i in range(max(0, len(h) - 12), len(h)):
        ring[i % 12] = np.frombuffer(h[i], dtype=np.int8)
    return (np.ascontiguousarray(game.board, np.int8).reshape(1, 90), np.array([game.current_player], np.int8),
            np.array([game.move_count], np.int32), np.array([game.no_capture_count], np.int32), ring[None])


class MCTS:
    def __init__(self, model, num_simulations: int = 200, c_puct: float = 1.5, device: str = 'cpu'):
        self.model = model
        self.num_simulations = num_simulations
        self.c_puct = c_puct
        self.device = device
        self._batch = None
        self.last_root = None

    def _predict(self, state):
        try:
            return self.model.predict(state, self.device)      # mcts.py:157-164
        except TypeError:
            return self.model.predict(state)

    def _evaluator(self, mb):
        t = mb.t
        state = mb.planes[0].cpu().numpy()
        probs, value = self._predict(state)
        p = t.from_numpy(np.ascontiguousarray(probs, np.float32).reshape(1, ACTION_SPACE)).to(mb.e.dev)
        v = t.tensor([float(value)], dtype=t.float32, device=mb.e.dev)
        return p, v

    def search(self, game: XiangqiGame, temperature: float = 1.0, add_noise: bool = True) -> np.ndarray:
        eng = engine()
        if self._batch is None:
            self._batch = xq_native.MctsBatch(eng, 1, node_capacity=(self.num_simulations + 2) * 130 + 1024)
        mb = self._batch
        mb.set_games(*game_to_arrays(game))
        noise = None
        n_legal = len(game.get_legal_actions())
        if n_legal == 0:
            return np.zeros(ACTION_SPACE)                        # mcts.py:111-112
        if add_noise:
            # same RNG stream as the reference (mcts.py:118): numpy's global legacy generator
            nz = np.zeros((1, xq_native.MAX_MOVES), np.float64)
            nz[0, :n_legal] = np.random.dirichlet([0.3] * n_legal)
            noise = mb.t.from_numpy(nz).to(eng.dev)
        acts, vis, n = mb.search(self._evaluator, self.num_simulations, self.c_puct, noise=noise, add_noise=add_noise)
        k = int(n[0])
        acts = acts[0, :k].cpu().numpy().astype(np.int64)
        counts = vis[0, :k].cpu().numpy()
        self.last_root = (acts, counts)
        probs = np.zeros(ACTION_SPACE)
        if temperature == 0:                                     # mcts.py:196-199, first max wins
            probs[acts[int(np.argmax(counts))]] = 1.0
            return probs
        probs[acts] = counts
        if probs.sum() > 0:                                      # mcts.py:201-203
            probs = probs ** (1.0 / temperature)
            probs /= probs.sum()
        return probs

    @staticmethod
    def _mask_and_normalize(policy_probs, legal_actions):
        """mcts.py:176-188 (float32 sequential sum, like the device kernel's policy_kind 0)."""
        total = sum(policy_probs[a] for a in legal_actions)
        if total > 0:
            return {a: policy_probs[a] / total for a in legal_actions}
        return {a: 1.0 / len(legal_actions) for a in legal_actions}

    @staticmethod
    def _get_action_probs(root: MCTSNode, temperature: float) -> np.ndarray:
        """mcts.py:190-206 for a host-side MCTSNode root."""
        probs = np.zeros(ACTION_SPACE)
        for a, child in root.children.items():
            probs[a] = child.visit_count
        if temperature == 0:
            best = max(root.children, key=lambda a: root.children[a].visit_count)
            probs = np.zeros(ACTION_SPACE)
            probs[best] = 1.0
        elif probs.sum() > 0:
            probs = probs ** (1.0 / temperature)
            probs /= probs.sum()
        return probs

    def get_action(self, game: XiangqiGame, temperature: float = 0.0, add_noise: bool = False) -> int:
        action_probs = self.search(game, temperature, add_noise)
        if temperature == 0:
            return int(np.argmax(action_probs))
        return int(np.random.choice(len(action_probs), p=action_probs))
