"""Batched evaluation arena: new model vs best model, all games on the GPU.

Replaces AlphaZeroTrainer._serial_evaluate (train.py:453-535), which plays `eval_games` games one after the
other with two MCTS objects.  Here every game is a slot of the device-resident loop (csrc/xq_mcts.cu,
xq_arena_play): a game's leaves go to the network of the player to move at the root, so each of the two forwards
of a step is sized to its own share of the leaves; `leaves_per_game` > 1 (opt-in, virtual loss) widens the batch of
the few evaluation games.  Rules kept from the reference: game i has the new model as red when i is even
(:474), moves are get_action(temperature=0, add_noise=False) with `eval_simulations` simulations (:481-483),
no random opening, no resignation, an undecided game after max_game_length plies is a draw (:496-498),
win_rate = (new_wins + 0.5 draws) / games (:512).

Multi-GPU: games are independent, so ranks take whole (red, black) PAIRS of games (keeps "even index = new
model is red" true locally) and the three counters are summed with one all-reduce.
"""
import ctypes as C

import numpy as np
import torch

import xq_native
from selfplay_engine import SelfPlayEngine

MAX_PLIES = 201


def shard_pairs(num_games: int, rank: int, world: int) -> int:
    """Games of this rank when the (2k, 2k+1) pairs are dealt round-robin; a final odd game goes with its pair slot."""
    pairs = (num_games + 1) // 2
    mine = pairs // world + (1 if rank < pairs % world else 0)
    games = 2 * mine
    if num_games % 2 and mine > 0 and (pairs - 1) % world == rank:
        games -= 1                       # the last pair only has its first (new model = red) game
    return games


class Arena:
    def __init__(self, eng: "xq_native.Engine", model_new, model_old, num_games: int, max_simulations: int,
                 leaves_per_game: int = 1):
        # NOTE: a context holds one self-play state; give the arena its own xq_native.Engine when a SelfPlayEngine on
        # `eng` must stay usable (AlphaZeroTrainer does)
        self.e = eng
        self.num_games = int(num_games)
        self.leaves_per_game = max(1, int(leaves_per_game))
        self.sp = SelfPlayEngine(eng, model_new, n_slots=self.num_games, max_games=self.num_games, sample_capacity=1,
                                 max_simulations=max_simulations, leaves_per_game=self.leaves_per_game)
        self.plan_new, self.net_new = self.sp.plan, self.sp.net
        self.sp.set_model(model_old)                   # second weight set with the same batch geometry
        self.plan_old, self.net_old = self.sp.plan, self.sp.net
        self.move_log = torch.full((self.num_games, MAX_PLIES), -1, dtype=torch.int16, device=eng.dev)

    def play(self, num_simulations: int, c_puct: float, max_game_length: int, chunk: int = 8):
        """-> dict(new_wins, old_wins, draws, winners int8[num_games], plies int16[num_games], moves int16[num_games,201])."""
        sp, e = self.sp, self.e
        sp.reset()                                     # also clears a live-games bound left by an earlier run
        self.move_log.fill_(-1)
        cfg = SelfPlayEngine.make_config(dict(num_simulations=num_simulations, c_puct=c_puct, max_game_length=max_game_length,
                                              random_opening_moves=0, enable_resign=False), self.num_games, seed=0,
                                         add_noise=False, leaves_per_game=self.leaves_per_game)
        played = 0
        limit = max_game_length + 2
        while played < limit:
            step = min(chunk, limit - played)
            e._check(e.L.xq_arena_play(e.h, C.byref(cfg), C.byref(self.plan_new), C.byref(self.plan_old), step,
                                       self.move_log.data_ptr(), e._stream()))
            played += step
            c = sp.counters()
            if c["error"]:
                raise xq_native.XqError(f"arena device error bits {c['error']}")
            if c["finished"] >= self.num_games:
                break
            # every arena game starts at ply 0: the games still alive are the unfinished ones (forwards sized to them)
            e._check(e.L.xq_selfplay_set_live_bound(e.h, max(1, self.num_games - c["finished"])))
        _, winner, plies = sp.fetch(0, 0)
        winner, plies = winner[:self.num_games], plies[:self.num_games]
        new_is_red = (np.arange(self.num_games) % 2) == 0
        decided = (winner == 1) | (winner == -1)
        new_won = decided & ((winner == 1) == new_is_red)
        return dict(new_wins=int(new_won.sum()), old_wins=int((decided & ~new_won).sum()),
                    draws=int((~decided).sum()), winners=winner.copy(), plies=plies.copy(),
                    moves=self.move_log.cpu().numpy())


def evaluate_models(eng, model_new, model_old, eval_games: int, eval_simulations: int, c_puct: float,
                    max_game_length: int, dist=None, leaves_per_game: int = 1):
    """The numbers of train.py:512-520: new_wins, old_wins, draws, win_rate (all ranks return the same dict).
    `eng`: the context the arena may take over (its previous SelfPlayEngine, if any, is superseded)."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    mine = shard_pairs(eval_games, rank, world)
    counts = torch.zeros(3, dtype=torch.int64, device=eng.dev)
    if mine > 0:
        r = Arena(eng, model_new, model_old, mine, eval_simulations, leaves_per_game).play(eval_simulations, c_puct, max_game_length)
        counts += torch.tensor([r["new_wins"], r["old_wins"], r["draws"]], dtype=torch.int64, device=eng.dev)
    if world > 1:
        dist.all_reduce(counts)
    nw, ow, dr = (int(x) for x in counts.tolist())
    return dict(new_wins=nw, old_wins=ow, draws=dr, win_rate=(nw + 0.5 * dr) / max(eval_games, 1))
