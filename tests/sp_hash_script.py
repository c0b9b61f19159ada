"""Launched by tests/test_selfplay_gpu.py: a small seeded self-play (and arena) run, prints a digest of the games.
The switches read once per process (XQ_SP_GRAPH, XQ_NET_PDL, XQ_NET_SMALL) are set by the caller's environment."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
import torch
import game
import model as M
from arena import Arena
from selfplay_engine import SelfPlayEngine, decode_samples


class Cfg:
    num_simulations = 20
    c_puct = 1.5
    temperature_threshold = 6
    max_game_length = 30
    random_opening_moves = 3
    enable_resign = True
    resign_threshold = -0.2
    resign_check_steps = 2
    num_games_per_iter = 60


eng = game.engine(0)
torch.manual_seed(3)
net = M.XiangqiNet(128, 2).eval()
sp = SelfPlayEngine(eng, net, n_slots=48, max_games=60)
sp.reset()
c = sp.play_games(SelfPlayEngine.make_config(Cfg(), 60, seed=11), chunk=4)
raw, winner, plies = sp.fetch(0, c["samples"])
dec = decode_samples(raw)
order = np.lexsort((dec["ply"], dec["uid"]))
h = hashlib.sha256()
for k in ("board", "side", "n", "actions", "probs", "played", "uid", "ply"):
    h.update(np.ascontiguousarray(dec[k][order]).tobytes())
h.update(winner[:60].tobytes())
h.update(plies[:60].tobytes())
torch.manual_seed(4)
old = M.XiangqiNet(128, 2).eval()
r = Arena(eng, net, old, 10, 16).play(16, 1.5, 24)
h.update(np.ascontiguousarray(r["moves"]).tobytes())
print("SP_DIGEST", c["samples"], c["finished"], c["error"], h.hexdigest())
