"""f4: the per-object path a demo / benchmark.py-style caller uses -- BASELINE configs[0]'s shape: ONE
game, XiangqiNet(128,6), 200 simulations per move -- through the drop-in classes: `model.predict()` latency, one
`MCTS(model, 200).search(game)` (one predict() per simulation, like the reference), and the same search in the lockstep
engine with a single slot (what parallel_self_play would do for one game).  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "xiangqi-alphazero_b200"))
import numpy as np
import torch
import game as G
import model as M
from mcts import MCTS
from selfplay_engine import SelfPlayEngine

torch.manual_seed(1)
net = M.XiangqiNet(128, 6).eval()
g = G.XiangqiGame()
state = g.get_state_for_nn()
for _ in range(5):
    net.predict(state)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    net.predict(state)
predict_ms = (time.perf_counter() - t0) / 200 * 1e3

m = MCTS(net, num_simulations=200, c_puct=1.5)
m.search(g, temperature=1.0, add_noise=True)
t0 = time.perf_counter()
for _ in range(3):
    m.search(g, temperature=1.0, add_noise=True)
search_s = (time.perf_counter() - t0) / 3

eng = G.engine(0)


class Cfg:
    num_simulations, c_puct, temperature_threshold, max_game_length = 200, 1.5, 30, 300
    random_opening_moves, enable_resign, resign_threshold, resign_check_steps = 0, False, -0.9, 5


sp = SelfPlayEngine(eng, net, n_slots=1, max_games=1, max_simulations=200)
sp.reset()
cfg = SelfPlayEngine.make_config(Cfg(), 1, seed=3, add_noise=True)
sp.play(cfg, 2)
torch.cuda.synchronize()
c0 = sp.counters()
t0 = time.perf_counter()
sp.play(cfg, 10)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
c1 = sp.counters()
print(json.dumps({"workload": "configs[0] shape: one game, XiangqiNet(128,6), 200 sims/move (per-object drop-in path, batch of one)",
                  "predict_ms": predict_ms, "mcts_search_200_sims_s": search_s, "mcts_sims_per_s": 200 / search_s,
                  "lockstep_engine_one_slot_sims_per_s": (c1["sims"] - c0["sims"]) / dt,
                  "lockstep_engine_ms_per_move": dt / 10 * 1e3}))
