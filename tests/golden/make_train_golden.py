"""Generate tests/golden/train_golden.npz by RUNNING THE REFERENCE's training step (build container only).

    python tests/golden/make_train_golden.py
    python tests/golden/make_train_golden.py 128 2 train_golden_128.npz

Imports /root/reference/training/train.py unmodified (with its game.py / model.py / parallel_selfplay.py and the
Cython engine from oracle/_ref), builds a replay buffer of (sample, mirrored sample) pairs with the reference's own
`_augment_data`, seeds torch and calls `AlphaZeroTrainer.train_network()` on the CPU in fp32.  The fixture holds the
sparse form of the samples, the per-call loss statistics and float64 checksums of every parameter before and after
training; tests/test_train_gpu.py feeds the same samples to the B200 trainer and compares.
"""
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/training"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

import train as reftrain                      # noqa: E402
import parallel_selfplay as refsp             # noqa: E402
from game import XiangqiGame                  # noqa: E402

assert reftrain.__file__.startswith(REF)

CHANNELS, BLOCKS, RECORDS, BATCH, EPOCHS, SEED = 16, 1, 150, 64, 2, 20261018
OUT = "train_golden.npz"
if len(sys.argv) > 1:        # python make_train_golden.py 128 2 train_golden_128.npz: the fixture of the hand-written step (128-wide tower)
    CHANNELS, BLOCKS, OUT = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]


def checksums(model):
    return np.array([[float(p.detach().double().sum()), float(p.detach().double().abs().sum())] for p in model.parameters()])


def main():
    random.seed(SEED)
    rs = np.random.RandomState(SEED)
    boards, sides, ns, acts, probs, zs, dense = [], [], [], [], [], [], []
    while len(boards) < RECORDS:
        g = XiangqiGame()
        for _ in range(rs.randint(0, 120)):
            done, _ = g.is_game_over()
            if done:
                break
            g.make_move(*random.choice(g.get_legal_moves()))
        done, _ = g.is_game_over()
        legal = g.get_legal_actions()
        if done or not legal:
            continue
        p32 = rs.dirichlet([0.5] * len(legal)).astype(np.float32)
        p32[rs.rand(len(legal)) < 0.3] = 0.0                  # unvisited moves
        if p32.sum() <= 0:
            p32[0] = 1.0
        p32 = (p32 / p32.sum()).astype(np.float32)
        z = float(rs.choice([-1.0, 0.0, 1.0]))
        pol = np.zeros(8100, np.float64)
        pol[legal] = p32
        dense.append((g.get_state_for_nn(), pol, z))
        a = np.full(128, -1, np.int16)
        a[:len(legal)] = legal
        q = np.zeros(128, np.float32)
        q[:len(legal)] = p32
        boards.append(g.board.reshape(90).copy())
        sides.append(g.current_player)
        ns.append(len(legal))
        acts.append(a)
        probs.append(q)
        zs.append(z)
    pairs = refsp._augment_data(dense)                         # the reference's own mirror augmentation
    assert len(pairs) == 2 * RECORDS

    cfg = reftrain.TrainingConfig()
    cfg.num_channels, cfg.num_res_blocks = CHANNELS, BLOCKS
    cfg.batch_size, cfg.num_epochs, cfg.min_buffer_size = BATCH, EPOCHS, 10
    cfg.device = 'cpu'
    cfg.checkpoint_dir = '/tmp/xq_train_golden'
    torch.manual_seed(SEED)
    tr = reftrain.AlphaZeroTrainer(cfg)
    init = checksums(tr.current_model)
    tr.replay_buffer.extend(pairs)
    torch.manual_seed(SEED + 1)
    stats1 = tr.train_network()
    after1 = checksums(tr.current_model)
    stats2 = tr.train_network()                               # second call: Adam moments and the LR schedule carry over
    after2 = checksums(tr.current_model)
    bn = np.array([[float(b.detach().double().sum())] for b in tr.current_model.buffers()])

    # (b) the reference's loss on ONE batch holding the whole buffer, initial weights, learning rate 0 (no update):
    # pins the CPU oracle's planes + mirroring + loss against train_network itself
    cfg0 = reftrain.TrainingConfig()
    cfg0.num_channels, cfg0.num_res_blocks = CHANNELS, BLOCKS
    cfg0.batch_size, cfg0.num_epochs, cfg0.min_buffer_size, cfg0.learning_rate = 2 * RECORDS, 1, 10, 0.0
    cfg0.device = 'cpu'
    cfg0.checkpoint_dir = '/tmp/xq_train_golden'
    torch.manual_seed(SEED)
    tr0 = reftrain.AlphaZeroTrainer(cfg0)
    tr0.replay_buffer.extend(pairs)
    stats0 = tr0.train_network()
    # (c) `_augment_data` outputs of the first 24 samples (mirrored twin: packed planes, policy as sorted (index, value))
    K = 24
    mir_planes = np.array([np.packbits(pairs[2 * i + 1][0].reshape(-1) > 0.5) for i in range(K)])
    mir_idx = np.full((K, 128), -1, np.int32)
    mir_val = np.zeros((K, 128), np.float32)
    for i in range(K):
        nz = np.nonzero(pairs[2 * i + 1][1] > 0)[0]
        mir_idx[i, :len(nz)] = nz
        mir_val[i, :len(nz)] = pairs[2 * i + 1][1][nz]
    np.savez_compressed(
        os.path.join(HERE, OUT),
        board=np.array(boards, np.int8), side=np.array(sides, np.int8), n=np.array(ns, np.uint8), actions=np.array(acts),
        probs=np.array(probs), z=np.array(zs, np.float32), init=init, after1=after1, after2=after2, buffers=bn,
        stats1=np.array([stats1['policy_loss'], stats1['value_loss'], stats1['total_loss'], stats1['learning_rate']]),
        stats2=np.array([stats2['policy_loss'], stats2['value_loss'], stats2['total_loss'], stats2['learning_rate']]),
        full_batch=np.array([stats0['policy_loss'], stats0['value_loss']]), mirror_planes=mir_planes, mirror_index=mir_idx,
        mirror_value=mir_val,
        meta=np.array([CHANNELS, BLOCKS, RECORDS, BATCH, EPOCHS, SEED]))
    print("stats1", stats1)
    print("stats2", stats2)


if __name__ == "__main__":
    main()
