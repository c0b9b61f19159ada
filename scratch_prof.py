import sys, torch
sys.path[:0] = ['xiangqi-alphazero_b200']
import game, model as M
eng = game.engine(0)
m = M.XiangqiNet(128, 2).eval()
net = M.B200Net(eng, m, max_batch=4096)
x = torch.zeros((4096, 15, 10, 9)); x[:, 14] = 1; x[:, 0, 0, 4] = 1
net.load_planes(x)
for _ in range(3): net.run()
torch.cuda.synchronize()
