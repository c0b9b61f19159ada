import sys, time, torch
sys.path[:0] = ['xiangqi-alphazero_b200']
import game, model as M
eng = game.engine(0)
C, R, B = 128, 2, 4096
m = M.XiangqiNet(C, R).eval()
net = M.B200Net(eng, m, max_batch=B)
x = torch.zeros((B, 15, 10, 9)); x[:, 14] = 1; x[:, 0, 0, 4] = 1
net.load_planes(x)
for _ in range(3): net.run()
torch.cuda.synchronize()
eng.set_timing(True)
names = ['input'] + [f'res{i//2}.{i%2}' for i in range(2*R)] + ['heads', 'fc']
out = []
for i in range(net.n_layers):
    ts = []
    for _ in range(5):
        net.run_layer(i); torch.cuda.synchronize(); ts.append(eng.last_kernel_ms())
    out.append(f"{names[i]}={min(ts)*1e3:.0f}us")
print(" ".join(out))
