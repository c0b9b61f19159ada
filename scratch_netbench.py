import sys, time, torch
sys.path[:0] = ['xiangqi-alphazero_b200']
import game, model as M
eng = game.engine(0)
for (C, R, B) in [(128, 6, 4096), (256, 20, 4096)]:
    m = M.XiangqiNet(C, R).eval()
    net = M.B200Net(eng, m, max_batch=B)
    x = torch.zeros((B, 15, 10, 9))
    x[:, 14] = 1
    x[:, 0, 0, 4] = 1
    net.load_planes(x)
    for _ in range(3): net.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n): net.run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = net.flops_per_board() * B
    print(f"C={C} R={R} B={B}: {ms:.3f} ms/forward, {fl/ms/1e9:.1f} TFLOP/s algorithmic, {B/ms*1e3:.0f} evals/s")
    eng.set_timing(True)
    names = ['input'] + [f'res{i//2}.{i%2}' for i in range(2*R)] + ['heads', 'fc']
    tot = 0
    out = []
    for i in range(net.n_layers):
        ts = []
        for _ in range(3):
            net.run_layer(i); torch.cuda.synchronize(); ts.append(eng.last_kernel_ms())
        t = min(ts); tot += t
        if i < 3 or i >= net.n_layers - 2: out.append(f"{names[i]}={t*1e3:.0f}us")
    print("   ", " ".join(out), f"sum={tot:.3f} ms")
    eng.set_timing(False)
    del net
