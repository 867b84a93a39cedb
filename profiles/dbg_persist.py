import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mercer_research_b200 import RCN, Padding, Pooling, RCNLayer
def fresh(imgs, labels, B):
    m = RCN(10, [RCNLayer.Convolve2D(Padding.Same), RCNLayer.Pool2D(Pooling.Max)], [30])
    m.scale_set = (20.0, 35.0)
    m.load_weights_and_bias(784)
    m.set_params(np.random.default_rng(1).standard_normal(m.n_params) * 0.05)
    m.epoch_bind(imgs, labels, B)
    return m
B = 1024
imgs = torch.randint(0, 256, (8192, 28, 28), dtype=torch.uint8, device="cuda")
labels = (torch.arange(8192, device="cuda") % 10).to(torch.int64)
for n in (1, 2, 3):
    a = fresh(imgs, labels, B); b = fresh(imgs, labels, B)
    for _ in range(n): a.epoch_step(3.0)
    b.epoch_run(3.0, n)
    torch.cuda.synchronize()
    ga, gb = a.get_gradients(), b.get_gradients()
    pa, pb = a.get_params(), b.get_params()
    dg = np.abs(ga - gb); dp = np.abs(pa - pb)
    print(f"steps {n}: grads max diff {dg.max():.3e} (W0 {dg[:23520].max():.3e}, small {dg[23520:].max():.3e}); params max diff {dp.max():.3e} (W0 {dp[:23520].max():.3e}, small {dp[23520:].max():.3e}); pos {a.epoch_position()} {b.epoch_position()}")
    if n == 1:
        bad = np.nonzero(dg[:23520] > 1e-9)[0]
        print("bad W0 grads:", len(bad), "cols", np.unique(bad // 30)[:40], "rows", np.unique(bad % 30)[:40])
        print("stats", a.last_batch_stats(), b.last_batch_stats())
