"""In-graph device timeline of the streamed host epoch (rcn_cuda_train_epoch_host): kernel A / kernel B spans and the
gaps between them for the last 64 steps of an epoch, per copy mode (RCN_CUDA_HOST_COPY) -- run once per mode."""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mercer_research_b200 as m

B, steps = 1024, 600
rng = np.random.default_rng(5)
images = torch.from_numpy(rng.integers(0, 256, size=(steps * B, 28, 28), dtype=np.uint8)).pin_memory()
labels = torch.from_numpy(rng.integers(0, 10, size=steps * B).astype(np.int64)).pin_memory()
model = m.RCN(10, [m.RCNLayer.Convolve2D(m.Padding.Same), m.RCNLayer.Pool2D(m.Pooling.Max)], [30])
model.load_weights_and_bias(784)
model.set_params(np.random.default_rng(6).standard_normal(model.n_params) * 0.05)
model.scale_set = (40.0, 60.0)
if os.environ.get("TL", "1") == "1":
    model.timeline_enable(True)
hi, hl = images.numpy(), labels.numpy()
for _ in range(3):
    model.train_epoch_host(hi, hl, B, 0.1)
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    model.train_epoch_host(hi, hl, B, 0.1)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
print(os.environ.get("RCN_CUDA_HOST_COPY", "dma"), "spg", os.environ.get("RCN_CUDA_HOST_STEPS_PER_GRAPH", "2"),
      "wall us/step", round(dt / steps * 1e6, 2), "M img/s", round(B * steps / dt / 1e6, 2))
if os.environ.get("TL", "1") == "1":
    s, e, n = model.timeline_read()
    na = int(n[0])
    idx = [(na - 60 + i) % 64 for i in range(58)]          # consecutive launches near the end of the last epoch
    a0 = s[0][idx].astype(np.int64); a1 = e[0][idx].astype(np.int64)
    b0 = s[1][idx].astype(np.int64); b1 = e[1][idx].astype(np.int64)
    print("  A span us mean/min/max", round((a1 - a0).mean() / 1e3, 2), (a1 - a0).min() / 1e3, (a1 - a0).max() / 1e3)
    print("  B span us mean", round((b1 - b0).mean() / 1e3, 2), " gap A->B", round((b0 - a1).mean() / 1e3, 2),
          " gap B->A(next)", round((a0[1:] - b1[:-1]).mean() / 1e3, 2), "max", (a0[1:] - b1[:-1]).max() / 1e3,
          " step period", round(np.diff(a0).mean() / 1e3, 2))
    print("  A spans:", ((a1 - a0) / 1e3).round(1)[:24].tolist())
    print("  periods:", (np.diff(a0) / 1e3).round(1)[:24].tolist())
