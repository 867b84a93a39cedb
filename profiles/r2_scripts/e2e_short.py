"""Short streamed epochs (the driver's K = 20): host wall time per rcn_cuda_train_epoch_host call against the device span
from kernel A of step 0 to the end of kernel B of the last step (in-graph timeline), per copy mode."""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import mercer_research_b200 as m

B, steps = 1024, int(os.environ.get("STEPS", "20"))
rng = np.random.default_rng(5)
images = torch.from_numpy(rng.integers(0, 256, size=(steps * B, 28, 28), dtype=np.uint8)).pin_memory()
labels = torch.from_numpy(rng.integers(0, 10, size=steps * B).astype(np.int64)).pin_memory()
model = m.RCN(10, [m.RCNLayer.Convolve2D(m.Padding.Same), m.RCNLayer.Pool2D(m.Pooling.Max)], [30])
model.load_weights_and_bias(784)
model.set_params(np.random.default_rng(6).standard_normal(model.n_params) * 0.05)
model.scale_set = (40.0, 60.0)
tl = os.environ.get("TL", "1") == "1"
if tl:
    model.timeline_enable(True)
hi, hl = images.numpy(), labels.numpy()
t_end = time.perf_counter() + 0.4
while time.perf_counter() < t_end:
    model.train_epoch_host(hi, hl, B, 0.1)
walls, spans, firsts = [], [], []
for _ in range(30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    model.train_epoch_host(hi, hl, B, 0.1)
    walls.append((time.perf_counter() - t0) * 1e6)
    if tl:
        s, e, n = model.timeline_read()
        na = int(n[0])
        idx = [(na - steps + i) % 64 for i in range(steps)]
        a0 = s[0][idx].astype(np.int64); b1 = e[1][idx].astype(np.int64)
        spans.append((b1[-1] - a0[0]) / 1e3)
        firsts.append(np.diff(a0)[:19] / 1e3)
walls = np.array(walls)
print(os.environ.get("RCN_CUDA_HOST_COPY", "dma"), "steps", steps, "TL", int(tl), "wall us/call median", round(float(np.median(walls)), 1),
      "min", round(float(walls.min()), 1), "-> M img/s", round(B * steps / np.median(walls), 2))
if tl:
    print("  device span A(0)..B(last) median", round(float(np.median(spans)), 1), "us; step periods (median over calls):",
          np.median(np.array(firsts), axis=0).round(1).tolist())
