"""Fixed cost vs per-step cost of the host-dataset loop (rcn_cuda_train_epoch_host): wall-clock per call for
n = 2 .. 2000 steps of the c2 workload over a pinned host dataset; a + b*n fit. One JSON line."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mercer_research_b200 import RCN  # noqa: E402

wl = bench.WORKLOADS["c2"]
B, H, W = wl["batch"], wl["H"], wl["W"]
L = bench.layer_shapes(wl)[0][1]
model = RCN(wl["classes"], wl["cfg"], wl["ff"])
model.load_weights_and_bias(L)
model.set_params(np.random.default_rng(1).standard_normal(model.n_params) * 0.05)
n_max = 2000
h_images = torch.randint(0, 256, (n_max * B, H, W), dtype=torch.uint8).pin_memory()
h_labels = (torch.arange(n_max * B) % wl["classes"]).to(torch.int64).pin_memory()
hi, hl = h_images.numpy(), h_labels.numpy()
model.gen_scales(model.flatten_feature_set(hi[:B]))
model.set_stream(torch.cuda.current_stream().cuda_stream)
rows = []
for n in (2, 8, 20, 64, 200, 2000):
    model.train_epoch_host(hi[:n * B], hl[:n * B], B, 3.0)          # warm-up at this size
    torch.cuda.synchronize()
    ts = []
    for _ in range(15 if n <= 200 else 5):
        t0 = time.perf_counter()
        model.train_epoch_host(hi[:n * B], hl[:n * B], B, 3.0)
        ts.append(time.perf_counter() - t0)
    rows.append((n, float(np.median(ts)) * 1e6, float(np.min(ts)) * 1e6))
ns = np.array([r[0] for r in rows], dtype=np.float64)
us = np.array([r[1] for r in rows])
b, a = np.polyfit(ns, us, 1)
print(json.dumps({"workload": "c2", "unroll": os.environ.get("RCN_CUDA_PREFETCH_UNROLL", "default"),
                  "ctas": os.environ.get("RCN_CUDA_PREFETCH_CTAS", "default"),
                  "host_steps_per_graph": os.environ.get("RCN_CUDA_HOST_STEPS_PER_GRAPH", "default"),
                  "calls": [{"steps": n, "median_us": round(m, 1), "min_us": round(lo, 1), "us_per_step": round(m / n, 2)}
                            for n, m, lo in rows],
                  "fit": {"fixed_us": round(float(a), 1), "us_per_step": round(float(b), 3)}}))
