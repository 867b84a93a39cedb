"""A short host-dataset epoch (rcn_cuda_train_epoch_host, c2 workload, pinned u8 images) -- the command profiled under
ncu for the streaming path's kernels (host_prefetch_kernel + the two step kernels)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mercer_research_b200 import RCN  # noqa: E402

wl = bench.WORKLOADS["c2"]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
B, H, W = wl["batch"], wl["H"], wl["W"]
model = RCN(wl["classes"], wl["cfg"], wl["ff"])
model.load_weights_and_bias(bench.layer_shapes(wl)[0][1])
model.set_params(np.random.default_rng(1).standard_normal(model.n_params) * 0.05)
h_images = torch.randint(0, 256, (steps * B, H, W), dtype=torch.uint8).pin_memory()
h_labels = (torch.arange(steps * B) % wl["classes"]).to(torch.int64).pin_memory()
model.gen_scales(model.flatten_feature_set(h_images.numpy()[:B]))
for _ in range(2):
    cost, hits = model.train_epoch_host(h_images.numpy(), h_labels.numpy(), B, 3.0)
print("ok", len(cost), float(cost[-1]), int(hits[-1]))
