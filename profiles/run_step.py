"""Runs a few eager training steps of one workload (default c2) -- the command profiled under ncu."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mercer_research_b200 import RCN  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
B, H, W = wl["batch"], wl["H"], wl["W"]
dev = torch.device("cuda", 0)
L = bench.layer_shapes(wl)[0][1]
model = RCN(wl["classes"], wl["cfg"], wl["ff"])
model.load_weights_and_bias(L)
model.set_params(np.random.default_rng(1).standard_normal(model.n_params))
images = torch.randint(0, 256, (4, B, H, W), dtype=torch.uint8, device=dev)
labels = (torch.arange(B, device=dev) % wl["classes"]).to(torch.int64)
model.gen_scales(model.flatten_feature_set(images[0]))
for i in range(steps):
    model.train_batch_images(images[i % 4], labels, 3.0)
torch.cuda.synchronize()
print("ok", model.last_batch_stats())
