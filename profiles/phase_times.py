"""Warm-L2 phase timings of one training step (config c2 by default) through the C ABI.

Each phase is captured into a CUDA graph that repeats it R times, replayed, and timed with CUDA events on the
launching stream; reported per repetition. Phases: features (rcn_cuda_features), backprop-from-features
(rcn_cuda_accumulate_gradients: kernels A+B+C or the generic GEMM chain), update (rcn_cuda_apply_gradients), and
the whole epoch step. Usage: python profiles/phase_times.py [workload]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mercer_research_b200 import RCN, _lib  # noqa: E402


def timed_graph(fn, reps=20, replays=20):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * replays) * 1e3  # us


def main():
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
    B, H, W = wl["batch"], wl["H"], wl["W"]
    dev = torch.device("cuda", 0)
    shapes = bench.layer_shapes(wl)
    L = shapes[0][1]
    model = RCN(wl["classes"], wl["cfg"], wl["ff"])
    model.load_weights_and_bias(L)
    model.set_params(np.random.default_rng(1).standard_normal(model.n_params))
    images = torch.randint(0, 256, (B, H, W), dtype=torch.uint8, device=dev)
    labels = (torch.arange(B, device=dev) % wl["classes"]).to(torch.int64)
    model.gen_scales(model.flatten_feature_set(images))
    feats = torch.empty((B, L), dtype=torch.float64, device=dev)

    def use_stream():
        model.set_stream(torch.cuda.current_stream().cuda_stream)

    def f_features():
        use_stream()
        model.flatten_feature_set(images, standardise=True, out=feats)

    def f_backprop():
        use_stream()
        model.accumulate_gradients(feats, labels=labels)

    def f_update():
        use_stream()
        model.apply_gradients(1e-9, B)

    def f_step():
        use_stream()
        model.accumulate_gradients_images(images, labels)
        model.apply_gradients(1e-9, B)

    out = {"workload": wl["desc"]}
    for name, fn in [("features_us", f_features), ("backprop_from_features_us", f_backprop), ("update_us", f_update),
                     ("whole_step_us", f_step)]:
        out[name] = round(timed_graph(fn), 3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
