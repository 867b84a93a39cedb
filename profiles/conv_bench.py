"""Wide conv stack (BASELINE config 4: 64-256 channels, 3x3 kernels, 64x64 inputs): forward / backward-data / backward-weight
of one layer through the extension ops on device tensors, tcgen05 integer-slice path vs the f64 DMMA implicit GEMM
(RCN_CUDA_GEMM=dmma in a second process). usage: python profiles/conv_bench.py [B Ci Co]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mercer_research_b200 import _lib, ext  # noqa: E402

B, Ci, Co = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 64, 64)
H = W = 64
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randn(B, H, W, Ci, dtype=torch.float64, device="cuda", generator=g).clamp_(min=0)
w = torch.randn(Co, 3, 3, Ci, dtype=torch.float64, device="cuda", generator=g) / (3.0 * Ci ** 0.5)
b = torch.randn(Co, dtype=torch.float64, device="cuda", generator=g)
dz = torch.randn(B, H, W, Co, dtype=torch.float64, device="cuda", generator=g)
flops = 2.0 * B * H * W * 9 * Ci * Co


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {"B": B, "Ci": Ci, "Co": Co, "H": H, "W": W, "gemm_env": os.environ.get("RCN_CUDA_GEMM", "auto"), "gflop_per_pass": flops / 1e9}
for name, fn in [("forward", lambda: ext.conv2d_forward(x, w, b, 1, 1)), ("backward_data", lambda: ext.conv2d_backward_data(dz, w, (H, W), 1)),
                 ("backward_weight", lambda: ext.conv2d_backward_weight(x, dz, 3, 3, 1))]:
    t = timed(fn)
    out[name + "_ms"] = round(t, 4)
    out[name + "_f64_equiv_tflops"] = round(flops / t / 1e9, 2)
_lib.profile_enable(True)
ext.conv2d_forward(x, w, b, 1, 1); ext.conv2d_backward_data(dz, w, (H, W), 1); ext.conv2d_backward_weight(x, dz, 3, 3, 1)
torch.cuda.synchronize()
out["kernels_ms"] = {k: round(v["total_ms"] / v["launches"], 4) for k, v in _lib.profile_report().items()}
_lib.profile_enable(False)
print(json.dumps(out))
