"""Feature-stage throughput (rcn_cuda_features: flatten_feature_set + standardise, rcn.rs:317-356,407-412) against
the HBM roofline.  Algorithmic bytes per image = H*W (u8 read) + L*8 (f64 features written), SURVEY.md 8d.
Inputs are device resident and larger than L2; CUDA events on the launching stream; JSON lines on stdout.
Usage: python profiles/features_bench.py [case-prefix]            (RCN_CUDA_FEATURES_STAGED=0 selects the one-image-per-CTA kernel)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mercer_research_b200 import RCN, Padding, Pooling, RCNLayer  # noqa: E402

CP = [RCNLayer.Convolve2D(Padding.Same), RCNLayer.Pool2D(Pooling.Max)]
CASES = [("c2 28x28 [C,P]", 28, 28, CP, 131072), ("c3 32x32 [C,P]x3", 32, 32, CP * 3, 65536),
         ("c5 64x64 [C,P]", 64, 64, CP, 16384), ("28x28 [C,P]x2", 28, 28, CP * 2, 131072),
         # BASELINE config 4's conv stack: four Same convolutions, 256 maps of 64x64 per image (the layered path's strip kernels)
         ("c4 64x64 [C]x4", 64, 64, [RCNLayer.Convolve2D(Padding.Same)] * 4, 128)]


def main():
    peak = 6549.4
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    dev = torch.device("cuda", 0)
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, H, W, cfg, B in CASES:
        if only and not name.startswith(only):
            continue
        model = RCN(10, cfg, [30])
        L = model.feature_len(H, W)
        imgs = torch.randint(0, 256, (B, H, W), dtype=torch.uint8, device=dev)
        out = torch.empty((B, L), dtype=torch.float64, device=dev)
        model.scale_set = (100.0, 50.0)
        for std in (False, True):
            for _ in range(3):
                model.flatten_feature_set(imgs, standardise=std, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                model.flatten_feature_set(imgs, standardise=std, out=out)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            bytes_ = B * (H * W + L * 8)
            gbs = bytes_ / us / 1e3
            print(json.dumps({"case": name, "batch": B, "standardise": std, "us": round(us, 2), "images_per_s": round(B / us * 1e6),
                              "algorithmic_GBps": round(gbs, 1), "frac_hbm": round(gbs / peak, 4), "peak_GBps": peak,
                              "staged": os.environ.get("RCN_CUDA_FEATURES_STAGED", "1") != "0"}), flush=True)


if __name__ == "__main__":
    main()
