#!/bin/bash
# conv-only strip kernels: feature parity tests + the c4 leg with and without them
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_features.py -m gpu -q -k "layered or features_u8" 2>&1 | tail -2
for v in 1 0; do
RCN_CUDA_CONV_STRIPS=$v timeout 300 python bench.py --steps 20 --warmup 5 --extra c4 > $OUT/r2t_c4_$v.json 2> $OUT/r2t_c4_$v.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2t_c4_$v.json"))
w = d["workloads"]["c4"]
print("strips=$v", round(w["value"] / 1e6, 3), "M img/s", round(w["ms_per_step"], 4), "ms  frac_hbm", round(w["roofline"]["frac"], 3), {k: round(x["avg_us"], 1) for k, x in w["roofline"]["kernels"].items()})
PY
done
