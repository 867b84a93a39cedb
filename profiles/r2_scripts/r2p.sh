#!/bin/bash
# single-GPU prewait (kernel A's front end under the tail of the previous kernel B): parity tests, then A/B of the bench line
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_configs.py -m gpu -q -x -k "c2 or c3 or epoch or host or generation or graph or smallnet or fused" 2>&1 | tail -3
for pw in 1 0 1 0; do
RCN_CUDA_PREWAIT=$pw timeout 300 python bench.py --steps 20 --warmup 5 --no-extra > $OUT/r2p_${pw}.json 2> $OUT/r2p_${pw}.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2p_${pw}.json"))
tl = d["roofline"].get("in_graph_timeline") or {}
print("prewait=$pw K=20", round(d["value"] / 1e6, 2), "M", round(d["ms_per_step"] * 1e3, 2), "us  e2e", round(d["e2e"]["value"] / 1e6, 2), "M |", {k: (round(v["us_mean"], 2) if isinstance(v, dict) else v) for k, v in tl.items()})
PY
done
for pw in 1 0; do
RCN_CUDA_PREWAIT=$pw timeout 300 python bench.py --steps 2000 --warmup 20 --no-extra > $OUT/r2p_${pw}_long.json 2> $OUT/r2p_${pw}_long.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2p_${pw}_long.json"))
print("prewait=$pw K=2000", round(d["value"] / 1e6, 2), "M", round(d["ms_per_step"] * 1e3, 2), "us  e2e", round(d["e2e"]["value"] / 1e6, 2), "M")
PY
done
