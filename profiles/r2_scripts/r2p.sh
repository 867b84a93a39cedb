#!/bin/bash
# single-GPU prewait modes (RCN_CUDA_PREWAIT: 0 nowhere, 1 streamed host epoch only, 2 every step): parity tests, bench line per mode
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q -x -k "c2 or epoch or host or generation or graph" 2>&1 | tail -2
for pw in 1 0 2; do
RCN_CUDA_PREWAIT=$pw timeout 300 python bench.py --steps 2000 --warmup 20 --no-extra > $OUT/r2p_${pw}_long.json 2> $OUT/r2p_${pw}_long.err
RCN_CUDA_PREWAIT=$pw timeout 300 python bench.py --steps 20 --warmup 5 --no-extra > $OUT/r2p_${pw}.json 2> $OUT/r2p_${pw}.err
python - <<PY
import json
for f, k in (("r2p_${pw}_long.json", 2000), ("r2p_${pw}.json", 20)):
    d = json.load(open("gpurun_out/" + f))
    tl = d["roofline"].get("in_graph_timeline") or {}
    print("prewait=$pw K=%d" % k, round(d["value"] / 1e6, 2), "M", round(d["ms_per_step"] * 1e3, 2), "us  e2e", round(d["e2e"]["value"] / 1e6, 2), "M |", {k2[:8]: (round(v["us_mean"], 2) if isinstance(v, dict) else round(v, 2)) for k2, v in tl.items()})
PY
done
