#!/bin/bash
# c3 with the big-tile backward-weight GEMM vs without; dense parity tests
set -u
OUT=gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python bench.py --workload c3 --steps 320 --no-extra > $OUT/r2r_c3_$tag.json 2> $OUT/r2r_c3_$tag.err; python - <<PY
import json
d = json.load(open("gpurun_out/r2r_c3_$tag.json"))
k = d["roofline"]["kernels"]
print("$tag", round(d["value"] / 1e6, 3), "M img/s", round(d["ms_per_step"] * 1e3, 1), "us", {n: round(v["avg_us"], 1) for n, v in k.items() if "gemm" in n or "reduce" in n})
PY
}
run big A=1
run nobig RCN_CUDA_WGRAD_BIG_TILE=0
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_configs.py tests/test_gpu_ext.py -m gpu -q -k "parity or c3 or gemm or conv or dmma" 2>&1 | tail -3
