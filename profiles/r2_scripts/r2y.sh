#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -q -x -k "host" 2>&1 | tail -2
RCN_CUDA_HOST_COPY=dma TL=0 timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -1
for rep in 1 2 3 4 5; do
for q in 1 0; do
RCN_BENCH_CLOCK_QUIET=$q timeout 300 python bench.py --steps 20 --warmup 5 --no-extra > $OUT/r2y_${q}_${rep}.json 2> $OUT/r2y_${q}_${rep}.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2y_${q}_${rep}.json"))
print("quiet=$q K=20", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), "M", round(d["e2e"]["ms_per_step"] * 1e3, 2), "us/step |", d["clocks"]["samples"], d["clocks"]["note"][-60:])
PY
done
done
