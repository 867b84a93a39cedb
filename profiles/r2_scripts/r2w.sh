#!/bin/bash
# host-dataset loop: copy engine vs SM pull; parity tests, then the driver's bench command per variant (c2 only)
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q -x -k "host or generation or epoch or c2" 2>&1 | tail -4
for v in "dma 0" "dma 40" "pull 0"; do
set -- $v
for rep in 1 2; do
RCN_CUDA_HOST_COPY=$1 RCN_CUDA_HOST_STEPS_PER_GRAPH=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-extra > $OUT/r2w_$1_$2_$rep.json 2> $OUT/r2w_$1_$2_$rep.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2w_$1_$2_$rep.json"))
print("$1 spg=$2 K=20", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), "M", round(d["e2e"]["ms_per_step"] * 1e3, 2), "us/step", round(d["e2e"]["h2d_GBps_per_gpu"], 1), "GB/s")
PY
done
done
for v in "dma 0" "dma 40" "pull 0"; do
set -- $v
RCN_CUDA_HOST_COPY=$1 RCN_CUDA_HOST_STEPS_PER_GRAPH=$2 timeout 300 python bench.py --steps 2000 --warmup 20 --no-extra > $OUT/r2w_$1_$2_long.json 2> $OUT/r2w_$1_$2_long.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2w_$1_$2_long.json"))
print("$1 spg=$2 K=2000", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), "M", round(d["e2e"]["ms_per_step"] * 1e3, 2), "us/step", round(d["e2e"]["h2d_GBps_per_gpu"], 1), "GB/s")
PY
done
