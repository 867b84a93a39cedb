#!/bin/bash
set -u
OUT=gpurun_out
nvidia-smi topo -m 2>&1 | head -8
nproc
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -q -x -k "host or generation" 2>&1 | tail -2
for pub in 0 1 2; do
echo "publish mode $pub"
RCN_CUDA_HS_PUBLISH=$pub RCN_CUDA_HOST_COPY=dma timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -2 | cut -c1-300
RCN_CUDA_HS_PUBLISH=$pub RCN_CUDA_HOST_COPY=dma TL=0 timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -1
RCN_CUDA_HS_PUBLISH=$pub RCN_CUDA_HOST_COPY=dma TL=0 STEPS=600 timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -1
done
RCN_CUDA_HOST_COPY=pull TL=0 timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -1
for aff in 1 0; do
for rep in 1 2 3; do
RCN_BENCH_AFFINITY=$aff timeout 300 python bench.py --steps 20 --warmup 5 --no-extra > $OUT/r2y_${aff}_${rep}.json 2> $OUT/r2y_${aff}_${rep}.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2y_${aff}_${rep}.json"))
print("affinity=$aff K=20", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), "M", round(d["e2e"]["ms_per_step"] * 1e3, 2), "us/step |", d["config"]["host_cpu_affinity"])
PY
done
done
timeout 300 python bench.py --steps 2000 --warmup 20 --no-extra > $OUT/r2y_long.json 2> $OUT/r2y_long.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2y_long.json"))
print("K=2000", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), "M", round(d["e2e"]["ms_per_step"] * 1e3, 2), "us/step", round(d["e2e"]["h2d_GBps_per_gpu"], 1), "GB/s")
PY
