#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -q -x -k "two or dp_ or host" 2>&1 | tail -2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 2000 --warmup 20 --no-extra > $OUT/r2_bench_n2_steps2000_final.json 2> $OUT/r2z_n2_long.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_n2_steps2000_final.json"))
p = d.get("parity") or {}
print("N=2 K=2000", round(d["value"] / 1e6, 2), "M", round(d["ms_per_step"] * 1e3, 2), "us  e2e", round(d["e2e"]["value"] / 1e6, 2), "M | parity", p.get("replicas_bit_identical"), p.get("max_rel_vs_single_gpu"), p.get("max_rel_vs_oracle"))
PY
