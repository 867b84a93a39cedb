#!/bin/bash
# 2 GPUs: data-parallel tests (incl. host-input epochs through the copy-engine path) + the bench line
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q -x -k "two or dp_ or host" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra > $OUT/r2z_n2.json 2> $OUT/r2z_n2.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2z_n2.json"))
print("N=2 K=20", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), "M", round(d["e2e"]["ms_per_step"] * 1e3, 2), "us/step |", d.get("parity"))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 2000 --warmup 20 --no-extra > $OUT/r2z_n2_long.json 2> $OUT/r2z_n2_long.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2z_n2_long.json"))
print("N=2 K=2000", round(d["value"] / 1e6, 2), "M  e2e", round(d["e2e"]["value"] / 1e6, 2), "M", round(d["e2e"]["ms_per_step"] * 1e3, 2), "us/step")
PY
