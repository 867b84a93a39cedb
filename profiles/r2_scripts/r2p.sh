#!/bin/bash
# 8 GPUs, final code: the driver's command line (K=20, W=5, all legs) and the steady-state c2 line (2000 steps)
set -u
OUT=gpurun_out
show() { python - <<PY
import json
d = json.load(open("gpurun_out/$1"))
print("$1", round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), "M", round(d["e2e"]["ms_per_step"] * 1e3, 2), "us", round(d["e2e"]["h2d_GBps_per_gpu"], 1), "GB/s", (d.get("parity") or {}))
print("    ", d["roofline"].get("in_graph_timeline"))
for k, v in d.get("workloads", {}).items():
    print("    ", k, round(v["value"] / 1e6, 3) if "value" in v else v, round(v.get("ms_per_step", 0), 4), v.get("exchange"))
PY
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/r2p_n8_driver_cmd.json 2> $OUT/r2p_n8_driver_cmd.err; echo "rc=$?"; show r2p_n8_driver_cmd.json
timeout 600 $TR bench.py --gpus 8 --no-extra > $OUT/r2p_n8_2000.json 2> $OUT/r2p_n8_2000.err; echo "rc=$?"; show r2p_n8_2000.json
