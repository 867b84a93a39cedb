#!/bin/bash
# 8 GPUs, final code: steady-state c2 (2000 steps), the driver's command line with all legs, and the same at 4 ranks
set -u
OUT=gpurun_out
show() { python - <<PY
import json
d = json.load(open("gpurun_out/$1"))
print("$1", round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), "M", d["config"]["steps_per_graph"], (d.get("parity") or {}).get("max_rel_vs_single_gpu"), (d.get("parity") or {}).get("replicas_bit_identical"))
print("    ", d["roofline"].get("in_graph_timeline"))
for k, v in d.get("workloads", {}).items():
    print("    ", k, round(v["value"] / 1e6, 3) if "value" in v else v, round(v.get("ms_per_step", 0), 4), v.get("exchange"))
PY
}
TR8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR8 bench.py --gpus 8 --no-extra > $OUT/r2s_n8_2000.json 2> $OUT/r2s_n8_2000.err; echo "rc=$?"; show r2s_n8_2000.json
timeout 600 $TR8 bench.py --gpus 8 --steps 20 --warmup 5 > $OUT/r2s_n8_driver_cmd.json 2> $OUT/r2s_n8_driver_cmd.err; echo "rc=$?"; show r2s_n8_driver_cmd.json
timeout 600 $TR4 bench.py --gpus 4 --no-extra > $OUT/r2s_n4_2000.json 2> $OUT/r2s_n4_2000.err; echo "rc=$?"; show r2s_n4_2000.json
timeout 600 $TR4 bench.py --gpus 4 --steps 20 --warmup 5 > $OUT/r2s_n4_driver_cmd.json 2> $OUT/r2s_n4_driver_cmd.err; echo "rc=$?"; show r2s_n4_driver_cmd.json
