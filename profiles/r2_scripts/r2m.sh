#!/bin/bash
# 8 GPUs: full default bench (c2 weak + parity, c3 strong, c4, c5 NCCL)
set -u
OUT=gpurun_out
T0=$(date +%s)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 8 > $OUT/r2m_bench_n8.json 2> $OUT/r2m_bench_n8.err; echo "n8 rc=$?"
echo "bench wall $(( $(date +%s) - T0 )) s"; tail -4 $OUT/r2m_bench_n8.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2m_bench_n8.json"))
print("c2", round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2))
print("   ", d.get("parity"))
print("   ", d["roofline"].get("in_graph_timeline"))
for k, v in d.get("workloads", {}).items():
    if "error" in v: print(k, v); continue
    print(k, round(v["value"] / 1e6, 3), "M img/s", round(v["ms_per_step"], 4), "ms;", v.get("scaling"), v.get("exchange"), "e2e", v.get("e2e", {}).get("value"))
PY
