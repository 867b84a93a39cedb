#!/bin/bash
# 2 GPUs: DP tests (2-GPU variants) + full default bench at N=2
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "dp_ or nccl" > $OUT/r2i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2i_pytest.log
tail -6 $OUT/r2i_pytest.log
T0=$(date +%s)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 > $OUT/r2i_bench_n2.json 2> $OUT/r2i_bench_n2.err; echo "n2 rc=$?"
echo "bench wall $(( $(date +%s) - T0 )) s"; tail -4 $OUT/r2i_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2i_bench_n2.json"))
print("c2", round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2))
print("   ", d.get("parity"))
print("   ", d["roofline"].get("in_graph_timeline"))
for k, v in d.get("workloads", {}).items():
    if "error" in v: print(k, v); continue
    print(k, round(v["value"] / 1e6, 3), "M img/s", round(v["ms_per_step"], 4), "ms;", v.get("scaling"), v.get("exchange"), "e2e", v.get("e2e", {}).get("value"))
PY
