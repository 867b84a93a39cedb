#!/bin/bash
# 2 GPUs: DP tests, then the driver's command line at N=1 and N=2
set -u
OUT=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "dp_ or nccl" > $OUT/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/r2o_pytest.log
show() { python - <<PY
import json
d = json.load(open("gpurun_out/$1"))
print("$1", round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), "M", round(d["e2e"]["ms_per_step"] * 1e3, 2), "us", round(d["e2e"]["h2d_GBps_per_gpu"], 1), "GB/s", (d.get("parity") or {}).get("max_rel_vs_single_gpu"))
print("    ", d["roofline"].get("in_graph_timeline"))
for k, v in d.get("workloads", {}).items():
    print("    ", k, round(v["value"] / 1e6, 3) if "value" in v else v, round(v.get("ms_per_step", 0), 4))
PY
}
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2o_n1.json 2> $OUT/r2o_n1.err; show r2o_n1.json
python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra > $OUT/r2o_n1b.json 2> $OUT/r2o_n1b.err; show r2o_n1b.json
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --steps 20 --warmup 5 > $OUT/r2o_n2.json 2> $OUT/r2o_n2.err; echo "n2 rc=$?"; show r2o_n2.json
$TR bench.py --gpus 2 --no-extra > $OUT/r2o_n2_2000.json 2> $OUT/r2o_n2_2000.err; echo "n2 rc=$?"; show r2o_n2_2000.json
