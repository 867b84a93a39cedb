#!/bin/bash
set -u
OUT=gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2_bench_n1_driver_cmd.json 2> $OUT/r2_bench_n1_driver_cmd.err; echo "rc=$?"
python bench.py > $OUT/r2_bench_n1_default_steps2000.json 2> $OUT/r2_bench_n1_default.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_n1_driver_cmd.json", "r2_bench_n1_default_steps2000.json"):
    d = json.load(open("gpurun_out/" + f))
    print(f, round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), "M; frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"])
    for k, v in d.get("workloads", {}).items():
        print("    ", k, round(v["value"] / 1e6, 3), round(v["ms_per_step"], 4), v["roofline"]["kernel"], round(v["roofline"]["frac"], 3), v["roofline"].get("traffic"))
PY
