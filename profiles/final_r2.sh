#!/bin/bash
# Round-2 final pass on one GPU: the whole GPU test suite, smoke(), then the driver's command line and the default one.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2_bench_n1_driver_cmd.json 2> $OUT/r2_bench_n1_driver_cmd.err; echo "rc=$?"
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/r2_bench_reference_arm.json 2> /dev/null; echo "ref rc=$?"
python bench.py > $OUT/r2_bench_n1_default_steps2000.json 2> $OUT/r2_bench_n1_default.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("r2_bench_n1_driver_cmd.json", "r2_bench_n1_default_steps2000.json"):
    d = json.load(open("gpurun_out/" + f))
    print(f, round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), "M; frac", round(d["roofline"]["frac"], 4), "traffic", d["roofline"]["traffic"])
    print("    ", d["roofline"].get("in_graph_timeline"))
    for k, v in d.get("workloads", {}).items():
        print("    ", k, round(v["value"] / 1e6, 3), round(v["ms_per_step"], 4), v["roofline"]["kernel"], round(v["roofline"]["frac"], 3))
r = json.load(open("gpurun_out/r2_bench_reference_arm.json"))
print("reference arm", round(r["value"]), r["cpu_baseline"]["cores"], "cores")
PY
