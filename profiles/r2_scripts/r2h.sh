#!/bin/bash
# full default bench line (c2 + c3/c4/c5 legs), 1 GPU
set -u
OUT=gpurun_out
T0=$(date +%s)
timeout 900 python bench.py > $OUT/r2h_bench.json 2> $OUT/r2h_bench.err; echo "bench rc=$?"
echo "bench wall $(( $(date +%s) - T0 )) s"; tail -3 $OUT/r2h_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2h_bench.json"))
print("c2", round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), "frac", round(d["roofline"]["frac"], 4))
print("   ", d["roofline"].get("in_graph_timeline"))
for k, v in d.get("workloads", {}).items():
    if "error" in v: print(k, v); continue
    print(k, round(v["value"] / 1e6, 3), "M img/s", round(v["ms_per_step"], 4), "ms; e2e", v.get("e2e", {}).get("value"), v["roofline"]["kernel"], round(v["roofline"]["frac"], 4), v["clocks"]["reasons"] if v.get("clocks") else None)
PY
