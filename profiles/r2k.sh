#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r2k_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2k_pytest.log
tail -5 $OUT/r2k_pytest.log
timeout 600 python bench.py > $OUT/r2k_bench.json 2> $OUT/r2k_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2k_bench.json"))
print("c2", round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), "frac", round(d["roofline"]["frac"], 4))
for k, v in d.get("workloads", {}).items():
    if "error" in v: print(k, v); continue
    print(k, round(v["value"] / 1e6, 3), "M img/s", round(v["ms_per_step"], 4), "ms; e2e", v.get("e2e", {}).get("value"), v["roofline"]["kernel"], round(v["roofline"]["frac"], 4))
    if k == "c3":
        for kk, vv in v["roofline"]["kernels"].items(): print("     ", kk, vv)
PY
