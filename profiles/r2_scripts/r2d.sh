#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/r2e_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2e_pytest.log
tail -12 $OUT/r2e_pytest.log
timeout 300 python bench.py --no-extra > $OUT/r2e_bench.json 2> $OUT/r2e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2e_bench.json"))
print(round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), d["roofline"].get("in_graph_timeline"))
PY
RCN_CUDA_LIB=profiles/_build/librcn_cuda_phases.so timeout 120 python profiles/sn_phases.py > $OUT/r2e_phases.txt 2>&1
tail -12 $OUT/r2e_phases.txt
