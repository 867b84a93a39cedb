#!/bin/bash
# Round-2 pass B (GPU box): full GPU suite (all failures), c2 bench with the 48-column kernel B vs the 64-column one, timeline.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/r2b_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2b_pytest.log
tail -25 $OUT/r2b_pytest.log
timeout 300 python bench.py --no-extra > $OUT/r2b_bench_cw48.json 2> $OUT/r2b_bench_cw48.err; echo "bench rc=$?"
RCN_CUDA_WGRAD_CW=64 timeout 300 python bench.py --no-extra > $OUT/r2b_bench_cw64.json 2> $OUT/r2b_bench_cw64.err; echo "bench rc=$?"
python - <<'PY'
import json
for t in ("cw48", "cw64"):
    try:
        d = json.load(open(f"gpurun_out/r2b_bench_{t}.json"))
        print(t, round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), d["roofline"].get("in_graph_timeline"))
    except Exception as e:
        print(t, "failed", e)
PY
RCN_CUDA_LIB=profiles/_build/librcn_cuda_phases.so timeout 120 python profiles/sn_phases.py > $OUT/r2b_phases.txt 2>&1
timeout 120 python profiles/sn_phases.py timeline >> $OUT/r2b_phases.txt 2>&1
tail -12 $OUT/r2b_phases.txt
