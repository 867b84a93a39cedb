#!/bin/bash
# Runs ON THE GPU BOX (gpurun): GPU test suite after the stale-graph fix, the e2e fixed-cost probe, and the
# steps-per-graph sweep of the host-dataset loop (RCN_CUDA_HOST_STEPS_PER_GRAPH) through bench.py's e2e leg.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 240 python -m pytest tests -m gpu -x -q > $OUT/tests_r1s_b.log 2>&1; echo "tests rc=$?" | tee -a $OUT/tests_r1s_b.log
rm -f $OUT/e2e_fixed_cost_r1s.jsonl $OUT/e2e_fixed_cost_r1s.err
for g in 2 4 8; do
  RCN_CUDA_HOST_STEPS_PER_GRAPH=$g timeout 100 python profiles/e2e_fixed_cost.py >> $OUT/e2e_fixed_cost_r1s.jsonl 2>> $OUT/e2e_fixed_cost_r1s.err
  RCN_CUDA_HOST_STEPS_PER_GRAPH=$g timeout 200 python bench.py > $OUT/bench_r1s_c2_hspg$g.json 2> /dev/null; echo "bench hspg$g rc=$?"
done
tail -4 $OUT/tests_r1s_b.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1s_c2_hspg*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), d['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
for l in open('gpurun_out/e2e_fixed_cost_r1s.jsonl'):
    d=json.loads(l); print(d['fit'], [(c['steps'], c['us_per_step']) for c in d['calls']])
PY
tail -3 $OUT/e2e_fixed_cost_r1s.err
