#!/bin/bash
# Runs ON THE GPU BOX (gpurun): full GPU test suite, smoke, the default bench, then the host-prefetch A/B
# (RCN_CUDA_PREFETCH_UNROLL / _CTAS) through bench.py's e2e leg and profiles/e2e_fixed_cost.py.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 240 python -m pytest tests -m gpu -x -q > $OUT/tests_r1s.log 2>&1; echo "tests rc=$?" | tee -a $OUT/tests_r1s.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke_r1s.log 2>&1; echo "smoke rc=$?"
timeout 200 python bench.py > $OUT/bench_r1s_c2.json 2> $OUT/bench_r1s_c2.err; echo "bench rc=$?"
RCN_CUDA_PREFETCH_UNROLL=1 timeout 200 python bench.py > $OUT/bench_r1s_c2_unroll1.json 2> /dev/null; echo "bench unroll1 rc=$?"
RCN_CUDA_PREFETCH_CTAS=64 timeout 200 python bench.py > $OUT/bench_r1s_c2_ctas64.json 2> /dev/null; echo "bench ctas64 rc=$?"
timeout 100 python bench.py --steps 20 --warmup 3 > $OUT/bench_r1s_c2_steps20.json 2> /dev/null; echo "bench20 rc=$?"
for u in 8 1; do RCN_CUDA_PREFETCH_UNROLL=$u timeout 100 python profiles/e2e_fixed_cost.py >> $OUT/e2e_fixed_cost_r1s.jsonl 2>> $OUT/e2e_fixed_cost_r1s.err; done
RCN_CUDA_PREFETCH_CTAS=16 timeout 100 python profiles/e2e_fixed_cost.py >> $OUT/e2e_fixed_cost_r1s.jsonl 2>> $OUT/e2e_fixed_cost_r1s.err
tail -3 $OUT/tests_r1s.log; cat $OUT/smoke_r1s.log | tail -2
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1s_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), d['ms_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
cat $OUT/e2e_fixed_cost_r1s.jsonl | cut -c1-900
