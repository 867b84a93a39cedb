#!/bin/bash
# Runs ON THE GPU BOX (gpurun): the bench line as the driver will run it, plus the steps-per-graph comparison.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 150 python bench.py > $OUT/bench_r1s_d_c2.json 2> $OUT/bench_r1s_d_c2.err; echo "bench rc=$?"
timeout 150 python bench.py --steps-per-graph 16 > $OUT/bench_r1s_d_c2_spg16.json 2> /dev/null; echo "bench spg16 rc=$?"
timeout 100 python bench.py --steps 20 --warmup 3 > $OUT/bench_r1s_d_c2_steps20.json 2> /dev/null; echo "bench20 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1s_d_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), d['ms_per_step'], d['config']['steps_per_graph'], round(d['e2e']['h2d_GBps_per_gpu'],2), d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 $OUT/bench_r1s_d_c2.err
