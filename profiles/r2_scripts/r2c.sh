#!/bin/bash
# Round-2 pass C (GPU box, 2 GPUs): data-parallel tests incl. the 2-GPU variants, c2 bench at N=1 and N=2 with the front end
# ahead of the exchange (default) and without it (RCN_CUDA_DP_PREWAIT=0).
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -k "dp_ or nccl or configs" > $OUT/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2c_pytest.log
tail -8 $OUT/r2c_pytest.log
timeout 300 python bench.py --no-extra > $OUT/r2c_bench_n1.json 2> $OUT/r2c_bench_n1.err; echo "n1 rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 2 --no-extra > $OUT/r2c_bench_n2.json 2> $OUT/r2c_bench_n2.err; echo "n2 rc=$?"
RCN_CUDA_DP_PREWAIT=0 timeout 400 $TR bench.py --gpus 2 --no-extra > $OUT/r2c_bench_n2_noprewait.json 2> $OUT/r2c_bench_n2_noprewait.err; echo "n2 noprewait rc=$?"
tail -3 $OUT/r2c_bench_n2.err
python - <<'PY'
import json
for t in ("n1", "n2", "n2_noprewait"):
    try:
        d = json.load(open(f"gpurun_out/r2c_bench_{t}.json"))
        print(t, round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), d.get("parity"))
        print("   ", d["roofline"].get("in_graph_timeline"))
    except Exception as e:
        print(t, "failed", e)
PY
