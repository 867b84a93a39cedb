#!/bin/bash
# Round-2 evidence pass (runs ON THE GPU BOX, one GPU): plain bench runs first (the driver's command line and the default one),
# then the ncu launch list of the same command, one `--set full` capture of the c2 step kernels (default cache control and
# `--cache-control none`: kernel B's features are L2-resident in the real step), the phase clocks of kernel A.
# Outputs -> gpurun_out/r2_*; profiles/summarize_rep.py / summarize_launches.py turn them into the committed summaries.
set -u
OUT=gpurun_out
mkdir -p $OUT
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2_bench_driver_cmd.json 2> $OUT/r2_bench_driver_cmd.err; echo "driver-cmd bench rc=$?"
python bench.py > $OUT/r2_bench_default.json 2> $OUT/r2_bench_default.err; echo "default bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2_launches_c2.csv \
    python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra > $OUT/r2_ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:^(smallnet)" -s 4 -c 4 -f -o $OUT/r2_full_c2 \
    python profiles/run_step.py c2 6 > $OUT/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none --cache-control none -k "regex:^(smallnet)" -s 4 -c 4 -f -o $OUT/r2_full_c2_warm \
    python profiles/run_step.py c2 6 > $OUT/r2_ncu_full_warm.log 2>&1; echo "ncu full (cache-control none) rc=$?"
RCN_CUDA_LIB=profiles/_build/librcn_cuda_phases.so python profiles/sn_phases.py > $OUT/r2_phases.txt 2>&1
python profiles/sn_phases.py timeline >> $OUT/r2_phases.txt 2>&1
python profiles/features_bench.py > $OUT/r2_features_bench.jsonl 2> $OUT/r2_features_bench.err
tail -14 $OUT/r2_phases.txt
python - <<'PY'
import json
for f in ("r2_bench_driver_cmd.json", "r2_bench_default.json"):
    d = json.load(open("gpurun_out/" + f))
    print(f, round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), "M; frac", round(d["roofline"]["frac"], 4))
    for k, v in d.get("workloads", {}).items():
        print("    ", k, round(v["value"] / 1e6, 3), round(v["ms_per_step"], 4), v["roofline"]["kernel"], round(v["roofline"]["frac"], 3))
PY
