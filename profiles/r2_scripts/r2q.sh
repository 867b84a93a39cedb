#!/bin/bash
# evidence pass after the strip kernels / single-GPU prewait: ncu --set full of the c4 feature stage and of the c2 step kernels
set -u
OUT=gpurun_out
bash profiles/capture_r2_c4.sh
python profiles/run_step.py c2 6 > $OUT/r2b_run_step_c2.log 2>&1 || { echo "run_step failed"; tail -3 $OUT/r2b_run_step_c2.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:^(smallnet)" -s 4 -c 4 -f -o $OUT/r2b_full_c2 \
    python profiles/run_step.py c2 6 > $OUT/r2b_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none --cache-control none -k "regex:^(smallnet)" -s 4 -c 4 -f -o $OUT/r2b_full_c2_warm \
    python profiles/run_step.py c2 6 > $OUT/r2b_ncu_full_warm.log 2>&1; echo "ncu full (cache-control none) rc=$?"
ls -la $OUT/*.ncu-rep
