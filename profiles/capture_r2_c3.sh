#!/bin/bash
# ncu --set full of the c3 step's kernels (f64 DMMA GEMMs, reductions, feature kernel); plain run first.
set -u
OUT=gpurun_out
python profiles/run_step.py c3 4 > $OUT/r2_run_step_c3.log 2>&1 || { echo "run_step failed"; tail -3 $OUT/r2_run_step_c3.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:(gemm_f64|bias_grad|reduce_splits|batch_stats|features_cp|sgd_update)" -s 10 -c 11 -f -o $OUT/r2_full_c3 \
    python profiles/run_step.py c3 4 > $OUT/r2_ncu_full_c3.log 2>&1; echo "ncu rc=$?"
ls -la $OUT/r2_full_c3.ncu-rep
