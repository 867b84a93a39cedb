#!/bin/bash
# Runs ON THE GPU BOX (gpurun): plain run first, then the ncu launch list of the same command, then one
# `--set full` capture of the step's kernels.  usage: profiles/capture.sh <tag> [workload] [extra bench args]
# Outputs land in gpurun_out/ (scratch); profiles/summarize_rep.py turns them into the committed summaries.
set -u
TAG=${1:-r1}; WL=${2:-c2}; shift 2 || true
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --workload $WL --steps 20 --warmup 3 $*"
$CMD > $OUT/bench_${TAG}_${WL}_plain.json 2> $OUT/bench_${TAG}_${WL}_plain.err || { echo "plain run failed"; tail -5 $OUT/bench_${TAG}_${WL}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_${TAG}_${WL}.csv \
    $CMD > $OUT/ncu_${TAG}_${WL}.log 2>&1
python profiles/run_step.py $WL 4 > $OUT/run_step_${TAG}_${WL}.log 2>&1 || { echo "run_step failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:^(smallnet|features_|gemm_f64|sgd_update|bias_grad|reduce_splits|batch_stats|skinny_|rcn_)" \
    -s 3 -c 9 -f -o $OUT/full_${TAG}_${WL} python profiles/run_step.py $WL 4 > $OUT/ncu_full_${TAG}_${WL}.log 2>&1
ls -la $OUT | tail -8
