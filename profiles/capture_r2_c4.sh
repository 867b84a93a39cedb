#!/bin/bash
# ncu --set full of the c4 feature stage (four stacked Same convolutions, the layered path's strip kernels); plain run first.
set -u
OUT=gpurun_out
python profiles/features_bench.py c4 > $OUT/r2_features_bench_c4.jsonl 2> $OUT/r2_features_bench_c4.err || { echo "plain run failed"; tail -3 $OUT/r2_features_bench_c4.err; exit 1; }
cat $OUT/r2_features_bench_c4.jsonl
ncu --set full --clock-control none --import-source on -k "regex:(conv_same|convert_images)" -s 8 -c 5 -f -o $OUT/r2_full_c4 \
    python profiles/features_bench.py c4 > $OUT/r2_ncu_full_c4.log 2>&1; echo "ncu rc=$?"
ls -la $OUT/r2_full_c4.ncu-rep
