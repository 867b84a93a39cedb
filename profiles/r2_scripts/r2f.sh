#!/bin/bash
set -u
OUT=gpurun_out
for occ in 6 4 6 4; do
RCN_CUDA_CONV_OCC=$occ python profiles/features_bench.py c4 2>/dev/null | tail -1 | sed "s/^/occ=$occ /"
done
timeout 300 python -m pytest tests/test_gpu_features.py -m gpu -q -x 2>&1 | tail -2
bash profiles/final_r2.sh
