#!/bin/bash
set -u
T="tests/test_gpu_dense.py::test_dp_epoch_steps_two_ranks_prewait"
for v in "" "RCN_CUDA_DP_PREWAIT=0" "RCN_CUDA_PDL=0" "RCN_CUDA_L2_PREFETCH=0" "RCN_CUDA_DP_FUSED_PUSH=0"; do
  echo "=== $v"
  env $v timeout 200 python -m pytest "$T" -q -m gpu 2>&1 | grep -E "passed|failed|elements off" | head -4
done
