#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_dense.py -m gpu -q -x -k "host" 2>&1 | tail -2
for rep in 1 2; do
RCN_CUDA_HOST_COPY=dma TL=0 timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -1
RCN_CUDA_HOST_COPY=dma TL=0 STEPS=600 timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -1
RCN_CUDA_HOST_COPY=dma RCN_CUDA_HOST_STEPS_PER_GRAPH=20 TL=0 STEPS=600 timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -1 | sed "s/^/spg20 /"
RCN_CUDA_HOST_COPY=pull TL=0 STEPS=600 timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -1
done
