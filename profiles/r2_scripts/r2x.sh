#!/bin/bash
set -u
for v in dma pull; do
RCN_CUDA_HOST_COPY=$v timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -2 | cut -c1-300
RCN_CUDA_HOST_COPY=$v TL=0 timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -1
done
RCN_CUDA_HOST_COPY=dma RCN_CUDA_HOST_STEPS_PER_GRAPH=5 TL=0 timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -1
RCN_CUDA_HOST_COPY=dma RCN_CUDA_HOST_STEPS_PER_GRAPH=10 TL=0 timeout 200 python profiles/r2_scripts/e2e_short.py 2>&1 | tail -1
