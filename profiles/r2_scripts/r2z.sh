#!/bin/bash
# 2 GPUs: data-parallel tests with the final library (the PREWAIT kernel now stores its features after the wait)
set -u
timeout 900 python -m pytest tests/test_gpu_dense.py -m gpu -q -x -k "two or dp_" 2>&1 | tail -3
