#!/bin/bash
# Round-2 pass A (runs ON THE GPU BOX): full GPU test suite, default bench line, kernel-A phase clocks + in-graph timeline,
# one `--set full` capture of the c2 step kernels.  Outputs -> gpurun_out/.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r2a_pytest.log
tail -15 $OUT/r2a_pytest.log
timeout 600 python bench.py > $OUT/r2a_bench.json 2> $OUT/r2a_bench.err; echo "bench rc=$?"
tail -3 $OUT/r2a_bench.err
RCN_CUDA_LIB=profiles/_build/librcn_cuda_phases.so timeout 120 python profiles/sn_phases.py > $OUT/r2a_phases.txt 2>&1
timeout 120 python profiles/sn_phases.py timeline >> $OUT/r2a_phases.txt 2>&1
cat $OUT/r2a_phases.txt
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:^(smallnet)" -s 4 -c 4 -f -o $OUT/full_r2a_c2 \
    python profiles/run_step.py c2 6 > $OUT/r2a_ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la $OUT | tail -8
