#!/bin/bash
# Runs ON THE GPU BOX (gpurun): ncu launch list of the host-streaming epoch with the batched-load prefetch kernel and with
# the old load/store-paired loop, plus one --set full capture of the prefetch kernel.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 100 python profiles/run_host_epoch.py 24 > $OUT/run_host_epoch_r1s.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/run_host_epoch_r1s.log; exit 1; }
for u in 8 1; do
  RCN_CUDA_PREFETCH_UNROLL=$u timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/launches_r1s_host_epoch_unroll$u.csv python profiles/run_host_epoch.py 24 > $OUT/ncu_host_epoch_unroll$u.log 2>&1
done
timeout 150 ncu --set full --clock-control none --import-source on -k "regex:host_prefetch" -s 4 -c 3 -f -o $OUT/full_r1s_host_prefetch \
    python profiles/run_host_epoch.py 24 > $OUT/ncu_full_host_prefetch.log 2>&1
for u in 8 1; do python profiles/summarize_launches.py $OUT/launches_r1s_host_epoch_unroll$u.csv | head -8; done
ls -la $OUT | tail -5
