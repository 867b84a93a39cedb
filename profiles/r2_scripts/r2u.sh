#!/bin/bash
# ncu --set full of the c4 final-stage kernels
set -u
OUT=gpurun_out
cat > /tmp/c4_run.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from mercer_research_b200 import RCN
m = RCN(10, [1, 1, 1, 1], [30])
imgs = torch.randint(0, 256, (128, 64, 64), dtype=torch.uint8, device="cuda")
out = torch.empty((128, m.feature_len(64, 64)), dtype=torch.float64, device="cuda")
m.scale_set = (100.0, 50.0)
for _ in range(3):
    m.flatten_feature_set(imgs, standardise=True, out=out)
torch.cuda.synchronize()
print("ok")
PY
ncu --set full --clock-control none --import-source on -k "regex:(conv_same_final|stage_final)" -s 1 -c 1 -f -o $OUT/r2_full_c4_strips python /tmp/c4_run.py > $OUT/r2_ncu_c4.log 2>&1; echo "rc=$?"
RCN_CUDA_CONV_STRIPS=0 ncu --set full --clock-control none --import-source on -k "regex:(conv_same_final|stage_final)" -s 1 -c 1 -f -o $OUT/r2_full_c4_generic python /tmp/c4_run.py > $OUT/r2_ncu_c4b.log 2>&1; echo "rc=$?"
