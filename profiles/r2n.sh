#!/bin/bash
# the driver's own command line (K=20, W=5), plus the write-combined host-buffer experiment
set -u
OUT=gpurun_out
show() { python - <<PY
import json
d = json.load(open("gpurun_out/$1"))
print("$1", round(d["value"] / 1e6, 2), "M img/s", round(d["ms_per_step"] * 1e3, 2), "us; e2e", round(d["e2e"]["value"] / 1e6, 2), "M", round(d["e2e"]["ms_per_step"] * 1e3, 2), "us", d["e2e"]["host_buffers"], round(d["e2e"]["h2d_GBps_per_gpu"], 1), "GB/s")
PY
}
T0=$(date +%s)
python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/r2n_driver.json 2> $OUT/r2n_driver.err; echo "rc=$? wall $(( $(date +%s) - T0 )) s"; show r2n_driver.json
python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra --host-alloc wc > $OUT/r2n_driver_wc.json 2> $OUT/r2n_driver_wc.err; show r2n_driver_wc.json
python bench.py --no-extra --host-alloc wc > $OUT/r2n_2000_wc.json 2> $OUT/r2n_2000_wc.err; show r2n_2000_wc.json
python bench.py --no-extra > $OUT/r2n_2000.json 2> $OUT/r2n_2000.err; show r2n_2000.json
