"""Copy-engine host->device rate from pinned memory at the chunk sizes the host-dataset loop could use (steady state:
the link is warmed for 0.3 s first, then 40 back-to-back copies per size are timed with CUDA events)."""
import json
import time
import torch

dev = torch.device("cuda:0")
host = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
host.random_(0, 255)
ring = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
s = torch.cuda.Stream()
out = []
with torch.cuda.stream(s):
    t0 = time.time()
    while time.time() - t0 < 0.4:
        ring[: 8 << 20].copy_(host[: 8 << 20], non_blocking=True)
        s.synchronize()
    for nbytes in [802816, 2 * 802816, 4 * 802816, 5 * 802816, 8 * 802816, 16 * 802816, 32 * 802816, 64 << 20]:
        reps = 40 if nbytes < (32 << 20) else 8
        best = None
        for trial in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for r in range(reps):
                off = (r * nbytes) % ((256 << 20) - nbytes)
                off -= off % 4096
                ring[:nbytes].copy_(host[off: off + nbytes], non_blocking=True)
            e1.record(s)
            s.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            best = us if best is None else min(best, us)
        out.append({"bytes": nbytes, "us_per_copy": round(best, 2), "GBps": round(nbytes / best / 1e3, 2)})
        print(out[-1], flush=True)
json.dump({"h2d_pinned_copy_engine_steady": out}, open("gpurun_out/r2v_pcie_dma.json", "w"))
