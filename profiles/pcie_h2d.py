"""Host->device copy-engine bandwidth of this box (pinned memory, cudaMemcpyAsync, CUDA events): the PCIe ceiling of the
e2e path, next to the ~38 GB/s the SM-issued zero-copy pulls of host_prefetch_kernel reach. One JSON line."""
import json

import torch

dev = torch.device("cuda", 0)
out = []
for nbytes in (802816, 8 << 20, 256 << 20):
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    reps = 200 if nbytes < (64 << 20) else 20
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    out.append({"bytes": nbytes, "us_per_copy": round(us, 2), "GBps": round(nbytes / us * 1e-3, 2)})
print(json.dumps({"h2d_pinned_copy_engine": out}))
