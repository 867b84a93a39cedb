"""Times the GEMM building block (rcn_cuda_ext_gemm_f64) on device-resident operands: DMMA (impl 0) vs the tcgen05
integer-slice path (impl 1, slicing included), next to cuBLAS DGEMM (torch.matmul f64). Per-kernel split from the
library's launch profiler. usage: python profiles/gemm_bench.py [M N K]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mercer_research_b200 import _lib  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (4096, 8192, 4096)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
A = torch.randn(M, K, dtype=torch.float64, device=dev, generator=g)      # [m][k]  k-contiguous
Bt = torch.randn(N, K, dtype=torch.float64, device=dev, generator=g)     # [n][k]  k-contiguous
C = torch.empty(N, M, dtype=torch.float64, device=dev)                   # column-major M x N
lib = _lib.load()
stream = torch.cuda.current_stream().cuda_stream
flops = 2.0 * M * N * K


def run(impl):
    _lib.check(lib.rcn_cuda_ext_gemm_f64(0, stream, A.data_ptr(), K, 1, Bt.data_ptr(), K, 1, M, N, K, impl, C.data_ptr()))


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


out = {"M": M, "N": N, "K": K}
ref = A @ Bt.T
t = timed(lambda: torch.matmul(A, Bt.T))
out["cublas_dgemm_ms"] = t; out["cublas_dgemm_tflops"] = flops / t / 1e9
for impl, name in [(0, "dmma"), (1, "tcgen05_int8_slices")]:
    t = timed(lambda: run(impl))
    err = float((C.T - ref).abs().max() / ref.abs().max())
    out[name + "_ms"] = t
    out[name + "_f64_equiv_tflops"] = flops / t / 1e9
    out[name + "_max_err_rel_to_max"] = err
_lib.profile_enable(True)
for _ in range(3):
    run(1)
torch.cuda.synchronize()
prof = _lib.profile_report()
_lib.profile_enable(False)
out["tc_kernels_ms"] = {k: v["total_ms"] / v["launches"] for k, v in prof.items()}
mma = [v for k, v in out["tc_kernels_ms"].items() if "ext_gemm" in k]
if mma:
    out["tc_mma_kernel_int8_tops"] = 15 * flops / mma[0] / 1e9
print(json.dumps(out))
