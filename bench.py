#!/usr/bin/env python
"""bench.py -- training images/sec of rcn's hot path on N B200s (one process per GPU) + roofline + CPU baseline.

  python bench.py --gpus 1 --steps K --warmup W                       # this repo's CUDA path
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference --gpus N --steps K --warmup W      # reference CPU algorithm (oracle) on host cores

A "step" is one pass of the hot path over one minibatch: u8 images -> Sobel/ReLU/max-pool features -> standardise
-> sigmoid MLP forward -> backprop -> batch gradient sum -> (all-reduce over ranks) -> SGD update
(rcn/src/rcn.rs:317-356, 407-412, 260-314, 176-223).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs -> concrete shapes. `batch` is PER GPU (weak scaling).
WORKLOADS = {
    # configs[1]: MNIST-shaped CNN (28x28x1, one conv+pool+dense+softmax-position layer), batch 1024
    "c2": dict(desc="rcn MNIST-shaped CNN 28x28 [Conv(Same),Pool(Max)] 784-30-10, batch 1024 per GPU",
               H=28, W=28, cfg=[1, 3], ff=[30], classes=10, batch=1024),
    # configs[0]: the CPU-runnable case (same net, batch 32)
    "c1": dict(desc="rcn MNIST-shaped CNN 28x28 [Conv(Same),Pool(Max)] 784-30-10, batch 32 per GPU",
               H=28, W=28, cfg=[1, 3], ff=[30], classes=10, batch=32),
    # configs[2] in reference semantics: grayscale 32x32, [C,P]x3 -> 64 maps 4x4 = 1024 feats -> 1024-256-10
    "c3": dict(desc="CIFAR-shaped 32x32 gray [C,P]x3 1024-256-10, batch 4096 per GPU",
               H=32, W=32, cfg=[1, 3, 1, 3, 1, 3], ff=[256], classes=10, batch=4096),
    # configs[4]: 64x64 [C,P] -> 4096 feats -> 4096-4096-4096-10
    "c5": dict(desc="dense-heavy head 64x64 [C,P] 4096-4096-4096-10, batch 8192 per GPU",
               H=64, W=64, cfg=[1, 3], ff=[4096, 4096], classes=10, batch=8192),
}
_RESULT_FD = 1


def emit_result(line: dict):
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


ETA = 3.0  # main.rs:32
DATA_SEED, PARAM_SEED = 0x5EED, 0xC0FFEE


def feature_len(cfg, H, W):
    maps, h, w = 0, H, W
    for c in cfg:
        if c in (0, 1):
            maps = maps * 4 if maps else 4
            if c == 0:
                h, w = h - 2, w - 2
        elif maps:
            h, w = (h + 1) // 2, (w + 1) // 2
    return maps * h * w


def layer_shapes(wl):
    L = feature_len(wl["cfg"], wl["H"], wl["W"])
    sizes = [L] + wl["ff"] + [wl["classes"]]
    return [(sizes[i + 1], sizes[i]) for i in range(len(sizes) - 1)]


def dense_flops_per_image(shapes):
    fwd = sum(2 * r * c for r, c in shapes)
    bwd_w = fwd
    bwd_d = sum(2 * r * c for r, c in shapes[1:])
    return fwd, bwd_d, bwd_w


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p.get("bf16_tflops"), source="measured (MEASURED_PEAKS.json)",
                    sm_max_mhz=p.get("sm_max_mhz"))
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)", sm_max_mhz=1965.0)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's threaded train step (C++ restatement of rcn's CPU path -- NOT rustc output)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_train_steps(wl, steps, warmup, max_seconds=None):
    import oracle as O
    B, H, W = wl["batch"], wl["H"], wl["W"]
    shapes = layer_shapes(wl)
    net = O.Net(shapes)
    rng = np.random.default_rng(DATA_SEED)
    n_pool = 4
    images = rng.integers(0, 256, size=(n_pool, B, H, W), dtype=np.uint8)
    labels = (np.arange(B) % wl["classes"]).astype(np.int64)
    params = np.random.default_rng(PARAM_SEED).standard_normal(net.n_params)
    mean, sd = O.gen_scales(O.features_u8(wl["cfg"], images[0][:min(B, 256)]))
    cores = os.cpu_count() or 1
    for i in range(warmup):
        net.train_step_u8(wl["cfg"], params, images[i % n_pool], labels, mean, sd, ETA, cores)
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        net.train_step_u8(wl["cfg"], params, images[i % n_pool], labels, mean, sd, ETA, cores)
        done += 1
        if max_seconds and time.perf_counter() - t0 > max_seconds and done >= 3:
            break
    dt = time.perf_counter() - t0
    return done * B / dt, dt / done * 1e3, done, cores


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    # bounded sample: each step is one minibatch of the same workload on all host threads; cap the run at ~2 min
    B = wl["batch"]
    ips, ms, done, cores = cpu_train_steps(wl, args.steps, args.warmup, max_seconds=120.0)
    line = {
        "impl": "reference", "metric": "training images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "batch_per_step": B, "pixels": "u8", "eta": ETA,
                   "note": "C++ restatement of rcn's CPU path (oracle/rcn_oracle.cpp: per-sample matvec backprop, worker threads "
                           "+ mutex-ordered gradient sum as rcn.rs:176-223); the Rust crate cannot be built here (no rustc)"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{done} steps of batch {B} (features+fwd+bwd+SGD) on {cores} host threads"},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_result(line)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, dev):
    """cuBLAS DGEMM 8192^3 via torch.matmul -- the f64 denominator MEASURED_PEAKS.json lacks."""
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2 * n ** 3 / (best * 1e-3) / 1e12


def run_gpu(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from mercer_research_b200 import RCN, _lib
    from mercer_research_b200.trainer import DataParallelTrainer

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, H, W = wl["batch"], wl["H"], wl["W"]
    shapes = layer_shapes(wl)
    L = shapes[0][1]
    n_params = sum(r * c + r for r, c in shapes)

    model = RCN(wl["classes"], wl["cfg"], wl["ff"], device=local_rank)
    assert model.feature_len(H, W) == L
    model.load_weights_and_bias(L)
    assert model.layer_shapes == shapes
    model.set_params(np.random.default_rng(PARAM_SEED).standard_normal(n_params))   # same replica on every rank
    trainer = DataParallelTrainer(model, eta=ETA)

    # synthetic dataset resident in HBM, larger than L2 (126 MB) so that consecutive steps never hit in L2
    pool_bytes = 192 << 20
    n_batches = max(4, -(-pool_bytes // (B * H * W)))
    g = torch.Generator(device=dev); g.manual_seed(DATA_SEED + rank)
    images = torch.randint(0, 256, (n_batches, B, H, W), dtype=torch.uint8, device=dev, generator=g)
    labels = (torch.arange(B, device=dev) % wl["classes"]).to(torch.int64)
    raw = model.flatten_feature_set(images[0][:min(B, 1024)])
    model.gen_scales(raw)                       # (mean, sd) are fixed inputs on the streaming path
    if world > 1:                               # every replica must use the same scale_set
        ms = torch.tensor(model.scale_set, dtype=torch.float64, device=dev)
        dist.broadcast(ms, 0)
        model.scale_set = tuple(ms.tolist())
    del raw

    stream = torch.cuda.current_stream(dev)
    model.set_stream(stream.cuda_stream)

    # epoch mode: the resident dataset is walked in chunks_exact(B) steps by a device-side cursor (rcn.rs:144-149),
    # so ONE captured CUDA graph (all kernels + the all-reduce) replays for every step.
    all_labels = labels.repeat(n_batches)
    trainer.bind_dataset(images.view(n_batches * B, H, W), all_labels, B)
    kernels_per_step = None
    if not args.no_graph:
        l_before = _lib.kernel_launches()
        trainer.capture(warmup=3)
        kernels_per_step = (_lib.kernel_launches() - l_before) // 4   # 3 warm-up steps + 1 captured step

    def step(i):
        trainer.epoch_step()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    l0 = _lib.kernel_launches()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record(stream)
    barrier()
    clk = clocks.stop() if rank == 0 else None
    launches = _lib.kernel_launches() - l0
    if kernels_per_step is not None:
        launches = kernels_per_step * args.steps   # graph replays re-launch the captured kernels
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- e2e: the reference-facing call with HOST buffers (pinned), H2D + D2H inside the timed region -------------
    n_host = 8
    h_images = torch.randint(0, 256, (n_host, B, H, W), dtype=torch.uint8).pin_memory()
    h_labels = (torch.arange(B) % wl["classes"]).to(torch.int64).pin_memory()
    hi = [h_images[i].numpy() for i in range(n_host)]
    hl = h_labels.numpy()

    def e2e_step(i):
        return trainer.step_images_host(hi[i % n_host], hl)   # returns (cost, hits) read back from the device

    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for i in range(args.steps):
        e2e_step(i)
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B / (e2e_ms / args.steps * 1e-3)
    h2d = B * H * W + B * 8
    d2h = 16

    # every collective is done: tear the process group down on ALL ranks together, then rank 0 alone continues with
    # the single-GPU profiling pass and the CPU baseline (nothing below uses torch.distributed)
    if world > 1:
        trainer.graph = None            # the captured graph holds NCCL work: drop it before leaving the group
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if rank != 0:
            # destroy_process_group() was seen to block with a captured NCCL graph on this stack (torch 2.11 / NCCL
            # 2.28): leave without running the destructors; every collective has completed at the barrier above.
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)
    if rank != 0:
        return

    print("[bench] rank 0: collective phases done, profiling pass", file=sys.stderr, flush=True)
    # ---- per-kernel durations (CUDA events on the launching stream) for the roofline; separate pass ---------------
    prof_steps = min(args.steps, 50)
    _lib.profile_enable(True)
    for i in range(prof_steps):
        model.accumulate_gradients_images(images[i % n_batches], labels)
        model.apply_gradients(ETA, B * world)
    torch.cuda.synchronize()
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    pk = peaks()
    fwd, bwd_d, bwd_w = dense_flops_per_image(shapes)
    alg = {  # algorithmic work per launch of each kernel
        "features_fused_kernel": ("hbm", B * (H * W * 1 + L * 8)),
        "dense_forward_gemm": ("tensor", B * fwd / len(shapes)),
        "dense_backward_data_gemm": ("tensor", B * bwd_d / max(1, len(shapes) - 1)),
        "dense_backward_weight_gemm": ("tensor", B * bwd_w / len(shapes)),
        "sgd_update_kernel": ("hbm", 3 * n_params * 8),
    }
    kernels = {k: {"launches_per_step": v["launches"] / prof_steps, "avg_us": v["total_ms"] / v["launches"] * 1e3,
                   "share": None} for k, v in prof.items()}
    tot = sum(v["total_ms"] for v in prof.values())
    for k, v in prof.items():
        kernels[k]["share"] = round(v["total_ms"] / tot, 4)
    top = max(prof, key=lambda k: prof[k]["total_ms"])
    fp64_peak = measure_fp64_peak(torch, dev)
    bound, work = alg.get(top, ("hbm", 0))
    dur_s = prof[top]["total_ms"] / prof[top]["launches"] * 1e-3
    if bound == "hbm":
        achieved, peak, unit = work / dur_s / 1e9, pk["hbm_gbs"], "GB/s"
    else:
        achieved, peak, unit = work / dur_s / 1e12, fp64_peak, "TFLOP/s"
    roofline = {"kernel": top, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                "frac": achieved / peak if peak else None, "traffic": None,
                "peak_source": pk["source"] if bound == "hbm" else "torch.matmul f64 8192^3 measured in this run (f64 DMMA path; "
                                                                   "MEASURED_PEAKS.json has no f64 figure)",
                "precision": "f64", "fp64_dgemm_tflops_measured": fp64_peak, "kernels": kernels}

    print("[bench] rank 0: cpu baseline", file=sys.stderr, flush=True)
    # ---- CPU baseline on this box's host cores (bounded sample, ~10-20 s) ------------------------------------------
    cpu_ips, cpu_ms, cpu_done, cores = cpu_train_steps(wl, steps=10 ** 6, warmup=2, max_seconds=12.0)

    line = {
        "metric": "training images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "batch_per_gpu": B, "global_batch": B * world, "pixels": "u8", "eta": ETA,
                   "params": n_params, "parallelism": f"dp{world}",
                   "l2_policy": f"inputs rotate over {n_batches} resident batches = {n_batches * B * H * W / 2 ** 20:.0f} MiB > 126 MB L2",
                   "step": trainer.describe(), "cuda_graph": not args.no_graph},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "host_buffers": "pinned"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{cpu_done} steps of batch {B} (features+fwd+bwd+SGD) on {cores} host threads, "
                                   "oracle/rcn_oracle.cpp (C++ restatement of rcn's CPU path, not rustc output)"},
    }
    emit_result(line)
    if world > 1:
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    # stdout must carry exactly ONE JSON line: route everything else that writes to fd 1 (e.g. NCCL's version banner)
    # to stderr and keep the real stdout for the result line.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return
    if world != args.gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    run_gpu(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
