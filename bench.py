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
import math
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
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).

    NVML is polled in-process every few ms (a timed region of this workload can be shorter than one `nvidia-smi -lms`
    period); samples carry a host timestamp and only those inside [mark_begin, mark_end] are reported. Falls back to an
    `nvidia-smi -lms 100` subprocess when pynvml is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, period_s=0.004):
        self.gpu, self.period = gpu_index, period_s
        self.samples, self.windows = [], []
        self.stop_flag = False
        self.thread = None
        self.max_mhz = None
        self.mode = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if self.gpu < len(ids) and ids[self.gpu].isdigit():
                    idx = int(ids[self.gpu])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
        except Exception:
            self.mode = "nvidia-smi"
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                              "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._poll_smi, daemon=True)
                self.thread.start()
            except Exception:
                self.mode = None

    def _poll_nvml(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), sm, mask))
            except Exception:
                pass
            time.sleep(self.period)

    def _poll_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            try:
                sm, mx = float(f[0]), float(f[1])
            except (ValueError, IndexError):
                continue
            self.max_mhz = mx
            mask = 0
            for bit, v in zip([0x8, 0x40, 0x20, 0x4], f[2:6]):
                if v.lower().startswith("active"):
                    mask |= bit
            self.samples.append((time.perf_counter(), sm, mask))

    def mark_begin(self):
        self._t0 = time.perf_counter()

    def mark_end(self):
        self.windows.append((self._t0, time.perf_counter()))

    def stop(self):
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        time.sleep(0.02 if self.mode == "nvml" else 0.15)
        self.stop_flag = True
        if self.mode == "nvidia-smi":
            self.proc.terminate()
        inside = [x for x in self.samples if any(a <= x[0] <= b for a, b in self.windows)]
        note = "samples inside the timed regions"
        if not inside and self.windows:      # region shorter than one sampling period: nearest samples under the same load
            a, b = self.windows[0][0], self.windows[-1][1]
            inside = [x for x in self.samples if a - 0.25 <= x[0] <= b + 0.05]
            note = "timed region shorter than a sampling period: samples within 250 ms before it (warm-up, same load)"
        mask = 0
        for x in inside:
            mask |= x[2]
        sm = [x[1] for x in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(name for bit, name in self.REASONS.items() if mask & bit), "samples": len(sm),
                "source": self.mode, "note": note}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's threaded train step (C++ restatement of rcn's CPU path -- NOT rustc output)
# ---------------------------------------------------------------------------------------------------------------------
def cpu_sample_batch(wl, cores):
    """Bounded CPU sample: the full minibatch when one CPU step is small, else a sub-batch of the same workload sized
    to ~2e10 flop per step (the per-sample cost of the reference algorithm does not depend on the batch size)."""
    shapes = layer_shapes(wl)
    fwd, bwd_d, bwd_w = dense_flops_per_image(shapes)
    per_image = fwd + bwd_d + bwd_w + 12 * 4 * wl["H"] * wl["W"]
    b = int(2e10 // per_image)
    b = max(4 * cores, b - b % max(1, cores))
    return min(wl["batch"], b)


def cpu_train_steps(wl, steps, warmup, max_seconds=None, batch=None):
    import oracle as O
    B, H, W = batch or wl["batch"], wl["H"], wl["W"]
    shapes = layer_shapes(wl)
    net = O.Net(shapes)
    rng = np.random.default_rng(DATA_SEED)
    n_pool = 4
    images = rng.integers(0, 256, size=(n_pool, B, H, W), dtype=np.uint8)
    labels = (np.arange(B) % wl["classes"]).astype(np.int64)
    params = np.random.default_rng(PARAM_SEED).standard_normal(net.n_params)
    mean, sd = O.gen_scales(O.features_u8(wl["cfg"], images[0][:min(B, 256)]))
    cores = os.cpu_count() or 1
    t_w = time.perf_counter()
    for i in range(warmup):
        net.train_step_u8(wl["cfg"], params, images[i % n_pool], labels, mean, sd, ETA, cores)
        if max_seconds and time.perf_counter() - t_w > max_seconds / 2:
            break
    t0 = time.perf_counter()
    done = 0
    for i in range(steps):
        net.train_step_u8(wl["cfg"], params, images[i % n_pool], labels, mean, sd, ETA, cores)
        done += 1
        if max_seconds and time.perf_counter() - t0 > max_seconds:
            break
    dt = time.perf_counter() - t0
    return done * B / dt, dt / done * 1e3, done, cores, B


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    # bounded sample: each step is one (sub-)minibatch of the same workload on all host threads; cap the run at ~2 min
    cores = os.cpu_count() or 1
    Bs = cpu_sample_batch(wl, cores)
    ips, ms, done, cores, Bs = cpu_train_steps(wl, args.steps, args.warmup, max_seconds=100.0, batch=Bs)
    sample = (f"{done} steps of batch {Bs} (features+fwd+bwd+SGD) on {cores} host threads" +
              ("" if Bs == wl["batch"] else f"; sub-batch of the {wl['batch']}-image minibatch, per-sample cost is batch-independent"))
    line = {
        "impl": "reference", "metric": "training images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "batch_per_step": Bs, "pixels": "u8", "eta": ETA,
                   "note": "C++ restatement of rcn's CPU path (oracle/rcn_oracle.cpp: per-sample matvec backprop, worker threads "
                           "+ mutex-ordered gradient sum as rcn.rs:176-223); the Rust crate cannot be built here (no rustc)"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_result(line)


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
_WC_KEEP = []   # write-combined allocations stay alive for the life of the process


def _wc_pinned_copy(a):
    """numpy array -> the same bytes in cudaHostAllocWriteCombined memory (viewed as a numpy array), or None if the CUDA
    runtime cannot be reached from here. CPU reads of such memory are slow; the bench only writes it once."""
    import ctypes
    try:
        import torch
        rt = ctypes.CDLL(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so.12"))
    except OSError:
        try:
            rt = ctypes.CDLL("libcudart.so.12")
        except OSError:
            return None
    p = ctypes.c_void_p()
    rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    if rt.cudaHostAlloc(ctypes.byref(p), a.nbytes, 0x01 | 0x04) != 0 or not p.value:   # portable | write-combined
        return None
    buf = (ctypes.c_uint8 * a.nbytes).from_address(p.value)
    out = np.frombuffer(buf, dtype=a.dtype).reshape(a.shape)
    out[...] = a
    _WC_KEEP.append((rt, p, buf))
    return out


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
    trainer = DataParallelTrainer(model, eta=ETA, exchange=args.exchange)

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
    step_desc = trainer.describe()
    kernels_per_step = None
    if not args.no_graph:
        l_before = _lib.kernel_launches()
        # the timed region is a whole number of replays: the largest divisor of K that is <= --steps-per-graph
        if world == 1:
            spg = max(d for d in range(1, max(1, min(args.steps_per_graph, args.steps)) + 1) if args.steps % d == 0)
        else:   # multi-GPU runs keep the graph shapes they were validated with on this pool (gcd with at most 8)
            spg = max(1, math.gcd(args.steps, min(args.steps_per_graph, 8)))
        trainer.capture(warmup=3, steps_per_graph=spg)
        kernels_per_step = (_lib.kernel_launches() - l_before) // (3 + spg)   # 3 warm-up steps + spg captured steps
    else:
        spg = 1

    def step(i):                                # warm-up unit: one graph replay (= spg steps) or one eager step
        trainer.epoch_step()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # warm-up: at least W steps and at least ~0.3 s of the same load, so clocks have ramped before the timed region
    n_warm = max(args.warmup, 3)
    t_w = time.perf_counter()
    for i in range(n_warm):
        step(i)
    torch.cuda.synchronize()
    per_step = max((time.perf_counter() - t_w) / n_warm, 1e-6)
    extra = torch.tensor([min(200000, int(0.3 / per_step))], dtype=torch.int64, device=dev)
    if world > 1:
        dist.broadcast(extra, 0)            # every rank must run the same number of (all-reducing) steps
    for i in range(int(extra.item())):
        step(n_warm + i)
    n_warm += int(extra.item())
    barrier()
    l0 = _lib.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.mark_begin()
    e0.record(stream)
    trainer.epoch_steps(args.steps)             # exactly K steps: K / spg replays of the spg-step graph
    e1.record(stream)
    barrier()
    clocks.mark_end()
    launches = _lib.kernel_launches() - l0
    if kernels_per_step is not None and launches == 0:
        launches = kernels_per_step * args.steps   # graph replays re-launch the captured kernels (the counter only sees
                                                   # launches made through the library, e.g. the persistent step kernel)
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- e2e: the reference-facing call with HOST buffers (pinned): the chunks_exact loop of rcn.rs:147-149 over a
    # host-resident dataset; every step's H2D copy and the D2H read of its (cost, hits) are inside the timed region -----
    n_host = max(2, min(args.steps, (512 << 20) // (B * H * W)))
    h_images = torch.randint(0, 256, (n_host * B, H, W), dtype=torch.uint8).pin_memory()
    h_labels = (torch.arange(n_host * B) % wl["classes"]).to(torch.int64).pin_memory()
    hi, hl = h_images.numpy(), h_labels.numpy()
    host_alloc = "pinned"
    if args.host_alloc == "wc":
        # opt-in experiment: the same dataset in WRITE-COMBINED pinned memory (cudaHostAllocWriteCombined). The e2e path is
        # bound by the rate at which the GPU can read host memory (38 GB/s from ordinary pinned pages);
        # write-combined pages are not snooped in the CPU caches on the way out. Not the default until it is measured.
        wc = _wc_pinned_copy(hi)
        if wc is not None:
            hi, host_alloc = wc, "pinned, write-combined"

    def e2e_steps(n):
        done = 0
        while done < n:
            m = min(n_host, n - done)
            cost, hits = trainer.train_epoch_host(hi[:m * B], hl[:m * B], B)
            assert len(cost) == m
            done += m

    e2e_steps(min(n_host, args.steps))          # warm-up at the timed call's size: buffers sized, step graphs captured
    barrier()
    clocks.mark_begin()
    e0.record(stream)
    e2e_steps(args.steps)
    e1.record(stream)
    barrier()
    clocks.mark_end()
    clk = clocks.stop() if rank == 0 else None
    e2e_ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * B / (e2e_ms / args.steps * 1e-3)
    h2d = B * H * W + B * 8
    d2h = 16

    # ---- per-kernel durations (CUDA events on the launching stream) for the roofline; separate eager pass that EVERY
    # rank runs (the exchange kernel / all-reduce needs all of them); rank 0 reports its own records ------------------
    prof_steps = min(args.steps, 50)
    barrier()
    _lib.profile_enable(True)
    for i in range(prof_steps):
        trainer.step_images(images[i % n_batches], labels)
    torch.cuda.synchronize()
    prof = _lib.profile_report()
    _lib.profile_enable(False)
    barrier()

    # every collective is done: tear the process group down on ALL ranks together, then rank 0 alone continues with
    # the single-GPU profiling pass and the CPU baseline (nothing below uses torch.distributed)
    if world > 1:
        trainer.graph = None            # the captured graph holds NCCL work: drop it before leaving the group
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if trainer.p2p:                 # rank 0's profiling pass below must not wait for peers that have left
            model.dp_shutdown()
            trainer.p2p = False
            dist.barrier()
        if rank != 0:
            # destroy_process_group() was seen to block with a captured NCCL graph on this stack (torch 2.11 / NCCL
            # 2.28): leave without running the destructors; every collective has completed at the barrier above.
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(0)
    if rank != 0:
        return

    print("[bench] rank 0: collective phases done", file=sys.stderr, flush=True)
    pk = peaks()
    fwd, bwd_d, bwd_w = dense_flops_per_image(shapes)
    sum_rows = sum(r for r, _ in shapes)
    act_bytes = 2 * sum_rows * 8                      # a_l and delta_l written once per sample
    int8_peak = 2.0 * (pk["bf16_tflops"] or 1590.0)   # kind::i8 issues at twice the bf16 rate; bf16 is the measured figure

    def on_tc(M, N, K):                               # mirrors use_tensor_cores() in csrc/dense.cu
        tiles = -(-M // 128) * -(-N // 96)
        return M >= 256 and N >= 256 and K >= 512 and K <= 16384 and tiles >= 74 and float(M) * N * K >= 4.0e9

    # algorithmic work per step of each kernel NAME: name -> [bound, bytes, flops]; the per-launch
    # figure is this divided by the launches per step the profiler counted
    alg = {
        "features_fused_kernel": ["hbm", B * (H * W * 1 + L * 8), B * 48 * H * W],
        "smallnet_fwd_bwd_kernel(fused features)": ["hbm", B * (H * W * 1 + 8 + L * 8 + act_bytes) + n_params * 8,
                                                    B * (48 * H * W + fwd + bwd_d)],
        "smallnet_fwd_bwd_kernel": ["hbm", B * (L * 8 + 8 + act_bytes) + n_params * 8, B * (fwd + bwd_d)],
        "smallnet_wgrad_kernel": ["hbm", B * (L + shapes[0][0]) * 8 + n_params * 8, B * bwd_w],
        "smallnet_wgrad_kernel(+SGD update)": ["hbm", B * (L + shapes[0][0]) * 8 + 3 * n_params * 8, B * bwd_w + 2 * n_params],
        "smallnet_wgrad_kernel(+exchange+SGD update)": ["hbm", B * (L + shapes[0][0]) * 8 + 3 * n_params * 8, B * bwd_w + 2 * n_params],
        "features_cp_kernel": ["hbm", B * (H * W * 1 + L * 8), B * 48 * H * W],
        "sgd_update_kernel": ["hbm", 3 * n_params * 8, 2 * n_params],
        "dp_allreduce_sgd_kernel": ["hbm", 3 * n_params * 8, 2 * n_params],
        "bias_grad_kernel": ["hbm", B * sum_rows * 8, B * sum_rows],
        # skinny output layer (rows <= 16 behind a wide layer): each kernel streams the wide activation matrix once
        "skinny_forward_kernel": ["hbm", B * (shapes[-1][1] + 2 * shapes[-1][0]) * 8, 2 * B * shapes[-1][0] * shapes[-1][1]],
        "skinny_backward_data_kernel": ["hbm", B * (2 * shapes[-1][1] + shapes[-1][0]) * 8, 2 * B * shapes[-1][0] * shapes[-1][1]],
        "skinny_backward_weight_kernel": ["hbm", B * (shapes[-1][1] + shapes[-1][0]) * 8, 2 * B * shapes[-1][0] * shapes[-1][1]],
    }
    for tag in ("forward", "backward_data", "backward_weight"):
        for suffix in ("", "(tcgen05 int8 slices)"):
            alg[f"dense_{tag}_gemm{suffix}"] = ["tensor", 0, 0]
    slice_bytes = 0
    for li, (r, c) in enumerate(shapes):
        f = 2.0 * r * c * B
        tc_f, tc_w = on_tc(r, B, c), on_tc(r, c, B)
        alg["dense_forward_gemm" + ("(tcgen05 int8 slices)" if tc_f else "")][2] += f
        alg["dense_backward_weight_gemm" + ("(tcgen05 int8 slices)" if tc_w else "")][2] += f
        slice_bytes += (13 * (r * c + c * B) if tc_f else 0) + (13 * (r * B + c * B) if tc_w else 0)
        if li >= 1:
            rp = shapes[li - 1][0]                    # backward-data into layer li-1: M = rp, K = r, N = B
            tc_d = on_tc(rp, B, r)
            alg["dense_backward_data_gemm" + ("(tcgen05 int8 slices)" if tc_d else "")][2] += 2.0 * rp * r * B
            slice_bytes += 13 * (rp * r + r * B) if tc_d else 0
    alg["ozaki_slice"] = ["hbm", slice_bytes, 0]      # both slicing kernels together: 8 B read + 5 B written per element
    kernels = {k: {"launches_per_step": v["launches"] / prof_steps, "avg_us": v["total_ms"] / v["launches"] * 1e3,
                   "share": None} for k, v in prof.items()}
    tot = sum(v["total_ms"] for v in prof.values())
    fp64_peak = measure_fp64_peak(torch, dev)
    for k, v in prof.items():
        kernels[k]["share"] = round(v["total_ms"] / tot, 4)
        key = "ozaki_slice" if k.startswith("ozaki_slice") else k
        if key in alg:
            d_s = v["total_ms"] / prof_steps * 1e-3          # seconds of this kernel name per step
            if key == "ozaki_slice":
                d_s = sum(x["total_ms"] for n_, x in prof.items() if n_.startswith("ozaki_slice")) / prof_steps * 1e-3
            bnd, nbytes, nflops = alg[key]
            kernels[k]["bound"] = bnd
            if nbytes:
                kernels[k]["GBps"] = round(nbytes / d_s / 1e9, 1)
                kernels[k]["frac_hbm"] = round(nbytes / d_s / 1e9 / pk["hbm_gbs"], 4)
            if nflops and "tcgen05" in k:
                kernels[k]["f64_equiv_TFLOPs"] = round(nflops / d_s / 1e12, 2)
                kernels[k]["int8_TOPs"] = round(15 * nflops / d_s / 1e12, 1)     # 15 digit-plane products per f64 product
                kernels[k]["frac_int8_tensor"] = round(15 * nflops / d_s / 1e12 / int8_peak, 4)
            elif nflops:
                kernels[k]["TFLOPs_f64"] = round(nflops / d_s / 1e12, 3)
                kernels[k]["frac_fp64"] = round(nflops / d_s / 1e12 / fp64_peak, 4)
    top = max(prof, key=lambda k: prof[k]["total_ms"])
    bound, nbytes, nflops = alg.get(top, ("hbm", 0, 0))
    dur_s = prof[top]["total_ms"] / prof_steps * 1e-3        # all launches of the top kernel in one step
    peak_source = pk["source"]
    if bound == "hbm":
        achieved, peak, unit = nbytes / dur_s / 1e9, pk["hbm_gbs"], "GB/s"
    elif "tcgen05" in top:
        achieved, peak, unit = 15 * nflops / dur_s / 1e12, int8_peak, "TFLOP/s"
        peak_source = ("int8 tensor peak taken as 2 x the measured bf16 cuBLAS figure of MEASURED_PEAKS.json (kind::i8 issues at twice "
                       "the bf16 rate; nominal 4500); achieved = 15 exact int8 digit-plane products per f64 product, in int8 TOP/s")
    else:
        achieved, peak, unit = nflops / dur_s / 1e12, fp64_peak, "TFLOP/s"
        peak_source = "torch.matmul f64 8192^3 measured in this run (f64 DMMA path; MEASURED_PEAKS.json has no f64 figure)"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed `ncu --set full` capture
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get(args.workload, {}).get(top)
    step_bytes = B * (H * W + 8) + 3 * n_params * 8           # compulsory HBM traffic of a whole step (intermediates on chip / in L2)
    step_flops = B * (48 * H * W + fwd + bwd_d + bwd_w)
    roofline = {"kernel": top, "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                "frac": achieved / peak if peak else None, "traffic": traffic,
                "peak_source": peak_source,
                "precision": "f64", "fp64_dgemm_tflops_measured": fp64_peak,
                "durations": "CUDA events around every launch on the launching stream, separate eager pass of "
                             f"{prof_steps} steps (the timed region replays a CUDA graph)",
                "whole_step": {"ms": ms_step, "compulsory_GBps": step_bytes / (ms_step * 1e-3) / 1e9,
                               "TFLOPs_f64": step_flops / (ms_step * 1e-3) / 1e12,
                               "frac_fp64": step_flops / (ms_step * 1e-3) / 1e12 / fp64_peak},
                "kernels": kernels}

    print("[bench] rank 0: cpu baseline", file=sys.stderr, flush=True)
    # ---- CPU baseline on this box's host cores (bounded sample, ~10-20 s) ------------------------------------------
    Bs = cpu_sample_batch(wl, os.cpu_count() or 1)
    cpu_ips, cpu_ms, cpu_done, cores, Bs = cpu_train_steps(wl, steps=10 ** 6, warmup=2, max_seconds=12.0, batch=Bs)
    # BASELINE.json configs[0] (the reference's own CPU-runnable case: the same network at batch 32), ~3 s more
    c1_ips, _, c1_done, _, _ = cpu_train_steps(WORKLOADS["c1"], steps=10 ** 6, warmup=2, max_seconds=3.0)

    line = {
        "metric": "training images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "batch_per_gpu": B, "global_batch": B * world, "pixels": "u8", "eta": ETA,
                   "params": n_params, "parallelism": f"dp{world}",
                   "l2_policy": f"inputs rotate over {n_batches} resident batches = {n_batches * B * H * W / 2 ** 20:.0f} MiB > 126 MB L2",
                   "step": step_desc, "cuda_graph": not args.no_graph, "steps_per_graph": spg,
                   # the W requested warm-up steps plus ~0.3 s of the same load so that clocks have ramped
                   "warmup_steps_run": n_warm * spg},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps, "host_buffers": host_alloc,
                # PCIe bytes per second this rank pulled inside the timed region (the e2e path's own bound: SM-issued
                # zero-copy reads reach ~38 GB/s on this pool's boxes, profiles/r1s_full_host_prefetch.txt)
                "h2d_GBps_per_gpu": h2d / (e2e_ms / args.steps * 1e-3) * 1e-9,
                "api": "rcn_cuda_train_epoch_host (chunks_exact loop over a pinned host dataset: the GPU pulls chunk k+1 over PCIe "
                       "while chunk k trains, one CUDA graph launch per step, per-step cost/hits written back to host memory)"},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": f"{cpu_done} steps of batch {Bs} (features+fwd+bwd+SGD) on {cores} host threads, "
                                   "oracle/rcn_oracle.cpp (C++ restatement of rcn's CPU path, not rustc output)",
                         "configs0_batch32": {"value": c1_ips, "unit": "images/s", "cores": cores,
                                              "sample": f"{c1_done} steps of batch 32, same network, same {cores} host threads"}},
    }
    emit_result(line)
    if world > 1:
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--steps-per-graph", type=int, default=16, help="consecutive steps captured into one CUDA graph (the timed region replays it steps/this times)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying a CUDA graph")
    ap.add_argument("--host-alloc", default="pinned", choices=["pinned", "wc"],
                    help="e2e leg: ordinary pinned host memory (default) or write-combined pinned memory (experiment)")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU gradient exchange: fused NVLink peer-memory kernel, NCCL all-reduce, or by size")
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
