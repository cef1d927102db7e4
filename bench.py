#!/usr/bin/env python
"""Headline benchmark: R(2+1)D-18 training clips/s (16x112x112 clips, bf16 compute) on N x B200.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 4 --warmup 1      # CPU arm (oracle port of the reference)

A step is exactly the reference's training iteration (main.py:170-207) on one synthetic batch of 22 clips per
GPU: zero_grad -> model(X) -> MSELoss -> train-time nearest-class accuracy (main.py:182-185, here on the GPU) ->
backward -> Adam step.  `value` is timed with the batch already resident in HBM; `e2e` repeats the measurement
through the same public API with pinned HOST buffers (H2D of the clip batch and targets and D2H of the loss
inside the timed region).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# NCCL's own log (NCCL_DEBUG=INFO/VERSION) is left alone: _capture_stdout() routes everything libraries print on fd 1
# to stderr, so stdout still carries exactly one JSON line.

_REAL_STDOUT = None


def _capture_stdout():
    """Route fd 1 to stderr while libraries (NCCL banners, warnings) may print; the JSON line goes to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj) -> None:
    line = json.dumps(obj) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(line)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, line.encode())


METRIC = "R(2+1)D-18 train clips/s (16x112^2, bf16)"
FLOP_PER_CLIP = 242.449e9          # fwd + dgrad + wgrad, SURVEY.md section 8(d) (no dgrad for stem.0)
N_TRAIN_CLASSES = 664              # train-time class table (Kinetics after the tau-filter, any C <= 700)


# ----------------------------------------------------------------------------------------------------
# helpers
# ----------------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(tflops_sustained=p.get("bf16_tflops_sustained", 1400.0), tflops_burst=p.get("bf16_tflops", 1590.0),
                    hbm_gbs=p.get("hbm_gbs", 6650.0), source="MEASURED_PEAKS.json (of measured)")
    return dict(tflops_sustained=1400.0, tflops_burst=1590.0, hbm_gbs=6650.0,
                source="B200_PROFILING.md fallback (of fallback)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while a timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def conv_flops(op) -> float:
    """Algorithmic FLOPs of one convolution pass (fprop, dgrad or wgrad): 2*M*N*K with unpadded channel counts."""
    k = op.kernel[0] * op.kernel[1] * op.kernel[2]
    return 2.0 * op.out_positions * op.cout * op.cin * k


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's training step on the host cores
# ----------------------------------------------------------------------------------------------------
def cpu_reference_steps(steps: int, warmup: int, bs: int = 2):
    """main.py:170-207 restated on the CPU oracle (fp32, torch CPU kernels on all host threads): forward,
    MSELoss, nearest-class accuracy, backward, Adam.  Returns (clips_per_s, seconds_per_step list, threads)."""
    import numpy as np
    import torch
    import torch.nn.functional as F
    from oracle import nearest_oracle, video_oracle as vo

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = vo.synthetic_state_dict_r2plus1d(0)
    params = {k: v.requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))}
    optimizer = torch.optim.Adam(list(params.values()), lr=1e-3)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(bs, 1, 3, 16, 112, 112, generator=g)
    cls = F.normalize(torch.randn(N_TRAIN_CLASSES, 300, generator=g))
    labels = torch.randint(0, N_TRAIN_CLASSES, (bs,), generator=g)
    z = cls[labels]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        optimizer.zero_grad()
        emb = vo.model_forward(sd, x, train=True)
        loss = vo.mse_loss(emb, z)
        pred = nearest_oracle.nearest_class(emb.detach().numpy(), cls.numpy(), 1)[:, 0]
        _acc = float(np.mean(pred == labels.numpy()))
        loss.backward()
        optimizer.step()
        float(loss.detach())
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return bs * len(times) / total, times, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    bs = 2
    cps, times, threads = cpu_reference_steps(args.steps, args.warmup, bs)
    ms = 1e3 * sum(times) / len(times)
    sample = (f"{len(times)} steps x {bs} clips (BASELINE.json config 1: bs=2x3x16x112x112, fp32) of the same training "
              f"step on the host cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": cps, "unit": "clips/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "R(2+1)D-18 training step, 16x112x112 clips, bs=22/GPU (CPU arm: bounded sample, bs=2 per step)",
                   "network": "r2plus1d_18", "per_gpu_batch": 22, "cpu_step_batch": bs},
        "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "oracle/ port of the reference path (the Python reference cannot travel to the GPU box); "
                "torch CPU kernels, all host threads",
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------
# library arm: the reference's iteration on stock PyTorch (cuDNN / cuBLAS) on the same GPU -- the path to beat
# ----------------------------------------------------------------------------------------------------
def library_baseline(network: str, batch: int, steps: int, warmup: int, device):
    from oracle import library_arm
    return library_arm.run(network, batch, steps, warmup, device)


def run_library(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.cuda.set_device(0)
    res = library_baseline(args.network, args.batch, args.steps, args.warmup, torch.device("cuda", 0))
    best = res["variants"].get(res["best"], {}) if res["best"] else {}
    emit({
        "impl": "library", "metric": METRIC if "2plus1d" in args.network else f"{args.network} train clips/s (16x112^2, bf16)",
        "value": res["clips_per_s"], "unit": "clips/s", "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": best.get("ms_per_step"), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if res["best"] and "bf16" in res["best"] else "fp16", "data": "synthetic",
        "config": {"workload": f"{args.network} end-to-end training step bs={args.batch}, 16x112x112 synthetic clips, stock "
                               f"PyTorch modules through cuDNN/cuBLAS (best of two variants: {res['best']})",
                   "network": args.network, "per_gpu_batch": args.batch},
        "library_baseline": res, "gpu_launches": 0,
    })


# ----------------------------------------------------------------------------------------------------
# zero-shot evaluation workload (BASELINE.json configs[4]): 10k clip embeddings vs 101 / 51 / 200 class tables
# ----------------------------------------------------------------------------------------------------
def nearest_workload(dev, steps: int = 20, warmup: int = 3, n_rows: int = 10_000, k: int = 5):
    """main.py:316-325's cdist + argsort[:, :5] for N = 10 000 embeddings against C in {101, 51, 200} class vectors
    (UCF101 / HMDB51 / ActivityNet tables), D = 300.  Per C: device time of zsv_nearest_class with the inputs resident
    (CUDA events around one call, L2 flushed before every call), the same call end to end from pinned HOST arrays
    (H2D of both tables and D2H of the int64 indices inside the timed region), scipy on one host core (the
    reference's own call), and the parity of the result (top-1 indices and top-5 sets) against it."""
    import numpy as np
    import torch
    from scipy.spatial.distance import cdist
    from zeroshotvideoclassification_b200 import _lib, ops

    peaks = load_peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # 2x the 126 MB L2
    rng = np.random.default_rng(5)
    out = {"rows": n_rows, "dim": 300, "k": k, "l2": "flushed (256 MB write) before every timed call", "tables": {}}
    # what the FP64 units of this GPU deliver: cuBLAS DGEMM 4096^3, measured here (the kernel issues one DFMA per
    # (row, class, k) in scipy's order, so its bound is the FP64 pipe, not HBM)
    a64 = torch.randn(4096, 4096, dtype=torch.float64, device=dev)
    for _ in range(2):
        torch.mm(a64, a64)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(3):
        torch.mm(a64, a64)
    g1.record()
    torch.cuda.synchronize()
    fp64_peak = 3 * 2.0 * 4096 ** 3 / (g0.elapsed_time(g1) * 1e-3) / 1e12
    out["fp64_peak_tflops"] = fp64_peak
    out["fp64_peak_source"] = "cuBLAS DGEMM 4096^3 timed in this run"
    del a64
    launches0 = _lib.launch_count()
    for C in (101, 51, 200):
        emb = rng.standard_normal((n_rows, 300)).astype(np.float32)
        emb /= np.linalg.norm(emb, axis=1, keepdims=True)
        cls = rng.standard_normal((C, 300)).astype(np.float32)
        cls /= np.linalg.norm(cls, axis=1, keepdims=True)
        emb_h, cls_h = torch.from_numpy(emb).pin_memory(), torch.from_numpy(cls).pin_memory()
        idx_h = torch.empty((n_rows, k), dtype=torch.int64).pin_memory()
        emb_d, cls_d = emb_h.to(dev), cls_h.to(dev)
        dev_us, e2e_us = [], []
        for i in range(warmup + steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            idx = ops.nearest_class(emb_d, cls_d, k)
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                dev_us.append(1e3 * e0.elapsed_time(e1))
        for i in range(warmup + steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            idx = ops.nearest_class(emb_h.to(dev, non_blocking=True), cls_h.to(dev, non_blocking=True), k)
            idx_h.copy_(idx, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            if i >= warmup:
                e2e_us.append(1e3 * e0.elapsed_time(e1))
        t0 = time.perf_counter()
        d = cdist(emb, cls, "cosine")                      # main.py:321
        top = d.argsort(1)[:, :k]                          # main.py:322-324
        cpu_s = time.perf_counter() - t0
        got = idx_h.numpy()
        us, eus = statistics.median(dev_us), statistics.median(e2e_us)
        alg_bytes = (n_rows + C) * 300 * 4 + n_rows * k * 8
        out["tables"][str(C)] = {
            "device_us": us, "device_us_best": min(dev_us), "gb_per_s": alg_bytes / us / 1e3,
            "frac_of_hbm_peak": alg_bytes / us / 1e3 / peaks["hbm_gbs"], "algorithmic_bytes": alg_bytes,
            "fp64_gflops": 2.0 * n_rows * C * 300 / us / 1e3,
            "frac_of_fp64_peak": 2.0 * n_rows * C * 300 / us / 1e6 / fp64_peak,
            "e2e_us": eus, "h2d_bytes": (n_rows + C) * 300 * 4, "d2h_bytes": n_rows * k * 8,
            "cpu_scipy_s": cpu_s, "cpu_cores": 1, "speedup_e2e_vs_scipy": cpu_s * 1e6 / eus,
            "top1_bit_equal": bool(np.array_equal(got[:, 0], d.argmin(1))),
            "top5_sets_equal": bool(all(set(a) == set(b) for a, b in zip(got.tolist(), top.tolist()))),
        }
    out["gpu_launches"] = _lib.launch_count() - launches0
    del flush
    return out


def run_nearest(args):
    import torch
    from zeroshotvideoclassification_b200 import _lib, build
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    build.build()
    _lib.load()
    torch.cuda.set_device(0)
    res = nearest_workload(torch.device("cuda", 0), args.steps, max(args.warmup, 3))
    t200 = res["tables"]["200"]
    total_us = sum(t["device_us"] for t in res["tables"].values())
    total_e2e = sum(t["e2e_us"] for t in res["tables"].values())
    total_cpu = sum(t["cpu_scipy_s"] for t in res["tables"].values())
    emit({
        "metric": "zero-shot nearest-class search, 10k embeddings x {101,51,200} classes, top-5 (rows/s)",
        "value": 3 * res["rows"] / (total_us / 1e6), "unit": "rows/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_us / 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE.json configs[4]: 10k x 300 unit-norm embeddings vs 101/51/200-class tables, "
                               "cosine distance (scipy cdist order, fp64) + 5 smallest per row", "l2": res["l2"]},
        "roofline": {"bound": "hbm", "achieved": t200["gb_per_s"], "peak": load_peaks()["hbm_gbs"], "unit": "GB/s",
                     "frac": t200["frac_of_hbm_peak"], "traffic": None,
                     "note": "C=200 table; (N+C)*300*4 + N*5*8 algorithmic bytes; the kernel is bound by the FP64 pipe "
                             "(N*C*300 DFMAs in scipy's summation order), not by HBM: see fp64",
                     "fp64": {"achieved_tflops": t200["fp64_gflops"] / 1e3, "peak_tflops": res["fp64_peak_tflops"],
                              "frac": t200["frac_of_fp64_peak"], "peak_source": res["fp64_peak_source"]}},
        "cpu_baseline": {"value": 3 * res["rows"] / total_cpu, "unit": "rows/s", "cores": 1, "kind": "reference",
                         "sample": "scipy cdist(...,'cosine') + argsort, the reference's own call (main.py:321-324), all three tables"},
        "e2e": {"value": 3 * res["rows"] / (total_e2e / 1e6), "unit": "rows/s",
                "h2d_bytes_per_step": sum(t["h2d_bytes"] for t in res["tables"].values()),
                "d2h_bytes_per_step": sum(t["d2h_bytes"] for t in res["tables"].values())},
        "gpu_launches": res["gpu_launches"], "nearest": res,
    })


def subprocess_line(extra_args, timeout=600):
    """Run this script again with other arguments (another network / workload) and return its JSON line."""
    cmd = [sys.executable, os.path.abspath(__file__)] + extra_args
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    try:
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": f"rc {r.returncode}: {r.stderr[-300:]}"}
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    from zeroshotvideoclassification_b200 import _lib, build, default_opt, get_network, ops
    from zeroshotvideoclassification_b200 import dist as zdist
    from zeroshotvideoclassification_b200.accuracy import nearest_class
    from zeroshotvideoclassification_b200.graph import GraphedStep

    build.build()
    _lib.load()
    rank, local_rank, world = zdist.init_from_env("nccl")
    if world != args.gpus:
        if rank == 0:
            sys.stderr.write(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE\n")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    B = args.batch
    torch.manual_seed(0)
    model = get_network(default_opt(args.network)).to(dev).train()
    zdist.broadcast_module(model)
    criterion = torch.nn.MSELoss().to(dev)
    # main.py:131's torch.optim.Adam; capturable keeps its step counter on the device so the iteration can be graphed
    if args.optimizer == "zsv":
        # main.py:131's Adam on the C ABI: same update rule and state layout; the convolution weights' bf16 images for
        # the next iteration come out of the same pass (zsv_adam_pack_step), step counters and lr live on the device
        from zeroshotvideoclassification_b200.optim import FusedAdam
        optimizer = FusedAdam(model.parameters(), lr=1e-3, model=model)
    else:
        optimizer = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=args.graph, fused=True)
    sync = zdist.GradSync(bucket_bytes=int(float(os.environ.get("ZSV_BUCKET_MB", "8")) * (1 << 20))) if world > 1 else None
    zdist.set_grad_sync(sync)
    # parameters whose gradients do not come out of the backbone Function (GradSync covers those): the MLP head of
    # network.Model; C3D has no bucketed sync, all of its gradients are reduced after backward
    head_params = list(model.output2emb_proj.parameters()) if hasattr(model, "output2emb_proj") else list(model.parameters())
    if sync is not None:
        sync.attach(head_params)          # all-reduced in place from a post-accumulate hook, during backward

    g = torch.Generator().manual_seed(1 + rank)
    x_host = torch.randn(B, 1, 3, 16, 112, 112, generator=g).pin_memory()
    cls = F.normalize(torch.randn(N_TRAIN_CLASSES, 300, generator=torch.Generator().manual_seed(7)))
    labels = torch.randint(0, N_TRAIN_CLASSES, (B,), generator=g)
    z_host = cls[labels].contiguous().pin_memory()
    x_dev = x_host.to(dev)
    z_dev = z_host.to(dev)
    cls_dev = cls.to(dev)
    labels_dev = labels.to(dev)
    acc_sum = torch.zeros((), device=dev)
    acc_stream = torch.cuda.Stream(device=dev)

    def step(X, Z):
        optimizer.zero_grad(set_to_none=True)
        out = model(X)
        emb = out[0] if isinstance(out, tuple) else out
        loss = criterion(emb, Z)
        # main.py:182-185 without the host round trip; the accuracy depends only on the embeddings and nothing depends
        # on it, so it runs on its own stream beside the backward pass (joined before the step ends)
        cur = torch.cuda.current_stream(dev)
        acc_stream.wait_stream(cur)
        with torch.cuda.stream(acc_stream):
            pred = nearest_class(emb.detach(), cls_dev, 1)[:, 0]
            acc_sum.add_((pred == labels_dev).float().mean())
        loss.backward()
        if sync is not None:
            sync.finish()                 # no-op when the backbone's backward already waited for everything
        optimizer.step()
        cur.wait_stream(acc_stream)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (eager) + capture of the iteration into one CUDA graph ----
    n_warm = args.warmup if args.quick else max(args.warmup, 3)
    for _ in range(n_warm):
        step(x_dev, z_dev)
    barrier()
    gstep = None
    graph_note = "eager (--no-graph)"
    if args.graph:
        try:
            gstep = GraphedStep(step, (x_dev, z_dev), device=dev, warmup=1 if args.quick else 2,
                                capture_error_mode=args.capture_mode)
            graph_note = "whole iteration (zero_grad..Adam) replayed as one CUDA graph"
        except Exception as exc:           # reported, never silent: the eager path below is the same kernels
            graph_note = f"eager: graph capture failed ({type(exc).__name__}: {str(exc)[:200]})"
            sys.stderr.write(graph_note + "\n")
            gstep = None
            torch.cuda.synchronize()
    if world > 1:
        ok = torch.tensor([1 if gstep is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and gstep is not None:
            gstep, graph_note = None, "eager: graph capture failed on another rank"

    def run_resident():
        if gstep is not None:
            return gstep(*gstep.static_inputs)     # inputs already in the captured HBM buffers: no copy
        return step(x_dev, z_dev)

    for _ in range(0 if args.quick else 2):
        run_resident()
    barrier()

    # ---- timed region 1: inputs resident in HBM ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    t_host0 = time.perf_counter()
    for _ in range(args.steps):
        loss = run_resident()
    host_ms_per_step = 1e3 * (time.perf_counter() - t_host0) / args.steps    # enqueue time, no synchronisation
    ev1.record()
    barrier()
    launches = _lib.launch_count() - launches0
    if gstep is not None:
        launches = gstep.launches_per_replay * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total / 1e3)
    final_loss = float(loss.detach())

    # ---- per-kernel timing: CUDA events cannot be recorded inside a graph replay, so the same iteration is run
    # kernel by kernel with an event pair (on the launching stream) around every convolution call ----
    prof = []
    prof_steps = max(1, min(args.steps, 5))
    from zeroshotvideoclassification_b200 import engine as zengine
    overlap_saved = zengine.OVERLAP_WGRAD
    zengine.OVERLAP_WGRAD = False       # kernels timed one at a time: nothing runs beside them
    ops.set_profile(prof)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for _ in range(prof_steps):
        # Eager enqueue is host-bound (~15 ms of Python per 12 ms step), so an event pair recorded around a call would
        # also time the launch latency of an idle GPU (and the host gap inside calls that launch two kernels).  A spin
        # kernel ahead of every profiled step lets the host run a whole step ahead: the kernels then execute back to
        # back and the events are taken at kernel boundaries on the device.
        torch.cuda._sleep(int(15e6))     # measured 30-90 ms on B200: longer than the host needs for one step
        step(x_dev, z_dev)
    pe1.record()
    barrier()
    ops.set_profile(None)
    zengine.OVERLAP_WGRAD = overlap_saved
    prof_ms_total = pe0.elapsed_time(pe1)

    # per-kernel accounting from the CUDA events recorded around every convolution call in the timed region
    kinds = {}
    for kind, op, e0, e1 in prof:
        d = kinds.setdefault(kind, {"ms": 0.0, "flops": 0.0, "calls": 0})
        d["ms"] += e0.elapsed_time(e1)
        d["flops"] += conv_flops(op)
        d["calls"] += 1
    if args.layer_table:
        # per-layer table (stderr): which convolutions run far from the tensor roofline
        tab = {}
        for kind, op, e0, e1 in prof:
            key = (kind, op.cin, op.cout, op.kernel, op.stride, op.T, op.H)
            d = tab.setdefault(key, [0.0, 0.0, 0])
            d[0] += e0.elapsed_time(e1)
            d[1] += conv_flops(op)
            d[2] += 1
        if rank == 0:
            sys.stderr.write("kind  cin->cout kernel stride TxH : calls/step  us/call  TFLOP/s  ms/step\n")
            for key, (ms, fl, n) in sorted(tab.items(), key=lambda kv: -kv[1][0]):
                sys.stderr.write(f"{key[0]:5s} {key[1]:4d}->{key[2]:4d} {key[3]} {key[4]} {key[5]}x{key[6]} : "
                                 f"{n / prof_steps:4.0f} {1e3 * ms / n:9.1f} {fl / (ms / 1e3) / 1e12:8.1f} "
                                 f"{ms / prof_steps:7.3f}\n")
    peaks = load_peaks()
    # algorithmic conv FLOPs per clip of the profiled iteration (242.449 GFLOP for R(2+1)D-18)
    flop_per_clip = sum(v["flops"] for v in kinds.values()) / prof_steps / B if kinds else FLOP_PER_CLIP
    # dram__bytes_read+write per launch of the same kernels, from the committed ncu capture -- only if that capture was
    # taken from THIS build of the kernels (the file records the csrc fingerprint build.py stamps the library with)
    traffic, traffic_src, tensor_pipe = None, None, None
    tpath = os.path.join(ROOT, "profiles", "r02_conv_dram_traffic.json")
    if os.path.exists(tpath) and "2plus1d" in args.network:
        tj = json.load(open(tpath))
        if tj.get("build_fingerprint") == build._fingerprint():
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), "profiles/r02_conv_dram_traffic.json (ncu, per launch, this build)"
            tensor_pipe = tj.get("tensor_pipe_active_pct_time_weighted")
        else:
            traffic_src = "profiles/r02_conv_dram_traffic.json is from another build of csrc/ (fingerprint differs): not reported"
    km = {k: {"ms_per_step": v["ms"] / prof_steps, "calls_per_step": v["calls"] / prof_steps,
              "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["ms"] > 0 else None,
              "share_of_step": (v["ms"] / prof_steps) / ms_per_step} for k, v in kinds.items()}
    dom_ms = sum(kinds[k]["ms"] for k in ("fprop", "dgrad") if k in kinds)
    dom_fl = sum(kinds[k]["flops"] for k in ("fprop", "dgrad") if k in kinds)
    dom_calls = sum(kinds[k]["calls"] for k in ("fprop", "dgrad") if k in kinds)
    achieved = dom_fl / (dom_ms / 1e3) / 1e12 if dom_ms > 0 else None
    roofline = {
        "bound": "tensor", "kernel": "igemm_halo_kernel + igemm_kmajor_kernel (conv fprop + dgrad: one tcgen05/TMA implicit-GEMM family, CTA pairs)",
        "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
        "frac": (achieved / peaks["tflops_sustained"]) if achieved else None, "traffic": traffic,
        "traffic_source": traffic_src,
        # sm__pipe_tensor_subpipe_hmma_cycles_active (counts tcgen05 work), time-weighted over the same launches, same capture
        "tensor_pipe_active_pct": tensor_pipe,
        "peak_source": peaks["source"] + ", bf16_tflops_sustained (kernel timed inside a long step)",
        "avg_launch_ms": dom_ms / dom_calls if dom_calls else None,
        "flops_per_launch": dom_fl / dom_calls if dom_calls else None,
        "share_of_step": (dom_ms / prof_steps) / ms_per_step if ms_per_step else None,
        "timed_in": f"{prof_steps} kernel-by-kernel iterations right after the timed region, CUDA events around every "
                    f"convolution call on the launching stream, weight gradients on the same stream, the host kept a step "
                    f"ahead by a spin kernel so that no launch latency is inside an interval "
                    f"({prof_ms_total / prof_steps:.2f} ms/step incl. the spin vs {ms_per_step:.2f} ms/step timed)",
        "by_kernel": km,
        "whole_step_frac_of_tensor_peak": (value / world) * flop_per_clip / (peaks["tflops_sustained"] * 1e12),
    }

    if args.quick:
        if rank == 0:
            emit({"metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world,
                  "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_per_step, "quick": True,
                  "roofline": roofline, "gpu_launches": launches, "clocks": clocks,
                  "host_enqueue_ms_per_step": host_ms_per_step, "launch": graph_note})
        return

    # ---- timed region 2: end to end with host buffers ----
    x_stage = torch.empty_like(x_dev)
    z_stage = torch.empty_like(z_dev)

    def e2e_step():
        if gstep is not None:
            l = gstep()                     # consumes the prefetched batch: D2D into the captured buffers + one graph launch
            gstep.prefetch(x_host, z_host)  # H2D of the next pinned host batch overlaps this replay
        else:
            x_stage.copy_(x_host, non_blocking=True)
            z_stage.copy_(z_host, non_blocking=True)
            l = step(x_stage, z_stage)
        return float(l.detach())            # D2H read of the step's loss (main.py:207)

    if gstep is not None:
        gstep.prefetch(x_host, z_host)
    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (float(t.item()) / 1e3)
    h2d = x_host.numel() * 4 + z_host.numel() * 4
    e2e = {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
           "ms_per_step": float(t.item()) / args.steps}

    # ---- data-parallel correctness on the hardware (N > 1) ----
    consistency = None
    if world > 1:
        consistency = replica_consistency(model, step, x_dev, z_dev, zdist, dist, dev, world)

    if rank != 0:
        return
    extras = {}
    if world == 1 and args.extras and "2plus1d" in args.network:
        # free the step's captured pools before other workloads run on this GPU
        gstep = None
        torch.cuda.empty_cache()
        extras["library_baseline"] = library_baseline(args.network, B, min(args.steps, 20), 3, dev)
        lib = extras["library_baseline"].get("clips_per_s")
        extras["speedup_vs_library"] = (e2e_value / lib) if lib else None
        extras["nearest"] = nearest_workload(dev, 10, 3)
        c3d = subprocess_line(["--network", "c3d", "--no-extras", "--no-cpu-baseline", "--library", "--steps",
                               str(min(args.steps, 20)), "--warmup", "3"])
        extras["c3d"] = {k: c3d.get(k) for k in ("metric", "value", "unit", "ms_per_step", "e2e", "roofline", "gpu_launches",
                                                 "final_loss", "library_baseline", "speedup_vs_library", "error",
                                                 "config") if k in c3d}
    elif world == 1 and args.library:
        extras["library_baseline"] = library_baseline(args.network, B, min(args.steps, 20), 3, dev)
        lib = extras["library_baseline"].get("clips_per_s")
        extras["speedup_vs_library"] = (e2e_value / lib) if lib else None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cps, times, threads = cpu_reference_steps(20, 1, 2)      # ~10-12 s of host work on the GPU box's cores
        cpu = {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
               "sample": f"{len(times)} steps x 2 clips (bs=2x3x16x112x112 fp32, BASELINE.json config 1) of the same "
                         f"training step; {sum(times):.1f} s of CPU work"}
    metric = METRIC if "2plus1d" in args.network else f"{args.network} train clips/s (16x112^2, bf16)"
    line = {
        "metric": metric, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps,
        "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": (f"R(2+1)D-18 end-to-end training step bs={B}/GPU, 16x112x112 synthetic clips, "
                                f"Word2Vec-300 MSE regression, Adam (BASELINE.json configs[{1 if world == 1 else 2}])")
                               if "2plus1d" in args.network else
                               (f"{args.network} end-to-end training step bs={B}/GPU, 16x112x112 synthetic clips, Word2Vec-300 "
                                f"MSE regression, Adam" + (" (BASELINE.json configs[3])" if "c3d" in args.network else "")),
                   "network": args.network, "per_gpu_batch": B, "global_batch": B * world, "clip": "3x16x112x112",
                   "parallelism": f"dp{world}" if world > 1 else "single",
                   "l2": "per-step working set (~6 GB of activations) is far larger than the 126 MB L2; no flush needed",
                   "loss_scaling": "none (bf16)", "weights": "random init (resnet.py:226-236)",
                   "launch": graph_note, "optimizer": ("zsv FusedAdam: torch.optim.Adam's update rule (main.py:131) fused with the bf16 "
                                 "weight re-pack (zsv_adam_pack_step)" if args.optimizer == "zsv"
                                 else "torch.optim.Adam(fused=True) (main.py:131)")},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "final_loss": final_loss, "host_enqueue_ms_per_step": host_ms_per_step,
    }
    if sync is not None:
        line["allreduce_bytes_per_step"] = sync.bytes_per_step
        line["allreduce_collectives_per_step"] = sync.collectives_per_step
    if consistency is not None:
        line["replicas_consistent"] = consistency["ok"]
        line["replica_check"] = consistency
    line.update(extras)
    emit(line)


def replica_consistency(model, step, x_dev, z_dev, zdist, dist, dev, world):
    """Two checks on the real N-GPU job (every rank calls this).
    (1) Gradient exchange: on the same weights and each rank's own batch, the gradients the overlapped, bucketed
        GradSync path leaves in param.grad must equal the mean over ranks of the purely local gradients (computed with
        the exchange switched off, then averaged by ONE plain all-reduce of the whole flat vector).
    (2) Replicas stay identical: after all the optimizer steps of this run, a checksum of every parameter has the same
        value on every rank (all-reduce MIN == all-reduce MAX)."""
    import torch
    params = [p for p in model.parameters() if p.requires_grad]
    saved = {id(p): p.detach().clone() for p in params}

    def grads_of_one_backward():
        torch.manual_seed(4321)          # same Dropout mask in both passes (C3D, network.py:124,167)
        for p in params:
            p.grad = None
        out = model(x_dev)
        emb = out[0] if isinstance(out, tuple) else out
        torch.nn.functional.mse_loss(emb, z_dev).backward()
        live = [p for p in params if p.grad is not None]
        return live, torch.cat([p.grad.detach().reshape(-1).float() for p in live])

    sync = zdist.active_grad_sync()
    head = list(model.output2emb_proj.parameters()) if hasattr(model, "output2emb_proj") else params
    zdist.set_grad_sync(None)            # exchange off: purely local gradients
    if sync is not None:
        sync.detach()
    live, local = grads_of_one_backward()
    dist.all_reduce(local, op=dist.ReduceOp.SUM)
    local /= world
    zdist.set_grad_sync(sync)            # exchange on: arena slices + head hooks, overlapped with backward
    if sync is not None:
        sync.attach(head)
    live2, _ = grads_of_one_backward()
    if sync is not None:
        sync.finish()
    synced = torch.cat([p.grad.detach().reshape(-1).float() for p in live2])
    scale = float(local.abs().max())
    err = float((synced - local).abs().max()) / max(scale, 1e-30)
    with torch.no_grad():                       # the check must not advance training
        for p in params:
            p.copy_(saved[id(p)])
            p.grad = None
    chk = torch.stack([p.detach().double().sum() for p in params]).sum().reshape(1)
    chk2 = torch.stack([p.detach().double().abs().sum() for p in params]).sum().reshape(1)
    v = torch.cat([chk, chk2])
    lo, hi = v.clone(), v.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    errt = torch.tensor([err], device=dev, dtype=torch.float64)
    dist.all_reduce(errt, op=dist.ReduceOp.MAX)
    err = float(errt.item())
    return {"ok": same and err <= 1e-5, "param_checksums_equal_on_all_ranks": same,
            "grad_sync_vs_mean_of_local_max_rel_err": err, "grad_tolerance": 1e-5, "gradients_compared": int(local.numel()),
            "param_checksum": float(v[0].item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "library"],
                    help="b200: this repo; reference: the CPU arm (oracle port); library: stock PyTorch/cuDNN on the GPU")
    ap.add_argument("--workload", default="train", choices=["train", "nearest"],
                    help="train: the training step (headline); nearest: BASELINE.json configs[4], 10k x {101,51,200} top-5")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="N=1 main line only: skip the library baseline, the nearest-class workload and the C3D line")
    ap.add_argument("--library", action="store_true", help="also time the stock-PyTorch (cuDNN) arm of this network")
    ap.add_argument("--batch", type=int, default=22, help="clips per GPU (README.md:45)")
    ap.add_argument("--network", default="r2plus1d_18")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--optimizer", default="zsv", choices=["torch", "zsv"],
                    help="zsv: FusedAdam on the C ABI, Adam + bf16 weight re-pack in one pass (default); "
                         "torch: torch.optim.Adam(fused=True), the reference's optimizer object unchanged")
    ap.add_argument("--layer-table", action="store_true", help="print a per-layer conv timing table to stderr")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="enqueue every iteration kernel by kernel instead of replaying one CUDA graph")
    ap.add_argument("--capture-mode", default="global", choices=["global", "thread_local", "relaxed"])
    ap.add_argument("--quick", action="store_true",
                    help="profiling aid (ncu): honour --warmup below 3, skip the e2e and CPU-baseline legs")
    args = ap.parse_args()
    _capture_stdout()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "library":
        run_library(args)
    elif args.workload == "nearest":
        run_nearest(args)
    else:
        run_b200(args)
    try:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
