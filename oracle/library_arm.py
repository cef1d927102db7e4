"""MEASUREMENT INFRASTRUCTURE ONLY -- the "library path to beat": the reference's own training iteration on stock
PyTorch modules (cuDNN / cuBLAS kernels) on the same B200.

The reference ships no GPU kernel of its own: ``network.Model`` / ``network.C3D`` are ``nn.Conv3d`` /
``nn.BatchNorm3d`` / ``nn.Linear`` trees that reach cuDNN (resnet.py:40-53,94-98,170-186; network.py:102-132), driven by
main.py:170-207 under ``torch.cuda.amp.autocast()`` + ``GradScaler``.  The Python reference cannot travel to the GPU
box, so the backbone is the functional restatement in ``oracle/video_oracle.py`` (pinned against the reference by the
golden fixtures) executed ON CUDA: F.conv3d / F.batch_norm / F.linear are the very ATen ops the modules call.

Two variants, both bs = 22 clips of 3x16x112x112 and the full main.py iteration (zero_grad, forward, MSELoss,
nearest-class accuracy, backward, Adam):
  * "fp16_autocast_as_written": main.py:137,172,195-203 literally -- fp16 autocast, GradScaler, NCDHW tensors, default
    cuDNN settings, torch.optim.Adam defaults, accuracy through scipy on the host (main.py:182-185), loss.item().
  * "bf16_channels_last_3d_tuned": the best stock-PyTorch configuration we know for this model -- bf16 autocast (no
    scaler), channels_last_3d weights and activations, cudnn.benchmark, fused capturable Adam, the iteration replayed as
    one CUDA graph (host overhead removed), no host round trip.
None of the repo's kernels run here.  Only ``bench.py`` imports this file.
"""
from __future__ import annotations

import time
from typing import Dict

import torch
import torch.nn.functional as F

from . import video_oracle as vo

N_TRAIN_CLASSES = 664


def _c3d_state_dict(seed: int, device) -> Dict[str, torch.Tensor]:
    """network.C3D's live parameters (network.py:102-132) with nn.Conv3d / nn.Linear default init."""
    torch.manual_seed(seed)
    sd = {}
    for name, cin, cout in (("conv1", 3, 64), ("conv2", 64, 128), ("conv3a", 128, 256), ("conv3b", 256, 256),
                            ("conv4a", 256, 512), ("conv4b", 512, 512), ("conv5a", 512, 512), ("conv5b", 512, 512)):
        m = torch.nn.Conv3d(cin, cout, 3, padding=1)
        sd[name + ".weight"], sd[name + ".bias"] = m.weight.detach(), m.bias.detach()
    for name, cin, cout in (("fc6", 8192, 4096), ("regressor", 4096, 300)):
        m = torch.nn.Linear(cin, cout)
        sd[name + ".weight"], sd[name + ".bias"] = m.weight.detach(), m.bias.detach()
    return {k: v.to(device) for k, v in sd.items()}


def _time_variant(network: str, variant: str, batch: int, steps: int, warmup: int, device) -> dict:
    import numpy as np
    tuned = variant == "bf16_channels_last_3d_tuned"
    torch.backends.cudnn.benchmark = tuned
    if "c3d" in network:
        sd = _c3d_state_dict(0, device)
    else:
        sd = {k: v.to(device) for k, v in vo.synthetic_state_dict_r2plus1d(0).items()}
    if tuned:
        sd = {k: (v.contiguous(memory_format=torch.channels_last_3d) if v.dim() == 5 else v) for k, v in sd.items()}
    params = [v.requires_grad_(True) for k, v in sd.items()
              if v.is_floating_point() and not k.endswith(("running_mean", "running_var"))]
    optimizer = (torch.optim.Adam(params, lr=1e-3, fused=True, capturable=True) if tuned
                 else torch.optim.Adam(params, lr=1e-3))
    scaler = None if tuned else torch.amp.GradScaler("cuda")
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 1, 3, 16, 112, 112, generator=g).to(device)
    cls = F.normalize(torch.randn(N_TRAIN_CLASSES, 300, generator=torch.Generator().manual_seed(7)))
    labels = torch.randint(0, N_TRAIN_CLASSES, (batch,), generator=g)
    z = cls[labels].contiguous().to(device)
    cls_np, labels_np = cls.numpy().astype(np.float64), labels.numpy()
    dtype = torch.bfloat16 if tuned else torch.float16

    def forward(X):
        if "c3d" in network:
            return vo.c3d_forward(sd, X, train=True, dropout_p=0.10)
        if tuned:
            X = X.reshape(-1, *X.shape[2:]).contiguous(memory_format=torch.channels_last_3d).unsqueeze(1)
        return vo.model_forward(sd, X, train=True)

    def step(X, Z):
        optimizer.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=dtype):
            emb = forward(X)
            loss = F.mse_loss(emb.float(), Z)
        if tuned:
            loss.backward()
            optimizer.step()
        else:
            from scipy.spatial.distance import cdist
            pred = cdist(emb.detach().float().cpu().numpy(), cls_np, "cosine").argmin(1)      # main.py:182-185
            _acc = float((pred == labels_np).mean())
            scaler.scale(loss).backward()
            scaler.step(optimizer)
            scaler.update()
            loss.item()                                                                       # main.py:207
        return loss

    note = "eager"
    run = lambda: step(x, z)
    for _ in range(max(warmup, 3)):
        run()
    torch.cuda.synchronize()
    if tuned:
        try:
            from zeroshotvideoclassification_b200.graph import GraphedStep   # generic capture helper, no repo kernels
            gs = GraphedStep(step, (x, z), device=device, warmup=1)
            run = lambda: gs(*gs.static_inputs)
            note = "whole iteration replayed as one CUDA graph"
            for _ in range(2):
                run()
        except Exception as exc:
            note = f"eager (graph capture failed: {type(exc).__name__}: {str(exc)[:120]})"
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"clips_per_s": batch * 1e3 / ms, "ms_per_step": ms, "steps": steps, "launch": note,
           "host_ms_per_step": 1e3 * (time.perf_counter() - t0) / steps, "final_loss": float(loss.detach()),
           "peak_mem_gb": torch.cuda.max_memory_allocated(device) / 2 ** 30}
    del sd, params, optimizer, x, z
    torch.cuda.empty_cache()
    return out


def run(network: str, batch: int, steps: int, warmup: int, device) -> dict:
    """Times both variants; returns {"variants": {...}, "best": name, "clips_per_s": best value, ...}."""
    res = {}
    for variant in ("fp16_autocast_as_written", "bf16_channels_last_3d_tuned"):
        try:
            res[variant] = _time_variant(network, variant, batch, steps, warmup, device)
        except Exception as exc:       # reported in the JSON line, never silent
            res[variant] = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
            torch.cuda.synchronize()
            torch.cuda.empty_cache()
    torch.backends.cudnn.benchmark = False
    ok = {k: v for k, v in res.items() if "clips_per_s" in v}
    best = max(ok, key=lambda k: ok[k]["clips_per_s"]) if ok else None
    return {"what": "the reference's training iteration on stock PyTorch (cuDNN/cuBLAS) on this GPU: "
                    "oracle/video_oracle.py run on CUDA, none of the repo's kernels",
            "network": network, "per_gpu_batch": batch, "variants": res, "best": best,
            "clips_per_s": ok[best]["clips_per_s"] if best else None,
            "cudnn": torch.backends.cudnn.version(), "torch": torch.__version__}
