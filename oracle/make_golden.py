"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/* from the UNMODIFIED reference.

Run in the build container (needs /root/reference, scipy):

    python -m oracle.make_golden

It imports the reference's own ``resnet.py`` / ``network.py`` (with ``sys.modules`` stubs for the unused
top-level imports ``gensim`` and ``clip``, network.py:8,19), runs its forward/backward on CPU fp32 through
stock PyTorch, and restates ``compute_accuracy`` (main.py:316-325) with scipy exactly as written there
(main.py itself cannot be imported: it parses argv and builds datasets at import time, main.py:55-137).
The reference cannot travel to the GPU box, so the outputs are committed as small fixtures.
"""
from __future__ import annotations

import json
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def import_reference():
    gensim = types.ModuleType("gensim")
    gm = types.ModuleType("gensim.models")
    gm.KeyedVectors = type("KeyedVectors", (), {})
    gensim.models = gm
    sys.modules.setdefault("gensim", gensim)
    sys.modules.setdefault("gensim.models", gm)
    sys.modules.setdefault("clip", types.ModuleType("clip"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import network  # noqa: E402  (the reference's network.py)
    return network


def checksum(t: torch.Tensor):
    t = t.detach().double().flatten()
    # (fp32 linspace: the end point of a >2^24-element tensor rounds up by one, hence the clamp)
    idx = torch.linspace(0, t.numel() - 1, steps=min(8, t.numel())).long().clamp_(max=t.numel() - 1)
    return {"sum": float(t.sum()), "abs": float(t.abs().sum()), "sq": float((t * t).sum()),
            "samples": [float(v) for v in t[idx]]}


def synthetic_batch(B, T, H, W, seed, n_classes=101):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 1, 3, T, H, W, generator=g)
    cls = F.normalize(torch.randn(n_classes, 300, generator=g))
    labels = torch.randint(0, n_classes, (B,), generator=g)
    return x, cls[labels], labels


def reference_step(network, B, T, H, W, seed, arch="r2plus1d_18"):
    """main.py:170-195 on the reference modules (fp32 CPU, no optimizer step), with forward hooks recording
    every conv / block output."""
    torch.manual_seed(seed)
    opt = SimpleNamespace(network=arch, fixconvs=False, nopretrained=False)
    model = network.get_network(opt).train()
    init = {k: checksum(v) for k, v in model.state_dict().items() if v.is_floating_point()}
    keys = {k: list(v.shape) for k, v in model.state_dict().items()}
    x, z, _ = synthetic_batch(B, T, H, W, seed + 100)
    acts = {}

    def hook(name):
        def fn(_m, _i, o):
            acts[name] = checksum(o)
        return fn

    for name, m in model.model.named_modules():
        if isinstance(m, torch.nn.Conv3d) or (name.startswith("layer") and name.count(".") == 1):
            m.register_forward_hook(hook(name))
    emb, none = model(x)
    assert none is None
    loss = torch.nn.MSELoss()(emb, z)
    loss.backward()
    grads = {k: checksum(p.grad) for k, p in model.named_parameters() if p.grad is not None}
    dead = sorted(k for k, p in model.named_parameters() if p.grad is None)
    bn_after = {k: checksum(v) for k, v in model.state_dict().items()
                if k.endswith(("running_mean", "running_var")) and k.startswith("model.")}
    return dict(config=dict(B=B, T=T, H=H, W=W, seed=seed, arch=arch), emb=emb.detach().numpy().tolist(), loss=float(loss),
                acts=acts, grads=grads, dead=dead, bn_after=bn_after), init, keys


def reference_c3d_step(network, B, T, H, W, seed):
    """main.py:170-195 on the reference's network.C3D (network.py:95-180; fp32 CPU, no optimizer step) with
    Dropout p set to 0 (network.py:124: the mask is not reproducible across devices), forward hooks on every
    conv / pool / linear."""
    torch.manual_seed(seed)
    opt = SimpleNamespace(network="c3d", fixconvs=False, nopretrained=False)
    model = network.get_network(opt).train()
    model.dropout.p = 0.0
    init = {k: checksum(v) for k, v in model.state_dict().items() if v.is_floating_point()}
    keys = {k: list(v.shape) for k, v in model.state_dict().items()}
    x, z, _ = synthetic_batch(B, T, H, W, seed + 100)
    acts = {}

    def hook(name):
        def fn(_m, _i, o):
            acts[name] = checksum(o)
        return fn

    for name, m in model.named_modules():
        if isinstance(m, (torch.nn.Conv3d, torch.nn.MaxPool3d, torch.nn.Linear)):
            m.register_forward_hook(hook(name))
    emb = model(x)
    loss = torch.nn.MSELoss()(emb, z)
    loss.backward()
    grads = {k: checksum(p.grad) for k, p in model.named_parameters() if p.grad is not None}
    dead = sorted(k for k, p in model.named_parameters() if p.grad is None)
    return dict(config=dict(B=B, T=T, H=H, W=W, seed=seed, arch="c3d"), emb=emb.detach().numpy().tolist(),
                loss=float(loss), acts=acts, grads=grads, dead=dead), init, keys


def nearest_fixture():
    """main.py:321-322 restated with scipy exactly as the reference calls it."""
    from scipy.spatial.distance import cdist
    rng = np.random.default_rng(2026)
    out = {}
    for C in (51, 101, 200):
        emb = rng.standard_normal((400, 300)).astype(np.float32)
        emb /= np.linalg.norm(emb, axis=1, keepdims=True)
        cls = rng.standard_normal((C, 300)).astype(np.float32)
        cls /= np.linalg.norm(cls, axis=1, keepdims=True)
        emb[:10] = cls[:10]                                  # exact hits: distance 0
        emb[10:20] = cls[:10] + 1e-4 * emb[10:20]            # near ties
        d = cdist(emb, cls, "cosine")
        out[f"emb_{C}"] = emb
        out[f"cls_{C}"] = cls
        out[f"dist_{C}"] = d
        out[f"argmin_{C}"] = d.argmin(1)
        out[f"top5_{C}"] = d.argsort(1)[:, :5]
    np.savez_compressed(os.path.join(OUT, "nearest_scipy.npz"), **out)


def transform_fixture():
    """The reference's own clip transform (auxiliary/transforms.py:41-56) on small random frames: the validation
    pipeline as get_transform(True) builds it, and the training pipeline with its random decisions pinned
    (crop origin (5, 9), flipped).  transforms.py:3 imports imageio for GIF dumps only; it is stubbed."""
    import importlib
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    if REF not in sys.path:
        sys.path.insert(0, REF)
    tr = importlib.import_module("auxiliary.transforms")
    g = torch.Generator().manual_seed(123)
    out = {}
    frames = torch.randint(0, 256, (1, 68, 90, 3), generator=g, dtype=torch.uint8)       # upscaled: short side 68 -> 128
    out["landscape_frames"] = frames.numpy()
    out["landscape_val"] = tr.get_transform(True)(frames).numpy()
    frames = torch.randint(0, 256, (1, 200, 150, 3), generator=g, dtype=torch.uint8)     # downscaled: short side 150 -> 128
    vid = tr.Resize(128)(tr.ToFloatTensorInZeroOne()(frames))
    out["portrait_frames"] = frames.numpy()
    out["portrait_train_5_9_flip"] = tr.crop(vid, 5, 9, 112, 112).flip(dims=(-1,)).numpy()
    out["portrait_resized_hw"] = np.array(vid.shape[-2:])
    np.savez_compressed(os.path.join(OUT, "clip_transform.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    network = import_reference()
    small, init, keys = reference_step(network, 2, 8, 32, 32, seed=0)
    json.dump(keys, open(os.path.join(OUT, "r2plus1d_state_dict_keys.json"), "w"), indent=0)
    json.dump(init, open(os.path.join(OUT, "r2plus1d_init_seed0.json"), "w"))
    json.dump(small, open(os.path.join(OUT, "r2plus1d_step_small.json"), "w"))
    full, _, _ = reference_step(network, 2, 16, 112, 112, seed=0)   # BASELINE.json config 1
    json.dump(full, open(os.path.join(OUT, "r2plus1d_step_bs2_16x112.json"), "w"))
    # r3d_18 (network.py:28-30): the third backbone get_network can select
    r3d, r3d_init, r3d_keys = reference_step(network, 2, 8, 32, 32, seed=0, arch="r3d_18")
    json.dump(r3d_keys, open(os.path.join(OUT, "r3d_state_dict_keys.json"), "w"), indent=0)
    json.dump(r3d_init, open(os.path.join(OUT, "r3d_init_seed0.json"), "w"))
    json.dump(r3d, open(os.path.join(OUT, "r3d_step_small.json"), "w"))
    # C3D (network.py:95-180): fc6 needs 8192 = 512 x 1 x 4 x 4 features, so the clip is the full 16 x 112 x 112
    c3d, c3d_init, c3d_keys = reference_c3d_step(network, 2, 16, 112, 112, seed=0)
    json.dump(c3d_keys, open(os.path.join(OUT, "c3d_state_dict_keys.json"), "w"), indent=0)
    json.dump(c3d_init, open(os.path.join(OUT, "c3d_init_seed0.json"), "w"))
    json.dump(c3d, open(os.path.join(OUT, "c3d_step_bs2_16x112.json"), "w"))
    nearest_fixture()
    transform_fixture()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
