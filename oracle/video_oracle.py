"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the R(2+1)D-18 / C3D embedding networks.

A functional, state-dict driven restatement (plain PyTorch CPU, fp32 or fp64) of the arithmetic the
reference reaches through its nn.Module tree.  Every function cites the reference lines it follows.
It is deliberately slow and simple; parity tests compare the CUDA path against it.

Pinned against the reference itself by ``oracle/make_golden.py`` -> ``tests/golden/`` (the reference has
no tests of its own; see oracle/__init__.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

BN_EPS = 1e-5        # nn.BatchNorm3d default (resnet.py:48)
BN_MOMENTUM = 0.1    # nn.BatchNorm3d default


def midplanes(inplanes: int, planes: int) -> int:
    """Width of the factorised (2+1)D convolution -- resnet.py:91."""
    return (inplanes * planes * 3 * 3 * 3) // (inplanes * 3 * 3 + 3 * planes)


def batchnorm_train(x: Tensor, sd: Dict[str, Tensor], prefix: str, trace: Optional[dict] = None,
                    update_running: bool = True) -> Tensor:
    """nn.BatchNorm3d in training mode (resnet.py:48,95,97,182,185,272) through the same ATen op the module
    calls (F.batch_norm): batch mean / biased variance over (N,T,H,W); running stats updated with momentum 0.1
    and the unbiased variance; num_batches_tracked += 1."""
    rm = sd[prefix + ".running_mean"] if update_running else None
    rv = sd[prefix + ".running_var"] if update_running else None
    y = F.batch_norm(x, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], training=True,
                     momentum=BN_MOMENTUM, eps=BN_EPS)
    if update_running:
        with torch.no_grad():
            sd[prefix + ".num_batches_tracked"].add_(1)
    if trace is not None:
        with torch.no_grad():
            trace[prefix + ":mean"] = x.mean((0, 2, 3, 4))
            trace[prefix + ":var"] = x.var((0, 2, 3, 4), unbiased=False)
    return y


def batchnorm_eval(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """nn.BatchNorm3d in eval mode (main.py:229): running statistics."""
    shape = (1, -1, 1, 1, 1)
    inv = torch.rsqrt(sd[prefix + ".running_var"].view(shape) + BN_EPS)
    return (x - sd[prefix + ".running_mean"].view(shape)) * inv * sd[prefix + ".weight"].view(shape) + \
        sd[prefix + ".bias"].view(shape)


class _RoundBf16Both(torch.autograd.Function):
    """bf16 storage point of an activation: rounds the value in forward and the gradient in backward."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


class _RoundBf16Fwd(torch.autograd.Function):
    """bf16 copy of an fp32 master value (weights, input clip): rounds in forward, gradient passes in fp32."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


class _Net:
    """Shared plumbing: conv / bn helpers that record per-layer tensors when tracing.

    ``emulate_bf16`` places bf16 rounding at exactly the points where the B200 path stores bf16 (conv outputs,
    post-activation tensors, their gradients, bf16 weight copies) while all arithmetic stays fp32 -- the same
    mixed-precision contract as north_star ("bf16 compute, fp32 accumulate").  A randomly initialised 37-BN-deep
    network amplifies any perturbation ~1e2-1e3x (numerics), so plain fp32 vs bf16 comparisons
    only bound the error loosely; this mode checks the composition of the kernels tightly."""

    def __init__(self, sd: Dict[str, Tensor], train: bool, trace: Optional[dict], emulate_bf16: bool = False):
        self.sd, self.train, self.trace, self.emu = sd, train, trace, emulate_bf16

    def act(self, x: Tensor, name: Optional[str] = None) -> Tensor:
        """Storage point of a post-activation tensor (bf16 on the B200 path)."""
        x = _RoundBf16Both.apply(x) if self.emu else x
        if self.trace is not None and name is not None:
            if x.requires_grad:
                x.retain_grad()
            self.trace[name + ":out"] = x
        return x

    def conv(self, x: Tensor, name: str, stride, padding) -> Tensor:
        w = self.sd[name + ".weight"]
        if self.emu:
            w = _RoundBf16Fwd.apply(w)
        y = F.conv3d(x, w, self.sd.get(name + ".bias"), stride=stride, padding=padding)
        y = _RoundBf16Both.apply(y) if self.emu else y
        if self.trace is not None:
            if y.requires_grad:
                y.retain_grad()
            self.trace[name] = y
            self.trace[name + ":in"] = x.detach()      # teacher-forced per-layer tests feed the kernels exactly this
        return y

    def bn(self, x: Tensor, name: str) -> Tensor:
        y = batchnorm_train(x, self.sd, name, self.trace) if self.train else batchnorm_eval(x, self.sd, name)
        if self.trace is not None:
            if y.requires_grad:
                y.retain_grad()
            self.trace[name] = y
        return y


def _conv2plus1d(net: _Net, x: Tensor, prefix: str, stride: int) -> Tensor:
    """Conv2Plus1D (resnet.py:37-57): spatial 1x3x3 (stride (1,s,s)) -> BN(mid) -> ReLU -> temporal 3x1x1
    (stride (s,1,1)); no bias."""
    x = net.conv(x, prefix + ".0", (1, stride, stride), (0, 1, 1))
    x = net.act(F.relu(net.bn(x, prefix + ".1")), prefix + ".1")
    return net.conv(x, prefix + ".3", (stride, 1, 1), (1, 0, 0))


def _basic_block(net: _Net, x: Tensor, prefix: str, stride: int, has_ds: bool) -> Tensor:
    """BasicBlock.forward (resnet.py:102-113)."""
    residual = x
    out = _conv2plus1d(net, x, prefix + ".conv1.0", stride)
    out = net.act(F.relu(net.bn(out, prefix + ".conv1.1")), prefix + ".conv1.1")
    out = _conv2plus1d(net, out, prefix + ".conv2.0", 1)
    out = net.bn(out, prefix + ".conv2.1")
    if has_ds:  # resnet.py:268-273: 1x1x1 conv, stride (s,s,s), then BN
        residual = net.conv(x, prefix + ".downsample.0", (stride, stride, stride), (0, 0, 0))
        residual = net.bn(residual, prefix + ".downsample.1")
    out = net.act(F.relu(out + residual))
    if net.trace is not None:
        if out.requires_grad:
            out.retain_grad()
        net.trace[prefix] = out
    return out


def r2plus1d_18_features(sd: Dict[str, Tensor], x: Tensor, train: bool = True, trace: Optional[dict] = None,
                         prefix: str = "model.", emulate_bf16: bool = False) -> Tensor:
    """VideoResNet.forward up to layer4 (resnet.py:243-249) for r2plus1d_18 (resnet.py:342-362):
    x [B,3,T,H,W] -> f [B,512,T/8,H/16,W/16]."""
    net = _Net({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}, train, trace, emulate_bf16)
    if emulate_bf16:
        x = _RoundBf16Fwd.apply(x)
    # R2Plus1dStem (resnet.py:176-187)
    x = net.conv(x, "stem.0", (1, 2, 2), (0, 3, 3))
    x = net.act(F.relu(net.bn(x, "stem.1")))
    x = net.conv(x, "stem.3", (1, 1, 1), (1, 0, 0))
    x = net.act(F.relu(net.bn(x, "stem.4")))
    if trace is not None:
        trace["stem"] = x
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):  # _make_layer (resnet.py:258-281), layers=[2,2,2,2]
        x = _basic_block(net, x, f"layer{li}.0", stride, has_ds=(li != 1))
        x = _basic_block(net, x, f"layer{li}.1", 1, has_ds=False)
    return x


def _simple_block(net: _Net, x: Tensor, prefix: str, stride: int, has_ds: bool) -> Tensor:
    """BasicBlock.forward (resnet.py:102-113) over Conv3DSimple 3x3x3 convolutions (resnet.py:18-34)."""
    residual = x
    out = net.conv(x, prefix + ".conv1.0", (stride, stride, stride), (1, 1, 1))
    out = net.act(F.relu(net.bn(out, prefix + ".conv1.1")), prefix + ".conv1.1")
    out = net.conv(out, prefix + ".conv2.0", (1, 1, 1), (1, 1, 1))
    out = net.bn(out, prefix + ".conv2.1")
    if has_ds:
        residual = net.conv(x, prefix + ".downsample.0", (stride, stride, stride), (0, 0, 0))
        residual = net.bn(residual, prefix + ".downsample.1")
    out = net.act(F.relu(out + residual))
    if net.trace is not None:
        if out.requires_grad:
            out.retain_grad()
        net.trace[prefix] = out
    return out


def r3d_18_features(sd: Dict[str, Tensor], x: Tensor, train: bool = True, trace: Optional[dict] = None,
                    prefix: str = "model.", emulate_bf16: bool = False) -> Tensor:
    """VideoResNet.forward up to layer4 (resnet.py:243-249) for r3d_18 (resnet.py:293-314): BasicStem
    (resnet.py:165-173), then layers=[2,2,2,2] of BasicBlock over Conv3DSimple."""
    net = _Net({k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}, train, trace, emulate_bf16)
    if emulate_bf16:
        x = _RoundBf16Fwd.apply(x)
    x = net.conv(x, "stem.0", (1, 2, 2), (1, 3, 3))
    x = net.act(F.relu(net.bn(x, "stem.1")))
    if trace is not None:
        trace["stem"] = x
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        x = _simple_block(net, x, f"layer{li}.0", stride, has_ds=(li != 1))
        x = _simple_block(net, x, f"layer{li}.1", 1, has_ds=False)
    return x


def embedding_head(sd: Dict[str, Tensor], feats: Tensor) -> Tensor:
    """network.py:595-596 with MLP network.py:613-618: mean over (T,H,W), Linear-ReLU-Linear, F.normalize."""
    f = feats.mean(dim=(2, 3, 4))
    h = F.relu(F.linear(f, sd["output2emb_proj.layers.0.weight"], sd["output2emb_proj.layers.0.bias"]))
    o = F.linear(h, sd["output2emb_proj.layers.1.weight"], sd["output2emb_proj.layers.1.bias"])
    return F.normalize(o)


def model_forward(sd: Dict[str, Tensor], x: Tensor, train: bool = True, trace: Optional[dict] = None,
                  emulate_bf16: bool = False, arch: str = "r2plus1d_18") -> Tensor:
    """network.Model.forward (network.py:533-600), live lines only: x [B,nc,3,T,H,W] -> emb [B*nc,300].
    arch selects the backbone like network.get_network (network.py:28-33)."""
    bs, nc = x.shape[:2]
    x = x.reshape(bs * nc, *x.shape[2:])
    backbone = r3d_18_features if "r3d" in arch else r2plus1d_18_features
    feats = backbone(sd, x, train, trace, emulate_bf16=emulate_bf16)
    if trace is not None:
        trace["feats"] = feats
    return embedding_head(sd, feats)


def c3d_forward(sd: Dict[str, Tensor], x: Tensor, train: bool = False, emulate_bf16: bool = False,
                trace: Optional[dict] = None, dropout_p: float = 0.10) -> Tensor:
    """network.C3D.forward (network.py:143-180).  Dropout(p=0.1) (network.py:124,167) is only applied in train
    mode; parity tests run with train=False or dropout_p=0 because the mask is not reproducible across devices.
    emulate_bf16: bf16 rounding at the storage points of the B200 path (input clip, weight copies, every
    conv+bias+ReLU output and its gradient); arithmetic stays fp32.  trace: records the output of every
    conv (pre-ReLU, as a forward hook on the nn.Conv3d sees it), pool and linear by module name."""
    bs, nc = x.shape[:2]
    h = x.reshape(bs * nc, *x.shape[2:])
    if emulate_bf16:
        h = _RoundBf16Fwd.apply(h)

    def keep(name, t):
        if trace is not None:
            if t.requires_grad:
                t.retain_grad()
            trace[name] = t
        return t

    def conv(name, t):
        w = sd[name + ".weight"]
        if emulate_bf16:
            w = _RoundBf16Fwd.apply(w)
        if trace is not None:
            trace[name + ":in"] = t.detach()
        y = F.relu(keep(name, F.conv3d(t, w, sd[name + ".bias"], padding=1)))
        y = _RoundBf16Both.apply(y) if emulate_bf16 else y
        return keep(name + ":out", y)

    h = keep("pool1", F.max_pool3d(conv("conv1", h), (1, 2, 2), (1, 2, 2)))
    h = keep("pool2", F.max_pool3d(conv("conv2", h), (2, 2, 2), (2, 2, 2)))
    h = keep("pool3", F.max_pool3d(conv("conv3b", conv("conv3a", h)), (2, 2, 2), (2, 2, 2)))
    h = keep("pool4", F.max_pool3d(conv("conv4b", conv("conv4a", h)), (2, 2, 2), (2, 2, 2)))
    h = keep("pool5", F.max_pool3d(conv("conv5b", conv("conv5a", h)), (2, 2, 2), (2, 2, 2), padding=(0, 1, 1)))
    h = h.reshape(-1, 8192)
    h = F.relu(keep("fc6", F.linear(h, sd["fc6.weight"], sd["fc6.bias"])))
    if train and dropout_p > 0:
        h = F.dropout(h, p=dropout_p, training=True)
    h = h.reshape(bs, nc, -1).mean(1).reshape(bs, -1)
    h = keep("regressor", F.linear(h, sd["regressor.weight"], sd["regressor.bias"]))
    return F.normalize(h, dim=-1)


def c3d_train_step_grads(sd: Dict[str, Tensor], x: Tensor, target: Tensor, trace: Optional[dict] = None,
                         emulate_bf16: bool = False) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """One forward + backward of main.py:170-195 on network.C3D with Dropout p = 0 (no optimizer):
    returns (emb, loss, grads by state-dict key); fc7 / fc8 (network.py:121-122, never used by forward) get none."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    emb = c3d_forward(params, x, train=True, emulate_bf16=emulate_bf16, trace=trace, dropout_p=0.0)
    loss = mse_loss(emb, target)
    loss.backward()
    return emb.detach(), loss.detach(), {k: p.grad for k, p in params.items() if p.grad is not None}


def mse_loss(emb: Tensor, target: Tensor) -> Tensor:
    """nn.MSELoss() with mean reduction (main.py:130,179)."""
    return ((emb - target) ** 2).mean()


def train_step_grads(sd: Dict[str, Tensor], x: Tensor, target: Tensor, trace: Optional[dict] = None,
                     loss_scale: float = 1.0, emulate_bf16: bool = False,
                     arch: str = "r2plus1d_18") -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """One forward + backward of main.py:170-195 (no optimizer): returns (emb, loss, grads by state-dict key)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.is_floating_point()
              and not k.endswith(("running_mean", "running_var"))}
    work = dict(sd)
    work.update(params)
    emb = model_forward(work, x, train=True, trace=trace, emulate_bf16=emulate_bf16, arch=arch)
    loss = mse_loss(emb, target)
    (loss * loss_scale).backward()
    grads = {k: p.grad for k, p in params.items() if p.grad is not None}
    return emb.detach(), loss.detach(), grads


# ----------------------------------------------------------------------------------------------------
# single ops in channels-first fp32, used by per-kernel parity tests
# ----------------------------------------------------------------------------------------------------
def conv3d(x: Tensor, w: Tensor, bias: Optional[Tensor], stride, padding) -> Tensor:
    return F.conv3d(x, w, bias, stride=stride, padding=padding)


def conv3d_grads(x: Tensor, w: Tensor, dy: Tensor, stride, padding) -> Tuple[Tensor, Tensor]:
    """(dx, dw) of conv3d by autograd."""
    x = x.detach().clone().requires_grad_(True)
    w = w.detach().clone().requires_grad_(True)
    F.conv3d(x, w, None, stride=stride, padding=padding).backward(dy)
    return x.grad, w.grad


def synthetic_state_dict_r2plus1d(seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Random-init state dict with the reference's key names, shapes and init scheme (resnet.py:226-236:
    Kaiming-normal fan_out for convs, BN weight 1 / bias 0; nn.Linear default init for the MLP).  Only the
    live parameters of network.Model are created (the dead Transformer branch, network.py:500-514, never
    reaches the output).  NOT bit-identical to constructing the reference modules (different RNG order); use
    the golden fixtures for that."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}

    def conv(name, cout, cin, k):
        fan_out = cout * k[0] * k[1] * k[2]
        sd[name + ".weight"] = torch.randn(cout, cin, *k, generator=g, dtype=dtype) * (2.0 / fan_out) ** 0.5

    def bn(name, c):
        sd[name + ".weight"] = torch.ones(c, dtype=dtype)
        sd[name + ".bias"] = torch.zeros(c, dtype=dtype)
        sd[name + ".running_mean"] = torch.zeros(c, dtype=dtype)
        sd[name + ".running_var"] = torch.ones(c, dtype=dtype)
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def linear(name, cout, cin):
        bound = 1.0 / cin ** 0.5
        sd[name + ".weight"] = (torch.rand(cout, cin, generator=g, dtype=dtype) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand(cout, generator=g, dtype=dtype) * 2 - 1) * bound

    conv("model.stem.0", 45, 3, (1, 7, 7)); bn("model.stem.1", 45)
    conv("model.stem.3", 64, 45, (3, 1, 1)); bn("model.stem.4", 64)
    inplanes = 64
    for li, planes in ((1, 64), (2, 128), (3, 256), (4, 512)):
        for bi in range(2):
            p = f"model.layer{li}.{bi}"
            cin = inplanes if bi == 0 else planes
            mid1 = midplanes(cin, planes)
            conv(p + ".conv1.0.0", mid1, cin, (1, 3, 3)); bn(p + ".conv1.0.1", mid1)
            conv(p + ".conv1.0.3", planes, mid1, (3, 1, 1)); bn(p + ".conv1.1", planes)
            # resnet.py:91,97: conv2 reuses the block's midplanes (computed from the block's inplanes)
            conv(p + ".conv2.0.0", mid1, planes, (1, 3, 3)); bn(p + ".conv2.0.1", mid1)
            conv(p + ".conv2.0.3", planes, mid1, (3, 1, 1)); bn(p + ".conv2.1", planes)
            if bi == 0 and li != 1:
                conv(p + ".downsample.0", planes, cin, (1, 1, 1)); bn(p + ".downsample.1", planes)
        inplanes = planes
    linear("output2emb_proj.layers.0", 512, 512)
    linear("output2emb_proj.layers.1", 300, 512)
    return sd
