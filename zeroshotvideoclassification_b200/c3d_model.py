"""C3D (network.C3D, network.py:95-180) on libzsv_b200.so: 8 x [3x3x3 conv + bias + ReLU] with 5 max-pools as one
autograd Function over channels-last bf16 buffers, then fc6 / regressor as fp32 weight-streaming linears.

Module tree, parameter names and construction order are the reference's (conv1 .. conv5b, fc6, fc7, fc8, regressor;
fc7/fc8 are constructed but unused there too, network.py:121-127,168-172), so state dicts are interchangeable.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops

# (conv name, Cin, Cout, pool after it or None): network.py:102-118
_C3D_LAYERS = [
    ("conv1", 3, 64, ((1, 2, 2), (0, 0, 0))),
    ("conv2", 64, 128, ((2, 2, 2), (0, 0, 0))),
    ("conv3a", 128, 256, None),
    ("conv3b", 256, 256, ((2, 2, 2), (0, 0, 0))),
    ("conv4a", 256, 512, None),
    ("conv4b", 512, 512, ((2, 2, 2), (0, 0, 0))),
    ("conv5a", 512, 512, None),
    ("conv5b", 512, 512, ((2, 2, 2), (0, 1, 1))),
]

_conv_cache: Dict[tuple, ops.Conv3d] = {}


def _conv(N, T, H, W, cin, cout, layout):
    key = (N, T, H, W, cin, cout, layout)
    op = _conv_cache.get(key)
    if op is None:
        op = ops.Conv3d(N, T, H, W, cin, cout, (3, 3, 3), (1, 1, 1), (1, 1, 1), layout)
        _conv_cache[key] = op
    return op


_plan_cache: Dict[tuple, tuple] = {}


def _trunk_plan(N, T, H, W):
    """Geometry of the 8 convolutions for a clip batch of this shape and the one-launch weight re-pack for them."""
    key = (N, T, H, W)
    hit = _plan_cache.get(key)
    if hit is not None:
        return hit
    convs, dims = [], (T, H, W)
    for i, (name, cin, cout, pool) in enumerate(_C3D_LAYERS):
        convs.append(_conv(N, *dims, cin, cout, _lib.X_WFOLD if i == 0 else _lib.X_NDHWC))
        if pool is not None:
            (kt, kh, kw), (pt, ph, pw) = pool
            dims = ((dims[0] + 2 * pt - kt) // kt + 1, (dims[1] + 2 * ph - kh) // kh + 1, (dims[2] + 2 * pw - kw) // kw + 1)
    plan = ops.PackPlan(convs, [i > 0 for i in range(len(convs))])
    _plan_cache[key] = (convs, plan)
    return convs, plan


class _TrunkFn(torch.autograd.Function):
    """conv1 .. pool5 (network.py:147-163): x [B,3,T,H,W] fp32 -> bf16 [B,T',H',W',512]."""

    @staticmethod
    def forward(ctx, x, *wb):
        from . import engine
        need_grad = any(ctx.needs_input_grad[1:])
        N, _, T, H, W = x.shape
        a = ops.repack_input(x, _lib.X_WFOLD, 1)
        convs, plan = _trunk_plan(N, T, H, W)
        weights = [wb[2 * i].detach() for i in range(len(_C3D_LAYERS))]
        pub = engine.published_for(weights[0])
        if pub is not None:              # images written by the optimizer step (zsv_adam_pack_step): nothing to re-pack
            wfs, wds = pub.wfs, pub.wds
        else:
            wfs, wds = plan.pack(weights)
        tape = []
        for i, (name, cin, cout, pool) in enumerate(_C3D_LAYERS):
            b = wb[2 * i + 1].detach().float().contiguous()
            op = convs[i]
            y, _, _ = op.fprop(a, wfs[i], stats=False, bias=b, relu=True)
            rec = dict(op=op, x=a, wd=wds[i], out=y, pool=None)
            a = y
            if pool is not None:
                k, p = pool
                pooled, am = ops.maxpool3d_fwd(y, cout, k, p)
                rec["pool"] = (k, p, am, tuple(y.shape), pooled)
                rec["out"] = None        # backward takes the ReLU mask from the pooled tensor: the big one can go
                a = pooled
            tape.append(rec)
        ctx.tape = tape if need_grad else None
        ctx.n = len(wb)
        return a

    @staticmethod
    def backward(ctx, g):
        if ctx.tape is None:
            return (None,) + (None,) * ctx.n
        grads: List[Optional[torch.Tensor]] = [None] * ctx.n
        g = g.contiguous()
        for i in reversed(range(len(_C3D_LAYERS))):
            rec = ctx.tape[i]
            name, cin, cout, _ = _C3D_LAYERS[i]
            op = rec["op"]
            want = ctx.needs_input_grad[1 + 2 * i] or ctx.needs_input_grad[2 + 2 * i]
            # pool + ReLU backward and the bias gradient (column sums of dz) in one pass over the tensor
            if rec["pool"] is not None:
                k, p, am, shape, pooled = rec["pool"]
                res = ops.maxpool3d_bwd(g, am, shape, cout, k, p, relu_pooled=pooled, want_bias=want)
            else:
                res = ops.relu_bwd(g, rec["out"], cout, want_bias=want)
            dz, db = res if want else (res, None)
            if want:
                dw, _ = op.wgrad(rec["x"], dz)
                grads[2 * i], grads[2 * i + 1] = dw, db
            if i > 0:
                g = op.dgrad(dz, rec["wd"])
        ctx.tape = None
        return (None,) + tuple(grads)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, relu):
        xc, wc, bc = x.detach().float().contiguous(), w.detach().float().contiguous(), b.detach().float().contiguous()
        out = ops.linear_fwd(xc, wc, bc, relu)
        ctx.save_for_backward(xc, wc, out if relu else None)
        ctx.relu = relu
        return out

    @staticmethod
    def backward(ctx, dy):
        xc, wc, act = ctx.saved_tensors
        dx, dw, db = ops.linear_bwd(dy.float().contiguous(), xc, wc, act if ctx.relu else None,
                                    need_dx=ctx.needs_input_grad[0], need_dw=ctx.needs_input_grad[1])
        return dx, dw, db, None


class _L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, o):
        emb, onorm = ops.l2norm_fwd(o.detach().float().contiguous())
        ctx.save_for_backward(emb, onorm)
        return emb

    @staticmethod
    def backward(ctx, demb):
        emb, onorm = ctx.saved_tensors
        return ops.l2norm_bwd(demb.float().contiguous(), emb, onorm)


class C3D(nn.Module):
    """network.C3D (network.py:95-180)."""

    def __init__(self, fixconvs: bool = False, nopretrained: bool = True):
        super().__init__()
        for name, cin, cout, pool in _C3D_LAYERS:
            setattr(self, name, nn.Conv3d(cin, cout, kernel_size=(3, 3, 3), padding=(1, 1, 1)))
            if pool is not None:
                k, p = pool
                setattr(self, "pool" + name[4], nn.MaxPool3d(kernel_size=k, stride=k, padding=p))
        self.fc6 = nn.Linear(8192, 4096)
        self.fc7 = nn.Linear(4096, 4096)
        self.fc8 = nn.Linear(4096, 487)
        self.dropout = nn.Dropout(p=0.10)
        self.relu = nn.ReLU()
        self.softmax = nn.Softmax()
        if nopretrained:
            self.load_state_dict(torch.load("./assets/c3d.pickle"))     # network.py:129-130 (never taken: main.py:42)
        self.regressor = nn.Linear(4096, 300)
        if fixconvs:
            for name, *_ in _C3D_LAYERS:
                for p in getattr(self, name).parameters():
                    p.requires_grad = False
            for p in self.fc6.parameters():
                p.requires_grad = False

    def packed_plan(self):
        """(PackPlan, conv weights) for ``optim.FusedAdam(model=...)``: the layout of the bf16 weight images."""
        _, plan = _trunk_plan(1, 16, 112, 112)
        return plan, [getattr(self, name).weight for name, *_ in _C3D_LAYERS]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        ops._require_cuda(x, "C3D.forward")
        bs, nc = x.shape[:2]
        x = x.reshape(bs * nc, *x.shape[2:])
        wb = []
        for name, *_ in _C3D_LAYERS:
            m = getattr(self, name)
            wb += [m.weight, m.bias]
        if not torch.is_grad_enabled():
            wb = [t.detach() for t in wb]
        feats = _TrunkFn.apply(x, *wb)                       # bf16 [B,1,4,4,512]
        from .engine import features_to_ncdhw
        h = features_to_ncdhw(feats, 512).reshape(-1, 8192)  # network.py:165 flattens in (C,T,H,W) order
        h = _LinearFn.apply(h, self.fc6.weight, self.fc6.bias, True)
        h = self.dropout(h)                                  # network.py:167 (identity in eval mode)
        h = h.reshape(bs, nc, -1).mean(1).reshape(bs, -1)    # network.py:174-176
        h = _LinearFn.apply(h, self.regressor.weight, self.regressor.bias, False)
        return _L2NormFn.apply(h)
