"""Multi-tensor Adam on the C ABI (zsv_adam_step), a drop-in for the reference's ``torch.optim.Adam`` (main.py:131).

Same update rule, same state layout (``step``, ``exp_avg``, ``exp_avg_sq`` per parameter, so ``state_dict()``
interchanges with torch.optim.Adam's), step counters on the device: the update is one or two kernel launches and can be
captured into the CUDA graph of the training iteration.  Optional: main.py works unchanged with torch's optimizer.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._tables = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        from . import engine
        engine.note_weights_changed()        # parameters are written through raw pointers (no version-counter bump)
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("FusedAdam: fp32 CUDA parameters and gradients only -- there is no CPU path")
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps = [self.state[p]["step"] for p in ps]
            torch._foreach_add_(steps, 1)
            n = len(ps)
            arr = lambda: (C.c_void_p * n)()
            pa, ga, ma, va, na = arr(), arr(), arr(), arr(), (C.c_longlong * n)()
            keep = []
            for i, p in enumerate(ps):
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                keep.append(g)
                st = self.state[p]
                pa[i], ga[i], ma[i], va[i], na[i] = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), \
                    st["exp_avg_sq"].data_ptr(), p.numel()
            b1, b2 = group["betas"]
            check(lib.zsv_adam_step(n, pa, ga, ma, va, na, steps[0].data_ptr(), float(group["lr"]), float(b1), float(b2),
                                    float(group["eps"]), float(group["weight_decay"]),
                                    torch.cuda.current_stream().cuda_stream), "zsv_adam_step")
        return loss
