"""Multi-tensor Adam on the C ABI (zsv_adam_step / zsv_adam_pack_step), a drop-in for the reference's
``torch.optim.Adam`` (main.py:131).

Same update rule, same state layout (``step``, ``exp_avg``, ``exp_avg_sq`` per parameter, so ``state_dict()``
interchanges with torch.optim.Adam's).  Step counters AND learning rates live on the device, so the update can be captured
into the CUDA graph of the training iteration and a learning-rate schedule (main.py:133,374: MultiStepLR) still reaches
the replays (``GraphedStep`` refreshes the device scalars before every replay; nothing to call).

With ``model=`` (a ``network.Model`` / ``C3D`` of this package) the convolution weights are updated by
``zsv_adam_pack_step``: the same pass that writes p / exp_avg / exp_avg_sq also writes the bf16 weight images the next
forward and backward read (engine.PackedWeights), so the per-step re-pack of all fp32 master weights disappears.
``grad_scale`` multiplies every gradient first (1/world after a summed all-reduce, 1/scale of a GradScaler).
Optional: main.py works unchanged with torch's optimizer.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import AdamHyper, ConvDesc, check

_live: "weakref.WeakSet[FusedAdam]" = weakref.WeakSet()


def sync_all_lr() -> None:
    """Copy every live FusedAdam's ``group['lr']`` into its device scalar if it changed (outside any capture)."""
    for opt in list(_live):
        opt.sync_lr()


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 model: Optional[torch.nn.Module] = None, grad_scale: float = 1.0):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = float(grad_scale)
        self._lr_dev: Dict[int, torch.Tensor] = {}
        self._lr_host: Dict[int, float] = {}
        self._packed = None                    # engine.PackedWeights of `model`
        self._conv_index: Dict[int, int] = {}  # id(param) -> position in the packed plan
        self._model = model
        _live.add(self)

    # -- learning rate on the device -----------------------------------------------------------------------------
    def sync_lr(self) -> None:
        if torch.cuda.is_available() and torch.cuda.is_current_stream_capturing():
            return                              # a captured fill would pin the rate again
        for gi, group in enumerate(self.param_groups):
            lr = float(group["lr"])
            t = self._lr_dev.get(gi)
            if t is None:
                dev = next((p.device for p in group["params"] if p.is_cuda), None)
                if dev is None:
                    continue
                t = self._lr_dev[gi] = torch.full((), lr, dtype=torch.float32, device=dev)
                self._lr_host[gi] = lr
            elif self._lr_host[gi] != lr:
                t.fill_(lr)
                self._lr_host[gi] = lr

    # -- packed convolution weights of `model` ---------------------------------------------------------------------
    def _setup_packed(self) -> None:
        from . import engine
        model = self._model
        self._model = None
        if model is None:
            return
        if hasattr(model, "model") and hasattr(model.model, "arch"):          # network.Model: backbone convolutions
            arch = model.model.arch
            names, plan = engine._network_plan(1, 16, 112, 112, True, arch)
            lookup = dict(model.model.named_parameters())
            weights = [lookup[n + ".weight"] for n in names]
        elif hasattr(model, "packed_plan"):                                  # C3D
            plan, weights = model.packed_plan()
        else:
            raise RuntimeError("FusedAdam(model=...): expected a Model / C3D of this package")
        if not all(w.is_cuda for w in weights):
            raise RuntimeError("FusedAdam(model=...): move the model to its CUDA device first -- there is no CPU path")
        self._packed = engine.PackedWeights(plan, weights)
        self._packed.refresh()
        engine.publish_packed(self._packed)
        self._conv_index = {id(w): i for i, w in enumerate(weights)}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        from . import engine
        engine.note_weights_changed()        # parameters are written through raw pointers (no version-counter bump)
        if self._model is not None:
            self._setup_packed()
        self.sync_lr()
        stream = torch.cuda.current_stream().cuda_stream
        pw = self._packed
        if pw is not None and not pw.fresh():
            pw.refresh()                     # someone else changed a weight since the last step: images restart from it
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
            torch._foreach_add_([self.state[p]["step"] for p in ps], 1)
            b1, b2 = group["betas"]
            hyper = AdamHyper(float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                              float(group["weight_decay"]), self.grad_scale, self._lr_dev[gi].data_ptr())
            keep = []

            def grad_of(p):
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                keep.append(g)
                return g

            packed: List[torch.nn.Parameter] = []
            plain: List[torch.nn.Parameter] = []
            for p in ps:
                i = self._conv_index.get(id(p))
                if i is not None and p.is_contiguous() and pw.plan.convs[i].x_layout == _lib.X_NDHWC:
                    packed.append(p)
                else:
                    plain.append(p)
            if packed:
                n = len(packed)
                descs = (ConvDesc * n)()
                arr = lambda: (C.c_void_p * n)()
                pa, ga, ma, va, sa, wfa, wda = arr(), arr(), arr(), arr(), arr(), arr(), arr()
                base = pw.buf.data_ptr()
                for k, p in enumerate(packed):
                    i = self._conv_index[id(p)]
                    C.memmove(C.byref(descs[k]), C.byref(pw.plan.descs[i]), C.sizeof(ConvDesc))
                    st = self.state[p]
                    pa[k], ga[k], ma[k], va[k], sa[k] = p.data_ptr(), grad_of(p).data_ptr(), st["exp_avg"].data_ptr(), \
                        st["exp_avg_sq"].data_ptr(), st["step"].data_ptr()
                    wfa[k] = base + 2 * pw.plan.wf_off[i]
                    wda[k] = None if pw.plan.wd_off[i] is None else base + 2 * pw.plan.wd_off[i]
                check(lib.zsv_adam_pack_step(n, descs, pa, ga, ma, va, sa, wfa, wda, C.byref(hyper), stream),
                      "zsv_adam_pack_step")
            if plain:
                n = len(plain)
                arr = lambda: (C.c_void_p * n)()
                pa, ga, ma, va, sa, na = arr(), arr(), arr(), arr(), arr(), (C.c_longlong * n)()
                for k, p in enumerate(plain):
                    st = self.state[p]
                    pa[k], ga[k], ma[k], va[k], sa[k], na[k] = p.data_ptr(), grad_of(p).data_ptr(), \
                        st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), st["step"].data_ptr(), p.numel()
                check(lib.zsv_adam_step(n, pa, ga, ma, va, na, sa, C.byref(hyper), stream), "zsv_adam_step")
                # convolutions of the model that went through the plain update (the W-folded first layer): one small
                # re-pack into their slot of the published images
                for p in plain:
                    i = self._conv_index.get(id(p))
                    if i is not None:
                        conv = pw.plan.convs[i]
                        wd = pw.wds[i]
                        check(lib.zsv_conv3d_pack_weight(C.byref(conv.desc), p.data_ptr(), pw.wfs[i].data_ptr(),
                                                         None if wd is None else wd.data_ptr(), stream),
                              "zsv_conv3d_pack_weight")
        if pw is not None:
            pw.mark()
        return loss
