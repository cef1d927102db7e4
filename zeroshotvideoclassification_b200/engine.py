"""Whole-network execution of the R(2+1)D-18 hot path on libzsv_b200.so.

One ``torch.autograd.Function`` spans the backbone (resnet.py:243-249) and one the embedding head
(network.py:595-596); inside them the step is a fixed sequence of C-ABI kernel launches on the current
stream over channels-last bf16 buffers:

  forward : repack -> [conv fprop (+BN partial stats in the epilogue) -> BN finalize -> BN apply(+ReLU,+residual)]*
  backward: [BN backward (reduce, apply) -> wgrad -> dgrad(+residual gradient)]* in reverse order

``Tensor.backward()`` at main.py:195 therefore reaches exactly these kernels; parameter gradients come back
as fp32 tensors in the state-dict layout so GradScaler/Adam (main.py:195-203) work unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib, ops
from ._lib import cpad


# ----------------------------------------------------------------------------------------------------
# static description of the network (parameter names relative to the VideoResNet module)
# ----------------------------------------------------------------------------------------------------
@dataclass
class ConvSpec:
    name: str            # state-dict prefix of the nn.Conv3d, e.g. "layer1.0.conv1.0.0"
    bn: str              # prefix of the BatchNorm3d that follows it
    cin: int
    cout: int
    kernel: Tuple[int, int, int]
    stride: Tuple[int, int, int]
    padding: Tuple[int, int, int]


@dataclass
class BlockSpec:
    prefix: str
    convs: List[ConvSpec]                 # spatial1, temporal1, spatial2, temporal2
    downsample: Optional[ConvSpec]


def _r2plus1d_specs():
    stem = [ConvSpec("stem.0", "stem.1", 3, 45, (1, 7, 7), (1, 2, 2), (0, 3, 3)),
            ConvSpec("stem.3", "stem.4", 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0))]
    blocks = []
    cin = 64
    for li, cout in ((1, 64), (2, 128), (3, 256), (4, 512)):
        for bi in range(2):
            s = 2 if (li > 1 and bi == 0) else 1
            mid = (cin * cout * 27) // (cin * 9 + 3 * cout)       # resnet.py:91
            p = f"layer{li}.{bi}"
            convs = [ConvSpec(p + ".conv1.0.0", p + ".conv1.0.1", cin, mid, (1, 3, 3), (1, s, s), (0, 1, 1)),
                     ConvSpec(p + ".conv1.0.3", p + ".conv1.1", mid, cout, (3, 1, 1), (s, 1, 1), (1, 0, 0)),
                     ConvSpec(p + ".conv2.0.0", p + ".conv2.0.1", cout, mid, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
                     ConvSpec(p + ".conv2.0.3", p + ".conv2.1", mid, cout, (3, 1, 1), (1, 1, 1), (1, 0, 0))]
            ds = None
            if s != 1 or cin != cout:
                ds = ConvSpec(p + ".downsample.0", p + ".downsample.1", cin, cout, (1, 1, 1), (s, s, s), (0, 0, 0))
            blocks.append(BlockSpec(p, convs, ds))
            cin = cout
    return stem, blocks


def _r3d_specs():
    """r3d_18 (resnet.py:293-314): BasicStem 3x7x7 (resnet.py:165-173), BasicBlocks of two 3x3x3 Conv3DSimple
    (resnet.py:18-34), stride (s,s,s) in the first convolution of layers 2-4."""
    stem = [ConvSpec("stem.0", "stem.1", 3, 64, (3, 7, 7), (1, 2, 2), (1, 3, 3))]
    blocks = []
    cin = 64
    for li, cout in ((1, 64), (2, 128), (3, 256), (4, 512)):
        for bi in range(2):
            s = 2 if (li > 1 and bi == 0) else 1
            p = f"layer{li}.{bi}"
            convs = [ConvSpec(p + ".conv1.0", p + ".conv1.1", cin, cout, (3, 3, 3), (s, s, s), (1, 1, 1)),
                     ConvSpec(p + ".conv2.0", p + ".conv2.1", cout, cout, (3, 3, 3), (1, 1, 1), (1, 1, 1))]
            ds = None
            if s != 1 or cin != cout:
                ds = ConvSpec(p + ".downsample.0", p + ".downsample.1", cin, cout, (1, 1, 1), (s, s, s), (0, 0, 0))
            blocks.append(BlockSpec(p, convs, ds))
            cin = cout
    return stem, blocks


ARCH_SPECS = {"r2plus1d_18": _r2plus1d_specs(), "r3d_18": _r3d_specs()}
STEM_SPECS, BLOCK_SPECS = ARCH_SPECS["r2plus1d_18"]


def all_conv_specs(arch: str = "r2plus1d_18") -> List[ConvSpec]:
    stem, blocks = ARCH_SPECS[arch]
    out = list(stem)
    for b in blocks:
        out.extend(b.convs)
        if b.downsample is not None:
            out.append(b.downsample)
    return out


def param_names(arch: str = "r2plus1d_18") -> List[str]:
    """Differentiable parameters of the backbone in the order they are passed to the autograd Function."""
    names = []
    for c in all_conv_specs(arch):
        names += [c.name + ".weight", c.bn + ".weight", c.bn + ".bias"]
    return names


# ----------------------------------------------------------------------------------------------------
# geometry cache
# ----------------------------------------------------------------------------------------------------
_conv_cache: Dict[tuple, ops.Conv3d] = {}


def _conv_for(spec: ConvSpec, N, T, H, W, layout=_lib.X_NDHWC) -> ops.Conv3d:
    key = (N, T, H, W, spec.cin, spec.cout, spec.kernel, spec.stride, spec.padding, layout)
    op = _conv_cache.get(key)
    if op is None:
        op = ops.Conv3d(N, T, H, W, spec.cin, spec.cout, spec.kernel, spec.stride, spec.padding, layout)
        _conv_cache[key] = op
    return op


import os

# fuse the reduction pass of BatchNorm backward into the epilogue of the dgrad that produces its input gradient:
# 0 = never, 1 = only where the tile's main loop is long (the round-1 rule: spatial 1x3x3 dgrads), 2 = also the shallow
# dgrads of layers 2-4 (default since the fused epilogue became one shared-memory pass over the staged tile and a
# TMA-loaded y tile), 3 = every stride-1 dgrad
FUSE_BN_BWD = int(os.environ.get("ZSV_FUSE_BN_BWD", "2"))

# run the weight-gradient GEMMs on a second stream: they only depend on dy, nothing in the backward chain depends on
# them, and being tensor / L2 bound they overlap with the HBM-bound BatchNorm-backward kernels of the next layer
OVERLAP_WGRAD = os.environ.get("ZSV_OVERLAP_WGRAD", "1") != "0"
_side_streams: Dict[int, "torch.cuda.Stream"] = {}


def _side_stream(device) -> "torch.cuda.Stream":
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _side_streams.get(idx)
    if st is None:
        st = _side_streams[idx] = torch.cuda.Stream(device=device)
    return st


_plan_cache: Dict[tuple, tuple] = {}


def _network_plan(N, T, H, W, need_grad: bool, arch: str = "r2plus1d_18"):
    """Geometry of every convolution for a clip batch of this shape (in all_conv_specs() order) and the one-launch
    weight re-pack for them."""
    key = (N, T, H, W, need_grad, arch)
    hit = _plan_cache.get(key)
    if hit is not None:
        return hit
    convs, nd = {}, {}

    def add(spec, dims, layout=_lib.X_NDHWC, need_dgrad=True):
        op = _conv_for(spec, dims[0], dims[1], dims[2], dims[3], layout)
        convs[spec.name], nd[spec.name] = op, need_grad and need_dgrad
        return (dims[0], op.To, op.Ho, op.Wo)

    stem, blocks = ARCH_SPECS[arch]
    d = add(stem[0], (N, T, H, W), _lib.X_WFOLD, need_dgrad=False)
    for sp in stem[1:]:
        d = add(sp, d)
    for b in blocks:
        d_in = d
        for c in b.convs:
            d = add(c, d)
        if b.downsample is not None:
            add(b.downsample, d_in)
    names = [c.name for c in all_conv_specs(arch)]
    plan = ops.PackPlan([convs[n] for n in names], [nd[n] for n in names])
    _plan_cache[key] = (names, plan)
    return names, plan


_eval_pack_cache: Dict[tuple, tuple] = {}


class PackedWeights:
    """bf16 weight images of a network kept current by the optimizer instead of being re-packed by every forward.

    ``optim.FusedAdam(model=...)`` updates the fp32 master weights and writes their bf16 images in the same kernel
    (zsv_adam_pack_step) into ``buf``; the forward pass takes its operands from here while ``fresh()``: every weight
    still has the address and version counter recorded when the images were written.  Anything that changes a weight
    through PyTorch (load_state_dict, an in-place op, another optimizer) bumps the version and the next forward (or
    ``ensure_packed_fresh`` before a graph replay) re-packs."""

    def __init__(self, plan: "ops.PackPlan", weights: List[torch.Tensor]):
        self.plan, self.weights = plan, list(weights)
        self.buf = torch.empty(plan.total, dtype=torch.bfloat16, device=weights[0].device)
        self.wfs, self.wds = plan.views(self.buf)
        self.stamp = None

    def _now(self):
        return tuple((w.data_ptr(), w._version) for w in self.weights)

    def fresh(self) -> bool:
        return self.stamp is not None and self.stamp == self._now()

    def mark(self) -> None:
        self.stamp = self._now()

    def refresh(self) -> None:
        self.plan.pack([w.detach() for w in self.weights], out=self.buf)
        self.mark()


_published: Dict[int, PackedWeights] = {}      # key: data_ptr of the first convolution weight of the network


def publish_packed(pw: PackedWeights) -> None:
    _published[pw.weights[0].data_ptr()] = pw


def published_for(first_weight: torch.Tensor) -> Optional[PackedWeights]:
    pw = _published.get(first_weight.data_ptr())
    return pw if pw is not None and pw.fresh() else None


def ensure_packed_fresh() -> None:
    """Before a CUDA-graph replay: a captured forward reads the published images without any Python running, so images
    whose master weights were changed behind the optimizer's back are re-packed now."""
    for pw in _published.values():
        if not pw.fresh():
            pw.refresh()


def note_weights_changed() -> None:
    """Drop the folded inference weights.  Tensor version counters do not see writes made through raw pointers
    (zsv_bn_finalize updates running_mean / running_var, zsv_adam_step the parameters, a CUDA-graph replay does both
    without running any Python), so every such site calls this: a training-mode forward, FusedAdam.step and
    GraphedStep.__call__.  Re-folding is one launch per evaluate() (main.py:224-257)."""
    _eval_pack_cache.clear()


def _folded_weights(tensors: Dict[str, torch.Tensor], names, specs, plan, key):
    """bf16 weight images with the inference-time BatchNorm folded in, plus the per-channel biases.  evaluate()
    (main.py:224-257) runs many batches on fixed weights, so the result is cached until a parameter or a running
    statistic changes (tensor version counters) or moves."""
    src = []
    for c in specs:
        src += [tensors[c.name + ".weight"], tensors[c.bn + ".weight"], tensors[c.bn + ".bias"],
                tensors[c.bn + ".running_mean"], tensors[c.bn + ".running_var"]]
    stamp = tuple((t.data_ptr(), t._version) for t in src)
    hit = _eval_pack_cache.get(key)
    if hit is not None and hit[0] == stamp:
        return hit[1]
    wfs, biases = plan.pack_folded([tensors[c.name + ".weight"] for c in specs],
                                   [(tensors[c.bn + ".weight"], tensors[c.bn + ".bias"], tensors[c.bn + ".running_mean"],
                                     tensors[c.bn + ".running_var"]) for c in specs])
    packed = {n: (wf, b) for n, wf, b in zip(names, wfs, biases)}
    _eval_pack_cache[key] = (stamp, packed)
    return packed


# ----------------------------------------------------------------------------------------------------
# tape records
# ----------------------------------------------------------------------------------------------------
@dataclass
class UnitRec:
    """conv -> BN(train) [-> ReLU] with everything backward needs."""
    spec: ConvSpec
    op: ops.Conv3d
    x: torch.Tensor                 # conv input (bf16)
    wd: Optional[torch.Tensor]      # packed dgrad weights
    y: torch.Tensor                 # raw conv output (bf16)
    mean: torch.Tensor
    invstd: torch.Tensor
    out: Optional[torch.Tensor]     # post-activation output; None for a bare BN feeding a block tail
    relu: bool
    scale: Optional[torch.Tensor] = None   # forward scale / shift: backward recomputes the ReLU mask from y with them
    shift: Optional[torch.Tensor] = None
    table: Optional[torch.Tensor] = None   # (scale, shift, invstd, -mean*invstd) per channel for the fused BN backward


@dataclass
class BlockRec:
    units: List[UnitRec]            # spatial1, temporal1, spatial2, temporal2(tail, out=None)
    ds: Optional[UnitRec]
    out: torch.Tensor               # block output after add + ReLU
    x: torch.Tensor                 # block input


class BackboneRunner:
    """Executes the backbone given a flat dict of parameter / buffer tensors (names relative to VideoResNet)."""

    def __init__(self, tensors: Dict[str, torch.Tensor], train: bool, need_grad: bool, arch: str = "r2plus1d_18"):
        self.arch = arch
        self.stem_specs, self.block_specs = ARCH_SPECS[arch]
        self.t = tensors
        self.train = train
        self.need_grad = need_grad
        self.stem_recs: List[UnitRec] = []
        self.block_recs: List[BlockRec] = []
        self._side_keep: List[tuple] = []       # tensors in use by weight gradients running on the side stream
        self.arena: Optional[ops.GradArena] = None   # flat gradient buffer of the backward pass in progress
        self.packed: Dict[str, tuple] = {}      # bf16 weight images of the whole network, packed in one launch
        self.nbt: List[torch.Tensor] = []       # num_batches_tracked counters, bumped once per forward in one launch

    # -- forward building blocks -------------------------------------------------------------------
    def _unit(self, spec: ConvSpec, x: torch.Tensor, dims, relu: bool, apply_now: bool = True, layout=_lib.X_NDHWC,
              need_dgrad: bool = True):
        N, T, H, W = dims
        op = _conv_for(spec, N, T, H, W, layout)
        pk = self.packed.get(spec.name)
        if pk is None:       # a unit driven on its own (block-level tests): pack just this convolution
            pk = op.pack(self.t[spec.name + ".weight"], need_dgrad=self.need_grad and need_dgrad)
        wf, wd = pk
        gamma, beta = self.t[spec.bn + ".weight"], self.t[spec.bn + ".bias"]
        rm, rv = self.t[spec.bn + ".running_mean"], self.t[spec.bn + ".running_var"]
        table = None
        if self.train:
            y, ps, pq = op.fprop(x, wf, stats=True)
            scale, shift, mean, invstd, table = ops.bn_finalize(ps, pq, spec.cout, op.out_positions, gamma, beta, rm, rv,
                                                                want_table=True)
            nbt = self.t.get(spec.bn + ".num_batches_tracked")
            if nbt is not None:
                self.nbt.append(nbt)
        else:
            y, _, _ = op.fprop(x, wf, stats=False)
            scale, shift = ops.bn_eval_scale_shift(spec.cout, gamma, beta, rm, rv)
            mean = invstd = None
        out = ops.bn_apply(y, scale, shift, spec.cout, relu) if apply_now else None
        rec = UnitRec(spec, op, x, wd, y, mean, invstd, out, relu, scale, shift, table) if self.need_grad else None
        return out, y, (scale, shift), rec, (op.To, op.Ho, op.Wo)

    @staticmethod
    def _input(x: torch.Tensor, wpad_left: int):
        """fp32 NCDHW clips (main.py:167) or the bf16 W-folded tensor ops.clip_transform produced -> (folded, N, T, H, W)."""
        if x.dtype == torch.bfloat16 and x.dim() == 5 and x.shape[-1] == 8:
            N, T, H, Wp, _ = x.shape
            return x.contiguous(), N, T, H, Wp - 8
        N, _, T, H, W = x.shape
        return ops.repack_input(x, _lib.X_WFOLD, wpad_left), N, T, H, W

    def forward(self, x_ncdhw: torch.Tensor) -> torch.Tensor:
        folded, N, T, H, W = self._input(x_ncdhw, self.stem_specs[0].padding[2])
        names, plan = _network_plan(N, T, H, W, self.need_grad, self.arch)
        if not self.train and not self.need_grad:
            return self._forward_folded(folded, (N, T, H, W), names, plan)
        if self.train:
            note_weights_changed()       # this forward rewrites the running statistics through raw pointers
        pub = published_for(self.t[names[0] + ".weight"])
        if pub is not None:          # images written by the optimizer step (zsv_adam_pack_step): nothing to re-pack
            wfs, wds = pub.wfs, pub.wds
        else:
            wfs, wds = plan.pack([self.t[n + ".weight"] for n in names])
        self.packed = {n: (wf, wd) for n, wf, wd in zip(names, wfs, wds)}
        s0 = self.stem_specs[0]
        a, _, _, rec, d = self._unit(s0, folded, (N, T, H, W), True, layout=_lib.X_WFOLD, need_dgrad=False)
        self.stem_recs.append(rec)
        for sp in self.stem_specs[1:]:
            a, _, _, rec, d = self._unit(sp, a, (N, *d), True)
            self.stem_recs.append(rec)
        dims = (N, *d)
        for b in self.block_specs:
            a, dims = self._block(b, a, dims)
        if self.nbt:
            torch._foreach_add_(self.nbt, 1)
        return a

    # -- inference: BatchNorm folded into the convolutions ------------------------------------------
    def _forward_folded(self, folded: torch.Tensor, dims, names, plan) -> torch.Tensor:
        """model.eval() under no_grad (evaluate(), main.py:224-257): every conv -> BN -> (+shortcut) -> ReLU group is ONE
        kernel -- running-statistics BatchNorm folded into the packed weights and a bias, residual add and ReLU in the
        convolution's epilogue; no statistics, no tape, no separate normalisation passes."""
        N, T, H, W = dims
        specs = all_conv_specs(self.arch)
        key = (self.arch, N, T, H, W, str(folded.device), self.t[specs[0].name + ".weight"].data_ptr())
        packed = _folded_weights(self.t, names, specs, plan, key)

        def conv(spec, x, dims, relu, layout=_lib.X_NDHWC, addend=None):
            op = _conv_for(spec, dims[0], dims[1], dims[2], dims[3], layout)
            wf, bias = packed[spec.name]
            y, _, _ = op.fprop(x, wf, stats=False, bias=bias, relu=relu, addend=addend)
            return y, (dims[0], op.To, op.Ho, op.Wo)

        s0 = self.stem_specs[0]
        a, d = conv(s0, folded, (N, T, H, W), True, layout=_lib.X_WFOLD)
        for sp in self.stem_specs[1:]:
            a, d = conv(sp, a, d, True)
        for b in self.block_specs:
            x_in, d_in = a, d
            for c in b.convs[:-1]:
                a, d = conv(c, a, d, True)
            shortcut = x_in
            if b.downsample is not None:
                shortcut, _ = conv(b.downsample, x_in, d_in, False)
            a, d = conv(b.convs[-1], a, d, True, addend=shortcut)      # relu(bn(conv) + shortcut), resnet.py:110-111
        return a

    def _block(self, b: BlockSpec, x: torch.Tensor, dims):
        """BasicBlock.forward (resnet.py:102-113): every convolution but the last is conv -> BN -> ReLU; the last one's
        BatchNorm is applied together with the shortcut (identity or conv -> BN) and the final ReLU."""
        N = dims[0]
        recs = []
        a, d = x, dims[1:]
        for c in b.convs[:-1]:
            a, _, _, r, d = self._unit(c, a, (N, *d), True)
            recs.append(r)
        last = b.convs[-1]
        _, y_t, (sc, sh), r, d = self._unit(last, a, (N, *d), False, apply_now=False)
        recs.append(r)
        ds_rec = None
        if b.downsample is not None:
            _, y_d, (sc_d, sh_d), ds_rec, _ = self._unit(b.downsample, x, dims, False, apply_now=False)
            out = ops.bn_apply(y_t, sc, sh, last.cout, True, y2=y_d, scale2=sc_d, shift2=sh_d)
        else:
            out = ops.bn_apply(y_t, sc, sh, last.cout, True, residual=x)
        if self.need_grad:
            self.block_recs.append(BlockRec(recs, ds_rec, out, x))
        return out, (N, *d)

    # -- backward ----------------------------------------------------------------------------------
    def _conv_bwd(self, rec: UnitRec, dy, grads, want, addend=None, need_dx: bool = True,
                  producer: Optional[UnitRec] = None):
        """wgrad + dgrad of rec's convolution.  `producer` is the conv->BN(->ReLU) unit whose output is this
        convolution's input: when given, the dgrad epilogue also does the reduction pass of that BatchNorm's backward
        and the return value is (dz, partial, rows) for ``_unit_bwd`` instead of the plain gradient."""
        if want.get(rec.spec.name + ".weight", True):
            if OVERLAP_WGRAD:
                main = torch.cuda.current_stream(dy.device)
                side = _side_stream(dy.device)
                side.wait_stream(main)                     # dy (and x) are complete on the main stream
                dw_out = self._take(rec.spec.cout, rec.spec.cin, *rec.spec.kernel)    # allocated on the main stream
                with torch.cuda.stream(side):
                    dw, _ = rec.op.wgrad(rec.x, dy, out=dw_out)
                # x / dy were allocated on the main stream: keep them referenced until the main stream has joined
                # the side stream (join_side), so the caching allocator cannot hand their memory out early
                self._side_keep.append((rec.x, dy, dw))
                grads[rec.spec.name + ".weight"] = dw
            else:
                grads[rec.spec.name + ".weight"], _ = rec.op.wgrad(
                    rec.x, dy, out=self._take(rec.spec.cout, rec.spec.cin, *rec.spec.kernel))
        if not need_dx:
            return None
        # Round 1's fused epilogue (a shuffle reduction per 16-column chunk) only paid where the tile's main loop is long
        # compared with its epilogue (spatial 1x3x3: few output channels, deep reduction) and tripled the time of the
        # wide, shallow temporal dgrads; mode 1 keeps that rule for A/B runs.
        # With the one-pass epilogue (round 2) fusing wins wherever the tile's reduction is at least 384 deep; on the
        # shallow-K dgrads (temporal 144->64 and the stem's 45->64, K = 192) the kernel is epilogue-bound and the fused
        # epilogue costs more than the separate reduction pass (144->64: +140 us against 105 us; 45->64: 163 us fused
        # against 71 us + ~60 us; measured, profiles/r02_fuse_modes.txt): those keep the two-pass form.
        op = rec.op
        kdepth = op.cout * op.kernel[0] * op.kernel[1] * op.kernel[2]
        deep = kdepth >= 8 * op.cin
        fuse = FUSE_BN_BWD and op.stride == (1, 1, 1) and (deep or (FUSE_BN_BWD == 2 and kdepth >= 384) or FUSE_BN_BWD >= 3)
        if producer is not None and fuse:
            return rec.op.dgrad_bn_fused(dy, rec.wd, addend, producer.y, producer.table, producer.relu)
        return rec.op.dgrad(dy, rec.wd, addend)

    def join_side(self, device) -> None:
        """Main stream waits for the weight gradients issued so far on the side stream."""
        if self._side_keep:
            torch.cuda.current_stream(device).wait_stream(_side_stream(device))
            self._side_keep = []

    def _unit_bwd(self, rec: UnitRec, gin, grads, want, addend=None, need_dx: bool = True,
                  producer: Optional[UnitRec] = None):
        """gin: gradient w.r.t. rec.out (post BN/ReLU), or the (dz, partial, rows) triple of a fused dgrad.
        Returns the gradient w.r.t. rec.x (same convention, depending on `producer`)."""
        gamma = self.t[rec.spec.bn + ".weight"]
        if isinstance(gin, tuple):
            dz, partial, nrows = gin
            dy, dgamma, dbeta = ops.bn_bwd_finish(dz, rec.y, rec.mean, rec.invstd, gamma, partial, nrows, rec.spec.cout,
                                                  grad_out=self._take(2, rec.spec.cout))
        else:
            dy, _, _, dgamma, dbeta, _, _ = ops.bn_bwd(gin, None, 2 if rec.relu else 0, rec.y, rec.mean, rec.invstd,
                                                       gamma, rec.spec.cout, mask_scale=rec.scale, mask_shift=rec.shift,
                                                       grad_out=self._take(2, rec.spec.cout))
        grads[rec.spec.bn + ".weight"] = dgamma
        grads[rec.spec.bn + ".bias"] = dbeta
        return self._conv_bwd(rec, dy, grads, want, addend, need_dx, producer)

    def block_backward(self, brec: BlockRec, g, grads, want, producer: Optional[UnitRec] = None):
        """g: gradient w.r.t. the block output -> gradient w.r.t. the block input (resnet.py:102-113 reversed).
        `producer`: the unit that produced the block input when it is a plain conv->BN->ReLU (the stem)."""
        u = brec.units
        tail, ds = u[-1], brec.ds
        if ds is not None:
            dy_t, dy_d, _, dg, db, dg2, db2 = ops.bn_bwd(
                g, brec.out, True, tail.y, tail.mean, tail.invstd, self.t[tail.spec.bn + ".weight"],
                tail.spec.cout, y2=ds.y, mean2=ds.mean, invstd2=ds.invstd, gamma2=self.t[ds.spec.bn + ".weight"],
                grad_out=self._take(4, tail.spec.cout))
            grads[ds.spec.bn + ".weight"], grads[ds.spec.bn + ".bias"] = dg2, db2
            dz = None
        else:
            dy_t, _, dz, dg, db, _, _ = ops.bn_bwd(
                g, brec.out, True, tail.y, tail.mean, tail.invstd, self.t[tail.spec.bn + ".weight"],
                tail.spec.cout, want_dz=True, grad_out=self._take(2, tail.spec.cout))
        grads[tail.spec.bn + ".weight"], grads[tail.spec.bn + ".bias"] = dg, db
        ga = self._conv_bwd(tail, dy_t, grads, want, producer=u[-2])         # grad w.r.t. the previous unit's output
        for i in range(len(u) - 2, 0, -1):
            ga = self._unit_bwd(u[i], ga, grads, want, producer=u[i - 1])
        if ds is not None:
            gx = self._unit_bwd(u[0], ga, grads, want)                       # main branch
            return self._conv_bwd(ds, dy_d, grads, want, addend=gx)          # + projected shortcut
        return self._unit_bwd(u[0], ga, grads, want, addend=dz, producer=producer)   # + identity shortcut

    def _take(self, *shape) -> Optional[torch.Tensor]:
        """Slice of this backward pass's gradient arena (None outside ``backward``: the op allocates its own)."""
        return self.arena.take(*shape) if self.arena is not None else None

    def backward(self, g: torch.Tensor, want: Dict[str, bool]) -> Dict[str, torch.Tensor]:
        """g: gradient w.r.t. the backbone output (bf16 NDHWC).  Returns grads keyed by parameter name; all of them are
        views of ONE flat fp32 arena laid out in the order they are produced, so the data-parallel exchange all-reduces
        contiguous slices in place."""
        if not self.train:
            raise NotImplementedError("backward through eval-mode BatchNorm is not part of the hot path "
                                      "(the reference only back-propagates in train mode, main.py:143,195)")
        from . import dist as zdist
        sync = zdist.active_grad_sync()
        grads: Dict[str, torch.Tensor] = {}
        g = g.contiguous()
        # one slot per convolution: its weight, (dgamma, dbeta) of its BatchNorm (4 rows when two share a pass)
        numels = []
        for c in all_conv_specs(self.arch):
            numels += [c.cout * c.cin * c.kernel[0] * c.kernel[1] * c.kernel[2], 4 * c.cout]
        self.arena = ops.GradArena(g.device, numels)
        mark = 0
        for bi in range(len(self.block_recs) - 1, -1, -1):
            brec = self.block_recs[bi]
            # the first block's input is the stem's conv->BN->ReLU output: its BatchNorm backward is fused as well
            g = self.block_backward(brec, g, grads, want, producer=self.stem_recs[-1] if bi == 0 else None)
            dev = g[0].device if isinstance(g, tuple) else g.device
            if sync is not None:
                self.join_side(dev)
                sync.submit_range(self.arena.buf, mark, self.arena.off)   # overlaps the next block's backward
                mark = self.arena.off
            elif len(self._side_keep) > 12:
                self.join_side(dev)       # bound the tensors pinned for the side stream (single GPU)
        for si in range(len(self.stem_recs) - 1, 0, -1):
            g = self._unit_bwd(self.stem_recs[si], g, grads, want, producer=self.stem_recs[si - 1])
        self._unit_bwd(self.stem_recs[0], g, grads, want, need_dx=False)
        self.join_side(self.stem_recs[0].y.device)
        if sync is not None:
            sync.submit_range(self.arena.buf, mark, self.arena.off)
            sync.finish()                 # averaged in place: the views in `grads` now hold the global gradients
        self.arena = None
        return grads


# ----------------------------------------------------------------------------------------------------
# autograd glue
# ----------------------------------------------------------------------------------------------------
_BUFFER_SUFFIXES = (".running_mean", ".running_var", ".num_batches_tracked")


def _module_tensors(module: torch.nn.Module) -> Dict[str, torch.Tensor]:
    t = {k: v for k, v in module.named_parameters()}
    t.update({k: v for k, v in module.named_buffers()})
    return t


class _BackboneFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, *params):
        arch = getattr(module, "arch", "r2plus1d_18")
        names = param_names(arch)
        tensors = {k: v.detach() for k, v in _module_tensors(module).items()}
        for n, p in zip(names, params):
            tensors[n] = p.detach()
        need_grad = any(ctx.needs_input_grad[2:])
        runner = BackboneRunner(tensors, train=module.training, need_grad=need_grad, arch=arch)
        feats = runner.forward(x)
        ctx.runner = runner if need_grad else None
        ctx.consumed = False
        ctx.names = names
        ctx.want = {n: ctx.needs_input_grad[2 + i] for i, n in enumerate(names)}
        ctx.param_meta = [(p.dtype, p.shape) for p in params]
        return feats

    @staticmethod
    def backward(ctx, g):
        if ctx.runner is None:
            if ctx.consumed:
                raise RuntimeError("backbone backward called a second time: the activation tape of this forward was "
                                   "released by the first backward (retain_graph is not supported on this path; "
                                   "the reference back-propagates once per iteration, main.py:195)")
            return (None, None) + tuple(None for _ in ctx.names)
        grads = ctx.runner.backward(g, ctx.want)
        ctx.runner = None
        ctx.consumed = True
        out = []
        for n, (dt, shape) in zip(ctx.names, ctx.param_meta):
            gr = grads.get(n) if ctx.want[n] else None
            if gr is not None:
                gr = gr.reshape(shape)
                if gr.dtype != dt:
                    gr = gr.to(dt)
            out.append(gr)
        return (None, None) + tuple(out)


def backbone_forward(module: torch.nn.Module, x: torch.Tensor) -> torch.Tensor:
    """R(2+1)D-18 trunk on a ``VideoResNet18``-shaped module: x [B,3,T,H,W] fp32 CUDA -> bf16 [B,T',H',W',512]."""
    ops._require_cuda(x, "backbone_forward")
    _lib.load()
    lookup = dict(module.named_parameters())
    params = [lookup[n] for n in param_names(getattr(module, "arch", "r2plus1d_18"))]
    if not torch.is_grad_enabled():
        params = [p.detach() for p in params]
    return _BackboneFn.apply(module, x, *params)


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, w1, b1, w2, b2):
        channels = w1.shape[1]
        w1c, b1c, w2c, b2c = (t.detach().float().contiguous() for t in (w1, b1, w2, b2))
        emb, saved = ops.head_fwd(feats.detach().contiguous(), channels, w1c, b1c, w2c, b2c)
        ctx.save_for_backward(emb, *saved, w1c, w2c)
        ctx.feat_shape = tuple(feats.shape)
        ctx.channels = channels
        return emb

    @staticmethod
    def backward(ctx, demb):
        emb, pooled, hidden, onorm, w1c, w2c = ctx.saved_tensors
        need_w = any(ctx.needs_input_grad[1:])
        dw1, db1, dw2, db2, dfeat = ops.head_bwd(demb, emb, (pooled, hidden, onorm), ctx.feat_shape, ctx.channels,
                                                 w1c, w2c, need_wgrad=need_w, need_dfeat=ctx.needs_input_grad[0])
        return dfeat, dw1, db1, dw2, db2


def head_forward(feats, w1, b1, w2, b2) -> torch.Tensor:
    """mean-pool -> Linear -> ReLU -> Linear -> L2 normalise (network.py:595-596) -> fp32 [B,300]."""
    return _HeadFn.apply(feats, w1, b1, w2, b2)


class _ToNCDHW(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, channels):
        ctx.shape = tuple(feats.shape)
        return ops.ndhwc_to_ncdhw(feats.contiguous(), channels)

    @staticmethod
    def backward(ctx, g):
        return ops.ncdhw_to_ndhwc(g.contiguous()), None


def features_to_ncdhw(feats: torch.Tensor, channels: int = 512) -> torch.Tensor:
    return _ToNCDHW.apply(feats, channels)
