"""Tensor-level wrappers over the C ABI (one Python function per entry point of include/zsv_b200.h).

Tensors are torch CUDA tensors used purely as typed device buffers: activations are bf16 ``[N,T,H,W,cpad(C)]``
(channels-last, pitch padded to 8), statistics and parameters fp32.  Every call enqueues on the current
torch CUDA stream and never synchronises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import BnBwdFuse, BnFold, ConvDesc, check, cpad, ptr

BN_EPS = 1e-5
BN_MOMENTUM = 0.1


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor -- this package has no CPU path "
                           "(the CPU oracle lives under oracle/ and is test infrastructure only)")


_workspace = {}
_profile = None   # when a list: (kind, conv, start_event, end_event) appended around every convolution call


def set_profile(sink) -> None:
    """bench.py: record CUDA events (on the launching stream) around each convolution call into ``sink``."""
    global _profile
    _profile = sink


class _Timed:
    def __init__(self, kind, conv):
        self.kind, self.conv = kind, conv

    def __enter__(self):
        if _profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _profile is not None and exc[0] is None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _profile.append((self.kind, self.conv, self.e0, e1))
        return False


def workspace(nbytes: int, device, tag: str = "default") -> torch.Tensor:
    """Grow-only scratch buffer per (device, tag); stream-ordered reuse on the current stream.  Buffers start zeroed:
    kernels that keep self-resetting ticket counters in their workspace (zsv_bn_finalize) rely on it."""
    key = (str(device), tag)
    buf = _workspace.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspace[key] = buf
    return buf


class GradArena:
    """One flat fp32 buffer for all parameter gradients of one backward pass, carved in the order the gradients are
    produced (deepest block first).  The data-parallel exchange all-reduces contiguous slices of it in place -- no
    gather copy, no copy back (dist.GradSync.submit_range) -- and every ``param.grad`` is a view of it."""

    ALIGN = 64      # elements: every tensor starts on a 256-byte boundary

    def __init__(self, device, numels):
        self.capacity = sum((n + self.ALIGN - 1) // self.ALIGN * self.ALIGN for n in numels)
        self.buf = torch.empty(max(self.capacity, 1), dtype=torch.float32, device=device)
        self.off = 0

    def take(self, *shape) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= d
        if self.off + n > self.capacity:
            raise RuntimeError("GradArena: more gradients than the plan announced")
        out = self.buf[self.off:self.off + n].view(*shape)
        self.off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        return out


class Conv3d:
    """One convolution of fixed geometry: descriptor + derived sizes.  Mirrors an ``nn.Conv3d`` of the
    reference (resnet.py:40-52,181-184,271; network.py:102-117)."""

    def __init__(self, N, T, H, W, cin, cout, kernel, stride, padding, x_layout=_lib.X_NDHWC):
        self.lib = _lib.load()
        self.desc = ConvDesc(N, T, H, W, cin, cout, *kernel, *stride, *padding, x_layout)
        out = (C.c_int32 * 3)()
        check(self.lib.zsv_conv3d_out_shape(C.byref(self.desc), out), "zsv_conv3d_out_shape")
        self.N, self.T, self.H, self.W, self.cin, self.cout = N, T, H, W, cin, cout
        self.To, self.Ho, self.Wo = int(out[0]), int(out[1]), int(out[2])
        self.kernel, self.stride, self.padding = tuple(kernel), tuple(stride), tuple(padding)
        self.x_layout = x_layout
        self.stat_rows = self.lib.zsv_conv3d_stat_rows(C.byref(self.desc))
        self.wf_bytes = self.lib.zsv_conv3d_packed_weight_bytes(C.byref(self.desc), 0)
        self.wd_bytes = self.lib.zsv_conv3d_packed_weight_bytes(C.byref(self.desc), 1)
        self.wgrad_ws = self.lib.zsv_conv3d_wgrad_workspace(C.byref(self.desc))
        self.out_positions = N * self.To * self.Ho * self.Wo

    # -- weights ---------------------------------------------------------------------------------
    def pack(self, w: torch.Tensor, need_dgrad: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        _require_cuda(w, "Conv3d.pack")
        w = w.detach()
        if w.dtype != torch.float32 or not w.is_contiguous():
            w = w.float().contiguous()
        wf = torch.empty(self.wf_bytes // 2, dtype=torch.bfloat16, device=w.device)
        wd = None
        if need_dgrad and self.wd_bytes:
            wd = torch.empty(self.wd_bytes // 2, dtype=torch.bfloat16, device=w.device)
        check(self.lib.zsv_conv3d_pack_weight(C.byref(self.desc), ptr(w), ptr(wf), ptr(wd), _stream()),
              "zsv_conv3d_pack_weight")
        return wf, wd

    # -- forward ---------------------------------------------------------------------------------
    def fprop(self, x, wf, stats: bool = True, bias=None, relu: bool = False, addend=None):
        y = torch.empty((self.N, self.To, self.Ho, self.Wo, cpad(self.cout)), dtype=torch.bfloat16, device=x.device)
        ps = pq = None
        if stats:
            ps = torch.empty((self.stat_rows, cpad(self.cout)), dtype=torch.float32, device=x.device)
            pq = torch.empty_like(ps)
        with _Timed("fprop", self):
            check(self.lib.zsv_conv3d_fprop(C.byref(self.desc), ptr(x), ptr(wf), ptr(y), ptr(ps), ptr(pq), ptr(bias),
                                            ptr(addend), int(relu), _stream()), "zsv_conv3d_fprop")
        return y, ps, pq

    # -- backward --------------------------------------------------------------------------------
    def dgrad(self, dy, wd, addend=None):
        dx = torch.empty((self.N, self.T, self.H, self.W, cpad(self.cin)), dtype=torch.bfloat16, device=dy.device)
        with _Timed("dgrad", self):
            check(self.lib.zsv_conv3d_dgrad(C.byref(self.desc), ptr(dy), ptr(wd), ptr(dx), ptr(addend), None, _stream()),
                  "zsv_conv3d_dgrad")
        return dx

    def dgrad_bn_fused(self, dy, wd, addend, y_prev, bn_table, relu: bool):
        """dgrad whose epilogue also does the first pass of the backward of the BatchNorm (+ReLU) that produced this
        convolution's input (zsv_bn_bwd_fuse).  Returns (dz, partial, rows): dz = g * ReLU mask (bf16, shape of the
        input), partial = fp32 [rows][4][cpad(Cin)] per-CTA sums for ``bn_bwd_finish``."""
        dz = torch.empty((self.N, self.T, self.H, self.W, cpad(self.cin)), dtype=torch.bfloat16, device=dy.device)
        cap = 8 * self.lib.zsv_sm_count()
        partial = torch.empty((cap, 4, cpad(self.cin)), dtype=torch.float32, device=dy.device)
        f = BnBwdFuse(ptr(y_prev), ptr(bn_table), int(relu), ptr(partial), cap, 0)
        with _Timed("dgrad", self):
            check(self.lib.zsv_conv3d_dgrad(C.byref(self.desc), ptr(dy), ptr(wd), ptr(dz), ptr(addend), C.byref(f),
                                            _stream()), "zsv_conv3d_dgrad")
        return dz, partial, int(f.rows_written)

    def wgrad(self, x, dy, want_bias: bool = False, out=None):
        dw = out if out is not None else torch.empty((self.cout, self.cin, *self.kernel), dtype=torch.float32,
                                                     device=dy.device)
        ws = workspace(self.wgrad_ws, dy.device, "wgrad")
        with _Timed("wgrad", self):
            check(self.lib.zsv_conv3d_wgrad(C.byref(self.desc), ptr(x), ptr(dy), ptr(dw), ptr(ws), ws.numel(),
                                            _stream()), "zsv_conv3d_wgrad")
        db = bias_grad(dy, self.cout) if want_bias else None
        return dw, db


class PackPlan:
    """Re-packs the weights of a fixed list of convolutions in ONE kernel launch (zsv_conv3d_pack_weights): after
    every optimizer step (main.py:203) all 37 fp32 master weights of R(2+1)D-18 need fresh bf16 images."""

    def __init__(self, convs, need_dgrad):
        self.lib = _lib.load()
        self.convs = list(convs)
        n = self.n = len(self.convs)
        self.descs = (ConvDesc * n)()
        for i, c in enumerate(self.convs):
            C.memmove(C.byref(self.descs[i]), C.byref(c.desc), C.sizeof(ConvDesc))
        self.wf_off, self.wd_off, off = [], [], 0
        for c, nd in zip(self.convs, need_dgrad):
            self.wf_off.append(off)
            off += (c.wf_bytes // 2 + 7) & ~7                 # 16-byte aligned images (TMA global address alignment)
            if nd and c.wd_bytes:
                self.wd_off.append(off)
                off += (c.wd_bytes // 2 + 7) & ~7
            else:
                self.wd_off.append(None)
        self.total = off

    def _arrays(self):
        # per call: a plan is cached globally and may be used from several threads / devices at once
        n = self.n
        return (C.c_void_p * n)(), (C.c_void_p * n)(), (C.c_void_p * n)()

    def views(self, buf):
        """([wf...], [wd or None...]) views of a packed buffer laid out by this plan."""
        wfs = [buf[o:o + c.wf_bytes // 2] for o, c in zip(self.wf_off, self.convs)]
        wds = [None if o is None else buf[o:o + c.wd_bytes // 2] for o, c in zip(self.wd_off, self.convs)]
        return wfs, wds

    def pack(self, weights, out=None):
        """weights: fp32 CUDA tensors in conv order -> ([wf...], [wd or None...]) views of one bf16 buffer (``out`` or
        a fresh one)."""
        dev = weights[0].device
        buf = out if out is not None else torch.empty(self.total, dtype=torch.bfloat16, device=dev)
        base = buf.data_ptr()
        keep = []
        w_arr, wf_arr, wd_arr = self._arrays()
        for i, w in enumerate(weights):
            _require_cuda(w, "PackPlan.pack")
            w = w.detach()
            if w.dtype != torch.float32 or not w.is_contiguous():
                w = w.float().contiguous()
                keep.append(w)
            w_arr[i] = w.data_ptr()
            wf_arr[i] = base + 2 * self.wf_off[i]
            wd_arr[i] = None if self.wd_off[i] is None else base + 2 * self.wd_off[i]
        check(self.lib.zsv_conv3d_pack_weights(self.n, self.descs, w_arr, wf_arr, wd_arr, _stream()),
              "zsv_conv3d_pack_weights")
        return self.views(buf)

    def pack_folded(self, weights, bns, eps: float = BN_EPS):
        """Inference: weights with the BatchNorm that follows each convolution folded in (running statistics).
        bns: (gamma, beta, running_mean, running_var) per convolution.  -> ([wf...], [bias fp32 [Cout]...])."""
        dev = weights[0].device
        buf = torch.empty(self.total, dtype=torch.bfloat16, device=dev)
        nb = [c.cout for c in self.convs]
        biases = torch.empty(sum(nb), dtype=torch.float32, device=dev)
        base, keep, off = buf.data_ptr(), [], 0
        folds = (BnFold * self.n)()
        out_b = []
        w_arr, wf_arr, _ = self._arrays()
        for i, (w, (gamma, beta, rm, rv)) in enumerate(zip(weights, bns)):
            _require_cuda(w, "PackPlan.pack_folded")
            tens = []
            for t in (w, gamma, beta, rm, rv):
                t = t.detach()
                if t.dtype != torch.float32 or not t.is_contiguous():
                    t = t.float().contiguous()
                    keep.append(t)
                tens.append(t)
            w_arr[i] = tens[0].data_ptr()
            wf_arr[i] = base + 2 * self.wf_off[i]
            b = biases[off:off + nb[i]]
            off += nb[i]
            out_b.append(b)
            folds[i] = BnFold(tens[1].data_ptr(), tens[2].data_ptr(), tens[3].data_ptr(), tens[4].data_ptr(), b.data_ptr(), eps)
        check(self.lib.zsv_conv3d_pack_weights_folded(self.n, self.descs, w_arr, wf_arr, folds, _stream()),
              "zsv_conv3d_pack_weights_folded")
        wfs = [buf[o:o + c.wf_bytes // 2] for o, c in zip(self.wf_off, self.convs)]
        return wfs, out_b


# ------------------------------------------------------------------------------------------------
# layout
# ------------------------------------------------------------------------------------------------
def repack_input(x: torch.Tensor, layout: int = _lib.X_NDHWC, wpad_left: int = 0) -> torch.Tensor:
    """fp32 NCDHW -> bf16 channels-last (network.py:534-535 reshapes, main.py:167 uploads)."""
    _require_cuda(x, "repack_input")
    lib = _lib.load()
    x = x.detach()
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.float().contiguous()
    N, Cc, T, H, W = x.shape
    Wp = W + 8 if layout == _lib.X_WFOLD else W
    out = torch.empty((N, T, H, Wp, cpad(Cc)), dtype=torch.bfloat16, device=x.device)
    check(lib.zsv_repack_input(ptr(x), ptr(out), N, Cc, T, H, W, layout, wpad_left, _stream()), "zsv_repack_input")
    return out


def clip_transform(frames: torch.Tensor, crop_ij: torch.Tensor, flip: Optional[torch.Tensor] = None,
                   resize_short: int = 128, crop: int = 112, wpad_left: int = 3) -> torch.Tensor:
    """uint8 frames [N,T,Hs,Ws,3] (CUDA) -> the transform of auxiliary/transforms.py:41-56 (normalise, Resize(128),
    crop 112 at crop_ij[n] = (i, j), optional flip) -> bf16 W-folded clips [N,T,112,120,8] that ``Model.forward``
    accepts in place of the fp32 ``[B,nc,3,T,H,W]`` batch."""
    _require_cuda(frames, "clip_transform")
    lib = _lib.load()
    if frames.dtype != torch.uint8 or frames.dim() != 5 or frames.shape[-1] != 3:
        raise RuntimeError("clip_transform: expected uint8 frames [N,T,H,W,3]")
    frames = frames.contiguous()
    N, T, Hs, Ws, _ = frames.shape
    crop_ij = crop_ij.to(device=frames.device, dtype=torch.int32).contiguous()
    if flip is not None:
        flip = flip.to(device=frames.device, dtype=torch.uint8).contiguous()
    out = torch.empty((N, T, crop, crop + 8, 8), dtype=torch.bfloat16, device=frames.device)
    check(lib.zsv_clip_transform(ptr(frames), ptr(out), N, T, Hs, Ws, resize_short, crop, ptr(crop_ij), ptr(flip),
                                 wpad_left, _stream()), "zsv_clip_transform")
    return out


def ndhwc_to_ncdhw(x: torch.Tensor, channels: int) -> torch.Tensor:
    lib = _lib.load()
    N, T, H, W, _ = x.shape
    out = torch.empty((N, channels, T, H, W), dtype=torch.float32, device=x.device)
    check(lib.zsv_ndhwc_to_ncdhw(ptr(x), ptr(out), N, channels, T, H, W, _stream()), "zsv_ndhwc_to_ncdhw")
    return out


def ncdhw_to_ndhwc(x: torch.Tensor) -> torch.Tensor:
    return repack_input(x, _lib.X_NDHWC, 0)


# ------------------------------------------------------------------------------------------------
# batch norm
# ------------------------------------------------------------------------------------------------
def bn_finalize(ps, pq, channels: int, count: int, gamma, beta, running_mean, running_var,
                momentum: float = BN_MOMENTUM, eps: float = BN_EPS, want_table: bool = False):
    """-> (scale, shift, mean, invstd[, table]); table = fp32 [cpad(C)][4] constants for the fused BN backward."""
    lib = _lib.load()
    cp = cpad(channels)
    out = torch.empty((8 if want_table else 4, cp), dtype=torch.float32, device=ps.device)
    scale, shift, mean, invstd = out[0], out[1], out[2], out[3]
    table = out[4:8].view(cp, 4) if want_table else None
    ws = workspace(lib.zsv_bn_finalize_workspace(channels), ps.device, "bn_finalize")
    check(lib.zsv_bn_finalize(ptr(ps), ptr(pq), ps.shape[0], channels, count, ptr(gamma), ptr(beta),
                              ptr(running_mean), ptr(running_var), momentum, eps, ptr(scale), ptr(shift), ptr(mean),
                              ptr(invstd), ptr(table), ptr(ws), ws.numel(), _stream()), "zsv_bn_finalize")
    if want_table:
        return scale, shift, mean, invstd, table
    return scale, shift, mean, invstd


def bn_bwd_finish(dz, y, mean, invstd, gamma, partial, partial_rows: int, channels: int, grad_out=None):
    """Second pass of BatchNorm backward from the sums a fused dgrad left behind -> (dy, dgamma, dbeta).
    grad_out: optional fp32 [2][channels] buffer that receives (dbeta, dgamma) (a slice of the gradient arena)."""
    lib = _lib.load()
    rows = y.numel() // y.shape[-1]
    dy = torch.empty_like(y)
    dgb = grad_out if grad_out is not None else torch.empty((2, channels), dtype=torch.float32, device=y.device)
    ws = workspace(4 * cpad(channels) * 4, y.device, "bn_bwd_finish")
    check(lib.zsv_bn_bwd_finish(ptr(dz), ptr(y), ptr(mean), ptr(invstd), ptr(gamma), ptr(partial), partial_rows, ptr(dy),
                                ptr(dgb[1]), ptr(dgb[0]), rows, channels, ptr(ws), ws.numel(), _stream()),
          "zsv_bn_bwd_finish")
    return dy, dgb[1], dgb[0]


def bn_eval_scale_shift(channels: int, gamma, beta, running_mean, running_var, eps: float = BN_EPS):
    lib = _lib.load()
    cp = cpad(channels)
    out = torch.empty((2, cp), dtype=torch.float32, device=running_mean.device)
    check(lib.zsv_bn_eval_scale_shift(channels, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), eps,
                                      ptr(out[0]), ptr(out[1]), _stream()), "zsv_bn_eval_scale_shift")
    return out[0], out[1]


def bn_apply(y, scale, shift, channels: int, relu: bool, y2=None, scale2=None, shift2=None, residual=None):
    lib = _lib.load()
    out = torch.empty_like(y)
    rows = y.numel() // y.shape[-1]
    check(lib.zsv_bn_apply(ptr(y), ptr(scale), ptr(shift), ptr(y2), ptr(scale2), ptr(shift2), ptr(residual), ptr(out),
                           rows, channels, int(relu), _stream()), "zsv_bn_apply")
    return out


def bn_bwd(g, out, relu, y, mean, invstd, gamma, channels: int, y2=None, mean2=None, invstd2=None, gamma2=None,
           want_dz: bool = False, mask_scale=None, mask_shift=None, grad_out=None):
    """Returns (dy, dy2, dz, dgamma, dbeta, dgamma2, dbeta2).  relu: False/0, True/1 (mask from `out`) or 2 (mask
    recomputed from y with the forward's scale/shift)."""
    relu = int(relu)
    lib = _lib.load()
    rows = y.numel() // y.shape[-1]
    dy = torch.empty_like(y)
    dy2 = torch.empty_like(y2) if y2 is not None else None
    dz = torch.empty_like(y) if want_dz else None
    # (dgamma, dbeta[, dgamma2, dbeta2]) rows; grad_out: a slice of the gradient arena with that many rows
    dgb = grad_out if grad_out is not None else torch.empty((4, channels), dtype=torch.float32, device=y.device)
    ws_bytes = lib.zsv_bn_bwd_workspace(channels)
    ws = workspace(ws_bytes, y.device, "bn_bwd")
    has2 = y2 is not None
    check(lib.zsv_bn_bwd(ptr(g), ptr(out) if relu == 1 else None, relu, ptr(mask_scale), ptr(mask_shift), ptr(y),
                         ptr(mean), ptr(invstd), ptr(gamma),
                         ptr(y2), ptr(mean2), ptr(invstd2), ptr(gamma2), ptr(dy), ptr(dy2), ptr(dz), ptr(dgb[0]),
                         ptr(dgb[1]), ptr(dgb[2]) if has2 else None, ptr(dgb[3]) if has2 else None, rows, channels,
                         ptr(ws), ws.numel(), _stream()), "zsv_bn_bwd")
    return dy, dy2, dz, dgb[0], dgb[1], (dgb[2] if has2 else None), (dgb[3] if has2 else None)


# ------------------------------------------------------------------------------------------------
# head / loss / nearest class / pooling
# ------------------------------------------------------------------------------------------------
def head_fwd(feat, channels: int, w1, b1, w2, b2, eps: float = 1e-12):
    lib = _lib.load()
    B = feat.shape[0]
    P = feat.numel() // (B * feat.shape[-1])
    Hd, E = w1.shape[0], w2.shape[0]
    dev = feat.device
    pooled = torch.empty((B, channels), dtype=torch.float32, device=dev)
    hidden = torch.empty((B, Hd), dtype=torch.float32, device=dev)
    onorm = torch.empty((B,), dtype=torch.float32, device=dev)
    emb = torch.empty((B, E), dtype=torch.float32, device=dev)
    check(lib.zsv_head_fwd(ptr(feat), B, P, channels, ptr(w1), ptr(b1), Hd, ptr(w2), ptr(b2), E, eps, ptr(pooled),
                           ptr(hidden), ptr(onorm), ptr(emb), _stream()), "zsv_head_fwd")
    return emb, (pooled, hidden, onorm)


def head_bwd(demb, emb, saved, feat_shape, channels: int, w1, w2, need_wgrad: bool = True, need_dfeat: bool = True,
             eps: float = 1e-12):
    lib = _lib.load()
    pooled, hidden, onorm = saved
    B, E = emb.shape
    Hd = w1.shape[0]
    P = 1
    for s in feat_shape[1:-1]:
        P *= s
    dev = emb.device
    dw1 = torch.empty_like(w1) if need_wgrad else None
    db1 = torch.empty(Hd, dtype=torch.float32, device=dev) if need_wgrad else None
    dw2 = torch.empty_like(w2) if need_wgrad else None
    db2 = torch.empty(E, dtype=torch.float32, device=dev) if need_wgrad else None
    dfeat = torch.empty(feat_shape, dtype=torch.bfloat16, device=dev) if need_dfeat else None
    scratch = torch.empty(lib.zsv_head_bwd_scratch(B, channels, Hd, E), dtype=torch.uint8, device=dev)
    demb = demb.float().contiguous()
    check(lib.zsv_head_bwd(ptr(demb), ptr(emb), ptr(onorm), ptr(pooled), ptr(hidden), B, P, channels, ptr(w1), Hd,
                           ptr(w2), E, eps, ptr(dw1), ptr(db1), ptr(dw2), ptr(db2), ptr(dfeat), ptr(scratch),
                           scratch.numel(), _stream()), "zsv_head_bwd")
    return dw1, db1, dw2, db2, dfeat


def mse_fwd_bwd(emb, target, grad_scale: float = 1.0, want_grad: bool = True):
    lib = _lib.load()
    B, E = emb.shape
    loss = torch.empty(1, dtype=torch.float32, device=emb.device)
    demb = torch.empty_like(emb) if want_grad else None
    check(lib.zsv_mse_fwd_bwd(ptr(emb), ptr(target), B, E, grad_scale, ptr(loss), ptr(demb), _stream()),
          "zsv_mse_fwd_bwd")
    return loss, demb


def nearest_class(emb: torch.Tensor, cls: torch.Tensor, k: int = 1, return_dist: bool = False):
    """int64 [N,k] indices of the k nearest class vectors by cosine distance (main.py:183, 321-322)."""
    _require_cuda(emb, "nearest_class")
    lib = _lib.load()
    emb = emb.detach().float().contiguous()
    cls = cls.detach().float().contiguous().to(emb.device)
    N, D = emb.shape
    Cn = cls.shape[0]
    idx = torch.empty((N, k), dtype=torch.int64, device=emb.device)
    dist = torch.empty((N, k), dtype=torch.float64, device=emb.device) if return_dist else None
    if N == 0:   # empty batch (every sample of a batch was filtered, main.py:157-158)
        return (idx, dist) if return_dist else idx
    check(lib.zsv_nearest_class(ptr(emb), ptr(cls), N, Cn, D, k, ptr(idx), ptr(dist), _stream()),
          "zsv_nearest_class")
    return (idx, dist) if return_dist else idx


def maxpool3d_fwd(x, channels: int, kernel, padding=(0, 0, 0)):
    lib = _lib.load()
    N, T, H, W, cp = x.shape
    kt, kh, kw = kernel
    pt, ph, pw = padding
    To, Ho, Wo = (T + 2 * pt - kt) // kt + 1, (H + 2 * ph - kh) // kh + 1, (W + 2 * pw - kw) // kw + 1
    y = torch.empty((N, To, Ho, Wo, cp), dtype=torch.bfloat16, device=x.device)
    am = torch.empty((N, To, Ho, Wo, cp), dtype=torch.uint8, device=x.device)
    check(lib.zsv_maxpool3d_fwd(ptr(x), ptr(y), ptr(am), N, T, H, W, channels, kt, kh, kw, pt, ph, pw, _stream()),
          "zsv_maxpool3d_fwd")
    return y, am


def maxpool3d_bwd(dy, argmax, in_shape, channels: int, kernel, padding=(0, 0, 0), relu_pooled=None,
                  want_bias: bool = False):
    """-> dx, or (dx, db) with want_bias.  relu_pooled: the pooling output when the pooled tensor was a ReLU output (its
    ReLU backward is fused); db: bias gradient of the convolution in front of that ReLU, from the same pass."""
    lib = _lib.load()
    N, T, H, W, cp = in_shape
    kt, kh, kw = kernel
    pt, ph, pw = padding
    dx = torch.empty(in_shape, dtype=torch.bfloat16, device=dy.device)
    db = ws = None
    if want_bias:
        db = torch.empty(channels, dtype=torch.float32, device=dy.device)
        ws = workspace(lib.zsv_bias_grad_workspace(channels), dy.device, "bias_grad")
    check(lib.zsv_maxpool3d_bwd(ptr(dy), ptr(argmax), ptr(relu_pooled), ptr(dx), N, T, H, W, channels, kt, kh, kw,
                                pt, ph, pw, ptr(db), ptr(ws), ws.numel() if ws is not None else 0, _stream()),
          "zsv_maxpool3d_bwd")
    return (dx, db) if want_bias else dx


def relu_bwd(g, out, channels: int, want_bias: bool = False):
    lib = _lib.load()
    dz = torch.empty_like(g)
    rows = g.numel() // g.shape[-1]
    db = ws = None
    if want_bias:
        db = torch.empty(channels, dtype=torch.float32, device=g.device)
        ws = workspace(lib.zsv_bias_grad_workspace(channels), g.device, "bias_grad")
    check(lib.zsv_relu_bwd(ptr(g), ptr(out), ptr(dz), rows, channels, ptr(db), ptr(ws),
                           ws.numel() if ws is not None else 0, _stream()), "zsv_relu_bwd")
    return (dz, db) if want_bias else dz


def bias_grad(dy, channels: int):
    lib = _lib.load()
    rows = dy.numel() // dy.shape[-1]
    db = torch.empty(channels, dtype=torch.float32, device=dy.device)
    ws = workspace(lib.zsv_bias_grad_workspace(channels), dy.device, "bias_grad")
    check(lib.zsv_bias_grad(ptr(dy), ptr(db), rows, channels, ptr(ws), ws.numel(), _stream()), "zsv_bias_grad")
    return db


def linear_fwd(x, w, b, relu: bool):
    lib = _lib.load()
    B, K = x.shape
    J = w.shape[0]
    out = torch.empty((B, J), dtype=torch.float32, device=x.device)
    ws = workspace(lib.zsv_linear_workspace(B, K, J), x.device, "linear")
    check(lib.zsv_linear_fwd(ptr(x), ptr(w), ptr(b), ptr(out), B, K, J, int(relu), ptr(ws), ws.numel(), _stream()),
          "zsv_linear_fwd")
    return out


def linear_bwd(dy, x, w, act=None, need_dx=True, need_dw=True):
    lib = _lib.load()
    B, K = x.shape
    J = w.shape[0]
    dev = x.device
    dx = torch.empty((B, K), dtype=torch.float32, device=dev) if need_dx else None
    dw = torch.empty((J, K), dtype=torch.float32, device=dev) if need_dw else None
    db = torch.empty((J,), dtype=torch.float32, device=dev) if need_dw else None
    ws = workspace(lib.zsv_linear_workspace(B, K, J), dev, "linear")
    check(lib.zsv_linear_bwd(ptr(dy), ptr(x), ptr(w), ptr(act), B, K, J, ptr(dx), ptr(dw), ptr(db), ptr(ws), ws.numel(),
                             _stream()), "zsv_linear_bwd")
    return dx, dw, db


def l2norm_fwd(o, eps: float = 1e-12):
    lib = _lib.load()
    B, E = o.shape
    emb = torch.empty_like(o)
    onorm = torch.empty((B,), dtype=torch.float32, device=o.device)
    check(lib.zsv_l2norm_fwd(ptr(o), ptr(emb), ptr(onorm), B, E, eps, _stream()), "zsv_l2norm_fwd")
    return emb, onorm


def l2norm_bwd(demb, emb, onorm, eps: float = 1e-12):
    lib = _lib.load()
    B, E = emb.shape
    dout = torch.empty_like(emb)
    check(lib.zsv_l2norm_bwd(ptr(demb), ptr(emb), ptr(onorm), ptr(dout), B, E, eps, _stream()), "zsv_l2norm_bwd")
    return dout
