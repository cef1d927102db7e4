"""Teacher-forced per-layer parity at BASELINE.json config-1 shapes (bs = 2 x 3 x 16 x 112 x 112).

north_star: "per-layer activations and gradients match within 1e-2 relative error (bf16 compute, fp32 accumulate)".
End to end a randomly initialised 37-BatchNorm network amplifies one-ulp differences chaotically, so
the claim is tested where it is well defined: EVERY convolution and EVERY BatchNorm of the network is fed the fp32
oracle's own input and the oracle's own upstream gradient (a trace of one full forward + backward of the model on the
CPU, oracle/video_oracle.py, itself pinned to the reference by tests/golden), and each kernel's output must be within
1e-2 of the oracle's (max |a-b| / max |b|; weight gradients 5e-3, see TOL_WGRAD).  The same for the 8 convolutions of C3D, and bs = 22
full-shape cases for the dgrad / wgrad of layer 1 and layer 4.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import video_oracle as vo
from oracle.make_golden import synthetic_batch
from tests.helpers import cpad, from_ndhwc, rel_err, to_ndhwc

pytestmark = pytest.mark.gpu

TOL = 1e-2          # north_star: activations and gradients
# Weight gradients.  On operands that are ALREADY bf16 on both sides (tests/test_gpu_conv.py, the bs = 22 cases below) the
# kernel is within 2e-3 of fp32 autograd: that is the fp32-accumulation error.  Here the oracle's operands are fp32 and
# the kernel sees them rounded to bf16 (its storage format): dW = sum over 1e4..1e6 positions of x * dy with random
# signs is ~sqrt(N) * sigma, and the two roundings (2^-9 each, uncorrelated) leave ~2^-9 * sqrt(2) = 2.8e-3 of that
# scale whatever N is -- measured 2.3e-3 .. 3.5e-3 (max-norm) on every layer of both networks.  The gate is therefore
# 5e-3: well inside north_star's 1e-2 and just above the rounding floor of the storage format.
TOL_WGRAD = 5e-3
TOL_WGRAD_BF16_OPERANDS = 2e-3


def _wgrad_tol(cin, cout, kernel):
    return TOL_WGRAD


@pytest.fixture(scope="module")
def r2plus1d_trace():
    from zeroshotvideoclassification_b200 import video_models as vm
    torch.manual_seed(0)
    model = vm.get_network(vm.default_opt("r2plus1d_18"))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, z, _ = synthetic_batch(2, 16, 112, 112, 100)
    trace = {}
    _, _, grads = vo.train_step_grads(sd, x, z, trace=trace)
    return sd, trace, grads


def _conv_case(spec, sd, trace, prefix="model."):
    """fprop / dgrad / wgrad of one convolution on the oracle's input and upstream gradient."""
    from zeroshotvideoclassification_b200 import _lib, ops
    w = sd[prefix + spec.name + ".weight"]
    x_in = trace[spec.name + ":in"]
    y_ref = trace[spec.name].detach()
    dy_ref = trace[spec.name].grad
    dx_ref, dw_ref = vo.conv3d_grads(x_in, w, dy_ref, spec.stride, spec.padding)
    N, _, T, H, W = x_in.shape
    first = spec.cin == 3
    layout = _lib.X_WFOLD if first else _lib.X_NDHWC
    op = ops.Conv3d(N, T, H, W, spec.cin, spec.cout, spec.kernel, spec.stride, spec.padding, layout)
    wf, wd = op.pack(w.cuda(), need_dgrad=not first)
    xd = ops.repack_input(x_in.cuda(), _lib.X_WFOLD, spec.padding[2]) if first else to_ndhwc(x_in)
    y, _, _ = op.fprop(xd, wf, stats=False)
    dyd = to_ndhwc(dy_ref)
    dw, _ = op.wgrad(xd, dyd)
    out = {"fprop": rel_err(from_ndhwc(y, spec.cout), y_ref), "wgrad": rel_err(dw.cpu(), dw_ref)}
    if not first:
        out["dgrad"] = rel_err(from_ndhwc(op.dgrad(dyd, wd), spec.cin), dx_ref)
    torch.cuda.synchronize()
    return out


def test_every_convolution_teacher_forced(r2plus1d_trace):
    from zeroshotvideoclassification_b200 import engine
    sd, trace, _ = r2plus1d_trace
    specs = engine.all_conv_specs("r2plus1d_18")
    assert len(specs) == 37
    worst, bad = {}, []
    for spec in specs:
        errs = _conv_case(spec, sd, trace)
        for k, v in errs.items():
            tol = _wgrad_tol(spec.cin, spec.cout, spec.kernel) if k == "wgrad" else TOL
            if not v <= tol:
                bad.append((spec.name, k, v, tol))
            worst[k] = max(worst.get(k, 0.0), v)
    print("worst teacher-forced conv errors", worst)
    assert not bad, bad


def test_every_batchnorm_teacher_forced(r2plus1d_trace):
    """BatchNorm3d forward (batch statistics, normalise, running statistics) and backward (dx, dgamma, dbeta) of all 37
    instances on the oracle's conv output and the oracle's gradient w.r.t. the BatchNorm output."""
    from zeroshotvideoclassification_b200 import engine, ops
    sd, trace, grads = r2plus1d_trace
    worst, bad = {}, []
    for spec in engine.all_conv_specs("r2plus1d_18"):
        C = spec.cout
        y_ref = trace[spec.name].detach()
        bn_out = trace[spec.bn].detach()
        g_bn = trace[spec.bn].grad                       # gradient w.r.t. the BatchNorm output
        dy_ref = trace[spec.name].grad                   # gradient w.r.t. its input
        gamma, beta = sd["model." + spec.bn + ".weight"].cuda(), sd["model." + spec.bn + ".bias"].cuda()
        N, _, T, H, W = y_ref.shape
        rows = N * T * H * W
        yd = to_ndhwc(y_ref)
        flat = yd.float().reshape(N * T, H * W, cpad(C))  # the conv epilogue's partial sums, one row per (n, t) slab
        ps, pq = flat.sum(1).contiguous(), (flat * flat).sum(1).contiguous()
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        scale, shift, mean, invstd = ops.bn_finalize(ps, pq, C, rows, gamma, beta, rm, rv)
        out = ops.bn_apply(yd, scale, shift, C, False)
        dy, _, _, dg, db, _, _ = ops.bn_bwd(to_ndhwc(g_bn), None, 0, yd, mean, invstd, gamma, C)
        torch.cuda.synchronize()
        errs = {
            "out": rel_err(from_ndhwc(out, C), bn_out),
            "mean": rel_err(mean[:C].cpu(), trace[spec.bn + ":mean"]),
            "var": rel_err((1.0 / invstd[:C].cpu() ** 2 - 1e-5), trace[spec.bn + ":var"]),
            "dx": rel_err(from_ndhwc(dy, C), dy_ref),
            "dgamma": rel_err(dg.cpu(), grads["model." + spec.bn + ".weight"]),
            "dbeta": rel_err(db.cpu(), grads["model." + spec.bn + ".bias"]),
        }
        # running statistics after this one training-mode forward (momentum 0.1, unbiased variance); same max-norm
        # measure (a channel whose mean is ~0 has no meaningful element-wise relative error)
        errs["running_mean"] = rel_err(rm.cpu(), 0.1 * trace[spec.bn + ":mean"])
        errs["running_var"] = rel_err(rv.cpu(), 0.9 + 0.1 * trace[spec.bn + ":var"] * rows / (rows - 1))
        for k, v in errs.items():
            if not v <= TOL:
                bad.append((spec.bn, k, v))
            worst[k] = max(worst.get(k, 0.0), v)
    print("worst teacher-forced BatchNorm errors", worst)
    assert not bad, bad


def test_c3d_convolutions_teacher_forced():
    """The 8 convolutions (3x3x3 + bias + ReLU) of network.C3D (network.py:102-117) on the oracle's inputs / gradients."""
    from zeroshotvideoclassification_b200 import _lib, ops, video_models as vm
    torch.manual_seed(0)
    model = vm.get_network(vm.default_opt("c3d"))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, z, _ = synthetic_batch(2, 16, 112, 112, 100)
    trace = {}
    vo.c3d_train_step_grads(sd, x, z, trace=trace)
    names = ["conv1", "conv2", "conv3a", "conv3b", "conv4a", "conv4b", "conv5a", "conv5b"]
    worst, bad = {}, []
    for i, name in enumerate(names):
        w, b = sd[name + ".weight"], sd[name + ".bias"]
        x_in = trace[name + ":in"]
        out_ref = trace[name + ":out"].detach()
        dz_ref = trace[name].grad                         # gradient w.r.t. the pre-ReLU output
        dx_ref, dw_ref = vo.conv3d_grads(x_in, w, dz_ref, (1, 1, 1), (1, 1, 1))
        N, cin, T, H, W = x_in.shape
        cout = w.shape[0]
        first = i == 0
        layout = _lib.X_WFOLD if first else _lib.X_NDHWC
        op = ops.Conv3d(N, T, H, W, cin, cout, (3, 3, 3), (1, 1, 1), (1, 1, 1), layout)
        wf, wd = op.pack(w.cuda(), need_dgrad=not first)
        xd = ops.repack_input(x_in.cuda(), _lib.X_WFOLD, 1) if first else to_ndhwc(x_in)
        y, _, _ = op.fprop(xd, wf, stats=False, bias=b.cuda(), relu=True)
        dzd = to_ndhwc(dz_ref)
        dw, db = op.wgrad(xd, dzd, want_bias=True)
        errs = {"fprop": rel_err(from_ndhwc(y, cout), out_ref), "wgrad": rel_err(dw.cpu(), dw_ref),
                "bias_grad": rel_err(db.cpu(), dz_ref.sum((0, 2, 3, 4)))}
        if not first:
            errs["dgrad"] = rel_err(from_ndhwc(op.dgrad(dzd, wd), cin), dx_ref)
        torch.cuda.synchronize()
        for k, v in errs.items():
            tol = _wgrad_tol(cin, cout, (3, 3, 3)) if k == "wgrad" else TOL
            if not v <= tol:
                bad.append((name, k, v, tol))
            worst[k] = max(worst.get(k, 0.0), v)
    print("worst teacher-forced C3D errors", worst)
    assert not bad, bad


@pytest.mark.parametrize("cin,cout,kernel,padding,dims", [
    (64, 144, (1, 3, 3), (0, 1, 1), (16, 56, 56)),      # layer1 spatial
    (144, 64, (3, 1, 1), (1, 0, 0), (16, 56, 56)),      # layer1 temporal
    (512, 1152, (1, 3, 3), (0, 1, 1), (2, 7, 7)),       # layer4.1 spatial
    (1152, 512, (3, 1, 1), (1, 0, 0), (2, 7, 7)),       # layer4.1 temporal
])
def test_bs22_full_shape_dgrad_wgrad(cin, cout, kernel, padding, dims):
    """BASELINE.json config-2 batch (22 clips): dgrad and wgrad of the first and the last residual stage against the CPU
    oracle (F.conv3d autograd, fp32) on the same bf16-rounded operands."""
    from zeroshotvideoclassification_b200 import ops
    g = torch.Generator().manual_seed(cin + cout)
    T, H, W = dims
    N = 22
    x = torch.randn(N, cin, T, H, W, generator=g).to(torch.bfloat16).float()
    k = kernel[0] * kernel[1] * kernel[2]
    w = (torch.randn(cout, cin, *kernel, generator=g) * (2.0 / (cout * k)) ** 0.5).to(torch.bfloat16).float()
    dy = torch.randn(N, cout, T, H, W, generator=g).to(torch.bfloat16).float()
    dx_ref, dw_ref = vo.conv3d_grads(x, w, dy, (1, 1, 1), padding)
    op = ops.Conv3d(N, T, H, W, cin, cout, kernel, (1, 1, 1), padding)
    wf, wd = op.pack(w.cuda())
    xd, dyd = to_ndhwc(x), to_ndhwc(dy)
    dx = op.dgrad(dyd, wd)
    dw, _ = op.wgrad(xd, dyd)
    torch.cuda.synchronize()
    assert rel_err(from_ndhwc(dx, cin), dx_ref) <= TOL
    assert rel_err(dw.cpu(), dw_ref) <= TOL_WGRAD_BF16_OPERANDS
