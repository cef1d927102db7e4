"""GPU parity of the HBM-bound kernels: layout conversion, BatchNorm finalize/apply/backward, head, loss."""
import pytest
import torch
import torch.nn.functional as F

from tests.helpers import bf16_round, cpad, from_ndhwc, rel_err, to_ndhwc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C", [3, 45, 64, 230])
def test_layout_roundtrip(C):
    from zeroshotvideoclassification_b200 import ops
    x = bf16_round(torch.randn(2, C, 3, 5, 7))
    a = ops.ncdhw_to_ndhwc(x.cuda())
    assert a.shape == (2, 3, 5, 7, cpad(C))
    assert torch.equal(from_ndhwc(a, C), x)
    if cpad(C) != C:
        assert float(a[..., C:].float().abs().max()) == 0.0
    back = ops.ndhwc_to_ncdhw(a, C)
    assert torch.equal(back.cpu(), x)


def test_wfold_repack_layout():
    from zeroshotvideoclassification_b200 import _lib, ops
    x = bf16_round(torch.randn(2, 3, 2, 6, 10))
    a = ops.repack_input(x.cuda(), _lib.X_WFOLD, 3).cpu().float()
    assert a.shape == (2, 2, 6, 18, 8)
    assert torch.equal(a[:, :, :, 3:13, :3].permute(0, 4, 1, 2, 3), x)
    assert float(a[:, :, :, :3].abs().max()) == 0.0 and float(a[:, :, :, 13:].abs().max()) == 0.0
    assert float(a[..., 3:].abs().max()) == 0.0


def _bn_reference(y, gamma, beta, eps=1e-5):
    mean = y.mean((0, 2, 3, 4))
    var = y.var((0, 2, 3, 4), unbiased=False)
    sh = (1, -1, 1, 1, 1)
    return (y - mean.view(sh)) / torch.sqrt(var.view(sh) + eps) * gamma.view(sh) + beta.view(sh), mean, var


@pytest.mark.parametrize("C,relu,mode", [(45, True, "plain"), (64, True, "residual"), (128, True, "two"),
                                         (230, False, "plain"), (1152, True, "plain")])
def test_bn_forward_backward(C, relu, mode):
    """conv-epilogue statistics are emulated with one partial row per (n,t) slab; compares finalize, apply and the
    two-pass backward with autograd on the fp32 reference formula (resnet.py:48,95-98,102-113)."""
    from zeroshotvideoclassification_b200 import ops
    g = torch.Generator().manual_seed(C)
    N, T, H, W = 2, 3, 6, 5
    y = bf16_round(torch.randn(N, C, T, H, W, generator=g) * 1.5 + 0.3)
    gamma = torch.rand(C, generator=g) + 0.5
    beta = torch.randn(C, generator=g) * 0.2
    y2 = bf16_round(torch.randn(N, C, T, H, W, generator=g)) if mode == "two" else None
    gamma2 = torch.rand(C, generator=g) + 0.5
    beta2 = torch.randn(C, generator=g) * 0.2
    res = bf16_round(torch.randn(N, C, T, H, W, generator=g)) if mode == "residual" else None
    gout = bf16_round(torch.randn(N, C, T, H, W, generator=g))

    # reference (fp32 autograd)
    yr = y.clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    o, mean_ref, var_ref = _bn_reference(yr, gr, br)
    extra = []
    if y2 is not None:
        y2r = y2.clone().requires_grad_(True)
        g2r = gamma2.clone().requires_grad_(True)
        b2r = beta2.clone().requires_grad_(True)
        o = o + _bn_reference(y2r, g2r, b2r)[0]
        extra = [y2r, g2r, b2r]
    if res is not None:
        rr = res.clone().requires_grad_(True)
        o = o + rr
        extra = [rr]
    if relu:
        o = F.relu(o)
    o.backward(gout)

    dev = "cuda"
    yd = to_ndhwc(y)
    rows = N * T * H * W
    cp = cpad(C)

    def partials(t):
        flat = t.float().reshape(N * T, H * W, cp)
        return flat.sum(1).contiguous(), (flat * flat).sum(1).contiguous()

    ps, pq = partials(yd)
    rm = torch.zeros(C, device=dev)
    rv = torch.ones(C, device=dev)
    scale, shift, mean, invstd = ops.bn_finalize(ps, pq, C, rows, gamma.to(dev), beta.to(dev), rm, rv)
    assert torch.allclose(mean[:C].cpu(), mean_ref.detach(), atol=2e-5, rtol=1e-5)
    assert torch.allclose(invstd[:C].cpu(), 1 / torch.sqrt(var_ref.detach() + 1e-5), rtol=1e-4)
    assert torch.allclose(rm.cpu(), 0.1 * mean_ref.detach(), atol=1e-5, rtol=1e-4)
    assert torch.allclose(rv.cpu(), 0.9 + 0.1 * var_ref.detach() * rows / (rows - 1), rtol=1e-4)
    kw = {}
    if y2 is not None:
        y2d = to_ndhwc(y2)
        ps2, pq2 = partials(y2d)
        scale2, shift2, mean2, invstd2 = ops.bn_finalize(ps2, pq2, C, rows, gamma2.to(dev), beta2.to(dev), None, None)
        kw = dict(y2=y2d, scale2=scale2, shift2=shift2)
    if res is not None:
        kw = dict(residual=to_ndhwc(res))
    out = ops.bn_apply(yd, scale, shift, C, relu, **kw)
    torch.cuda.synchronize()
    assert rel_err(from_ndhwc(out, C), o.detach()) < 1e-2

    gd = to_ndhwc(gout)
    if y2 is not None:
        dy, dy2, dz, dg, db, dg2, db2 = ops.bn_bwd(gd, out, relu, yd, mean, invstd, gamma.to(dev), C, y2=y2d,
                                                   mean2=mean2, invstd2=invstd2, gamma2=gamma2.to(dev))
    else:
        dy, dy2, dz, dg, db, dg2, db2 = ops.bn_bwd(gd, out, relu, yd, mean, invstd, gamma.to(dev), C,
                                                   want_dz=res is not None)
    torch.cuda.synchronize()
    # the ReLU mask is taken from the bf16 output; elements whose fp32 pre-activation is within bf16 rounding of
    # zero may flip, so compare with a tolerance relative to the tensor scale
    assert rel_err(from_ndhwc(dy, C), yr.grad) < 2e-2
    assert rel_err(dg.cpu(), gr.grad) < 5e-3
    assert rel_err(db.cpu(), br.grad) < 5e-3
    if y2 is not None:
        assert rel_err(from_ndhwc(dy2, C), extra[0].grad) < 2e-2
        assert rel_err(dg2.cpu(), extra[1].grad) < 5e-3
        assert rel_err(db2.cpu(), extra[2].grad) < 5e-3
    if res is not None:
        assert rel_err(from_ndhwc(dz, C), extra[0].grad) < 1e-2
    if mode == "plain" and relu:
        # relu=2: the mask is recomputed from y with the forward's scale/shift instead of re-reading `out`
        r2 = ops.bn_bwd(gd, None, 2, yd, mean, invstd, gamma.to(dev), C, mask_scale=scale, mask_shift=shift)
        torch.cuda.synchronize()
        assert torch.equal(r2[0], dy) and torch.equal(r2[3], dg) and torch.equal(r2[4], db)


def test_bn_eval_scale_shift():
    from zeroshotvideoclassification_b200 import ops
    C = 45
    g = torch.Generator().manual_seed(0)
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    rm, rv = torch.randn(C, generator=g), torch.rand(C, generator=g) + 0.1
    sc, sh = ops.bn_eval_scale_shift(C, gamma.cuda(), beta.cuda(), rm.cuda(), rv.cuda())
    inv = 1 / torch.sqrt(rv + 1e-5)
    assert torch.allclose(sc[:C].cpu(), gamma * inv, rtol=1e-5)
    assert torch.allclose(sh[:C].cpu(), beta - rm * gamma * inv, rtol=1e-4, atol=1e-6)
    assert float(sc[C:].abs().max()) == 0.0


@pytest.mark.parametrize("B", [1, 5, 22])
def test_head_forward_backward(B):
    """pool -> MLP -> normalize (network.py:595-596) and its backward vs fp32 autograd."""
    from zeroshotvideoclassification_b200 import engine
    g = torch.Generator().manual_seed(B)
    C, Hd, E = 512, 512, 300
    feat = bf16_round(torch.randn(B, C, 2, 7, 7, generator=g).abs())
    w1 = (torch.rand(Hd, C, generator=g) * 2 - 1) / C ** 0.5
    b1 = (torch.rand(Hd, generator=g) * 2 - 1) / C ** 0.5
    w2 = (torch.rand(E, Hd, generator=g) * 2 - 1) / Hd ** 0.5
    b2 = (torch.rand(E, generator=g) * 2 - 1) / Hd ** 0.5
    demb = torch.randn(B, E, generator=g)

    fr = feat.clone().requires_grad_(True)
    ps = [t.clone().requires_grad_(True) for t in (w1, b1, w2, b2)]
    o = F.normalize(F.linear(F.relu(F.linear(fr.mean((2, 3, 4)), ps[0], ps[1])), ps[2], ps[3]))
    o.backward(demb)

    fd = to_ndhwc(feat).requires_grad_(True)
    pd = [t.clone().cuda().requires_grad_(True) for t in (w1, b1, w2, b2)]
    emb = engine.head_forward(fd, *pd)
    emb.backward(demb.cuda())
    torch.cuda.synchronize()
    assert rel_err(emb.detach().cpu(), o.detach()) < 1e-4
    for got, ref in zip(pd, ps):
        assert rel_err(got.grad.cpu(), ref.grad) < 1e-3
    assert rel_err(from_ndhwc(fd.grad, C), fr.grad) < 1e-2


def test_mse_loss_and_grad():
    from zeroshotvideoclassification_b200 import ops
    g = torch.Generator().manual_seed(0)
    emb = F.normalize(torch.randn(22, 300, generator=g))
    tgt = F.normalize(torch.randn(22, 300, generator=g))
    loss, demb = ops.mse_fwd_bwd(emb.cuda(), tgt.cuda(), grad_scale=65536.0)
    assert abs(float(loss) - float(((emb - tgt) ** 2).mean())) < 1e-7
    assert torch.allclose(demb.cpu(), 65536.0 * 2 * (emb - tgt) / emb.numel(), rtol=1e-5, atol=1e-8)


def test_maxpool_forward_backward():
    from zeroshotvideoclassification_b200 import ops
    g = torch.Generator().manual_seed(0)
    for (shape, k, p) in (((2, 64, 4, 8, 8), (1, 2, 2), (0, 0, 0)), ((1, 128, 4, 8, 8), (2, 2, 2), (0, 0, 0)),
                          ((2, 512, 2, 7, 7), (2, 2, 2), (0, 1, 1))):
        x = bf16_round(torch.randn(*shape, generator=g))
        xr = x.clone().requires_grad_(True)
        ref = F.max_pool3d(xr, k, k, padding=p)
        dy = bf16_round(torch.randn(ref.shape, generator=g))
        ref.backward(dy)
        C = shape[1]
        y, am = ops.maxpool3d_fwd(to_ndhwc(x), C, k, p)
        assert torch.equal(from_ndhwc(y, C), ref.detach())
        dx = ops.maxpool3d_bwd(to_ndhwc(dy), am, tuple(to_ndhwc(x).shape), C, k, p)
        assert torch.equal(from_ndhwc(dx, C), xr.grad)
        assert am.dtype == torch.uint8


def test_pool_and_relu_backward_with_fused_bias_gradient():
    """C3D backward of conv+bias -> ReLU [-> MaxPool3d] (network.py:147-163): the pooling backward takes the ReLU mask
    from the POOLED tensor and emits the bias gradient (column sums of dz) in the same pass; likewise zsv_relu_bwd."""
    from zeroshotvideoclassification_b200 import ops
    g = torch.Generator().manual_seed(5)
    for (shape, k, p) in (((2, 64, 4, 8, 8), (1, 2, 2), (0, 0, 0)), ((3, 128, 4, 8, 6), (2, 2, 2), (0, 0, 0)),
                          ((2, 512, 2, 7, 7), (2, 2, 2), (0, 1, 1)), ((1, 24, 2, 4, 4), (2, 2, 2), (0, 0, 0))):
        pre = bf16_round(torch.randn(*shape, generator=g))
        pre[:, :, :, :2, :2] = -pre[:, :, :, :2, :2].abs()          # some windows are all-negative: pooled value 0
        xr = pre.clone().requires_grad_(True)
        act = F.relu(xr)
        ref = F.max_pool3d(act, k, k, padding=p)
        dy = bf16_round(torch.randn(ref.shape, generator=g))
        ref.backward(dy)
        C = shape[1]
        a = to_ndhwc(act.detach())
        y, am = ops.maxpool3d_fwd(a, C, k, p)
        dz, db = ops.maxpool3d_bwd(to_ndhwc(dy), am, tuple(a.shape), C, k, p, relu_pooled=y, want_bias=True)
        torch.cuda.synchronize()
        got = from_ndhwc(dz, C)
        # where a window is all zeros after the ReLU, torch sends dy to its first element and the ReLU backward then
        # zeroes it: same result, dz = 0
        assert torch.equal(got, xr.grad), (shape, float((got - xr.grad).abs().max()))
        assert rel_err(db.cpu(), xr.grad.sum((0, 2, 3, 4))) < 1e-5
        # plain ReLU backward with the bias gradient
        xr2 = pre.clone().requires_grad_(True)
        out = F.relu(xr2)
        gy = bf16_round(torch.randn(shape, generator=g))
        out.backward(gy)
        dz2, db2 = ops.relu_bwd(to_ndhwc(gy), to_ndhwc(out.detach()), C, want_bias=True)
        assert torch.equal(from_ndhwc(dz2, C), xr2.grad)
        assert rel_err(db2.cpu(), xr2.grad.sum((0, 2, 3, 4))) < 1e-5
        assert torch.equal(from_ndhwc(ops.relu_bwd(to_ndhwc(gy), to_ndhwc(out.detach()), C), C), xr2.grad)


@pytest.mark.parametrize("B,K,J", [(22, 8192, 4096), (22, 4096, 300), (22, 512, 512), (5, 512, 300), (1, 37, 5),
                                   (30, 1030, 77), (50, 256, 64)])
def test_linear_forward_backward(B, K, J):
    """zsv_linear_fwd / zsv_linear_bwd (C3D fc6 / regressor, network.py:120,132,166,178; the MLP head's layers) vs
    torch on the CPU: full C3D sizes, odd sizes (scalar path), more rows than one 24-row pass."""
    from zeroshotvideoclassification_b200 import ops
    g = torch.Generator().manual_seed(B * 7 + J)
    x = torch.randn(B, K, generator=g)
    w = torch.randn(J, K, generator=g) / K ** 0.5
    b = torch.randn(J, generator=g)
    dy = torch.randn(B, J, generator=g)
    for relu in (False, True):
        xr, wr, br = (t.clone().requires_grad_(True) for t in (x, w, b))
        ref = F.linear(xr, wr, br)
        if relu:
            ref = F.relu(ref)
        ref.backward(dy)
        out = ops.linear_fwd(x.cuda(), w.cuda(), b.cuda(), relu)
        dx, dw, db = ops.linear_bwd(dy.cuda(), x.cuda(), w.cuda(), out if relu else None)
        torch.cuda.synchronize()
        assert rel_err(out.cpu(), ref.detach()) < 2e-5
        assert rel_err(dx.cpu(), xr.grad) < 2e-5
        assert rel_err(dw.cpu(), wr.grad) < 2e-5
        assert rel_err(db.cpu(), br.grad) < 2e-5


def test_clip_transform_matches_oracle_and_feeds_the_model():
    """zsv_clip_transform (uint8 frames -> normalise -> Resize(128) -> crop 112 -> flip -> bf16 W-folded) vs the transform
    oracle (pinned to the reference's transforms), per clip crop origins / flips, up- and down-scaling; and the model
    gives the same embeddings from the GPU-transformed clips as from the fp32 batch the reference loader would build."""
    import numpy as np
    import torch
    from oracle import transform_oracle as to
    from zeroshotvideoclassification_b200 import ops
    from zeroshotvideoclassification_b200 import video_models as vm
    g = torch.Generator().manual_seed(8)
    for (T, H, W) in ((4, 68, 90), (3, 200, 150), (2, 128, 171)):
        N = 3
        frames = torch.randint(0, 256, (N, T, H, W, 3), generator=g, dtype=torch.uint8)
        scale = 128.0 / min(H, W)
        hr, wr = int(np.floor(H * scale)), int(np.floor(W * scale))
        ij = torch.tensor([[0, 0], [hr - 112, wr - 112], list(to.center_crop_origin(hr, wr))], dtype=torch.int32)
        flip = torch.tensor([0, 1, 1], dtype=torch.uint8)
        out = ops.clip_transform(frames.cuda(), ij, flip)
        torch.cuda.synchronize()
        assert out.shape == (N, T, 112, 120, 8)
        got = out[:, :, :, 3:115, :3].float().cpu().permute(0, 4, 1, 2, 3)             # [N,3,T,112,112]
        for n in range(N):
            ref = to.clip_transform(frames[n], ij[n].tolist(), bool(flip[n]))
            assert float((got[n] - ref).abs().max()) < 2.5e-3                           # bf16 rounding of values in [-0.5, 0]
        assert float(out[:, :, :, :3].abs().max()) == 0.0 and float(out[:, :, :, 115:].abs().max()) == 0.0
        assert float(out[..., 3:].abs().max()) == 0.0
    # end to end: same embeddings as from the fp32 clips (bf16 rounding of the input is where both paths start)
    torch.manual_seed(0)
    model = vm.get_network(vm.default_opt("r2plus1d_18")).cuda().eval()
    frames = torch.randint(0, 256, (2, 8, 136, 180, 3), generator=g, dtype=torch.uint8)
    ij = torch.tensor([[3, 17], [8, 40]], dtype=torch.int32)
    flip = torch.tensor([1, 0], dtype=torch.uint8)
    x32 = torch.stack([to.clip_transform(frames[n], ij[n].tolist(), bool(flip[n])) for n in range(2)])[:, None]
    with torch.no_grad():
        emb_ref, _ = model(x32.cuda())
        emb_gpu, _ = model(ops.clip_transform(frames.cuda(), ij, flip))
    assert float((emb_ref - emb_gpu).abs().max() / emb_ref.abs().max()) < 2e-2


def test_fused_adam_matches_torch_adam():
    """zsv_adam_step == torch.optim.Adam (main.py:131) over several steps, ragged tensor sizes, with and without weight
    decay; state dicts interchange."""
    import torch
    from zeroshotvideoclassification_b200.optim import FusedAdam
    for wd in (0.0, 1e-2):
        g = torch.Generator().manual_seed(3)
        shapes = [(64, 3, 1, 7, 7), (45,), (1,), (300, 512), (7, 13, 3)] + [(17,)] * 80      # > 72 tensors: two launches
        ref = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
        mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
        o_ref = torch.optim.Adam(ref, lr=1e-3, weight_decay=wd)
        o_mine = FusedAdam(mine, lr=1e-3, weight_decay=wd)
        for _ in range(5):
            for a, b in zip(ref, mine):
                gr = torch.randn(a.shape, generator=g).cuda()
                a.grad, b.grad = gr.clone(), gr.clone()
            o_ref.step()
            o_mine.step()
        for a, b in zip(ref, mine):
            assert torch.allclose(a, b, rtol=2e-6, atol=2e-7)
        sd = o_mine.state_dict()
        assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["state"][0]["step"]) == 5.0
        o_ref.load_state_dict(sd)          # same state layout as torch.optim.Adam


def test_fused_adam_lr_schedule_and_late_parameters():
    """The learning rate is read from a device scalar that follows group['lr'] (main.py:133,374: MultiStepLR), and every
    tensor has its own step counter: a parameter that starts receiving gradients later gets its own bias correction."""
    from zeroshotvideoclassification_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(4)
    ref = [torch.nn.Parameter(torch.randn(33, 7, generator=g).cuda()), torch.nn.Parameter(torch.randn(129, generator=g).cuda())]
    mine = [torch.nn.Parameter(p.detach().clone()) for p in ref]
    o_ref, o_mine = torch.optim.Adam(ref, lr=1e-2), FusedAdam(mine, lr=1e-2)
    s_ref = torch.optim.lr_scheduler.MultiStepLR(o_ref, [2, 4], gamma=0.1)
    s_mine = torch.optim.lr_scheduler.MultiStepLR(o_mine, [2, 4], gamma=0.1)
    for it in range(6):
        for i, (a, b) in enumerate(zip(ref, mine)):
            if i == 1 and it < 2:
                a.grad = b.grad = None           # the second tensor joins at step 2
                continue
            gr = torch.randn(a.shape, generator=g).cuda()
            a.grad, b.grad = gr.clone(), gr.clone()
        o_ref.step(), o_mine.step()
        s_ref.step(), s_mine.step()
    for a, b in zip(ref, mine):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7)
    assert float(o_mine.state[mine[0]]["step"]) == 6.0 and float(o_mine.state[mine[1]]["step"]) == 4.0


@pytest.mark.parametrize("network", ["r2plus1d_18", "c3d"])
def test_fused_adam_writes_the_packed_weight_images(network):
    """FusedAdam(model=...) == torch.optim.Adam on every parameter, and the bf16 weight images it leaves behind
    (zsv_adam_pack_step, engine.PackedWeights) are bit-identical to re-packing the updated fp32 weights; the next
    forward uses them (no re-pack launch) and a weight changed behind the optimizer's back is noticed."""
    from zeroshotvideoclassification_b200 import _lib, engine, video_models as vm
    from zeroshotvideoclassification_b200.optim import FusedAdam
    torch.manual_seed(0)
    m_ref = vm.get_network(vm.default_opt(network)).cuda().train()
    torch.manual_seed(0)
    m_mine = vm.get_network(vm.default_opt(network)).cuda().train()
    if network == "c3d":
        m_ref.dropout.p = m_mine.dropout.p = 0.0
    o_ref = torch.optim.Adam(m_ref.parameters(), lr=1e-3)
    o_mine = FusedAdam(m_mine.parameters(), lr=1e-3, model=m_mine)
    g = torch.Generator().manual_seed(2)
    shape = (2, 1, 3, 16, 112, 112) if network == "c3d" else (2, 1, 3, 8, 32, 32)
    for it in range(3):
        x = torch.randn(*shape, generator=g).cuda()
        z = F.normalize(torch.randn(2, 300, generator=g)).cuda()
        o_mine.zero_grad(set_to_none=True)
        out = m_mine(x)
        F.mse_loss(out[0] if isinstance(out, tuple) else out, z).backward()
        # both optimizers get the SAME gradients (two bf16 networks one ulp apart drift chaotically)
        for a, b in zip(m_ref.parameters(), m_mine.parameters()):
            a.grad = None if b.grad is None else b.grad.detach().clone()
        o_ref.step()
        o_mine.step()
    for (k, a), (_, b) in zip(m_ref.named_parameters(), m_mine.named_parameters()):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), k
    pw = o_mine._packed
    assert pw is not None and pw.fresh()
    wfs, wds = pw.plan.pack([w.detach() for w in pw.weights])
    for i, (wf, wd) in enumerate(zip(wfs, wds)):
        assert torch.equal(wf, pw.wfs[i]), i
        assert (wd is None) == (pw.wds[i] is None) and (wd is None or torch.equal(wd, pw.wds[i])), i
    # the next forward takes the published images: no pack kernel among its launches
    n0 = _lib.launch_count()
    with torch.no_grad():
        m_mine(x)
    n_pub = _lib.launch_count() - n0
    with torch.no_grad():
        pw.weights[1].mul_(1.0)                               # in-place op: version bump -> images are stale
    assert not pw.fresh() and engine.published_for(pw.weights[0]) is None
    n0 = _lib.launch_count()
    with torch.no_grad():
        m_mine(x)
    assert _lib.launch_count() - n0 > n_pub                   # this forward re-packed
    engine.ensure_packed_fresh()
    assert pw.fresh()
