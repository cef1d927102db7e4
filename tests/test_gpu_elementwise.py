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
