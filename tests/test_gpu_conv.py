"""GPU parity of the implicit-GEMM convolution kernels (fprop / dgrad / wgrad) against the CPU oracle.

Tolerances: operands are bf16-rounded identically on both sides, accumulation is fp32, outputs are rounded to
bf16 (fprop/dgrad) -> 1e-2 relative to the tensor scale (north_star), wgrad fp32 -> 2e-3.
"""
import pytest
import torch

from oracle import video_oracle as vo
from tests.helpers import bf16_round, cpad, from_ndhwc, rel_err, to_ndhwc

pytestmark = pytest.mark.gpu

TOL = 1e-2

# (name, N, T, H, W, Cin, Cout, kernel, stride, padding)
CASES = [
    ("spatial_64_64", 2, 2, 16, 16, 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("spatial_64_144", 2, 4, 24, 24, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("temporal_144_64", 2, 6, 12, 12, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("temporal_45_64", 2, 4, 10, 10, 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("spatial_s2_64_230", 2, 2, 20, 20, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ("temporal_s2_230_128", 1, 8, 6, 6, 230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0)),
    ("downsample_64_128", 2, 4, 12, 12, 64, 128, (1, 1, 1), (2, 2, 2), (0, 0, 0)),
    ("spatial_256_460", 1, 2, 14, 14, 256, 460, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("spatial_odd_hw", 3, 3, 7, 7, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("spatial_s2_odd", 1, 2, 15, 13, 64, 96, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    ("c3d_27tap", 1, 4, 8, 8, 64, 128, (3, 3, 3), (1, 1, 1), (1, 1, 1)),
]


def _make(case, seed=0):
    name, N, T, H, W, cin, cout, k, s, p = case
    g = torch.Generator().manual_seed(seed)
    x = bf16_round(torch.randn(N, cin, T, H, W, generator=g))
    w = bf16_round(torch.randn(cout, cin, *k, generator=g) * (2.0 / (cin * k[0] * k[1] * k[2])) ** 0.5)
    return x, w


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fprop_matches_oracle(case):
    from zeroshotvideoclassification_b200 import ops
    name, N, T, H, W, cin, cout, k, s, p = case
    x, w = _make(case)
    ref = vo.conv3d(x, w, None, s, p)
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    wf, _ = op.pack(w.cuda(), need_dgrad=False)
    y, ps, pq = op.fprop(to_ndhwc(x), wf, stats=True)
    torch.cuda.synchronize()
    got = from_ndhwc(y, cout)
    assert got.shape == ref.shape
    assert rel_err(got, ref) < TOL, name
    # pad lanes are written as exact zeros
    if cpad(cout) != cout:
        assert float(y[..., cout:].float().abs().max()) == 0.0
    # BatchNorm partial statistics: sums of the bf16-rounded outputs
    yb = y[..., :cout].float().reshape(-1, cout).double()
    s1 = ps.double().sum(0)[:cout].cpu()
    s2 = pq.double().sum(0)[:cout].cpu()
    assert torch.allclose(s1, yb.sum(0).cpu(), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s2, (yb * yb).sum(0).cpu(), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_dgrad_matches_oracle(case):
    from zeroshotvideoclassification_b200 import ops
    name, N, T, H, W, cin, cout, k, s, p = case
    x, w = _make(case)
    y = vo.conv3d(x, w, None, s, p)
    g = torch.Generator().manual_seed(1)
    dy = bf16_round(torch.randn(y.shape, generator=g))
    dx_ref, _ = vo.conv3d_grads(x, w, dy, s, p)
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    _, wd = op.pack(w.cuda(), need_dgrad=True)
    dx = op.dgrad(to_ndhwc(dy), wd)
    torch.cuda.synchronize()
    assert rel_err(from_ndhwc(dx, cin), dx_ref) < TOL, name
    # fused residual-gradient add
    add = bf16_round(torch.randn(x.shape, generator=g))
    dx2 = op.dgrad(to_ndhwc(dy), wd, addend=to_ndhwc(add))
    torch.cuda.synchronize()
    assert rel_err(from_ndhwc(dx2, cin), dx_ref + add) < TOL, name


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_wgrad_matches_oracle(case):
    from zeroshotvideoclassification_b200 import ops
    name, N, T, H, W, cin, cout, k, s, p = case
    x, w = _make(case)
    y = vo.conv3d(x, w, None, s, p)
    g = torch.Generator().manual_seed(2)
    dy = bf16_round(torch.randn(y.shape, generator=g))
    _, dw_ref = vo.conv3d_grads(x, w, dy, s, p)
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    dw, _ = op.wgrad(to_ndhwc(x), to_ndhwc(dy))
    torch.cuda.synchronize()
    assert dw.shape == dw_ref.shape
    assert rel_err(dw.cpu(), dw_ref) < 2e-3, name


def test_stem_wfold_fprop_and_wgrad():
    """First layer (resnet.py:181): Cin=3, 1x7x7, stride (1,2,2) through the W-folded input layout."""
    from zeroshotvideoclassification_b200 import _lib, ops
    N, T, H, W, cin, cout = 2, 3, 32, 32, 3, 45
    k, s, p = (1, 7, 7), (1, 2, 2), (0, 3, 3)
    g = torch.Generator().manual_seed(3)
    x = bf16_round(torch.randn(N, cin, T, H, W, generator=g))
    w = bf16_round(torch.randn(cout, cin, *k, generator=g) * 0.1)
    ref = vo.conv3d(x, w, None, s, p)
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p, x_layout=_lib.X_WFOLD)
    xin = ops.repack_input(x.cuda(), _lib.X_WFOLD, p[2])
    wf, wd = op.pack(w.cuda(), need_dgrad=False)
    assert wd is None
    y, _, _ = op.fprop(xin, wf, stats=True)
    torch.cuda.synchronize()
    assert rel_err(from_ndhwc(y, cout), ref) < TOL
    dy = bf16_round(torch.randn(ref.shape, generator=g))
    _, dw_ref = vo.conv3d_grads(x, w, dy, s, p)
    dw, _ = op.wgrad(xin, to_ndhwc(dy))
    torch.cuda.synchronize()
    assert rel_err(dw.cpu(), dw_ref) < 2e-3


def test_bias_relu_epilogue():
    """C3D convolutions carry a bias and a ReLU (network.py:102-117,147-162)."""
    from zeroshotvideoclassification_b200 import ops
    N, T, H, W, cin, cout = 1, 4, 8, 8, 64, 64
    k, s, p = (3, 3, 3), (1, 1, 1), (1, 1, 1)
    g = torch.Generator().manual_seed(4)
    x = bf16_round(torch.randn(N, cin, T, H, W, generator=g))
    w = bf16_round(torch.randn(cout, cin, *k, generator=g) * 0.03)
    b = torch.randn(cout, generator=g)
    ref = torch.relu(vo.conv3d(x, w, b, s, p))
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    wf, _ = op.pack(w.cuda(), need_dgrad=False)
    y, _, _ = op.fprop(to_ndhwc(x), wf, stats=False, bias=b.cuda(), relu=True)
    torch.cuda.synchronize()
    assert rel_err(from_ndhwc(y, cout), ref) < TOL


def test_fprop_full_size_layer1_linearity():
    """BASELINE-size layer1 spatial conv (bs=22, 64->144, 16x56x56): too large for the CPU oracle in seconds,
    so check against torch's fp32 conv on the same device and the size-independent linearity property."""
    from zeroshotvideoclassification_b200 import ops
    N, T, H, W, cin, cout = 22, 16, 56, 56, 64, 144
    k, s, p = (1, 3, 3), (1, 1, 1), (0, 1, 1)
    g = torch.Generator(device="cuda").manual_seed(5)
    xa = torch.randn(N, T, H, W, cin, generator=g, device="cuda").to(torch.bfloat16)
    w = bf16_round(torch.randn(cout, cin, *k, generator=g, device="cuda") * 0.05)
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    wf, _ = op.pack(w, need_dgrad=False)
    ya, _, _ = op.fprop(xa, wf, stats=False)
    ref = torch.nn.functional.conv3d(xa.float().permute(0, 4, 1, 2, 3), w, None, s, p).permute(0, 2, 3, 4, 1)
    assert rel_err(ya.float().cpu(), ref.cpu()) < TOL
    # linearity: conv(2x) == 2 conv(x) exactly in bf16 (power-of-two scaling commutes with rounding)
    yb, _, _ = op.fprop((xa.float() * 2).to(torch.bfloat16), wf, stats=False)
    assert torch.equal(yb.float(), ya.float() * 2)


def test_batched_weight_pack_matches_per_conv_pack():
    """zsv_conv3d_pack_weights (one tiled launch for a list of convolutions) writes the same bf16 images as
    zsv_conv3d_pack_weight per convolution, including the W-folded first-layer image and a conv without dgrad image."""
    import torch
    from zeroshotvideoclassification_b200 import _lib, ops
    g = torch.Generator().manual_seed(3)
    geoms = [(3, 45, (1, 7, 7), (1, 2, 2), (0, 3, 3), _lib.X_WFOLD, False),
             (45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), _lib.X_NDHWC, True),
             (64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), _lib.X_NDHWC, True),
             (64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1), _lib.X_NDHWC, False),
             (64, 128, (1, 1, 1), (2, 2, 2), (0, 0, 0), _lib.X_NDHWC, True),
             (128, 230, (1, 3, 3), (1, 1, 1), (0, 1, 1), _lib.X_NDHWC, True),      # Cout pad lanes in the dgrad image
             (230, 128, (3, 1, 1), (1, 1, 1), (1, 0, 0), _lib.X_NDHWC, True),      # Cin pad lanes, several ci chunks
             (64, 72, (3, 3, 3), (1, 1, 1), (1, 1, 1), _lib.X_NDHWC, True),        # 27 taps (C3D): 32-channel chunks
             (9, 19, (1, 3, 3), (1, 1, 1), (0, 1, 1), _lib.X_NDHWC, True)]         # everything ragged
    convs, ws, nd = [], [], []
    for cin, cout, k, s, p, layout, need in geoms:
        convs.append(ops.Conv3d(2, 8, 32, 32, cin, cout, k, s, p, layout))
        ws.append(torch.randn(cout, cin, *k, generator=g).cuda())
        nd.append(need)
    plan = ops.PackPlan(convs, nd)
    n0 = _lib.launch_count()
    wfs, wds = plan.pack(ws)
    assert _lib.launch_count() - n0 == 2       # the tiled kernel + the element-wise one for the W-folded image
    for c, w, need, wf, wd in zip(convs, ws, nd, wfs, wds):
        wf1, wd1 = c.pack(w, need_dgrad=need)
        assert torch.equal(wf.view(torch.int16), wf1.view(torch.int16))
        if need:
            assert torch.equal(wd.view(torch.int16), wd1.view(torch.int16))
        else:
            assert wd is None


@pytest.mark.parametrize("geom", [
    # cin, cout, kernel, stride, padding, T, H, W, relu, addend
    (48, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 8, 16, 16, True, False),     # temporal halo kernel
    (64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), 4, 24, 24, True, True),     # spatial halo kernel + shortcut gradient
    (230, 128, (3, 1, 1), (2, 1, 1), (1, 0, 0), 8, 12, 12, True, False),   # temporal stride 2: parity-class launches
    (144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), 4, 12, 12, False, False),   # BatchNorm without ReLU
    (300, 520, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, 10, 10, True, False),   # wide layers: several N tiles
    (256, 96, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, 14, 14, True, True),     # 256 gradient channels: N tiles capped at 192
    (512, 80, (3, 1, 1), (1, 1, 1), (1, 0, 0), 2, 7, 7, True, False),      # 512 gradient channels, tiny planes (layer 4)
], ids=["temporal", "spatial+addend", "temporal-s2", "norelu", "wide", "n256", "n512"])
def test_dgrad_with_fused_bn_backward_matches_two_pass(geom):
    """dgrad with zsv_bn_bwd_fuse + zsv_bn_bwd_finish == plain dgrad followed by the two-pass zsv_bn_bwd
    (same masks, same rounding points; only the order of the fp32 partial sums differs)."""
    import torch
    from zeroshotvideoclassification_b200 import ops
    cin, cout, k, s, p, T, H, W, relu, use_add = geom
    N = 3
    g = torch.Generator().manual_seed(11)
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    cp_in = ops.cpad(cin)
    w = (torch.randn(cout, cin, *k, generator=g) * 0.05).cuda()
    _, wd = op.pack(w)
    dyo = torch.zeros(N, op.To, op.Ho, op.Wo, ops.cpad(cout), dtype=torch.bfloat16)
    dyo[..., :cout] = torch.randn(N, op.To, op.Ho, op.Wo, cout, generator=g).to(torch.bfloat16)
    dyo = dyo.cuda()
    # the BatchNorm that produced this conv's input: raw y, batch statistics, affine parameters
    y = torch.zeros(N, T, H, W, cp_in, dtype=torch.bfloat16)
    y[..., :cin] = (torch.randn(N, T, H, W, cin, generator=g) * 1.5 + 0.3).to(torch.bfloat16)
    y = y.cuda()
    yf = y[..., :cin].float().reshape(-1, cin)
    mean = torch.zeros(cp_in, device="cuda")
    invstd = torch.zeros(cp_in, device="cuda")
    mean[:cin] = yf.mean(0)
    invstd[:cin] = 1.0 / torch.sqrt(yf.var(0, unbiased=False) + 1e-5)
    gamma = (0.5 + torch.rand(cin, generator=g)).cuda()
    beta = (0.2 * torch.randn(cin, generator=g)).cuda()
    scale = torch.zeros(cp_in, device="cuda")
    shift = torch.zeros(cp_in, device="cuda")
    scale[:cin] = gamma * invstd[:cin]
    shift[:cin] = beta - mean[:cin] * scale[:cin]
    table = torch.stack([scale, shift, invstd, -mean * invstd], dim=1).contiguous()
    addend = None
    if use_add:
        addend = torch.zeros_like(y)
        addend[..., :cin] = torch.randn(N, T, H, W, cin, generator=g).to(torch.bfloat16).cuda()

    gin = op.dgrad(dyo, wd, addend)
    ref_dy, _, _, ref_dg, ref_db, _, _ = ops.bn_bwd(gin, None, 2 if relu else 0, y, mean, invstd, gamma, cin,
                                                    mask_scale=scale, mask_shift=shift)
    dz, partial, rows = op.dgrad_bn_fused(dyo, wd, addend, y, table, relu)
    assert rows >= 1
    got_dy, got_dg, got_db = ops.bn_bwd_finish(dz, y, mean, invstd, gamma, partial, rows, cin)
    torch.cuda.synchronize()
    # dz is the masked plain gradient, bit for bit
    mask = (y.float() * scale + shift) > 0 if relu else torch.ones_like(y, dtype=torch.bool)
    want_dz = torch.where(mask, gin, torch.zeros_like(gin))
    # (an element whose pre-activation rounds differently with / without FMA may flip its mask: allow a handful)
    assert float((dz.view(torch.int16) != want_dz.view(torch.int16)).float().mean()) < 1e-4
    assert rel_err(got_db.cpu(), ref_db.cpu()) < 1e-4
    assert rel_err(got_dg.cpu(), ref_dg.cpu()) < 1e-4
    assert rel_err(got_dy[..., :cin].float().cpu(), ref_dy[..., :cin].float().cpu()) < 1e-2
    assert float(got_dy[..., cin:].abs().max()) == 0.0 if cp_in > cin else True


def test_fprop_bias_addend_relu_epilogue():
    """Folded-BatchNorm inference epilogue: y = relu(conv(x) + bias + addend) in one kernel (resnet.py:110-111)."""
    import torch
    from zeroshotvideoclassification_b200 import ops
    N, T, H, W, cin, cout = 2, 4, 12, 12, 144, 64
    g = torch.Generator().manual_seed(21)
    x = bf16_round(torch.randn(N, cin, T, H, W, generator=g))
    w = bf16_round(torch.randn(cout, cin, 3, 1, 1, generator=g) * 0.05)
    bias = torch.randn(cout, generator=g)
    add = bf16_round(torch.randn(N, cout, T, H, W, generator=g))
    ref = torch.relu(vo.conv3d(x, w, None, (1, 1, 1), (1, 0, 0)) + bias.view(1, -1, 1, 1, 1) + add)
    op = ops.Conv3d(N, T, H, W, cin, cout, (3, 1, 1), (1, 1, 1), (1, 0, 0))
    wf, _ = op.pack(w.cuda(), need_dgrad=False)
    y, _, _ = op.fprop(to_ndhwc(x), wf, stats=False, bias=bias.cuda(), relu=True, addend=to_ndhwc(add))
    torch.cuda.synchronize()
    assert rel_err(from_ndhwc(y, cout), ref) < TOL


PAIR_WGRAD_CASES = [c for c in CASES if c[0] in ("spatial_s2_64_230", "temporal_s2_230_128", "spatial_256_460",
                                                  "spatial_odd_hw", "c3d_27tap", "downsample_64_128")] + [
    # layer-4 shapes: 7x7 planes (a 7x1x2x8 position box fills its 112-row slot), several N tiles, ragged last M tile
    ("l4_spatial_512_600", 9, 2, 7, 7, 512, 600, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    ("l4_temporal_921_512", 9, 2, 7, 7, 921, 512, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    ("l4_ds_256_512", 5, 4, 14, 14, 256, 512, (1, 1, 1), (2, 2, 2), (0, 0, 0)),
    # a box that does not fill its slot (5*3*2 = 30 rows in a 32-row slot: zeroed tail rows) and N/2 not a panel multiple
    ("ragged_box", 1, 2, 3, 5, 200, 176, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
]


@pytest.mark.parametrize("case", PAIR_WGRAD_CASES, ids=lambda c: c[0])
@pytest.mark.parametrize("splits", ["", "1", "3"], ids=["plan", "split1", "split3"])
def test_wgrad_cta_pair_kernel(case, splits, monkeypatch):
    """wgrad_pair_kernel (cluster of two CTAs, cta_group::2 MMAs with M = 256, MN-major operands, each CTA loading its
    own two x panels and half of the dy columns; forced on with ZSV_WGRAD_PAIR=2 so that every geometry it can run is
    covered): the weight gradient must match the oracle with the planned split-K factor, without split-K and with a
    forced one, and agree with the generic single-CTA kernel (ZSV_WGRAD_PAIR=0)."""
    from zeroshotvideoclassification_b200 import ops
    name, N, T, H, W, cin, cout, k, s, p = case
    x, w = _make(case)
    y = vo.conv3d(x, w, None, s, p)
    g = torch.Generator().manual_seed(2)
    dy = bf16_round(torch.randn(y.shape, generator=g))
    _, dw_ref = vo.conv3d_grads(x, w, dy, s, p)
    monkeypatch.setenv("ZSV_WGRAD_PAIR", "2")
    if splits:
        monkeypatch.setenv("ZSV_DEBUG_WGRAD_SPLITS", splits)
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)          # (the workspace size is fixed at construction)
    dw, _ = op.wgrad(to_ndhwc(x), to_ndhwc(dy))
    torch.cuda.synchronize()
    assert rel_err(dw.cpu(), dw_ref) < 2e-3, name
    monkeypatch.setenv("ZSV_WGRAD_PAIR", "0")
    op0 = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    dw0, _ = op0.wgrad(to_ndhwc(x), to_ndhwc(dy))
    torch.cuda.synchronize()
    assert rel_err(dw.cpu(), dw0.cpu()) < 1e-3, name


@pytest.mark.parametrize("case", [c for c in CASES if c[0] in ("spatial_s2_64_230", "spatial_256_460", "temporal_s2_230_128",
                                                                 "downsample_64_128", "c3d_27tap", "spatial_s2_odd")],
                         ids=lambda c: c[0])
@pytest.mark.parametrize("pair", ["1", "0"])
def test_cta_pair_and_single_cta_kernels(case, pair, monkeypatch):
    """igemm_kmajor_kernel<pair> (the default: cluster of two CTAs, tcgen05 cta_group::2 MMAs with M = 256, each CTA
    loading half of the B tile) and the single-CTA kernel (ZSV_2CTA=0): fprop with statistics and dgrad must match the
    oracle, including an odd number of M tiles (the peer's last tile is out of range) and several N tiles."""
    from zeroshotvideoclassification_b200 import ops
    monkeypatch.setenv("ZSV_2CTA", pair)
    name, N, T, H, W, cin, cout, k, s, p = case
    x, w = _make(case)
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    wf, wd = op.pack(w.cuda(), need_dgrad=True)
    y, ps, pq = op.fprop(to_ndhwc(x), wf, stats=True)
    ref = vo.conv3d(x, w, None, s, p)
    assert rel_err(from_ndhwc(y, cout), ref) < TOL
    yb = y[..., :cout].float().reshape(-1, cout).double()
    assert torch.allclose(ps.double().sum(0)[:cout].cpu(), yb.sum(0).cpu(), rtol=1e-4, atol=1e-2)
    g = torch.Generator().manual_seed(4)
    dy = bf16_round(torch.randn(ref.shape, generator=g))
    dx_ref, _ = vo.conv3d_grads(x, w, dy, s, p)
    dx = op.dgrad(to_ndhwc(dy), wd)
    torch.cuda.synchronize()
    assert rel_err(from_ndhwc(dx, cin), dx_ref) < TOL


@pytest.mark.parametrize("case", [c for c in CASES if c[0] in ("spatial_64_64", "spatial_64_144", "temporal_144_64",
                                                                 "temporal_45_64", "spatial_odd_hw")],
                         ids=lambda c: c[0])
def test_halo_cta_pair_kernel(case, monkeypatch):
    """igemm_halo_kernel<pair> (cluster of two CTAs, M = 256, half of the resident weight rows per CTA; chosen by default
    where the weight image starves the activation ring, forced here with ZSV_HALO_2CTA=1): fprop with statistics, dgrad
    and dgrad with the fused BatchNorm backward must match the oracle / the single-CTA kernel, odd M-tile counts
    included (the peer's last tile is out of range)."""
    from zeroshotvideoclassification_b200 import ops
    name, N, T, H, W, cin, cout, k, s, p = case
    x, w = _make(case)
    g = torch.Generator().manual_seed(4)
    ref = vo.conv3d(x, w, None, s, p)
    dy = bf16_round(torch.randn(ref.shape, generator=g))
    tab = torch.rand(cpad(cin), 4, generator=g).cuda()
    monkeypatch.setenv("ZSV_HALO_2CTA", "0")
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)       # (the statistics row count is fixed at construction)
    wf, wd = op.pack(w.cuda(), need_dgrad=True)
    y0, ps0, _ = op.fprop(to_ndhwc(x), wf, stats=True)
    dx0 = op.dgrad(to_ndhwc(dy), wd)
    dz0, part0, r0 = op.dgrad_bn_fused(to_ndhwc(dy), wd, None, to_ndhwc(x), tab, True)
    monkeypatch.setenv("ZSV_HALO_2CTA", "1")
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    y, ps, pq = op.fprop(to_ndhwc(x), wf, stats=True)
    dx = op.dgrad(to_ndhwc(dy), wd)
    dz, part, r1 = op.dgrad_bn_fused(to_ndhwc(dy), wd, None, to_ndhwc(x), tab, True)
    torch.cuda.synchronize()
    assert rel_err(from_ndhwc(y, cout), ref) < TOL
    # (not bit-equal: where the single-CTA plan does not fit, the generic kernel runs and sums K in another order)
    for a, b in ((y, y0), (dx, dx0), (dz, dz0)):
        assert rel_err(a.float().cpu(), b.float().cpu()) < 5e-3
    assert torch.allclose(ps.double().sum(0), ps0.double().sum(0), rtol=1e-3, atol=5e-2)
    assert torch.allclose(part[:r1, :2].double().sum(0), part0[:r0, :2].double().sum(0), rtol=1e-3, atol=5e-2)   # rows 2-3: finish pass
    dx_ref, _ = vo.conv3d_grads(x, w, dy, s, p)
    assert rel_err(from_ndhwc(dx, cin), dx_ref) < TOL


@pytest.mark.parametrize("case", [c for c in CASES if c[0] in ("spatial_64_64", "spatial_64_144", "c3d_27tap")] +
                         [("spatial_w56_ragged_h", 1, 3, 13, 56, 72, 40, (1, 3, 3), (1, 1, 1), (0, 1, 1))],
                         ids=lambda c: c[0])
def test_halo_w_taps_by_descriptor_offset(case, monkeypatch):
    """Halo kernel with the W taps served from ONE widened box (descriptor start offsets of one row, 8-row groups
    (8 + kw - 1) rows apart) against the same kernel with kw W-shifted copies (ZSV_HALO_WSHIFT=0): same MMAs in the
    same order, so fprop / dgrad / fused-BN dgrad are bit-identical; and both match the oracle."""
    from zeroshotvideoclassification_b200 import ops
    name, N, T, H, W, cin, cout, k, s, p = case
    x, w = _make(case)
    g = torch.Generator().manual_seed(4)
    ref = vo.conv3d(x, w, None, s, p)
    dy = bf16_round(torch.randn(ref.shape, generator=g))
    tab = torch.rand(cpad(cin), 4, generator=g).cuda()
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("ZSV_HALO_WSHIFT", mode)
        op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
        wf, wd = op.pack(w.cuda(), need_dgrad=True)
        y, ps, pq = op.fprop(to_ndhwc(x), wf, stats=True)
        dx = op.dgrad(to_ndhwc(dy), wd)
        dz, part, r = op.dgrad_bn_fused(to_ndhwc(dy), wd, None, to_ndhwc(x), tab, True)
        torch.cuda.synchronize()
        outs[mode] = (y, dx, dz, ps.double().sum(0), part[:r, :2].double().sum(0))
    assert rel_err(from_ndhwc(outs["1"][0], cout), ref) < TOL
    dx_ref, _ = vo.conv3d_grads(x, w, dy, s, p)
    assert rel_err(from_ndhwc(outs["1"][1], cin), dx_ref) < TOL
    for a, b in zip(outs["0"][:3], outs["1"][:3]):
        assert torch.equal(a, b)
    assert torch.allclose(outs["0"][3], outs["1"][3], rtol=1e-6, atol=1e-4)
    assert torch.allclose(outs["0"][4], outs["1"][4], rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("geom", [
    (3, 8, 24, 24, 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),     # narrow tiles: split epilogue, single CTA
    (3, 8, 24, 24, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),    # CTA pair; dgrad narrow (split), fprop wide
    (2, 8, 24, 24, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),    # pair chosen for the narrow temporal fprop
    (4, 4, 14, 14, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1)),   # generic K-major kernel, wide tiles
], ids=lambda g: f"{g[4]}_{g[5]}")
def test_passes_are_run_to_run_deterministic(geom):
    """The warp-specialised kernels hand tiles between roles (TMA producer, MMA issuer, convert / finish epilogue groups)
    through mbarriers; a protocol bug shows up as run-to-run differences.  Twelve runs of every pass on the same inputs
    must be bit-identical (tools/stress_determinism.py is the longer version at layer-1 sizes)."""
    from zeroshotvideoclassification_b200 import ops
    N, T, H, W, cin, cout, k, s, p = geom
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(N, T, H, W, cpad(cin), device="cuda", generator=g).to(torch.bfloat16)
    x[..., cin:] = 0
    w = torch.randn(cout, cin, *k, device="cuda", generator=g) * 0.05
    wf, wd = op.pack(w)
    y0, _, _ = op.fprop(x, wf, stats=True)
    dy = torch.randn(y0.shape, device="cuda", generator=g).to(torch.bfloat16)
    tab = torch.rand(cpad(cin), 4, device="cuda", generator=g)
    ref = None
    for _ in range(12):
        y, ps, pq = op.fprop(x, wf, stats=True)
        dz, part, r = op.dgrad_bn_fused(dy, wd, None, x, tab, True)
        dw, _ = op.wgrad(x, dy)
        torch.cuda.synchronize()
        outs = [t.clone() for t in (y, ps.sum(0), pq.sum(0), op.dgrad(dy, wd), dz, part[:r, :2].sum(0), dw)]
        if ref is None:
            ref = outs
        else:
            for i, (a, b) in enumerate(zip(outs, ref)):
                assert torch.equal(a, b), f"output {i} differs between runs"
