"""Whole-model parity: R(2+1)D-18 forward + backward on the CUDA path vs the CPU fp32 oracle on identical
random-init weights and synthetic clips (north_star: per-layer activations / gradients within 1e-2 relative,
bf16 compute with fp32 accumulate)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import video_oracle as vo
from tests.helpers import from_ndhwc, rel_err, rms_rel_err

pytestmark = pytest.mark.gpu


def _setup(B, T, H, W, seed=0):
    from zeroshotvideoclassification_b200 import video_models as vm
    torch.manual_seed(seed)
    model = vm.get_network(vm.default_opt("r2plus1d_18"))
    # make BN affine parameters non-trivial so their gradients are exercised
    g = torch.Generator().manual_seed(seed + 1)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm3d):
            m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
            m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x = torch.randn(B, 1, 3, T, H, W, generator=g)
    cls = F.normalize(torch.randn(101, 300, generator=g))
    labels = torch.randint(0, 101, (B,), generator=g)
    return model, sd, x, cls[labels]


def test_state_dict_contract():
    """Names/shapes of the drop-in module tree are those of the reference's network.Model (golden key list)."""
    import json, os
    from zeroshotvideoclassification_b200 import video_models as vm
    model = vm.get_network(vm.default_opt("r2plus1d_18"))
    path = os.path.join(os.path.dirname(__file__), "golden", "r2plus1d_state_dict_keys.json")
    golden = json.load(open(path))
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert got == golden


@pytest.mark.parametrize("shape", [(4, 8, 64, 64), (2, 16, 112, 112)], ids=["small", "full_clip_bs2"])
def test_forward_backward_vs_oracle(shape):
    B, T, H, W = shape
    model, sd, x, z = _setup(B, T, H, W)
    # ---- oracle (CPU fp32) ----
    trace = {}
    emb_ref, loss_ref, grads_ref = vo.train_step_grads(sd, x, z, trace=trace)
    # ---- CUDA path, exactly main.py:170-195 without the optimizer ----
    model = model.cuda().train()
    crit = torch.nn.MSELoss()
    out = model(x.cuda())
    emb = out[0] if isinstance(out, tuple) else out
    loss = crit(emb, z.cuda())
    loss.backward()
    torch.cuda.synchronize()

    assert emb.shape == (B, 300) and emb.dtype == torch.float32
    e_emb = rel_err(emb.detach().cpu(), emb_ref)
    e_loss = abs(float(loss) - float(loss_ref)) / abs(float(loss_ref))
    print(f"emb rel err {e_emb:.3e}  loss {float(loss):.6f} vs {float(loss_ref):.6f}")
    assert e_emb < 2e-2
    assert e_loss < 2e-2

    worst = {}
    for name, p in model.named_parameters():
        if name not in grads_ref:
            assert p.grad is None, f"{name} is dead in the reference (network.py:500-517) but got a gradient"
            continue
        assert p.grad is not None, name
        assert p.grad.dtype == torch.float32 and p.grad.shape == p.shape
        worst[name] = (rms_rel_err(p.grad.cpu(), grads_ref[name]), rel_err(p.grad.cpu(), grads_ref[name]))
    bad = {k: v for k, v in worst.items() if not (v[0] < 3e-2)}
    top = sorted(worst.items(), key=lambda kv: -kv[1][0])[:8]
    print("worst gradient errors (rms-rel, max-rel):", top)
    assert not bad, bad

    # BatchNorm running statistics were updated like the reference (momentum 0.1, unbiased variance)
    sd_after = model.state_dict()
    for k in ("model.stem.1.running_mean", "model.layer2.0.downsample.1.running_var", "model.layer4.1.conv2.1.running_var"):
        ref_key = k
        assert rel_err(sd_after[k].cpu(), sd[ref_key]) < 2e-2, k   # sd was updated in place by the oracle
    assert int(sd_after["model.stem.1.num_batches_tracked"]) == 1


def test_per_layer_activations_small():
    """Per-layer activation parity through the BackboneRunner tape (every conv output and block output)."""
    from zeroshotvideoclassification_b200 import engine
    B, T, H, W = 2, 8, 48, 48
    model, sd, x, z = _setup(B, T, H, W, seed=3)
    trace = {}
    with torch.no_grad():
        vo.model_forward({k: v.clone() for k, v in sd.items()}, x, train=True, trace=trace)
    model = model.cuda().train()
    tensors = {k: v.detach() for k, v in engine._module_tensors(model.model).items()}
    runner = engine.BackboneRunner(tensors, train=True, need_grad=True)
    feats = runner.forward(x[:, 0].cuda())
    torch.cuda.synchronize()
    errs = {}
    for rec in runner.stem_recs:
        errs[rec.spec.name] = rel_err(from_ndhwc(rec.y, rec.spec.cout), trace[rec.spec.name])
    for b in runner.block_recs:
        for rec in b.units + ([b.ds] if b.ds is not None else []):
            errs[rec.spec.name] = rel_err(from_ndhwc(rec.y, rec.spec.cout), trace[rec.spec.name])
    for spec, b in zip(engine.BLOCK_SPECS, runner.block_recs):
        errs[spec.prefix] = rel_err(from_ndhwc(b.out, spec.convs[3].cout), trace[spec.prefix])
    print("per-layer activation rel err:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < 3e-2, errs
    assert rel_err(from_ndhwc(feats, 512), trace["feats"]) < 3e-2


def test_eval_mode_and_no_grad():
    """evaluate() path (main.py:229-250): eval-mode BN uses running statistics, no autograd state kept."""
    B, T, H, W = 3, 8, 32, 32
    model, sd, x, z = _setup(B, T, H, W, seed=5)
    g = torch.Generator().manual_seed(9)
    for k, v in sd.items():   # non-trivial running statistics
        if k.endswith("running_mean"):
            v.copy_(0.1 * torch.randn(v.shape, generator=g))
        if k.endswith("running_var"):
            v.copy_(0.5 + torch.rand(v.shape, generator=g))
    model.load_state_dict(sd)
    ref = vo.model_forward(sd, x, train=False)
    model = model.cuda().eval()
    with torch.no_grad():
        emb, none = model(x.cuda())
    assert none is None and not emb.requires_grad
    assert rel_err(emb.cpu(), ref) < 2e-2
    sd_after = model.state_dict()
    assert torch.equal(sd_after["model.stem.1.running_mean"].cpu(), sd["model.stem.1.running_mean"])


def test_variable_batch_and_fixconvs():
    """main.py:157-158 filters broken samples -> any B >= 1; --fixconvs freezes the backbone (network.py:482-484)."""
    from types import SimpleNamespace
    from zeroshotvideoclassification_b200 import video_models as vm
    torch.manual_seed(0)
    model = vm.get_network(SimpleNamespace(network="r2plus1d_18", fixconvs=True, nopretrained=False)).cuda().train()
    for B in (1, 3):
        x = torch.randn(B, 1, 3, 8, 32, 32, device="cuda")
        emb, _ = model(x)
        emb.square().mean().backward()
        assert emb.shape == (B, 300)
    assert model.model.stem[0].weight.grad is None
    assert model.output2emb_proj.layers[0].weight.grad is not None
