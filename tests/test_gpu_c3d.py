"""C3D (network.py:95-180) on the CUDA path vs the CPU oracle (eval mode: Dropout is not reproducible across
devices).  C3D has no BatchNorm, so there is no chaotic amplification: plain fp32 oracle."""
import pytest
import torch
import torch.nn.functional as F

from oracle import video_oracle as vo
from tests.helpers import rel_err, rms_rel_err

pytestmark = pytest.mark.gpu


def test_c3d_forward_backward_vs_oracle():
    from types import SimpleNamespace
    from zeroshotvideoclassification_b200 import video_models as vm
    torch.manual_seed(0)
    model = vm.get_network(SimpleNamespace(network="c3d", fixconvs=False, nopretrained=False))
    model.dropout.p = 0.0
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 1, 3, 16, 112, 112, generator=g)
    z = F.normalize(torch.randn(2, 300, generator=g))

    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    emb_ref = vo.c3d_forward(params, x, train=False)
    loss_ref = vo.mse_loss(emb_ref, z)
    loss_ref.backward()

    model = model.cuda().train()
    emb = model(x.cuda())
    loss = torch.nn.MSELoss()(emb, z.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert emb.shape == (2, 300)
    e = rel_err(emb.detach().cpu(), emb_ref.detach())
    print("c3d emb rel err", e, "loss", float(loss.detach()), float(loss_ref))
    assert e < 2e-2
    assert abs(float(loss.detach()) - float(loss_ref)) < 2e-2 * abs(float(loss_ref))
    errs = {}
    for name, p in model.named_parameters():
        ref = params[name].grad
        if ref is None:
            assert p.grad is None, name          # fc7 / fc8 are dead (network.py:168-172)
            continue
        errs[name] = rms_rel_err(p.grad.cpu(), ref)
    print({k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) < 8e-2, errs


def test_c3d_multi_clip_eval():
    """--evaluate uses several clips per video; C3D averages them (network.py:174-176)."""
    from types import SimpleNamespace
    from zeroshotvideoclassification_b200 import video_models as vm
    torch.manual_seed(1)
    model = vm.get_network(SimpleNamespace(network="c3d", fixconvs=False, nopretrained=False))
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x = torch.randn(2, 2, 3, 16, 112, 112, generator=torch.Generator().manual_seed(2))
    ref = vo.c3d_forward(sd, x, train=False)
    model = model.cuda().eval()
    with torch.no_grad():
        emb = model(x.cuda())
    assert emb.shape == (2, 300)
    assert rel_err(emb.cpu(), ref) < 2e-2
