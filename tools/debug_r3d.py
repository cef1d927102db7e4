"""Bisect aid: r3d_18 gradient error vs the rounding-matched oracle, per parameter."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from oracle import video_oracle as vo
from tests.helpers import rel_err, rms_rel_err
from zeroshotvideoclassification_b200 import video_models as vm

B, T, H, W = 3, 8, 64, 64
torch.manual_seed(5)
model = vm.get_network(vm.default_opt("r3d_18"))
g = torch.Generator().manual_seed(6)
for name, m in model.named_modules():
    if isinstance(m, torch.nn.BatchNorm3d):
        m.weight.data = 0.5 + torch.rand(m.weight.shape, generator=g)
        m.bias.data = 0.2 * torch.randn(m.bias.shape, generator=g)
        if name.endswith("conv2.1"):
            m.weight.data *= 0.1
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
x = torch.randn(B, 1, 3, T, H, W, generator=g)
cls = F.normalize(torch.randn(101, 300, generator=g))
z = cls[torch.randint(0, 101, (B,), generator=g)]
emb_emu, loss_emu, grads_emu = vo.train_step_grads({k: v.clone() for k, v in sd.items()}, x, z, emulate_bf16=True, arch="r3d_18")
emb32, loss32, grads32 = vo.train_step_grads({k: v.clone() for k, v in sd.items()}, x, z, arch="r3d_18")
model = model.cuda().train()
emb, _ = model(x.cuda())
loss = torch.nn.MSELoss()(emb, z.cuda())
loss.backward()
torch.cuda.synchronize()
print("emb err", rel_err(emb.detach().cpu(), emb_emu), "emu-vs-fp32 emb", rel_err(emb_emu, emb32))
for name, p in model.named_parameters():
    if name in grads_emu:
        print(f"{name:45s} ours-vs-emu {rms_rel_err(p.grad.cpu(), grads_emu[name]):.3e}   emu-vs-fp32 {rms_rel_err(grads_emu[name], grads32[name]):.3e}")
