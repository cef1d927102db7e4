"""GPU debugging aid: one residual block forward+backward, stage-by-stage error vs the rounding-matched oracle."""
import sys
import torch

sys.path.insert(0, ".")
from oracle import video_oracle as vo
from tests.helpers import bf16_round, from_ndhwc, rel_err, rms_rel_err, to_ndhwc
from tests.test_gpu_model import _block_tensors
from zeroshotvideoclassification_b200 import engine, ops

bi = int(sys.argv[1]) if len(sys.argv) > 1 else 1
spec = engine.BLOCK_SPECS[bi]
g = torch.Generator().manual_seed(100 + bi)
cin = spec.convs[0].cin
N, T, H, W = 3, 4, 12, 12
x = bf16_round(torch.randn(N, cin, T, H, W, generator=g).abs())
t = _block_tensors(spec, g)
stride = spec.convs[0].stride[1]
params = {k: v.clone().requires_grad_(True) for k, v in t.items() if v.is_floating_point() and "running" not in k}
work = {k: v.clone() for k, v in t.items()}
work.update(params)
xr = x.clone().requires_grad_(True)
trace = {}
net = vo._Net(work, True, trace, emulate_bf16=True)
out_ref = vo._basic_block(net, xr, spec.prefix, stride, spec.downsample is not None)
gout = bf16_round(torch.randn(out_ref.shape, generator=g))
out_ref.backward(gout)


def show(name, got, ref):
    print(f"{name:40s} max-rel {rel_err(got, ref):.3e}  rms-rel {rms_rel_err(got, ref):.3e}  |ref|max {float(ref.abs().max()):.3e}")


tens = {k: v.cuda() for k, v in t.items()}
runner = engine.BackboneRunner(tens, train=True, need_grad=True)
out, dims = runner._block(spec, to_ndhwc(x), (N, T, H, W))
brec = runner.block_recs[0]
print("---- forward ----")
for rec in brec.units + ([brec.ds] if brec.ds else []):
    show("y " + rec.spec.name, from_ndhwc(rec.y, rec.spec.cout), trace[rec.spec.name].detach())
    if rec.out is not None:
        show("out " + rec.spec.bn, from_ndhwc(rec.out, rec.spec.cout), trace[rec.spec.bn + ":out"].detach())
show("block out", from_ndhwc(out, spec.convs[3].cout), out_ref.detach())

print("---- backward ----")
grads = {}
want = {}
gd = to_ndhwc(gout)
tail, ds = brec.units[3], brec.ds
if ds is not None:
    dy_t, dy_d, _, dg, db, dg2, db2 = ops.bn_bwd(gd, brec.out, True, tail.y, tail.mean, tail.invstd,
                                                 tens[tail.spec.bn + ".weight"], tail.spec.cout, y2=ds.y, mean2=ds.mean,
                                                 invstd2=ds.invstd, gamma2=tens[ds.spec.bn + ".weight"])
    show("dy downsample conv", from_ndhwc(dy_d, ds.spec.cout), trace[ds.spec.name].grad)
    dz = None
else:
    dy_t, _, dz, dg, db, _, _ = ops.bn_bwd(gd, brec.out, True, tail.y, tail.mean, tail.invstd,
                                           tens[tail.spec.bn + ".weight"], tail.spec.cout, want_dz=True)
show("dy tail conv (T2)", from_ndhwc(dy_t, tail.spec.cout), trace[tail.spec.name].grad)
show("dgamma tail", dg.cpu(), params[tail.spec.bn + ".weight"].grad)
ga = runner._conv_bwd(tail, dy_t, grads, want)
show("g wrt S2 act", from_ndhwc(ga, tail.spec.cin), trace[brec.units[2].spec.bn + ":out"].grad)
show("dw T2", grads[tail.spec.name + ".weight"].cpu(), params[tail.spec.name + ".weight"].grad)
u = brec.units[2]
dy, _, _, dgm, dbt, _, _ = ops.bn_bwd(ga, u.out, True, u.y, u.mean, u.invstd, tens[u.spec.bn + ".weight"], u.spec.cout)
show("dy S2 conv", from_ndhwc(dy, u.spec.cout), trace[u.spec.name].grad)
ga = runner._conv_bwd(u, dy, grads, want)
show("g wrt T1 act", from_ndhwc(ga, u.spec.cin), trace[brec.units[1].spec.bn + ":out"].grad)
u = brec.units[1]
dy, _, _, dgm, dbt, _, _ = ops.bn_bwd(ga, u.out, True, u.y, u.mean, u.invstd, tens[u.spec.bn + ".weight"], u.spec.cout)
show("dy T1 conv", from_ndhwc(dy, u.spec.cout), trace[u.spec.name].grad)
ga = runner._conv_bwd(u, dy, grads, want)
show("g wrt S1 act", from_ndhwc(ga, u.spec.cin), trace[brec.units[0].spec.bn + ":out"].grad)
u = brec.units[0]
dy, _, _, dgm, dbt, _, _ = ops.bn_bwd(ga, u.out, True, u.y, u.mean, u.invstd, tens[u.spec.bn + ".weight"], u.spec.cout)
show("dy S1 conv", from_ndhwc(dy, u.spec.cout), trace[u.spec.name].grad)
gx = runner._conv_bwd(u, dy, grads, want, addend=dz)
if ds is not None:
    gx = runner._conv_bwd(ds, dy_d, grads, want, addend=gx)
show("g wrt block input", from_ndhwc(gx, cin), xr.grad)
for k in sorted(grads):
    show("dw " + k, grads[k].cpu().reshape(params[k].shape), params[k].grad)
