import os, sys, torch
sys.path.insert(0, "/root/repo")
os.chdir("/root/repo")
from tests.test_gpu_conv import CASES, _make, vo, bf16_round, to_ndhwc, from_ndhwc, cpad
from zeroshotvideoclassification_b200 import ops
case = [c for c in CASES if c[0] == "spatial_64_144"][0]
name, N, T, H, W, cin, cout, k, s, p = case
print(case)
x, w = _make(case)
g = torch.Generator().manual_seed(4)
ref = vo.conv3d(x, w, None, s, p)
dy = bf16_round(torch.randn(ref.shape, generator=g))
tab = torch.rand(cpad(cin), 4, generator=g).cuda()
out = {}
for ny in ("1", "2"):
    os.environ["ZSV_NYBUF"] = ny
    for mode in ("0", "1"):
        os.environ["ZSV_HALO_WSHIFT"] = mode
        op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
        wf, wd = op.pack(w.cuda(), need_dgrad=True)
        dz, part, r = op.dgrad_bn_fused(to_ndhwc(dy), wd, None, to_ndhwc(x), tab, True)
        torch.cuda.synchronize()
        out[(ny, mode)] = dz.clone()
base = out[("1", "1")]
for key, v in out.items():
    d = (v.float() - base.float()).abs()
    idx = (d > 0).nonzero()
    print(key, "ndiff", idx.shape[0])
    if idx.shape[0]:
        print(" first", idx[:6].tolist(), " last", idx[-3:].tolist())
        for dim in range(5):
            print("  dim", dim, "unique", idx[:, dim].unique().tolist()[:40])
