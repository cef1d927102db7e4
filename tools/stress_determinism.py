"""Race check without a sanitizer: the warp-specialised kernels hand tiles between roles with mbarriers, so a protocol bug
shows up as run-to-run differences.  Every pass of a few layer shapes (both epilogue variants, CTA pairs, fused BatchNorm
backward) is run 30 times on fresh random inputs' and all outputs must be bit-identical to the first run."""
import sys

import torch

sys.path.insert(0, ".")
from zeroshotvideoclassification_b200 import ops

SPECS = [  # N,T,H,W,Cin,Cout,k,s,p
    (4, 16, 56, 56, 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    (4, 16, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    (4, 16, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0)),
    (6, 8, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
    (6, 8, 28, 28, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1)),
    (8, 2, 7, 7, 512, 1152, (1, 3, 3), (1, 1, 1), (0, 1, 1)),
]
bad = 0
for N, T, H, W, cin, cout, k, s, p in SPECS:
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, T, H, W, ops.cpad(cin), device="cuda", generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, *k, device="cuda", generator=g) * 0.05
    wf, wd = op.pack(w)
    y0, ps0, pq0 = op.fprop(x, wf, stats=True)
    dy = torch.randn(y0.shape, device="cuda", generator=g).to(torch.bfloat16)
    tab = torch.rand(ops.cpad(cin), 4, device="cuda", generator=g)
    ref = None
    for it in range(30):
        y, ps, pq = op.fprop(x, wf, stats=True)
        dx = op.dgrad(dy, wd)
        outs = [y, ps.sum(0), pq.sum(0), dx]
        if s == (1, 1, 1):
            dz, part, r = op.dgrad_bn_fused(dy, wd, None, x, tab, True)
            outs += [dz, part[:r, :2].sum(0)]
        dw, _ = op.wgrad(x, dy)
        outs.append(dw)
        torch.cuda.synchronize()
        outs = [o.clone() for o in outs]
        if ref is None:
            ref = outs
        else:
            for i, (a, b) in enumerate(zip(outs, ref)):
                if not torch.equal(a, b):
                    bad += 1
                    print(f"spec {cin}->{cout} {k} run {it}: output {i} differs ({int((a != b).sum())} elements)")
    print(f"{cin}->{cout} {k} s{s}: 30 runs identical" if bad == 0 else f"{cin}->{cout}: DIFFERENCES")
print("determinism ok" if bad == 0 else f"{bad} differences")
sys.exit(1 if bad else 0)
