"""GPU micro-benchmark of the HBM-bound BatchNorm kernels: achieved GB/s vs algorithmic bytes."""
import sys

import torch

sys.path.insert(0, ".")
from zeroshotvideoclassification_b200 import ops


def bench(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for rows, C in ((22 * 16 * 56 * 56, 144), (22 * 16 * 56 * 56, 64), (22 * 8 * 28 * 28, 288), (22 * 8 * 28 * 28, 128),
                (22 * 4 * 14 * 14, 576), (22 * 2 * 7 * 7, 1152)):
    cp = ops.cpad(C)
    y = torch.randn(rows, cp, device="cuda").to(torch.bfloat16)
    g = torch.randn(rows, cp, device="cuda").to(torch.bfloat16)
    res = torch.randn(rows, cp, device="cuda").to(torch.bfloat16)
    sc = torch.rand(cp, device="cuda") + 0.5
    sh = torch.randn(cp, device="cuda") * 0.1
    mean = torch.zeros(cp, device="cuda")
    invstd = torch.ones(cp, device="cuda")
    gamma = torch.ones(C, device="cuda")
    nbytes = rows * cp * 2
    t_a = bench(lambda: ops.bn_apply(y, sc, sh, C, True))
    t_ar = bench(lambda: ops.bn_apply(y, sc, sh, C, True, residual=res))
    out = ops.bn_apply(y, sc, sh, C, True)
    t_b2 = bench(lambda: ops.bn_bwd(g, None, 2, y, mean, invstd, gamma, C, mask_scale=sc, mask_shift=sh))
    t_b1 = bench(lambda: ops.bn_bwd(g, out, 1, y, mean, invstd, gamma, C, want_dz=True))
    print(f"rows {rows:8d} C {C:5d}: apply {t_a:7.1f} us {2 * nbytes / t_a / 1e3:6.0f} GB/s | apply+res {t_ar:7.1f} us "
          f"{3 * nbytes / t_ar / 1e3:6.0f} GB/s | bwd(mask from y) {t_b2:7.1f} us {5 * nbytes / t_b2 / 1e3:6.0f} GB/s | "
          f"bwd(tail,+dz) {t_b1:7.1f} us {8 * nbytes / t_b1 / 1e3:6.0f} GB/s")
