"""Run a fixed list of single convolutions (bs=22 shapes of R(2+1)D-18) once each through the C ABI, for
`ncu --set full -k regex:igemm|wgrad_mnmajor` captures.  Prints the launch order so report IDs can be matched."""
import sys

import torch

sys.path.insert(0, ".")
from zeroshotvideoclassification_b200 import _lib, ops

N = 22
LAYERS = [
    # name, T,H,W, cin,cout, k, s, p, which passes
    ("stem.3 temporal 45->64", 16, 56, 56, 45, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), "fdw"),
    ("layer1 temporal 144->64", 16, 56, 56, 144, 64, (3, 1, 1), (1, 1, 1), (1, 0, 0), "fdw"),
    ("layer1 spatial 64->144", 16, 56, 56, 64, 144, (1, 3, 3), (1, 1, 1), (0, 1, 1), "fdw"),
    ("layer2.0 spatial s2 64->230", 16, 56, 56, 64, 230, (1, 3, 3), (1, 2, 2), (0, 1, 1), "d"),
    ("layer4.1 temporal 1152->512", 2, 7, 7, 1152, 512, (3, 1, 1), (1, 1, 1), (1, 0, 0), "w"),
    ("layer2.1 spatial 128->288", 8, 28, 28, 128, 288, (1, 3, 3), (1, 1, 1), (0, 1, 1), "f"),
]
order = []
for name, T, H, W, cin, cout, k, s, p, which in LAYERS:
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    x = torch.randn(N, T, H, W, ops.cpad(cin), device="cuda").to(torch.bfloat16)
    if ops.cpad(cin) != cin:
        x[..., cin:] = 0
    w = torch.randn(cout, cin, *k, device="cuda") * 0.05
    wf, wd = op.pack(w)
    dy = torch.randn(N, op.To, op.Ho, op.Wo, ops.cpad(cout), device="cuda").to(torch.bfloat16)
    if ops.cpad(cout) != cout:
        dy[..., cout:] = 0
    torch.cuda.synchronize()
    if "f" in which:
        op.fprop(x, wf, stats=True)
        order.append(f"fprop {name}")
    if "d" in which:
        n0 = _lib.launch_count()
        op.dgrad(dy, wd)
        order += [f"dgrad {name}"] * (_lib.launch_count() - n0)
    if "w" in which:
        op.wgrad(x, dy)
        order.append(f"wgrad {name}")
    torch.cuda.synchronize()
    del x, dy
for i, o in enumerate(order):
    print(i, o)
