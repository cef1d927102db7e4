"""One convolution pass a few times, for an `ncu -k regex:... -s 2 -c 1` capture.
usage: python tools/ncu_one.py "N,T,H,W,Cin,Cout,kt,kh,kw,st,sh,sw,pt,ph,pw" fprop|dgrad|dgrad_fused|wgrad"""
import sys

import torch

sys.path.insert(0, ".")
from zeroshotvideoclassification_b200 import ops

v = [int(t) for t in sys.argv[1].split(",")]
N, T, H, W, cin, cout = v[:6]
k, s, p = tuple(v[6:9]), tuple(v[9:12]), tuple(v[12:15])
op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
x = torch.randn(N, T, H, W, ops.cpad(cin), device="cuda").to(torch.bfloat16)
w = torch.randn(cout, cin, *k, device="cuda") * 0.05
wf, wd = op.pack(w)
dy = torch.randn(N, op.To, op.Ho, op.Wo, ops.cpad(cout), device="cuda").to(torch.bfloat16)
tab = torch.rand(ops.cpad(cin), 4, device="cuda")
for _ in range(4):
    if sys.argv[2] == "fprop":
        op.fprop(x, wf, stats=True)
    elif sys.argv[2] == "dgrad":
        op.dgrad(dy, wd)
    elif sys.argv[2] == "dgrad_fused":
        op.dgrad_bn_fused(dy, wd, None, x, tab, True)
    else:
        op.wgrad(x, dy)
torch.cuda.synchronize()
