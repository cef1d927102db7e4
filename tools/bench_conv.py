"""GPU micro-benchmark of single convolutions through the C ABI (fprop / dgrad / wgrad): time and TFLOP/s.

usage: python tools/bench_conv.py "N,T,H,W,Cin,Cout,kt,kh,kw,st,sh,sw,pt,ph,pw" ...
"""
import sys

import torch

sys.path.insert(0, ".")
from zeroshotvideoclassification_b200 import ops


def bench(fn, iters=10):
    """us per call, timed as a CUDA-graph replay of `iters` back-to-back calls (no host launch overhead in the number)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * iters) * 1e3  # us


for spec in sys.argv[1:]:
    v = [int(t) for t in spec.split(",")]
    N, T, H, W, cin, cout = v[:6]
    k, s, p = tuple(v[6:9]), tuple(v[9:12]), tuple(v[12:15])
    op = ops.Conv3d(N, T, H, W, cin, cout, k, s, p)
    x = torch.randn(N, T, H, W, ops.cpad(cin), device="cuda").to(torch.bfloat16)
    w = torch.randn(cout, cin, *k, device="cuda") * 0.05
    wf, wd = op.pack(w)
    y, _, _ = op.fprop(x, wf, stats=True)
    dy = torch.randn_like(y)
    flops = 2.0 * op.out_positions * cout * cin * k[0] * k[1] * k[2]
    t_f = bench(lambda: op.fprop(x, wf, stats=True))
    t_fn = bench(lambda: op.fprop(x, wf, stats=False))
    t_d = bench(lambda: op.dgrad(dy, wd))
    t_w = bench(lambda: op.wgrad(x, dy))
    t_df = float("nan")
    if wd is not None and s == (1, 1, 1):
        tab = torch.rand(ops.cpad(cin), 4, device="cuda")           # (mean, invstd, gamma, beta) rows of bn_finalize
        t_df = bench(lambda: op.dgrad_bn_fused(dy, wd, None, x, tab, True))
    print(f"{spec:48s} fprop {t_f:8.1f} us {flops / t_f / 1e6:7.1f} TF/s (no stats {t_fn:8.1f} us) | dgrad {t_d:8.1f} us {flops / t_d / 1e6:7.1f} TF/s"
          f" (bn-fused {t_df:8.1f} us) | wgrad {t_w:8.1f} us {flops / t_w / 1e6:7.1f} TF/s")
