"""Per-node latency of a chain of small dependent kernels replayed from a CUDA graph (ZSV_PDL=0|1): does the
programmatic-dependent-launch attribute survive stream capture, and what does it buy?"""
import os
import sys

import torch

sys.path.insert(0, ".")
from zeroshotvideoclassification_b200 import ops

dev = torch.device("cuda")
x = torch.randn(22, 2, 7, 7, 512, device=dev).to(torch.bfloat16)          # a layer-4 activation (2.2 MB)
scale = torch.ones(512, device=dev)
shift = torch.zeros(512, device=dev)
n = 400
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        y = x
        for _ in range(8):
            y = ops.bn_apply(y, scale, shift, 512, True)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        y = x
        for _ in range(n):
            y = ops.bn_apply(y, scale, shift, 512, True)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    g.replay()
e1.record()
torch.cuda.synchronize()
print(f"ZSV_PDL={os.environ.get('ZSV_PDL', '1')}: {1e3 * e0.elapsed_time(e1) / (10 * n):.2f} us per dependent bn_apply node (2.2 MB tensor)")
