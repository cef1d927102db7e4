"""The evaluation-size nearest-class call a few times, for an `ncu -k regex:nearest -s 2 -c 1` capture; prints the
back-to-back device time per call (CUDA events around 20 calls).
usage: python tools/ncu_nearest.py [N] [C]"""
import sys

import torch

sys.path.insert(0, ".")
from zeroshotvideoclassification_b200 import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 101
emb = torch.nn.functional.normalize(torch.randn(N, 300, device="cuda"))
cls = torch.nn.functional.normalize(torch.randn(C, 300, device="cuda"))
for _ in range(4):
    ops.nearest_class(emb, cls, 5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.nearest_class(emb, cls, 5)
e1.record()
torch.cuda.synchronize()
print(f"N={N} C={C}: {1e3 * e0.elapsed_time(e1) / 20:.1f} us per call (20 back to back)")
