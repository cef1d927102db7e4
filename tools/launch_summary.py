"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of the last step."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
idx = [i for i, n in enumerate(names) if "ncdhw_to_ndhwc" in n]
sel = rows[idx[-1]:] if idx else rows


def short(n):
    m = re.search(r"(\w+_kernel|\w+Kernel\w*|multi_tensor\w*|\w+elementwise\w*)", n)
    return m.group(1) if m else n[:60]


agg = collections.defaultdict(lambda: [0, 0.0])
for r in sel:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    agg[short(r["Kernel Name"])][0] += 1
    agg[short(r["Kernel Name"])][1] += v
tot = sum(v[1] for v in agg.values())
print(f"one training step (last in capture): {len(sel)} launches, {tot:.0f} us of kernel time (serialised, cold cache)")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  x{c:4d}  {n}")
