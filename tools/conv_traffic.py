"""DRAM traffic and tensor-pipe activity of the implicit-GEMM conv kernels (fprop + dgrad) of one training step, from
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,
sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active -k regex:igemm --csv` output.
Writes profiles/r02_conv_dram_traffic.json (with the fingerprint of the csrc/ build it was taken from), which bench.py
reports as roofline.traffic / roofline.tensor_pipe_active_pct when the fingerprint matches the library it runs.
(sm__pipe_tensor_subpipe_hmma_cycles_active counts tcgen05 UTCHMMA work on sm_100 -- checked against tools/mma_rate.cu --
sm__inst_executed_pipe_tensor_subpipe_hmma does not.)"""
import collections
import csv
import json
import sys

sys.path.insert(0, ".")
from zeroshotvideoclassification_b200 import build

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_traffic.csv"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_conv_dram_traffic.json"
per_step = int(sys.argv[3]) if len(sys.argv) > 3 else 85      # igemm launches of one R(2+1)D-18 step (fprop 37 + dgrad 36 + parity classes)
with open(src) as f:
    rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
by_id = collections.OrderedDict()
for r in rows:
    d = by_id.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    d[r["Metric Name"]] = v * scale
launches = list(by_id.values())[-per_step:]
tot = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in launches)
us = sum(l.get("gpu__time_duration.sum", 0) for l in launches)
TP = "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"
tp = sum(l.get(TP, 0) * l.get("gpu__time_duration.sum", 0) for l in launches) / us if us else None
top = sorted(launches, key=lambda l: -l.get("gpu__time_duration.sum", 0))[:5]
out = {"kernel": "igemm_kmajor_kernel + igemm_halo_kernel (conv fprop + dgrad)", "launches": len(launches),
       "dram_bytes_per_launch": tot / max(1, len(launches)), "dram_bytes_per_step": tot,
       "kernel_us_per_step": us, "tensor_pipe_active_pct_time_weighted": tp,
       "top5_launches": [{"kernel": ("halo" if "halo" in l["name"] else "kmajor"), "us": l.get("gpu__time_duration.sum"),
                          "dram_mb": (l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0)) / 1e6,
                          "tensor_pipe_active_pct": l.get(TP)} for l in top],
       "build_fingerprint": build._fingerprint(),
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum," + TP +
                 " --clock-control none -k regex:igemm on `bench.py --quick --no-graph --steps 1 --warmup 1` (last step of "
                 "the capture; cold-cache, serialised launches)"}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out)[:600])
