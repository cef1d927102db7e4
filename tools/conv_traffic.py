"""Mean DRAM traffic per launch of the implicit-GEMM conv kernels (fprop + dgrad) of one training step, from
`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:igemm --csv` output.
Writes profiles/r01_conv_dram_traffic.json, which bench.py reports as roofline.traffic."""
import collections
import csv
import json
import sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/conv_traffic.csv"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/r01_conv_dram_traffic.json"
per_step = int(sys.argv[3]) if len(sys.argv) > 3 else 85      # igemm launches of one R(2+1)D-18 step (71 + 14)
with open(src) as f:
    rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
by_id = collections.OrderedDict()
for r in rows:
    d = by_id.setdefault(r["ID"], {})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    d[r["Metric Name"]] = v * scale
launches = list(by_id.values())[-per_step:]
tot = sum(l.get("dram__bytes_read.sum", 0) + l.get("dram__bytes_write.sum", 0) for l in launches)
out = {"kernel": "igemm_kmajor_kernel + igemm_halo_kernel (conv fprop + dgrad)", "launches": len(launches),
       "dram_bytes_per_launch": tot / max(1, len(launches)), "dram_bytes_per_step": tot,
       "kernel_us_per_step": sum(l.get("gpu__time_duration.sum", 0) for l in launches),
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                 "-k regex:igemm on `bench.py --quick --no-graph --steps 1 --warmup 1` (last step of the capture)"}
json.dump(out, open(dst, "w"), indent=1)
print(out)
