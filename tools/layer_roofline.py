"""Per-convolution roofline from a `bench.py --layer-table` file: for every (pass, layer) the measured time against
the tensor floor (algorithmic FLOPs / sustained bf16 peak) and the HBM floor (algorithmic bytes / copy bandwidth).

Algorithmic bytes (bf16 activations with channels padded to 8, each tensor touched once):
  fprop: read x, write y            dgrad: read dy, write dx (+ read y_prev when the BatchNorm reduction is fused)
  wgrad: read x, read dy            weights are negligible against the activations on every layer but layer 4.
usage: python tools/layer_roofline.py profiles/r01_layer_table_v9.txt [batch]"""
import json
import re
import sys

path = sys.argv[1]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 22
try:
    peaks = json.load(open("MEASURED_PEAKS.json"))
    tf, bw, src = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"], "MEASURED_PEAKS.json"
except OSError:
    tf, bw, src = 1401.9, 6542.7, "B200_PROFILING.md fallback"
cpad = lambda c: (c + 7) // 8 * 8
pat = re.compile(r"(\w+)\s+(\d+)->\s*(\d+) \((\d+), (\d+), (\d+)\) \((\d+), (\d+), (\d+)\) (\d+)x(\d+)\s*:\s+(\d+)\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)")
rows = []
for line in open(path):
    m = pat.match(line)
    if not m:
        continue
    kind = m.group(1)
    cin, cout, kt, kh, kw, st, sh, sw, T, H, calls = (int(m.group(i)) for i in (2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12))
    us, tflops = float(m.group(13)), float(m.group(14))
    W = H
    To, Ho, Wo = -(-T // st), -(-H // sh), -(-W // sw)          # "same"-style padding on this network
    pin, pout = batch * T * H * W, batch * To * Ho * Wo
    flops = 2.0 * pout * cout * cin * kt * kh * kw
    xin = 16 if (cin == 3) else cpad(cin)                        # the stem reads the W-folded image (8 dw x 8 c per 4 px)
    bx, by = pin * xin * 2, pout * cpad(cout) * 2
    wbytes = cout * cin * kt * kh * kw * (2 if kind != "wgrad" else 4)
    nbytes = bx + by + wbytes
    t_tensor, t_hbm = flops / (tf * 1e12) * 1e6, nbytes / (bw * 1e9) * 1e6
    floor = max(t_tensor, t_hbm)
    rows.append((calls * us, kind, f"{cin}->{cout}", f"{kt}x{kh}x{kw}/{st}{sh}{sw}", f"{T}x{H}", calls, us, t_tensor, t_hbm,
                 "tensor" if t_tensor >= t_hbm else "hbm", floor / us))
rows.sort(reverse=True)
print(f"peaks: {tf} TFLOP/s sustained bf16, {bw} GB/s copy ({src}); batch {batch}")
print("pass   layer        kernel/stride  TxH    calls  us/call  tensor-floor  hbm-floor  bound   floor/measured")
tot_meas = tot_floor = 0.0
for tot, kind, lay, ks, th, calls, us, tt, thb, bound, frac in rows:
    print(f"{kind:6s} {lay:12s} {ks:14s} {th:6s} {calls:4d} {us:9.1f} {tt:11.1f} {thb:10.1f}  {bound:6s} {frac:8.2f}")
    tot_meas += calls * us
    tot_floor += calls * max(tt, thb)
print(f"all convolutions: measured {tot_meas / 1e3:.2f} ms per step, sum of floors {tot_floor / 1e3:.2f} ms, ratio {tot_floor / tot_meas:.2f}")
