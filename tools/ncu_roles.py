"""Where the warps of a warp-specialised kernel spend their samples: reads the SASS page of an ncu report
(--set full --import-source on) and prints, per barrier-wait / TMA / MMA / TMEM-load instruction, the stall samples, plus the
hottest lines.  usage: python tools/ncu_roles.py report.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    for key in ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"):
        if key in hdr:
            print(f"{key:90s} {r[hdr.index(key)]} {rows[1][hdr.index(key)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]
i_s, i_src, i_ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stalls = [(x, hdr.index(x)) for x in hdr if x.startswith("stall_") and "Not Issued" not in x]
data = []
for n, r in enumerate(rows[h + 1:]):
    try:
        data.append((n, int(r[i_s]), int(r[i_ex]), r[i_src], r))
    except (ValueError, IndexError):
        pass
tot = sum(d[1] for d in data)
print(f"total samples {tot}")
print("--- hottest lines")
for n, s, ex, srcl, r in sorted(data, key=lambda d: -d[1])[:top]:
    st = sorted([(int(r[i] or 0), x) for x, i in stalls], reverse=True)[:2]
    print(f"{n:5d} {s:6d} {100.0 * s / tot:5.1f}% ex={ex:9d} {srcl[:70]:70s} {st}")
print("--- sync / TMA / MMA instructions")
for n, s, ex, srcl, r in data:
    if any(k in srcl for k in ("SYNCS.PHASECHK", "UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "UTMASTG", "UCGABAR_WAIT")) and ex > 0:
        print(f"{n:5d} {s:6d} ex={ex:9d} {srcl[:90]}")

# samples between consecutive named barriers / role boundaries (phases of the epilogue)
marks = [(n, srcl.strip()[:44]) for n, s_, ex, srcl, r in data if ("BAR.SYNC" in srcl or "UCGABAR_WAIT" in srcl) and ex > 0]
print("--- samples between barrier instructions (line ranges)")
prev = 0
for n, lab in marks + [(len(data), "end")]:
    tot_s = sum(d[1] for d in data if prev <= d[0] < n + 2)
    tot_x = sum(d[2] for d in data if prev <= d[0] < n + 2)
    print(f"{prev:5d}-{n:5d} samples {tot_s:6d}  instr {tot_x:10d}   up to: {lab}")
    prev = n + 2
