"""Top stall sites of one kernel from `ncu -i rep --page source --csv --kernel-id :::N` output (SASS view)."""
import csv, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
rows = rows[starts[which]:starts[which + 1]]
print(rows[0][1][:80])
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:top]
for i in sorted(order):
    r = data[i]
    st = {s[6:]: int(r[ix[s]] or 0) for s in stalls if int(r[ix[s]] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/tot:5.1f}% ex={r[ix['Instructions Executed']]:>8s} {r[ix['Source']].strip()[:70]:70s} {st}")
