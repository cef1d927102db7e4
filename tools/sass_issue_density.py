"""Instruction density of the MMA-issue code of each tcgen05 kernel in libzsv_b200.so: for every kernel, the SASS
instructions between consecutive UTCHMMA instructions (cuobjdump -sass, no GPU needed).  The issuing thread of the small-N
convolutions is a scalar bottleneck (profiles/r02_issue_loop.txt), so this is the figure to watch when touching those loops.
usage: python tools/sass_issue_density.py [kernel-name-substring]"""
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parents[1] / "zeroshotvideoclassification_b200" / "libzsv_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
want = sys.argv[1] if len(sys.argv) > 1 else ""
name, n, gaps, r2ur = None, 0, [], 0
res = {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name and gaps:
            res[name] = (gaps, r2ur)
        name, n, gaps, r2ur = m.group(1), 0, [], 0
        continue
    if not re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        continue
    n += 1
    if "UTCHMMA" in line:
        gaps.append(n)
        n = 0
if name and gaps:
    res[name] = (gaps, r2ur)
for k, (g, _) in res.items():
    if want in k:
        short = re.sub(r"^_ZN3zsv\d+_GLOBAL__N__[0-9a-f_]+\d+", "", k)[:60]
        print(f"{short:60s} {len(g):3d} UTCHMMA, instructions before each: {g[1:]}")
