#!/usr/bin/env python
"""Join an ncu capture's per-SASS-instruction counters with nvdisasm's line table (same build) and aggregate per
source line: warp instructions executed, thread instructions, stall samples.

usage: python scripts/ncu_by_line.py <report.ncu-rep> <object.o> <kernel-name-substring> [top_n]
"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os
from collections import defaultdict

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instructions of the kernel in order, each with the innermost (last seen) file:line
lines, cur, on = [], ("?", 0), False
for l in dis:
    if l.startswith(".text."):
        on = kname in l
        continue
    if not on:
        continue
    ms = re.match(r"^(\$__internal_\d+_\$)?(__cuda_\w+):", l)
    if ms:
        cur = ("subroutine " + ms.group(2), 0)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if "inlined at" not in l:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout)))
hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
h = src[hi]
ix = {k: h.index(k) for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "Source")}
rows = src[hi + 1:]
if len(rows) != len(lines):
    print(f"warning: {len(rows)} instructions in the report vs {len(lines)} in the object (different build?)", file=sys.stderr)
agg = defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0]
for r, loc in zip(rows, lines):
    try:
        s, e, t = int(float(r[ix["# Samples"]])), int(float(r[ix["Instructions Executed"]])), int(float(r[ix["Thread Instructions Executed"]]))
    except ValueError:
        continue
    a = agg[loc]
    a[0] += s; a[1] += e; a[2] += t; a[3] += 1
    tot[0] += s; tot[1] += e; tot[2] += t
print(f"# {rep}: kernel *{kname}*: {tot[1]} warp instructions, {tot[2]} thread instructions, {tot[0]} stall samples")
print(f"{'file:line':28s} {'sass':>5s} {'warp instr':>13s} {'%':>6s} {'thread instr':>14s} {'lanes':>6s} {'samples':>9s} {'%':>6s}")
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{loc[0] + ':' + str(loc[1]):28s} {a[3]:5d} {a[1]:13d} {100.0 * a[1] / max(tot[1], 1):6.2f} {a[2]:14d} {a[2] / max(a[1], 1):6.1f} {a[0]:9d} {100.0 * a[0] / max(tot[0], 1):6.2f}")
