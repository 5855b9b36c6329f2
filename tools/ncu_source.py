#!/usr/bin/env python
"""Per-source-line totals from an .ncu-rep source page (needs -lineinfo and --import-source on):
     python tools/ncu_source.py rep.ncu-rep <kernel regex> <instance> [top]
Prints warp instructions executed and stall samples aggregated by SASS opcode class and the hottest SASS addresses."""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, rx, inst = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{rx}:{inst}"],
                     stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ia, isrc, iinst, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ops, samp = Counter(), Counter()
total = 0
lines = []
seen = set()
for r in rows[2:]:
    if len(r) <= iinst:
        continue
    if r[ia] in seen:             # the page lists every instruction twice (two views): count each address once
        continue
    seen.add(r[ia])
    try:
        n = int(r[iinst]); s = int(r[isamp])
    except ValueError:
        continue
    src = r[isrc].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:2]) if op.startswith(("LD", "ST")) else "")
    ops[op] += n
    samp[op] += s
    total += n
    lines.append((s, n, src))
print(f"total warp instructions: {total}")
for op, n in ops.most_common(top):
    print(f"{op:14s} {n:12d} {100.0 * n / total:6.2f}%   samples {samp[op]}")
print("-- hottest by stall samples")
for s, n, src in sorted(lines, reverse=True)[:top]:
    print(f"{s:7d} {n:10d}  {src}")
