#!/usr/bin/env python
"""Per CUDA source line: warp instructions executed and stall samples, from an .ncu-rep with imported source
     python tools/ncu_lines.py rep.ncu-rep <kernel regex> <instance> [top]"""
import csv
import io
import subprocess
import sys

rep, rx, inst = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-id",
                      f"::regex:{rx}:{inst}"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, data = "", None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 8 and r[0] == "Line No":
        hdr = r
        ii, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    elif hdr and len(r) > ii and r[0] not in ("", "Line No"):
        try:
            data.append((int(r[ii]), int(r[isamp]), fname, r[0], r[1].strip()))
        except ValueError:
            pass
tot, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print("warp instructions", tot, " stall samples", ts)
print("-- by instructions")
for d in sorted(data, key=lambda x: -x[0])[:top]:
    print(f"{d[0] / tot * 100:5.1f}% instr {d[1] / max(ts, 1) * 100:5.1f}% samp  {d[2]}:{d[3]:>5}  {d[4][:120]}")
print("-- by stall samples")
for d in sorted(data, key=lambda x: -x[1])[:top]:
    print(f"{d[0] / tot * 100:5.1f}% instr {d[1] / max(ts, 1) * 100:5.1f}% samp  {d[2]}:{d[3]:>5}  {d[4][:120]}")
