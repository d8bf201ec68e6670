#!/usr/bin/env python
"""Per-source-line executed warp instructions of a --import-source ncu capture (top lines):
   python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr, tot, lines = None, None, 0, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit() and "Instructions Executed" in hdr:
        try:
            n = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        d = {"# Samples": r[hdr.index("# Samples")]}
        if n:
            lines.append((n, cur, int(r[0]), r[1].strip()[:110], d.get("# Samples", "")))
            tot += n
lines.sort(reverse=True)
print(f"total executed warp instructions attributed to source lines: {tot}")
for n, f, ln, src, smp in lines[:top]:
    print(f"{100.0 * n / tot:5.1f}%  {n:12d}  {f}:{ln:<5d} {src}")
