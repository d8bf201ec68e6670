#!/usr/bin/env python
"""Warp-state samples per SASS instruction of a --import-source ncu capture: where warps WAIT.
   python tools/ncu_stalls.py gpurun_out/x.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if "Source" in r and any("Samples" in c for c in r))
isrc = hdr.index("Source")
isamp = next(i for i, c in enumerate(hdr) if c.strip() == "# Samples")
stall_cols = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") or c.startswith("Stall")]
data = []
for n, r in enumerate(rows):
    if len(r) != len(hdr) or r is hdr:
        continue
    try:
        s = int(r[isamp])
    except ValueError:
        continue
    data.append((s, n, r))
tot = sum(s for s, _, _ in data)
print(f"total samples {tot}; columns: {[c for _, c in stall_cols][:30]}")
for s, n, r in sorted(data, reverse=True)[:top]:
    why = sorted(((int(r[i]) if r[i].isdigit() else 0, c) for i, c in stall_cols), reverse=True)[:3]
    print(f"{100.0 * s / tot:5.1f}%  {r[isrc][:70]:70s} " + " ".join(f"{c.replace('stall_', '')}:{v}" for v, c in why if v))
