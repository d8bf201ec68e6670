#!/usr/bin/env python
"""List the loops (backward branches) of one kernel in a .so with their instruction mix:
   python tools/sass_loops.py hackrfdiags_b200/libhrd_b200.so rx_wbfm_kernelILi0E"""
import re
import subprocess
import sys

so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
body = next(b for b in blocks if pat in b.split("\n", 1)[0])
ins = []
for l in body.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
idx = {a: i for i, (a, _) in enumerate(ins)}
print(f"{len(ins)} instructions")
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in idx:
            ops = {}
            for _, x in ins[idx[tgt]:i + 1]:
                x = re.sub(r"^@!?U?P\d\s+", "", x)
                op = x.split()[0].split(".")[0]
                ops[op] = ops.get(op, 0) + 1
            top = sorted(ops.items(), key=lambda kv: -kv[1])[:14]
            print(f"loop {tgt:#06x}..{a:#06x}: {i + 1 - idx[tgt]:5d} instr  " + " ".join(f"{k}:{v}" for k, v in top))
