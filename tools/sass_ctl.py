#!/usr/bin/env python
"""Decode the scheduling control bits (stall, yield, write/read scoreboard, wait mask) of a kernel's SASS:
   python tools/sass_ctl.py hackrfdiags_b200/libhrd_b200.so 'rx_kernelILi1ELi0' [filter-regex]
Shows global loads and every instruction that waits on a scoreboard a global load writes -- i.e. whether
software-prefetched LDGs really stay in flight (two LDGs on ONE counting scoreboard do not)."""
import re
import subprocess
import sys

so, pat = sys.argv[1], sys.argv[2]
flt = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
body = next(b for b in blocks if pat in b.split("\n", 1)[0]).splitlines()
ins = []
i = 0
while i < len(body) - 1:
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/", body[i])
    m2 = re.match(r"\s*/\* (0x[0-9a-f]{16}) \*/", body[i + 1])
    if m and m2:
        ctl = int(m2.group(1), 16) >> 41
        ins.append((int(m.group(1), 16), m.group(2).strip(), ctl & 0xf, (ctl >> 4) & 1, (ctl >> 5) & 7, (ctl >> 8) & 7, (ctl >> 11) & 0x3f))
        i += 2
    else:
        i += 1
ldg_sb = {w for _, t, _, _, w, _, _ in ins if t.split()[-0].startswith(("LDG", "@")) and "LDG" in t and w != 7}
print(f"{len(ins)} instructions; global loads write scoreboards {sorted(ldg_sb)}")
for a, t, s, y, w, r, wm in ins:
    waits_ldg = any(wm >> b & 1 for b in ldg_sb)
    if ("LDG" in t or waits_ldg) and (not flt or flt.search(t)):
        print(f"{a:05x} stall={s:2d} wbar={w if w != 7 else '-'} rbar={r if r != 7 else '-'} wait={wm:06b}  {t[:90]}")
