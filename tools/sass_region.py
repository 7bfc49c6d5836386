#!/usr/bin/env python
"""Opcode histogram of the SASS between two markers of one kernel: python tools/sass_region.py lib.so mangled-name [first-op last-op]"""
import subprocess, sys, re
from collections import Counter
lib, fn = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", "-fun", fn, lib], capture_output=True, text=True).stdout
L = [m.group(1).strip() for m in (re.match(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", l) for l in out.splitlines()) if m]
print(len(L), "instructions")
def op(l):
    t = l.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
if len(sys.argv) > 3:
    idx = [i for i, l in enumerate(L) if sys.argv[3] in l]
    blocks, cur = [], [idx[0]]
    for i in idx[1:]:
        if i - cur[-1] > 120: blocks.append(cur); cur = [i]
        else: cur.append(i)
    blocks.append(cur)
    for b in blocks:
        s = max((i for i in range(b[0]) if L[i].startswith("BAR") or "SYNCS" in L[i]), default=0)
        e = b[-1] + 30
        print(f"region {s}..{e} ({e - s} instr, {len(b)} markers):", Counter(op(l) for l in L[s:e]).most_common(14))
else:
    print(Counter(op(l) for l in L).most_common(25))
