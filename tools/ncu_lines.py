#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares from `ncu --page source --csv --print-source cuda`.
usage: python tools/ncu_lines.py export.csv <function-substring> <file-basename> [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
func, fname = sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
cur_f = cur_file = None; hdr = None; items = []
for r in rows:
    if not r: continue
    if r[0] == "Function Name": cur_f = r[1]; continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and cur_f and func in cur_f:
        try:
            inst = int(r[hdr.index("Instructions Executed")]); samp = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        if inst or samp:
            items.append((cur_file.split("/")[-1], int(r[0]) if r[0].isdigit() else -1, inst, samp, r[1].strip()[:90]))
ti = sum(i[2] for i in items); ts = sum(i[3] for i in items)
print("total warp-instr", ti, "samples", ts)
for i in sorted(items, key=lambda x: -x[3])[:top]:
    print(f"{i[0]:16s} L{i[1]:<4d} inst {100*i[2]/ti:5.1f}%  samples {100*i[3]/ts:5.1f}%  {i[4]}")
import json
json.dump(items, open("/tmp/ncu_lines.json", "w"))
