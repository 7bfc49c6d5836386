#!/usr/bin/env python
"""Per-region instruction and stall-sample shares of one kernel from an .ncu-rep source page (SASS view).
usage: python tools/ncu_regions.py report.ncu-rep kernel-regex [top-N instructions]
Regions are runs of SASS instructions with the same execution count (loops and phases separate naturally)."""
import csv, io, subprocess, sys
from collections import Counter
rep, kre = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) >= len(h) - 2 and r[0].startswith("0x")]
first = data[0][0]
for i in range(1, len(data)):                      # some reports list the function twice
    if data[i][0] == first:
        data = data[:i]
        break
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_s = sum(int(r[ix["# Samples"]]) for r in data)
print(f"{len(data)} SASS instructions, {tot_i} warp instructions executed, {tot_s} samples")
agg = Counter()
for r in data:
    for s in stalls:
        agg[s] += int(r[ix[s]])
print("stall samples: " + ", ".join(f"{k[6:]} {100 * v / tot_s:.1f}%" for k, v in agg.most_common(9)))
segs, cur = [], None
for i, r in enumerate(data):
    ie = int(r[ix["Instructions Executed"]])
    t = r[ix["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    if cur is None or abs(ie - cur["ie"]) > 0.15 * max(cur["ie"], 1000):
        cur = {"ie": ie, "start": i, "n": 0, "sum": 0, "samp": 0, "ops": Counter(), "st": Counter()}
        segs.append(cur)
    cur["n"] += 1; cur["sum"] += ie; cur["samp"] += int(r[ix["# Samples"]]); cur["ops"][op] += ie
    for s in stalls:
        cur["st"][s[6:]] += int(r[ix[s]])
for s in segs:
    if s["sum"] > tot_i * 0.004 or s["samp"] > tot_s * 0.01:
        ops = ", ".join(f"{k}:{v * 100 // max(s['sum'], 1)}" for k, v in s["ops"].most_common(6))
        st = ", ".join(f"{k}:{v * 100 // max(s['samp'], 1)}" for k, v in s["st"].most_common(3))
        print(f"@{s['start']:5d} n={s['n']:4d} exec={s['ie']:7d} instr={100 * s['sum'] / tot_i:5.1f}% samples={100 * s['samp'] / tot_s:5.1f}% | {ops} | {st}")
print()
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
    st = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{data.index(r):5d} {100 * int(r[ix['# Samples']]) / tot_s:5.2f}% exec={r[ix['Instructions Executed']]:>8s} {r[ix['Source']][:58]:58s} {st}")
