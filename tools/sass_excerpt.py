#!/usr/bin/env python3
"""Per-kernel SASS evidence for profiles/: instruction totals and the counts of the Blackwell-specific mnemonics
(tcgen05 MMA = UTC*MMA, TMEM loads/stores = LDTM/STTM, TMA = UTMALDG/UBLKCP, mbarrier = SYNCS, ...), plus the first
lines around each tensor-core / TMA instruction.  Usage: python tools/sass_excerpt.py [lib.so] > profiles/rNN_sass_excerpt.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "audio_processor_b200/libb2a.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF",
        "SYNCS", "HMMA", "IMMA", "LDSM", "LDGSTS", "REDUX", "MUFU", "SHFL", "BAR", "LDG", "STG", "LDS", "STS", "ATOM", "RED", "CCTL", "NANOSLEEP"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
funcs, cur = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = funcs.setdefault(m.group(1), [])
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur is not None:
        cur.append((m.group(1), m.group(2).strip()))
print(f"# SASS summary of {LIB} (cuobjdump -sass, sm_100a)\n")
print(f"{'instr':>6}  kernel   [mnemonic counts]")
for name, ins in sorted(funcs.items(), key=lambda kv: -len(kv[1])):
    cnt = collections.Counter()
    for _, text in ins:
        op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0].split(".")[0]
        for k in KEYS:
            if op == k or (k in ("BAR",) and op == "BAR") or (k.startswith("UTC") and op.startswith(k)):
                cnt[k] += 1
                break
    short = re.sub(r"\(.*", "", demangle(name))
    print(f"{len(ins):>6}  {short}")
    print("        " + ", ".join(f"{k} {cnt[k]}" for k in KEYS if cnt[k]))
print("\n# excerpts: every tcgen05 MMA / TMA / TMEM instruction form that appears, once per kernel\n")
for name, ins in funcs.items():
    seen, lines = set(), []
    for addr, text in ins:
        op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0]
        if re.match(r"(UTC.*MMA|UTCBAR|LDTM|STTM|UTMALDG|UBLKCP|UTCATOMSWS|HMMA|LDSM|SYNCS)", op) and op not in seen:
            seen.add(op)
            lines.append(f"    /*{addr}*/  {text}")
    if lines:
        print(re.sub(r"\(.*", "", demangle(name)))
        print("\n".join(lines))
