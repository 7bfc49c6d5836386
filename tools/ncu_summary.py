#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / profiles/ quote.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
PAT = re.compile(r"^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
                 r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|launch__(registers_per_thread|grid_size|block_size|shared_mem_per_block_dynamic|occupancy_limit_\w+)|"
                 r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|"
                 r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|sm__inst_executed_pipe_(fma|alu|lsu|xu|tensor\w*|uniform)\.avg\.pct_of_peak_sustained_active|"
                 r"sm__pipe_tensor\w*cycles_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed_pipe_tensor\w*\.sum|"
                 r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?|"
                 r"lts__t_sector_hit_rate\.pct|sm__icc_request_hit_rate\.pct|lts__t_bytes\.sum|sm__cycles_active\.avg)$")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if filt and filt not in d.get("Kernel Name", ""):
        continue
    print("----")
    for k, u in zip(hdr, units):
        if PAT.match(k):
            v = d[k]
            if k.startswith("smsp__average_warps_issue_stalled") and float(v or 0) < 0.05:
                continue
            print(f"  {k} = {v[:90]} {u}")
