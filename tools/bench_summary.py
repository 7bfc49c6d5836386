#!/usr/bin/env python
"""print the headline fields of bench JSON lines: python tools/bench_summary.py gpurun_out/bench_*_TAG.json"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d.get("roofline", {})
    print(f"{f}: value={d['value']:.1f} {d['unit']} ms/step={d['ms_per_step']:.3f} launches={d.get('gpu_launches')}")
    print("   stages_ms", d.get("stages_ms"))
    if r:
        print(f"   roofline {r.get('kernel')}: {r['achieved']:.0f} GB/s frac={r['frac']:.3f} kernel_ms={r.get('kernel_ms'):.3f} | pipeline frac={r.get('pipeline_frac'):.3f}")
    if d.get("e2e"):
        e = d["e2e"]; print(f"   e2e {e['value']:.1f} ms/step={e.get('ms_per_step')}, h2d={e['h2d_bytes_per_step']/1e6:.0f}MB d2h={e['d2h_bytes_per_step']/1e6:.0f}MB")
    if d.get("cpu_baseline"):
        c = d["cpu_baseline"]; print(f"   cpu {c['value']:.4f} cores={c['cores']} :: {c['sample'][:150]}")
    print("   clocks", d.get("clocks"))
