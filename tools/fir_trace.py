"""pipeline timeline of the tcgen05 FIR (CTA 0): env B2A_FIR_TRACE hands the kernel a device buffer, it records clock64 at
converter (wait start / slot acquired / piece published), MMA-issuer (item start / waits done / issued) and epilogue
(accumulator full) events.  Profiling aid only."""
import os, sys, torch
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
buf = torch.zeros(2048 + 4 * 2048 + 1024, dtype=torch.int64, device="cuda")
os.environ["B2A_FIR_TRACE"] = str(buf.data_ptr())
from audio_processor_b200 import ops, synth
x = synth.synth_clip(2, 44100, 2, 3600.0, 0.2, device="cuda")
for _ in range(3): ops.resample(x, 44100, want_energy=True)
torch.cuda.synchronize()
t = buf.cpu().numpy()
c0 = t[:960].reshape(-1, 3); c1 = t[1024:1024 + 960].reshape(-1, 3)
n = int((c0[:, 2] > 0).sum())
t0 = c0[0, 0]
print("pieces traced", n)
for name, c in (("cvt warp 0", c0), ("cvt warp last", c1)):
    c = c[:n]
    wait = c[:, 1] - c[:, 0]; conv = c[:, 2] - c[:, 1]; per = np.diff(c[:, 2])
    print(f"{name}: PE wait mean {wait[9:].mean():.0f} cyc, convert+publish mean {conv[9:].mean():.0f}, piece period mean {per[9:].mean():.0f}")
    print("  first 30 pieces (wait, convert, t_publish):", [(int(a), int(b), int(cc - t0)) for a, b, cc in zip(wait[:30], conv[:30], c[:30, 2])])
for w in range(4):
    it = t[2048 + w * 2048: 2048 + w * 2048 + 680 * 3].reshape(-1, 3)
    m = int((it[:, 2] > 0).sum()); it = it[:m]
    print(f"issuer {w}: items {m}, wait mean {(it[:,1]-it[:,0])[10:].mean():.0f}, issue mean {(it[:,2]-it[:,1])[10:].mean():.0f}, item period {np.diff(it[:,0])[10:].mean():.0f}")
    if w == 0: print("  first 30 items (t_start, wait, issue):", [(int(a - t0), int(b - a), int(c - b)) for a, b, c in it[:30]])
e = t[2048 + 4 * 2048: 2048 + 4 * 2048 + 1000]; e = e[e > 0]
print("epilogue DF times (first 25):", [int(v - t0) for v in e[:25]], "block period mean", np.diff(e)[10:].mean())
