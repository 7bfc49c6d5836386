"""time b2a_resample on the cfg2 clip with FIR phases masked (profiling aid; set B2A_FIR_PHASES before import)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_processor_b200 import ops, synth
x = synth.synth_clip(2, 44100, 2, 3600.0, 0.2, device="cuda")
for _ in range(5): ops.resample(x, 44100, want_energy=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(30): ops.resample(x, 44100, want_energy=True)
b.record(); torch.cuda.synchronize()
print("B2A_FIR_IMPL=%s B2A_FIR_PHASES=%s  resample %.1f us" % (os.environ.get("B2A_FIR_IMPL", "tmem"), os.environ.get("B2A_FIR_PHASES", "all"), a.elapsed_time(b) / 30 * 1e3))
