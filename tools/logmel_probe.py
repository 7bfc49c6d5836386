"""time b2a_log_mel on cfg2-sized trimmed s16 PCM and on a cfg3 slice (profiling aid)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_processor_b200 import ops, synth
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
x = synth.synth_clip(2, 16000, 1, 2900.0, 0.0, device="cuda")            # ~46.4 M samples, like cfg2 after trimming
print("log-mel s16 %d samples 80 mel: %.1f us" % (x.numel(), timeit(lambda: ops.log_mel(x, 80))))
y = synth.noise_batch(3, 512, 480000, device="cuda")
print("log-mel f32 [512,480000] 128 mel: %.1f us" % timeit(lambda: ops.log_mel(y, 128), 10))
