import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from audio_processor_b200 import ops, synth
x = synth.synth_clip(2, 44100, 2, 3600.0, 0.2, device="cuda")
plan = ops.PipelinePlan(x.shape[0], 44100, 2, x.dtype, n_mels=80)
r = plan.run(x, min_silence_len=1000, silence_thresh=-40, keep_silence=200)
mel = r.mel
mx, mn = float(mel.max()), float(mel.min())
print("mel max %.4f min %.4f floor %.4f frac_at_floor %.5f" % (mx, mn, mx - 2.0, float((mel <= mx - 2.0 + 1e-6).float().mean())))
per_frame_min = mel.min(0).values
T = mel.shape[1]; tiles = (T + 31) // 32
pm = torch.nn.functional.pad(per_frame_min, (0, tiles * 32 - T), value=9.0).view(tiles, 32).min(1).values
print("tiles", tiles, "tiles needing clamp", int((pm <= mx - 2.0 + 1e-6).sum()))
