"""A/B timing of libb2a build variants (tools/probes/_bin/libb2a_<name>.so): b2a_log_mel on 288 000 frames of s16 (the cfg2 clip
after trimming) and b2a_pipeline on a 10-minute 44.1 kHz stereo clip with ~20 % silence (gather path).  CUDA events, 20 runs."""
import ctypes as C, glob, os, sys
import numpy as np, torch
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE)); sys.path.insert(0, ROOT)
from audio_processor_b200 import _abi, synth
libs = sys.argv[1:] or sorted(glob.glob(os.path.join(HERE, "_bin", "libb2a_v*.so")))
n = 160 * 288000
g = torch.Generator(device="cuda").manual_seed(1)
x = (torch.randn(n, generator=g, device="cuda") * 3000).clamp(-32768, 32767).to(torch.int16)
clip = synth.synth_clip(2, 44100, 2, 3600.0, 0.20, device="cuda")
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for path in libs:
    lib = C.CDLL(path)
    for name, (res, args) in _abi.SIGNATURES.items():
        if hasattr(lib, name):
            fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
    T = n // 160
    out = torch.empty((80, T), dtype=torch.float32, device="cuda")
    wsb = lib.b2a_log_mel_workspace_bytes(1, n, 0)
    ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda")
    def lm():
        rc = lib.b2a_log_mel(C.c_void_p(x.data_ptr()), 0, 1, n, n, None, 0, 80, 0, C.c_void_p(out.data_ptr()), None, C.c_void_p(ws.data_ptr()), wsb, None)
        assert rc == 0, lib.b2a_last_error()
    t_lm = timeit(lm)
    n_in = int(clip.shape[0]); cap = 8192
    n16 = lib.b2a_resample_out_len(n_in, 44100, 16000)
    pcm = torch.empty(n16 + 64, dtype=torch.int16, device="cuda"); mel = torch.empty(80 * ((n16 + 16) // 160), dtype=torch.float32, device="cuda")
    ns = torch.zeros((cap, 2), dtype=torch.int32, device="cuda"); kp = torch.zeros((cap, 2), dtype=torch.int32, device="cuda"); info = torch.zeros(8, dtype=torch.int64, device="cuda")
    pwsb = lib.b2a_pipeline_workspace_bytes(n_in, 44100, 0, cap); pws = torch.empty(pwsb + 512, dtype=torch.uint8, device="cuda")
    pptr = C.c_void_p(pws.data_ptr() + (-pws.data_ptr()) % 256)
    prm = _abi.SilenceParams(1000, 200, 1, 0, -40.0)
    def pipe():
        rc = lib.b2a_pipeline(C.c_void_p(clip.data_ptr()), 0, 2, 44100, n_in, C.byref(prm), 80, 0, cap, C.c_void_p(pcm.data_ptr()), C.c_void_p(mel.data_ptr()),
                              C.c_void_p(ns.data_ptr()), C.c_void_p(kp.data_ptr()), C.c_void_p(info.data_ptr()), pptr, pwsb, None)
        assert rc == 0, lib.b2a_last_error()
    t_p = timeit(pipe)
    # cfg3-shaped: 512 x 30 s of f32, 128 mels (an eighth of cfg3)
    B3, n3 = 512, 480000
    if "x3" not in globals():
        globals()["x3"] = torch.randn((B3, n3), generator=g, device="cuda") * 0.1
    out3 = torch.empty((B3, 128, n3 // 160), dtype=torch.float32, device="cuda")
    wsb3 = lib.b2a_log_mel_workspace_bytes(B3, n3, 0)
    ws3 = torch.empty(wsb3 + 256, dtype=torch.uint8, device="cuda")
    def lm3():
        rc = lib.b2a_log_mel(C.c_void_p(x3.data_ptr()), 1, B3, n3, n3, None, 0, 128, 0, C.c_void_p(out3.data_ptr()), None, C.c_void_p(ws3.data_ptr()), wsb3, None)
        assert rc == 0, lib.b2a_last_error()
    t_3 = timeit(lm3, 10)
    print("%-28s log_mel(288k frames s16, 80) %7.1f us   pipeline(1 h clip) %7.1f us   log_mel(512 x 30 s f32, 128) %7.1f us" % (os.path.basename(path), t_lm, t_p, t_3), flush=True)
