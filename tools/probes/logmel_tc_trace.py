"""Event timeline of logmel_tc_kernel's CTA 0 (profiling build, tools/probes/build_profile_lib.py).
Prints, per tile, when each phase of worker warps 0 / 5 / 15 and of the issuer started, in cycles from the kernel's first event."""
import ctypes as C, os, sys
import numpy as np, torch
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE)); sys.path.insert(0, ROOT)
from audio_processor_b200 import _abi
lib = C.CDLL(os.path.join(HERE, "_bin", "libb2a_prof.so"))
for name, (res, args) in _abi.SIGNATURES.items():
    fn = getattr(lib, name); fn.restype = res; fn.argtypes = args
lib.b2a_debug_tc_trace.argtypes = [C.c_void_p, C.c_int, C.c_int]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 160 * 128 * 148 * 8          # 8 tiles per CTA
g = torch.Generator(device="cuda").manual_seed(1)
x = (torch.randn(n, generator=g, device="cuda") * 3000).clamp(-32768, 32767).to(torch.int16)
T = n // 160
out = torch.empty((80, T), dtype=torch.float32, device="cuda")
wsb = lib.b2a_log_mel_workspace_bytes(1, n, 0)
ws = torch.empty(wsb + 256, dtype=torch.uint8, device="cuda")
def run():
    rc = lib.b2a_log_mel(C.c_void_p(x.data_ptr()), 0, 1, n, n, None, 0, 80, 0, C.c_void_p(out.data_ptr()), None, C.c_void_p(ws.data_ptr()), wsb, None)
    assert rc == 0, lib.b2a_last_error()
for _ in range(3): run()
torch.cuda.synchronize()
lib.b2a_debug_tc_trace(None, 0, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); run(); b.record(); torch.cuda.synchronize()
print("log_mel call: %.1f us for %d frames (%d tiles)" % (a.elapsed_time(b) * 1e3, T, (T + 127) // 128))
buf = np.zeros(16384, dtype=np.uint64)
lib.b2a_debug_tc_trace(buf.ctypes.data_as(C.c_void_p), 16384, 0)
cnt = int(buf[0]); rec = buf[1:min(cnt, 16383) + 1]
ev = (rec >> np.uint64(56)).astype(int); wp = ((rec >> np.uint64(48)) & np.uint64(0xff)).astype(int); clk = (rec & np.uint64(0xffffffffffff)).astype(np.int64)
t0 = clk.min()
names = {1: "fetch", 2: "fetch_end", 3: "rawfull", 4: "kstep", 5: "store_wait", 6: "store", 7: "store_end", 8: "accfull", 12: "epi_end", 9: "iss_full0", 10: "iss_tile", 11: "iss_commit"}
order = np.argsort(clk, kind="stable")
print("records", cnt)
for w in (0, 5, 15, 16):
    print("---- warp", w)
    line = []
    for i in order:
        if wp[i] != w: continue
        line.append("%s@%d" % (names.get(ev[i], ev[i]), clk[i] - t0))
        if ev[i] in (12, 11): print("  " + " ".join(line)); line = []
        if len(line) > 40: print("  " + " ".join(line)); line = []
    if line: print("  " + " ".join(line))
