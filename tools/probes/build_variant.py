"""Build tools/probes/_bin/libb2a_v<name>.so from the product sources with extra -D flags:  build_variant.py <name> [-DX=1 ...]"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from audio_processor_b200 import build as B
name, extra = sys.argv[1], sys.argv[2:]
OBJ = os.path.join(HERE, "_bin", "var_obj_" + name); LIB = os.path.join(HERE, "_bin", "libb2a_v%s.so" % name)
os.makedirs(OBJ, exist_ok=True)
B.gen_mel(); B.gen_mel_tc()
def one(s):
    obj = os.path.join(OBJ, s[:-3] + ".o")
    r = subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *extra, "-I", os.path.join(ROOT, "include"), "-c", os.path.join(B.CSRC, s), "-o", obj], capture_output=True, text=True)
    if r.returncode: raise RuntimeError(r.stderr)
    return obj
with ThreadPoolExecutor(8) as ex: objs = list(ex.map(one, B.SOURCES))
subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart"], check=True)
print(LIB)
