"""Build tests/emu/libb2a_emu.so: the kernel sources compiled by g++ against the fiber
emulation in cuda_emu.h.  TEST-ONLY — lets `-m "not gpu"` tests exercise kernel index math,
barrier structure and integer exactness on a box without a GPU.  The product never loads it.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "audio_processor_b200", "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libb2a_emu.so")

sys.path.insert(0, ROOT)
from audio_processor_b200.build import SOURCES, gen_mel, gen_mel_tc  # noqa: E402

FLAGS = ["-O1", "-std=c++17", "-fPIC", "-DB2A_EMU", "-I", HERE, "-I", os.path.join(ROOT, "include"),
         "-include", os.path.join(HERE, "cuda_emu.h"), "-x", "c++", "-mfma", "-mavx2", "-ffp-contract=fast",
         "-Wno-unused-function", "-Wno-attributes"]


def build(force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    gen_mel()
    gen_mel_tc()
    dep_m = 0.0
    for d in (CSRC, HERE, os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "probes")):       # probes: the tensor-core log-mel variant
        for f in os.listdir(d):
            if f.endswith((".cuh", ".h", ".inc")):
                dep_m = max(dep_m, os.path.getmtime(os.path.join(d, f)))
    jobs = []
    srcs = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(HERE, "cuda_emu.cpp")]
    for src in srcs:
        obj = os.path.join(OBJ, os.path.basename(src).rsplit(".", 1)[0] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), dep_m):
            jobs.append((src, obj))

    def one(job):
        src, obj = job
        r = subprocess.run(["g++", *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ (emu) failed for {os.path.basename(src)}:\n{r.stderr[-6000:]}")

    if jobs:
        with ThreadPoolExecutor(max_workers=8) as ex:
            list(ex.map(one, jobs))
    objs = [os.path.join(OBJ, os.path.basename(s).rsplit(".", 1)[0] + ".o") for s in srcs]
    if jobs or not os.path.exists(LIB):
        r = subprocess.run(["g++", "-shared", "-o", LIB, *objs], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"emu link failed:\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
