"""Build libb2a.so (sm_100a) in-tree with nvcc.  No GPU needed: nvcc cross-compiles.

    python -m audio_processor_b200.build            # incremental
    python -m audio_processor_b200.build --force

The mel table header csrc/mel_tables_gen.inc is regenerated from tools/gen_mel_tables.cpp
(the same double-precision design code the runtime uses, csrc/mel_design.h).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libb2a.so")

SOURCES = [
    "b2a_host.cu", "logmel.cu", "silence.cu", "resample.cu", "pipeline.cu", "remap.cu", "fir_dispatch.cu",
    "fir_mma_44100.cu", "fir_mma_48000.cu", "fir_tmem_44100.cu", "fir_tmem_48000.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found: libb2a.so can only be built with the CUDA toolkit")
    return cand


def _deps_mtime() -> float:
    m = 0.0
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            if f.endswith((".cuh", ".h", ".inc")):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def gen_mel(force: bool = False) -> None:
    """csrc/mel_tables_gen.inc from tools/gen_mel_tables.cpp (+ csrc/mel_design.h, the design the runtime also uses)."""
    out = os.path.join(CSRC, "mel_tables_gen.inc")
    src = os.path.join(ROOT, "tools", "gen_mel_tables.cpp")
    dep = max(os.path.getmtime(src), os.path.getmtime(os.path.join(CSRC, "mel_design.h")))
    if not force and os.path.exists(out) and os.path.getmtime(out) >= dep:
        return
    os.makedirs(OBJ, exist_ok=True)
    exe = os.path.join(OBJ, "gen_mel_tables")
    subprocess.run(["g++", "-O2", "-o", exe, src], check=True)
    subprocess.run([exe, out], check=True)


def gen_mel_tc(force: bool = False) -> None:
    """tools/probes/_bin/mel_tc_tables_gen.inc (the filterbank as the tensor-core PROBE's epilogue unrolls it) from
    tools/probes/gen_mel_tc_tables.cpp: needed by profiling builds and the test-only emulation library, not by libb2a.so."""
    os.makedirs(os.path.join(ROOT, "tools", "probes", "_bin"), exist_ok=True)
    out = os.path.join(ROOT, "tools", "probes", "_bin", "mel_tc_tables_gen.inc")
    src = os.path.join(ROOT, "tools", "probes", "gen_mel_tc_tables.cpp")
    dep = max(os.path.getmtime(src), os.path.getmtime(os.path.join(CSRC, "mel_design.h")))
    if not force and os.path.exists(out) and os.path.getmtime(out) >= dep:
        return
    os.makedirs(OBJ, exist_ok=True)
    exe = os.path.join(OBJ, "gen_mel_tc_tables")
    subprocess.run(["g++", "-O2", "-o", exe, src], check=True)
    subprocess.run([exe, out], check=True)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    gen_mel(force)
    dep_m = _deps_mtime()
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), dep_m):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, *os.environ.get("B2A_NVCC_EXTRA", "").split(), "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]   # B2A_NVCC_EXTRA: experiment knobs (-D...)
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {os.path.basename(src)}:\n{r.stdout}\n{r.stderr}")
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            logs = list(ex.map(compile_one, jobs))
        if verbose:
            for (src, _), lg in zip(jobs, logs):
                print(f"== {os.path.basename(src)}\n{lg}")
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
