import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_json():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


def golden_arrays():
    return np.load(os.path.join(GOLDEN_DIR, "golden_arrays.npz"))


def tone_pair(rate, n):
    t = np.arange(n) / float(rate)
    L = np.round(8000 * np.sin(2 * np.pi * 1000 * t)).astype(np.int16)
    R = np.round(8000 * np.sin(2 * np.pi * 3000 * t)).astype(np.int16)
    return np.stack([L, R], 1)


def piecewise(parts, extra=0):
    a = [np.full(int(round(d * 16000)), amp, dtype=np.int16) for d, amp in parts]
    if extra:
        a.append(np.zeros(extra, dtype=np.int16))
    return np.concatenate(a)


def random_speechlike(rng, n, lo_amp=20, hi_amp=6000, min_span=1500, max_span=40000):
    """piecewise-amplitude noise with stray lengths: exercises every silence code path"""
    out = np.zeros(n, dtype=np.int16)
    pos = 0
    loud = bool(rng.integers(0, 2))
    while pos < n:
        span = int(rng.integers(min_span, max_span))
        amp = hi_amp if loud else lo_amp
        kind = rng.integers(0, 3)
        seg_len = min(span, n - pos)
        if kind == 0:
            seg = np.full(seg_len, amp if loud else int(rng.integers(0, lo_amp + 1)))
        else:
            seg = rng.standard_normal(seg_len) * amp
        out[pos:pos + seg_len] = np.clip(np.round(seg), -32768, 32767).astype(np.int16)
        pos += span
        loud = not loud
    return out


def energy_oracle(y16, spm=16):
    n = len(y16)
    ne = (n + spm - 1) // spm
    yy = np.zeros(ne * spm, np.int64)
    yy[:n] = y16
    return (yy * yy).reshape(ne, spm).sum(1)
