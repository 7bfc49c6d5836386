"""float64 restatement of FFmpeg libswresample's default resampler + downmix.

TEST INFRASTRUCTURE (see oracle/__init__.py).

What it restates: the arithmetic reached from the reference's
``ffmpeg -y -i IN -ar 16000 -ac 1 -c:a pcm_s16le OUT``
(/root/reference/app/services/audio_processor.py:912-923).  The algorithm lives in
FFmpeg's libswresample (not vendored in /root/reference; the Dockerfile installs an
unpinned apt ffmpeg, /root/reference/Dockerfile:9).  Pinned version here: FFmpeg 8.0.1,
libswresample 6.1.100 (oracle/swr_ref.py drives the real library).  Published
algorithm restated below (libswresample/resample.c: build_filter, swri_resample;
libswresample/rematrix.c for the 0.5/0.5 stereo→mono matrix; audioconvert.c for
s16<->float scaling and lrintf quantisation):

  defaults  filter_size=32, phase_shift=10, linear_interp=1, exact_rational=1,
            cutoff=0.97 (swr engine), Kaiser window beta=9, no dither.
  ratio     g = gcd(in, out); L = out/g phases; M = in/g.
  design    factor = min(out*cutoff/in, 1); taps = ceil(32/factor) rounded up to even;
            center = (taps-1)//2;
            h[ph][i] = sinc(pi*((i-center) - ph/L)*factor) * I0(beta*sqrt(max(1-w^2,0))),
            w = 2*((i-center) - ph/L)/taps ... normalised so every phase sums to 1.
  run       y[m] = sum_i h[ph][i] * x[idx - center + i],  idx = (m*M)//L, ph = (m*M)%L
  edges     x[-k] = x[k] (reflect) before the start; x[n-1+k] = x[n-k] (symmetric) after
            the end;  n_out = what one-shot swr_convert + flush returns (out_len below restates the
            library's buffering: ceil((n_in - taps/2 + R)*L/M) with R the flush reflection, taps/2 or
            one less) — equal to the real library on every length swept.
  convert   s16 in: x/32768 ; stereo→mono: 0.5*L + 0.5*R (float) when resampling;
            s16 out: clip(rint(32768*y)) round-half-even.
            Same-rate s16 stereo → mono stays integer: (L + R + 1) >> 1.

Pinning: tests/test_oracle_resample.py checks this file against the real library
(≤1 LSB, ≥99.8 % identical on s16; ≤2e-6 on float) and against SURVEY A.5 R1/R2.
"""
from __future__ import annotations

from math import gcd

import numpy as np

CUTOFF = 0.97
FILTER_SIZE = 32
KAISER_BETA = 9.0


def ratio(in_rate: int, out_rate: int) -> tuple[int, int]:
    g = gcd(in_rate, out_rate)
    return out_rate // g, in_rate // g  # L (phases), M


def n_taps(in_rate: int, out_rate: int) -> int:
    factor = min(out_rate * CUTOFF / in_rate, 1.0)
    t = int(np.ceil(FILTER_SIZE / factor))
    return (t + 1) & ~1


def out_len(n_in: int, in_rate: int, out_rate: int) -> int:
    """Samples a one-shot ``swr_convert(all input)`` + flush returns (what ``ffmpeg -i IN -ar out_rate`` writes).

    libswresample/swresample.c resample() + resample.c (invert_initial_buffer, swri_resample, resample_flush), restated:
    the stream starts with ``center`` reflected samples; a call emits every output whose ``taps`` window fits in what is
    buffered, i.e. outputs m with m*M < (1 + V - taps)*L for V virtual samples; at flush the library appends
    R = (min(buffered, taps) + 1) // 2 mirrored samples, where ``buffered`` is what the first call left unconsumed
    (taps-1 down to taps-1-ceil(M/L)+1, so R is taps/2 or one less), and emits what fits then.  Inputs shorter than
    taps + 1 sit in the initial buffer until flush (buffered = n_in) and yield nothing unless n_in + R >= taps + 1.
    Checked against the real library on 25 030 (length, rate, channels, format) cases: 0 mismatches
    (tests/test_oracle.py::test_out_len_matches_library_sweep)."""
    if n_in <= 0:
        return 0
    L, M = ratio(in_rate, out_rate)
    if L == 1 and M == 1:
        return n_in
    taps = n_taps(in_rate, out_rate)
    center = (taps - 1) // 2
    if n_in < taps + 1:
        buffered = n_in
    else:
        n1 = max(0, -((-(n_in - taps // 2) * L) // M))      # outputs of the first call
        buffered = center + n_in - (n1 * M) // L
    refl = (min(buffered, taps) + 1) // 2
    if n_in < taps + 1 and n_in + refl < taps + 1:
        return 0
    return max(0, -((-(n_in - taps // 2 + refl) * L) // M))


def design(in_rate: int, out_rate: int) -> np.ndarray:
    """[L, taps] float64 filter bank (each phase normalised to unit DC gain)."""
    L, _ = ratio(in_rate, out_rate)
    factor = min(out_rate * CUTOFF / in_rate, 1.0)
    taps = n_taps(in_rate, out_rate)
    center = (taps - 1) // 2
    i = np.arange(taps, dtype=np.float64)[None, :]
    ph = np.arange(L, dtype=np.float64)[:, None]
    x = np.pi * ((i - center) - ph / L) * factor
    with np.errstate(invalid="ignore", divide="ignore"):
        y = np.where(x == 0.0, 1.0, np.sin(x) / x)
    w = 2.0 * x / (factor * taps * np.pi)
    y = y * np.i0(KAISER_BETA * np.sqrt(np.maximum(1.0 - w * w, 0.0)))
    y /= y.sum(axis=1, keepdims=True)
    return y


def to_mono_float(pcm: np.ndarray) -> np.ndarray:
    """decode + downmix to float64 in nominal ±1.0 (exact for s16 input)."""
    a = np.asarray(pcm)
    if a.ndim == 1:
        a = a[:, None]
    if a.dtype == np.int16:
        f = a.astype(np.float64) / 32768.0
    else:
        f = a.astype(np.float64)
    if f.shape[1] == 1:
        return f[:, 0]
    if f.shape[1] == 2:
        return 0.5 * f[:, 0] + 0.5 * f[:, 1]
    raise ValueError("only mono/stereo are in scope")


def quantise_s16(y: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(np.asarray(y, dtype=np.float64) * 32768.0), -32768, 32767).astype(np.int16)


def resample_float(pcm: np.ndarray, in_rate: int, out_rate: int = 16000,
                   taps_dtype=np.float32) -> np.ndarray:
    """float64 mono output before quantisation.  Taps are rounded to float32 first,
    as the library stores them (set taps_dtype=np.float64 to keep them exact)."""
    x = to_mono_float(pcm)
    n = x.shape[0]
    L, M = ratio(in_rate, out_rate)
    if L == 1 and M == 1:
        return x.copy()
    h = design(in_rate, out_rate).astype(taps_dtype).astype(np.float64)
    taps = h.shape[1]
    center = (taps - 1) // 2
    n_out = out_len(n, in_rate, out_rate)
    if n_out == 0:
        return np.zeros(0, dtype=np.float64)
    # extended signal: reflect head (edge not repeated), symmetric tail (edge repeated); a clip shorter than the filter
    # (n_out > 0 needs n > center) is extended the same way, as far as its own length allows
    head = x[1:center + 1][::-1]
    tail = x[::-1][:taps]
    xe = np.concatenate([head, x, tail, np.zeros(max(0, taps - len(tail)))])
    m = np.arange(n_out, dtype=np.int64)
    idx = (m * M) // L
    ph = (m * M) % L
    y = np.zeros(n_out, dtype=np.float64)
    # accumulate tap by tap (vectorised over outputs)
    for i in range(taps):
        y += h[ph, i] * xe[idx + i]
    return y


def convert(pcm: np.ndarray, in_rate: int, out_rate: int = 16000) -> np.ndarray:
    """int16 mono at out_rate, i.e. what the reference's WAV holds."""
    a = np.asarray(pcm)
    if a.ndim == 1:
        a = a[:, None]
    L, M = ratio(in_rate, out_rate)
    if L == 1 and M == 1:
        if a.dtype == np.int16:
            if a.shape[1] == 1:
                return a[:, 0].copy()
            s = a[:, 0].astype(np.int32) + a[:, 1].astype(np.int32) + 1
            return (s >> 1).astype(np.int16)
        return quantise_s16(to_mono_float(a))
    return quantise_s16(resample_float(a, in_rate, out_rate))
