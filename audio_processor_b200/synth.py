"""Synthetic clips for benchmarks and parity tests (SURVEY.md §8d recipe).  torch is used as a data
generator only; parity is always computed on the very bytes produced here (copied D2H), never on a
re-generated stream.

synth_clip(seed, sr, channels, dur_s, sil): alternating *speech* and *gap* spans.
  speech span length  ~ 1.5 s + Exp(mean 4 s)
  gap span length     ~ 1.2 s + Exp(mean max(sil/(1-sil)*5.5 s - 1.2 s, 0.05 s))      (every gap > min_silence_len)
  speech = 0.08 * sum of 5 sinusoids (180, 360, 900, 2100, 3300 Hz, random phase per channel) + N(0, 0.02)
  gap    = N(0, 0.0008)                                                                (about -62 dBFS)
  -> * 32768, round, clip to int16; stereo channels share the span layout, noise is independent.
"""
from __future__ import annotations

import math

import numpy as np

TONES_HZ = (180.0, 360.0, 900.0, 2100.0, 3300.0)


def span_layout(seed: int, dur_s: float, sil: float):
    """-> sorted array of span boundaries (seconds) and a bool per span (True = speech)."""
    rng = np.random.default_rng(seed)
    gap_extra = max(sil / max(1.0 - sil, 1e-6) * 5.5 - 1.2, 0.05)
    t, speech = 0.0, bool(rng.integers(0, 2))
    edges, kinds = [0.0], []
    while t < dur_s:
        d = 1.5 + rng.exponential(4.0) if speech else 1.2 + rng.exponential(gap_extra)
        t += d
        edges.append(min(t, dur_s))
        kinds.append(speech)
        speech = not speech
    return np.asarray(edges), np.asarray(kinds, dtype=bool)


def synth_clip(seed: int, sr: int, channels: int, dur_s: float, sil: float, device="cpu", chunk: int = 1 << 24):
    """int16 tensor [n] (mono) or [n, channels] on `device`."""
    import torch
    n = int(round(dur_s * sr))
    edges, kinds = span_layout(seed, dur_s, sil)
    edge_samp = torch.as_tensor(np.round(edges * sr).astype(np.int64), device=device)
    kind_t = torch.as_tensor(kinds, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed) * 7919 + 17)
    phases = torch.rand((channels, len(TONES_HZ)), generator=gen, device=device) * (2 * math.pi)
    out = torch.empty((n, channels), dtype=torch.int16, device=device)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        idx = torch.arange(lo, hi, device=device, dtype=torch.int64)
        span = torch.searchsorted(edge_samp, idx, right=True) - 1
        span.clamp_(0, kind_t.numel() - 1)
        is_speech = kind_t[span]
        tsec = idx.to(torch.float64) / sr
        for c in range(channels):
            tone = torch.zeros(hi - lo, dtype=torch.float32, device=device)
            for k, f in enumerate(TONES_HZ):
                tone += torch.sin((2 * math.pi * f) * tsec + phases[c, k].double()).float()
            noise = torch.randn(hi - lo, generator=gen, device=device)
            x = torch.where(is_speech, 0.08 * tone + 0.02 * noise, 0.0008 * noise)
            out[lo:hi, c] = torch.clamp(torch.round(x * 32768.0), -32768, 32767).to(torch.int16)
    return out[:, 0].contiguous() if channels == 1 else out


def noise_batch(seed: int, batch: int, n: int, device="cpu", std: float = 0.1):
    """cfg3-style input: [batch, n] float32 = N(0, std) clipped to +-1."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    x = torch.empty((batch, n), dtype=torch.float32, device=device)
    rows = max(1, (1 << 26) // max(n, 1))
    for lo in range(0, batch, rows):
        hi = min(batch, lo + rows)
        x[lo:hi] = torch.randn((hi - lo, n), generator=gen, device=device).mul_(std).clamp_(-1.0, 1.0)
    return x
