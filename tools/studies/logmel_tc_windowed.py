"""CPU numerics study for the round-2 tensor-core log-mel (nothing here is on the product path).

Formulation C (what csrc/logmel_tc.cu implements): Whisper's own order — periodic Hann in the TIME domain in f32 —
followed by the two exact DFT folds, so that even and odd bins decouple completely (no 3-tap in the epilogue):

    xw[n] = w[n] * x[n]                     (f32; x scaled to the s16 / 4 domain so that every folded value fits f16's range)
    ge[n] = xw[n] + xw[n+200],  go[n] = xw[n] - xw[n+200]          n = 0..199     (time aliasing: even / odd bins)
    ae[n] = ge[n] + ge[200-n]   (n = 0..100; ae[0] = ge[0], ae[100] = ge[100]),   ao[n] = ge[n] - ge[200-n]   (n = 1..99)
    de[n] = go[n] - go[200-n]   (n = 0..99;  de[0] = go[0]),                      do[n] = go[n] + go[200-n]   (n = 1..100; do[100] = go[100])
    Re X[2j]   =  sum ae[n] cos(pi j n / 100)          Im X[2j]   = -sum ao[n] sin(pi j n / 100)
    Re X[2j+1] =  sum de[n] cos(pi (2j+1) n / 200)     Im X[2j+1] = -sum do[n] sin(pi (2j+1) n / 200)
    every operand v -> f16 planes hi = f16(v), lo = f16(v - hi); basis c -> f16(c), f16(c - f16(c));  four plane products
    per k-step of 16 into one fp32 accumulator that truncates toward zero (tools/probes/umma_accum_probe.cu).

Run:  python tools/studies/logmel_tc_windowed.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import whisper_logmel as W

F32 = np.float32


def frames_of(x, padding=0):
    x = np.concatenate([x, np.zeros(padding, x.dtype)])
    x = np.concatenate([x[1:201][::-1], x, x[-201:-1][::-1]])
    T = (len(x) - 400) // 160 + 1
    idx = np.arange(T)[:, None] * 160 + np.arange(400)[None, :]
    return x[idx][:-1]


def trunc32(x64):
    f = x64.astype(F32)
    over = np.abs(f.astype(np.float64)) > np.abs(x64)
    return np.where(over, np.nextafter(f, F32(0)), f).astype(F32)


def split16(v):
    hi = v.astype(np.float16)
    lo = (v - hi.astype(F32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64)


def mma_chain(A, B, truncate=True):
    """A [T, K] f32 values, B [K, N] float64 basis; K padded to 16s; per k-step hh, hl, lh, ll accumulate, fp32 truncating"""
    K = (A.shape[1] + 15) // 16 * 16
    A = np.pad(A, ((0, 0), (0, K - A.shape[1]))); B = np.pad(B, ((0, K - B.shape[0]), (0, 0)))
    ah, al = split16(A.astype(F32))
    bh = B.astype(np.float16).astype(np.float64); bl = (B - bh).astype(np.float16).astype(np.float64)
    acc = np.zeros((A.shape[0], B.shape[1]), F32)
    for k0 in range(0, K, 16):
        for a, b in ((ah, bh), (ah, bl), (al, bh), (al, bl)):
            s = acc.astype(np.float64) + a[:, k0:k0 + 16] @ b[k0:k0 + 16]
            acc = trunc32(s) if truncate else s.astype(F32)
    return acc


def run(x_scaled, n_mels, out_scale, padding=0):
    """x_scaled: f32 samples already in the s16/4 domain; out_scale: factor back to Whisper's +-1 domain"""
    fr = frames_of(x_scaled.astype(F32), padding)
    n = np.arange(400)
    w = (0.5 - 0.5 * np.cos(2 * np.pi * n / 400)).astype(F32)
    xw = (fr * w).astype(F32)
    ge = (xw[:, :200] + xw[:, 200:]).astype(F32); go = (xw[:, :200] - xw[:, 200:]).astype(F32)
    ae = ge[:, :101].copy(); ae[:, 1:100] = ge[:, 1:100] + ge[:, 199:100:-1]
    ao = (ge[:, 1:100] - ge[:, 199:100:-1]).astype(F32)
    de = go[:, :100].copy(); de[:, 1:100] = go[:, 1:100] - go[:, 199:100:-1]
    do = np.zeros((fr.shape[0], 100), F32); do[:, :99] = go[:, 1:100] + go[:, 199:100:-1]; do[:, 99] = go[:, 100]
    m = np.arange(101)
    j = np.arange(101); jo = np.arange(100)
    Ce = np.cos(np.pi * m[:, None] * j[None, :] / 100)
    Se = -np.sin(np.pi * m[1:100, None] * j[None, :] / 100)
    Co = np.cos(np.pi * m[:100, None] * (2 * jo[None, :] + 1) / 200)
    So = -np.sin(np.pi * m[1:101, None] * (2 * jo[None, :] + 1) / 200)
    T = fr.shape[0]
    re = np.zeros((T, 201), F32); im = np.zeros((T, 201), F32)
    re[:, 0::2] = mma_chain(ae, Ce); im[:, 0::2] = mma_chain(ao, Se)
    re[:, 1::2] = mma_chain(de, Co); im[:, 1::2] = mma_chain(do, So)
    power = (re * re + im * im).astype(F32)
    mel = (power @ (W.mel_filters(n_mels).T.astype(F32) * F32(out_scale * out_scale))).astype(F32)
    lg = np.log10(np.maximum(mel, 1e-10))
    lg = np.maximum(lg, lg.max() - 8.0)
    return ((lg + 4.0) / 4.0).T


if __name__ == "__main__":
    rng = np.random.default_rng(3)
    t = np.arange(20 * 16000) / 16000
    sp = sum(np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28)) for f in (180, 360, 900, 2100, 3300)) * 0.08
    gate = (np.floor(t / 2.5) % 2 == 0)
    y = np.where(gate, sp + rng.normal(0, 0.02, t.size), rng.normal(0, 0.0008, t.size))
    cases = {
        "speech+gaps (s16)": np.clip(np.rint(y * 32768), -32768, 32767).astype(np.int16),
        "full-scale tone + noise floor (s16)": np.clip(np.rint(32000 * np.sin(2 * np.pi * 1234.5 * np.arange(160000) / 16000) + rng.normal(0, 2, 160000)), -32768, 32767).astype(np.int16),
        "white noise (s16)": np.clip(np.rint(rng.normal(0, 3000, 160000)), -32768, 32767).astype(np.int16),
        "very quiet noise, +-3 LSB (s16)": np.clip(np.rint(rng.normal(0, 1.5, 160000)), -32768, 32767).astype(np.int16),
        "N(0, 0.1) clipped (f32, cfg3)": np.clip(rng.normal(0, 0.1, 160000), -1, 1).astype(F32),
        "tone 0.9 + 1e-5 noise (f32)": (0.9 * np.sin(2 * np.pi * 777.7 * np.arange(160000) / 16000) + rng.normal(0, 1e-5, 160000)).astype(F32),
    }
    for name, p in cases.items():
        for n_mels in (80, 128):
            if p.dtype == np.int16:
                wav = p.astype(F32) / 32768.0
                xs, osc = p.astype(F32) * F32(0.25), 4.0 / 32768.0
            else:
                wav, xs, osc = p, p * F32(8192.0), 1.0 / 8192.0
            ref = W.log_mel_spectrogram(torch.from_numpy(wav), n_mels).numpy()
            f32 = W.log_mel_spectrogram(torch.from_numpy(wav), n_mels, dtype=torch.float32).numpy()
            got = run(xs, n_mels, osc)
            print(f"{name:40s} n_mels {n_mels:3d}: torch.stft f32 {np.abs(f32 - ref).max():.2e}   formulation C {np.abs(got - ref).max():.2e}")
