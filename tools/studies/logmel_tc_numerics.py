"""CPU numerics study for a tensor-core log-mel (round-2 candidate; nothing here is on the product path).

Question: how many f16 operand planes does a DFT-as-GEMM need to keep the Whisper log-mel within 1e-4 of the float64
oracle, when the accumulator is fp32?  Formulation (per frame of 400 reflect-padded s16 samples x[n], periodic Hann w):

    E[n] = x[n] + x[400-n], O[n] = x[n] - x[400-n]   (n = 1..199; E[200] = x[200]; w[0] = 0)      17-bit integers
    Re[k] =  sum_n E[n] * (w[n] cos(2 pi k n / 400))       K = 200, N = 201
    Im[k] = -sum_n O[n] * (w[n] sin(2 pi k n / 400))       K = 199, N = 199
    E, O -> planes  hi = rint(v / 2^s), lo = v - 2^s hi  (exact in f16);  basis -> f16 hi + f16 (residual * 2^11)

Run:  python tools/studies/logmel_tc_numerics.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import whisper_logmel as W
from audio_processor_b200 import synth


def frames_of(pcm_s16, padding=0):
    x = np.concatenate([pcm_s16.astype(np.int64), np.zeros(padding, np.int64)])
    x = np.concatenate([x[1:201][::-1], x, x[-201:-1][::-1]])
    T = (len(x) - 400) // 160 + 1
    idx = np.arange(T)[:, None] * 160 + np.arange(400)[None, :]
    return x[idx][:-1]                                         # whisper drops the last frame


def planes_int(v, shift):
    hi = np.rint(v / (1 << shift))
    lo = v - hi * (1 << shift)
    assert np.abs(hi).max() <= 2048 and np.abs(lo).max() <= 2048
    return hi.astype(np.float32), lo.astype(np.float32)


def planes_basis(B, nplanes):
    out, r, scale = [], B.copy(), 1.0
    for _ in range(nplanes):
        h = (r * scale).astype(np.float16).astype(np.float64)
        out.append((h.astype(np.float32), scale))
        r = r - h / scale
        scale *= 2048.0
    return out


TRUNCATE = False        # True: emulate the tcgen05 accumulator the probe saw (tools/probes/umma_accum_probe.cu)


def trunc32(x64):
    """float64 -> float32 rounding toward zero."""
    f = x64.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(x64)
    return np.where(over, np.nextafter(f, np.float32(0)), f).astype(np.float32)


def mm32(a, b):
    if not TRUNCATE:
        return a.astype(np.float32) @ b.astype(np.float32)      # fp32 accumulate (blocked order; a proxy for the MMA)
    # K in steps of 16: the 16 products of a step are summed exactly, the running fp32 accumulator truncates toward zero
    acc = np.zeros((a.shape[0], b.shape[1]), np.float32)
    for k0 in range(0, a.shape[1], 16):
        acc = trunc32(acc.astype(np.float64) + a[:, k0:k0 + 16].astype(np.float64) @ b[k0:k0 + 16].astype(np.float64))
    return acc


def run(pcm, n_mels, shift, nb, drop_lolo):
    fr = frames_of(pcm)
    n = np.arange(1, 201)
    E = fr[:, 1:201].copy(); E[:, :199] += fr[:, 399:200:-1]
    O = fr[:, 1:200] - fr[:, 399:200:-1]
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)
    k = np.arange(201)
    C = w[1:201, None] * np.cos(2 * np.pi * n[:, None] * k[None, :] / 400)
    S = -w[1:200, None] * np.sin(2 * np.pi * n[:199, None] * k[None, :] / 400)
    out = []
    for X, B in ((E, C), (O, S)):
        xh, xl = planes_int(X, shift)
        acc = np.zeros((X.shape[0], 201), np.float32)
        for pi_, (bp, sc) in enumerate(planes_basis(B, nb)):
            # each (x plane, basis plane) pair = one MMA pass into its own fp32 accumulator, summed small-to-large in the epilogue
            acc_p = mm32(xh, bp) * np.float32((1 << shift) / sc)
            if not (drop_lolo and pi_ == nb - 1):
                acc_p = acc_p + mm32(xl, bp) * np.float32(1.0 / sc)
            acc = acc + acc_p
        out.append(acc / np.float32(32768.0))
    power = out[0].astype(np.float32) ** 2 + out[1].astype(np.float32) ** 2
    mel = power @ W.mel_filters(n_mels).T.astype(np.float32)
    lg = np.log10(np.maximum(mel, 1e-10))
    lg = np.maximum(lg, lg.max() - 8.0)
    return ((lg + 4.0) / 4.0).T


if __name__ == "__main__":
    for seed, sil in ((1, 0.25), (7, 0.5)):
        pcm = synth.clip_numpy(seed, 16000, 1, 20.0, sil).reshape(-1) if hasattr(synth, "clip_numpy") else None
        if pcm is None:
            rng = np.random.default_rng(seed)
            t = np.arange(20 * 16000) / 16000
            sp = sum(np.sin(2 * np.pi * f * t + rng.uniform(0, 6.28)) for f in (180, 360, 900, 2100, 3300)) * 0.08
            gate = (np.floor(t / 2.5) % 2 == 0)
            y = np.where(gate, sp + rng.normal(0, 0.02, t.size), rng.normal(0, 0.0008, t.size))
            pcm = np.clip(np.rint(y * 32768), -32768, 32767).astype(np.int16)
        ref = W.log_mel_spectrogram(torch.from_numpy(pcm.astype(np.float32) / 32768.0), 80).numpy()
        for shift, nb, drop in ((8, 2, False), (6, 2, True), (6, 2, False), (8, 1, False), (6, 3, True)):
            got = run(pcm, 80, shift, nb, drop)
            passes = 2 * nb - (1 if drop else 0)
            print(f"seed {seed}: x split 2^{shift}, basis planes {nb}, lo*lo {'dropped' if drop else 'kept'}: "
                  f"{passes} MMA passes, max |err| = {np.abs(got - ref).max():.2e}")


# ---- formulation B: unwindowed DFT, radix-2 in k + symmetric folds in n (four K<=101 x N<=101 products, basis 4x smaller:
# it fits one SM's shared memory), Hann applied in the frequency domain as the 3-tap  Xw[k] = X[k]/2 - (X[k-1] + X[k+1])/4 ----
def run_radix2(pcm, n_mels, shift, nb):
    fr = frames_of(pcm)
    a = fr[:, :200] + fr[:, 200:]
    d = fr[:, :200] - fr[:, 200:]
    n = np.arange(101)
    ae = a[:, :101].copy(); ae[:, 1:100] += a[:, 199:100:-1]                 # n = 0..100
    ao = a[:, 1:100] - a[:, 199:100:-1]                                      # n = 1..99
    de = d[:, :100].copy(); de[:, 1:100] -= d[:, 199:100:-1]                 # n = 0..99
    do = d[:, 1:101].copy(); do[:, :99] += d[:, 199:100:-1]                  # n = 1..100
    j = np.arange(101)
    Bc_e = np.cos(2 * np.pi * n[:, None] * j[None, :] / 200)                 # [101, 101]
    Bs_e = -np.sin(2 * np.pi * n[1:100, None] * j[None, :] / 200)            # [99, 101]
    jo = np.arange(100)
    Bc_o = np.cos(2 * np.pi * n[:100, None] * (2 * jo[None, :] + 1) / 400)   # [100, 100]
    Bs_o = -np.sin(2 * np.pi * n[1:101, None] * (2 * jo[None, :] + 1) / 400)  # [100, 100]

    def gemm(X, B):
        assert np.abs(X).max() < (1 << 18)
        xh, xl = planes_int(X, shift)
        acc = np.zeros((X.shape[0], B.shape[1]), np.float32)
        for bp, sc in reversed(planes_basis(B, nb)):
            acc = acc + (mm32(xl, bp) * np.float32(1.0 / sc) + mm32(xh, bp) * np.float32((1 << shift) / sc))
        return acc

    T = fr.shape[0]
    re = np.zeros((T, 203), np.float32); im = np.zeros((T, 203), np.float32)   # slot k + 1; slots 0 and 202 = mirrored bins
    re[:, 1:203:2] = gemm(ae, Bc_e); im[:, 1:203:2] = gemm(ao, Bs_e)
    re[:, 2:202:2] = gemm(de, Bc_o); im[:, 2:202:2] = gemm(do, Bs_o)
    re[:, 0] = re[:, 2]; im[:, 0] = -im[:, 2]; re[:, 202] = re[:, 200]; im[:, 202] = -im[:, 200]
    h, q = np.float32(0.5), np.float32(0.25)
    wr = (h * re[:, 1:202] - q * (re[:, 0:201] + re[:, 2:203])) / np.float32(32768.0)
    wi = (h * im[:, 1:202] - q * (im[:, 0:201] + im[:, 2:203])) / np.float32(32768.0)
    power = wr * wr + wi * wi
    mel = power @ W.mel_filters(n_mels).T.astype(np.float32)
    lg = np.log10(np.maximum(mel, 1e-10))
    lg = np.maximum(lg, lg.max() - 8.0)
    return ((lg + 4.0) / 4.0).T


if __name__ == "__main__":
    rng = np.random.default_rng(3)
    cases = {"speech+gaps": pcm,
             "full-scale tone + noise floor": np.clip(np.rint(32000 * np.sin(2 * np.pi * 1234.5 * np.arange(160000) / 16000)
                                                              + rng.normal(0, 2, 160000)), -32768, 32767).astype(np.int16),
             "white noise": np.clip(np.rint(rng.normal(0, 3000, 160000)), -32768, 32767).astype(np.int16)}
    for name, p in cases.items():
        ref = W.log_mel_spectrogram(torch.from_numpy(p.astype(np.float32) / 32768.0), 80).numpy()
        f32 = W.log_mel_spectrogram(torch.from_numpy(p.astype(np.float32) / 32768.0), 80, dtype=torch.float32).numpy()
        print(f"{name}: torch.stft f32 vs f64 max |err| = {np.abs(f32 - ref).max():.2e}")
        print(f"  windowed fold  (4 passes, K 200+199, N 201): {np.abs(run(p, 80, 6, 2, False) - ref).max():.2e}")
        for shift, nb in ((7, 2), (6, 2), (6, 3)):
            print(f"  radix-2 unwindowed, x split 2^{shift}, basis planes {nb} ({2 * nb} passes, 4 x [K<=101, N<=101]): "
                  f"{np.abs(run_radix2(p, 80, shift, nb) - ref).max():.2e}")
        TRUNCATE = True
        print(f"  radix-2 unwindowed, x split 2^7, 4 passes, accumulator truncating toward zero every 16 products: "
              f"{np.abs(run_radix2(p, 80, 7, 2) - ref).max():.2e}")
        TRUNCATE = False
