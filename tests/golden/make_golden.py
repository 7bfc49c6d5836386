"""Mint the golden vectors under tests/golden/ (run in the build container; commits the outputs).

The reference has no tests or fixtures for this path, so the vectors come from the engines
the reference reaches: the real FFmpeg 8.0.1 libswresample in this image (oracle/swr_ref.py), stdlib
audioop via the literal pydub restatement (oracle/pydub_silence.py), and torch.stft float64
(oracle/whisper_logmel.py).  SURVEY.md Appendix A.5 lists the same known answers.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import pydub_silence as ps, swr_ref, whisper_logmel as wl  # noqa: E402


def tone_pair(rate, n):
    t = np.arange(n) / float(rate)
    L = np.round(8000 * np.sin(2 * np.pi * 1000 * t)).astype(np.int16)
    R = np.round(8000 * np.sin(2 * np.pi * 3000 * t)).astype(np.int16)
    return np.stack([L, R], 1)


def sig(parts, extra=0):
    a = [np.full(int(round(d * 16000)), amp, dtype=np.int16) for d, amp in parts]
    if extra:
        a.append(np.zeros(extra, dtype=np.int16))
    return np.concatenate(a)


SILENCE_CASES = {
    "K1": ([(3, 1000), (2, 0), (3, 1000)], 0, (1000, -40, 200, 1)),
    "K2": ([(3, 1000), (2, 0), (3, 1000)], 0, (1000, -16, 100, 1)),
    "K3": ([(2, 1000), (0.9, 0), (2, 1000)], 0, (1000, -40, 200, 1)),
    "K4": ([(1, 8000), (1.2, 0), (0.3, 8000), (1.2, 0), (1, 8000)], 0, (1000, -40, 500, 1)),
    "K4b": ([(1, 8000), (1.2, 0), (0.3, 8000), (1.2, 0), (1, 8000)], 0, (1000, -40, 700, 1)),
    "K4c": ([(3, 1000), (2, 0), (3, 1000)], 0, (1000, -40, 200, 10)),
    "K5": ([(1.5, 0), (2, 8000), (1.5, 0)], 0, (1000, -40, 100, 1)),
    "K6": ([(3, 0)], 0, (1000, -40, 100, 1)),
    "K7": ([(0.5, 0)], 0, (1000, -40, 100, 1)),
    "K8": ([(2, 8000), (1.5, 0)], 7, (1000, -40, 100, 1)),
    "K9": ([(2, 8000), (1.5, 0)], 9, (1000, -40, 100, 1)),
}


def main():
    assert swr_ref.available(), "the bundled libswresample is required to mint resampler vectors"
    out = {}
    # --- resampler: R1/R2 tone pairs + one second of seeded noise per rate, through the real library
    res = {"ffmpeg": swr_ref.versions()[0], "swresample": swr_ref.versions()[1]}
    arrays = {}
    for name, rate, n in (("R1", 44100, 4410), ("R2", 48000, 4800)):
        x = tone_pair(rate, n)
        y = swr_ref.convert(x, rate)
        arrays[f"{name}_out"] = y
        res[name] = dict(rate=rate, n_in=n, n_out=int(len(y)), first16=y[:16].tolist(), last8=y[-8:].tolist(),
                         sum=int(y.astype(np.int64).sum()), sumsq=int((y.astype(np.int64) ** 2).sum()))
    rng = np.random.default_rng(20261018)
    for name, rate in (("N441", 44100), ("N480", 48000), ("N220", 22050)):
        x = (rng.standard_normal((rate // 2, 2)) * 5000).clip(-32768, 32767).astype(np.int16)
        arrays[f"{name}_in"] = x
        arrays[f"{name}_out"] = swr_ref.convert(x, rate)
        res[name] = dict(rate=rate, n_in=int(len(x)), n_out=int(len(arrays[f"{name}_out"])))
    xm = (rng.standard_normal(22050) * 0.25).astype(np.float32)
    arrays["F441_in"] = xm
    arrays["F441_out"] = swr_ref.convert(xm, 44100)
    out["resample"] = res
    # --- silence: literal pydub loop over audioop
    sil = {}
    for k, (parts, extra, (W, th, keep, step)) in SILENCE_CASES.items():
        x = sig(parts, extra)
        s = ps.Segment(x)
        sil[k] = dict(parts=parts, extra=extra, params=[W, th, keep, step], len_ms=len(s),
                      silent=ps.detect_silence(s, W, th, step), nonsilent=ps.detect_nonsilent(s, W, th, step),
                      kept=ps.kept_ranges(s, W, th, keep, step),
                      n_keep=int(len(ps.strip_silence(s, min_silence_len=W, silence_thresh=th, keep_silence=keep,
                                                      seek_step=step))))
    out["silence"] = sil
    out["thresholds"] = {str(db): ps.db_to_float(db) * 32768.0 for db in (-16, -30, -40, -50, -60)}
    # --- log-mel: float64 oracle on 1 s of 0.5*sin(2*pi*1000 t) and on seeded noise
    t = np.arange(16000) / 16000.0
    tone = (0.5 * np.sin(2 * np.pi * 1000 * t)).astype(np.float32)
    noise = (rng.standard_normal(16000 * 2 + 123) * 0.1).astype(np.float32)
    arrays["mel_noise_in"] = noise
    mel = {}
    for nm in (80, 128):
        m = wl.log_mel_spectrogram(tone, nm).numpy()
        arrays[f"mel_tone_{nm}"] = m.astype(np.float32)
        mel[f"M{nm}"] = dict(shape=list(m.shape), max=float(m.max()), argmax_mel=int(m.argmax() // m.shape[1]),
                             min=float(m.min()), mean=float(m.mean()))
        arrays[f"mel_noise_{nm}_pad480"] = wl.log_mel_spectrogram(noise, nm, padding=480).numpy().astype(np.float32)
    out["logmel"] = mel
    f = wl.mel_filters(80)
    out["filterbank"] = dict(nnz80=int((f != 0).sum()), nnz128=int((wl.mel_filters(128) != 0).sum()), f_1_0=float(f[0, 1]))
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **arrays)
    print("wrote golden.json, golden_arrays.npz")


if __name__ == "__main__":
    main()
