"""The oracles against the engines the reference reaches and against the frozen golden vectors.
(No GPU, no product code: this pins the checker itself.)"""
import numpy as np
import pytest
import torch

from oracle import pydub_silence as ps, resample_oracle as ro, swr_ref, whisper_logmel as wl
from tests import helpers as H

G = H.golden_json()
A = H.golden_arrays()
needs_swr = pytest.mark.skipif(not swr_ref.available(), reason="bundled libswresample not found")


# ---------------- resampler ----------------
@pytest.mark.parametrize("name", ["R1", "R2"])
def test_resample_oracle_vs_golden_tones(name):
    g = G["resample"][name]
    y = ro.convert(H.tone_pair(g["rate"], g["n_in"]), g["rate"])
    ref = A[f"{name}_out"]
    assert g["n_out"] == len(ref) == len(y)
    d = np.abs(y.astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d == 0).mean() >= 0.998
    assert ref[:16].tolist() == g["first16"] and ref[-8:].tolist() == g["last8"]
    assert int(ref.astype(np.int64).sum()) == g["sum"]


@pytest.mark.parametrize("name,rate", [("N441", 44100), ("N480", 48000), ("N220", 22050)])
def test_resample_oracle_vs_golden_noise(name, rate):
    y = ro.convert(A[f"{name}_in"], rate)
    ref = A[f"{name}_out"]
    assert len(y) == len(ref)
    d = np.abs(y.astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d == 0).mean() >= 0.995


def test_resample_oracle_float_mono_golden():
    y = ro.convert(A["F441_in"], 44100)
    assert len(y) == len(A["F441_out"])
    d = np.abs(y.astype(int) - A["F441_out"].astype(int))
    assert d.max() <= 1 and (d == 0).mean() >= 0.995


@needs_swr
@pytest.mark.parametrize("rate", [44100, 48000, 32000, 22050])
def test_resample_oracle_vs_live_library(rate):
    rng = np.random.default_rng(rate)
    x = (rng.standard_normal((rate + 13, 2)) * 6000).clip(-32768, 32767).astype(np.int16)
    lib16 = swr_ref.convert(x, rate)
    o = ro.convert(x, rate)
    assert len(o) == ro.out_len(len(x), rate, 16000) == len(lib16)
    d = np.abs(o.astype(int) - lib16.astype(int))
    assert d.max() <= 1 and (d == 0).mean() >= 0.995
    xm = x[:, 0].copy()
    libf = swr_ref.convert(xm, rate, out_fmt="flt")
    assert np.abs(libf - ro.resample_float(xm, rate)[:len(libf)]).max() < 2e-6     # filter design matches the library's taps


@needs_swr
@pytest.mark.parametrize("rate", [44100, 48000, 22050, 32000, 8000, 24000, 11025, 96000])
def test_out_len_matches_library_sweep(rate):
    """one-shot swr_convert + flush length for >= 1 000 consecutive input lengths per rate (round-1 VERDICT: the old
    ceil(n*L/M) overshot by one on a third of them), plus clips around and below the filter length, mono and stereo,
    s16 and f32"""
    taps = ro.n_taps(rate, 16000)
    lengths = list(range(1, 2 * taps + 40)) + list(range(5000, 6000)) + [rate * 60, rate * 60 + 1, rate * 3 + 7]
    bad = []
    for n in lengths:
        ch, dt = (1 if n % 2 else 2), (np.int16 if n % 3 else np.float32)
        x = np.zeros((n, ch) if ch > 1 else (n,), dtype=dt)
        got = len(swr_ref.convert(x, rate))
        if got != ro.out_len(n, rate, 16000):
            bad.append((n, got, ro.out_len(n, rate, 16000)))
    assert not bad, bad[:10]


@needs_swr
@pytest.mark.parametrize("rate", [44100, 48000, 22050, 8000])
def test_short_clip_values_match_library(rate):
    """clips shorter than the filter: the library still emits samples once n_in + reflection fills it; same extension rule"""
    taps = ro.n_taps(rate, 16000)
    rng = np.random.default_rng(rate)
    for n in range(taps // 2, taps + 12):
        x = (rng.standard_normal((n, 2)) * 8000).astype(np.int16)
        a, b = swr_ref.convert(x, rate), ro.convert(x, rate)
        assert len(a) == len(b), (n, len(a), len(b))
        if len(a):
            assert np.abs(a.astype(int) - b.astype(int)).max() <= 1


@needs_swr
def test_same_rate_paths_match_library():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((4001, 2)) * 9000).clip(-32768, 32767).astype(np.int16)
    assert np.array_equal(ro.convert(x, 16000), swr_ref.convert(x, 16000))          # (L+R+1)>>1
    assert np.array_equal(ro.convert(x[:, 0].copy(), 16000), x[:, 0])               # identity
    f = np.array([0.5, 1.5, 2.5, 32766.5, 32768.0, -49152.0], dtype=np.float32) / 32768.0
    assert ro.convert(f, 16000).tolist() == [0, 2, 2, 32766, 32767, -32768]          # round-half-even + clip


def test_filter_geometry():
    assert ro.ratio(44100, 16000) == (160, 441) and ro.n_taps(44100, 16000) == 92
    assert ro.ratio(48000, 16000) == (1, 3) and ro.n_taps(48000, 16000) == 100
    assert ro.out_len(44100 * 60, 44100, 16000) == 960000
    h = ro.design(44100, 16000)
    assert h.shape == (160, 92) and np.allclose(h.sum(1), 1.0)


# ---------------- silence ----------------
@pytest.mark.parametrize("name", sorted(G["silence"].keys()))
def test_silence_oracle_vs_golden(name):
    g = G["silence"][name]
    x = H.piecewise([tuple(p) for p in g["parts"]], g["extra"])
    W, th, keep, step = g["params"]
    s = ps.Segment(x)
    assert len(s) == g["len_ms"]
    assert ps.detect_silence(s, W, th, step) == g["silent"]
    assert ps.detect_nonsilent(s, W, th, step) == g["nonsilent"]
    assert ps.kept_ranges(s, W, th, keep, step) == g["kept"]
    # exact-integer vectorised form == literal audioop loop
    assert ps.detect_silence_fast(x, 16000, W, th, step) == g["silent"]
    assert ps.detect_nonsilent_fast(x, 16000, W, th, step) == g["nonsilent"]
    assert ps.kept_ranges_fast(x, 16000, W, th, keep, step) == g["kept"]
    assert len(ps.strip_silence_fast(x, 16000, min_silence_len=W, silence_thresh=th, keep_silence=keep, seek_step=step)) == g["n_keep"]


def test_threshold_table_and_rms_floor():
    for db, v in G["thresholds"].items():
        assert abs(ps.db_to_float(float(db)) * 32768.0 - v) < 1e-9
    assert int(np.floor(G["thresholds"]["-16"])) == 5193 and abs(G["thresholds"]["-40"] - 327.68) < 1e-9
    for T in (1, 10, 327, 328, 1000, 32767):
        x = np.full(16000, T, dtype=np.int16)
        x[-1] = T - 1
        assert ps.Segment(x).rms == T - 1          # audioop floors sqrt(mean square)


def test_silence_fast_equals_literal_random():
    rng = np.random.default_rng(7)
    for trial in range(40):
        n = int(rng.integers(9000, 70000)) + int(rng.integers(0, 16))
        x = H.random_speechlike(rng, n, min_span=800, max_span=12000)
        W = int(rng.choice([100, 250, 500, 1000]))
        th = float(rng.choice([-16, -30, -40, -50, -60.5]))
        keep = rng.choice([0, 50, 100, 200, 700])
        keep = bool(keep % 3) if trial % 11 == 0 else int(keep)
        step = int(rng.choice([1, 1, 3, 10, 25, 300, 1200]))
        s = ps.Segment(x)
        assert ps.detect_silence(s, W, th, step) == ps.detect_silence_fast(x, 16000, W, th, step), (trial, W, th, step)
        assert ps.detect_nonsilent(s, W, th, step) == ps.detect_nonsilent_fast(x, 16000, W, th, step)
        assert ps.kept_ranges(s, W, th, keep, step) == ps.kept_ranges_fast(x, 16000, W, th, keep, step)
        a = ps.strip_silence(s, min_silence_len=W, silence_thresh=th, keep_silence=keep, seek_step=step)
        b = ps.strip_silence_fast(x, 16000, min_silence_len=W, silence_thresh=th, keep_silence=keep, seek_step=step)
        assert np.array_equal(a, b)


# ---------------- log-mel ----------------
@pytest.mark.parametrize("nm", [80, 128])
def test_logmel_oracle_vs_golden(nm):
    t = np.arange(16000) / 16000.0
    tone = (0.5 * np.sin(2 * np.pi * 1000 * t)).astype(np.float32)
    m = wl.log_mel_spectrogram(tone, nm).numpy()
    g = G["logmel"][f"M{nm}"]
    assert list(m.shape) == g["shape"]
    assert abs(m.max() - g["max"]) < 1e-9 and int(m.argmax() // m.shape[1]) == g["argmax_mel"]
    assert abs(m.min() - (m.max() - 2.0)) < 1e-9 and abs(m.mean() - g["mean"]) < 1e-9
    assert np.abs(m - A[f"mel_tone_{nm}"]).max() < 1e-6
    mn = wl.log_mel_spectrogram(A["mel_noise_in"], nm, padding=480).numpy()
    assert np.abs(mn - A[f"mel_noise_{nm}_pad480"]).max() < 1e-6


def test_survey_known_answers():
    assert abs(G["logmel"]["M80"]["max"] - 1.439652) < 1e-6 and G["logmel"]["M80"]["argmax_mel"] == 26
    assert abs(G["logmel"]["M128"]["max"] - 1.474446) < 1e-6 and G["logmel"]["M128"]["argmax_mel"] == 42
    assert G["filterbank"]["nnz80"] == 391 and G["filterbank"]["nnz128"] == 394
    assert abs(G["filterbank"]["f_1_0"] - 0.024862594) < 1e-8


def test_logmel_oracle_vs_independent_restatements():
    from transformers import WhisperFeatureExtractor
    from transformers.audio_utils import mel_filter_bank
    for nm in (80, 128):
        hf = mel_filter_bank(num_frequency_bins=201, num_mel_filters=nm, min_frequency=0.0, max_frequency=8000.0,
                             sampling_rate=16000, norm="slaney", mel_scale="slaney").T
        assert np.abs(wl.mel_filters_f64(nm) - hf).max() < 1e-12
    rng = np.random.default_rng(1)
    a = (rng.standard_normal(480000) * 0.1).astype(np.float32)
    hfm = WhisperFeatureExtractor(feature_size=80)(a, sampling_rate=16000, return_tensors="np").input_features[0]
    assert np.abs(hfm - wl.log_mel_spectrogram(a, 80).numpy()).max() < 2e-5
    m32 = wl.log_mel_spectrogram(a, 80, dtype=torch.float32).numpy()            # what the reference executes
    assert np.abs(m32 - wl.log_mel_spectrogram(a, 80).numpy()).max() < 5e-5
    assert torch.all(wl.log_mel_spectrogram(np.zeros(16000, np.float32), 80) == -1.5)
    assert wl.log_mel_spectrogram(a[:960000 // 2], 80, padding=480000).shape == (80, 6000)
