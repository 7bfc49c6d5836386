"""Kernel logic on the TEST-ONLY CPU emulation (tests/emu): the very .cu sources, compiled by g++ against a fiber
model of the CUDA execution model, checked against the oracles.  This is how index math, barrier structure and
integer exactness are verified on a box without a GPU; the GPU parity tests proper are in test_gpu_parity.py."""
import numpy as np
import pytest

from oracle import pydub_silence as ps, resample_oracle as ro, swr_ref, whisper_logmel as wl
from tests import helpers as H

G = H.golden_json()
A = H.golden_arrays()


def _cmp16(y, ref, min_exact):
    assert len(y) == len(ref), (len(y), len(ref))       # exactly the library's one-shot length
    d = np.abs(y.astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d == 0).mean() >= min_exact, (d.max(), (d == 0).mean())


@pytest.mark.parametrize("name,rate,min_exact", [("R1", 44100, 0.998), ("R2", 48000, 0.998), ("N441", 44100, 0.998),
                                                 ("N480", 48000, 0.998), ("N220", 22050, 0.995)])
def test_emu_resample_golden(emu, name, rate, min_exact):
    x = H.tone_pair(rate, G["resample"][name]["n_in"]) if name.startswith("R") else A[f"{name}_in"]
    _cmp16(emu.resample(x, rate), A[f"{name}_out"], min_exact)


@pytest.mark.parametrize("rate,ch,n", [(44100, 2, 14112 * 3 + 1234), (44100, 1, 14112 * 2 + 77), (48000, 2, 1632 * 4 + 777),
                                       (48000, 1, 1632 * 3 + 5), (32000, 2, 9000), (11025, 1, 6000)])
def test_emu_resample_fast_and_generic_paths(emu, rate, ch, n):
    rng = np.random.default_rng(rate + ch)
    x = (rng.standard_normal((n, ch)) * 6000).clip(-32768, 32767).astype(np.int16)
    x = x[:, 0].copy() if ch == 1 else x
    y, yf, en = emu.resample(x, rate, want_f32=True, want_energy=True)
    assert len(y) == ro.out_len(n, rate, 16000)
    assert np.abs(yf - ro.resample_float(x, rate)).max() <= 1e-5          # normalised-PCM gate (float64 restatement)
    d = np.abs(y.astype(int) - ro.convert(x, rate).astype(int))
    assert d.max() <= 1
    if swr_ref.available():
        _cmp16(y, swr_ref.convert(x, rate), 0.998 if rate in (44100, 48000) else 0.95)   # named pairs: SURVEY A.4 gate
    assert np.array_equal(en.astype(np.int64), H.energy_oracle(y))         # exact-integer epilogue


def test_emu_resample_float_input_and_same_rate(emu):
    rng = np.random.default_rng(5)
    xf = (rng.standard_normal((20000, 2)) * 0.2).astype(np.float32)
    y = emu.resample(xf, 44100)
    assert np.abs(y.astype(int) - ro.convert(xf, 44100).astype(int)).max() <= 1
    x = (rng.standard_normal((4001, 2)) * 9000).clip(-32768, 32767).astype(np.int16)
    y, en = emu.resample(x, 16000, want_energy=True)
    assert np.array_equal(y, ro.convert(x, 16000)) and np.array_equal(en.astype(np.int64), H.energy_oracle(y))
    assert np.array_equal(emu.resample(x[:, 0].copy(), 16000), x[:, 0])
    f = np.array([0.5, 1.5, 2.5, 32766.5, 32768.0, -49152.0] * 40, dtype=np.float32) / 32768.0
    assert emu.resample(f, 16000)[:6].tolist() == [0, 2, 2, 32766, 32767, -32768]


@pytest.mark.parametrize("name", sorted(G["silence"].keys()))
def test_emu_silence_golden(emu, name):
    g = G["silence"][name]
    x = H.piecewise([tuple(p) for p in g["parts"]], g["extra"])
    W, th, keep, step = g["params"]
    r = emu.detect(x, 16000, W, th, keep, step)
    assert r["silent"] == g["silent"] and r["nonsilent"] == g["nonsilent"] and r["kept"] == g["kept"]
    assert int(r["info"][4]) == g["len_ms"] and len(r["compact"]) == g["n_keep"]


def test_emu_silence_random_properties(emu):
    rng = np.random.default_rng(11)
    for trial in range(24):
        n = int(rng.integers(20000, 150000)) + int(rng.integers(0, 16))      # up to ~9.4 s: several 2048-ms tiles
        x = H.random_speechlike(rng, n, min_span=800, max_span=20000)
        W = int(rng.choice([100, 250, 500, 1000]))
        th = float(rng.choice([-16, -30, -40, -50, -60.5]))
        keep = [0, 50, 100, 200, 700, True, False][trial % 7]
        step = int(rng.choice([1, 1, 1, 3, 10, 25, 300, 1200]))
        r = emu.detect(x, 16000, W, th, keep, step)
        assert r["silent"] == ps.detect_silence_fast(x, 16000, W, th, step), (trial, W, th, step)
        assert r["nonsilent"] == ps.detect_nonsilent_fast(x, 16000, W, th, step), (trial, W, th, step)
        assert r["kept"] == ps.kept_ranges_fast(x, 16000, W, th, keep, step), (trial, W, th, keep, step)
        st = ps.strip_silence_fast(x, 16000, min_silence_len=W, silence_thresh=th, keep_silence=keep, seek_step=step)
        assert np.array_equal(r["compact"], st)


def test_emu_silence_edge_cases(emu):
    z = np.zeros(0, dtype=np.int16)
    r = emu.detect(np.zeros(5, dtype=np.int16))           # len_ms == 0
    assert r["nonsilent"] == [[0, 0]] and len(r["compact"]) == 0
    x = np.full(16000 * 3, 9000, dtype=np.int16)          # nothing silent
    r = emu.detect(x, 16000, 1000, -40, 100, 1)
    assert r["silent"] == [] and r["nonsilent"] == [[0, 3000]] and np.array_equal(r["compact"], x)
    r = emu.detect(x[:8000], 16000, 1000, -40, 100, 1)    # shorter than the window
    assert r["nonsilent"] == [[0, 500]]
    assert len(z) == 0


@pytest.mark.parametrize("nm", [80, 128])
def test_emu_logmel_golden(emu, nm):
    t = np.arange(16000) / 16000.0
    tone = (0.5 * np.sin(2 * np.pi * 1000 * t)).astype(np.float32)
    assert np.abs(emu.log_mel(tone, nm) - A[f"mel_tone_{nm}"]).max() <= 1e-4
    assert np.abs(emu.log_mel(A["mel_noise_in"], nm, padding=480) - A[f"mel_noise_{nm}_pad480"]).max() <= 1e-4


def test_emu_logmel_cases(emu):
    rng = np.random.default_rng(2)
    # tone + -70 dB noise (worst probed dynamic range), zeros, click, batch with whole-call max and per-clip max
    t = np.arange(16000 * 2) / 16000.0
    hd = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.5 * 10 ** (-70 / 20) * rng.standard_normal(len(t))).astype(np.float32)
    assert np.abs(emu.log_mel(hd, 80) - wl.log_mel_spectrogram(hd, 80).numpy()).max() <= 1e-4
    assert np.all(emu.log_mel(np.zeros(16000, np.float32), 80) == -1.5)
    click = np.zeros(8000, np.float32); click[4000] = 1.0
    assert np.abs(emu.log_mel(click, 128, padding=4000) - wl.log_mel_spectrogram(click, 128, padding=4000).numpy()).max() <= 1e-4
    b = (rng.standard_normal((3, 4000)) * np.array([[0.3], [0.01], [0.0003]])).astype(np.float32)
    assert np.abs(emu.log_mel(b, 80) - wl.log_mel_spectrogram(b, 80).numpy()).max() <= 1e-4
    assert np.abs(emu.log_mel(b, 80, norm_mode=1) - wl.log_mel_spectrogram(b, 80, per_clip_max=True).numpy()).max() <= 1e-4
    s = (rng.standard_normal(5000) * 3000).astype(np.int16)
    assert np.abs(emu.log_mel(s, 80) - wl.log_mel_spectrogram(s.astype(np.float32) / 32768.0, 80).numpy()).max() <= 1e-4


def _windows_reference(mel, n_frames, content, seek0, stride, n_windows, dtype):
    """transcribe's loop: mel[:, seek:seek + n_frames] limited to the content frames, pad_or_trim with zeros, cast"""
    out = np.zeros((n_windows, mel.shape[0], n_frames), dtype)
    for w in range(n_windows):
        s = seek0 + w * stride
        e = min(s + n_frames, content)
        if e > s:
            out[w, :, :e - s] = mel[:, s:e].astype(dtype)
    return out


def test_emu_mel_windows(emu):
    rng = np.random.default_rng(4)
    mel = (rng.standard_normal((80, 7321 + 3000)) * 0.7).astype(np.float32)
    mel[3, 10] = 65519.0; mel[4, 11] = 6.0e-8; mel[5, 12] = 2049.0 / 2048.0      # f16 rounding edges: max, subnormal, tie
    w = emu.mel_windows(mel)                                                   # transcribe defaults: 3 windows, last one padded
    assert w.shape == (3, 80, 3000) and np.array_equal(w, _windows_reference(mel, 3000, 7321, 0, 3000, 3, np.float32))
    h = emu.mel_windows(mel, half=True)
    assert h.dtype == np.float16 and np.array_equal(h.view(np.uint16), _windows_reference(mel, 3000, 7321, 0, 3000, 3, np.float16).view(np.uint16))
    # odd window length (unpaired stores), overlapping stride, a start offset, windows entirely past the content
    for half in (False, True):
        dt = np.float16 if half else np.float32
        got = emu.mel_windows(mel[:5], n_frames=301, content=1000, seek0=7, stride=150, n_windows=9, half=half)
        assert np.array_equal(got.view(np.uint16 if half else np.uint32),
                              _windows_reference(mel[:5], 301, 1000, 7, 150, 9, dt).view(np.uint16 if half else np.uint32))
    assert emu.mel_windows(mel[:2, :3000]).shape == (0, 2, 3000)               # no content frames: no windows
    with pytest.raises(RuntimeError):
        emu.mel_windows(mel, content=mel.shape[1] + 1)


def test_emu_pipeline(emu):
    from audio_processor_b200 import synth
    x = synth.synth_clip(3, 44100, 2, 7.0, 0.35).numpy()
    r = emu.pipeline(x, 44100, n_mels=80, padding=0)
    full = emu.resample(x, 44100)
    kw = dict(min_silence_len=1000, silence_thresh=-40, keep_silence=200, seek_step=1)
    assert r["nonsilent"] == ps.detect_nonsilent_fast(full, 16000, 1000, -40, 1)
    assert r["kept"] == ps.kept_ranges_fast(full, 16000, **kw)
    trimmed = ps.strip_silence_fast(full, 16000, **kw)
    assert np.array_equal(r["pcm"], trimmed) and 0 < len(trimmed) < len(full)
    assert np.abs(r["mel"] - wl.log_mel_spectrogram(trimmed.astype(np.float32) / 32768.0, 80).numpy()).max() <= 1e-4
    r2 = emu.pipeline(x, 44100, n_mels=128, padding=1600, trim=False)
    assert np.array_equal(r2["pcm"], full)
    assert np.abs(r2["mel"] - wl.log_mel_spectrogram(full.astype(np.float32) / 32768.0, 128, padding=1600).numpy()).max() <= 1e-4


@pytest.mark.parametrize("rate,tile_frames,ch", [(44100, 7056, 2), (48000, 7680, 2), (44100, 7056, 1), (48000, 7680, 1)])
def test_emu_fir_pipeline_many_tiles_per_cta(emu, rate, tile_frames, ch, monkeypatch):
    """the tensor-core FIR's producer/consumer pipeline (mbarrier phases, raw-buffer refills two tiles ahead, deferred
    copy-out) with only 2 persistent CTAs, i.e. 5-6 tiles per CTA — the steady state of the real grid"""
    monkeypatch.setenv("B2A_FIR_GRID", "2")
    rng = np.random.default_rng(rate)
    n = tile_frames * 11 + 999
    x = (rng.standard_normal((n, ch)) * 6000).clip(-32768, 32767).astype(np.int16)
    x = x[:, 0].copy() if ch == 1 else x
    y, en = emu.resample(x, rate, want_energy=True)
    assert len(y) == ro.out_len(n, rate, 16000)
    assert np.abs(y.astype(int) - ro.convert(x, rate).astype(int)).max() <= 1
    if swr_ref.available():
        _cmp16(y, swr_ref.convert(x, rate), 0.998)
    assert np.array_equal(en.astype(np.int64), H.energy_oracle(y))


@pytest.mark.parametrize("rate,grid", [(44100, "4"), (48000, "8")])
def test_emu_fir_tcgen05_tiles(emu, rate, grid, monkeypatch):
    """index logic of the tcgen05 FIR (fir_tmem.cuh) under the functional emulation of tcgen05.mma / TMEM: column ring
    with wrap and mirrored chunk, class tiles (every 4th run, shifted filter banks), accumulator ring, 3 spans over 4-8 persistent
    CTAs, and the hand-over to the mma.sync kernel behind the last span.  (B2A_FIR_GRID exists in the emulation and
    -DB2A_PROFILE builds only; the release library reads no environment variable.)"""
    monkeypatch.setenv("B2A_FIR_GRID", grid)
    S = 441 if rate == 44100 else 480
    rng = np.random.default_rng(rate)
    n = S * 512 * 3 + 12345
    x = (rng.standard_normal((n, 2)) * 6000).clip(-32768, 32767).astype(np.int16)
    y, en = emu.resample(x, rate, want_energy=True)
    assert len(y) == ro.out_len(n, rate, 16000)
    assert np.abs(y.astype(int) - ro.convert(x, rate).astype(int)).max() <= 1
    if swr_ref.available():
        _cmp16(y, swr_ref.convert(x, rate), 0.998)
    assert np.array_equal(en.astype(np.int64), H.energy_oracle(y))


@pytest.mark.parametrize("keep,pad,extra", [(0, 0, 37), (30, 480, 0), (0, 0, 160 * 32 + 5)])
def test_emu_pipeline_fused_compaction_many_short_segments(emu, keep, pad, extra):
    """the log-mel tile loader gathers the kept ranges itself and writes the trimmed PCM (no compaction kernel): bursts of
    120-300 ms put three and more segments under one 5360-sample tile (per-sample segment search), the clip lengths leave
    a partial last hop / a whole extra tile of trimmed samples behind the last frame, with and without right padding"""
    rng = np.random.default_rng(keep + pad + extra)
    parts = []
    for i in range(28):
        n_b = int(rng.integers(120, 300)) * 16
        parts.append((rng.standard_normal(n_b) * 4000).astype(np.int16))
        parts.append((rng.standard_normal(int(rng.integers(140, 260)) * 16) * 3).astype(np.int16))
    parts.append((rng.standard_normal(16 * 400 + extra) * 4000).astype(np.int16))
    x = np.concatenate(parts)
    kw = dict(min_silence_len=100, silence_thresh=-50, keep_silence=keep, seek_step=1)
    r = emu.pipeline(x, 16000, n_mels=80, padding=pad, **kw)
    assert r["kept"] == ps.kept_ranges_fast(x, 16000, **kw) and len(r["kept"]) >= 20
    trimmed = ps.strip_silence_fast(x, 16000, **kw)
    assert np.array_equal(r["pcm"], trimmed)
    ref = wl.log_mel_spectrogram(trimmed.astype(np.float32) / 32768.0, 80, padding=pad).numpy()
    assert r["mel"].shape == ref.shape and np.abs(r["mel"] - ref).max() <= 1e-4


# ---- the tensor-core log-mel PROBE (tools/probes/logmel_tc.cuh): compiled into the emulation library only, selected with
# B2A_LM_IMPL=tc; the release library ships the FFT kernel for every input (profiles/r02_logmel_tc.md) ----
def test_emu_logmel_tc_many_tiles_per_cta(emu, monkeypatch):
    """the tensor-core log-mel probe with 2 persistent CTAs over 7 tiles of 128 frames: operand-ring and
    accumulator barrier parities across tiles, raw-tile reuse, a partial last tile, whole-call maximum"""
    monkeypatch.setenv("B2A_LM_IMPL", "tc")
    monkeypatch.setenv("B2A_LM_GRID", "2")
    rng = np.random.default_rng(5)
    n = 160 * 128 * 6 + 160 * 50 + 77
    t = np.arange(n) / 16000.0
    x = (6000 * np.sin(2 * np.pi * 523.0 * t) * (1 + 0.5 * np.sin(2 * np.pi * 0.7 * t)) + rng.standard_normal(n) * 40).astype(np.int16)
    ref = wl.log_mel_spectrogram(x.astype(np.float32) / 32768.0, 80).numpy()
    got = emu.log_mel(x, 80)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-4
    monkeypatch.setenv("B2A_LM_GRID", "3")
    ref = wl.log_mel_spectrogram(x.astype(np.float32) / 32768.0, 128, padding=160 * 300).numpy()
    got = emu.log_mel(x, 128, padding=160 * 300)
    assert got.shape == ref.shape and np.abs(got - ref).max() <= 1e-4


def test_emu_logmel_tc_batch_unaligned_and_short(emu, monkeypatch):
    monkeypatch.setenv("B2A_LM_IMPL", "tc")
    _logmel_batch_unaligned_and_short(emu)


def test_emu_logmel_s16_batch_unaligned_and_short(emu):
    _logmel_batch_unaligned_and_short(emu)


def _logmel_batch_unaligned_and_short(emu):
    """batches of 16-bit clips (work item = clip x tile, per-clip and whole-call maximum), rows whose stride breaks the
    16-byte alignment of the bulk copies (lanes copy those rows themselves), clips shorter than one tile"""
    rng = np.random.default_rng(6)
    b = (rng.standard_normal((3, 160 * 130 + 3)) * np.array([[3000], [100], [3]])).astype(np.int16)
    f = b.astype(np.float32) / 32768.0
    assert np.abs(emu.log_mel(b, 80) - wl.log_mel_spectrogram(f, 80).numpy()).max() <= 1e-4
    assert np.abs(emu.log_mel(b, 128, norm_mode=1) - wl.log_mel_spectrogram(f, 128, per_clip_max=True).numpy()).max() <= 1e-4
    s = (rng.standard_normal(401) * 2000).astype(np.int16)
    assert np.abs(emu.log_mel(s, 80) - wl.log_mel_spectrogram(s.astype(np.float32) / 32768.0, 80).numpy()).max() <= 1e-4
    assert np.all(emu.log_mel(np.zeros(16000, np.int16), 80) == -1.5)


def test_emu_logmel_tc_tone_over_noise_floor(emu, monkeypatch):
    monkeypatch.setenv("B2A_LM_IMPL", "tc")
    _logmel_tone_over_noise_floor(emu)


def test_emu_logmel_s16_tone_over_noise_floor(emu):
    _logmel_tone_over_noise_floor(emu)


def _logmel_tone_over_noise_floor(emu):
    """the worst probed dynamic range for the f16-plane DFT: a full-scale tone over a +-2 LSB noise floor"""
    rng = np.random.default_rng(7)
    n = 16000 * 3
    x = np.clip(np.rint(32000 * np.sin(2 * np.pi * 1234.5 * np.arange(n) / 16000) + rng.normal(0, 2, n)), -32768, 32767).astype(np.int16)
    for nm in (80, 128):
        assert np.abs(emu.log_mel(x, nm) - wl.log_mel_spectrogram(x.astype(np.float32) / 32768.0, nm).numpy()).max() <= 1e-4


def test_emu_pipeline_tc_gather_window_overflow(emu, monkeypatch):
    monkeypatch.setenv("B2A_LM_IMPL", "tc")
    _pipeline_many_ranges_under_one_tile(emu)


def test_emu_pipeline_many_ranges_under_one_tile(emu):
    _pipeline_many_ranges_under_one_tile(emu)


def _pipeline_many_ranges_under_one_tile(emu):
    """more than 32 kept ranges under one 128-frame tile (40 ms bursts): the producer's cached range window overflows and
    the rows behind it take the per-sample path; result identical to the oracle's trimmed PCM and log-mel"""
    rng = np.random.default_rng(8)
    parts = []
    for i in range(60):
        parts.append((rng.standard_normal(16 * 25) * 5000).astype(np.int16))
        parts.append((rng.standard_normal(16 * 30) * 2).astype(np.int16))
    x = np.concatenate(parts)
    kw = dict(min_silence_len=20, silence_thresh=-50, keep_silence=2, seek_step=1)
    r = emu.pipeline(x, 16000, n_mels=80, padding=0, **kw)
    assert r["kept"] == ps.kept_ranges_fast(x, 16000, **kw) and len(r["kept"]) >= 50
    trimmed = ps.strip_silence_fast(x, 16000, **kw)
    assert np.array_equal(r["pcm"], trimmed)
    ref = wl.log_mel_spectrogram(trimmed.astype(np.float32) / 32768.0, 80).numpy()
    assert r["mel"].shape == ref.shape and np.abs(r["mel"] - ref).max() <= 1e-4


@pytest.mark.parametrize("n", [7056 * 32 + 5000, 7056 * 32 + 3708, 7056 * 16 + 6990])
def test_emu_fir_mono_last_tile_stays_inside_the_clip(emu, n):
    """mono input: the last 16-run tile of the mma.sync kernel must not read past the clip (round 1 sized the tile count with
    stereo's 4 bytes per frame: on lengths up to 3 600 frames behind a tile boundary the tail came out wrong)"""
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 6000).clip(-32768, 32767).astype(np.int16)
    y, en = emu.resample(x, 44100, want_energy=True)
    assert np.abs(y.astype(int) - ro.convert(x, 44100).astype(int)).max() <= 1
    if swr_ref.available():
        _cmp16(y, swr_ref.convert(x, 44100), 0.998)
    assert np.array_equal(en.astype(np.int64), H.energy_oracle(y))


def test_emu_silence_many_blocks_and_sparse_seek(emu):
    """the one-launch silence kernel over several blocks (>= 4 096 ms each): run counts exchanged between blocks, runs that
    cross block boundaries, seek_step larger than min_silence_len (pydub still merges consecutive silent candidates, also
    across a block boundary), windows longer than a block's halo rounding"""
    rng = np.random.default_rng(21)
    for trial, (W, th, keep, step) in enumerate([(100, -40, 50, 300), (250, -30, 0, 1200), (1000, -40, 200, 1), (500, -50, True, 25),
                                                 (100, -40, 100, 4100), (2500, -40, 700, 3)]):
        n = 16 * int(rng.integers(20000, 42000)) + int(rng.integers(0, 16))
        x = H.random_speechlike(rng, n, min_span=800, max_span=60000)
        r = emu.detect(x, 16000, W, th, keep, step)
        assert r["silent"] == ps.detect_silence_fast(x, 16000, W, th, step), (trial, W, th, step)
        assert r["nonsilent"] == ps.detect_nonsilent_fast(x, 16000, W, th, step), (trial, W, th, step)
        assert r["kept"] == ps.kept_ranges_fast(x, 16000, W, th, keep, step), (trial, W, th, keep, step)


@pytest.mark.parametrize("keep,pad,extra", [(0, 0, 37), (30, 480, 0)])
def test_emu_pipeline_tc_fused_compaction(emu, keep, pad, extra, monkeypatch):
    """the probe's gathering fetch + trimmed-PCM write-back on bursts of 120-300 ms (several kept ranges per tile)"""
    monkeypatch.setenv("B2A_LM_IMPL", "tc")
    test_emu_pipeline_fused_compaction_many_short_segments(emu, keep, pad, extra)


def test_emu_pipeline_batch_matches_single_calls(emu):
    """b2a_pipeline_batch: three clips of different length, rate and channel count in ONE call == three b2a_pipeline calls"""
    rng = np.random.default_rng(31)
    clips = []
    for rate, ch, secs in ((44100, 2, 3.1), (16000, 1, 5.0), (48000, 2, 2.3)):
        n = int(rate * secs) + int(rng.integers(0, 50))
        t = np.arange(n) / rate
        env = (np.floor(t / 1.3) % 2 == 0)
        x = (np.where(env, 5000 * np.sin(2 * np.pi * 330 * t), 0)[:, None] + rng.standard_normal((n, ch)) * 5).astype(np.int16)
        clips.append((x[:, 0].copy() if ch == 1 else x, rate))
    kw = dict(min_silence_len=400, silence_thresh=-40, keep_silence=100, seek_step=1)
    got = emu.pipeline_batch(clips, n_mels=80, padding=0, **kw)
    for (pcm, rate), g in zip(clips, got):
        one = emu.pipeline(pcm, rate, n_mels=80, padding=0, **kw)
        assert g["kept"] == one["kept"] and g["nonsilent"] == one["nonsilent"]
        assert np.array_equal(g["pcm"], one["pcm"]) and np.array_equal(g["mel"], one["mel"])
    assert sum(len(g["kept"]) for g in got) >= 5 and all(0 < len(g["pcm"]) for g in got)


def test_emu_remap_times_matches_oracle(emu):
    """b2a_kept_offsets + b2a_remap_times against oracle/remap.py (trimmed-timeline seconds -> original recording)"""
    from oracle import remap as orm
    kept = [[0, 3093], [4907, 8000], [9100, 9900], [15000, 15016]]
    times = [0.0, 0.001, 3.0929, 3.093, 3.0931, 5.0, 6.186, 6.1861, 6.986, 7.0, 7.002, 7.5, 100.0]
    got, koff = emu.remap_times(times, kept)
    assert koff.tolist() == [0, 3093 * 16, 6186 * 16, 6986 * 16, 7002 * 16]
    ref = [orm.remap_time(t, kept) for t in times]
    assert np.abs(got - np.array(ref)).max() <= 1e-9, (got, ref)
    got0, _ = emu.remap_times([1.5], [])
    assert got0.tolist() == [1.5]
