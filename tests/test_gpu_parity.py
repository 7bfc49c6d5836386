"""GPU parity tests proper: the CUDA path (through the C ABI, libb2a.so) against the oracle on the same bytes.
Gates (SURVEY.md A.4 / BASELINE.md §5): silence ranges, keep-mask and compacted PCM bit-exact; 16 kHz PCM <= 1 LSB
from libswresample with >= 99.8 % identical samples and <= 1e-5 (normalised float) from the float64 restatement;
log-mel <= 1e-4 abs from the float64 oracle."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

G = H.golden_json()
A = H.golden_arrays()
MEL_TOL = 1e-4
PCM_TOL = 1e-5


@pytest.fixture(scope="module")
def T():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def ops(T):
    from audio_processor_b200 import _lib, ops as o
    _lib.lib()          # fails loudly if libb2a.so is missing — there is no fallback to test instead
    return o


def _cmp16(y, ref, min_exact=0.998):
    assert len(y) == len(ref), (len(y), len(ref))       # exactly the library's one-shot length
    d = np.abs(y.astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d == 0).mean() >= min_exact, (d.max(), (d == 0).mean())


# ---------------------------------------------------------------- conversion
@pytest.mark.parametrize("name,rate,min_exact", [("R1", 44100, 0.998), ("R2", 48000, 0.998), ("N441", 44100, 0.998),
                                                 ("N480", 48000, 0.998), ("N220", 22050, 0.995)])
def test_resample_golden(ops, name, rate, min_exact):
    x = H.tone_pair(rate, G["resample"][name]["n_in"]) if name.startswith("R") else A[f"{name}_in"]
    y, _, _ = ops.resample(x, rate)
    _cmp16(y.cpu().numpy(), A[f"{name}_out"], min_exact)


@pytest.mark.parametrize("rate,ch,secs", [(44100, 2, 12.3), (44100, 1, 7.7), (48000, 2, 9.1), (48000, 1, 3.3), (32000, 2, 2.0),
                                          (22050, 1, 2.0), (8000, 1, 2.0)])
def test_resample_vs_oracles(ops, T, rate, ch, secs):
    from oracle import resample_oracle as ro, swr_ref
    rng = np.random.default_rng(rate * 3 + ch)
    n = int(rate * secs) + 13
    x = (rng.standard_normal((n, ch)) * 5000).clip(-32768, 32767).astype(np.int16)
    x = x[:, 0].copy() if ch == 1 else x
    y, yf, en = ops.resample(T.from_numpy(x).cuda(), rate, want_f32=True, want_energy=True)
    y, yf, en = y.cpu().numpy(), yf.cpu().numpy(), en.cpu().numpy()
    assert len(y) == ro.out_len(n, rate, 16000)
    assert np.abs(yf - ro.resample_float(x, rate)).max() <= PCM_TOL
    assert np.abs(y.astype(int) - ro.convert(x, rate).astype(int)).max() <= 1
    if swr_ref.available():
        _cmp16(y, swr_ref.convert(x, rate), 0.998 if rate in (44100, 48000) else 0.95)
    assert np.array_equal(en, H.energy_oracle(y))



@pytest.mark.parametrize("rate,spans", [(44100, 3), (48000, 5), (44100, 85), (48000, 120), (44100, 301), (48000, 300)])
def test_resample_tcgen05_tiles(ops, T, rate, spans):
    """the tcgen05 FIR (fir_tmem.cuh): 512-run spans = 4 class tiles of 128 rows; 85 / 120 spans over 37 span lanes give
    3-4 tiles per persistent CTA (column ring wrap, mbarrier phases, accumulator ring) + the mma.sync kernel behind the
    last span — from 300 spans on (25 minutes) those 200 runs are computed by the FIR kernel's spare warps instead
    (fir_dispatch.cu) — against the float64 restatement (<= 1 LSB), the real libswresample (>= 99.8 % identical, SAME
    length) and the exact per-millisecond energies"""
    from oracle import resample_oracle as ro, swr_ref
    S = 441 if rate == 44100 else 480
    rng = np.random.default_rng(rate + spans)
    n = S * 512 * spans + S * 200 + 999
    x = (rng.standard_normal((n, 2)) * 6000).clip(-32768, 32767).astype(np.int16)
    y, _, en = ops.resample(T.from_numpy(x).cuda(), rate, want_energy=True)
    y, en = y.cpu().numpy(), en.cpu().numpy()
    assert len(y) == ro.out_len(n, rate, 16000)
    if swr_ref.available():
        ref = swr_ref.convert(x, rate)
        _cmp16(y, ref, 0.998)
    if spans <= 8:
        assert np.abs(y.astype(int) - ro.convert(x, rate).astype(int)).max() <= 1
    assert np.array_equal(en, H.energy_oracle(y))

@pytest.mark.parametrize("rate,ch", [(44100, 2), (48000, 2), (44100, 1), (22050, 1), (32000, 2), (8000, 1)])
def test_resample_length_sweep_vs_library(ops, T, rate, ch):
    """round-1 VERDICT #1: the output length must be what one-shot swr_convert + flush returns on EVERY input length
    (the old ceil(n L / M) overshot by one on a third of them).  60 consecutive lengths through the real kernels (the
    tcgen05 spans end at different places for each) + short clips around the filter length, against the real library"""
    from oracle import resample_oracle as ro, swr_ref
    if not swr_ref.available():
        pytest.skip("bundled libswresample not found")
    rng = np.random.default_rng(rate + ch)
    taps = ro.n_taps(rate, 16000)
    base = (441 if rate == 44100 else 480) * 512 + 5000
    lengths = list(range(base, base + 60)) + [taps // 2, taps - 1, taps, taps + 1, taps + 7, 2 * taps + 3, 1]
    for n in lengths:
        x = (rng.standard_normal((n, ch)) * 6000).clip(-32768, 32767).astype(np.int16)
        x = x[:, 0].copy() if ch == 1 else x
        ref = swr_ref.convert(x, rate)
        assert ops.resample_out_len(n, rate) == len(ref), (n, len(ref))
        if len(ref) == 0:
            continue
        y, _, en = ops.resample(T.from_numpy(x).cuda(), rate, want_energy=True)
        y = y.cpu().numpy()
        assert len(y) == len(ref), (n, len(y), len(ref))
        d = np.abs(y.astype(int) - ref.astype(int))
        assert d.max() <= 1, (n, d.max())
        assert np.array_equal(en.cpu().numpy(), H.energy_oracle(y))


def test_resample_float_input_same_rate_and_unaligned(ops, T):
    from oracle import resample_oracle as ro
    rng = np.random.default_rng(5)
    xf = (rng.standard_normal((30000, 2)) * 0.2).astype(np.float32)
    y, _, _ = ops.resample(xf, 44100)
    assert np.abs(y.cpu().numpy().astype(int) - ro.convert(xf, 44100).astype(int)).max() <= 1
    x = (rng.standard_normal((40001, 2)) * 9000).clip(-32768, 32767).astype(np.int16)
    y, _, en = ops.resample(x, 16000, want_energy=True)
    assert np.array_equal(y.cpu().numpy(), ro.convert(x, 16000))                       # (L+R+1)>>1, bit exact
    assert np.array_equal(en.cpu().numpy(), H.energy_oracle(y.cpu().numpy()))
    y, _, _ = ops.resample(x[:, 0].copy(), 16000)
    assert np.array_equal(y.cpu().numpy(), x[:, 0])                                     # identity
    # a 4-byte-aligned (not 16-byte) device pointer must take the table-driven kernel and still be right
    big = T.from_numpy((rng.standard_normal((14112 * 3 + 1, 2)) * 4000).astype(np.int16)).cuda()
    view = big[1:]
    y2, _, _ = ops.resample(view, 44100)
    assert np.abs(y2.cpu().numpy().astype(int) - ro.convert(view.cpu().numpy(), 44100).astype(int)).max() <= 1


def test_resample_linearity_full_minute(ops, T):
    """size-independent property at a BASELINE size (60 s): FIR is linear, so R(a)+R(b) ~= R(a+b) before rounding"""
    from audio_processor_b200 import synth
    a = synth.synth_clip(1, 44100, 2, 60.0, 0.25, device="cuda")
    b = synth.synth_clip(2, 44100, 2, 60.0, 0.25, device="cuda")
    half = lambda t: (t.to(T.int32) // 2).to(T.int16)
    _, fa, _ = ops.resample(half(a), 44100, want_s16=False, want_f32=True)
    _, fb, _ = ops.resample(half(b), 44100, want_s16=False, want_f32=True)
    _, fs, _ = ops.resample(half(a) + half(b), 44100, want_s16=False, want_f32=True)
    assert fs.shape[0] == 960000
    assert float((fa + fb - fs).abs().max()) <= 2e-6


# ---------------------------------------------------------------- silence
@pytest.mark.parametrize("name", sorted(G["silence"].keys()))
def test_silence_golden(ops, name):
    g = G["silence"][name]
    x = H.piecewise([tuple(p) for p in g["parts"]], g["extra"])
    W, th, keep, step = g["params"]
    r = ops.detect(x, 16000, W, th, keep, step)
    assert r.silent == g["silent"] and r.nonsilent == g["nonsilent"] and r.kept == g["kept"]
    assert r.len_ms == g["len_ms"] and r.n_keep == g["n_keep"]


def test_silence_random_bit_exact(ops, T):
    from oracle import pydub_silence as ps
    rng = np.random.default_rng(13)
    for trial in range(40):
        n = int(rng.integers(20000, 900000)) + int(rng.integers(0, 16))
        x = H.random_speechlike(rng, n, min_span=800, max_span=60000)
        W = int(rng.choice([100, 250, 500, 1000, 2500]))
        th = float(rng.choice([-16, -30, -40, -50, -60.5]))
        keep = [0, 50, 100, 200, 700, True, False][trial % 7]
        step = int(rng.choice([1, 1, 1, 3, 10, 25, 300, 1200]))
        d = T.from_numpy(x).cuda()
        r = ops.detect(d, 16000, W, th, keep, step)
        assert r.silent == ps.detect_silence_fast(x, 16000, W, th, step), (trial, W, th, step)
        assert r.nonsilent == ps.detect_nonsilent_fast(x, 16000, W, th, step), (trial, W, th, step)
        assert r.kept == ps.kept_ranges_fast(x, 16000, W, th, keep, step), (trial, W, th, keep, step)
        out = ops.compact(d, r)[: r.n_keep].cpu().numpy()
        st = ps.strip_silence_fast(x, 16000, min_silence_len=W, silence_thresh=th, keep_silence=keep, seek_step=step)
        assert np.array_equal(out, st)


def test_silence_literal_pydub_loop_small(ops):
    """the oracle of record (one audioop.rms per window) on a short clip"""
    from oracle import pydub_silence as ps
    rng = np.random.default_rng(3)
    x = H.random_speechlike(rng, 16000 * 6 + 9, min_span=3000, max_span=30000)
    s = ps.Segment(x)
    r = ops.detect(x, 16000, 500, -35, 150, 1)
    assert r.silent == ps.detect_silence(s, 500, -35, 1)
    assert r.nonsilent == ps.detect_nonsilent(s, 500, -35, 1)
    assert r.kept == ps.kept_ranges(s, 500, -35, 150, 1)


def test_silence_edges_and_api(ops, T):
    from audio_processor_b200 import silence as S
    r = ops.detect(np.zeros(5, dtype=np.int16))
    assert r.nonsilent == [[0, 0]] and r.n_keep == 0
    x = np.full(16000 * 3, 9000, dtype=np.int16)
    r = ops.detect(x, 16000, 1000, -40, 100, 1)
    assert r.silent == [] and r.nonsilent == [[0, 3000]] and r.n_keep == len(x)
    k1 = H.piecewise([(3, 1000), (2, 0), (3, 1000)])
    seg = S.AudioSegment(k1)
    assert S.detect_silence(seg, 1000, -40) == [[2893, 5107]]
    assert S.detect_nonsilent(seg, 1000, -40) == [[0, 2893], [5107, 8000]]
    chunks = S.split_on_silence(seg, 1000, -40, keep_silence=200)
    assert [len(c) for c in chunks] == [3093, 3093]
    st = S.strip_silence(seg, 1000, -40, keep_silence=200)
    assert np.array_equal(st.get_array_of_samples(), np.concatenate([c.get_array_of_samples() for c in chunks]))
    with pytest.raises(RuntimeError):
        ops.detect(x, 44100)                       # 44.1 samples per ms: unsupported, loudly


def test_silence_long_clip_several_rounds_of_chunks(ops, T):
    """clips beyond about an hour take several rounds of chunks per block (the prefix sums of a chunk live in shared memory):
    2.6 h of 16 kHz audio against the vectorised oracle, plus a 10-second window (the largest supported, smallest chunks)"""
    from oracle import pydub_silence as ps
    rng = np.random.default_rng(77)
    n = int(2.6 * 3600 * 16000) + 5
    x = H.random_speechlike(rng, n, min_span=8000, max_span=400000)
    d = T.from_numpy(x).cuda()
    for W, th, keep, step in [(1000, -40.0, 200, 1), (10000, -35.0, 700, 1), (700, -40.0, 100, 7)]:
        r = ops.detect(d, 16000, W, th, keep, step)
        assert r.silent == ps.detect_silence_fast(x, 16000, W, th, step), (W, th, step)
        assert r.nonsilent == ps.detect_nonsilent_fast(x, 16000, W, th, step), (W, th, step)
        assert r.kept == ps.kept_ranges_fast(x, 16000, W, th, keep, step), (W, th, keep, step)


def test_silence_idempotent_full_hour(ops, T):
    """size-independent properties at the cfg2 size (1 h @ 16 kHz): sorted, disjoint, inside the clip, and
    trimming the trimmed audio with keep_silence >= min_silence_len/2 removes (almost) nothing more."""
    from audio_processor_b200 import synth
    x = synth.synth_clip(2, 16000, 1, 3600.0, 0.2, device="cuda")
    r = ops.detect(x, 16000, 1000, -40, 200, 1)
    kept = np.asarray(r.kept)
    assert len(kept) > 100 and (kept[:, 0] < kept[:, 1]).all() and (kept[1:, 0] >= kept[:-1, 1]).all()
    assert kept[0, 0] >= 0 and kept[-1, 1] <= r.len_ms == 3600000
    assert r.n_keep == int(((kept[:, 1] - kept[:, 0]) * 16).sum())
    out = ops.compact(x, r)[: r.n_keep]
    # checksum of checksums: compacted audio is exactly the concatenation of the kept slices
    xs = x.to(T.int64)
    csum = T.cumsum(xs * xs, 0)
    tot = sum(int(csum[e * 16 - 1] - (csum[s * 16 - 1] if s else 0)) for s, e in kept.tolist())
    assert tot == int((out.to(T.int64) ** 2).sum())


# ---------------------------------------------------------------- log-mel
@pytest.mark.parametrize("nm", [80, 128])
def test_logmel_golden(ops, nm):
    t = np.arange(16000) / 16000.0
    tone = (0.5 * np.sin(2 * np.pi * 1000 * t)).astype(np.float32)
    m = ops.log_mel(tone, nm).cpu().numpy()
    assert np.abs(m - A[f"mel_tone_{nm}"]).max() <= MEL_TOL
    g = G["logmel"][f"M{nm}"]
    assert abs(m.max() - g["max"]) <= MEL_TOL and int(m.argmax() // m.shape[1]) == g["argmax_mel"]
    mn = ops.log_mel(A["mel_noise_in"], nm, padding=480).cpu().numpy()
    assert np.abs(mn - A[f"mel_noise_{nm}_pad480"]).max() <= MEL_TOL


def test_logmel_cases(ops, T):
    from oracle import whisper_logmel as wl
    rng = np.random.default_rng(2)
    t = np.arange(16000 * 2) / 16000.0
    hd = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.5 * 10 ** (-70 / 20) * rng.standard_normal(len(t))).astype(np.float32)
    assert np.abs(ops.log_mel(hd, 80).cpu().numpy() - wl.log_mel_spectrogram(hd, 80).numpy()).max() <= MEL_TOL
    assert bool((ops.log_mel(np.zeros(16000, np.float32), 80) == -1.5).all())
    click = np.zeros(8000, np.float32); click[4000] = 1.0
    assert np.abs(ops.log_mel(click, 128, padding=4000).cpu().numpy() - wl.log_mel_spectrogram(click, 128, padding=4000).numpy()).max() <= MEL_TOL
    b = (rng.standard_normal((5, 48000)) * np.array([[0.3], [0.01], [0.0003], [0.1], [0.9]])).astype(np.float32)
    assert np.abs(ops.log_mel(b, 80).cpu().numpy() - wl.log_mel_spectrogram(b, 80).numpy()).max() <= MEL_TOL
    assert np.abs(ops.log_mel(b, 128, per_clip_max=True).cpu().numpy() - wl.log_mel_spectrogram(b, 128, per_clip_max=True).numpy()).max() <= MEL_TOL
    s = (rng.standard_normal(60 * 16000) * 3000).astype(np.int16)                        # cfg1 shape: 60 s -> [80, 6000]
    m = ops.log_mel(s, 80).cpu().numpy()
    assert m.shape == (80, 6000)
    assert np.abs(m - wl.log_mel_spectrogram(s.astype(np.float32) / 32768.0, 80).numpy()).max() <= MEL_TOL
    m = ops.log_mel(s, 80, padding=480000).cpu().numpy()                                    # transcribe's call
    assert m.shape == (80, 9000)
    assert np.abs(m - wl.log_mel_spectrogram(s.astype(np.float32) / 32768.0, 80, padding=480000).numpy()).max() <= MEL_TOL


def test_logmel_whisper_api(ops, T, tmp_path):
    from audio_processor_b200 import wavio, whisper_audio as wa
    from oracle import whisper_logmel as wl
    rng = np.random.default_rng(9)
    x = (rng.standard_normal(16000 * 3) * 2500).astype(np.int16)
    p = str(tmp_path / "c.wav")
    wavio.write_wav_s16(p, x, 16000)
    a = wa.load_audio(p)
    assert a.dtype == np.float32 and np.array_equal(a, x.astype(np.float32) / 32768.0)
    m = wa.log_mel_spectrogram(p, n_mels=80, padding=wa.N_SAMPLES)
    assert tuple(m.shape) == (80, (len(x) + wa.N_SAMPLES) // 160)
    assert np.abs(m.cpu().numpy() - wl.log_mel_spectrogram(a, 80, padding=wa.N_SAMPLES).numpy()).max() <= MEL_TOL
    assert wa.pad_or_trim(T.zeros(10), 4).shape == (4,) and wa.pad_or_trim(np.zeros((2, 3)), 5).shape == (2, 5)
    assert np.abs(wa.mel_filters(None, 80).numpy() - wl.mel_filters(80)).max() <= 1e-9


def test_mel_windows_match_transcribe_loop(ops, T):
    """b2a_mel_windows == whisper.transcribe's slice + pad_or_trim + cast, window by window (bit-exact, f32 and f16), against
    the numpy restatement in oracle/whisper_logmel.py (encoder_windows) — and the package's own window iterator agrees"""
    from audio_processor_b200 import whisper_audio as wa
    from oracle import whisper_logmel as wl
    g = T.Generator(device="cuda").manual_seed(5)
    audio = (T.randn(16000 * 73, device="cuda", generator=g) * 0.1).clamp(-1, 1)
    mel = wa.log_mel_spectrogram(audio, 80, padding=wa.N_SAMPLES)
    mel_h = mel.cpu().numpy()
    for dtype, npdt in ((T.float32, np.float32), (T.float16, np.float16)):
        ref = wl.encoder_windows(mel_h, dtype=npdt)
        got = wa.mel_windows(mel, dtype=dtype)
        assert got.shape == ref.shape == (3, 80, 3000) and got.dtype == dtype
        assert np.array_equal(got.cpu().numpy(), ref)
        for w, (_, seg) in enumerate(wa.mel_segments(mel, dtype=dtype)):
            assert np.array_equal(seg.cpu().numpy(), ref[w])
    # overlapping grid with an odd window length and windows past the content
    got = ops.mel_windows(mel, 301, content_frames=1000, seek0=7, stride=150, n_windows=9, dtype=T.float16)
    ref = wl.encoder_windows(mel_h, 301, content_frames=1000, seek0=7, stride=150, n_windows=9, dtype=np.float16)
    assert np.array_equal(got.cpu().numpy(), ref)
    with pytest.raises(RuntimeError):
        ops.mel_windows(mel, content_frames=mel.shape[1] + 1)


def test_logmel_batch_4096_property(ops, T):
    """cfg3 shape [4096, 480000] f32 -> [4096, 128, 3000]; spot-check rows against the oracle and the floor invariant"""
    from audio_processor_b200 import synth
    from oracle import whisper_logmel as wl
    x = synth.noise_batch(3, 4096, 480000, device="cuda")
    m = ops.log_mel(x, 128)
    assert tuple(m.shape) == (4096, 128, 3000)
    gmax = float(m.max())
    assert abs(float(m.min()) - max(float(m.min()), gmax - 2.0)) < 1e-6         # nothing below the floor (max-8)/4
    rows = [0, 1777, 4095]
    sub = x[rows].cpu().numpy()
    ref = wl.log_mel_spectrogram(sub, 128).numpy()
    # the floor of the full batch can only be >= the floor of 3 rows; noise never reaches it, so rows agree exactly in form
    assert np.abs(m[rows].cpu().numpy() - ref).max() <= MEL_TOL


# ---------------------------------------------------------------- whole path
@pytest.mark.parametrize("rate,ch,secs,nm,pad", [(44100, 2, 20.0, 80, 0), (48000, 2, 15.0, 80, 0), (16000, 1, 60.0, 80, 0),
                                                 (44100, 1, 9.0, 128, 480000)])
def test_pipeline_vs_oracle(ops, T, rate, ch, secs, nm, pad):
    from audio_processor_b200 import synth
    from oracle import pydub_silence as ps, whisper_logmel as wl
    x = synth.synth_clip(rate + ch, rate, ch, secs, 0.35, device="cuda")
    r = ops.pipeline(x, rate, n_mels=nm, padding=pad, min_silence_len=1000, silence_thresh=-40, keep_silence=200)
    full, _, _ = ops.resample(x, rate)
    full = full.cpu().numpy()
    kw = dict(min_silence_len=1000, silence_thresh=-40, keep_silence=200, seek_step=1)
    assert r.nonsilent == ps.detect_nonsilent_fast(full, 16000, 1000, -40, 1)
    assert r.kept == ps.kept_ranges_fast(full, 16000, **kw)
    trimmed = ps.strip_silence_fast(full, 16000, **kw)
    assert np.array_equal(r.pcm.cpu().numpy(), trimmed) and 0 < len(trimmed) < len(full)
    ref = wl.log_mel_spectrogram(trimmed.astype(np.float32) / 32768.0, nm, padding=pad).numpy()
    assert tuple(r.mel.shape) == ref.shape and np.abs(r.mel.cpu().numpy() - ref).max() <= MEL_TOL


def test_pipeline_tail_millisecond_follows_library_length(ops, T):
    """round-1 VERDICT #1: lengths where ceil(n L / M) was one sample longer than the library AND that sample moves
    F mod 16 across 8, i.e. pydub's len_ms = round(1000 F / 16000) and with it the last kept range.  The kept table, the
    trimmed PCM and the mel must equal the oracle run on the LIBRARY's 16 kHz PCM."""
    from oracle import pydub_silence as ps, resample_oracle as ro, swr_ref, whisper_logmel as wl
    if not swr_ref.available():
        pytest.skip("bundled libswresample not found")
    rng = np.random.default_rng(77)
    found = 0
    for rate, L, M in ((48000, 1, 3), (44100, 160, 441)):
        n0 = rate * 6
        for n in range(n0, n0 + 2000):
            old = -((-n * L) // M)
            new = ro.out_len(n, rate, 16000)
            if old == new or ps.len_ms(new, 16000) == ps.len_ms(old, 16000):     # need: the extra sample changed pydub's clip length
                continue
            found += 1
            t = np.arange(n) / rate
            x = (rng.standard_normal((n, 2)) * 40).astype(np.int16)
            x[: n // 3] += (5000 * np.sin(2 * np.pi * 300 * t[: n // 3]))[:, None].astype(np.int16)
            x[-n // 4:] += (5000 * np.sin(2 * np.pi * 500 * t[-n // 4:]))[:, None].astype(np.int16)     # loud up to the last sample
            lib16 = swr_ref.convert(x, rate)
            assert len(lib16) == new
            kw = dict(min_silence_len=1000, silence_thresh=-40, keep_silence=200, seek_step=1)
            r = ops.pipeline(T.from_numpy(x).cuda(), rate, n_mels=80, padding=0, **kw)
            mine16 = ops.resample(T.from_numpy(x).cuda(), rate)[0].cpu().numpy()
            assert len(mine16) == len(lib16) and np.abs(mine16.astype(int) - lib16.astype(int)).max() <= 1
            assert ps.len_ms(len(lib16), 16000) != ps.len_ms(old, 16000)                 # the old length changed pydub's clip length
            assert r.kept == ps.kept_ranges_fast(mine16, 16000, **kw) and r.kept[-1][1] == ps.len_ms(len(lib16), 16000)
            trimmed = ps.strip_silence_fast(mine16, 16000, **kw)
            assert np.array_equal(r.pcm.cpu().numpy(), trimmed)
            ref = wl.log_mel_spectrogram(trimmed.astype(np.float32) / 32768.0, 80).numpy()
            assert np.abs(r.mel.cpu().numpy() - ref).max() <= MEL_TOL
            break
    assert found >= 2


def test_logmel_s16_cases(ops, T):
    """16-bit input (what the pipeline feeds: Whisper reads the WAV back as int16 / 32768): full-scale tone over a +-2 LSB
    noise floor (the worst probed dynamic range), hundreds of tiles per persistent CTA with a partial last tile, batches
    with whole-call and per-clip maxima, rows off the 16-byte grid, a clip shorter than one tile, zeros, transcribe's
    30-second zero padding"""
    from oracle import whisper_logmel as wl
    rng = np.random.default_rng(7)
    n = 16000 * 3
    x = np.clip(np.rint(32000 * np.sin(2 * np.pi * 1234.5 * np.arange(n) / 16000) + rng.normal(0, 2, n)), -32768, 32767).astype(np.int16)
    for nm in (80, 128):
        got = ops.log_mel(T.from_numpy(x).cuda(), nm).cpu().numpy()
        assert np.abs(got - wl.log_mel_spectrogram(x.astype(np.float32) / 32768.0, nm).numpy()).max() <= MEL_TOL
    n = 160 * 128 * 400 + 160 * 50 + 77                      # 401 tiles: 2-3 per CTA
    t = np.arange(n) / 16000.0
    x = (6000 * np.sin(2 * np.pi * 523.0 * t) * (1 + 0.5 * np.sin(2 * np.pi * 0.7 * t)) + rng.standard_normal(n) * 40).astype(np.int16)
    for nm, pad in ((80, 0), (128, 480000)):
        ref = wl.log_mel_spectrogram(x.astype(np.float32) / 32768.0, nm, padding=pad).numpy()
        got = ops.log_mel(T.from_numpy(x).cuda(), nm, padding=pad).cpu().numpy()
        assert got.shape == ref.shape and np.abs(got - ref).max() <= MEL_TOL
    b = (rng.standard_normal((5, 160 * 130 + 3)) * np.array([[3000], [100], [3], [900], [30]])).astype(np.int16)
    f = b.astype(np.float32) / 32768.0
    assert np.abs(ops.log_mel(T.from_numpy(b).cuda(), 80).cpu().numpy() - wl.log_mel_spectrogram(f, 80).numpy()).max() <= MEL_TOL
    assert np.abs(ops.log_mel(T.from_numpy(b).cuda(), 128, per_clip_max=True).cpu().numpy()
                  - wl.log_mel_spectrogram(f, 128, per_clip_max=True).numpy()).max() <= MEL_TOL
    s = (rng.standard_normal(401) * 2000).astype(np.int16)
    assert np.abs(ops.log_mel(T.from_numpy(s).cuda(), 80).cpu().numpy() - wl.log_mel_spectrogram(s.astype(np.float32) / 32768.0, 80).numpy()).max() <= MEL_TOL
    odd = T.from_numpy(np.concatenate([np.zeros(3, np.int16), s])).cuda()[3:]      # 6 bytes off the 16-byte grid
    assert np.abs(ops.log_mel(odd, 80).cpu().numpy() - wl.log_mel_spectrogram(s.astype(np.float32) / 32768.0, 80).numpy()).max() <= MEL_TOL
    assert np.all(ops.log_mel(T.zeros(16000, dtype=T.int16, device="cuda"), 80).cpu().numpy() == -1.5)


def test_pipeline_many_ranges_under_one_tile(ops, T):
    """hundreds of 25 ms bursts: many kept ranges under every log-mel tile (the gathering tile loader's per-sample path)"""
    from oracle import pydub_silence as ps, whisper_logmel as wl
    rng = np.random.default_rng(8)
    parts = []
    for i in range(400):
        parts.append((rng.standard_normal(16 * 25) * 5000).astype(np.int16))
        parts.append((rng.standard_normal(16 * 30) * 2).astype(np.int16))
    x = np.concatenate(parts)
    kw = dict(min_silence_len=20, silence_thresh=-50, keep_silence=2, seek_step=1)
    r = ops.pipeline(T.from_numpy(x).cuda(), 16000, n_mels=80, padding=0, **kw)
    assert r.kept == ps.kept_ranges_fast(x, 16000, **kw) and len(r.kept) >= 350
    trimmed = ps.strip_silence_fast(x, 16000, **kw)
    assert np.array_equal(r.pcm.cpu().numpy(), trimmed)
    ref = wl.log_mel_spectrogram(trimmed.astype(np.float32) / 32768.0, 80).numpy()
    assert tuple(r.mel.shape) == ref.shape and np.abs(r.mel.cpu().numpy() - ref).max() <= MEL_TOL


@pytest.mark.parametrize("keep,pad,extra", [(0, 0, 37), (30, 480, 0), (0, 0, 160 * 32 + 5)])
def test_pipeline_fused_compaction_many_short_segments(ops, T, keep, pad, extra):
    """the log-mel tile loader gathers the kept ranges itself and writes the trimmed PCM (no compaction kernel): three and
    more segments under one tile, partial last hop / an extra tile of trimmed samples behind the last frame, padding"""
    from oracle import pydub_silence as ps, whisper_logmel as wl
    rng = np.random.default_rng(keep + pad + extra)
    parts = []
    for i in range(28):
        parts.append((rng.standard_normal(int(rng.integers(120, 300)) * 16) * 4000).astype(np.int16))
        parts.append((rng.standard_normal(int(rng.integers(140, 260)) * 16) * 3).astype(np.int16))
    parts.append((rng.standard_normal(16 * 400 + extra) * 4000).astype(np.int16))
    x = np.concatenate(parts)
    kw = dict(min_silence_len=100, silence_thresh=-50, keep_silence=keep, seek_step=1)
    r = ops.pipeline(T.from_numpy(x).cuda(), 16000, n_mels=80, padding=pad, **kw)
    assert r.kept == ps.kept_ranges_fast(x, 16000, **kw) and len(r.kept) >= 20
    trimmed = ps.strip_silence_fast(x, 16000, **kw)
    assert np.array_equal(r.pcm.cpu().numpy(), trimmed)
    ref = wl.log_mel_spectrogram(trimmed.astype(np.float32) / 32768.0, 80, padding=pad).numpy()
    assert tuple(r.mel.shape) == ref.shape and np.abs(r.mel.cpu().numpy() - ref).max() <= MEL_TOL


def test_pipeline_graph_replay_matches_direct_calls(ops, T):
    """PipelinePlan.run(graph=True): first use eager, second captures, later ones replay one CUDA graph of the same
    b2a_pipeline call — same bytes out as the direct call, and the launch counter still sees every kernel"""
    from audio_processor_b200 import synth
    x = synth.synth_clip(5, 44100, 2, 9.0, 0.3, device="cuda")
    kw = dict(min_silence_len=1000, silence_thresh=-40, keep_silence=200)
    ref = ops.PipelinePlan(int(x.shape[0]), 44100, 2, x.dtype, n_mels=80).run(x, **kw)
    pcm0, mel0, kept0 = ref.pcm.clone(), ref.mel.clone(), ref.kept
    plan = ops.PipelinePlan(int(x.shape[0]), 44100, 2, x.dtype, n_mels=80)
    per_call = None
    for i in range(4):
        n0 = ops.launch_count()
        plan.pcm.zero_(); plan.mel.zero_()
        r = plan.run(x, graph=True, **kw)
        T.cuda.synchronize()
        n = ops.launch_count() - n0
        per_call = n if per_call is None else per_call
        assert n == per_call and n >= 5
        assert r.kept == kept0 and T.equal(r.pcm, pcm0) and T.equal(r.mel, mel0)
    assert len(plan._graphs) == 1


def test_pipeline_notrim_and_service(ops, T, tmp_path):
    from audio_processor_b200 import synth, wavio
    from audio_processor_b200.service import AudioFrontend
    from oracle import pydub_silence as ps, resample_oracle as ro, whisper_logmel as wl
    x = synth.synth_clip(5, 44100, 2, 12.0, 0.3, device="cuda")
    r = ops.pipeline(x, 44100, n_mels=80, trim=False)
    full, _, _ = ops.resample(x, 44100)
    assert np.array_equal(r.pcm.cpu().numpy(), full.cpu().numpy())
    ref = wl.log_mel_spectrogram(full.cpu().numpy().astype(np.float32) / 32768.0, 80).numpy()
    assert np.abs(r.mel.cpu().numpy() - ref).max() <= MEL_TOL
    # the reference's helpers, on files
    src = str(tmp_path / "meeting.wav.orig")
    wavio.write_wav_s16(src, x.cpu().numpy(), 44100)
    fe = AudioFrontend()
    wav = fe.convert_to_wav(src)
    assert wav == str(tmp_path / "meeting.wav.wav") or wav.endswith(".wav")
    y, sr = wavio.read_wav(wav)
    assert sr == 16000 and y.ndim == 1 and np.abs(y.astype(int) - ro.convert(x.cpu().numpy(), 44100).astype(int)).max() <= 1
    out = fe.preprocess_audio(wav)
    assert out != wav and out.endswith(".wav")
    z, _ = wavio.read_wav(out)
    assert np.array_equal(z, ps.strip_silence_fast(y, 16000, min_silence_len=1000, silence_thresh=-40, keep_silence=200, seek_step=1))
    assert fe.last_segments == ps.kept_ranges_fast(y, 16000, 1000, -40, 200, 1)
    import subprocess
    bad = tmp_path / "x.m4a"; bad.write_bytes(b"not audio")
    with pytest.raises(subprocess.CalledProcessError):
        fe.convert_to_wav(str(bad))


def test_pipeline_full_hour_properties(ops, T):
    """cfg2 at full size: 1 h 44.1 kHz stereo.  Oracle-free invariants + a sampled window against the oracle."""
    from audio_processor_b200 import synth
    from oracle import pydub_silence as ps, resample_oracle as ro, whisper_logmel as wl
    x = synth.synth_clip(2, 44100, 2, 3600.0, 0.2, device="cuda")
    plan = ops.PipelinePlan(x.shape[0], 44100, 2, x.dtype, n_mels=80)
    r = plan.run(x, min_silence_len=1000, silence_thresh=-40, keep_silence=200)
    kept = np.asarray(r.kept)
    assert (kept[:, 0] < kept[:, 1]).all() and (kept[1:, 0] >= kept[:-1, 1]).all() and kept[-1, 1] <= 3600000
    assert r.n_keep == int(((kept[:, 1] - kept[:, 0]) * 16).sum()) and r.n_frames == r.n_keep // 160
    mel = r.mel
    assert bool(T.isfinite(mel).all()) and float(mel.min()) >= float(mel.max()) - 2.0 - 1e-6
    # sampled window: first 30 s of input (head incl. reflect) against libswresample-equivalent oracle
    head = x[: 44100 * 30].cpu().numpy()
    full, _, _ = ops.resample(x, 44100)
    o = ro.convert(head, 44100)
    d = np.abs(full[: len(o) - 200].cpu().numpy().astype(int) - o[:-200].astype(int))
    assert d.max() <= 1 and (d == 0).mean() >= 0.998
    # ranges recomputed by the oracle from the GPU's own 16 kHz PCM: bit exact over the whole hour
    f16 = full.cpu().numpy()
    assert r.kept == ps.kept_ranges_fast(f16, 16000, 1000, -40, 200, 1)
    trimmed = ps.strip_silence_fast(f16, 16000, min_silence_len=1000, silence_thresh=-40, keep_silence=200, seek_step=1)
    assert np.array_equal(r.pcm.cpu().numpy(), trimmed)
    # mel of a middle slice vs the oracle on the same trimmed samples.  seg = trimmed[160a:160b] => oracle frame j is
    # centred on trimmed sample 160(a+j), i.e. it IS frame a+j of the full clip except near the slice edges (reflect).
    a, b = 100000, 103000
    seg = trimmed[a * 160: b * 160].astype(np.float32) / 32768.0
    refm = wl.log_mel_spectrogram(seg, 80).numpy()[:, 2:-2]
    sub = mel[:, a + 2: b - 2].cpu().numpy()
    assert sub.shape == refm.shape
    # the two floors differ (slice max vs whole-clip max): compare where neither side is clamped
    mask = (refm > refm.max() - 1.9) & (sub > float(mel.max()) - 1.9)
    assert mask.mean() > 0.5 and np.abs(sub - refm)[mask].max() <= MEL_TOL


def test_clip_stream_pipelined_api(ops, T):
    """ClipStream.submit()/result(): same results as the one-shot pipeline, with two clips in flight"""
    from audio_processor_b200 import synth
    from audio_processor_b200.service import AudioFrontend
    fe = AudioFrontend(min_silence_len=1000, silence_thresh=-40, keep_silence=200)
    clips = [synth.synth_clip(40 + i, 44100, 2, 12.0, 0.3, device="cpu") for i in range(4)]
    cs = fe.stream(int(clips[0].shape[0]), 44100, 2, clips[0].dtype, depth=2)
    got, pending = [], []
    for c in clips:
        pending.append(cs.submit(c.pin_memory()))
        if len(pending) > 1:
            pcm, mel, kept = cs.result(pending.pop(0))
            got.append((pcm.clone(), mel.clone(), kept))
    while pending:
        pcm, mel, kept = cs.result(pending.pop(0))
        got.append((pcm.clone(), mel.clone(), kept))
    for c, (pcm, mel, kept) in zip(clips, got):
        r = ops.pipeline(c.cuda(), 44100, n_mels=80, min_silence_len=1000, silence_thresh=-40, keep_silence=200)
        assert kept == r.kept and T.equal(pcm, r.pcm.cpu()) and T.equal(mel, r.mel.cpu())
    with pytest.raises(RuntimeError):
        cs.submit(clips[0]); cs.submit(clips[1]); cs.submit(clips[2])      # third clip would overwrite an uncollected slot


def test_pipeline_degenerate_clips(ops, T):
    """all-silent clip (nothing kept, empty mel), and a clip shorter than one FIR tile (table-driven kernel only)"""
    from oracle import pydub_silence as ps, resample_oracle as ro, whisper_logmel as wl
    quiet = T.zeros((44100 * 3, 2), dtype=T.int16, device="cuda")
    r = ops.pipeline(quiet, 44100, n_mels=80, min_silence_len=1000, silence_thresh=-40, keep_silence=200)
    assert r.kept == [] and r.nonsilent == [] and r.n_keep == 0 and r.n_frames == 0 and tuple(r.mel.shape) == (80, 0)
    rng = np.random.default_rng(17)
    short = (rng.standard_normal((44100 // 2, 2)) * 4000).astype(np.int16)          # 0.5 s: below min_silence_len, all kept
    r = ops.pipeline(short, 44100, n_mels=80, min_silence_len=1000, silence_thresh=-40, keep_silence=200)
    y = ro.convert(short, 44100)
    pcm = r.pcm.cpu().numpy()
    assert r.kept == ps.kept_ranges_fast(pcm if len(pcm) == len(y) else y, 16000, 1000, -40, 200, 1)
    assert abs(len(pcm) - len(y)) <= 16 and np.abs(pcm[:len(y) - 16].astype(int) - y[:len(y) - 16].astype(int)).max() <= 1
    ref = wl.log_mel_spectrogram(pcm.astype(np.float32) / 32768.0, 80).numpy()
    assert np.abs(r.mel.cpu().numpy() - ref).max() <= MEL_TOL


def test_second_device_in_one_process(ops, T):
    """per-device table caches and kernel attributes: the same process drives cuda:1 after cuda:0 (needs 2 GPUs)"""
    if T.cuda.device_count() < 2:
        pytest.skip("single-GPU box")
    from audio_processor_b200 import synth
    x0 = synth.synth_clip(7, 44100, 2, 10.0, 0.3, device="cuda:0")
    r0 = ops.pipeline(x0, 44100, n_mels=80, min_silence_len=1000, silence_thresh=-40, keep_silence=200)
    with T.cuda.device(1):
        x1 = x0.to("cuda:1")
        r1 = ops.pipeline(x1, 44100, n_mels=80, min_silence_len=1000, silence_thresh=-40, keep_silence=200)
        assert r1.kept == r0.kept and T.equal(r1.pcm.cpu(), r0.pcm.cpu()) and T.equal(r1.mel.cpu(), r0.mel.cpu())


# ---------------------------------------------------------------- batches, threads, remap
def _cfg4_like_clip(T, seed, secs):
    from audio_processor_b200 import synth
    return synth.synth_clip(seed, 48000, 2, secs, 0.50, device="cuda")


def test_pipeline_batch_cfg4_shape_vs_oracle(ops, T):
    """BASELINE configs[3] in miniature: 48 kHz stereo clips with ~50 % silence and DIFFERENT lengths through ONE
    b2a_pipeline_batch call (the library forks them over its internal streams), each checked against the oracle run on the
    same bytes: ranges and trimmed PCM bit-exact, log-mel <= 1e-4; then the same batch replayed as one CUDA graph"""
    from oracle import pydub_silence as ps, whisper_logmel as wl
    clips = [_cfg4_like_clip(T, 40 + i, secs) for i, secs in enumerate((21.0, 9.5, 33.3, 14.2, 6.1))]
    plans = [ops.PipelinePlan(int(c.shape[0]), 48000, 2, c.dtype, n_mels=80, padding=0) for c in clips]
    batch = ops.PipelineBatch(plans)
    kw = dict(min_silence_len=1000, silence_thresh=-40, keep_silence=200, seek_step=1)
    for rep in range(3):                                     # eager, capture, replay
        res = batch.run(clips, graph=True, **kw)
        T.cuda.synchronize()
        for c, r in zip(clips, res):
            full = ops.resample(c, 48000)[0].cpu().numpy()
            assert r.kept == ps.kept_ranges_fast(full, 16000, **kw) and r.nonsilent == ps.detect_nonsilent_fast(full, 16000, 1000, -40, 1)
            trimmed = ps.strip_silence_fast(full, 16000, **kw)
            assert np.array_equal(r.pcm.cpu().numpy(), trimmed) and 0.2 * len(full) < len(trimmed) < 0.9 * len(full)
            ref = wl.log_mel_spectrogram(trimmed.astype(np.float32) / 32768.0, 80).numpy()
            assert np.abs(r.mel.cpu().numpy() - ref).max() <= MEL_TOL


def test_c_abi_from_six_threads(ops, T):
    """the reference calls the front-end from ThreadPoolExecutor workers (audio_processor.py:56): six Python threads, each
    with its own stream and its own clips, drive b2a_resample / b2a_detect_silence / b2a_log_mel / b2a_pipeline concurrently
    (ctypes releases the GIL during the calls); every result is checked against the oracle"""
    import threading
    from audio_processor_b200 import synth
    from audio_processor_b200.service import AudioFrontend
    from oracle import pydub_silence as ps, resample_oracle as ro, whisper_logmel as wl
    fe = AudioFrontend(min_silence_len=700, silence_thresh=-40, keep_silence=150)       # ONE shared instance, as in the reference
    kw = dict(min_silence_len=700, silence_thresh=-40, keep_silence=150, seek_step=1)
    errors = []

    def worker(tid):
        try:
            rate = (44100, 48000, 16000)[tid % 3]
            ch = 2 if rate != 16000 else 1
            stream = T.cuda.Stream()
            with T.cuda.stream(stream):
                for it in range(4):
                    x = synth.synth_clip(100 * tid + it, rate, ch, 6.0 + tid + 0.37 * it, 0.35, device="cuda")
                    y16, _, en = ops.resample(x, rate, want_energy=True)
                    r = ops.detect(y16, 16000, energy=en, **kw)
                    mel = ops.log_mel(y16, 80)
                    pcm_t, mel_t, kept = fe.process_pcm(x, rate)
                    stream.synchronize()
                    full = y16.cpu().numpy()
                    ref16 = ro.convert(x.cpu().numpy(), rate)
                    assert len(full) == len(ref16) and np.abs(full.astype(int) - ref16.astype(int)).max() <= 1
                    want = ps.kept_ranges_fast(full, 16000, **kw)
                    assert r.kept == want and kept == want and fe.last_segments == want
                    trimmed = ps.strip_silence_fast(full, 16000, **kw)
                    assert np.array_equal(pcm_t.cpu().numpy(), trimmed)
                    assert np.abs(mel.cpu().numpy() - wl.log_mel_spectrogram(full.astype(np.float32) / 32768.0, 80).numpy()).max() <= MEL_TOL
                    assert np.abs(mel_t.cpu().numpy() - wl.log_mel_spectrogram(trimmed.astype(np.float32) / 32768.0, 80).numpy()).max() <= MEL_TOL
        except Exception as e:                                # noqa: BLE001
            import traceback
            errors.append((tid, traceback.format_exc()))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors[0][1]
    # results of earlier calls stay valid when the same thread runs the next clip of the same shape
    a = synth.synth_clip(900, 44100, 2, 5.0, 0.3, device="cuda")
    b = synth.synth_clip(901, 44100, 2, 5.0, 0.3, device="cuda")
    ra, rb = fe.process_pcm(a, 44100), fe.process_pcm(b, 44100)
    ra2 = fe.process_pcm(a, 44100)
    assert T.equal(ra[1], ra2[1]) and T.equal(ra[0], ra2[0]) and ra[1].data_ptr() != rb[1].data_ptr() and ra[2] == ra2[2]


def test_remap_times_on_device(ops, T):
    """b2a_remap_times: Whisper-style segment times on the trimmed timeline -> original recording, against oracle/remap.py,
    with the kept table of a real pipeline run (and the derived offsets against the silence detector's own)"""
    from audio_processor_b200 import synth
    from oracle import remap as orm
    x = synth.synth_clip(77, 44100, 2, 40.0, 0.4, device="cuda")
    plan = ops.PipelinePlan(int(x.shape[0]), 44100, 2, x.dtype, n_mels=80, padding=0)
    r = plan.run(x, min_silence_len=1000, silence_thresh=-40, keep_silence=200)
    kept = r.kept
    assert len(kept) >= 3
    rng = np.random.default_rng(5)
    total_s = sum(e - s for s, e in kept) / 1000.0
    times = np.concatenate([rng.uniform(0, total_s, 500), [0.0, total_s, total_s + 3.0],
                            np.cumsum([e - s for s, e in kept]) / 1000.0])            # incl. times exactly on the cuts
    got = ops.remap_times(times, plan.kept, plan.info).cpu().numpy()
    ref = np.array([orm.remap_time(t, kept) for t in times])
    assert np.abs(got - ref).max() <= 1e-9
    y16, _, en = ops.resample(x, 44100, want_energy=True)
    d = ops.detect(y16, 16000, 1000, -40, 200, 1, energy=en)
    got2 = ops.remap_times(times, d.kept_ms, d.info, kept_off=d.kept_off).cpu().numpy()
    assert np.array_equal(got, got2)
    from audio_processor_b200 import service
    segs = [{"start": float(a), "end": float(a) + 0.5, "text": "x"} for a in times[:20]]
    host = service.remap_segments(segs, kept)
    assert all(abs(h["start"] - orm.remap_time(s["start"], kept)) <= 1e-9 for h, s in zip(host, segs))


def test_convert_to_wav_from_compressed_containers(ops, T, tmp_path):
    """f4: convert_to_wav("x.m4a") / ("x.flac") — the call process_audio makes for every non-WAV upload
    (audio_processor.py:1040-1044): host decode (libavcodec) -> b2a_resample on the GPU -> 16 kHz mono s16 WAV, equal to
    what the real libswresample makes of the same decoded PCM (<= 1 LSB), i.e. to the WAV route on the decoded samples"""
    import os
    import shutil
    from audio_processor_b200 import avdecode, wavio
    from audio_processor_b200.service import AudioFrontend
    from oracle import resample_oracle as ro, swr_ref
    if not avdecode.available():
        pytest.skip("bundled FFmpeg libraries not found")
    fe = AudioFrontend()
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name in ("tone_stereo_22k.m4a", "tone_stereo_22k.flac"):
        src = str(tmp_path / name)
        shutil.copy(os.path.join(here, name), src)
        out = fe.convert_to_wav(src)
        assert out == os.path.splitext(src)[0] + ".wav" and os.path.exists(out)
        y, rate = wavio.read_wav(out)
        assert rate == 16000 and y.ndim == 1 and y.dtype == np.int16
        pcm, in_rate = avdecode.decode_audio(src)
        ref = swr_ref.convert(pcm, in_rate) if swr_ref.available() else ro.convert(pcm, in_rate)
        assert len(y) == len(ref) and np.abs(y.astype(int) - ref.astype(int)).max() <= 1
        # the same decoded PCM through the WAV route gives the same file
        wav_in = str(tmp_path / (name + ".pcm.wav"))
        if pcm.dtype == np.int16:
            import wave
            w = wave.open(wav_in, "wb"); w.setnchannels(2); w.setsampwidth(2); w.setframerate(in_rate); w.writeframes(pcm.tobytes()); w.close()
            y2, _ = wavio.read_wav(fe.convert_to_wav(wav_in))
            assert np.array_equal(y2, y)
    import subprocess
    bad = str(tmp_path / "noise.m4a")
    open(bad, "wb").write(b"not audio at all" * 100)
    with pytest.raises(subprocess.CalledProcessError):
        fe.convert_to_wav(bad)


def test_logmel_row_limit_is_reported(ops, T):
    """a clip whose [n_mels][T] block would reach 2^31 values is refused with B2A_EUNSUPPORTED before anything is launched
    (the kernel addresses a clip's block with 32-bit offsets; include/b2a.h)"""
    import ctypes as C
    from audio_processor_b200 import _lib
    lib = _lib.lib()
    x = T.zeros(1024, dtype=T.int16, device="cuda")
    out = T.zeros(1024, dtype=T.float32, device="cuda")
    n = 160 * ((1 << 31) // 128)                      # T * 128 == 2^31
    n0 = ops.launch_count()
    rc = lib.b2a_log_mel(C.c_void_p(x.data_ptr()), 0, 1, n, n, None, 0, 128, 0, C.c_void_p(out.data_ptr()), None,
                         C.c_void_p(out.data_ptr()), 1 << 40, None)
    assert rc == -2 and b"2^31" in lib.b2a_last_error()
    assert ops.launch_count() == n0
    assert lib.b2a_log_mel(C.c_void_p(x.data_ptr()), 0, 1, n - 160, n - 160, None, 0, 128, 0, C.c_void_p(out.data_ptr()), None,
                           C.c_void_p(out.data_ptr()), 16, None) == -3      # one frame fewer passes the limit (and then fails on the workspace size)


def test_process_audio_file_equals_the_two_helpers(ops, T, tmp_path):
    """AudioFrontend.process_audio_file: process_audio's front half (audio_processor.py:1039-1080) as ONE read, one
    b2a_pipeline call and one write — same WAV bytes as convert_to_wav followed by preprocess_audio, same kept table, and
    the log-mel Whisper would compute from that WAV (padding = N_SAMPLES as transcribe passes it)"""
    import shutil
    from audio_processor_b200 import synth, wavio, whisper_audio
    from audio_processor_b200.service import AudioFrontend
    from oracle import whisper_logmel as wl
    x = synth.synth_clip(9, 44100, 2, 11.0, 0.35, device="cuda").cpu().numpy()
    a_dir, b_dir = tmp_path / "a", tmp_path / "b"
    a_dir.mkdir(); b_dir.mkdir()
    src_a, src_b = str(a_dir / "upload.wav"), str(b_dir / "upload.wav")
    wavio.write_wav_s16(src_a, x, 44100)
    shutil.copy(src_a, src_b)
    fe = AudioFrontend()
    two_step, kept2 = fe.preprocess_audio_segments(fe.convert_to_wav(src_a))
    path, mel, kept = fe.process_audio_file(src_b)
    assert kept == kept2 and path.endswith(".trimmed.wav")
    y2, _ = wavio.read_wav(two_step)
    y, sr = wavio.read_wav(path)
    assert sr == 16000 and np.array_equal(y, y2)
    ref = wl.log_mel_spectrogram(y.astype(np.float32) / 32768.0, 80, padding=whisper_audio.N_SAMPLES).numpy()
    assert mel.shape == ref.shape and np.abs(mel.cpu().numpy() - ref).max() <= MEL_TOL
    # a canonical WAV without silence comes back under its own name
    tone = (np.sin(np.arange(16000 * 3) * 0.05) * 8000).astype(np.int16)
    src_c = str(tmp_path / "tone.wav")
    wavio.write_wav_s16(src_c, tone, 16000)
    path_c, mel_c, kept_c = fe.process_audio_file(src_c)
    assert path_c == src_c and kept_c == [[0, 3000]] and mel_c.shape == (80, (len(tone) + whisper_audio.N_SAMPLES) // 160)
    import subprocess
    bad = tmp_path / "x.m4a"; bad.write_bytes(b"not audio")
    with pytest.raises(subprocess.CalledProcessError):
        fe.process_audio_file(str(bad))
