"""float64 restatement of openai-whisper's audio front-end.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Reached from the reference at /root/reference/app/services/audio_processor.py:1076-1080
(``model_to_use.transcribe(audio_path, ...)``): ``whisper.transcribe`` calls
``whisper.audio.log_mel_spectrogram(audio, model.dims.n_mels, padding=N_SAMPLES)`` on the CPU.
openai-whisper is not vendored under /root/reference (requirements.txt:25, unpinned; n_mels=128
needs >= v20231106, pinned here as **openai-whisper v20231117**) and is absent from this image.
Published algorithm restated (whisper/audio.py):

    SAMPLE_RATE=16000  N_FFT=400  HOP_LENGTH=160  CHUNK_LENGTH=30  N_SAMPLES=480000  N_FRAMES=3000
    load_audio:  int16 PCM -> float32 / 32768.0
    log_mel_spectrogram(audio, n_mels=80, padding=0):
        audio  = F.pad(audio, (0, padding))
        window = torch.hann_window(N_FFT)                       # periodic
        stft   = torch.stft(audio, N_FFT, HOP_LENGTH, window=window, return_complex=True)
        magnitudes = stft[..., :-1].abs() ** 2
        mel_spec = mel_filters(n_mels) @ magnitudes              # librosa slaney mel, f32
        log_spec = torch.clamp(mel_spec, min=1e-10).log10()
        log_spec = torch.maximum(log_spec, log_spec.max() - 8.0) # max over the whole call
        log_spec = (log_spec + 4.0) / 4.0
    pad_or_trim(array, length=N_SAMPLES, axis=-1)

``mel_filters`` is whisper/assets/mel_filters.npz = librosa.filters.mel(sr=16000, n_fft=400,
n_mels) (slaney scale, slaney norm) stored as float32; regenerated here in float64 and rounded
to float32.  Pinning: tests/test_oracle_logmel.py cross-checks this file against
transformers' WhisperFeatureExtractor (an independent restatement) and
transformers.audio_utils.mel_filter_bank, against torch.stft float32, and against the
SURVEY A.5 known answers M80/M128.
"""
from __future__ import annotations

import numpy as np
import torch

SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE
N_FRAMES = N_SAMPLES // HOP_LENGTH


def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filters_f64(n_mels: int, sr: int = SAMPLE_RATE, n_fft: int = N_FFT) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, htk=False, norm='slaney') in float64: [n_mels, 1+n_fft//2]"""
    n_freqs = 1 + n_fft // 2
    fftfreqs = np.linspace(0.0, sr / 2.0, n_freqs)
    mel_pts = np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2)
    mel_f = _mel_to_hz(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, n_freqs), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return w * enorm[:, None]


def mel_filters(n_mels: int) -> np.ndarray:
    """float32 filterbank as whisper ships it."""
    assert n_mels in (80, 128), f"Unsupported n_mels: {n_mels}"
    return mel_filters_f64(n_mels).astype(np.float32)


def load_audio_from_s16(pcm: np.ndarray) -> np.ndarray:
    return np.asarray(pcm, dtype=np.int16).astype(np.float32) / 32768.0


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    if torch.is_tensor(array):
        if array.shape[axis] > length:
            array = array.index_select(dim=axis, index=torch.arange(length, device=array.device))
        if array.shape[axis] < length:
            pad_widths = [(0, 0)] * array.ndim
            pad_widths[axis] = (0, length - array.shape[axis])
            array = torch.nn.functional.pad(array, [p for sizes in pad_widths[::-1] for p in sizes])
    else:
        if array.shape[axis] > length:
            array = array.take(indices=range(length), axis=axis)
        if array.shape[axis] < length:
            pad_widths = [(0, 0)] * array.ndim
            pad_widths[axis] = (0, length - array.shape[axis])
            array = np.pad(array, pad_widths)
    return array


def log_mel_spectrogram(audio, n_mels: int = 80, padding: int = 0, dtype=torch.float64,
                        per_clip_max: bool = False) -> torch.Tensor:
    """Whisper's function restated; ``dtype`` float64 = oracle of record, float32 = what the
    reference actually executes on the CPU.  ``per_clip_max`` reproduces HF's per-row max
    (feature_extraction_whisper.py) for batched input instead of Whisper's whole-call max."""
    if not torch.is_tensor(audio):
        audio = torch.from_numpy(np.asarray(audio))
    audio = audio.to(dtype)
    if padding > 0:
        audio = torch.nn.functional.pad(audio, (0, padding))
    window = torch.hann_window(N_FFT, dtype=torch.float64).to(dtype)
    stft = torch.stft(audio, N_FFT, HOP_LENGTH, window=window, return_complex=True)
    magnitudes = stft[..., :-1].abs() ** 2
    filters = torch.from_numpy(mel_filters(n_mels)).to(dtype)
    mel_spec = filters @ magnitudes
    log_spec = torch.clamp(mel_spec, min=1e-10).log10()
    if per_clip_max and log_spec.dim() == 3:
        mx = log_spec.amax(dim=(1, 2), keepdim=True)
        log_spec = torch.maximum(log_spec, mx - 8.0)
    else:
        log_spec = torch.maximum(log_spec, log_spec.max() - 8.0)
    log_spec = (log_spec + 4.0) / 4.0
    return log_spec


def encoder_windows(mel: np.ndarray, n_frames: int = N_FRAMES, *, content_frames=None, seek0: int = 0, stride=None,
                    n_windows=None, dtype=np.float32) -> np.ndarray:
    """The windows ``whisper.transcribe`` feeds the encoder, restated in numpy (whisper/transcribe.py main loop):

        content_frames = mel.shape[-1] - N_FRAMES            # the mel was computed with padding = N_SAMPLES
        while seek < content_frames:
            segment_size = min(N_FRAMES, content_frames - seek)
            mel_segment = mel[:, seek : seek + segment_size]
            mel_segment = pad_or_trim(mel_segment, N_FRAMES).to(device).to(dtype)

    for every window of a uniform grid at once (window w starts at ``seek0 + w * stride``; transcribe without timestamp
    seeking advances by ``n_frames``).  Frames at or beyond ``content_frames`` read as zero; the cast is numpy's
    round-to-nearest-even (``astype(np.float16)`` = torch ``.half()``).  Returns [n_windows, n_mels, n_frames]."""
    mel = np.asarray(mel, dtype=np.float32)
    T = mel.shape[-1]
    content = max(T - N_FRAMES, 0) if content_frames is None else int(content_frames)
    stride = n_frames if stride is None else int(stride)
    if n_windows is None:
        n_windows = max((content - seek0 + stride - 1) // stride, 0)
    out = np.zeros((n_windows, mel.shape[0], n_frames), dtype=dtype)
    for w in range(n_windows):
        s = seek0 + w * stride
        e = min(s + n_frames, content)
        if e > s:
            out[w, :, : e - s] = mel[:, s:e].astype(dtype)
    return out
