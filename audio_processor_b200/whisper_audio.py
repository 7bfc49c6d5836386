"""Drop-in for the ``whisper.audio`` names the reference reaches through ``model.transcribe``
(/root/reference/app/services/audio_processor.py:1076-1080): same names, signatures and constants as
openai-whisper ``whisper/audio.py``; the arithmetic runs in libb2a's sm_100a kernels.

    from audio_processor_b200 import whisper_audio
    mel = whisper_audio.log_mel_spectrogram("meeting.wav", n_mels=80, padding=whisper_audio.N_SAMPLES)
    whisper_audio.patch_whisper()      # optional: rebind whisper.audio / whisper.transcribe to these
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np

from . import ops, wavio

# whisper/audio.py hard-coded audio hyperparameters
SAMPLE_RATE = 16000
N_FFT = 400
HOP_LENGTH = 160
CHUNK_LENGTH = 30
N_SAMPLES = CHUNK_LENGTH * SAMPLE_RATE  # 480000 samples in a 30-second chunk
N_FRAMES = N_SAMPLES // HOP_LENGTH  # 3000 frames in a mel spectrogram input
N_SAMPLES_PER_TOKEN = HOP_LENGTH * 2
FRAMES_PER_SECOND = SAMPLE_RATE // HOP_LENGTH
TOKENS_PER_SECOND = SAMPLE_RATE // N_SAMPLES_PER_TOKEN


def load_audio(file: str, sr: int = SAMPLE_RATE) -> np.ndarray:
    """whisper.audio.load_audio: file -> mono float32 waveform at ``sr`` in [-1, 1].

    whisper shells out to ``ffmpeg -i file -f s16le -ac 1 -acodec pcm_s16le -ar sr -``; here the file is parsed (WAV) or
    decoded (m4a, mp3, flac, ...: bundled libavcodec, avdecode.py) on the host and downmix + resampling + s16 quantisation
    run on the GPU, then ``/ 32768.0`` as whisper does."""
    torch = ops.require_cuda()
    pcm, rate = wavio.read_audio(file)
    s16, _, _ = ops.resample(pcm, rate, sr)
    return (s16.to(torch.float32) / 32768.0).cpu().numpy()


def pad_or_trim(array, length: int = N_SAMPLES, *, axis: int = -1):
    """whisper.audio.pad_or_trim (numpy arrays and torch tensors)."""
    import torch
    if torch.is_tensor(array):
        if array.shape[axis] > length:
            array = array.index_select(dim=axis, index=torch.arange(length, device=array.device))
        if array.shape[axis] < length:
            pad_widths = [(0, 0)] * array.ndim
            pad_widths[axis] = (0, length - array.shape[axis])
            array = torch.nn.functional.pad(array, [pad for sizes in pad_widths[::-1] for pad in sizes])
    else:
        if array.shape[axis] > length:
            array = array.take(indices=range(length), axis=axis)
        if array.shape[axis] < length:
            pad_widths = [(0, 0)] * array.ndim
            pad_widths[axis] = (0, length - array.shape[axis])
            array = np.pad(array, pad_widths)
    return array


def mel_filters(device=None, n_mels: int = 80):
    """The float32 [n_mels, 201] slaney filterbank (whisper loads it from assets/mel_filters.npz)."""
    import ctypes as C
    import torch
    from ._lib import check, lib
    assert n_mels in {80, 128}, f"Unsupported n_mels: {n_mels}"
    buf = np.empty((n_mels, 201), dtype=np.float32)
    check(lib().b2a_mel_filters(n_mels, buf.ctypes.data_as(C.c_void_p), buf.size))
    t = torch.from_numpy(buf)
    return t.to(device) if device is not None else t


def log_mel_spectrogram(audio: Union[str, np.ndarray, "object"], n_mels: int = 80, padding: int = 0,
                        device: Optional[Union[str, "object"]] = None):
    """whisper.audio.log_mel_spectrogram: -> torch.Tensor [n_mels, T] (or [B, n_mels, T] for 2-D input).

    Like whisper, the floor ``max - 8`` is taken over everything this call returns.  The result lives on the
    CUDA device that did the work (``device`` selects it; whisper's default of "same device as the input"
    becomes "current CUDA device" for host inputs, since there is no CPU path)."""
    torch = ops.require_cuda()
    if isinstance(audio, str):
        audio = load_audio(audio)
    if isinstance(audio, np.ndarray):
        audio = torch.from_numpy(np.ascontiguousarray(audio))
    if not torch.is_tensor(audio):
        raise TypeError("audio must be a path, numpy array or torch tensor")
    if audio.dtype not in (torch.float32, torch.int16):
        audio = audio.to(torch.float32)
    if device is not None:
        audio = audio.to(device)
    elif not audio.is_cuda:
        audio = audio.cuda()
    return ops.log_mel(audio, n_mels=n_mels, padding=padding)


def mel_segments(mel, n_frames: int = N_FRAMES, dtype=None):
    """The 30-second windows ``whisper.transcribe`` feeds the encoder: ``mel[:, seek:seek + n_frames]`` for
    seek = 0, n_frames, ... over the content frames, each zero-padded to ``n_frames`` by pad_or_trim (transcribe computes
    the mel with ``padding=N_SAMPLES`` and treats the last N_FRAMES as padding; pass such a mel here).  Yields
    (seek, segment) with segment [n_mels, n_frames] on the mel's device (optionally cast to ``dtype``), no host round trip."""
    content_frames = int(mel.shape[-1]) - N_FRAMES
    seek = 0
    while seek < content_frames:
        size = min(n_frames, content_frames - seek)
        seg = pad_or_trim(mel[:, seek:seek + size], n_frames)
        yield seek, (seg.to(dtype) if dtype is not None else seg)      # e.g. torch.float16 for an fp16 encoder
        seek += size


def mel_windows(mel, n_frames: int = N_FRAMES, dtype=None):
    """Every window ``mel_segments`` would yield, as one [n_windows, n_mels, n_frames] tensor produced by a single kernel
    (slice + zero pad + cast fused; ``dtype`` torch.float16 for an fp16 encoder) — the batched form of transcribe's loop
    for callers that run the encoder over all 30-second windows at once."""
    return ops.mel_windows(mel, n_frames, dtype=dtype)


def patch_whisper() -> None:
    """Rebind openai-whisper's front-end to this module.

    ``whisper/transcribe.py`` does ``from .audio import log_mel_spectrogram, pad_or_trim, ...`` and ``whisper/__init__.py``
    does ``from .transcribe import transcribe``: the names live in three module namespaces, and ``whisper.transcribe`` as an
    attribute of the package is the FUNCTION, not the submodule — so the submodules are fetched from ``sys.modules`` by
    their dotted names (importlib), and every one that binds a name gets it replaced."""
    import importlib
    import sys
    importlib.import_module("whisper")
    for modname in ("whisper.audio", "whisper.transcribe", "whisper"):
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            continue
        mod = sys.modules.get(modname, mod)
        for name in ("log_mel_spectrogram", "load_audio", "pad_or_trim"):
            if hasattr(mod, name):
                setattr(mod, name, globals()[name])
