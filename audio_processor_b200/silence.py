"""pydub-style silence API on the GPU.

Mirrors ``pydub.silence`` (pydub 0.25.1): same function names, argument order, defaults and return
shapes, so the silence strip the reference intends in ``preprocess_audio``
(/root/reference/app/services/audio_processor.py:305-314, call site :1046-1047) can be written exactly as
one would with pydub:

    seg = AudioSegment.from_wav(path)                                     # ours: silence.AudioSegment
    chunks = split_on_silence(seg, min_silence_len=1000, silence_thresh=-40, keep_silence=200)

Ranges are bit-identical to pydub's (tests compare against a literal restatement on ``audioop``).
Only mono 16-bit segments at a whole number of samples per millisecond are supported on the GPU path;
anything else raises (there is no CPU fallback).
"""
from __future__ import annotations

from typing import List, Union

import numpy as np

from . import ops


class AudioSegment:
    """Minimal stand-in for pydub.AudioSegment: raw mono int16 PCM + frame_rate.

    ``data`` may live on the host (numpy / bytes) or on the device (torch CUDA tensor); it is uploaded
    once and cached."""

    sample_width = 2
    channels = 1

    def __init__(self, data, frame_rate: int = 16000):
        import torch
        if isinstance(data, (bytes, bytearray, memoryview)):
            data = np.frombuffer(data, dtype=np.int16)
        if isinstance(data, np.ndarray):
            if data.dtype != np.int16 or data.ndim != 1:
                raise TypeError("AudioSegment expects mono int16 samples")
            self._host = np.ascontiguousarray(data)
            self._dev = None
        elif torch.is_tensor(data):
            if data.dtype != torch.int16 or data.dim() != 1:
                raise TypeError("AudioSegment expects mono int16 samples")
            self._host = None if data.is_cuda else data.numpy()
            self._dev = data if data.is_cuda else None
        else:
            raise TypeError("unsupported data type")
        self.frame_rate = int(frame_rate)
        self.frame_width = 2

    # -- construction helpers -------------------------------------------------------------
    @classmethod
    def from_wav(cls, path: str) -> "AudioSegment":
        from . import wavio
        pcm, rate = wavio.read_wav(path)
        if pcm.ndim != 1 or pcm.dtype != np.int16:
            raise ValueError("from_wav: GPU silence path needs a mono 16-bit WAV (convert_to_wav output)")
        return cls(pcm, rate)

    # -- pydub surface ----------------------------------------------------------------------
    def frame_count(self, ms=None):
        if ms is not None:
            return ms * (self.frame_rate / 1000.0)
        return float(self.n_frames)

    @property
    def n_frames(self) -> int:
        return int(self._dev.shape[0] if self._dev is not None else self._host.shape[0])

    def __len__(self) -> int:
        return round(1000 * (self.frame_count() / self.frame_rate))

    @property
    def max_possible_amplitude(self):
        return 32768.0

    def device_samples(self):
        if self._dev is None:
            torch = ops.require_cuda()
            self._dev = torch.from_numpy(self._host).cuda()
        return self._dev

    def get_array_of_samples(self) -> np.ndarray:
        if self._host is None:
            self._host = self._dev.cpu().numpy()
        return self._host

    @property
    def raw_data(self) -> bytes:
        return self.get_array_of_samples().tobytes()

    def __getitem__(self, ms: slice) -> "AudioSegment":
        """seg[start_ms:end_ms] with pydub's clamping and its zero-fill of a rounded-up last millisecond."""
        if not isinstance(ms, slice):
            raise TypeError("only millisecond slices are supported")
        L = len(self)
        start = 0 if ms.start is None else ms.start
        end = L if ms.stop is None else ms.stop
        start, end = min(start, L), min(end, L)
        if start < 0:
            start = L - abs(start)
        if end < 0:
            end = L - abs(end)
        a, b = int(self.frame_count(ms=start)), int(self.frame_count(ms=end))
        x = self.get_array_of_samples()
        piece = x[a:min(b, len(x))]
        if len(piece) < b - a:
            piece = np.concatenate([piece, np.zeros(b - a - len(piece), dtype=np.int16)])
        return AudioSegment(piece, self.frame_rate)

    def __add__(self, other: "AudioSegment") -> "AudioSegment":
        return AudioSegment(np.concatenate([self.get_array_of_samples(), other.get_array_of_samples()]), self.frame_rate)


def _detect(audio_segment: AudioSegment, min_silence_len, silence_thresh, keep_silence, seek_step) -> ops.SilenceResult:
    return ops.detect(audio_segment.device_samples(), audio_segment.frame_rate, min_silence_len, silence_thresh,
                      keep_silence, seek_step)


def detect_silence(audio_segment: AudioSegment, min_silence_len: int = 1000, silence_thresh: float = -16,
                   seek_step: int = 1) -> List[List[int]]:
    return _detect(audio_segment, min_silence_len, silence_thresh, 0, seek_step).silent


def detect_nonsilent(audio_segment: AudioSegment, min_silence_len: int = 1000, silence_thresh: float = -16,
                     seek_step: int = 1) -> List[List[int]]:
    return _detect(audio_segment, min_silence_len, silence_thresh, 0, seek_step).nonsilent


def split_on_silence(audio_segment: AudioSegment, min_silence_len: int = 1000, silence_thresh: float = -16,
                     keep_silence: Union[int, bool] = 100, seek_step: int = 1) -> List[AudioSegment]:
    res = _detect(audio_segment, min_silence_len, silence_thresh, keep_silence, seek_step)
    return [audio_segment[s:e] for s, e in res.kept]


def strip_silence(audio_segment: AudioSegment, min_silence_len: int = 1000, silence_thresh: float = -16,
                  keep_silence: Union[int, bool] = 100, seek_step: int = 1) -> AudioSegment:
    """sum(split_on_silence(...)) without leaving the device: ranges + stream compaction kernels."""
    res = _detect(audio_segment, min_silence_len, silence_thresh, keep_silence, seek_step)
    buf = ops.compact(audio_segment.device_samples(), res, audio_segment.frame_rate)
    return AudioSegment(buf[: res.n_keep], audio_segment.frame_rate)
