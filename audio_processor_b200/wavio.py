"""RIFF/WAVE reader + writer (host side of convert_to_wav / preprocess_audio).

The reference shells out to ffmpeg for every container (app/services/audio_processor.py:912-923) and
whisper re-reads the 16-bit WAV it wrote.  In this tier only PCM WAV is decoded on the host (16-bit
and 24-bit int, 32-bit float, WAVE_FORMAT_EXTENSIBLE wrappers of those); compressed inputs are out of scope
 and raise.
"""
from __future__ import annotations

import struct
from typing import Tuple

import numpy as np

WAVE_FORMAT_PCM = 1
WAVE_FORMAT_IEEE_FLOAT = 3
WAVE_FORMAT_EXTENSIBLE = 0xFFFE


class UnsupportedAudio(RuntimeError):
    pass


def read_wav(path: str) -> Tuple[np.ndarray, int]:
    """-> (samples [n] or [n, C] as int16 or float32, sample_rate)"""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 12 or data[0:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise UnsupportedAudio(f"{path}: not a RIFF/WAVE file")
    pos = 12
    fmt = None
    payload = None
    while pos + 8 <= len(data):
        cid = data[pos:pos + 4]
        size = struct.unpack_from("<I", data, pos + 4)[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, rate, _brate, _align, bits = struct.unpack_from("<HHIIHH", body, 0)
            if tag == WAVE_FORMAT_EXTENSIBLE and len(body) >= 26:
                tag = struct.unpack_from("<H", body, 24)[0]
            fmt = (tag, ch, rate, bits)
        elif cid == b"data":
            payload = body
        pos += 8 + size + (size & 1)
    if fmt is None or payload is None:
        raise UnsupportedAudio(f"{path}: missing fmt/data chunk")
    tag, ch, rate, bits = fmt
    if tag == WAVE_FORMAT_PCM and bits == 16:
        a = np.frombuffer(payload, dtype="<i2", count=len(payload) // 2)
    elif tag == WAVE_FORMAT_IEEE_FLOAT and bits == 32:
        a = np.frombuffer(payload, dtype="<f4", count=len(payload) // 4)
    elif tag == WAVE_FORMAT_PCM and bits == 24:
        # packed little-endian s24 -> float32 in [-1, 1) (exact: 24 bits fit the f32 significand); goes down the float path
        raw = np.frombuffer(payload, dtype=np.uint8, count=(len(payload) // 3) * 3).reshape(-1, 3).astype(np.int32)
        v = raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)
        v = np.where(v >= (1 << 23), v - (1 << 24), v)
        a = (v.astype(np.float32) / np.float32(1 << 23)).astype(np.float32)
    else:
        raise UnsupportedAudio(f"{path}: WAV format tag {tag} / {bits} bits is not supported (s16, s24 and f32 only)")
    if ch > 1:
        a = a[: (len(a) // ch) * ch].reshape(-1, ch)
    return np.ascontiguousarray(a), int(rate)


def write_wav_s16(path: str, samples: np.ndarray, sample_rate: int = 16000) -> None:
    """mono/stereo int16 -> canonical 44-byte-header PCM WAV (what `-c:a pcm_s16le` produces, minus LIST tags)."""
    a = np.ascontiguousarray(samples, dtype="<i2")
    ch = 1 if a.ndim == 1 else a.shape[1]
    payload = a.tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(payload)) + b"WAVE"
    hdr += b"fmt " + struct.pack("<IHHIIHH", 16, WAVE_FORMAT_PCM, ch, sample_rate, sample_rate * ch * 2, ch * 2, 16)
    hdr += b"data" + struct.pack("<I", len(payload))
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(payload)


def read_audio(path: str) -> Tuple[np.ndarray, int]:
    """Any input the service accepts -> (PCM [n] or [n, C] as int16 or float32, sample_rate).

    WAV (s16 / s24 / f32) is parsed here; everything else — the m4a / mp3 downloads ``process_audio`` hands to
    ``convert_to_wav`` (/root/reference/app/services/audio_processor.py:1040-1044), flac, ogg, WAV flavours this parser does
    not know — is demuxed and decoded on the host by the bundled libavformat / libavcodec (avdecode.py), as the reference's
    ``ffmpeg -i`` does (:912-920).  Resampling, downmix and quantisation stay on the GPU."""
    try:
        return read_wav(path)
    except UnsupportedAudio as wav_err:
        from . import avdecode
        try:
            return avdecode.decode_audio(path)
        except avdecode.DecodeError as e:
            raise UnsupportedAudio(f"{wav_err}; decoder: {e}") from e
