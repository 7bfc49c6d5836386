"""ctypes driver for the real FFmpeg libswresample that ships in this image.

TEST INFRASTRUCTURE (see oracle/__init__.py).  This is the *engine* behind the
reference's ``ffmpeg -ar 16000 -ac 1 -c:a pcm_s16le`` call
(/root/reference/app/services/audio_processor.py:912-923): the ffmpeg CLI passes
no resampler options, so libswresample's defaults apply.  The library found
here is FFmpeg 8.0.1 / libswresample 6.1.100 (bundled with
opencv_python_headless); the same image runs on the GPU box.

Only plain C entry points are used: swr_alloc_set_opts2 / swr_init /
swr_convert / swr_free and av_channel_layout_default.
"""
from __future__ import annotations

import ctypes
import glob
import os
from typing import Optional

import numpy as np

AV_SAMPLE_FMT_S16 = 1
AV_SAMPLE_FMT_FLT = 3

_LIBS = None


def _libdir() -> Optional[str]:
    try:
        import cv2  # noqa: F401  (only to locate site-packages reliably)
        base = os.path.dirname(os.path.dirname(cv2.__file__))
    except Exception:
        import site
        cands = site.getsitepackages()
        base = cands[0] if cands else ""
    d = os.path.join(base, "opencv_python_headless.libs")
    return d if os.path.isdir(d) else None


def _first(d: str, pat: str) -> str:
    m = sorted(glob.glob(os.path.join(d, pat)))
    if not m:
        raise OSError(f"{pat} not found in {d}")
    return m[0]


def available() -> bool:
    try:
        _load()
        return True
    except OSError:
        return False


def _load():
    global _LIBS
    if _LIBS is not None:
        return _LIBS
    d = _libdir()
    if d is None:
        raise OSError("opencv_python_headless.libs not found (no bundled FFmpeg)")
    mode = ctypes.RTLD_GLOBAL
    # dependency order matters: the bundled .so files carry mangled sonames.
    # libavutil needs libdrm; libswresample needs libavutil.
    for pat in ("libdrm-*.so*",):
        try:
            ctypes.CDLL(_first(d, pat), mode=mode)
        except OSError:
            pass
    avutil = ctypes.CDLL(_first(d, "libavutil-*.so*"), mode=mode)
    swr = ctypes.CDLL(_first(d, "libswresample-*.so*"), mode=mode)

    avutil.av_version_info.restype = ctypes.c_char_p
    avutil.av_channel_layout_default.argtypes = [ctypes.c_void_p, ctypes.c_int]
    avutil.av_channel_layout_default.restype = None
    avutil.av_opt_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    avutil.av_opt_set.restype = ctypes.c_int
    swr.swresample_version.restype = ctypes.c_uint
    swr.swr_alloc_set_opts2.argtypes = [
        ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    swr.swr_alloc_set_opts2.restype = ctypes.c_int
    swr.swr_init.argtypes = [ctypes.c_void_p]
    swr.swr_init.restype = ctypes.c_int
    swr.swr_convert.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int,
                                ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]
    swr.swr_convert.restype = ctypes.c_int
    swr.swr_free.argtypes = [ctypes.POINTER(ctypes.c_void_p)]
    swr.swr_free.restype = None
    _LIBS = (avutil, swr)
    return _LIBS


class _AVChannelLayout(ctypes.Structure):
    _fields_ = [("order", ctypes.c_int), ("nb_channels", ctypes.c_int),
                ("mask", ctypes.c_uint64), ("opaque", ctypes.c_void_p)]


def versions() -> tuple[str, int]:
    avutil, swr = _load()
    return avutil.av_version_info().decode(), int(swr.swresample_version())


def convert(pcm: np.ndarray, in_rate: int, out_rate: int = 16000, out_fmt: str = "s16",
            internal_fmt: Optional[str] = None) -> np.ndarray:
    """Run ``pcm`` ([n] or [n, C]; int16 or float32, interleaved) through
    libswresample to ``out_rate`` mono, like ``ffmpeg -ar out_rate -ac 1``.

    out_fmt "s16" → int16 (what ``-c:a pcm_s16le`` stores), "flt" → float32
    (pre-quantisation, for the ≤1e-5 normalised-PCM gate).
    """
    avutil, swr = _load()
    a = np.ascontiguousarray(pcm)
    if a.ndim == 1:
        a = a[:, None]
    n_in, ch = a.shape
    if a.dtype == np.int16:
        in_fmt = AV_SAMPLE_FMT_S16
    elif a.dtype == np.float32:
        in_fmt = AV_SAMPLE_FMT_FLT
    else:
        raise TypeError("pcm must be int16 or float32")
    ofmt = AV_SAMPLE_FMT_S16 if out_fmt == "s16" else AV_SAMPLE_FMT_FLT
    odt = np.int16 if out_fmt == "s16" else np.float32

    lin, lout = _AVChannelLayout(), _AVChannelLayout()
    avutil.av_channel_layout_default(ctypes.byref(lin), ch)
    avutil.av_channel_layout_default(ctypes.byref(lout), 1)
    ctx = ctypes.c_void_p(None)
    rc = swr.swr_alloc_set_opts2(ctypes.byref(ctx), ctypes.byref(lout), ofmt, out_rate,
                                 ctypes.byref(lin), in_fmt, in_rate, 0, None)
    if rc < 0 or not ctx:
        raise RuntimeError(f"swr_alloc_set_opts2 failed ({rc})")
    try:
        if internal_fmt is not None:
            avutil.av_opt_set(ctx, b"internal_sample_fmt", internal_fmt.encode(), 0)
        rc = swr.swr_init(ctx)
        if rc < 0:
            raise RuntimeError(f"swr_init failed ({rc})")
        cap = int(n_in * out_rate // in_rate) + 4096
        out = np.empty(cap, dtype=odt)
        got = 0
        inp = (ctypes.c_void_p * 1)(a.ctypes.data)
        outp = (ctypes.c_void_p * 1)(out.ctypes.data)
        r = swr.swr_convert(ctx, outp, cap, inp, n_in)
        if r < 0:
            raise RuntimeError(f"swr_convert failed ({r})")
        got += r
        while True:  # flush
            outp = (ctypes.c_void_p * 1)(out.ctypes.data + got * out.itemsize)
            r = swr.swr_convert(ctx, outp, cap - got, None, 0)
            if r < 0:
                raise RuntimeError(f"swr_convert(flush) failed ({r})")
            if r == 0:
                break
            got += r
        return out[:got].copy()
    finally:
        swr.swr_free(ctypes.byref(ctx))
