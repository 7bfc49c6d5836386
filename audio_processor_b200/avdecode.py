"""Host-side decode of compressed containers (m4a / AAC, mp3, flac, ogg, ...) to PCM, for the path-level shims.

The service's real inputs are Drive downloads in m4a / mp3 (/root/reference/README.md:22); ``process_audio`` calls
``convert_to_wav`` exactly when the file is NOT a ``.wav`` (/root/reference/app/services/audio_processor.py:1040-1044) and
the reference lets the ``ffmpeg`` CLI demux + decode there (:912-920).  Codec work is host work:
this module drives the FFmpeg 8 ``libavformat`` / ``libavcodec`` shared objects that ship in this image (next to the
libswresample the oracle uses; there is no ``ffmpeg`` binary) through ctypes and hands the decoder's own PCM — float32 for
AAC / MP3, int16 for 16-bit PCM / FLAC — to ``b2a_resample`` on the GPU, which does what libswresample does behind the CLI
(downmix, polyphase resample, s16 quantisation).  Nothing is resampled or requantised on the host.

Only plain C entry points are used.  A handful of leading struct fields are read directly (their offsets have been stable
since FFmpeg 5 and are sanity-checked at run time); everything else goes through AVOptions."""
from __future__ import annotations

import ctypes as C
import glob
import os
from typing import Optional, Tuple

import numpy as np

AVMEDIA_TYPE_AUDIO = 1
AVERROR_EOF = -541478725          # FFERRTAG('E','O','F',' ')
AVERROR_EAGAIN = -11
# enum AVSampleFormat
_FMT = {0: ("u1", False), 1: ("<i2", False), 2: ("<i4", False), 3: ("<f4", False), 4: ("<f8", False),
        5: ("u1", True), 6: ("<i2", True), 7: ("<i4", True), 8: ("<f4", True), 9: ("<f8", True)}

# leading fields read directly (FFmpeg 5 .. 8, 64-bit)
_OFF_FMTCTX_NB_STREAMS = 44       # AVFormatContext: av_class, iformat, oformat, priv_data, pb (5 pointers), ctx_flags, nb_streams
_OFF_FMTCTX_STREAMS = 48          # AVStream **streams
_OFF_STREAM_INDEX = 8             # AVStream: av_class, index, id, codecpar
_OFF_STREAM_CODECPAR = 16
_OFF_PKT_STREAM_INDEX = 36        # AVPacket: buf, pts, dts, data, size, stream_index
_OFF_FRAME_EXTENDED_DATA = 96     # AVFrame: data[8], linesize[8], extended_data, width, height, nb_samples, format
_OFF_FRAME_NB_SAMPLES = 112
_OFF_FRAME_FORMAT = 116


class DecodeError(RuntimeError):
    pass


class _AVChannelLayout(C.Structure):
    _fields_ = [("order", C.c_int), ("nb_channels", C.c_int), ("mask", C.c_uint64), ("opaque", C.c_void_p)]


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


def _load():
    global _LIBS
    if _LIBS is not None:
        return _LIBS
    d = _libdir()
    if d is None:
        raise DecodeError("no FFmpeg libraries found (opencv_python_headless.libs): compressed input cannot be decoded here")
    mode = C.RTLD_GLOBAL
    loaded = {}
    # dependency order: the bundled .so files carry mangled sonames
    for pat in ("libdrm-*", "libcrypto-*", "libssl-*", "libvpx-*", "libaom-*", "libavutil-*", "libswresample-*", "libswscale-*",
                "libavcodec-*", "libavformat-*"):
        for f in sorted(glob.glob(os.path.join(d, pat + ".so*"))):
            try:
                loaded[pat] = C.CDLL(f, mode=mode)
            except OSError:
                pass
    try:
        avu, avc, avf = loaded["libavutil-*"], loaded["libavcodec-*"], loaded["libavformat-*"]
    except KeyError as e:
        raise DecodeError(f"bundled FFmpeg is incomplete: {e}") from e
    P, PP = C.c_void_p, C.POINTER(C.c_void_p)
    avf.avformat_open_input.argtypes = [PP, C.c_char_p, P, P]; avf.avformat_open_input.restype = C.c_int
    avf.avformat_find_stream_info.argtypes = [P, P]; avf.avformat_find_stream_info.restype = C.c_int
    avf.av_find_best_stream.argtypes = [P, C.c_int, C.c_int, C.c_int, PP, C.c_int]; avf.av_find_best_stream.restype = C.c_int
    avf.av_read_frame.argtypes = [P, P]; avf.av_read_frame.restype = C.c_int
    avf.avformat_close_input.argtypes = [PP]; avf.avformat_close_input.restype = None
    avc.avcodec_alloc_context3.argtypes = [P]; avc.avcodec_alloc_context3.restype = P
    avc.avcodec_parameters_to_context.argtypes = [P, P]; avc.avcodec_parameters_to_context.restype = C.c_int
    avc.avcodec_open2.argtypes = [P, P, P]; avc.avcodec_open2.restype = C.c_int
    avc.avcodec_send_packet.argtypes = [P, P]; avc.avcodec_send_packet.restype = C.c_int
    avc.avcodec_receive_frame.argtypes = [P, P]; avc.avcodec_receive_frame.restype = C.c_int
    avc.avcodec_free_context.argtypes = [PP]; avc.avcodec_free_context.restype = None
    avc.av_packet_alloc.argtypes = []; avc.av_packet_alloc.restype = P
    avc.av_packet_unref.argtypes = [P]; avc.av_packet_unref.restype = None
    avc.av_packet_free.argtypes = [PP]; avc.av_packet_free.restype = None
    avu.av_frame_alloc.argtypes = []; avu.av_frame_alloc.restype = P
    avu.av_frame_unref.argtypes = [P]; avu.av_frame_unref.restype = None
    avu.av_frame_free.argtypes = [PP]; avu.av_frame_free.restype = None
    avu.av_opt_get_int.argtypes = [P, C.c_char_p, C.c_int, C.POINTER(C.c_int64)]; avu.av_opt_get_int.restype = C.c_int
    avu.av_opt_get_chlayout.argtypes = [P, C.c_char_p, C.c_int, C.POINTER(_AVChannelLayout)]; avu.av_opt_get_chlayout.restype = C.c_int
    avu.av_strerror.argtypes = [C.c_int, C.c_char_p, C.c_size_t]; avu.av_strerror.restype = C.c_int
    avu.av_log_set_level.argtypes = [C.c_int]; avu.av_log_set_level.restype = None
    avu.av_log_set_level(16)            # AV_LOG_ERROR: no chatter on stderr from probing
    _LIBS = (avu, avc, avf)
    return _LIBS


def available() -> bool:
    try:
        _load()
        return True
    except DecodeError:
        return False


def _err(avu, rc: int) -> str:
    buf = C.create_string_buffer(256)
    avu.av_strerror(rc, buf, 256)
    return f"{buf.value.decode(errors='replace')} ({rc})"


def _i32(addr: int) -> int:
    return C.c_int32.from_address(addr).value


def _frame_to_array(frame: int, channels: int) -> np.ndarray:
    """one decoded AVFrame -> [nb_samples, channels] in the decoder's own sample type (copied)"""
    n = _i32(frame + _OFF_FRAME_NB_SAMPLES)
    fmt = _i32(frame + _OFF_FRAME_FORMAT)
    if fmt not in _FMT:
        raise DecodeError(f"unexpected sample format {fmt}")
    dt, planar = _FMT[fmt]
    itemsize = np.dtype(dt).itemsize
    ext = C.c_void_p.from_address(frame + _OFF_FRAME_EXTENDED_DATA).value
    if n <= 0 or not ext:
        return np.zeros((0, channels), dtype=dt)
    planes = (C.c_void_p * (channels if planar else 1)).from_address(ext)
    if planar:
        cols = [np.frombuffer(C.string_at(planes[c], n * itemsize), dtype=dt) for c in range(channels)]
        return np.stack(cols, axis=1)
    return np.frombuffer(C.string_at(planes[0], n * channels * itemsize), dtype=dt).reshape(n, channels).copy()


def decode_audio(path: str) -> Tuple[np.ndarray, int]:
    """Decode the best audio stream of ``path``.  Returns (pcm, sample_rate): pcm is [n] (mono) or [n, channels];
    int16 when the decoder produces 16-bit integers, float32 (nominal +-1.0) otherwise — the two formats b2a_resample takes."""
    avu, avc, avf = _load()
    if not os.path.isfile(path):
        raise DecodeError(f"no such file: {path}")
    fmt = C.c_void_p(None)
    rc = avf.avformat_open_input(C.byref(fmt), os.fsencode(path), None, None)
    if rc < 0:
        raise DecodeError(f"cannot open {os.path.basename(path)}: {_err(avu, rc)}")
    cctx = C.c_void_p(None)
    pkt = C.c_void_p(None)
    frame = C.c_void_p(None)
    try:
        rc = avf.avformat_find_stream_info(fmt, None)
        if rc < 0:
            raise DecodeError(f"no stream info: {_err(avu, rc)}")
        dec = C.c_void_p(None)
        idx = avf.av_find_best_stream(fmt, AVMEDIA_TYPE_AUDIO, -1, -1, C.byref(dec), 0)
        if idx < 0 or not dec:
            raise DecodeError(f"no decodable audio stream: {_err(avu, idx)}")
        nb = C.c_uint32.from_address(fmt.value + _OFF_FMTCTX_NB_STREAMS).value
        streams = C.c_void_p.from_address(fmt.value + _OFF_FMTCTX_STREAMS).value
        if not (0 <= idx < nb <= 256) or not streams:
            raise DecodeError("unexpected AVFormatContext layout (nb_streams / streams)")
        st = C.c_void_p.from_address(streams + 8 * idx).value
        par = C.c_void_p.from_address(st + _OFF_STREAM_CODECPAR).value
        if _i32(st + _OFF_STREAM_INDEX) != idx or not par or _i32(par) != AVMEDIA_TYPE_AUDIO:      # AVCodecParameters.codec_type comes first
            raise DecodeError("unexpected AVStream layout (index / codecpar)")
        cctx = C.c_void_p(avc.avcodec_alloc_context3(dec))
        if not cctx:
            raise DecodeError("avcodec_alloc_context3 failed")
        rc = avc.avcodec_parameters_to_context(cctx, par)
        if rc >= 0:
            rc = avc.avcodec_open2(cctx, dec, None)
        if rc < 0:
            raise DecodeError(f"cannot open the decoder: {_err(avu, rc)}")
        pkt = C.c_void_p(avc.av_packet_alloc())
        frame = C.c_void_p(avu.av_frame_alloc())
        chunks = []
        rate = channels = 0

        def drain():
            nonlocal rate, channels
            while True:
                r = avc.avcodec_receive_frame(cctx, frame)
                if r == AVERROR_EAGAIN or r == AVERROR_EOF:
                    return
                if r < 0:
                    raise DecodeError(f"decode error: {_err(avu, r)}")
                if not rate:                      # the decoder knows its output only after the first frame (e.g. AAC SBR doubles the rate)
                    v = C.c_int64(0)
                    if avu.av_opt_get_int(cctx, b"ar", 0, C.byref(v)) < 0 or v.value <= 0:
                        raise DecodeError("decoder reports no sample rate")
                    rate = int(v.value)
                    lay = _AVChannelLayout()
                    if avu.av_opt_get_chlayout(cctx, b"ch_layout", 0, C.byref(lay)) < 0 or not (1 <= lay.nb_channels <= 8):
                        raise DecodeError("decoder reports no channel layout")
                    channels = int(lay.nb_channels)
                chunks.append(_frame_to_array(frame.value, channels))
                avu.av_frame_unref(frame)

        while True:
            r = avf.av_read_frame(fmt, pkt)
            if r < 0:
                break
            if _i32(pkt.value + _OFF_PKT_STREAM_INDEX) == idx:
                r = avc.avcodec_send_packet(cctx, pkt)
                if r < 0 and r != AVERROR_EAGAIN:
                    avc.av_packet_unref(pkt)
                    raise DecodeError(f"decode error: {_err(avu, r)}")
                drain()
            avc.av_packet_unref(pkt)
        avc.avcodec_send_packet(cctx, None)       # flush
        drain()
        if not chunks or not rate:
            raise DecodeError("the stream holds no audio frames")
        a = np.concatenate(chunks, axis=0)
        if a.dtype == np.int16:
            pcm = a
        elif a.dtype == np.uint8:
            pcm = ((a.astype(np.float32) - 128.0) / 128.0).astype(np.float32)
        elif a.dtype == np.int32:
            pcm = (a.astype(np.float64) / 2147483648.0).astype(np.float32)
        else:
            pcm = a.astype(np.float32)
        if channels > 2:
            raise DecodeError(f"{channels} channels: only mono and stereo are in scope")
        return (np.ascontiguousarray(pcm[:, 0]) if channels == 1 else np.ascontiguousarray(pcm)), rate
    finally:
        if frame:
            avu.av_frame_free(C.byref(frame))
        if pkt:
            avc.av_packet_free(C.byref(pkt))
        if cctx:
            avc.avcodec_free_context(C.byref(cctx))
        avf.avformat_close_input(C.byref(fmt))
