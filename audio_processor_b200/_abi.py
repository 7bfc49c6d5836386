"""ctypes declarations of the C ABI in include/b2a.h (signatures only; no library is loaded here)."""
from __future__ import annotations

import ctypes as C

FMT_S16 = 0
FMT_F32 = 1
FMT_F16 = 2   # b2a_mel_windows output only
NORM_WHISPER = 0
NORM_PER_CLIP = 1

INFO_N_SILENT = 0
INFO_N_NONSILENT = 1
INFO_N_KEPT = 2
INFO_OVERFLOW = 3
INFO_LEN_MS = 4
INFO_N_KEEP = 5
INFO_N_FRAMES = 6
INFO_LEN = 8

OK = 0
EINVAL = -1
EUNSUPPORTED = -2
EWORKSPACE = -3
ECUDA = -4


class SilenceParams(C.Structure):
    _fields_ = [("min_silence_len", C.c_int32), ("keep_silence", C.c_int32), ("seek_step", C.c_int32),
                ("reserved", C.c_int32), ("silence_thresh", C.c_double)]


P = C.c_void_p


class ClipDesc(C.Structure):
    """b2a_clip_desc (include/b2a.h)"""
    _fields_ = [("d_in", C.c_void_p), ("fmt", C.c_int32), ("channels", C.c_int32), ("in_rate", C.c_int32), ("reserved", C.c_int32),
                ("n_in", C.c_int64), ("d_pcm_out", C.c_void_p), ("d_mel_out", C.c_void_p), ("d_nonsilent_ms", C.c_void_p),
                ("d_kept_ms", C.c_void_p), ("d_info", C.c_void_p), ("d_ws", C.c_void_p), ("ws_bytes", C.c_size_t)]


SIGNATURES = {
    "b2a_version": (C.c_int, []),
    "b2a_last_error": (C.c_char_p, []),
    "b2a_launch_count": (C.c_int64, []),
    "b2a_fir_schedule": (C.c_int, [C.c_int, P, C.c_int]),
    "b2a_resample_out_len": (C.c_int64, [C.c_int64, C.c_int, C.c_int]),
    "b2a_resample_ntaps": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int)]),
    "b2a_resample_taps": (C.c_int, [C.c_int, C.c_int, P, C.c_size_t]),
    "b2a_energy_len": (C.c_int64, [C.c_int64, C.c_int]),
    "b2a_resample": (C.c_int, [P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, P, P, P, P]),
    "b2a_energy_ms": (C.c_int, [P, C.c_int64, C.c_int, P, P]),
    "b2a_silence_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "b2a_detect_silence": (C.c_int, [P, C.c_int64, C.c_int, C.POINTER(SilenceParams), C.c_int32, P, P, P, P, P, P,
                                     C.c_size_t, P]),
    "b2a_compact": (C.c_int, [P, C.c_int64, C.c_int, P, P, P, P, C.c_int64, P]),
    "b2a_log_mel_frames": (C.c_int64, [C.c_int64, C.c_int64]),
    "b2a_log_mel_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64]),
    "b2a_log_mel": (C.c_int, [P, C.c_int, C.c_int64, C.c_int64, C.c_int64, P, C.c_int64, C.c_int, C.c_int, P, P, P,
                              C.c_size_t, P]),
    "b2a_mel_windows": (C.c_int, [P, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, P, P]),
    "b2a_mel_filters": (C.c_int, [C.c_int, P, C.c_size_t]),
    "b2a_pipeline_batch": (C.c_int, [C.POINTER(ClipDesc), C.c_int, C.POINTER(SilenceParams), C.c_int, C.c_int64, C.c_int32, P]),
    "b2a_kept_offsets": (C.c_int, [P, P, C.c_int, C.c_int32, P, P]),
    "b2a_remap_times": (C.c_int, [P, C.c_int64, P, P, P, C.c_int, P, P]),
    "b2a_pipeline_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int, C.c_int64, C.c_int32]),
    "b2a_pipeline": (C.c_int, [P, C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(SilenceParams), C.c_int, C.c_int64,
                               C.c_int32, P, P, P, P, P, P, C.c_size_t, P]),
}


def declare(lib: C.CDLL) -> C.CDLL:
    """Attach restype/argtypes for every symbol include/b2a.h declares; raises if one is missing."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError => the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    return lib
