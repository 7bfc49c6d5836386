"""numpy-facing driver for tests/emu/libb2a_emu.so (TEST-ONLY CPU emulation of the kernels).

Calls the same C ABI as the product (include/b2a.h) but with HOST buffers, because the emulated
"device" is the CPU.  Used by `-m "not gpu"` tests only.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from audio_processor_b200 import _abi  # noqa: E402

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        from tests.emu import build_emu
        path = build_emu.build()
        _LIB = _abi.declare(C.CDLL(path))
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"b2a error {rc}: {lib().b2a_last_error().decode()}")


def _aligned(n, dtype, align=256):
    raw = np.zeros(n * np.dtype(dtype).itemsize + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * np.dtype(dtype).itemsize].view(dtype)


def resample(pcm, in_rate, out_rate=16000, want_f32=False, want_energy=False):
    L = lib()
    a = np.ascontiguousarray(pcm)
    ch = 1 if a.ndim == 1 else a.shape[1]
    n_in = a.shape[0]
    fmt = _abi.FMT_S16 if a.dtype == np.int16 else _abi.FMT_F32
    src = _aligned(a.size, a.dtype)
    src[:] = a.reshape(-1)
    n_out = L.b2a_resample_out_len(n_in, in_rate, out_rate)
    out = _aligned(n_out + 64, np.int16)
    outf = _aligned(n_out + 64, np.float32) if want_f32 else None
    ne = L.b2a_energy_len(n_out, out_rate) if want_energy else 0
    en = _aligned(ne + 8, np.uint64) if want_energy else None
    if en is not None:
        en[:] = np.uint64(0xDEADBEEF)
    _check(L.b2a_resample(_p(src), fmt, ch, in_rate, n_in, out_rate, _p(out), _p(outf), _p(en), None))
    res = [out[:n_out].copy()]
    if want_f32:
        res.append(outf[:n_out].copy())
    if want_energy:
        res.append(en[:ne].copy())
    return res[0] if len(res) == 1 else tuple(res)


def energy_ms(pcm, sr=16000):
    L = lib()
    a = _aligned(len(pcm), np.int16)
    a[:] = pcm
    ne = L.b2a_energy_len(len(pcm), sr)
    en = _aligned(ne + 8, np.uint64)
    _check(L.b2a_energy_ms(_p(a), len(pcm), sr, _p(en), None))
    return en[:ne].copy()


def detect(pcm, sr=16000, min_silence_len=1000, silence_thresh=-16.0, keep_silence=100, seek_step=1, cap=4096,
           energy=None):
    """returns dict(silent, nonsilent, kept, kept_off, info, compact)"""
    L = lib()
    a = _aligned(len(pcm) + 64, np.int16)
    a[:len(pcm)] = pcm
    n = len(pcm)
    en = energy_ms(pcm, sr) if energy is None else energy
    en_a = _aligned(len(en) + 8, np.uint64)
    en_a[:len(en)] = en
    if isinstance(keep_silence, bool):
        keep = -1 if keep_silence else 0
    else:
        keep = int(keep_silence)
    prm = _abi.SilenceParams(int(min_silence_len), keep, int(seek_step), 0, float(silence_thresh))
    sil = _aligned(cap * 2, np.int32); ns = _aligned(cap * 2, np.int32); kp = _aligned(cap * 2, np.int32)
    koff = _aligned(cap + 2, np.int64)
    info = _aligned(_abi.INFO_LEN, np.int64)
    wsb = L.b2a_silence_workspace_bytes(n, sr)
    ws = _aligned(wsb + 256, np.uint8)
    _check(L.b2a_detect_silence(_p(en_a), n, sr, C.byref(prm), cap, _p(sil), _p(ns), _p(kp), _p(koff), _p(info),
                                _p(ws), wsb, None))
    n_s, n_n, n_k = int(info[0]), int(info[1]), int(info[2])
    out = _aligned(n + 64, np.int16)
    _check(L.b2a_compact(_p(a), n, sr, _p(kp), _p(koff), _p(info), _p(out), n + 64, None))
    n_keep = int(info[_abi.INFO_N_KEEP])
    return dict(silent=sil[:2 * n_s].reshape(-1, 2).tolist(), nonsilent=ns[:2 * n_n].reshape(-1, 2).tolist(),
                kept=kp[:2 * n_k].reshape(-1, 2).tolist(), kept_off=koff[:n_k + 1].copy(), info=info.copy(),
                compact=out[:n_keep].copy())


def log_mel(audio, n_mels=80, padding=0, norm_mode=0):
    L = lib()
    a = np.ascontiguousarray(audio)
    batch = 1 if a.ndim == 1 else a.shape[0]
    n = a.shape[-1]
    fmt = _abi.FMT_S16 if a.dtype == np.int16 else _abi.FMT_F32
    src = _aligned(a.size + 64, a.dtype)
    src[:a.size] = a.reshape(-1)
    T = L.b2a_log_mel_frames(n, padding)
    out = _aligned(batch * n_mels * T + 64, np.float32)
    wsb = L.b2a_log_mel_workspace_bytes(batch, n, padding)
    ws = _aligned(wsb + 256, np.uint8)
    _check(L.b2a_log_mel(_p(src), fmt, batch, n, n, None, padding, n_mels, norm_mode, _p(out), None, _p(ws), wsb, None))
    r = out[:batch * n_mels * T].reshape(batch, n_mels, T).copy()
    return r[0] if a.ndim == 1 else r


def mel_windows(mel, n_frames=3000, content=None, seek0=0, stride=None, n_windows=None, half=False):
    L = lib()
    m = np.ascontiguousarray(mel, dtype=np.float32)
    n_mels, T = m.shape
    content = max(T - 3000, 0) if content is None else content
    stride = n_frames if stride is None else stride
    if n_windows is None:
        n_windows = max((content - seek0 + stride - 1) // stride, 0)
    src = _aligned(m.size + 64, np.float32)
    src[:m.size] = m.reshape(-1)
    dt = np.float16 if half else np.float32
    out = _aligned(n_windows * n_mels * n_frames + 64, dt)
    out[:] = 7
    _check(L.b2a_mel_windows(_p(src), n_mels, T, content, seek0, stride, n_windows, n_frames,
                             _abi.FMT_F16 if half else _abi.FMT_F32, _p(out), None))
    assert np.all(out[n_windows * n_mels * n_frames:] == 7)            # nothing written past the last window
    return out[:n_windows * n_mels * n_frames].reshape(n_windows, n_mels, n_frames).copy()


def pipeline(pcm, in_rate, n_mels=80, padding=0, trim=True, min_silence_len=1000, silence_thresh=-40.0,
             keep_silence=200, seek_step=1, cap=4096):
    L = lib()
    a = np.ascontiguousarray(pcm)
    ch = 1 if a.ndim == 1 else a.shape[1]
    n_in = a.shape[0]
    fmt = _abi.FMT_S16 if a.dtype == np.int16 else _abi.FMT_F32
    src = _aligned(a.size + 64, a.dtype)
    src[:a.size] = a.reshape(-1)
    n16 = L.b2a_resample_out_len(n_in, in_rate, 16000)
    pcm_out = _aligned(n16 + 64, np.int16)
    Tcap = (n16 + 16 + padding) // 160
    mel = _aligned(n_mels * Tcap + 64, np.float32)
    ns = _aligned(cap * 2, np.int32); kp = _aligned(cap * 2, np.int32)
    info = _aligned(_abi.INFO_LEN, np.int64)
    wsb = L.b2a_pipeline_workspace_bytes(n_in, in_rate, padding, cap)
    ws = _aligned(wsb + 256, np.uint8)
    prm = _abi.SilenceParams(int(min_silence_len), -1 if keep_silence is True else int(keep_silence), int(seek_step), 0,
                             float(silence_thresh))
    _check(L.b2a_pipeline(_p(src), fmt, ch, in_rate, n_in, C.byref(prm) if trim else None, n_mels, padding, cap,
                          _p(pcm_out), _p(mel), _p(ns), _p(kp), _p(info), _p(ws), wsb, None))
    n_keep = int(info[_abi.INFO_N_KEEP]); T = int(info[_abi.INFO_N_FRAMES])
    n_n, n_k = int(info[1]), int(info[2])
    return dict(pcm=pcm_out[:n_keep].copy(), mel=mel[:n_mels * T].reshape(n_mels, T).copy(),
                nonsilent=ns[:2 * n_n].reshape(-1, 2).tolist(), kept=kp[:2 * n_k].reshape(-1, 2).tolist(), info=info.copy())


def pipeline_batch(clips, n_mels=80, padding=0, trim=True, min_silence_len=1000, silence_thresh=-40.0, keep_silence=200, seek_step=1,
                   cap=4096):
    """clips: list of (pcm, in_rate); ONE b2a_pipeline_batch call; returns one dict per clip like pipeline()"""
    L = lib()
    descs = (_abi.ClipDesc * len(clips))()
    hold = []
    for d, (pcm, in_rate) in zip(descs, clips):
        a = np.ascontiguousarray(pcm)
        ch = 1 if a.ndim == 1 else a.shape[1]
        n_in = a.shape[0]
        src = _aligned(a.size + 64, a.dtype); src[:a.size] = a.reshape(-1)
        n16 = L.b2a_resample_out_len(n_in, in_rate, 16000)
        pcm_out = _aligned(n16 + 64, np.int16)
        mel = _aligned(n_mels * ((n16 + 16 + padding) // 160) + 64, np.float32)
        ns = _aligned(cap * 2, np.int32); kp = _aligned(cap * 2, np.int32); info = _aligned(_abi.INFO_LEN, np.int64)
        wsb = L.b2a_pipeline_workspace_bytes(n_in, in_rate, padding, cap)
        ws = _aligned(wsb + 256, np.uint8)
        d.d_in, d.fmt, d.channels, d.in_rate, d.n_in = src.ctypes.data, (_abi.FMT_S16 if a.dtype == np.int16 else _abi.FMT_F32), ch, in_rate, n_in
        d.d_pcm_out, d.d_mel_out, d.d_nonsilent_ms, d.d_kept_ms, d.d_info = pcm_out.ctypes.data, mel.ctypes.data, ns.ctypes.data, kp.ctypes.data, info.ctypes.data
        d.d_ws, d.ws_bytes = ws.ctypes.data, wsb
        hold.append((src, pcm_out, mel, ns, kp, info, ws))
    prm = _abi.SilenceParams(int(min_silence_len), -1 if keep_silence is True else int(keep_silence), int(seek_step), 0, float(silence_thresh))
    _check(L.b2a_pipeline_batch(descs, len(clips), C.byref(prm) if trim else None, n_mels, padding, cap, None))
    out = []
    for (src, pcm_out, mel, ns, kp, info, ws) in hold:
        n_keep = int(info[_abi.INFO_N_KEEP]); T = int(info[_abi.INFO_N_FRAMES]); n_n, n_k = int(info[1]), int(info[2])
        out.append(dict(pcm=pcm_out[:n_keep].copy(), mel=mel[:n_mels * T].reshape(n_mels, T).copy(),
                        nonsilent=ns[:2 * n_n].reshape(-1, 2).tolist(), kept=kp[:2 * n_k].reshape(-1, 2).tolist()))
    return out


def remap_times(times, kept_ms, sr=16000):
    """b2a_kept_offsets + b2a_remap_times on host buffers"""
    L = lib()
    k = np.asarray(kept_ms, dtype=np.int32).reshape(-1, 2)
    cap = max(len(k), 1)
    kp = _aligned(cap * 2, np.int32); kp[:k.size] = k.reshape(-1)
    info = _aligned(_abi.INFO_LEN, np.int64); info[:] = 0; info[_abi.INFO_N_KEPT] = len(k)
    koff = _aligned(cap + 1, np.int64)
    _check(L.b2a_kept_offsets(_p(kp), _p(info), sr, cap, _p(koff), None))
    t = _aligned(max(len(times), 1), np.float64); t[:len(times)] = np.asarray(times, dtype=np.float64)
    out = _aligned(max(len(times), 1), np.float64)
    _check(L.b2a_remap_times(_p(t), len(times), _p(kp), _p(koff), _p(info), sr, _p(out), None))
    return out[:len(times)].copy(), koff[:len(k) + 1].copy()
