"""Tensor-level entry points over the C ABI (include/b2a.h).

PyTorch is plumbing here: it owns device memory (caching allocator), streams and
torch.distributed.  Every function enqueues on torch's current CUDA stream and returns
device tensors without synchronising; results whose size is data dependent (kept samples,
frame count, range tables) are exposed through small result objects that synchronise
lazily when the host asks for them.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Union

from . import _abi
from ._lib import check, lib, require_cuda

SAMPLE_RATE = 16000
DEFAULT_SEG_CAP = 8192


def _stream(torch) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t) -> Optional[C.c_void_p]:
    return None if t is None else C.c_void_p(t.data_ptr())


def _fmt(torch, t) -> int:
    if t.dtype == torch.int16:
        return _abi.FMT_S16
    if t.dtype == torch.float32:
        return _abi.FMT_F32
    raise TypeError(f"PCM must be int16 or float32, got {t.dtype}")


def _as_cuda_pcm(torch, pcm, device=None):
    import numpy as np
    if isinstance(pcm, np.ndarray):
        pcm = torch.from_numpy(np.ascontiguousarray(pcm))
    if not torch.is_tensor(pcm):
        raise TypeError("expected a torch.Tensor or numpy array")
    if not pcm.is_cuda:
        pcm = pcm.to(device if device is not None else "cuda", non_blocking=True)
    return pcm.contiguous()


_graph_launches = 0     # kernels launched through replays of captured pipeline graphs (the library only counts its own launch calls)


def launch_count() -> int:
    """CUDA kernels of libb2a launched by this process so far: the library's own launch calls plus the kernels inside
    every replayed pipeline graph (PipelinePlan.run(graph=True))."""
    return int(lib().b2a_launch_count()) + _graph_launches


def resample_out_len(n_in: int, in_rate: int, out_rate: int = SAMPLE_RATE) -> int:
    return int(lib().b2a_resample_out_len(int(n_in), int(in_rate), int(out_rate)))


def resample(pcm, in_rate: int, out_rate: int = SAMPLE_RATE, *, want_s16: bool = True, want_f32: bool = False,
             want_energy: bool = False):
    """[n] or [n, C] int16/float32 PCM -> mono at out_rate.  Returns (s16, f32, energy_ms) with None for
    outputs not requested.  Equivalent of ``ffmpeg -ar out_rate -ac 1 -c:a pcm_s16le`` on raw PCM
    (reference: app/services/audio_processor.py:912-923)."""
    torch = require_cuda()
    x = _as_cuda_pcm(torch, pcm)
    ch = 1 if x.dim() == 1 else int(x.shape[1])
    n_in = int(x.shape[0])
    with torch.cuda.device(x.device):
        n_out = resample_out_len(n_in, in_rate, out_rate)
        s16 = torch.empty(n_out + 16, dtype=torch.int16, device=x.device) if want_s16 else None
        f32 = torch.empty(n_out + 16, dtype=torch.float32, device=x.device) if want_f32 else None
        en = None
        if want_energy:
            ne = int(lib().b2a_energy_len(n_out, out_rate))
            en = torch.empty(ne + 2, dtype=torch.int64, device=x.device)
        check(lib().b2a_resample(_ptr(x), _fmt(torch, x), ch, int(in_rate), n_in, int(out_rate), _ptr(s16), _ptr(f32),
                                 _ptr(en), _stream(torch)))
    return (s16[:n_out] if s16 is not None else None, f32[:n_out] if f32 is not None else None,
            en[:-2] if en is not None else None)


def _params(min_silence_len, silence_thresh, keep_silence, seek_step) -> _abi.SilenceParams:
    if isinstance(keep_silence, bool):
        keep = -1 if keep_silence else 0
    else:
        keep = int(keep_silence)
        if keep < 0:
            raise ValueError("keep_silence must be >= 0 or a bool")
    return _abi.SilenceParams(int(min_silence_len), keep, int(seek_step), 0, float(silence_thresh))


@dataclass
class SilenceResult:
    """Device-side result of detect(); host views synchronise on first access."""
    silent_ms: "object"
    nonsilent_ms: "object"
    kept_ms: "object"
    kept_off: "object"
    info: "object"
    cap: int
    _host: Optional[dict] = None

    def _sync(self) -> dict:
        if self._host is None:
            info = self.info.cpu().tolist()
            if info[_abi.INFO_OVERFLOW]:
                raise RuntimeError(f"more than cap={self.cap} silence ranges; raise `cap`")
            ns, nn, nk = info[_abi.INFO_N_SILENT], info[_abi.INFO_N_NONSILENT], info[_abi.INFO_N_KEPT]
            self._host = dict(
                info=info,
                silent=self.silent_ms[:ns].cpu().tolist(),
                nonsilent=self.nonsilent_ms[:nn].cpu().tolist(),
                kept=self.kept_ms[:nk].cpu().tolist(),
            )
        return self._host

    @property
    def silent(self) -> List[List[int]]:
        return self._sync()["silent"]

    @property
    def nonsilent(self) -> List[List[int]]:
        return self._sync()["nonsilent"]

    @property
    def kept(self) -> List[List[int]]:
        return self._sync()["kept"]

    @property
    def len_ms(self) -> int:
        return int(self._sync()["info"][_abi.INFO_LEN_MS])

    @property
    def n_keep(self) -> int:
        return int(self._sync()["info"][_abi.INFO_N_KEEP])


def energy_ms(pcm16, sample_rate: int = SAMPLE_RATE):
    torch = require_cuda()
    x = _as_cuda_pcm(torch, pcm16)
    if x.dtype != torch.int16 or x.dim() != 1:
        raise TypeError("energy_ms expects mono int16 PCM")
    n = int(x.shape[0])
    with torch.cuda.device(x.device):
        ne = int(lib().b2a_energy_len(n, sample_rate))
        en = torch.empty(ne + 2, dtype=torch.int64, device=x.device)
        if n > 0:
            check(lib().b2a_energy_ms(_ptr(x), n, int(sample_rate), _ptr(en), _stream(torch)))
    return en[:ne]


def detect(pcm16, sample_rate: int = SAMPLE_RATE, min_silence_len: int = 1000, silence_thresh: float = -16,
           keep_silence: Union[int, bool] = 100, seek_step: int = 1, *, energy=None, cap: int = DEFAULT_SEG_CAP) -> SilenceResult:
    """pydub detect_silence / detect_nonsilent / split_on_silence ranges for mono int16 PCM, on the GPU."""
    torch = require_cuda()
    x = _as_cuda_pcm(torch, pcm16)
    if x.dtype != torch.int16 or x.dim() != 1:
        raise TypeError("detect expects mono int16 PCM")
    n = int(x.shape[0])
    prm = _params(min_silence_len, silence_thresh, keep_silence, seek_step)
    with torch.cuda.device(x.device):
        if energy is None:
            energy = energy_ms(x, sample_rate)
        dev = x.device
        # one zeroed block carved into the five tables (one memset launch instead of five: the call is launch bound)
        blk = torch.zeros(3 * cap + (cap + 2) + _abi.INFO_LEN, dtype=torch.int64, device=dev)
        sil, ns, kp = (blk[i * cap:(i + 1) * cap].view(torch.int32).view(cap, 2) for i in range(3))
        koff = blk[3 * cap:4 * cap + 2]
        info = blk[4 * cap + 2:]
        wsb = int(lib().b2a_silence_workspace_bytes(n, sample_rate))
        ws = torch.empty(wsb + 256, dtype=torch.uint8, device=dev)
        if energy.numel() == 0:
            energy = torch.zeros(2, dtype=torch.int64, device=dev)
        check(lib().b2a_detect_silence(_ptr(energy), n, int(sample_rate), C.byref(prm), int(cap), _ptr(sil), _ptr(ns),
                                       _ptr(kp), _ptr(koff), _ptr(info), _ptr(ws), wsb, _stream(torch)))
    return SilenceResult(sil, ns, kp, koff, info, cap)


def compact(pcm16, res: SilenceResult, sample_rate: int = SAMPLE_RATE):
    """Concatenate the kept ranges (pydub `+`).  Returns (buffer, res.info): the first info[N_KEEP] samples of
    `buffer` are valid; slicing needs the host value (res.n_keep)."""
    torch = require_cuda()
    x = _as_cuda_pcm(torch, pcm16)
    n = int(x.shape[0])
    with torch.cuda.device(x.device):
        out = torch.empty(n + 64, dtype=torch.int16, device=x.device)
        if n > 0:
            check(lib().b2a_compact(_ptr(x), n, int(sample_rate), _ptr(res.kept_ms), _ptr(res.kept_off), _ptr(res.info),
                                    _ptr(out), n + 64, _stream(torch)))
    return out


def log_mel(audio, n_mels: int = 80, padding: int = 0, *, per_clip_max: bool = False):
    """Whisper log-mel of [n] or [B, n] float32 (+-1.0) or int16 audio -> [n_mels, T] / [B, n_mels, T] float32."""
    torch = require_cuda()
    x = _as_cuda_pcm(torch, audio)
    if x.dim() not in (1, 2):
        raise ValueError("audio must be 1-D or 2-D")
    batch = 1 if x.dim() == 1 else int(x.shape[0])
    n = int(x.shape[-1])
    with torch.cuda.device(x.device):
        T = int(lib().b2a_log_mel_frames(n, padding))
        out = torch.empty((batch, n_mels, T), dtype=torch.float32, device=x.device)
        wsb = int(lib().b2a_log_mel_workspace_bytes(batch, n, padding))
        ws = torch.empty(wsb + 256, dtype=torch.uint8, device=x.device)
        check(lib().b2a_log_mel(_ptr(x), _fmt(torch, x), batch, n, n, None, int(padding), int(n_mels),
                                _abi.NORM_PER_CLIP if per_clip_max else _abi.NORM_WHISPER, _ptr(out), None, _ptr(ws), wsb,
                                _stream(torch)))
    return out[0] if x.dim() == 1 else out


def mel_windows(mel, n_frames: int = 3000, *, content_frames: Optional[int] = None, seek0: int = 0, stride: Optional[int] = None,
                n_windows: Optional[int] = None, dtype=None):
    """All encoder windows of a [n_mels, T] float32 mel in one launch: window w = ``mel[:, s : s + n_frames]`` with
    s = seek0 + w * stride, frames at or beyond ``content_frames`` zero (whisper's pad_or_trim), cast to ``dtype``
    (torch.float32 or torch.float16).  Defaults follow ``whisper.transcribe``: the mel carries 3000 frames of padding
    (content_frames = T - 3000), windows advance by n_frames and cover the content.  Returns [n_windows, n_mels, n_frames]."""
    torch = require_cuda()
    if not (mel.is_cuda and mel.dtype == torch.float32 and mel.dim() == 2):
        raise ValueError("mel must be a 2-D float32 CUDA tensor [n_mels, T]")
    mel = mel.contiguous()
    n_mels, T = int(mel.shape[0]), int(mel.shape[1])
    content = max(T - 3000, 0) if content_frames is None else int(content_frames)
    stride = int(n_frames) if stride is None else int(stride)
    if n_windows is None:
        if stride <= 0:
            raise ValueError("stride must be positive when n_windows is not given")
        n_windows = max((content - int(seek0) + stride - 1) // stride, 0)
    dtype = torch.float32 if dtype is None else dtype
    if dtype not in (torch.float32, torch.float16):
        raise ValueError("dtype must be torch.float32 or torch.float16")
    with torch.cuda.device(mel.device):
        out = torch.empty((int(n_windows), n_mels, int(n_frames)), dtype=dtype, device=mel.device)
        check(lib().b2a_mel_windows(_ptr(mel), n_mels, T, content, int(seek0), stride, int(n_windows), int(n_frames),
                                    _abi.FMT_F16 if dtype == torch.float16 else _abi.FMT_F32, _ptr(out), _stream(torch)))
    return out


def remap_times(times, kept_ms, info, kept_off=None, sample_rate: int = SAMPLE_RATE):
    """Seconds on the silence-stripped timeline -> seconds on the original recording, on the device (b2a_remap_times).
    ``times``: float64 tensor / array of any shape; ``kept_ms`` [cap, 2] int32 and ``info`` int64[8] as produced by
    detect() / PipelinePlan (device tensors); ``kept_off`` int64[cap + 1] if the caller has it (SilenceResult.kept_off),
    else it is derived from the range table."""
    torch = require_cuda()
    dev = kept_ms.device
    t = torch.as_tensor(times, dtype=torch.float64).to(dev).contiguous()
    out = torch.empty_like(t)
    with torch.cuda.device(dev):
        cap = int(kept_ms.shape[0])
        if kept_off is None:
            kept_off = torch.empty(cap + 1, dtype=torch.int64, device=dev)
            check(lib().b2a_kept_offsets(_ptr(kept_ms), _ptr(info), int(sample_rate), cap, _ptr(kept_off), _stream(torch)))
        check(lib().b2a_remap_times(_ptr(t), int(t.numel()), _ptr(kept_ms), _ptr(kept_off), _ptr(info), int(sample_rate), _ptr(out),
                                    _stream(torch)))
    return out


class PipelinePlan:
    """Pre-allocated buffers for running the whole path on clips of one shape (no allocation per call)."""

    def __init__(self, n_in: int, in_rate: int, channels: int, dtype, n_mels: int = 80, padding: int = 0,
                 cap: int = DEFAULT_SEG_CAP, device=None):
        torch = require_cuda()
        self.torch = torch
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.n_in, self.in_rate, self.channels, self.dtype = int(n_in), int(in_rate), int(channels), dtype
        self.n_mels, self.padding, self.cap = int(n_mels), int(padding), int(cap)
        self.fmt = _abi.FMT_S16 if dtype == torch.int16 else _abi.FMT_F32
        with torch.cuda.device(self.device):
            self.n16 = resample_out_len(self.n_in, self.in_rate, SAMPLE_RATE)
            self.t_cap = (self.n16 + 16 + self.padding) // 160
            self.pcm = torch.empty(self.n16 + 64, dtype=torch.int16, device=self.device)
            self.mel = torch.empty(self.n_mels * max(self.t_cap, 1), dtype=torch.float32, device=self.device)
            self.nonsilent = torch.zeros((self.cap, 2), dtype=torch.int32, device=self.device)
            self.kept = torch.zeros((self.cap, 2), dtype=torch.int32, device=self.device)
            self.info = torch.zeros(_abi.INFO_LEN, dtype=torch.int64, device=self.device)
            self.ws_bytes = int(lib().b2a_pipeline_workspace_bytes(self.n_in, self.in_rate, self.padding, self.cap))
            self.ws = torch.empty(self.ws_bytes + 512, dtype=torch.uint8, device=self.device)
            off = (-self.ws.data_ptr()) % 256
            self._ws_ptr = C.c_void_p(self.ws.data_ptr() + off)
        self._graphs = {}        # (input pointer, silence parameters) -> captured CUDA graph of the clip's launches (run(graph=True))
        self._warm = set()

    def run(self, pcm, *, trim: bool = True, min_silence_len: int = 1000, silence_thresh: float = -40,
            keep_silence: Union[int, bool] = 200, seek_step: int = 1, graph: bool = False) -> "PipelineResult":
        """Enqueue the whole path for one clip on the current stream.  graph=True replays a CUDA graph of the same
        b2a_pipeline call (captured on the second use of an input buffer with the same parameters): one launch instead
        of eight, for callers that keep their input in a fixed device buffer (ClipStream slots, the benchmark)."""
        global _graph_launches
        torch = self.torch
        x = pcm
        if (not x.is_cuda) or x.dtype != self.dtype or not x.is_contiguous():
            raise ValueError("PipelinePlan.run expects a contiguous CUDA tensor of the planned dtype")
        ch = 1 if x.dim() == 1 else int(x.shape[1])
        if int(x.shape[0]) != self.n_in or ch != self.channels:
            raise ValueError("clip shape differs from the plan")
        prm = _params(min_silence_len, silence_thresh, keep_silence, seek_step) if trim else None

        def launch():
            check(lib().b2a_pipeline(_ptr(x), self.fmt, ch, self.in_rate, self.n_in, C.byref(prm) if prm is not None else None,
                                     self.n_mels, self.padding, self.cap, _ptr(self.pcm), _ptr(self.mel),
                                     _ptr(self.nonsilent), _ptr(self.kept), _ptr(self.info), self._ws_ptr, self.ws_bytes,
                                     _stream(torch)))

        with torch.cuda.device(self.device):
            if not graph:
                launch()
                return PipelineResult(self)
            # keyed on the RESOLVED parameter struct: keep_silence=True (keep everything, -1) and keep_silence=1 (1 ms) hash
            # alike as Python values but bake different parameters into the captured graph
            key = (x.data_ptr(), bool(trim)) + ((prm.min_silence_len, prm.keep_silence, prm.seek_step, prm.silence_thresh) if prm is not None else ())
            g = self._graphs.get(key)
            if g is None:
                if key not in self._warm:                # first use: run eagerly (builds the device tables, sets kernel attributes)
                    self._warm.add(key)
                    launch()
                    return PipelineResult(self)
                g = torch.cuda.CUDAGraph()
                cap_stream = torch.cuda.Stream(device=self.device)
                cap_stream.wait_stream(torch.cuda.current_stream())
                n0 = int(lib().b2a_launch_count())
                with torch.cuda.graph(g, stream=cap_stream, capture_error_mode="thread_local"):   # other threads (NCCL watchdog, job workers) keep using CUDA
                    launch()
                torch.cuda.current_stream().wait_stream(cap_stream)
                g = (g, int(lib().b2a_launch_count()) - n0)      # the graph and the number of kernels it holds
                self._graphs[key] = g
                _graph_launches -= g[1]                           # the capture only recorded them; the replay below runs them
            g[0].replay()
            _graph_launches += g[1]
        return PipelineResult(self)


class PipelineBatch:
    """A rank's shard of clips through ONE b2a_pipeline_batch call per pass (one CUDA-graph replay with graph=True).

    ``plans[i]`` holds clip i's buffers (clips may differ in length, rate and channel count; n_mels, padding and cap are
    shared).  The library forks the clips over its internal streams and joins them back into the current stream."""

    def __init__(self, plans: List["PipelinePlan"]):
        if not plans:
            raise ValueError("PipelineBatch needs at least one plan")
        p0 = plans[0]
        for p in plans:
            if (p.n_mels, p.padding, p.cap, p.device) != (p0.n_mels, p0.padding, p0.cap, p0.device):
                raise ValueError("plans of one batch share n_mels, padding, cap and device")
        self.plans = plans
        self.torch = p0.torch
        self._graphs = {}
        self._warm = set()

    def run(self, inputs, *, trim: bool = True, min_silence_len: int = 1000, silence_thresh: float = -40,
            keep_silence: Union[int, bool] = 200, seek_step: int = 1, graph: bool = False) -> List["PipelineResult"]:
        global _graph_launches
        torch = self.torch
        if len(inputs) != len(self.plans):
            raise ValueError("one input per plan")
        descs = (_abi.ClipDesc * len(inputs))()
        for d, p, x in zip(descs, self.plans, inputs):
            if (not x.is_cuda) or x.dtype != p.dtype or not x.is_contiguous() or int(x.shape[0]) != p.n_in:
                raise ValueError("every input must be a contiguous CUDA tensor of its plan's shape and dtype")
            d.d_in, d.fmt, d.channels, d.in_rate, d.n_in = x.data_ptr(), p.fmt, p.channels, p.in_rate, p.n_in
            d.d_pcm_out, d.d_mel_out = p.pcm.data_ptr(), p.mel.data_ptr()
            d.d_nonsilent_ms, d.d_kept_ms, d.d_info = p.nonsilent.data_ptr(), p.kept.data_ptr(), p.info.data_ptr()
            d.d_ws, d.ws_bytes = p._ws_ptr.value, p.ws_bytes
        prm = _params(min_silence_len, silence_thresh, keep_silence, seek_step) if trim else None
        p0 = self.plans[0]

        def launch():
            check(lib().b2a_pipeline_batch(descs, len(inputs), C.byref(prm) if prm is not None else None, p0.n_mels, p0.padding,
                                           p0.cap, _stream(torch)))

        with torch.cuda.device(p0.device):
            if not graph:
                launch()
            else:
                key = tuple(x.data_ptr() for x in inputs) + (bool(trim),) + \
                    ((prm.min_silence_len, prm.keep_silence, prm.seek_step, prm.silence_thresh) if prm is not None else ())
                g = self._graphs.get(key)
                if g is None and key not in self._warm:      # first use: eager (device tables, kernel attributes, internal streams)
                    self._warm.add(key)
                    launch()
                else:
                    if g is None:
                        g = torch.cuda.CUDAGraph()
                        cap_stream = torch.cuda.Stream(device=p0.device)
                        cap_stream.wait_stream(torch.cuda.current_stream())
                        n0 = int(lib().b2a_launch_count())
                        with torch.cuda.graph(g, stream=cap_stream, capture_error_mode="thread_local"):
                            launch()
                        torch.cuda.current_stream().wait_stream(cap_stream)
                        g = (g, int(lib().b2a_launch_count()) - n0)
                        self._graphs[key] = g
                        _graph_launches -= g[1]
                    g[0].replay()
                    _graph_launches += g[1]
        return [PipelineResult(p) for p in self.plans]


class PipelineResult:
    """View over a PipelinePlan's output buffers (valid until the plan runs again)."""

    def __init__(self, plan: PipelinePlan):
        self.plan = plan
        self._info = None

    def info(self) -> List[int]:
        if self._info is None:
            self._info = self.plan.info.cpu().tolist()       # the one synchronising read of the path
            if self._info[_abi.INFO_OVERFLOW]:
                raise RuntimeError(f"more than cap={self.plan.cap} silence ranges; raise `cap`")
        return self._info

    @property
    def n_keep(self) -> int:
        return int(self.info()[_abi.INFO_N_KEEP])

    @property
    def n_frames(self) -> int:
        return int(self.info()[_abi.INFO_N_FRAMES])

    @property
    def pcm(self):
        """trimmed 16 kHz mono int16 PCM (device)"""
        return self.plan.pcm[: self.n_keep]

    @property
    def mel(self):
        """[n_mels, T] float32 Whisper log-mel (device)"""
        T = self.n_frames
        return self.plan.mel[: self.plan.n_mels * T].view(self.plan.n_mels, T)

    @property
    def nonsilent(self) -> List[List[int]]:
        return self.plan.nonsilent[: int(self.info()[_abi.INFO_N_NONSILENT])].cpu().tolist()

    @property
    def kept(self) -> List[List[int]]:
        return self.plan.kept[: int(self.info()[_abi.INFO_N_KEPT])].cpu().tolist()


def pipeline(pcm, in_rate: int, *, n_mels: int = 80, padding: int = 0, trim: bool = True, min_silence_len: int = 1000,
             silence_thresh: float = -40, keep_silence: Union[int, bool] = 200, seek_step: int = 1,
             cap: int = DEFAULT_SEG_CAP) -> PipelineResult:
    """One-shot convenience wrapper: convert -> strip silence -> log-mel for a single clip."""
    torch = require_cuda()
    x = _as_cuda_pcm(torch, pcm)
    ch = 1 if x.dim() == 1 else int(x.shape[1])
    plan = PipelinePlan(int(x.shape[0]), in_rate, ch, x.dtype, n_mels=n_mels, padding=padding, cap=cap, device=x.device)
    return plan.run(x, trim=trim, min_silence_len=min_silence_len, silence_thresh=silence_thresh,
                    keep_silence=keep_silence, seek_step=seek_step)
