"""Host-side mirror of the reference's hot-path helpers — same names, arguments and error behaviour as
``AudioProcessor.convert_to_wav`` / ``AudioProcessor.preprocess_audio``
(/root/reference/app/services/audio_processor.py:901-930 and :305-314), so the Flask job workers can
call them unchanged (INTEGRATION.md shows the two-line patch).

    fe = AudioFrontend()
    wav = fe.convert_to_wav("/tmp/job/meeting.wav")      # -> "/tmp/job/meeting.wav" rewritten as 16 kHz mono s16
    wav = fe.preprocess_audio(wav)                       # -> silence-stripped WAV (new path)
    mel = fe.log_mel(wav)                                # what model.transcribe computes first
"""
from __future__ import annotations

import logging
import os
import subprocess
import threading
from typing import List, Optional, Tuple, Union

import numpy as np

from . import _abi, ops, wavio, whisper_audio


class AudioFrontend:
    """GPU implementation of the convert -> strip-silence -> log-mel front-end.

    Silence parameters are pydub's (min_silence_len / silence_thresh / keep_silence / seek_step)."""

    def __init__(self, min_silence_len: int = 1000, silence_thresh: float = -40, keep_silence: Union[int, bool] = 200,
                 seek_step: int = 1, n_mels: int = 80, strip_silence: bool = True, device: Optional[str] = None):
        self.min_silence_len = min_silence_len
        self.silence_thresh = silence_thresh
        self.keep_silence = keep_silence
        self.seek_step = seek_step
        self.n_mels = n_mels
        self.strip_silence = strip_silence
        self.device = device
        # The reference calls these helpers from a 3-thread ThreadPoolExecutor (audio_processor.py:56) on ONE shared instance:
        # everything a call leaves behind (the kept table, the cached plans and their buffers) is per thread.
        self._tls = threading.local()

    @property
    def last_segments(self) -> Optional[List[List[int]]]:
        """kept [start_ms, end_ms] of the last preprocess_audio / process_pcm made by THIS thread (a job runs on one worker
        thread from convert to remap, audio_processor.py:1150-1179); prefer the table the call itself returns"""
        return getattr(self._tls, "last_segments", None)

    def _thread_plans(self) -> dict:
        plans = getattr(self._tls, "plans", None)
        if plans is None:
            plans = self._tls.plans = {}                        # PipelinePlan per clip shape (buffers reused across this thread's calls)
        return plans

    # -- convert_to_wav (audio_processor.py:901-930) ----------------------------------------------------
    def convert_to_wav(self, input_path: str) -> str:
        """轉換檔案為 WAV 格式 (16kHz 單聲道) — returns ``<dirname>/<stem>.wav`` and overwrites it (`-y`)."""
        logging.info(f"🔄 轉換檔案格式為 WAV: {os.path.basename(input_path)}")
        output_dir = os.path.dirname(input_path)
        output_filename = f"{os.path.splitext(os.path.basename(input_path))[0]}.wav"
        output_path = os.path.join(output_dir, output_filename)
        try:
            pcm, rate = wavio.read_audio(input_path)           # WAV parsed here; m4a / mp3 / ... decoded on the host by libavcodec
            s16, _, _ = ops.resample(self._to_device(pcm), rate, ops.SAMPLE_RATE)
            wavio.write_wav_s16(output_path, s16.cpu().numpy(), ops.SAMPLE_RATE)
            logging.info(f"✅ 檔案轉換完成: {output_filename}")
            return output_path
        except (wavio.UnsupportedAudio, OSError, RuntimeError) as e:
            # the reference lets subprocess.CalledProcessError escape (:928-930); keep the caller's except clause working
            logging.error(f"❌ 檔案轉換失敗: {e}")
            raise subprocess.CalledProcessError(1, ["b2a_resample", input_path], stderr=str(e).encode()) from e

    # -- preprocess_audio (audio_processor.py:305-314; silence strip intended at :1046) ----------------------------
    def preprocess_audio(self, audio_path: str) -> str:
        """預處理音頻: ensure 16 kHz mono WAV, then strip silence.  Returns the same path when nothing changed, else
        a new ``<stem>.trimmed.wav`` (the caller removes the old file when the path differs, :1048-1051)."""
        return self.preprocess_audio_segments(audio_path)[0]

    def preprocess_audio_segments(self, audio_path: str) -> Tuple[str, List[List[int]]]:
        """preprocess_audio that also RETURNS the kept [start_ms, end_ms] table (what remap_segments needs), instead of
        leaving it on the shared instance."""
        logging.info(f"🔄 預處理音頻: {os.path.basename(audio_path)}")
        if not audio_path.lower().endswith(".wav"):
            audio_path = self.convert_to_wav(audio_path)
        if not self.strip_silence:
            return audio_path, []
        pcm, rate = wavio.read_audio(audio_path)
        if rate != ops.SAMPLE_RATE or pcm.ndim != 1 or pcm.dtype != np.int16:
            audio_path = self.convert_to_wav(audio_path)
            pcm, rate = wavio.read_audio(audio_path)
        dev = self._to_device(pcm)
        res = ops.detect(dev, rate, self.min_silence_len, self.silence_thresh, self.keep_silence, self.seek_step)
        kept = res.kept
        self._tls.last_segments = kept
        if res.n_keep == len(pcm):
            logging.info("✅ 音頻預處理完成")
            return audio_path, kept
        out = ops.compact(dev, res, rate)[: res.n_keep]
        new_path = os.path.splitext(audio_path)[0] + ".trimmed.wav"
        wavio.write_wav_s16(new_path, out.cpu().numpy(), rate)
        logging.info("✅ 音頻預處理完成")
        return new_path, kept

    # -- what model.transcribe computes first (audio_processor.py:1076) ------------------------------------
    def log_mel(self, audio_path: str, padding: int = whisper_audio.N_SAMPLES):
        return whisper_audio.log_mel_spectrogram(audio_path, n_mels=self.n_mels, padding=padding, device=self.device)

    # -- all three without touching the filesystem in between ---------------------------------------------------
    def process_pcm(self, pcm, in_rate: int, padding: int = 0, copy: bool = True) -> Tuple["object", "object", List[List[int]]]:
        """(trimmed 16 kHz s16 PCM, log-mel, kept [start_ms, end_ms]) for raw PCM, one fused device pass.

        The work buffers are cached per thread and clip shape; the results are returned as fresh tensors (``copy=True``),
        so earlier results stay valid when the next clip of the same shape runs.  ``copy=False`` returns views of the
        thread's plan buffers, valid until this thread's next call with the same shape."""
        x = self._to_device(pcm).contiguous()
        ch = 1 if x.dim() == 1 else int(x.shape[1])
        key = (int(x.shape[0]), int(in_rate), ch, x.dtype, self.n_mels, int(padding), str(x.device))
        plans = self._thread_plans()
        plan = plans.get(key)
        if plan is None:
            if len(plans) >= 4:
                plans.clear()
            plan = plans[key] = ops.PipelinePlan(key[0], in_rate, ch, x.dtype, n_mels=self.n_mels, padding=padding,
                                                 device=x.device)
        r = plan.run(x, trim=self.strip_silence, min_silence_len=self.min_silence_len, silence_thresh=self.silence_thresh,
                     keep_silence=self.keep_silence, seek_step=self.seek_step)
        kept = r.kept
        self._tls.last_segments = kept
        if copy:
            return r.pcm.clone(), r.mel.clone(), kept
        return r.pcm, r.mel, kept

    # -- process_audio's front half on a file (audio_processor.py:1039-1080) in ONE device pass ------------------------
    def process_audio_file(self, input_path: str, padding: int = whisper_audio.N_SAMPLES) -> Tuple[str, "object", List[List[int]]]:
        """convert_to_wav + preprocess_audio + the log-mel of ``model.transcribe`` for one upload, fused: the file is read
        (WAV parsed, anything else decoded on the host) once, ``b2a_pipeline`` runs once, the trimmed 16 kHz mono WAV is
        written once.  Returns (wav path for the diarizer, log-mel [n_mels, T] on the device, kept [start_ms, end_ms]).

        The WAV lands where the two helpers would have left it: ``<stem>.wav`` for a non-WAV upload (convert_to_wav's name,
        :904-906), ``<stem>.trimmed.wav`` for a WAV input that had to change (rate, channels, sample format or removed
        silence), the input path itself for a 16 kHz mono s16 WAV from which nothing was removed.  Errors surface as
        ``subprocess.CalledProcessError`` like convert_to_wav's (:928-930)."""
        logging.info(f"🔄 預處理音頻: {os.path.basename(input_path)}")
        try:
            pcm, rate = wavio.read_audio(input_path)
        except (wavio.UnsupportedAudio, OSError, RuntimeError) as e:
            logging.error(f"❌ 檔案轉換失敗: {e}")
            raise subprocess.CalledProcessError(1, ["b2a_pipeline", input_path], stderr=str(e).encode()) from e
        out_pcm, mel, kept = self.process_pcm(pcm, rate, padding=padding, copy=False)
        is_wav = input_path.lower().endswith(".wav")
        canonical = is_wav and rate == ops.SAMPLE_RATE and pcm.ndim == 1 and pcm.dtype == np.int16
        if canonical and int(out_pcm.shape[0]) == len(pcm):
            out_path = input_path                                       # nothing to rewrite
        else:
            stem = os.path.splitext(input_path)[0]
            out_path = stem + (".trimmed.wav" if is_wav else ".wav")
            wavio.write_wav_s16(out_path, out_pcm.cpu().numpy(), ops.SAMPLE_RATE)
        logging.info("✅ 音頻預處理完成")
        return out_path, mel.clone(), kept

    def stream(self, n_in: int, in_rate: int, channels: int = 2, dtype=None, padding: int = 0, depth: int = 2) -> "ClipStream":
        """pipelined submit()/result() front-end for many same-shaped clips (uploads overlap kernels and downloads)"""
        return ClipStream(self, n_in, in_rate, channels, dtype, padding, depth)

    def _to_device(self, pcm):
        torch = ops.require_cuda()
        t = torch.from_numpy(np.ascontiguousarray(pcm)) if isinstance(pcm, np.ndarray) else pcm
        return t.to(self.device if self.device is not None else "cuda", non_blocking=True)


class ClipStream:
    """Pipelined host -> device -> host front-end for a stream of same-shaped clips (what a job worker pool feeds).

    ``submit(host_pcm)`` enqueues the H2D copy, the whole device path and the D2H of the info block on dedicated
    streams and returns a ticket at once; ``result(ticket)`` returns (trimmed int16 PCM, log-mel, kept ms ranges) as
    views of pinned host buffers.  With ``depth`` clips in flight the upload of clip i+1 overlaps the kernels and the
    download of clip i (PCIe is full duplex), so steady-state throughput is bounded by the larger transfer, not the sum.
    Results stay valid until ``depth`` further clips have been submitted."""

    def __init__(self, frontend: "AudioFrontend", n_in: int, in_rate: int, channels: int = 2, dtype=None, padding: int = 0,
                 depth: int = 2):
        torch = ops.require_cuda()
        self.torch, self.fe, self.depth = torch, frontend, int(depth)
        dtype = torch.int16 if dtype is None else dtype
        dev = torch.device(frontend.device if frontend.device is not None else f"cuda:{torch.cuda.current_device()}")
        self.dev, self.in_rate = dev, int(in_rate)
        shape = (int(n_in),) if channels == 1 else (int(n_in), int(channels))
        self.s_h2d, self.s_run, self.s_info, self.s_d2h = (torch.cuda.Stream(device=dev) for _ in range(4))
        self.slots = []
        for _ in range(self.depth):
            plan = ops.PipelinePlan(int(n_in), in_rate, channels, dtype, n_mels=frontend.n_mels, padding=padding, device=dev)
            self.slots.append(dict(
                plan=plan, dev_in=torch.empty(shape, dtype=dtype, device=dev),
                h_info=torch.empty(_abi.INFO_LEN, dtype=torch.int64).pin_memory(),
                h_pcm=torch.empty(plan.n16 + 64, dtype=torch.int16).pin_memory(),
                h_mel=torch.empty(plan.n_mels * max(plan.t_cap, 1), dtype=torch.float32).pin_memory(),
                h_kept=torch.empty((plan.cap, 2), dtype=torch.int32).pin_memory(),
                ev_in=torch.cuda.Event(), ev_run=torch.cuda.Event(), ev_info=torch.cuda.Event(), ev_out=torch.cuda.Event(),
                busy=False))
        self.n_submitted = 0

    def submit(self, host_pcm) -> int:
        torch, fe = self.torch, self.fe
        t = torch.from_numpy(np.ascontiguousarray(host_pcm)) if isinstance(host_pcm, np.ndarray) else host_pcm
        ticket = self.n_submitted
        sl = self.slots[ticket % self.depth]
        if sl["busy"]:
            raise RuntimeError("ClipStream: collect result(ticket) of the clip that used this slot before submitting more")
        with torch.cuda.stream(self.s_h2d):
            sl["dev_in"].copy_(t, non_blocking=True)
            sl["ev_in"].record()
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(sl["ev_in"])
            sl["plan"].run(sl["dev_in"], trim=fe.strip_silence, min_silence_len=fe.min_silence_len,
                           silence_thresh=fe.silence_thresh, keep_silence=fe.keep_silence, seek_step=fe.seek_step)
            sl["ev_run"].record()
        with torch.cuda.stream(self.s_info):          # own stream: a later clip's info copy must not queue ahead of this clip's download
            self.s_info.wait_event(sl["ev_run"])
            sl["h_info"].copy_(sl["plan"].info, non_blocking=True)
            sl["ev_info"].record()
        sl["busy"] = True
        self.n_submitted += 1
        return ticket

    def result(self, ticket: int):
        torch = self.torch
        sl = self.slots[ticket % self.depth]
        sl["ev_info"].synchronize()
        info = sl["h_info"].tolist()
        if info[_abi.INFO_OVERFLOW]:
            raise RuntimeError(f"more than cap={sl['plan'].cap} silence ranges; raise `cap`")
        n_keep, T, nk = int(info[_abi.INFO_N_KEEP]), int(info[_abi.INFO_N_FRAMES]), int(info[_abi.INFO_N_KEPT])
        plan = sl["plan"]
        with torch.cuda.stream(self.s_d2h):
            sl["h_pcm"][:n_keep].copy_(plan.pcm[:n_keep], non_blocking=True)
            sl["h_mel"][: plan.n_mels * T].copy_(plan.mel[: plan.n_mels * T], non_blocking=True)
            sl["h_kept"][:nk].copy_(plan.kept[:nk], non_blocking=True)
            sl["ev_out"].record()
        sl["ev_out"].synchronize()
        sl["busy"] = False
        return sl["h_pcm"][:n_keep], sl["h_mel"][: plan.n_mels * T].view(plan.n_mels, T), sl["h_kept"][:nk].tolist()

    def bytes_per_clip(self, ticket: int):
        """(h2d, d2h) bytes moved for a collected ticket — for benchmarks"""
        sl = self.slots[ticket % self.depth]
        info = sl["h_info"].tolist()
        plan = sl["plan"]
        return (sl["dev_in"].numel() * sl["dev_in"].element_size(),
                8 * _abi.INFO_LEN + 2 * int(info[_abi.INFO_N_KEEP]) + 4 * plan.n_mels * int(info[_abi.INFO_N_FRAMES]) + 8 * int(info[_abi.INFO_N_KEPT]))


def remap_time(t_trimmed_s: float, kept_ms: List[List[int]]) -> float:
    """Map a timestamp on the silence-stripped timeline back to the original recording:
    Whisper's segment start/end refer to the trimmed audio, diarization (audio_processor.py:1105-1145) to the
    original."""
    t_ms = t_trimmed_s * 1000.0
    acc = 0.0
    for s, e in kept_ms:
        d = e - s
        if t_ms <= acc + d:
            return (s + (t_ms - acc)) / 1000.0
        acc += d
    return (kept_ms[-1][1] / 1000.0) if kept_ms else t_trimmed_s


def remap_segments(segments, kept_ms: List[List[int]]):
    """Vectorised remap for Whisper's result: every ``{"start": s, "end": e, ...}`` of ``asr_result["segments"]`` (times on
    the silence-stripped timeline) gets original-recording times, so the speaker-overlap loop of the reference
    (audio_processor.py:1114-1145) compares like with like.  Returns new dicts; the input is not modified."""
    if not kept_ms:
        return [dict(s) for s in segments]
    k = np.asarray(kept_ms, dtype=np.float64)
    dur = k[:, 1] - k[:, 0]
    acc_end = np.cumsum(dur)                       # trimmed-timeline end of every kept range (ms)
    acc_start = acc_end - dur

    def one(t_s: float) -> float:
        t_ms = float(t_s) * 1000.0
        i = int(np.searchsorted(acc_end, t_ms, side="left"))
        if i >= len(k):
            return float(k[-1, 1]) / 1000.0
        return float(k[i, 0] + (t_ms - acc_start[i])) / 1000.0

    out = []
    for seg in segments:
        d = dict(seg)
        d["start"], d["end"] = one(seg["start"]), one(seg["end"])
        out.append(d)
    return out
