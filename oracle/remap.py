"""Trimmed-timeline -> original-timeline timestamp map, restated on the CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The consumer in the reference is the speaker-overlap loop of ``process_audio``
(/root/reference/app/services/audio_processor.py:1114-1145): it compares Whisper's ``segment["start"] / ["end"]``
with pyannote's diarization times.  The reference never trims (``preprocess_audio`` :305-314 only converts), so both
are on the same clock there; once silence is stripped (:1046-1051, the step this repo implements with pydub
``split_on_silence`` semantics) Whisper's times refer to the concatenation of the kept ranges.  The inverse of that
concatenation is fixed by pydub's own semantics (oracle/pydub_silence.py: kept ranges [s_i, e_i) in ms, concatenated in
order with crossfade 0):

    trimmed position t_ms in [A_i, A_i + (e_i - s_i)]  ->  s_i + (t_ms - A_i),     A_i = sum_{j<i} (e_j - s_j)

A time exactly on a cut belongs to the END of the earlier range (a segment that ends on the cut ends there in the
original too); times behind the last range map to its end.  Pure Python / float64, one timestamp at a time.
"""
from __future__ import annotations

from typing import List, Sequence


def remap_time(t_trimmed_s: float, kept_ms: Sequence[Sequence[int]]) -> float:
    if not kept_ms:
        return float(t_trimmed_s)
    t_ms = float(t_trimmed_s) * 1000.0
    acc = 0.0
    for s, e in kept_ms:
        d = float(e - s)
        if t_ms <= acc + d:
            return (float(s) + (t_ms - acc)) / 1000.0
        acc += d
    return float(kept_ms[-1][1]) / 1000.0


def remap_segments(segments: List[dict], kept_ms: Sequence[Sequence[int]]) -> List[dict]:
    """every {"start", "end", ...} of Whisper's ``result["segments"]`` with original-recording times (new dicts)"""
    out = []
    for seg in segments:
        d = dict(seg)
        d["start"] = remap_time(seg["start"], kept_ms)
        d["end"] = remap_time(seg["end"], kept_ms)
        out.append(d)
    return out
