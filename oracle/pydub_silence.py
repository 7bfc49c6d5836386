"""Literal restatement of pydub 0.25.1 silence detection on top of stdlib ``audioop``.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Why pydub: the reference advertises silence removal (README.md:17) and marks the
call site (/root/reference/app/services/audio_processor.py:1046 "音頻預處理 (移除靜音)",
:1047 ``preprocess_audio``), but ``preprocess_audio`` (:305-314) does not implement it.
BASELINE.json fixes the contract: "pydub-style silence parameters
min_silence_len/silence_thresh/keep_silence" with bit-exact segment boundaries.  pydub is
not a dependency of the reference and is absent from this image; pinned version restated
here: **pydub 0.25.1** — ``pydub/silence.py`` (detect_silence, detect_nonsilent,
split_on_silence), ``pydub/audio_segment.py`` (__len__, frame_count, _parse_position,
__getitem__ with its ≤2 ms zero-fill, rms, max_possible_amplitude, __add__ = append with
crossfade 0) and ``pydub/utils.py`` (db_to_float).  The per-window energy is the real C
routine pydub calls: CPython ``Modules/audioop.c: audioop_rms_impl`` via ``audioop.rms``.

Two implementations:
  * ``detect_silence`` / ``detect_nonsilent`` / ``split_on_silence`` — the literal loop
    (one ``audioop.rms`` per candidate window).  This is the oracle of record.
  * ``*_fast`` — exact-integer vectorised numpy (per-ms energies, prefix sums,
    ``sum(x^2) < n*(floor(thr)+1)^2``).  Proven identical to the literal loop by
    tests/test_oracle_silence.py; used where the literal loop would take minutes.
"""
from __future__ import annotations

import itertools
import warnings

import numpy as np

with warnings.catch_warnings():
    warnings.simplefilter("ignore", DeprecationWarning)
    import audioop  # stdlib C module (still present in CPython 3.12)


def db_to_float(db, using_amplitude=True):
    """pydub/utils.py: db_to_float"""
    db = float(db)
    if using_amplitude:
        return 10 ** (db / 20)
    return 10 ** (db / 10)


class TooManyMissingFrames(Exception):
    pass


class Segment:
    """Minimal stand-in for pydub.AudioSegment (raw PCM, 16-bit)."""

    def __init__(self, data, frame_rate: int = 16000, channels: int = 1, sample_width: int = 2):
        if isinstance(data, np.ndarray):
            if data.dtype != np.int16:
                raise TypeError("int16 PCM expected")
            data = np.ascontiguousarray(data).tobytes()
        self._data = bytes(data)
        self.frame_rate = int(frame_rate)
        self.channels = int(channels)
        self.sample_width = int(sample_width)
        self.frame_width = self.channels * self.sample_width

    # -- pydub/audio_segment.py -------------------------------------------------
    def frame_count(self, ms=None):
        if ms is not None:
            return ms * (self.frame_rate / 1000.0)
        return float(len(self._data) // self.frame_width)

    def __len__(self):
        return round(1000 * (self.frame_count() / self.frame_rate))

    def _parse_position(self, val):
        if val < 0:
            val = len(self) - abs(val)
        val = self.frame_count(ms=len(self)) if val == float("inf") else self.frame_count(ms=val)
        return int(val)

    def _spawn(self, data):
        return Segment(data, self.frame_rate, self.channels, self.sample_width)

    def __getitem__(self, millisecond):
        if not isinstance(millisecond, slice):
            raise TypeError("only slices are restated")
        start = millisecond.start if millisecond.start is not None else 0
        end = millisecond.stop if millisecond.stop is not None else len(self)
        start = min(start, len(self))
        end = min(end, len(self))
        start = self._parse_position(start) * self.frame_width
        end = self._parse_position(end) * self.frame_width
        data = self._data[start:end]
        expected_length = end - start
        missing_frames = (expected_length - len(data)) // self.frame_width
        if missing_frames:
            if missing_frames > self.frame_count(ms=2):
                raise TooManyMissingFrames(
                    "You should never be filling in more than 2 ms with silence here, "
                    "missing frames: %s" % missing_frames)
            silence = audioop.mul(data[:self.frame_width], self.sample_width, 0)
            data += (silence * missing_frames)
        return self._spawn(data)

    @property
    def rms(self):
        return audioop.rms(self._data, self.sample_width)

    @property
    def max_possible_amplitude(self):
        bits = self.sample_width * 8
        return (2 ** bits) / 2

    def __add__(self, other):  # append(crossfade=0)
        return self._spawn(self._data + other._data)

    def samples(self) -> np.ndarray:
        return np.frombuffer(self._data, dtype=np.int16)


# -- pydub/silence.py -------------------------------------------------------------
def detect_silence(audio_segment, min_silence_len=1000, silence_thresh=-16, seek_step=1):
    seg_len = len(audio_segment)
    if seg_len < min_silence_len:
        return []
    silence_thresh = db_to_float(silence_thresh) * audio_segment.max_possible_amplitude
    silence_starts = []
    last_slice_start = seg_len - min_silence_len
    slice_starts = range(0, last_slice_start + 1, seek_step)
    if last_slice_start % seek_step:
        slice_starts = itertools.chain(slice_starts, [last_slice_start])
    for i in slice_starts:
        audio_slice = audio_segment[i:i + min_silence_len]
        if audio_slice.rms <= silence_thresh:
            silence_starts.append(i)
    if not silence_starts:
        return []
    silent_ranges = []
    prev_i = silence_starts.pop(0)
    current_range_start = prev_i
    for silence_start_i in silence_starts:
        continuous = (silence_start_i == prev_i + seek_step)
        silence_has_gap = silence_start_i > (prev_i + min_silence_len)
        if not continuous and silence_has_gap:
            silent_ranges.append([current_range_start, prev_i + min_silence_len])
            current_range_start = silence_start_i
        prev_i = silence_start_i
    silent_ranges.append([current_range_start, prev_i + min_silence_len])
    return silent_ranges


def _nonsilent_from_silent(silent_ranges, len_seg):
    if not silent_ranges:
        return [[0, len_seg]]
    if silent_ranges[0][0] == 0 and silent_ranges[0][1] == len_seg:
        return []
    prev_end_i = 0
    nonsilent_ranges = []
    for start_i, end_i in silent_ranges:
        nonsilent_ranges.append([prev_end_i, start_i])
        prev_end_i = end_i
    if end_i != len_seg:
        nonsilent_ranges.append([prev_end_i, len_seg])
    if nonsilent_ranges[0] == [0, 0]:
        nonsilent_ranges.pop(0)
    return nonsilent_ranges


def detect_nonsilent(audio_segment, min_silence_len=1000, silence_thresh=-16, seek_step=1):
    silent_ranges = detect_silence(audio_segment, min_silence_len, silence_thresh, seek_step)
    return _nonsilent_from_silent(silent_ranges, len(audio_segment))


def _kept_from_nonsilent(nonsilent, keep_silence, len_seg):
    if isinstance(keep_silence, bool):
        keep_silence = len_seg if keep_silence else 0
    output_ranges = [[start - keep_silence, end + keep_silence] for (start, end) in nonsilent]
    for k in range(len(output_ranges) - 1):  # pairwise
        range_i, range_ii = output_ranges[k], output_ranges[k + 1]
        last_end = range_i[1]
        next_start = range_ii[0]
        if next_start < last_end:
            range_i[1] = (last_end + next_start) // 2
            range_ii[0] = range_i[1]
    return [[max(s, 0), min(e, len_seg)] for s, e in output_ranges]


def kept_ranges(audio_segment, min_silence_len=1000, silence_thresh=-16, keep_silence=100, seek_step=1):
    """The clamped [start_ms, end_ms] slices split_on_silence cuts (its last list comprehension)."""
    ns = detect_nonsilent(audio_segment, min_silence_len, silence_thresh, seek_step)
    return _kept_from_nonsilent(ns, keep_silence, len(audio_segment))


def split_on_silence(audio_segment, min_silence_len=1000, silence_thresh=-16, keep_silence=100, seek_step=1):
    return [audio_segment[s:e] for s, e in
            kept_ranges(audio_segment, min_silence_len, silence_thresh, keep_silence, seek_step)]


def strip_silence(audio_segment, **kw) -> np.ndarray:
    """Concatenate split_on_silence chunks with pydub's ``+`` (crossfade 0) → int16 samples."""
    chunks = split_on_silence(audio_segment, **kw)
    out = audio_segment._spawn(b"")
    for c in chunks:
        out = out + c
    return out.samples().copy()


# -- exact-integer vectorised form (mono int16 only) ----------------------------------------
def threshold_int(silence_thresh_db: float, n_window_samples: int) -> int:
    """rms <= thr  <=>  sum(x^2) < n * (floor(thr)+1)^2   (all integers)"""
    thr = db_to_float(silence_thresh_db) * 32768.0
    k = int(np.floor(thr)) + 1
    return int(n_window_samples) * k * k


def len_ms(n_frames: int, frame_rate: int) -> int:
    return round(1000 * (float(n_frames) / frame_rate))


def silent_starts_fast(x: np.ndarray, frame_rate, min_silence_len, silence_thresh, seek_step):
    """bool array over candidate start ms (dense 0..last), plus the candidate mask."""
    x = np.asarray(x)
    assert x.dtype == np.int16 and x.ndim == 1
    F = x.shape[0]
    L = len_ms(F, frame_rate)
    W = int(min_silence_len)
    if L < W:
        return None, None, L
    assert (frame_rate % 1000) == 0, "fast path restates integer samples-per-ms only"
    spm = frame_rate // 1000
    need = L * spm
    xx = np.zeros(need, dtype=np.int64)
    m = min(F, need)
    xx[:m] = x[:m]
    e = (xx * xx).reshape(L, spm).sum(axis=1)
    P = np.concatenate([[0], np.cumsum(e)])
    last = L - W
    E = P[W:W + last + 1] - P[0:last + 1]
    sil = E < threshold_int(silence_thresh, W * spm)
    cand = np.zeros(last + 1, dtype=bool)
    cand[::seek_step] = True
    cand[last] = True
    return sil & cand, cand, L


def detect_silence_fast(x, frame_rate=16000, min_silence_len=1000, silence_thresh=-16, seek_step=1):
    flags, _, L = silent_starts_fast(x, frame_rate, min_silence_len, silence_thresh, seek_step)
    if flags is None:
        return []
    starts = np.flatnonzero(flags)
    if starts.size == 0:
        return []
    W = int(min_silence_len)
    d = np.diff(starts)
    brk = (d != seek_step) & (d > W)
    first = np.concatenate([[0], np.flatnonzero(brk) + 1])
    lastk = np.concatenate([np.flatnonzero(brk), [starts.size - 1]])
    return [[int(starts[a]), int(starts[b]) + W] for a, b in zip(first, lastk)]


def detect_nonsilent_fast(x, frame_rate=16000, min_silence_len=1000, silence_thresh=-16, seek_step=1):
    sr = detect_silence_fast(x, frame_rate, min_silence_len, silence_thresh, seek_step)
    return _nonsilent_from_silent(sr, len_ms(len(x), frame_rate))


def kept_ranges_fast(x, frame_rate=16000, min_silence_len=1000, silence_thresh=-16, keep_silence=100,
                     seek_step=1):
    ns = detect_nonsilent_fast(x, frame_rate, min_silence_len, silence_thresh, seek_step)
    return _kept_from_nonsilent(ns, keep_silence, len_ms(len(x), frame_rate))


def strip_silence_fast(x, frame_rate=16000, **kw) -> np.ndarray:
    """Same samples as strip_silence(Segment(x)), incl. pydub's zero-fill of a rounded-up tail."""
    kept = kept_ranges_fast(x, frame_rate, **kw)
    spm = frame_rate / 1000.0
    out = []
    for s, e in kept:
        a, b = int(s * spm), int(e * spm)
        piece = x[a:min(b, len(x))]
        if len(piece) < b - a:
            piece = np.concatenate([piece, np.zeros(b - a - len(piece), dtype=np.int16)])
        out.append(piece)
    if not out:
        return np.zeros(0, dtype=np.int16)
    return np.concatenate(out)
