"""audio_processor_b200 — B200-native (sm_100a) implementation of dong881/audio-processor's data-parallel
front-end: any PCM -> 16 kHz mono s16 -> strip silence (pydub semantics) -> Whisper log-mel.

Only the hot path lives here: hand-written CUDA kernels behind the C ABI in include/b2a.h
(`csrc/`, built in-tree as libb2a.so) and the host-side mirror of the reference's call surface:

    service.AudioFrontend.convert_to_wav / preprocess_audio      (app/services/audio_processor.py:901-930, :305-314)
    silence.detect_silence / detect_nonsilent / split_on_silence  (pydub.silence)
    whisper_audio.log_mel_spectrogram / load_audio / pad_or_trim  (whisper.audio)

There is no CPU fallback: importing is cheap, but every op raises without libb2a.so and a CUDA device.
"""
from . import _abi  # noqa: F401

__version__ = "0.1.0"
__all__ = ["ops", "silence", "whisper_audio", "service", "wavio", "synth", "sharding"]
