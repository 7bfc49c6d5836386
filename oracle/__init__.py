"""oracle/ — CPU restatements of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``audio_processor_b200/`` may import this package.  The only
permitted users are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` — and there only
as the checker or as the CPU arm being timed, never as the product.

The reference (dong881/audio-processor) orchestrates the hot path but the
arithmetic lives in three third-party engines that are NOT vendored under
/root/reference:

* conversion   — ``ffmpeg -ar 16000 -ac 1 -c:a pcm_s16le``
                 (app/services/audio_processor.py:912-923) → FFmpeg libswresample.
                 ``swr_ref.py`` drives the real FFmpeg 8.0.1 libswresample 6.1.100
                 that ships in this image (opencv_python_headless.libs);
                 ``resample_oracle.py`` is a float64 restatement of its default
                 Kaiser windowed-sinc polyphase resampler.
* silence trim — intended at app/services/audio_processor.py:1046-1047
                 ("移除靜音"), contract = pydub 0.25.1 ``pydub/silence.py``.
                 ``pydub_silence.py`` restates it literally on top of the stdlib
                 ``audioop.rms`` (the C routine pydub itself calls).
* log-mel      — inside ``model.transcribe`` (app/services/audio_processor.py:1076-1080)
                 → openai-whisper ``whisper/audio.py:log_mel_spectrogram``.
                 ``whisper_logmel.py`` restates it on ``torch.stft`` in float64.

Parity pinning status: the reference has NO tests or golden vectors for this
path ("parity unpinned" by the reference itself).  The oracles are pinned
instead against (a) the real libswresample engine run in this image,
(b) stdlib audioop, (c) torch.stft and transformers' WhisperFeatureExtractor,
and (d) the known-answer vectors of SURVEY.md Appendix A.5, frozen under
tests/golden/ by tests/golden/make_golden.py.
"""
