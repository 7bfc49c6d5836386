/* b2a.h — C ABI of libb2a.so, the B200-native (sm_100a) audio front-end.
 *
 * Drop-in boundary for ONE hot path of dong881/audio-processor:
 *     any PCM input -> 16 kHz mono s16 -> strip silence -> Whisper log-mel.
 * The reference is pure Python and has no FFI of its own; each entry point below
 * replaces the native engine the reference reaches through a dependency, and is
 * what a ctypes/cffi binding on the reference side would bind.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch/C++ types.
 *   - every `d_*` pointer is DEVICE memory owned by the caller (e.g. the PyTorch
 *     caching allocator), including workspaces; the library never allocates or
 *     frees on the hot path and never synchronises: all work is enqueued on the
 *     caller's stream (`b2a_stream_t` is a cudaStream_t).  Small immutable tables
 *     (filter bank, twiddles, mel weights) are built once per process and device
 *     on first use (mutex-guarded) and live until exit.
 *   - return value: 0 = ok, negative = error (B2A_E*); message via b2a_last_error()
 *     (thread-local).  No C++ exception crosses the boundary.
 *   - re-entrant: no mutable global state besides the table cache; safe to call
 *     from the reference's 3 ThreadPoolExecutor workers
 *     (app/services/audio_processor.py:56) with per-call streams.
 */
#ifndef B2A_H_
#define B2A_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* b2a_stream_t; /* == cudaStream_t */

#define B2A_FMT_S16 0 /* interleaved signed 16-bit PCM */
#define B2A_FMT_F32 1 /* interleaved float32 PCM, nominal range +-1.0 */
#define B2A_FMT_F16 2 /* OUTPUT of b2a_mel_windows only: IEEE binary16 */

#define B2A_OK 0
#define B2A_EINVAL (-1)       /* bad argument */
#define B2A_EUNSUPPORTED (-2) /* valid request outside the implemented domain */
#define B2A_EWORKSPACE (-3)   /* workspace too small */
#define B2A_ECUDA (-4)        /* CUDA runtime error (message has the detail) */

#define B2A_NORM_WHISPER 0 /* floor at (max over the WHOLE call) - 8, whisper/audio.py */
#define B2A_NORM_PER_CLIP 1 /* floor per batch row (HF WhisperFeatureExtractor) */

/* layout of the int64 `d_info[B2A_INFO_LEN]` block written by the silence / pipeline calls */
#define B2A_INFO_N_SILENT 0    /* number of [start,end) ms ranges in d_silent_ms */
#define B2A_INFO_N_NONSILENT 1 /* ... in d_nonsilent_ms */
#define B2A_INFO_N_KEPT 2      /* ... in d_kept_ms */
#define B2A_INFO_OVERFLOW 3    /* !=0: more than `cap` ranges, tables truncated */
#define B2A_INFO_LEN_MS 4      /* pydub len(segment) in ms */
#define B2A_INFO_N_KEEP 5      /* kept samples (= length of the compacted PCM) */
#define B2A_INFO_N_FRAMES 6    /* log-mel frames T written (pipeline / device-length log-mel) */
#define B2A_INFO_RESERVED 7
#define B2A_INFO_LEN 8

int b2a_version(void);
const char* b2a_last_error(void);
/* diagnostics: number of CUDA kernels this library has launched in this process so far */
int64_t b2a_launch_count(void);
/* diagnostics (host only): the tcgen05 FIR's k-step issue schedule for in_rate 44100 / 48000 (csrc/fir_tc_common.cuh).
 * out[0] = items, out[1] = ring pieces per tile, out[2] = k-steps per block, out[3 + b] = first column chunk of block b
 * (b < 10), out[16 + j] = item j.  Returns the number of words written, or a negative error. */
int b2a_fir_schedule(int in_rate, uint32_t* out, int capacity_words);

/* ------------------------------------------------------------------------------------------
 * Conversion: replaces the libswresample work behind
 *   ffmpeg -y -i IN -ar 16000 -ac 1 -c:a pcm_s16le OUT
 * (/root/reference/app/services/audio_processor.py:912-923, convert_to_wav :901-930).
 * Fused PCM decode (s16 / f32) + stereo->mono downmix + polyphase Kaiser windowed-sinc
 * resampler with libswresample's default design (filter_size 32, cutoff 0.97, beta 9),
 * reflect head / symmetric tail, lrintf + clip quantisation.
 * ------------------------------------------------------------------------------------------ */

/* number of output samples for n_in input frames: exactly what libswresample returns for a one-shot swr_convert of the
 * whole input followed by a flush (= the sample count of the WAV `ffmpeg -i IN -ar out_rate` writes):
 * ceil((n_in - taps/2 + R) * L / M), L/M = out_rate/in_rate reduced, R = the library's flush reflection (taps/2 or one
 * less, from its buffering state; csrc/fir_design.h restates it).  0 for clips too short to fill the filter. */
int64_t b2a_resample_out_len(int64_t n_in, int in_rate, int out_rate);

/* filter geometry; returns taps per phase (<0 on error) and stores the phase count L */
int b2a_resample_ntaps(int in_rate, int out_rate, int* phases);

/* copy the float32 filter bank [phases][ntaps] the kernels use into HOST memory (for audits) */
int b2a_resample_taps(int in_rate, int out_rate, float* h_taps, size_t capacity_floats);

/* number of per-millisecond energy slots for n_out samples at out_rate (out_rate % 1000 == 0) */
int64_t b2a_energy_len(int64_t n_out, int out_rate);

/* d_in: [n_in, channels] interleaved, fmt B2A_FMT_*; channels 1 or 2.
 * d_out_s16 [n_out] and/or d_out_f32 [n_out] (pre-quantisation, nominal +-1) — either may be NULL.
 * d_energy_ms (optional, may be NULL): uint64[b2a_energy_len] = sum of squares of the QUANTISED s16
 * output per millisecond (last partial ms zero-extended) — the exact-integer input of the silence
 * detector, produced for free in the resampler epilogue. */
int b2a_resample(const void* d_in, int fmt, int channels, int in_rate, int64_t n_in, int out_rate,
                 int16_t* d_out_s16, float* d_out_f32, uint64_t* d_energy_ms, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Silence detection / trimming: the step the reference intends at
 * /root/reference/app/services/audio_processor.py:1046-1047 ("音頻預處理 (移除靜音)",
 * preprocess_audio :305-314).  Semantics = pydub 0.25.1 pydub/silence.py:
 * detect_silence / detect_nonsilent / split_on_silence (+ concatenation), bit-exact.
 * ------------------------------------------------------------------------------------------ */
typedef struct b2a_silence_params {
    int32_t min_silence_len; /* ms, pydub default 1000 */
    int32_t keep_silence;    /* ms >= 0; -1 == True (keep everything); pydub default 100 */
    int32_t seek_step;       /* ms, pydub default 1 */
    int32_t reserved;
    double silence_thresh;   /* dBFS, pydub default -16 */
} b2a_silence_params;

/* per-ms energies of an existing 16-bit mono PCM buffer (when it did not come from b2a_resample) */
int b2a_energy_ms(const int16_t* d_pcm, int64_t n, int sample_rate, uint64_t* d_energy_ms,
                  b2a_stream_t stream);

size_t b2a_silence_workspace_bytes(int64_t n_samples, int sample_rate);

/* d_energy_ms: uint64[b2a_energy_len(n_samples, sample_rate)].
 * d_silent_ms / d_nonsilent_ms / d_kept_ms: int32[cap][2] = [start_ms, end_ms) ranges
 *   (detect_silence, detect_nonsilent, and the clamped slices split_on_silence cuts).
 * d_kept_off: int64[cap+1] = sample offset of each kept range in the compacted output
 *   (d_kept_off[n_kept] == n_keep).
 * d_info: int64[B2A_INFO_LEN].  Any of the three range tables may be NULL (not wanted). */
int b2a_detect_silence(const uint64_t* d_energy_ms, int64_t n_samples, int sample_rate,
                       const b2a_silence_params* params, int32_t cap, int32_t* d_silent_ms,
                       int32_t* d_nonsilent_ms, int32_t* d_kept_ms, int64_t* d_kept_off,
                       int64_t* d_info, void* d_ws, size_t ws_bytes, b2a_stream_t stream);

/* stream compaction: concatenate pcm[kept ranges] (pydub `+`, crossfade 0; a tail that pydub
 * zero-fills is zero-filled).  d_out must hold n_samples + sample_rate/1000 samples. */
int b2a_compact(const int16_t* d_pcm, int64_t n_samples, int sample_rate, const int32_t* d_kept_ms,
                const int64_t* d_kept_off, const int64_t* d_info, int16_t* d_out, int64_t out_capacity,
                b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Whisper front-end: replaces whisper.audio.log_mel_spectrogram(audio, n_mels, padding)
 * (openai-whisper whisper/audio.py), reached from model.transcribe at
 * /root/reference/app/services/audio_processor.py:1076-1080.
 * Hann-400 STFT at hop 160 (reflect-padded, last frame dropped), 201 -> n_mels slaney mel,
 * log10(max(.,1e-10)), floor at max-8, (x+4)/4.
 * ------------------------------------------------------------------------------------------ */
int64_t b2a_log_mel_frames(int64_t n, int64_t padding); /* (n + padding) / 160 */

size_t b2a_log_mel_workspace_bytes(int64_t batch, int64_t n, int64_t padding);

/* d_audio: [batch] rows of `n` samples, consecutive rows `row_stride` samples apart; fmt S16
 *   (value/32768, whisper.load_audio) or F32.
 * d_n (optional, batch must be 1): device int64 holding the actual sample count (<= n), e.g. the
 *   N_KEEP slot of a d_info block, so trimming and log-mel chain without a host round trip.
 * d_out: float32 [batch][n_mels][T], T = (n_actual + padding)/160, T contiguous (whisper layout).
 * d_frames_out (optional): device int64 receiving T.
 * n_mels: 80 or 128.
 * Limit: one clip's [n_mels][T] block must stay below 2^31 values (46 hours of audio at 128 mels);
 *   longer rows return B2A_EUNSUPPORTED (split them: the floor couples only what one call sees). */
int b2a_log_mel(const void* d_audio, int fmt, int64_t batch, int64_t n, int64_t row_stride,
                const int64_t* d_n, int64_t padding, int n_mels, int norm_mode, float* d_out,
                int64_t* d_frames_out, void* d_ws, size_t ws_bytes, b2a_stream_t stream);

/* The encoder windows whisper.transcribe cuts from the mel (openai-whisper whisper/transcribe.py:
 * `mel_segment = mel[:, seek : seek + segment_size]`, `pad_or_trim(mel_segment, N_FRAMES).to(device).to(dtype)`;
 * reached from /root/reference/app/services/audio_processor.py:1076-1080), for every window of a uniform grid at once:
 * window w starts at frame seek0 + w * stride; frames at or beyond content_frames (transcribe: T - 3000, the mel was
 * computed with padding = 480000) read as zero.
 * d_mel: float32 [n_mels][T] (b2a_log_mel / b2a_pipeline output). d_out: [n_windows][n_mels][n_frames] of out_fmt
 * (B2A_FMT_F32, or B2A_FMT_F16 rounded to nearest even like torch's .half()). */
int b2a_mel_windows(const float* d_mel, int n_mels, int64_t T, int64_t content_frames, int64_t seek0, int64_t stride,
                    int n_windows, int n_frames, int out_fmt, void* d_out, b2a_stream_t stream);

/* the float32 [n_mels][201] slaney filterbank the kernel uses (HOST copy, for audits) */
int b2a_mel_filters(int n_mels, float* h_filters, size_t capacity_floats);

/* ------------------------------------------------------------------------------------------
 * Whole path for one clip, no host synchronisation between stages:
 *   convert_to_wav (:901-930) -> preprocess_audio (:305-314, silence trim) -> log_mel (:1076)
 * ------------------------------------------------------------------------------------------ */
size_t b2a_pipeline_workspace_bytes(int64_t n_in, int in_rate, int64_t padding, int32_t cap);

/* d_pcm_out: int16, capacity >= b2a_resample_out_len(...) + 16: trimmed 16 kHz mono PCM (the WAV
 *   payload the reference keeps for diarization, audio_processor.py:1105).
 * d_mel_out: float32, capacity >= n_mels * ((n_out16k + padding)/160): [n_mels][T].
 * d_kept_ms / d_nonsilent_ms: int32[cap][2]; d_info: int64[B2A_INFO_LEN].
 * params == NULL skips trimming (pure convert + log-mel). */
int b2a_pipeline(const void* d_in, int fmt, int channels, int in_rate, int64_t n_in,
                 const b2a_silence_params* params, int n_mels, int64_t padding, int32_t cap,
                 int16_t* d_pcm_out, float* d_mel_out, int32_t* d_nonsilent_ms, int32_t* d_kept_ms,
                 int64_t* d_info, void* d_ws, size_t ws_bytes, b2a_stream_t stream);

/* Many clips in one call (a rank's shard of a corpus: SURVEY.md section 8(b) "batched clip descriptors").  `clips` is a HOST
 * array; every clip brings its own outputs and workspace (b2a_pipeline_workspace_bytes for its n_in), all clips share
 * the silence parameters, n_mels, padding and cap.  The clips are independent problems: the library spreads them over
 * a small pool of internal streams forked from and joined back into `stream` (created once per device on first use,
 * like the tables), so the latency-bound stages of one clip (silence ranges, the log-mel floor pass) overlap the
 * bandwidth-bound stages of its neighbours.  Capturable into a CUDA graph on `stream` after one eager call.
 * Results are identical to n_clips calls of b2a_pipeline. */
typedef struct b2a_clip_desc {
    const void* d_in;        /* [n_in, channels] interleaved PCM */
    int32_t fmt;             /* B2A_FMT_S16 / B2A_FMT_F32 */
    int32_t channels;
    int32_t in_rate;
    int32_t reserved;
    int64_t n_in;
    int16_t* d_pcm_out;      /* as b2a_pipeline */
    float* d_mel_out;
    int32_t* d_nonsilent_ms;
    int32_t* d_kept_ms;
    int64_t* d_info;
    void* d_ws;
    size_t ws_bytes;
} b2a_clip_desc;

int b2a_pipeline_batch(const b2a_clip_desc* clips, int n_clips, const b2a_silence_params* params, int n_mels,
                       int64_t padding, int32_t cap, b2a_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Timestamps on the silence-stripped timeline -> the original recording: what the speaker-overlap loop of
 * process_audio (/root/reference/app/services/audio_processor.py:1114-1145) needs once preprocess_audio
 * (:1046-1051) has removed silence — Whisper's segment start / end refer to the trimmed audio, diarization to the
 * original.  d_t_in / d_t_out: double[n] seconds; d_kept_ms / d_kept_off / d_info: the tables of
 * b2a_detect_silence or b2a_pipeline (kept_off in samples of the trimmed clip at sample_rate).  A time exactly on a
 * cut maps to the end of the earlier kept range; times behind the last range map to its end.
 * ------------------------------------------------------------------------------------------ */
/* d_kept_off[k] (int64[cap + 1]) = samples of the trimmed clip before kept range k, from the range table alone
 * (b2a_detect_silence returns it; b2a_pipeline keeps its copy inside the workspace) */
int b2a_kept_offsets(const int32_t* d_kept_ms, const int64_t* d_info, int sample_rate, int32_t cap, int64_t* d_kept_off,
                     b2a_stream_t stream);

int b2a_remap_times(const double* d_t_in, int64_t n, const int32_t* d_kept_ms, const int64_t* d_kept_off,
                    const int64_t* d_info, int sample_rate, double* d_t_out, b2a_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B2A_H_ */
