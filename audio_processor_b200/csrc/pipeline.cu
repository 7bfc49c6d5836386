// pipeline.cu — the whole hot path for one clip, enqueued on one stream with no host round trip:
//   convert_to_wav (app/services/audio_processor.py:901-930)  -> resample + downmix + per-ms energy
//   preprocess_audio (:305-314, intended silence strip :1046)  -> pydub-exact ranges + compaction
//   whisper log_mel_spectrogram (via transcribe, :1076-1080)   -> STFT + mel + log + floor
// Data-dependent sizes (kept samples, frame count) stay on the device: the log-mel kernels read the
// N_KEEP slot of d_info, so nothing waits for the host between stages.
#include "b2a_common.cuh"

#include <mutex>

namespace b2a {

int resample_launch(const void* d_in, int fmt, int channels, int in_rate, i64 n_in, int out_rate, int16_t* d_out_s16,
                    float* d_out_f32, u64* d_energy, cudaStream_t stream);
int silence_launch(const u64* d_energy, i64 n_samples, int sample_rate, const b2a_silence_params* prm, int cap,
                   int32_t* d_silent, int32_t* d_nonsilent, int32_t* d_kept, i64* d_kept_off, i64* d_info, void* d_ws,
                   size_t ws_bytes, cudaStream_t stream);
int compact_launch(const int16_t* d_pcm, i64 n_samples, int sample_rate, const int32_t* d_kept, const i64* d_kept_off,
                   const i64* d_info, int16_t* d_out, i64 out_cap, cudaStream_t stream);
struct LogMelGather { const int32_t* kept_ms; const i64* kept_off; const i64* info; int16_t* trim_out; i64 n_src; };
int logmel_launch(const void* d_audio, int fmt, i64 batch, i64 n, i64 row_stride, const i64* d_n, i64 padding,
                  int n_mels, int norm_mode, float* d_out, i64* d_frames_out, void* d_ws, size_t ws_bytes,
                  cudaStream_t stream, const LogMelGather* gather);
size_t logmel_workspace_bytes(i64 batch, i64 n, i64 padding);
size_t silence_workspace_bytes(i64 n_samples, int sample_rate);

struct PipelineWs {
    size_t off_pcm, off_energy, off_sil, off_keptoff, off_logmel, total;
};

static PipelineWs pipeline_layout(i64 n_in, int in_rate, i64 padding, int cap) {
    PipelineWs w;
    i64 n16 = b2a_resample_out_len(n_in, in_rate, kSampleRate);
    size_t o = 0;
    w.off_pcm = o;      o += align_up((size_t)(n16 + 32) * 2, 256);
    w.off_energy = o;   o += align_up((size_t)(n16 / 16 + 2) * 8, 256);
    w.off_sil = o;      o += align_up(silence_workspace_bytes(n16, kSampleRate), 256);
    w.off_keptoff = o;  o += align_up((size_t)(cap + 2) * 8, 256);
    w.off_logmel = o;   o += align_up(logmel_workspace_bytes(1, n16 + 16, padding), 256);
    w.total = o;
    return w;
}

__global__ void pipeline_notrim_info_kernel(i64* info, i64 n16, int32_t* nonsilent, int32_t* kept, int cap) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        i64 len_ms = n16 / 16;
        for (int i = 0; i < B2A_INFO_LEN; i++) info[i] = 0;
        info[B2A_INFO_LEN_MS] = len_ms;
        info[B2A_INFO_N_KEEP] = n16;
        if (cap > 0) {
            info[B2A_INFO_N_NONSILENT] = 1;
            info[B2A_INFO_N_KEPT] = 1;
            if (nonsilent) { nonsilent[0] = 0; nonsilent[1] = (int32_t)len_ms; }
            if (kept) { kept[0] = 0; kept[1] = (int32_t)len_ms; }
        }
    }
}

int pipeline_launch(const void* d_in, int fmt, int channels, int in_rate, i64 n_in, const b2a_silence_params* prm,
                    int n_mels, i64 padding, int cap, int16_t* d_pcm_out, float* d_mel_out, int32_t* d_nonsilent,
                    int32_t* d_kept, i64* d_info, void* d_ws, size_t ws_bytes, cudaStream_t stream) {
    if (!d_in || !d_pcm_out || !d_mel_out || !d_info || !d_ws) { set_error("pipeline: null pointer"); return B2A_EINVAL; }
    if (cap <= 0 || padding < 0) { set_error("pipeline: bad cap/padding"); return B2A_EINVAL; }
    if (prm && (!d_nonsilent || !d_kept)) { set_error("pipeline: range tables required when trimming"); return B2A_EINVAL; }
    PipelineWs w = pipeline_layout(n_in, in_rate, padding, cap);
    if (ws_bytes < w.total) { set_error("pipeline: workspace too small (%zu < %zu)", ws_bytes, w.total); return B2A_EWORKSPACE; }
    if (((uintptr_t)d_ws) & 255) { set_error("pipeline: workspace must be 256-byte aligned"); return B2A_EINVAL; }
    const i64 n16 = b2a_resample_out_len(n_in, in_rate, kSampleRate);
    char* ws = (char*)d_ws;
    void* ws_logmel = ws + w.off_logmel;
    const size_t ws_logmel_bytes = w.total - w.off_logmel;
    int rc;
    if (!prm) {
        rc = resample_launch(d_in, fmt, channels, in_rate, n_in, kSampleRate, d_pcm_out, nullptr, nullptr, stream);
        if (rc) return rc;
        auto k = pipeline_notrim_info_kernel;
        B2A_LAUNCH(k, 1, 32, 0, stream, d_info, n16, d_nonsilent, d_kept, cap);
        B2A_CHECK_LAUNCH("pipeline_notrim_info_kernel");
        return logmel_launch(d_pcm_out, B2A_FMT_S16, 1, n16, n16, d_info + B2A_INFO_N_KEEP, padding, n_mels, B2A_NORM_WHISPER,
                             d_mel_out, d_info + B2A_INFO_N_FRAMES, ws_logmel, ws_logmel_bytes, stream, nullptr);
    }
    int16_t* pcm16 = (int16_t*)(ws + w.off_pcm);
    u64* energy = (u64*)(ws + w.off_energy);
    i64* kept_off = (i64*)(ws + w.off_keptoff);
    rc = resample_launch(d_in, fmt, channels, in_rate, n_in, kSampleRate, pcm16, nullptr, energy, stream);
    if (rc) return rc;
    rc = silence_launch(energy, n16, kSampleRate, prm, cap, nullptr, d_nonsilent, d_kept, kept_off, d_info, ws + w.off_sil,
                        w.off_keptoff - w.off_sil, stream);
    if (rc) return rc;
    // stream compaction is fused into the log-mel tile loader: it gathers the kept ranges from the untrimmed PCM and
    // writes the trimmed PCM as it goes.  The compacted length lives in d_info[N_KEEP]; n16 + 16 is its upper bound
    // (pydub may zero-fill < 1 ms)
    LogMelGather g;
    g.kept_ms = d_kept; g.kept_off = kept_off; g.info = d_info; g.trim_out = d_pcm_out; g.n_src = n16;
    return logmel_launch(pcm16, B2A_FMT_S16, 1, n16 + 16, n16 + 16, d_info + B2A_INFO_N_KEEP, padding, n_mels, B2A_NORM_WHISPER,
                         d_mel_out, d_info + B2A_INFO_N_FRAMES, ws_logmel, ws_logmel_bytes, stream, &g);
}

// ---- many clips per call: fork / join over a small pool of internal streams --------------------------------------
#ifndef B2A_BATCH_LANES
#define B2A_BATCH_LANES 8
#endif
constexpr int kBatchLanes = B2A_BATCH_LANES;   // clip i runs on lane i % lanes; lane 0 is the caller's stream
struct BatchPool {
    cudaStream_t side[kBatchLanes - 1];
    cudaEvent_t fork, join[kBatchLanes - 1];
    bool ready;
};
static std::mutex g_batch_mu;
static BatchPool g_batch_pool[64];        // per device

static BatchPool* batch_pool() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    BatchPool* bp = &g_batch_pool[dev];
    if (!bp->ready) {
        for (int i = 0; i < kBatchLanes - 1; i++) {
            if (cudaStreamCreateWithFlags(&bp->side[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&bp->join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&bp->fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        bp->ready = true;
    }
    return bp;
}

int pipeline_batch_launch(const b2a_clip_desc* clips, int n_clips, const b2a_silence_params* prm, int n_mels, i64 padding, int cap,
                          cudaStream_t stream) {
    if (n_clips < 0 || (n_clips > 0 && !clips)) { set_error("pipeline_batch: bad clip list"); return B2A_EINVAL; }
    auto one = [&](const b2a_clip_desc& c, cudaStream_t s) {
        return pipeline_launch(c.d_in, c.fmt, c.channels, c.in_rate, c.n_in, prm, n_mels, padding, cap, c.d_pcm_out, c.d_mel_out,
                               c.d_nonsilent_ms, c.d_kept_ms, (i64*)c.d_info, c.d_ws, c.ws_bytes, s);
    };
    if (n_clips <= 1) return n_clips == 1 ? one(clips[0], stream) : B2A_OK;
    // the pool (and its fork / join events) is shared by every caller on this device: enqueue under the lock
    std::lock_guard<std::mutex> lk(g_batch_mu);
    BatchPool* bp = batch_pool();
    if (!bp) { set_error("pipeline_batch: cannot create the internal streams"); return B2A_ECUDA; }
    const int lanes = n_clips < kBatchLanes ? n_clips : kBatchLanes;
    cudaError_t e = cudaEventRecord(bp->fork, stream);
    for (int l = 1; l < lanes && e == cudaSuccess; l++) e = cudaStreamWaitEvent(bp->side[l - 1], bp->fork, 0);
    if (e != cudaSuccess) return cuda_fail(e, "pipeline_batch: fork");
    int rc = B2A_OK;
    for (int i = 0; i < n_clips && rc == B2A_OK; i++) {
        const int l = i % lanes;
        rc = one(clips[i], l == 0 ? stream : bp->side[l - 1]);
    }
    // always join, also after an error: the caller's stream must not run ahead of work already enqueued on the side streams
    for (int l = 1; l < lanes; l++) {
        e = cudaEventRecord(bp->join[l - 1], bp->side[l - 1]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, bp->join[l - 1], 0);
        if (e != cudaSuccess && rc == B2A_OK) rc = cuda_fail(e, "pipeline_batch: join");
    }
    return rc;
}

}  // namespace b2a

extern "C" {

int b2a_pipeline_batch(const b2a_clip_desc* clips, int n_clips, const b2a_silence_params* params, int n_mels, int64_t padding,
                       int32_t cap, b2a_stream_t stream) {
    return b2a::pipeline_batch_launch(clips, n_clips, params, n_mels, padding, cap, (cudaStream_t)stream);
}

size_t b2a_pipeline_workspace_bytes(int64_t n_in, int in_rate, int64_t padding, int32_t cap) {
    if (n_in <= 0 || in_rate <= 0 || padding < 0 || cap <= 0) return 0;
    return b2a::pipeline_layout(n_in, in_rate, padding, cap).total;
}

int b2a_pipeline(const void* d_in, int fmt, int channels, int in_rate, int64_t n_in, const b2a_silence_params* params,
                 int n_mels, int64_t padding, int32_t cap, int16_t* d_pcm_out, float* d_mel_out, int32_t* d_nonsilent_ms,
                 int32_t* d_kept_ms, int64_t* d_info, void* d_ws, size_t ws_bytes, b2a_stream_t stream) {
    return b2a::pipeline_launch(d_in, fmt, channels, in_rate, n_in, params, n_mels, padding, cap, d_pcm_out, d_mel_out,
                                d_nonsilent_ms, d_kept_ms, (b2a::i64*)d_info, d_ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
