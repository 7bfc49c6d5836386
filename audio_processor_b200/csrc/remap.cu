// remap.cu — timestamps on the silence-stripped timeline back to the original recording, on the device.
//
// Consumer in the reference: the speaker-overlap loop of process_audio (app/services/audio_processor.py:1114-1145) compares
// Whisper's segment start / end (seconds on the audio Whisper saw = the TRIMMED clip once preprocess_audio strips silence,
// :1046-1051) with diarization times on the original clip.  The kept-range table this library already produces
// (d_kept_ms [n][2] in ms, d_kept_off [n + 1] in samples of the trimmed clip) is the map between the two:
//   trimmed time t (s) -> t_ms = 1000 t -> range i = first with end offset (kept_off[i+1] / spm ms) >= t_ms
//                      -> kept_ms[i][0] + (t_ms - kept_off[i] / spm)  ms;   beyond the last range: its end.
// (a time exactly on a cut maps to the END of the earlier range, as service.remap_segments does on the host.)
// One thread per timestamp, binary search in the offset table (L2-resident, a few KB).
#include "b2a_common.cuh"

namespace b2a {

__global__ void __launch_bounds__(256) remap_times_kernel(const double* __restrict__ t_in, i64 n, const int32_t* __restrict__ kept_ms,
                                                          const i64* __restrict__ kept_off, const i64* __restrict__ info, int spm,
                                                          double* __restrict__ t_out) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int n_kept = (int)info[B2A_INFO_N_KEPT];
    const double t = t_in[i];
    if (n_kept <= 0) { t_out[i] = t; return; }
    const double t_ms = t * 1000.0;
    // first range whose trimmed-timeline end is >= t_ms
    int lo = 0, hi = n_kept;                               // answer in [lo, hi]; hi = n_kept means "behind the last range"
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const double end_ms = (double)kept_off[mid + 1] / (double)spm;
        if (end_ms >= t_ms) hi = mid; else lo = mid + 1;
    }
    if (lo >= n_kept) { t_out[i] = (double)kept_ms[2 * (n_kept - 1) + 1] / 1000.0; return; }
    const double start_ms = (double)kept_off[lo] / (double)spm;
    t_out[i] = ((double)kept_ms[2 * lo] + (t_ms - start_ms)) / 1000.0;
}

// kept_off[k] = samples of the trimmed clip before kept range k (exclusive scan of the range lengths), kept_off[n] = total:
// for callers that hold only the range table (b2a_pipeline keeps its own copy in the workspace).  One block.
__global__ void __launch_bounds__(1024) kept_offsets_kernel(const int32_t* __restrict__ kept_ms, const i64* __restrict__ info, int spm,
                                                            i64* __restrict__ kept_off, int cap) {
    __shared__ i64 s_warp[32];
    __shared__ i64 s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int n = (int)info[B2A_INFO_N_KEPT];
    if (n > cap) n = cap;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int k = base + tid;
        const i64 len = k < n ? (i64)(kept_ms[2 * k + 1] - kept_ms[2 * k]) * spm : 0;
        i64 inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const i64 y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        i64 wpre = 0, tot = 0;
        for (int w = 0; w < 32; w++) { const i64 x = s_warp[w]; if (w < warp) wpre += x; tot += x; }
        const i64 carry = s_carry;
        if (k < n) kept_off[k] = carry + wpre + inc - len;
        __syncthreads();
        if (tid == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (tid == 0) kept_off[n] = s_carry;
}

int remap_launch(const double* d_t_in, i64 n, const int32_t* d_kept_ms, const i64* d_kept_off, const i64* d_info, int sample_rate,
                 double* d_t_out, cudaStream_t stream) {
    if (n < 0 || (n > 0 && (!d_t_in || !d_t_out)) || !d_kept_ms || !d_kept_off || !d_info) { set_error("remap: bad argument"); return B2A_EINVAL; }
    if (sample_rate <= 0 || sample_rate % 1000) { set_error("remap: bad sample rate"); return B2A_EUNSUPPORTED; }
    if (n == 0) return B2A_OK;
    auto k = remap_times_kernel;
    B2A_LAUNCH(k, (unsigned)((n + 255) / 256), 256, 0, stream, d_t_in, n, d_kept_ms, d_kept_off, d_info, sample_rate / 1000, d_t_out);
    B2A_CHECK_LAUNCH("remap_times_kernel");
    return B2A_OK;
}

}  // namespace b2a

extern "C" int b2a_kept_offsets(const int32_t* d_kept_ms, const int64_t* d_info, int sample_rate, int32_t cap, int64_t* d_kept_off,
                                b2a_stream_t stream) {
    if (!d_kept_ms || !d_info || !d_kept_off || cap <= 0 || sample_rate <= 0 || sample_rate % 1000) { b2a::set_error("kept_offsets: bad argument"); return B2A_EINVAL; }
    auto k = b2a::kept_offsets_kernel;
    B2A_LAUNCH(k, 1, 1024, 0, (cudaStream_t)stream, d_kept_ms, (const b2a::i64*)d_info, sample_rate / 1000, (b2a::i64*)d_kept_off, (int)cap);
    B2A_CHECK_LAUNCH("kept_offsets_kernel");
    return B2A_OK;
}

extern "C" int b2a_remap_times(const double* d_t_in, int64_t n, const int32_t* d_kept_ms, const int64_t* d_kept_off, const int64_t* d_info,
                               int sample_rate, double* d_t_out, b2a_stream_t stream) {
    return b2a::remap_launch(d_t_in, n, d_kept_ms, (const b2a::i64*)d_kept_off, (const b2a::i64*)d_info, sample_rate, d_t_out, (cudaStream_t)stream);
}
