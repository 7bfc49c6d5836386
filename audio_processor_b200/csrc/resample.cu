// resample.cu — conversion to 16 kHz mono s16: dispatch, generic/edge kernel, same-rate passthrough.
//
// Replaces the libswresample work behind convert_to_wav (app/services/audio_processor.py:901-930,
// ffmpeg command :912-920).  The tensor-core fast path (fir_mma.cuh) covers 44.1 kHz and 48 kHz s16 input; this
// file holds the table-driven kernel used for (a) the head/tail outputs whose windows need
// libswresample's reflect / symmetric edge extension, (b) every other rate pair and float input,
// and the same-rate paths (identity, (L+R+1)>>1 stereo s16 downmix, float quantisation).
#include "b2a_tables.cuh"
#include "fir_mma.cuh"
#include "resample_generic.cuh"

namespace b2a {

// tensor-core fast path (fir_mma.cuh), instantiated in its own translation units
int fir_fast_dispatch(int in_rate, int fmt, int channels, const void* d_in, i64 n_in, int16_t* d_out_s16, float* d_out_f32,
                      u64* d_energy, FirMmaPlan* plan, const GenericParams* edge, cudaStream_t stream);

__global__ void __launch_bounds__(256) resample_generic_kernel(GenericParams p) {
    resample_generic_output(p, (i64)blockIdx.x * blockDim.x + threadIdx.x);
}

// same-rate paths: libswresample stays in the sample domain
__global__ void __launch_bounds__(256) passthrough_kernel(const void* __restrict__ in, int fmt, int channels, i64 n,
                                                          int16_t* __restrict__ out_s16, float* __restrict__ out_f32,
                                                          u64* __restrict__ energy, int spm, i64 n_energy) {
    const i64 m = (i64)blockIdx.x * blockDim.x + threadIdx.x;   // blockDim multiple of 32, m0 multiple of 32
    const bool valid = m < n;
    int q = 0;
    if (valid) {
        float y;
        if (fmt == B2A_FMT_S16) {
            const int16_t* s = (const int16_t*)in;
            if (channels == 1) q = s[m];
            else q = ((int)s[2 * m] + (int)s[2 * m + 1] + 1) >> 1;          // swr s16 rematrix, rounds half up
            y = (float)q * (1.0f / 32768.0f);
            if (channels == 2 && out_f32) y = ((float)s[2 * m] + (float)s[2 * m + 1]) * (1.0f / 65536.0f);
        } else {
            const float* s = (const float*)in;
            y = channels == 1 ? s[m] : 0.5f * s[2 * m] + 0.5f * s[2 * m + 1];
            q = quant_s16(y * 32768.0f);
        }
        if (out_s16) out_s16[m] = (int16_t)q;
        if (out_f32) out_f32[m] = y;
    }
    if (energy && spm > 0) {
        u64 sq = valid ? (u64)(unsigned)(q * q) : 0ull;
        if (spm <= 32 && (32 % spm) == 0) {
            for (int o = 1; o < spm; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            if (((threadIdx.x & 31) % spm) == 0 && (m / spm) < n_energy) energy[m / spm] = sq;
        } else if (valid) {
            atomicAdd((unsigned long long*)&energy[m / spm], (unsigned long long)sq);
        }
    }
}

int resample_launch(const void* d_in, int fmt, int channels, int in_rate, i64 n_in, int out_rate, int16_t* d_out_s16,
                    float* d_out_f32, u64* d_energy, cudaStream_t stream) {
    if (!d_in) { set_error("resample: null input"); return B2A_EINVAL; }
    if (fmt != B2A_FMT_S16 && fmt != B2A_FMT_F32) { set_error("resample: bad fmt %d", fmt); return B2A_EINVAL; }
    if (channels != 1 && channels != 2) { set_error("resample: %d channels unsupported (mono/stereo only)", channels); return B2A_EUNSUPPORTED; }
    if (in_rate <= 0 || out_rate <= 0 || n_in <= 0) { set_error("resample: bad rate/length"); return B2A_EINVAL; }
    if (!d_out_s16 && !d_out_f32 && !d_energy) { set_error("resample: no output requested"); return B2A_EINVAL; }
    int spm = 0;
    i64 n_out = b2a_resample_out_len(n_in, in_rate, out_rate);
    i64 n_energy = 0;
    if (d_energy) {
        if (out_rate % 1000) { set_error("resample: per-ms energy needs out_rate %% 1000 == 0"); return B2A_EUNSUPPORTED; }
        spm = out_rate / 1000;
        n_energy = (n_out + spm - 1) / spm;
    }

    if (in_rate == out_rate) {
        const bool direct = spm > 0 && spm <= 32 && (32 % spm) == 0;
        if (d_energy && !direct) cudaMemsetAsync(d_energy, 0, (size_t)n_energy * 8, stream);
        auto k = passthrough_kernel;
        i64 cover = d_energy ? n_energy * spm : n_in;            // include the zero-extended last millisecond
        if (cover < n_in) cover = n_in;
        B2A_LAUNCH(k, (unsigned)((cover + 255) / 256), 256, 0, stream, d_in, fmt, channels, n_in, d_out_s16, d_out_f32, d_energy, spm, n_energy);
        B2A_CHECK_LAUNCH("passthrough_kernel");
        return B2A_OK;
    }

    const ResampleDesign* des = get_resample_design(in_rate, out_rate);
    if (!des) return B2A_EINVAL;
    if (n_out == 0) return B2A_OK;     // too short to fill the filter: libswresample returns no samples either (b2a_resample_out_len)

    GenericParams p;
    p.in = d_in; p.fmt = fmt; p.channels = channels; p.n_in = n_in; p.n_out = n_out;
    p.L = des->L; p.M = des->M; p.taps = des->taps; p.center = des->center; p.taps_dev = des->d_taps;
    p.out_s16 = d_out_s16; p.out_f32 = d_out_f32; p.energy = d_energy; p.spm = spm; p.n_energy = n_energy;
    p.energy_atomic = 0;
    p.lo0 = 0; p.hi0 = n_out; p.lo1 = p.hi1 = n_out;

    // fast path for the named rate pairs (needs 16-byte aligned buffers).  The tcgen05 kernel also computes the edge
    // outputs (in its otherwise idle warps) and then reports the whole clip as done.
    FirMmaPlan plan;
    plan.out_lo = plan.out_hi = 0;
    const bool aligned = ((((uintptr_t)d_in) | ((uintptr_t)d_out_s16) | ((uintptr_t)d_out_f32) | ((uintptr_t)d_energy)) & 15) == 0;
    if (aligned && out_rate == 16000 && n_in > des->taps) {    // shorter clips: every window needs the edge extension
        int rc = fir_fast_dispatch(in_rate, fmt, channels, d_in, n_in, d_out_s16, d_out_f32, d_energy, &plan, &p, stream);
        if (rc < 0) return rc;
    }
    if (plan.out_hi > plan.out_lo) { p.lo0 = 0; p.hi0 = plan.out_lo; p.lo1 = plan.out_hi; p.hi1 = n_out; }
    const bool direct = spm > 0 && spm <= 32 && (32 % spm) == 0;
    if (d_energy && !direct) {
        // (the fast path zeroes the table itself when it accumulates atomically)
        cudaMemsetAsync(d_energy, 0, (size_t)n_energy * 8, stream);     // only for output rates whose millisecond is not a power-of-two count (no fast path there)
        p.energy_atomic = 1;
    }
    const i64 total = resample_generic_total(p);
    if (total > 0) {
        auto k = resample_generic_kernel;
        B2A_LAUNCH(k, (unsigned)((total + 255) / 256), 256, 0, stream, p);
        B2A_CHECK_LAUNCH("resample_generic_kernel");
    }
    return B2A_OK;
}

}  // namespace b2a

extern "C" int b2a_resample(const void* d_in, int fmt, int channels, int in_rate, int64_t n_in, int out_rate, int16_t* d_out_s16,
                            float* d_out_f32, uint64_t* d_energy_ms, b2a_stream_t stream) {
    return b2a::resample_launch(d_in, fmt, channels, in_rate, n_in, out_rate, d_out_s16, d_out_f32, (b2a::u64*)d_energy_ms,
                                (cudaStream_t)stream);
}
