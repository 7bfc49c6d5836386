// resample_generic.cuh — the table-driven resampler body (one thread per output): edge outputs whose windows need
// libswresample's reflect / symmetric extension, every rate pair without a tensor-core kernel, float input.
// Shared by resample_generic_kernel (resample.cu) and the otherwise idle warps of the tcgen05 FIR (fir_tmem.cuh), which
// compute the clip's head and tail while the main pipeline streams.
#pragma once
#include "b2a_common.cuh"

namespace b2a {

struct GenericParams {
    const void* in;
    int fmt, channels;
    i64 n_in;
    i64 n_out;
    int L, M, taps, center;
    const float* taps_dev;        // [L][taps]
    int16_t* out_s16;
    float* out_f32;
    u64* energy;                  // nullable
    int spm;                      // samples per ms at the output rate (0: no energy)
    i64 n_energy;
    // output ranges [lo0,hi0) and [lo1,hi1) (multiples of 32 at the low ends); blocks cover them back to back
    i64 lo0, hi0, lo1, hi1;
    int energy_atomic;
};

__device__ __forceinline__ float generic_sample(const GenericParams& p, i64 k) {
    // libswresample edge handling: reflect before the start (edge not repeated),
    // symmetric after the end (edge repeated)
    if (k < 0) k = -k;
    if (k >= p.n_in) k = 2 * p.n_in - 1 - k;
    if (k < 0) k = 0;
    if (k >= p.n_in) k = p.n_in - 1;
    if (p.fmt == B2A_FMT_S16) {
        const int16_t* s = (const int16_t*)p.in;
        if (p.channels == 1) return (float)s[k] * (1.0f / 32768.0f);
        return ((float)s[2 * k] + (float)s[2 * k + 1]) * (1.0f / 65536.0f);   // exact: 0.5*L + 0.5*R
    } else {
        const float* s = (const float*)p.in;
        if (p.channels == 1) return s[k];
        return 0.5f * s[2 * k] + 0.5f * s[2 * k + 1];
    }
}

// one thread per output sample, taps from global memory (L2/L1 resident).  r = linear index over the two output ranges
// (the first padded to a multiple of 32); a converged warp must pass 32 consecutive r starting at a multiple of 32 (the
// per-millisecond energies are reduced with shuffles).
__device__ __forceinline__ void resample_generic_output(const GenericParams& p, i64 r) {
    const i64 span0 = p.hi0 - p.lo0;
    // block ranges are padded to multiples of 32 so a warp never straddles the two ranges
    const i64 span0p = (span0 + 31) / 32 * 32;
    i64 m;
    bool valid;
    if (r < span0p) { m = p.lo0 + r; valid = m < p.hi0; }
    else { m = p.lo1 + (r - span0p); valid = m < p.hi1; }
    int q = 0;
    if (valid) {
        const i64 t = m * p.M;
        const i64 idx = t / p.L;
        const int ph = (int)(t % p.L);
        const float* h = p.taps_dev + (size_t)ph * p.taps;
        float a0 = 0.f, a1 = 0.f;
        const i64 base = idx - p.center;
        int i = 0;
        if (base >= 0 && base + p.taps <= p.n_in) {
            // interior window: no edge extension, plain strided loads, four taps in flight per accumulator pair
            float b0 = 0.f, b1 = 0.f;
            if (p.fmt == B2A_FMT_S16 && p.channels == 2 && (((uintptr_t)p.in) & 3) == 0) {
                const int* s = (const int*)p.in + base;            // one 32-bit word per stereo frame
                // 16 taps per round with all 32 loads issued up front (same accumulator order as the 4-tap loop below): the
                // FIR's spare warps run this while the kernel saturates HBM, where a dependent round trip costs microseconds
                for (; i + 15 < p.taps; i += 16) {
                    int sv[16];
                    float hv[16];
#pragma unroll
                    for (int j = 0; j < 16; j++) { sv[j] = s[i + j]; hv[j] = h[i + j]; }
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        a0 = fmaf((float)__dp2a_lo(sv[j], 0x0101, 0), hv[j], a0);
                        a1 = fmaf((float)__dp2a_lo(sv[j + 1], 0x0101, 0), hv[j + 1], a1);
                        b0 = fmaf((float)__dp2a_lo(sv[j + 2], 0x0101, 0), hv[j + 2], b0);
                        b1 = fmaf((float)__dp2a_lo(sv[j + 3], 0x0101, 0), hv[j + 3], b1);
                    }
                }
                for (; i + 3 < p.taps; i += 4) {
                    a0 = fmaf((float)__dp2a_lo(s[i], 0x0101, 0), h[i], a0);
                    a1 = fmaf((float)__dp2a_lo(s[i + 1], 0x0101, 0), h[i + 1], a1);
                    b0 = fmaf((float)__dp2a_lo(s[i + 2], 0x0101, 0), h[i + 2], b0);
                    b1 = fmaf((float)__dp2a_lo(s[i + 3], 0x0101, 0), h[i + 3], b1);
                }
                for (; i < p.taps; i++) a0 = fmaf((float)__dp2a_lo(s[i], 0x0101, 0), h[i], a0);
                a0 = ((a0 + b0) + (a1 + b1)) * (1.0f / 65536.0f);   // exact power-of-two scale of 0.5*(L+R)/32768
                a1 = 0.f;
            } else {
                for (; i + 3 < p.taps; i += 4) {
                    a0 = fmaf(generic_sample(p, base + i), h[i], a0);
                    a1 = fmaf(generic_sample(p, base + i + 1), h[i + 1], a1);
                    b0 = fmaf(generic_sample(p, base + i + 2), h[i + 2], b0);
                    b1 = fmaf(generic_sample(p, base + i + 3), h[i + 3], b1);
                }
                for (; i < p.taps; i++) a0 = fmaf(generic_sample(p, base + i), h[i], a0);
                a0 += b0; a1 += b1;
            }
        } else {
            for (; i + 1 < p.taps; i += 2) {
                a0 = fmaf(generic_sample(p, base + i), h[i], a0);
                a1 = fmaf(generic_sample(p, base + i + 1), h[i + 1], a1);
            }
            if (i < p.taps) a0 = fmaf(generic_sample(p, base + i), h[i], a0);
        }
        const float y = a0 + a1;
        q = quant_s16(y * 32768.0f);
        if (p.out_s16) p.out_s16[m] = (int16_t)q;
        if (p.out_f32) p.out_f32[m] = y;
    }
    if (p.energy && p.spm > 0) {
        u64 sq = valid ? (u64)(unsigned)(q * q) : 0ull;
        if (!p.energy_atomic) {
            // a warp covers 32 consecutive outputs starting at a multiple of 32 and spm divides 32:
            // reduce per spm-lane group; the group leader is valid iff the millisecond has any output
            // in this range (a trailing partial millisecond is thereby zero-extended)
            for (int o = 1; o < p.spm; o <<= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            if ((m % p.spm) == 0 && valid) p.energy[m / p.spm] = sq;
        } else if (valid) {
            atomicAdd((unsigned long long*)&p.energy[m / p.spm], (unsigned long long)sq);
        }
    }
}

// total linear indices (rounded up to whole warps) covering both ranges of p
static inline i64 resample_generic_total(const GenericParams& p) {
    i64 span0 = (p.hi0 - p.lo0 + 31) / 32 * 32;
    i64 span1 = p.hi1 - p.lo1;
    if (p.spm > 0) span1 = (span1 + p.spm - 1) / p.spm * p.spm;     // cover the zero-extended tail of the last millisecond too
    return span0 + span1;
}

}  // namespace b2a
