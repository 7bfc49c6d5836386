// resample_fast.cuh — fused PCM decode + downmix + polyphase FIR with compile-time taps (sm_100a).
//
// Replaces libswresample's swr_convert inner loop behind the reference's
//   ffmpeg -ar 16000 -ac 1 -c:a pcm_s16le   (app/services/audio_processor.py:912-923).
//
// Mapping ("lane = run"): every thread produces NOUT consecutive output samples whose phase
// pattern is identical for all threads (NOUT is a multiple of the phase count L), so after full
// unrolling every tap is a compile-time constant and nvcc folds it into the immediate operand
// of an FFMA: the inner loop is FFMA-imm only (no tap loads, no tap registers).  A warp owns 32
// consecutive runs; their raw input frames (one contiguous range of the clip) are pulled into
// shared memory with 16-byte cp.async, and each lane walks its own row with a per-lane stride of
// S = NOUT*M/L frames, which is odd for both named rate pairs (441 and 51) => conflict-free LDS.
// The lane keeps a sliding register window of TAPS(+3) decoded samples; every input sample is
// decoded (s16 pair -> dp2a -> float) exactly once per lane.
//
// Epilogue per output: scale, lrintf+clip to s16 (swr audioconvert), optional float copy, and
// the per-millisecond sum of squares of the QUANTISED samples (uint64) that the silence detector
// consumes — so trimming never re-reads the PCM to find energies.
#pragma once
#include "b2a_common.cuh"
#ifndef B2A_FIR_TAPS_INCLUDED
#define B2A_FIR_TAPS_INCLUDED
#include "fir_taps_gen.inc"
#endif

namespace b2a {

template <int IN_RATE> struct FirTraits;
template <> struct FirTraits<44100> {
    static constexpr int L = 160, M = 441, TAPS = 92, NOUT = 160;
    __device__ static __forceinline__ float tap(int idx) { return kFirTaps_44100_16000[idx]; }
};
template <> struct FirTraits<48000> {
    static constexpr int L = 1, M = 3, TAPS = 100, NOUT = 17;
    __device__ static __forceinline__ float tap(int idx) { return kFirTaps_48000_16000[idx]; }
};

// geometry shared by host and device
template <int IN_RATE, int FB /*bytes per input frame*/>
struct FirGeom {
    using TR = FirTraits<IN_RATE>;
    static constexpr int L = TR::L, M = TR::M, TAPS = TR::TAPS, NOUT = TR::NOUT;
    static constexpr int CENTER = (TAPS - 1) / 2;
    static constexpr int S = NOUT * M / L;                  // input frames per run (441 / 51)
    static constexpr int SET_IN = 32 * S;                   // input frames per warp-set
    static constexpr int SET_OUT = 32 * NOUT;               // outputs per warp-set
    static constexpr int LAST_B = ((NOUT - 1) * M) / L;     // window start of the last output of a run
    static constexpr int SPAN = 31 * S + LAST_B + TAPS;     // frames a set touches, from (set start - CENTER)
    // byte offset of (set start - CENTER) modulo 16 is the same for every set (32*S*FB % 16 == 0)
    static constexpr int ALIGN_BYTES = ((16 - (CENTER * FB) % 16) % 16);   // bytes we start early
    static constexpr int ALIGN_FRAMES = ALIGN_BYTES / FB;
    static constexpr int TILE_BYTES = ((ALIGN_BYTES + SPAN * FB + 15) / 16) * 16;
    static constexpr int RING = ((TAPS + 3 + 3) / 4) * 4;
    static_assert((32 * S * FB) % 16 == 0, "set stride must keep 16-byte alignment");
    static_assert(ALIGN_BYTES % FB == 0, "alignment slack must be whole frames");
    static_assert(NOUT % L == 0, "a run must cover whole phase periods");
};

#ifndef B2A_EMU
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
#else
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 16); }
__device__ __forceinline__ void cp_async_wait_all() {}
#endif

// decode one input frame (already in shared memory) to the FIR's working float.
//  s16: integer sample (mono) or L+R (stereo) as float; the 2^-15 / 2^-16 scale is applied once per output.
//  f32: sample (mono) or 0.5*L + 0.5*R (stereo), exactly as swr's float rematrix.
template <int FMT, int CH>
__device__ __forceinline__ float fir_decode(const unsigned char* row, int k) {
    if (FMT == B2A_FMT_S16 && CH == 2) {
        int v = *(const int*)(row + 4 * k);
        return (float)__dp2a_lo(v, 0x0101, 0);
    } else if (FMT == B2A_FMT_S16 && CH == 1) {
        return (float)(*(const short*)(row + 2 * k));
    } else if (FMT == B2A_FMT_F32 && CH == 1) {
        return *(const float*)(row + 4 * k);
    } else {
        float2 v = *(const float2*)(row + 8 * k);
        return 0.5f * v.x + 0.5f * v.y;
    }
}
template <int FMT, int CH> __device__ __forceinline__ float fir_out_scale() {
    return FMT == B2A_FMT_S16 ? (CH == 2 ? (1.0f / 65536.0f) : (1.0f / 32768.0f)) : 1.0f;
}

template <int IN_RATE, int FMT, int CH>
struct FirRun {
    static constexpr int FB = (FMT == B2A_FMT_S16 ? 2 : 4) * CH;
    using G = FirGeom<IN_RATE, FB>;
    using TR = FirTraits<IN_RATE>;

    struct State {
        float w[G::RING];          // sliding window of decoded input (register ring after unrolling)
        unsigned pack[4];          // 8 quantised outputs awaiting a 16-byte store
        float fpack[4];
        u64 e_acc;                 // running sum of squares for the current millisecond
    };

    template <int J>
    __device__ static __forceinline__ void output(State& st, const unsigned char* row, i64 m0, int16_t* out_s16,
                                                  float* out_f32, u64* energy) {
        constexpr int B = (J * G::M) / G::L;                          // window start (frames from row origin)
        constexpr int PH = (J * G::M) % G::L;
        constexpr int PREV_END = (J == 0) ? 0 : (((J - 1) * G::M) / G::L + G::TAPS);
#pragma unroll
        for (int k = PREV_END; k < B + G::TAPS; k++) st.w[k % G::RING] = fir_decode<FMT, CH>(row, k);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int i = 0; i < G::TAPS; i += 4) {
            a0 = fmaf(st.w[(B + i) % G::RING], TR::tap(PH * G::TAPS + i), a0);
            a1 = fmaf(st.w[(B + i + 1) % G::RING], TR::tap(PH * G::TAPS + i + 1), a1);
            a2 = fmaf(st.w[(B + i + 2) % G::RING], TR::tap(PH * G::TAPS + i + 2), a2);
            a3 = fmaf(st.w[(B + i + 3) % G::RING], TR::tap(PH * G::TAPS + i + 3), a3);
        }
        const float y = ((a0 + a1) + (a2 + a3)) * fir_out_scale<FMT, CH>();
        const int q = quant_s16(y * 32768.0f);
        st.e_acc += (u64)(unsigned)(q * q);

        if (G::NOUT % 16 == 0) {
            // run is a whole number of milliseconds and of 16-byte output vectors
            if (J % 2 == 0) st.pack[(J % 8) / 2] = (unsigned)(q & 0xffff);
            else st.pack[(J % 8) / 2] |= ((unsigned)q << 16);
            if (J % 8 == 7 && out_s16) *(uint4*)(out_s16 + m0 + J - 7) = make_uint4(st.pack[0], st.pack[1], st.pack[2], st.pack[3]);
            if (out_f32) {
                st.fpack[J % 4] = y;
                if (J % 4 == 3) *(float4*)(out_f32 + m0 + J - 3) = make_float4(st.fpack[0], st.fpack[1], st.fpack[2], st.fpack[3]);
            }
            if (J % 16 == 15) {
                if (energy) energy[(m0 + J) / 16] = st.e_acc;
                st.e_acc = 0;
            }
        } else {
            if (out_s16) out_s16[m0 + J] = (int16_t)q;
            if (out_f32) out_f32[m0 + J] = y;
            // milliseconds straddle runs: integer atomics (exact, order-independent) into a zeroed table
            const bool ms_end = ((m0 + J) % 16 == 15) || (J == G::NOUT - 1);
            if (ms_end) {
                if (energy) atomicAdd((unsigned long long*)&energy[(m0 + J) / 16], (unsigned long long)st.e_acc);
                st.e_acc = 0;
            }
        }
    }

    template <int J0, int J1>
    __device__ static __forceinline__ void outputs(State& st, const unsigned char* row, i64 m0, int16_t* out_s16, float* out_f32,
                                                   u64* energy) {
        if constexpr (J1 - J0 == 1) {
            output<J0>(st, row, m0, out_s16, out_f32, energy);
        } else {
            outputs<J0, (J0 + J1) / 2>(st, row, m0, out_s16, out_f32, energy);
            outputs<(J0 + J1) / 2, J1>(st, row, m0, out_s16, out_f32, energy);
        }
    }
};

// one warp per block; block b handles warp-set (set0 + b): runs [32*set, 32*set+32)
template <int IN_RATE, int FMT, int CH>
__global__ void __launch_bounds__(32) fir_fast_kernel(const unsigned char* __restrict__ in, i64 set0, int16_t* __restrict__ out_s16,
                                                      float* __restrict__ out_f32, u64* __restrict__ energy) {
    using R = FirRun<IN_RATE, FMT, CH>;
    using G = typename R::G;
    B2A_DYN_SMEM(tile);
    const int lane = threadIdx.x;
    const i64 set = set0 + blockIdx.x;
    // first frame the set needs is (set*SET_IN - CENTER); start ALIGN_BYTES earlier => 16-byte aligned source
    const unsigned char* src = in + ((i64)set * G::SET_IN - G::CENTER) * R::FB - G::ALIGN_BYTES;
    for (int c = lane; c < G::TILE_BYTES / 16; c += 32) cp_async16(tile + 16 * c, src + 16 * (i64)c);
    cp_async_wait_all();
    __syncwarp();
    const unsigned char* row = tile + G::ALIGN_BYTES + (size_t)lane * G::S * R::FB;
    const i64 m0 = ((i64)set * 32 + lane) * G::NOUT;
    typename R::State st;
    st.e_acc = 0;
    st.pack[0] = st.pack[1] = st.pack[2] = st.pack[3] = 0;
    R::template outputs<0, G::NOUT>(st, row, m0, out_s16, out_f32, energy);
}

// host-side description of one instantiation
struct FirFastPlan {
    i64 set_first, set_count;     // warp-sets handled by the fast kernel
    i64 out_lo, out_hi;           // outputs [out_lo, out_hi) are produced by it
    bool energy_atomic;           // energy table must be zeroed first (ms straddle runs)
};

template <int IN_RATE, int FMT, int CH>
static inline FirFastPlan fir_fast_plan(i64 n_in, i64 n_out) {
    using R = FirRun<IN_RATE, FMT, CH>;
    using G = typename R::G;
    FirFastPlan p;
    p.energy_atomic = (G::NOUT % 16 != 0);
    // interior sets: s >= 1 (no reflect) and the 16-byte-rounded tile ends inside the input
    const i64 total_bytes = n_in * R::FB;
    i64 s_hi = 0;   // exclusive
    {
        // tile(s) = [ (s*SET_IN - CENTER)*FB - ALIGN_BYTES , +TILE_BYTES )
        i64 num = total_bytes + G::ALIGN_BYTES - G::TILE_BYTES + (i64)G::CENTER * R::FB;
        if (num >= 0) s_hi = num / ((i64)G::SET_IN * R::FB) + 1;
    }
    i64 s_out = n_out / G::SET_OUT;      // sets whose outputs all exist
    if (s_hi > s_out) s_hi = s_out;
    p.set_first = 1;
    p.set_count = s_hi > 1 ? s_hi - 1 : 0;
    p.out_lo = p.set_first * G::SET_OUT;
    p.out_hi = (p.set_first + p.set_count) * G::SET_OUT;
    if (p.set_count == 0) { p.out_lo = p.out_hi = 0; }
    return p;
}

template <int IN_RATE, int FMT, int CH>
static inline int fir_fast_launch(const void* d_in, const FirFastPlan& p, int16_t* d_out_s16, float* d_out_f32, u64* d_energy,
                                  cudaStream_t stream) {
    using R = FirRun<IN_RATE, FMT, CH>;
    using G = typename R::G;
    if (p.set_count <= 0) return B2A_OK;
    auto k = fir_fast_kernel<IN_RATE, FMT, CH>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::TILE_BYTES);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fir_fast_kernel)");
    B2A_LAUNCH(k, (unsigned)p.set_count, 32, G::TILE_BYTES, stream, (const unsigned char*)d_in, p.set_first, d_out_s16, d_out_f32, d_energy);
    B2A_CHECK_LAUNCH("fir_fast_kernel");
    return B2A_OK;
}

}  // namespace b2a

// Each instantiation lives in its own translation unit (fir_fast_*.cu) so the long unrolled kernels
// compile in parallel.  Returns 1 when the fast kernel was launched, 0 when nothing qualified, <0 on error.
#define B2A_DEFINE_FIR_FAST(NAME, RATE, FMT, CH)                                                                         \
    namespace b2a {                                                                                                      \
    int NAME(const void* d_in, i64 n_in, i64 n_out, int16_t* d_out_s16, float* d_out_f32, u64* d_energy,                 \
             FirFastPlan* plan, cudaStream_t stream) {                                                                   \
        *plan = fir_fast_plan<RATE, FMT, CH>(n_in, n_out);                                                               \
        if (plan->set_count <= 0) return 0;                                                                              \
        if (d_energy && plan->energy_atomic) {                                                                           \
            cudaError_t e = cudaMemsetAsync(d_energy, 0, (size_t)((n_out + 15) / 16) * 8, stream);                       \
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(energy)");                                        \
        }                                                                                                                \
        int rc = fir_fast_launch<RATE, FMT, CH>(d_in, *plan, d_out_s16, d_out_f32, d_energy, stream);                    \
        return rc < 0 ? rc : 1;                                                                                          \
    }                                                                                                                    \
    }
