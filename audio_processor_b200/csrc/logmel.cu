// logmel.cu — Whisper log-mel spectrogram on sm_100a.
//
// Replaces whisper.audio.log_mel_spectrogram (openai-whisper whisper/audio.py), which the
// reference reaches through model.transcribe at app/services/audio_processor.py:1076-1080:
//   reflect-pad 200 | 400-sample frames at hop 160 | periodic Hann | rFFT-400 | |.|^2 |
//   drop last frame | slaney mel [n_mels x 201] | log10(max(.,1e-10)) | max(., gmax-8) | (x+4)/4
//
// Kernel design (K4 "stft_mel", K5 "mel_floor"):
//  * persistent CTAs of 160 threads = 5 warps; each iteration handles a tile of 32 frames of one clip.
//    lane = frame, warp = role: every per-role quantity (twiddles, window slice, mel group) is warp-uniform,
//    so table reads are shared-memory broadcasts and per-lane addresses are "base + compile-time offset".
//  * the tile's 5360 samples are staged once in shared memory as f32 (reflect / zero-pad resolved at load,
//    16-byte vector loads on the interior path), skewed by 2 words per hop so that the 32 lanes of a warp,
//    whose frames start 160 samples apart, hit distinct banks.
//  * one real 400-point FFT per frame = one complex 200-point FFT of (even, odd) samples, split over the 5 roles:
//    role u does the radix-10 butterflies (2x5 prime-factor) of columns n2 = 4u .. 4u + 3, applies the stage twiddles,
//    transposes through shared memory (pitch 202 complex per frame; two neighbouring columns per 16-byte access), then the
//    radix-20 butterflies (4x5 prime-factor) of output residues {u, 10-u} mod 10 — so the real-FFT unpack pairs (k, 200-k)
//    stay inside one thread.  The unpack produces 4|X|^2 (16 flops per conjugate pair); the 1/4 lives in the mel weights.
//  * mel projection: warp u owns mels u, u + 5, u + 10, ... (lane = frame), so all five warps run the same fully unrolled
//    code over the sparse filterbank (mel_tables_gen.inc: weights as 16-byte shared-memory broadcasts at compile-time
//    offsets, conflict-free loads of the power bins at pitch 203).  log10, running max/min and the [n_mels][T] store
//    follow; a warp stores 32 consecutive frames of one mel = one 128-byte line.
//  * K5 applies Whisper's global floor in place and skips tiles whose minimum is already above it.
#include <cstdlib>
#include <utility>

#include "b2a_tables.cuh"
#include "f16_bits.h"
#if defined(B2A_PROFILE) || defined(B2A_EMU)
#include "../../tools/probes/logmel_tc.cuh"      // the tensor-core variant (probe; measured, not faster: see its header)
#define B2A_HAVE_LOGMEL_TC 1
#endif

namespace b2a {

constexpr int LM_FRAMES = 32;                       // frames per tile (lane = frame)
constexpr int LM_ROLES = 5;                         // warps per CTA (warp = FFT role / mel group)
constexpr int LM_THREADS = LM_ROLES * 32;
constexpr int LM_TILE = LM_FRAMES * kHop + 240;     // 5360 samples
constexpr int LM_SKEW = 2;                          // extra words per hop (bank de-phasing, keeps 8-byte alignment)
constexpr int LM_HOPW = kHop + LM_SKEW;             // words between consecutive frames in the skewed tile
constexpr int LM_TILE_WORDS = LM_TILE + LM_SKEW * (LM_TILE / kHop + 1);
// exchange pitch per frame, in complex values.  The buffer is accessed as 16-byte pairs of neighbouring columns: rows 1616 bytes
// apart keep every access aligned and the 8 lanes of a quarter warp on distinct 16-byte bank groups (101 f mod 8)
constexpr int LM_EXP = 202;
#ifndef LM_S16_CTAS
#define LM_S16_CTAS 3
#endif
constexpr int LM_GBATCH = 32;                       // gathered tiles whose segment descriptors are looked up together
constexpr int LM_PP = 203;                          // power-spectrum pitch per frame, in floats (odd)

#ifndef B2A_MEL_TABLES_INCLUDED
#define B2A_MEL_TABLES_INCLUDED
#include "mel_tables_gen.inc"
#endif
struct cpx { float r, i; };
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return {a.r + b.r, a.i + b.i}; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return {a.r - b.r, a.i - b.i}; }
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
__device__ __forceinline__ cpx cmuli_neg(cpx a) { return {a.i, -a.r}; }  // -i * a

// 5-point DFT (forward), in place
__device__ __forceinline__ void dft5(cpx& x0, cpx& x1, cpx& x2, cpx& x3, cpx& x4) {
    const float kS1 = 0.95105651629515357f, kS2 = 0.58778525229247313f, kC = 0.55901699437494742f;
    cpx t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    cpx t5 = cadd(t1, t2);
    cpx m1 = {x0.r - 0.25f * t5.r, x0.i - 0.25f * t5.i};
    cpx m2 = {(t1.r - t2.r) * kC, (t1.i - t2.i) * kC};
    cpx s = cadd(m1, m2), d = csub(m1, m2);
    cpx u = {kS1 * t3.r + kS2 * t4.r, kS1 * t3.i + kS2 * t4.i};
    cpx v = {kS2 * t3.r - kS1 * t4.r, kS2 * t3.i - kS1 * t4.i};
    cpx mu = cmuli_neg(u), mv = cmuli_neg(v);
    x0 = cadd(x0, t5);
    x1 = cadd(s, mu);
    x4 = csub(s, mu);
    x2 = cadd(d, mv);
    x3 = csub(d, mv);
}

__device__ __forceinline__ void dft4(cpx& x0, cpx& x1, cpx& x2, cpx& x3) {
    cpx a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = csub(x1, x3);
    cpx md = cmuli_neg(d);
    x0 = cadd(a, c);
    x1 = cadd(b, md);
    x2 = csub(a, c);
    x3 = csub(b, md);
}

// 10-point DFT via 2x5 prime-factor map: n = (5a+2b)%10, k = (5c+6d)%10.  in x[10] -> out y[10]
__device__ __forceinline__ void dft10(const cpx (&x)[10], cpx (&y)[10]) {
    cpx s[5], d[5];
#pragma unroll
    for (int b = 0; b < 5; b++) {
        cpx p = x[(2 * b) % 10], q = x[(5 + 2 * b) % 10];
        s[b] = cadd(p, q);
        d[b] = csub(p, q);
    }
    dft5(s[0], s[1], s[2], s[3], s[4]);
    dft5(d[0], d[1], d[2], d[3], d[4]);
#pragma unroll
    for (int e = 0; e < 5; e++) {
        y[(6 * e) % 10] = s[e];
        y[(5 + 6 * e) % 10] = d[e];
    }
}

// 20-point DFT via 4x5 prime-factor map: n = (5a+4b)%20, k = (5c+16d)%20.
__device__ __forceinline__ void dft20(const cpx (&x)[20], cpx (&y)[20]) {
    cpx r[4][5];
#pragma unroll
    for (int b = 0; b < 5; b++) {
        cpx a0 = x[(4 * b) % 20], a1 = x[(5 + 4 * b) % 20], a2 = x[(10 + 4 * b) % 20], a3 = x[(15 + 4 * b) % 20];
        dft4(a0, a1, a2, a3);
        r[0][b] = a0; r[1][b] = a1; r[2][b] = a2; r[3][b] = a3;
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        dft5(r[c][0], r[c][1], r[c][2], r[c][3], r[c][4]);
#pragma unroll
        for (int d = 0; d < 5; d++) y[(5 * c + 16 * d) % 20] = r[c][d];
    }
}

// real-FFT unpack of one conjugate pair + power, scaled by 4:  Z = FFT200(x_even + i x_odd), m = 200-k.
//   2X[k] = E + t, 2X[200-k] = conj(E - t)  with  E = Zk + conj(Zm),  t = W400^k * (-i)(Zk - conj(Zm)).
// The sums are formed BEFORE squaring (|E|^2 + |t|^2 +- 2Re(..) would cancel catastrophically in quiet bins).
__device__ __forceinline__ void unpack_pair4(cpx zk, cpx zm, float2 tw, float& pk, float& pm) {
    const float er = zk.r + zm.r, ei = zk.i - zm.i;
    const float orr = zk.r - zm.r, oi = zk.i + zm.i;
    const float tr = tw.x * oi - tw.y * orr;
    const float ti = tw.x * orr + tw.y * oi;
    const float ar = er + tr, ai = ei - ti;
    const float br = er - tr, bi = ei + ti;
    pk = fmaf(ar, ar, ai * ai);
    pm = fmaf(br, br, bi * bi);
}

struct LogMelParams {
    const void* audio;     // [batch] rows
    int fmt;               // B2A_FMT_*
    i64 row_stride;        // samples between rows
    i64 n;                 // samples per row (capacity when d_n != nullptr)
    const i64* d_n;        // optional device-side actual length (batch == 1)
    i64 padding;           // zeros appended on the right
    int n_mels;
    int batch;
    float* out;            // [batch][n_mels][T]
    i64* d_frames_out;     // optional
    int* gmax_key;         // [batch] (per-clip) or [1]
    int per_clip;
    int* tile_min;         // [batch][tiles_cap]: float_to_key of the minimum of every 32-frame tile
    i64 tiles_cap;         // tiles per clip at capacity
    const LogMelTables* tab;
    // fused stream compaction (pipeline, s16, batch == 1): when kept_ms != nullptr `audio` is the UNTRIMMED 16 kHz PCM of
    // n_src samples and the clip this kernel sees is the concatenation of the kept ranges (pydub `+`): trimmed index q in
    // [kept_off[s], kept_off[s+1]) lives at source sample 16 kept_ms[2s] + q - kept_off[s] (zero past n_src: pydub
    // zero-fills a rounded-up last millisecond).  Every tile also writes the 5120 trimmed samples it owns to trim_out,
    // so there is no separate compaction kernel and the trimmed PCM is never re-read.
    const int32_t* kept_ms;
    const i64* kept_off;
    const i64* info;       // B2A_INFO_N_KEPT
    int16_t* trim_out;
    i64 n_src;
};

__device__ __forceinline__ float load_sample(const void* audio, int fmt, i64 idx) {
    if (fmt == B2A_FMT_S16) return (float)((const int16_t*)audio)[idx] * (1.0f / 32768.0f);
    return ((const float*)audio)[idx];
}

// ---- mel projection of one frame: warp u owns mels u, u + 5, u + 10, ... (slot i = mel 5 i + u), lane = frame ----
// Every warp runs the same fully unrolled code (mel_tables_gen.inc): slot i takes nt[i] taps (the widest of its five
// filters, the others zero padded), its weights are 16-byte broadcast loads at a compile-time offset of the warp's block,
// and only the first power bin differs between warps (one broadcast load per slot).  History: unrolled per-warp bodies with
// immediate weights overflowed the instruction cache (profiles/r01_logmel_icache.md); a table-driven loop over each warp's
// contiguous group of mels then spent 22 instructions per mel on descriptors, pointers and loop control for 4-16 FFMAs and
// serialised every mel behind its descriptor load (20 % of the kernel's instructions for 5 % of its flops).
// lmax / lmin track log2(mel) (converted once per tile by the caller); only the store is predicated.
__device__ __forceinline__ float lg2_fast(float x) {          // x >= 1e-10: never denormal, one MUFU
#ifndef B2A_EMU
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#else
    return __log2f(x);
#endif
}
// (log10(x) + 4) / 4 from log2(x), with the same roundings as Whisper's two steps (silence comes out as exactly -1.5)
__device__ __forceinline__ float mel_out_value(float l2) { return fmaf(l2 * 0.30102999566398120f, 0.25f, 1.0f); }

// predicated global store: the address is formed unconditionally and the store stays one predicated STG (written as
// `if (ok) *p = v` with a 64-bit address the compiler branches around every store: 16 BSSY / BRA / BSYNC blocks per tile)
__device__ __forceinline__ void st_f32_if(float* p, float v, bool ok) {
#ifndef B2A_EMU
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"((unsigned)ok));
#else
    if (ok) *p = v;
#endif
}

// one slot: returns log2(mel), or NaN for a slot without a mel (fmaxf / fminf skip NaN).  Store addresses are a 64-bit base
// plus a 32-bit offset (one IMAD.WIDE each): a clip's [n_mels][T] block stays below 2^31 floats (checked at launch)
template <int NM, int I>
__device__ __forceinline__ float mel_slot(const float4* __restrict__ wq, const unsigned* __restrict__ so, const char* __restrict__ Pf,
                                          float* __restrict__ obase, unsigned o0, unsigned ostep, bool valid, int u) {
    constexpr int NT = MelC<NM>::nt[I], Q0 = MelC<NM>::qoff[I], NQ = (NT + 3) / 4;
    constexpr bool TAIL = 5 * I + 4 >= NM;                    // the last slot of 128 mels: warps 3 and 4 have no mel
    const float* pf = (const float*)(Pf + so[I]);
    float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        const float4 w = wq[Q0 + q];
        if (q == 0) {
            a0 = w.x * pf[0];
            if (NT > 1) a1 = w.y * pf[1];
        } else {
            a0 = fmaf(w.x, pf[4 * q], a0);
            if (4 * q + 1 < NT) a1 = fmaf(w.y, pf[4 * q + 1], a1);
        }
        if (4 * q + 2 < NT) a0 = fmaf(w.z, pf[4 * q + 2], a0);
        if (4 * q + 3 < NT) a1 = fmaf(w.w, pf[4 * q + 3], a1);
    }
    const float l2 = lg2_fast(fmaxf(a0 + a1, 1e-10f));
    const bool real = !TAIL || 5 * I + u < NM;
    st_f32_if(obase + (o0 + (unsigned)I * ostep), mel_out_value(l2), valid && real);
    return real ? l2 : __int_as_float(0x7fffffff);
}
// two slots and one three-input FMNMX3 each for the running maximum and minimum
template <int NM, int I2>
__device__ __forceinline__ void mel_pair(const float4* __restrict__ wq, const unsigned* __restrict__ so, const char* __restrict__ Pf,
                                         float* __restrict__ obase, unsigned o0, unsigned ostep, bool valid, int u, float& lmax, float& lmin) {
    const float x = mel_slot<NM, 2 * I2>(wq, so, Pf, obase, o0, ostep, valid, u);
    const float y = mel_slot<NM, 2 * I2 + 1>(wq, so, Pf, obase, o0, ostep, valid, u);
    lmax = fmaxf(fmaxf(lmax, x), y);
    lmin = fminf(fminf(lmin, x), y);
}
template <int NM, int... Is>
__device__ __forceinline__ void mel_slots(std::integer_sequence<int, Is...>, const float4* __restrict__ wq, const unsigned* __restrict__ so,
                                          const float* __restrict__ Pf, float* __restrict__ obase, unsigned o0, unsigned ostep, bool valid, int u,
                                          float& lmax, float& lmin) {
    static_assert(MelC<NM>::SLOTS == 2 * sizeof...(Is), "slots are walked in pairs");
    (mel_pair<NM, Is>(wq, so, (const char*)Pf, obase, o0, ostep, valid, u, lmax, lmin), ...);
}

#ifndef B2A_EMU
// 32-bit global load that stays where it is written (memory clobber): the next tile's prefetch must be issued before
// stage 1, not sunk to its first use behind it
__device__ __forceinline__ unsigned ldg_pinned(const unsigned* p) {
    unsigned v;
    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ldg_pinned16(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_drain() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory"); }
#else
static inline unsigned ldg_pinned(const unsigned* p) { return *p; }
static inline uint4 ldg_pinned16(const uint4* p) { return *p; }
static inline void cp_async8(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 8); }
static inline void cp_async_drain() {}
#endif

// Split barriers: two of the four block-wide barriers of a tile are arrive / wait pairs on shared-memory mbarriers (one
// arrival per warp): a warp ARRIVES as soon as it is done with the contested buffer, keeps working on registers, and WAITS
// only where it needs the others.  (bar.arrive + bar.sync on a named barrier cannot do this when every thread does both:
// the sync counts as an arrival too.)  The emulation build turns the pair into one full barrier at the wait.
#ifndef B2A_EMU
__device__ __forceinline__ void lm_bar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void lm_arrive(unsigned long long* bar, int lane) {
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void lm_wait(unsigned long long* bar, unsigned parity) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
#else
static inline void lm_bar_init(unsigned long long*, unsigned) {}
static inline void lm_arrive(unsigned long long*, int) {}
static inline void lm_wait(unsigned long long*, unsigned) { __syncthreads(); }
#endif

constexpr int LM_PAIRS = LM_TILE / 2;                                   // 2680 sample pairs per tile
constexpr int LM_PRE = (LM_PAIRS + LM_THREADS - 1) / LM_THREADS;        // 17 pairs per thread (f32 input: cp.async of pairs)
constexpr int LM_CHUNKS = LM_TILE / 8;                                  // s16 input: 670 chunks of 8 samples (16 bytes) per tile
constexpr int LM_PRE4 = (LM_CHUNKS + LM_THREADS - 1) / LM_THREADS;      // 5 chunks per thread
static_assert(LM_TILE % 8 == 0 && (kHop / 2) % 4 == 0, "a 16-byte chunk never straddles a hop (the skew changes between hops)");

constexpr int LM_MELW_WORDS = LM_ROLES * MelC<128>::QUADS * 4;            // mel weights, one block of QUADS quads per warp (128 mels: the larger)
constexpr int LM_MELOFF_WORDS = (LM_ROLES * MelC<128>::SLOTS + 3) / 4 * 4;  // first power bin (byte offset) of every (warp, slot)
static_assert(MelC<80>::QUADS <= MelC<128>::QUADS && MelC<80>::SLOTS <= MelC<128>::SLOTS, "shared-memory mel tables are sized for 128 mels");

// shared-memory footprint: the s16 variant keeps the tile as raw sample pairs (half the bytes) and both variants put the
// power spectra into the exchange buffer once it has been consumed => 67 KB (s16: 3 CTAs per SM) / 78 KB (f32: 2)
template <int FMT> struct LmSmem {
    static constexpr int TILE_WORDS = ((FMT == B2A_FMT_S16 ? (LM_TILE / 2 + (LM_TILE / kHop + 2)) : LM_TILE_WORDS) + 3) / 4 * 4;   // keeps the buffers behind it 16-byte aligned
    static constexpr int HOPW = FMT == B2A_FMT_S16 ? (kHop / 2 + 1) : LM_HOPW;      // words between frames (81: odd, conflict free)
    static constexpr int WORDS = TILE_WORDS + 2 * LM_FRAMES * LM_EXP + kNFFT + 2 * 200 + 2 * 202 + 12 + LM_MELW_WORDS + LM_MELOFF_WORDS + 12 * LM_GBATCH + 4 * LM_GBATCH;   // + LM_GBATCH x 6 int64 gather descriptors
    static constexpr int CTAS = FMT == B2A_FMT_S16 ? LM_S16_CTAS : 2;
};

template <int NM, int FMT, bool GATHER>
__global__ void __launch_bounds__(LM_THREADS, LmSmem<FMT>::CTAS) stft_mel_kernel(LogMelParams p) {
    using SM = LmSmem<FMT>;
    constexpr bool S16 = FMT == B2A_FMT_S16;
    B2A_DYN_SMEM(smem_raw);
    float* s_tile = (float*)smem_raw;                                   // SM::TILE_WORDS (f32 samples, or s16 sample pairs)
    unsigned* s_tile16 = (unsigned*)smem_raw;
    float2* s_ex = (float2*)(s_tile + SM::TILE_WORDS);                  // LM_FRAMES * LM_EXP complex
    float* s_P = (float*)s_ex;                                          // LM_FRAMES * LM_PP, reuses the exchange buffer
    float* s_win = (float*)(s_ex + LM_FRAMES * LM_EXP);                 // 400
    float2* s_tw200 = (float2*)(s_win + kNFFT);                         // 200
    float2* s_tw400 = s_tw200 + 200;                                    // 201 (+1 pad)
    float* s_red = (float*)(s_tw400 + 202);                             // 8
    unsigned long long* s_bar = (unsigned long long*)(s_red + 8);       // 2 mbarriers: [0] tile done (power spectra consumed, next samples in place), [1] exchange consumed
    float4* s_flat4 = (float4*)(s_red + 12);                            // mel weights: [warp][QUADS] quads (16-byte aligned)
    unsigned* s_moff = (unsigned*)(s_flat4 + LM_MELW_WORDS / 4);         // [warp][SLOTS] byte offset of the slot's first power bin
    int4* s_gi = (int4*)(s_moff + LM_MELOFF_WORDS);                       // [LM_GBATCH] interior gathered tiles in 32 bits: {first source chunk, split chunk, delta chunks, mode}
    i64* s_g = (i64*)(s_gi + LM_GBATCH);                          // [LM_GBATCH][6]: lo0, add0, bound1, add1, bound2, mode of the CTA's next gathered tiles

    const int tid = threadIdx.x;
    const LogMelTables* tab = p.tab;
    for (int i = tid; i < kNFFT; i += LM_THREADS) s_win[i] = tab->win[i] * (S16 ? (1.0f / 32768.0f) : 1.0f);   // s16 tile holds raw integers
    for (int i = tid; i < 200; i += LM_THREADS) s_tw200[i] = tab->tw200[i];
    for (int i = tid; i < kNBins; i += LM_THREADS) s_tw400[i] = tab->tw400[i];
    for (int i = tid; i < LM_ROLES * MelC<NM>::QUADS * 4; i += LM_THREADS) ((float*)s_flat4)[i] = NM == 80 ? kMelUW80[i] : kMelUW128[i];
    for (int i = tid; i < LM_ROLES * MelC<NM>::SLOTS; i += LM_THREADS) s_moff[i] = NM == 80 ? kMelUOff80[i] : kMelUOff128[i];
    for (int i = tid; i < LM_FRAMES * LM_EXP; i += LM_THREADS) s_ex[i] = make_float2(0.0f, 0.0f);   // padded taps may read slots no stage writes
    if (tid == 0) { lm_bar_init(s_bar, LM_ROLES); lm_bar_init(s_bar + 1, LM_ROLES); }                // visible to all after the first block-wide barrier

    i64 n_act = p.n;
    if (p.d_n) { n_act = *p.d_n; if (n_act > p.n) n_act = p.n; if (n_act < 0) n_act = 0; }
    const i64 ltot = n_act + p.padding;          // padded length
    const i64 T = ltot / kHop;                   // frames
    constexpr bool gather = GATHER;                                     // fused stream compaction: own instantiation, the plain kernel keeps its registers
    static_assert(!GATHER || S16, "fused compaction is an s16 path");
    const int n_seg = gather ? (int)p.info[B2A_INFO_N_KEPT] : 0;
    i64 tiles = (T + LM_FRAMES - 1) / LM_FRAMES;
    if (gather && tiles * (LM_FRAMES * kHop) < n_act) tiles++;          // a last partial hop still has trimmed samples to write
    if (blockIdx.x == 0 && tid == 0 && p.d_frames_out) *p.d_frames_out = T;

    const int u = tid >> 5, f = tid & 31;        // role (warp-uniform), frame within the tile
    const int k1a = u, k1b = (u == 0) ? 5 : 10 - u;
    // role u owns the columns n2 = 4u .. 4u + 3 (neighbours in the exchange buffer: one 16-byte store per pair)
    const float* ps = s_tile + LM_HOPW * f + 8 * u;          // f32 tile: + 40 n1 + 2 j + 2 (n1 / 4)
    const unsigned* ps16 = s_tile16 + SM::HOPW * f + 4 * u;   // s16 tile (words = sample pairs): + 20 n1 + j + (n1 / 4)
    const float* pw = s_win + 8 * u;                         // + 40 n1 + 2 j
    const float2* pt = s_tw200 + 40 * u;                     // + 10 j + k1
    float2* pex_w = s_ex + LM_EXP * f + 4 * u;               // + 20 k1 + j
    constexpr int JS = 1, JP = 2, JT = 10;                   // per-column steps of the tile (pairs), the f32 tile / window (floats), the twiddles
    const float2* pex_a = s_ex + LM_EXP * f + 20 * k1a;      // + n2
    const float2* pex_b = s_ex + LM_EXP * f + 20 * k1b;
    float* pP = s_P + LM_PP * f;
    const int elem = S16 ? 2 : 4;
    const float4* mel_wq = s_flat4 + u * MelC<NM>::QUADS;    // this warp's mels: u, u + 5, u + 10, ...
    const unsigned* mel_so = s_moff + u * MelC<NM>::SLOTS;

    float run_max = -3.0e38f;
    i64 prev_slot = -1;                          // tile_min slot of the previous work item (written one barrier later)
    const i64 n_work = tiles * p.batch;

    // where a work item's samples live: row base, first padded-domain index, and whether the fast (interior) path applies
    // work item -> (clip b, tile): one clip needs no division, batches stay in 32 bits (a 64-bit division is ~70 dependent
    // instructions at the top of every tile, ahead of the barrier everybody waits at)
    const bool one_clip = p.batch == 1;
    const bool small_work = tiles * p.batch < ((i64)1 << 31);
    auto split_work = [&](i64 work, int& b, i64& tile) {
        if (GATHER || one_clip) { b = 0; tile = work; }                      // (the gathering instantiation is always one clip)
        else if (small_work) { b = (int)((unsigned)work / (unsigned)tiles); tile = (i64)((unsigned)work - (unsigned)b * (unsigned)tiles); }
        else { b = (int)(work / tiles); tile = work - (i64)b * tiles; }
    };
    const bool src_vec = (((uintptr_t)p.audio) & 15) == 0;                 // gathered tiles: 16-byte loads from the untrimmed PCM
    const bool trim_vec = gather && (((uintptr_t)p.trim_out) & 15) == 0;   // ... and 16-byte stores of the trimmed PCM
    struct Src { const char* row; i64 q0; int mode; const i64* g; const int4* gi; };   // mode 0: generic (reflect / zero pad), 1: interior s16, 2: interior f32; g: gather descriptor (shared memory)
    // gather: source sample of trimmed index q (segment search; the two segments cached per tile cover an interior tile)
    auto seg_of = [&](i64 q) -> int {
        int lo = 0, hi = n_seg - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (p.kept_off[mid] <= q) lo = mid; else hi = mid - 1;
        }
        return lo;
    };
    // gathered tiles: every LM_GBATCH iterations the first LM_GBATCH threads look up the segments under the CTA's next
    // tiles (one binary search each, in parallel) and publish lo0, add0, bound1, add1, bound2, mode in shared memory
    auto lookup_batch = [&](i64 work0) {
        if (tid < LM_GBATCH) {
            const i64 work = work0 + (i64)tid * gridDim.x;
            if (work < tiles * p.batch) {
                const i64 q0 = (work % tiles) * (LM_FRAMES * kHop) - 200;
                i64 lo0 = 0, add0 = 0, add1 = 0, bound1 = (i64)1 << 62, bound2 = (i64)1 << 62;
                if (n_seg > 0) {
                    const int sg = seg_of(q0 > 0 ? q0 : 0);
                    lo0 = p.kept_off[sg];
                    add0 = (i64)p.kept_ms[2 * sg] * 16 - lo0;
                    if (sg + 1 < n_seg) { bound1 = p.kept_off[sg + 1]; add1 = (i64)p.kept_ms[2 * sg + 2] * 16 - bound1; }
                    if (sg + 2 < n_seg) bound2 = p.kept_off[sg + 2];
                }
                // interior: inside the clip, at most two segments under the tile, every source sample inside the buffer
                const bool interior = src_vec && q0 >= 0 && q0 + LM_TILE <= n_act && q0 + LM_TILE <= bound2 && add0 + q0 >= 0 &&
                                      add0 + (bound1 < q0 + LM_TILE ? bound1 : q0 + LM_TILE) <= p.n_src &&
                                      (bound1 >= q0 + LM_TILE || add1 + q0 + LM_TILE <= p.n_src);
                i64* g = s_g + 6 * tid;
                g[0] = lo0; g[1] = add0; g[2] = bound1; g[3] = add1; g[4] = bound2; g[5] = interior ? 1 : 0;
                // what the interior loader needs, in chunks of 8 samples and 32 bits (a clip's source offsets fit: n_src < 2^34):
                // chunk k of the tile is source chunk first + k (+ delta from the split on, where the second segment starts)
                const i64 split = (bound1 - q0) >> 3;
                s_gi[tid] = make_int4((int)((add0 + q0) >> 3), (int)(split < LM_CHUNKS ? split : LM_CHUNKS), (int)((add1 - add0) >> 3),
                                      interior && p.n_src < ((i64)1 << 34) ? 1 : 0);
            }
        }
    };
    auto locate = [&](i64 work, int slot) -> Src {
        Src r;
        int b;
        i64 tile;
        split_work(work, b, tile);
        r.row = GATHER ? (const char*)p.audio : (const char*)p.audio + (size_t)b * (size_t)p.row_stride * elem;
        r.q0 = tile * (LM_FRAMES * kHop) - 200;
        r.g = s_g + 6 * slot;
        r.gi = s_gi + slot;
        if (gather) {
            r.mode = r.gi->w;
        } else {
            const bool interior = r.q0 >= 0 && r.q0 + LM_TILE <= n_act && ((((uintptr_t)(r.row + r.q0 * elem)) & 7) == 0);
            r.mode = interior ? (S16 ? 1 : 2) : 0;
        }
        return r;
    };
    // trimmed-domain sample q of a gathered clip (0 <= q < n_act)
    auto gather_sample = [&](const Src& sc, i64 q) -> short {
        i64 si;
        if (q >= sc.g[0] && q < sc.g[2]) si = sc.g[1] + q;                    // the two segments cached for the tile
        else if (q >= sc.g[2] && q < sc.g[4]) si = sc.g[3] + q;
        else { const int sg = seg_of(q); si = (i64)p.kept_ms[2 * sg] * 16 + (q - p.kept_off[sg]); }
        return (si >= 0 && si < p.n_src) ? ((const short*)p.audio)[si] : (short)0;
    };
    // generic tile load: padded-domain index q = q0 + i, reflect at both ends, zeros past n_act
    auto load_generic = [&](const Src& sc) {
        for (int i = tid; i < LM_TILE; i += LM_THREADS) {
            i64 q = sc.q0 + i;
            const bool own = gather && q >= 0 && q < n_act && i >= 200 && i < 200 + LM_FRAMES * kHop;   // the tile's own 5120 samples
            if (q < 0) q = -q;
            if (q >= ltot) q = 2 * (ltot - 1) - q;
            if (S16) {
                short v = 0;
                if (q >= 0 && q < n_act) v = gather ? gather_sample(sc, q) : ((const short*)sc.row)[q];
                ((short*)s_tile16)[i + 2 * (i / kHop)] = v;                     // one word of skew per hop
                if (own) p.trim_out[q] = v;
            } else {
                float v = 0.0f;
                if (q >= 0 && q < n_act) v = ((const float*)sc.row)[q];
                s_tile[i + LM_SKEW * (i / kHop)] = v;
            }
        }
    };
    // gathered s16 tiles travel as 16-byte chunks of 8 samples: a tile starts 200 samples (400 bytes) before a multiple of 5120 samples,
    // segment bounds are multiples of 16 samples in both domains, so chunk k of a tile is aligned in the source, in the trimmed
    // output and against the segment split whenever the buffers themselves are (checked where the mode is chosen).
    // chunk k (pairs 4k .. 4k+3) lives at words 4k + k / 20 .. + 3 of the skewed tile (one word of skew per hop of 80 pairs)
    // k = tid + 160 i, so k / 20 = tid / 20 + 8 i: every address below is one per-thread base plus a compile-time offset
    static_assert(LM_THREADS % 20 == 0, "the chunk skew advances by whole words per round");
    unsigned* const st_chunk0 = s_tile16 + 4 * tid + tid / 20;
    auto store_s16_chunks = [&](const Src& sc, const uint4 (&pre)[LM_PRE4]) {
        uint4* const t16 = (uint4*)p.trim_out + ((sc.q0 >> 3) + tid);             // (used when trim_vec)
        unsigned* const t4 = (unsigned*)p.trim_out + ((sc.q0 >> 1) + 4 * tid);
#pragma unroll
        for (int i = 0; i < LM_PRE4; i++) {
            const int k0 = LM_THREADS * i;                                         // k = k0 + tid
            if (k0 + LM_THREADS <= LM_CHUNKS || tid < LM_CHUNKS - k0) {
                unsigned* d = st_chunk0 + (4 * k0 + k0 / 20);                      // raw pairs, converted where they are used
                d[0] = pre[i].x; d[1] = pre[i].y; d[2] = pre[i].z; d[3] = pre[i].w;
                // the tile's own 5120 samples: chunks 25 .. 664
                const bool own = (k0 >= 25 || tid >= 25 - k0) && (k0 + LM_THREADS <= 25 + LM_FRAMES * kHop / 8 || tid < 25 + LM_FRAMES * kHop / 8 - k0);
                if (own) {
                    if (trim_vec) t16[k0] = pre[i];
                    else {
                        unsigned* o = t4 + 4 * k0;
                        o[0] = pre[i].x; o[1] = pre[i].y; o[2] = pre[i].z; o[3] = pre[i].w;
                    }
                }
            }
        }
    };
    auto fetch_s16_chunks = [&](const Src& sc, uint4 (&pre)[LM_PRE4]) {
        const int4 d = *sc.gi;                                                  // {first source chunk, split, delta, mode}
        const uint4* g0 = (const uint4*)p.audio + ((unsigned)d.x + (unsigned)tid);
#pragma unroll
        for (int i = 0; i < LM_PRE4; i++) {
            const int k0 = LM_THREADS * i;                                         // k = k0 + tid
            if (k0 + LM_THREADS <= LM_CHUNKS || tid < LM_CHUNKS - k0) pre[i] = ldg_pinned16(g0 + (k0 + (k0 + tid >= d.y ? d.z : 0)));
        }
    };
    // plain (not gathered) s16 tiles keep the 4-byte form: pair pr (samples 2pr, 2pr+1) lives at word pr + pr / 80 of the skewed
    // tile; 17 loads per thread in flight measured 2 % faster than 5 wide ones there, and rows need only 4-byte alignment
    auto store_s16_pairs = [&](const unsigned (&pre)[LM_PRE]) {         // (per-thread base + compile-time offsets measured 2 % slower here: spills)
#pragma unroll
        for (int i = 0; i < LM_PRE; i++) {
            const int pr = tid + LM_THREADS * i;
            if (pr < LM_PAIRS) s_tile16[pr + pr / 80] = pre[i];
        }
    };
    auto fetch_s16_pairs = [&](const Src& sc, unsigned (&pre)[LM_PRE]) {
        const unsigned* g = (const unsigned*)(sc.row + sc.q0 * 2);
#pragma unroll
        for (int i = 0; i < LM_PRE; i++) {
            const int pr = tid + LM_THREADS * i;
            pre[i] = pr < LM_PAIRS ? ldg_pinned(g + pr) : 0u;
        }
    };
    auto copy_f32_pairs = [&](const Src& sc) {
        const float2* g = (const float2*)(sc.row + sc.q0 * 4);
#pragma unroll
        for (int i = 0; i < LM_PRE; i++) {
            const int pr = tid + LM_THREADS * i;
            if (pr < LM_PAIRS) cp_async8(s_tile + 2 * pr + LM_SKEW * (pr / 80), g + pr);
        }
    };

    // first tile of this CTA: synchronous
    if ((i64)blockIdx.x < n_work) {
        if (gather) { lookup_batch(blockIdx.x); __syncthreads(); }
        const Src sc = locate(blockIdx.x, 0);
        if (sc.mode == 1) {
            if constexpr (GATHER) { uint4 pre[LM_PRE4]; fetch_s16_chunks(sc, pre); store_s16_chunks(sc, pre); }
            else { unsigned pre[LM_PRE]; fetch_s16_pairs(sc, pre); store_s16_pairs(pre); }
        }
        else if (!S16 && sc.mode == 2) { copy_f32_pairs(sc); cp_async_drain(); }
        else load_generic(sc);
    }

    unsigned it = 0;                                  // this CTA's iteration count: mbarrier phase parities
    for (i64 work = blockIdx.x; work < n_work; work += gridDim.x, it++) {
        int b;
        i64 tile;
        split_work(work, b, tile);
        const i64 t0 = tile * LM_FRAMES;
        const i64 nwork = work + gridDim.x;
        Src nsc;
        nsc.mode = -1;
        nsc.g = s_g;
        nsc.gi = s_gi;
        const int nit = (int)it + 1;                                             // iteration index of the next tile
        if (gather && nit % LM_GBATCH == 0) {
            __syncthreads();                                                       // every reader of the old batch is done
            lookup_batch(nwork);
            __syncthreads();
        }
        if (nwork < n_work) nsc = locate(nwork, nit % LM_GBATCH);

        // s_tile (and, first time, the tables) visible; previous tile's s_red written and its power spectra consumed
        if (it == 0) __syncthreads(); else lm_wait(s_bar, (it - 1) & 1);
        if (tid == 0 && prev_slot >= 0) {
            const float l2 = fminf(fminf(fminf(s_red[0], s_red[1]), fminf(s_red[2], s_red[3])), s_red[4]);
            p.tile_min[prev_slot] = float_to_key(mel_out_value(l2));
        }
        // next tile's s16 samples travel global -> registers while stage 1 runs
        uint4 pre4[GATHER ? LM_PRE4 : 1];
        unsigned pre[GATHER ? 1 : LM_PRE];
        if (nsc.mode == 1) {
            if constexpr (GATHER) fetch_s16_chunks(nsc, pre4);
            else fetch_s16_pairs(nsc, pre);
        }

        // ---- stage 1: radix-10 butterflies over n1 for n2 = u + 5j, twiddle, transpose into s_ex ----
        // one column: window, radix-10 butterflies, stage twiddles -> v[k1]
        auto column = [&](int j, cpx (&v)[10]) {
            cpx x[10], y[10];
#pragma unroll
            for (int n1 = 0; n1 < 10; n1++) {
                float2 xv;
                if (S16) {
                    const unsigned w = ps16[JS * j + 20 * n1 + n1 / 4];
                    xv = make_float2((float)(short)(w & 0xffff), (float)(short)(w >> 16));
                } else {
                    xv = *(const float2*)(ps + JP * j + 40 * n1 + LM_SKEW * (n1 / 4));
                }
                const float2 wv = *(const float2*)(pw + JP * j + 40 * n1);
                x[n1].r = xv.x * wv.x;
                x[n1].i = xv.y * wv.y;
            }
            dft10(x, y);
#pragma unroll
            for (int k1 = 0; k1 < 10; k1++) {
                const float2 tw = pt[JT * j + k1];
                const cpx w = {tw.x, tw.y};
                v[k1] = (k1 == 0) ? y[0] : cmul(y[k1], w);
            }
        };
#pragma unroll 1
        for (int j = 0; j < 4; j += 2) {
            cpx va[10], vb[10];
            column(j, va);
            column(j + 1, vb);
#pragma unroll
            for (int k1 = 0; k1 < 10; k1++) *(float4*)(pex_w + j + 20 * k1) = make_float4(va[k1].r, va[k1].i, vb[k1].r, vb[k1].i);
        }
        __syncthreads();   // exchange complete; every stage-1 read of s_tile is done

        // ---- next tile -> s_tile (its registers are free again before stage 2 needs them) ----
        if (nsc.mode == 1) {
            if constexpr (GATHER) store_s16_chunks(nsc, pre4);
            else store_s16_pairs(pre);
        }
        else if (!S16 && nsc.mode == 2) copy_f32_pairs(nsc);
        else if (nsc.mode == 0) load_generic(nsc);

        // ---- stage 2: radix-20 butterflies over n2 for residues k1a, k1b; unpack conjugate pairs in registers ----
        {
            cpx za[20], zb[20];
            {
                cpx xa[20], xb[20];
#pragma unroll
                for (int n2 = 0; n2 < 20; n2 += 2) { const float4 v = *(const float4*)(pex_a + n2); xa[n2] = {v.x, v.y}; xa[n2 + 1] = {v.z, v.w}; }
#pragma unroll
                for (int n2 = 0; n2 < 20; n2 += 2) { const float4 v = *(const float4*)(pex_b + n2); xb[n2] = {v.x, v.y}; xb[n2 + 1] = {v.z, v.w}; }
                lm_arrive(s_bar + 1, f);       // this warp has consumed its part of the exchange buffer ...
                dft20(xa, za);
                dft20(xb, zb);
                lm_wait(s_bar + 1, it & 1);    // ... and the power spectra may overwrite it once every warp has
            }
            if (u != 0) {
                // Z[k] = za[j] (k = u+10j),  Z[200-k] = zb[19-j]
                const float2* ptw = s_tw400 + u;
                float* p0 = pP + u;
                float* p1 = pP + 200 - u;
#pragma unroll
                for (int j = 0; j < 20; j++) {
                    float pk, pm;
                    unpack_pair4(za[j], zb[19 - j], ptw[10 * j], pk, pm);
                    p0[10 * j] = pk;
                    p1[-10 * j] = pm;
                }
            } else {
                // residue 0 pairs with itself: (10j, 200-10j); residue 5: (5+10j, 195-10j)
#pragma unroll
                for (int j = 0; j <= 10; j++) {
                    float pk, pm;
                    unpack_pair4(za[j], za[(20 - j) % 20], s_tw400[10 * j], pk, pm);
                    pP[10 * j] = pk;
                    pP[200 - 10 * j] = pm;
                }
#pragma unroll
                for (int j = 0; j < 10; j++) {
                    float pk, pm;
                    unpack_pair4(zb[j], zb[19 - j], s_tw400[5 + 10 * j], pk, pm);
                    pP[5 + 10 * j] = pk;
                    pP[195 - 10 * j] = pm;
                }
            }
        }
        __syncthreads();   // power spectra complete

        // ---- mel projection + log10 + store: warp u = mels u, u + 5, ..., lane = frame ----
        {
            const i64 t = t0 + f;
            const bool valid = t < T;
            float* obase = p.out + (size_t)b * (size_t)NM * (size_t)T + (valid ? t : 0);
            float lmax = -3.0e38f, lmin = 3.0e38f;
            mel_slots<NM>(std::make_integer_sequence<int, MelC<NM>::SLOTS / 2>{}, mel_wq, mel_so, pP, obase, (unsigned)u * (unsigned)T, (unsigned)LM_ROLES * (unsigned)T, valid, u, lmax, lmin);
            if (valid) run_max = fmaxf(run_max, lmax * 0.30102999566398120f);
            // per-tile minimum (lets mel_floor skip tiles that need no clamping): published after the next barrier
            lmin = warp_reduce_min_f(valid ? lmin : 3.0e38f);
            if (f == 0) s_red[u] = lmin;
        }
        prev_slot = (i64)b * p.tiles_cap + tile;
        if (p.per_clip) {
            const float bm = warp_reduce_max_f(run_max);
            if (f == 0 && bm > -1.0e38f) atomicMax(p.gmax_key + b, float_to_key(bm));
            run_max = -3.0e38f;
        }
        if (!S16 && nsc.mode == 2) cp_async_drain();
        if (nwork < n_work) lm_arrive(s_bar, f);   // done with this tile's power spectra; the next tile's samples are in place
    }
    __syncthreads();
    if (tid == 0 && prev_slot >= 0) {
        const float l2 = fminf(fminf(fminf(s_red[0], s_red[1]), fminf(s_red[2], s_red[3])), s_red[4]);
        p.tile_min[prev_slot] = float_to_key(mel_out_value(l2));
    }
    if (!p.per_clip) {
        const float bm = warp_reduce_max_f(run_max);
        if (f == 0 && bm > -1.0e38f) atomicMax(p.gmax_key, float_to_key(bm));
    }
}

// K5: out = max(out, ((gmax - 8) + 4) / 4) in place; tiles whose minimum already clears the floor are skipped.
// A block screens `tpc` consecutive 32-frame tiles (one thread each: one load of the tile's minimum), collects the tiles below
// the floor in shared memory, and then all its warps walk that list in items of (tile, 16 rows): lane = frame, 16 loads in
// flight per lane.  The grid is a single wave (tpc = tiles / 296 for one clip), so the kernel is about four dependent memory
// latencies long.  (Round 1: one warp per tile walked its 80 rows alone while the warps of clean tiles exited: 16 us at cfg2
// in almost three waves of blocks.)
constexpr int LM_FLOOR_THREADS = 512;
template <int NM>
__global__ void __launch_bounds__(LM_FLOOR_THREADS, 2) mel_floor_kernel(LogMelParams p, int tpc) {
    constexpr int R = 16, GROUPS = NM / R;
    static_assert(NM % R == 0, "row groups of 16");
    __shared__ int s_count;
    __shared__ int s_b[LM_FLOOR_THREADS], s_tile[LM_FLOOR_THREADS];
    __shared__ float s_floor[LM_FLOOR_THREADS];
    i64 n_act = p.n;
    if (p.d_n) { n_act = *p.d_n; if (n_act > p.n) n_act = p.n; if (n_act < 0) n_act = 0; }
    const i64 T = (n_act + p.padding) / kHop;
    const i64 tiles = (T + LM_FRAMES - 1) / LM_FRAMES;
    const i64 total = tiles * p.batch;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_count = 0;
    __syncthreads();
    const i64 work = (i64)blockIdx.x * tpc + tid;
    if (tid < tpc && work < total) {
        const int b = (int)(work / tiles);
        const int tile = (int)(work - (i64)b * tiles);
        const float floor_l = ((key_to_float(p.gmax_key[p.per_clip ? b : 0]) - 8.0f) + 4.0f) * 0.25f;
        if (key_to_float(p.tile_min[(size_t)b * (size_t)p.tiles_cap + (size_t)tile]) < floor_l) {
            const int slot = atomicAdd(&s_count, 1);
            s_b[slot] = b; s_tile[slot] = tile; s_floor[slot] = floor_l;
        }
    }
    __syncthreads();
    const int items = s_count * GROUPS;
    for (int it = warp; it < items; it += LM_FLOOR_THREADS / 32) {
        const int e = it / GROUPS, g = it - e * GROUPS;
        const i64 t = (i64)s_tile[e] * LM_FRAMES + lane;
        if (t >= T) continue;
        const float floor_v = s_floor[e];
        float* q = p.out + ((size_t)s_b[e] * (size_t)NM + (size_t)g * R) * (size_t)T + t;
        float v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = q[(size_t)r * (size_t)T];
#pragma unroll
        for (int r = 0; r < R; r++)
            if (v[r] < floor_v) q[(size_t)r * (size_t)T] = floor_v;
    }
}

// gmax keys start at -inf; the per-group minima at +inf (the tensor-core kernel lowers them with atomicMin, one per role)
__global__ void logmel_init_kernel(int* gmax_key, int n, int* tile_min_key, i64 n_groups) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gmax_key[i] = float_to_key(-3.0e38f);
    if (tile_min_key && i < n_groups) tile_min_key[i] = float_to_key(3.0e38f);
}

template <int FMT> static size_t logmel_smem_bytes() { return (size_t)LmSmem<FMT>::WORDS * 4; }

static i64 logmel_tiles_cap(i64 n, i64 padding) {
    i64 T = (n + padding) / kHop;
    return (T + LM_FRAMES - 1) / LM_FRAMES + 1;     // + 1: fused compaction adds a tile for a last partial hop
}

// workspace layout: [gmax keys: batch ints, padded to 256 B][tile_min: batch * tiles_cap floats]
size_t logmel_workspace_bytes(i64 batch, i64 n, i64 padding) {
    return align_up((size_t)batch * 4, 256) + align_up((size_t)batch * (size_t)logmel_tiles_cap(n, padding) * 4, 256) + 256;
}

struct LogMelGather { const int32_t* kept_ms; const i64* kept_off; const i64* info; int16_t* trim_out; i64 n_src; };

int logmel_launch(const void* d_audio, int fmt, i64 batch, i64 n, i64 row_stride, const i64* d_n, i64 padding,
                  int n_mels, int norm_mode, float* d_out, i64* d_frames_out, void* d_ws, size_t ws_bytes,
                  cudaStream_t stream, const LogMelGather* gather = nullptr) {
    if (!d_audio || !d_out || !d_ws) { set_error("log_mel: null pointer"); return B2A_EINVAL; }
    if (fmt != B2A_FMT_S16 && fmt != B2A_FMT_F32) { set_error("log_mel: bad fmt %d", fmt); return B2A_EINVAL; }
    if (batch <= 0 || n < 0 || padding < 0 || row_stride < n) { set_error("log_mel: bad shape"); return B2A_EINVAL; }
    if (d_n && batch != 1) { set_error("log_mel: device-side length needs batch == 1"); return B2A_EINVAL; }
    if (!d_n && n + padding <= 200) { set_error("log_mel: need more than 200 samples (reflect pad), got %lld", (long long)(n + padding)); return B2A_EINVAL; }
    if (norm_mode != B2A_NORM_WHISPER && norm_mode != B2A_NORM_PER_CLIP) { set_error("log_mel: bad norm_mode"); return B2A_EINVAL; }
    if ((n + padding) / kHop * (i64)n_mels >= ((i64)1 << 31)) {           // 46 hours at 128 mels: the kernel addresses a clip's block with 32-bit offsets
        set_error("log_mel: a clip's [n_mels][T] block must stay below 2^31 values (n + padding = %lld)", (long long)(n + padding));
        return B2A_EUNSUPPORTED;
    }
    if (ws_bytes < logmel_workspace_bytes(batch, n, padding)) { set_error("log_mel: workspace too small"); return B2A_EWORKSPACE; }
    const LogMelTables* tab = get_logmel_tables(n_mels);
    if (!tab) return B2A_EINVAL;

    LogMelParams p;
    p.audio = d_audio; p.fmt = fmt; p.row_stride = row_stride; p.n = n; p.d_n = d_n; p.padding = padding;
    p.n_mels = n_mels; p.batch = (int)batch; p.out = d_out; p.d_frames_out = d_frames_out;
    p.gmax_key = (int*)d_ws;
    p.per_clip = (norm_mode == B2A_NORM_PER_CLIP);
    p.tile_min = (int*)((char*)d_ws + align_up((size_t)batch * 4, 256));
    p.tiles_cap = logmel_tiles_cap(n, padding);
    p.tab = tab;
    p.kept_ms = nullptr; p.kept_off = nullptr; p.info = nullptr; p.trim_out = nullptr; p.n_src = 0;
    if (gather) {
        if (fmt != B2A_FMT_S16 || batch != 1 || !d_n) { set_error("log_mel: fused compaction needs one s16 clip with a device-side length"); return B2A_EINVAL; }
        p.kept_ms = gather->kept_ms; p.kept_off = gather->kept_off; p.info = gather->info; p.trim_out = gather->trim_out; p.n_src = gather->n_src;
    }

    i64 work = p.tiles_cap * batch;
    if (work <= 0) { if (d_frames_out) cudaMemsetAsync(d_frames_out, 0, 8, stream); return B2A_OK; }
    int nkeys = (int)batch;
    const bool s16 = fmt == B2A_FMT_S16;
    bool use_tc = false;
#ifdef B2A_HAVE_LOGMEL_TC
    if (s16) { const char* impl = getenv("B2A_LM_IMPL"); use_tc = impl && impl[0] == 't'; }   // profiling / emulation builds only
#endif
    auto kinit = logmel_init_kernel;
    {
        const i64 n_groups = use_tc ? p.tiles_cap * batch : 0;   // (the tensor-core probe lowers the group minima with atomicMin)
        const i64 n_init = n_groups > nkeys ? n_groups : nkeys;
        B2A_LAUNCH(kinit, (unsigned)((n_init + 255) / 256), 256, 0, stream, p.gmax_key, nkeys, use_tc ? p.tile_min : (int*)nullptr, n_groups);
    }
#ifdef B2A_HAVE_LOGMEL_TC
    if (use_tc) {
        // the tensor-core probe kernel: 128-frame tiles, minima per 32-frame group as the FFT kernel
        const unsigned char* blob = get_logmel_tc_blob();
        if (!blob) return B2A_ECUDA;
        LogMelTcParams q;
        q.audio = (const int16_t*)d_audio; q.row_stride = row_stride; q.n = n; q.d_n = d_n; q.padding = padding; q.batch = (int)batch;
        q.out = d_out; q.d_frames_out = d_frames_out; q.gmax_key = p.gmax_key; q.per_clip = p.per_clip;
        q.tile_min_key = p.tile_min; q.groups_cap = p.tiles_cap; q.blob = blob;
        q.kept_ms = p.kept_ms; q.kept_off = p.kept_off; q.info = p.info; q.trim_out = p.trim_out; q.n_src = p.n_src;
        const i64 tiles128 = ((n + padding) / kHop + kTcFrames - 1) / kTcFrames + (gather ? 1 : 0);
        const i64 work_tc = tiles128 * batch;
        unsigned grid_tc = (unsigned)(work_tc < 148 ? work_tc : 148);
#if defined(B2A_PROFILE) || defined(B2A_EMU)
        if (const char* gs = getenv("B2A_LM_GRID")) {            // emulation / profiling builds only: few CTAs => many tiles per CTA
            const int gv = atoi(gs);
            if (gv > 0 && (unsigned)gv < grid_tc) grid_tc = (unsigned)gv;
        }
#endif
        static unsigned long long attr_mask = 0;             // per-device opt-in to > 48 KB of dynamic shared memory
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !((attr_mask >> dev) & 1ull)) {
            cudaError_t e = cudaFuncSetAttribute(logmel_tc_kernel<80, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(logmel_tc_kernel<80, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(logmel_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(logmel_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(logmel_tc_kernel)");
            if (dev >= 0 && dev < 64) attr_mask |= 1ull << dev;
        }
        if (grid_tc > 0) {
            if (n_mels == 80) {
                if (gather) { auto k = logmel_tc_kernel<80, true>; B2A_LAUNCH(k, grid_tc, kTcThreads, kTcSmemBytes, stream, q); }
                else { auto k = logmel_tc_kernel<80, false>; B2A_LAUNCH(k, grid_tc, kTcThreads, kTcSmemBytes, stream, q); }
            } else {
                if (gather) { auto k = logmel_tc_kernel<128, true>; B2A_LAUNCH(k, grid_tc, kTcThreads, kTcSmemBytes, stream, q); }
                else { auto k = logmel_tc_kernel<128, false>; B2A_LAUNCH(k, grid_tc, kTcThreads, kTcSmemBytes, stream, q); }
            }
            B2A_CHECK_LAUNCH("logmel_tc_kernel");
        }
    } else
#endif
    {
    size_t smem = s16 ? logmel_smem_bytes<B2A_FMT_S16>() : logmel_smem_bytes<B2A_FMT_F32>();
    auto k4 = gather ? (n_mels == 80 ? stft_mel_kernel<80, B2A_FMT_S16, true> : stft_mel_kernel<128, B2A_FMT_S16, true>)
              : n_mels == 80 ? (s16 ? stft_mel_kernel<80, B2A_FMT_S16, false> : stft_mel_kernel<80, B2A_FMT_F32, false>)
                             : (s16 ? stft_mel_kernel<128, B2A_FMT_S16, false> : stft_mel_kernel<128, B2A_FMT_F32, false>);
    {
        cudaError_t e = cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // idempotent, cheap
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(stft_mel)");
    }
    const i64 resident = 148 * (s16 ? LmSmem<B2A_FMT_S16>::CTAS : LmSmem<B2A_FMT_F32>::CTAS);   // persistent: every CTA resident
    i64 grid = work < resident ? work : resident;
    B2A_LAUNCH(k4, (unsigned)grid, LM_THREADS, smem, stream, p);
    B2A_CHECK_LAUNCH("stft_mel_kernel");
    }
    i64 tpc = (work + 148 * 2 - 1) / (148 * 2);                            // tiles per block: one wave of two blocks per SM ...
    tpc = tpc < 32 ? 32 : tpc > LM_FLOOR_THREADS ? LM_FLOOR_THREADS : tpc;    // ... of at least a warp's worth, at most a thread each
    const i64 grid5 = (work + tpc - 1) / tpc;
    if (n_mels == 80) {
        auto k5 = mel_floor_kernel<80>;
        B2A_LAUNCH(k5, (unsigned)grid5, LM_FLOOR_THREADS, 0, stream, p, (int)tpc);
    } else {
        auto k5 = mel_floor_kernel<128>;
        B2A_LAUNCH(k5, (unsigned)grid5, LM_FLOOR_THREADS, 0, stream, p, (int)tpc);
    }
    B2A_CHECK_LAUNCH("mel_floor_kernel");
    return B2A_OK;
}

// ---- encoder windows: transcribe's mel[:, seek : seek + n_frames] -> pad_or_trim -> dtype cast, all windows in one launch ----
__device__ __forceinline__ unsigned short f32_to_f16_rn(float v) {
#ifndef B2A_EMU
    unsigned short h;
    asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
    return h;
#else
    return b2a_f16::f32_to_f16(v);
#endif
}

// one block per (window, mel) row; a thread writes two neighbouring frames per round (one 4-byte store for f16 when the row
// starts on a 4-byte boundary, i.e. n_frames even).  Frames at or beyond `content` read as zero (pad_or_trim).
template <bool HALF>
__global__ void __launch_bounds__(256) mel_windows_kernel(const float* __restrict__ mel, int n_mels, i64 T, i64 content, i64 seek0,
                                                          i64 stride, int n_frames, void* __restrict__ out) {
    const i64 row = blockIdx.x;
    const int w = (int)(row / n_mels), m = (int)(row - (i64)w * n_mels);
    const i64 s = seek0 + (i64)w * stride;
    const float* src = mel + (size_t)m * (size_t)T;
    const bool paired = (n_frames & 1) == 0;
    for (int f = 2 * threadIdx.x; f < n_frames; f += 2 * blockDim.x) {
        const i64 t0 = s + f, t1 = t0 + 1;
        const float v0 = (t0 >= 0 && t0 < content) ? src[t0] : 0.0f;
        const float v1 = (f + 1 < n_frames && t1 >= 0 && t1 < content) ? src[t1] : 0.0f;
        const size_t o = (size_t)row * (size_t)n_frames + (size_t)f;
        if (HALF) {
            unsigned short* q = (unsigned short*)out;
            if (paired) *(unsigned*)(q + o) = (unsigned)f32_to_f16_rn(v0) | ((unsigned)f32_to_f16_rn(v1) << 16);
            else { q[o] = f32_to_f16_rn(v0); if (f + 1 < n_frames) q[o + 1] = f32_to_f16_rn(v1); }
        } else {
            float* q = (float*)out;
            if (paired) *(float2*)(q + o) = make_float2(v0, v1);
            else { q[o] = v0; if (f + 1 < n_frames) q[o + 1] = v1; }
        }
    }
}

int mel_windows_launch(const float* d_mel, int n_mels, i64 T, i64 content, i64 seek0, i64 stride, int n_win, int n_frames,
                       int out_fmt, void* d_out, cudaStream_t stream) {
    if (!d_mel || !d_out) { set_error("mel_windows: null pointer"); return B2A_EINVAL; }
    if (n_mels <= 0 || T < 0 || content < 0 || content > T || seek0 < 0 || stride < 0 || n_win < 0 || n_frames <= 0) {
        set_error("mel_windows: bad shape (n_mels=%d T=%lld content=%lld seek0=%lld stride=%lld n_win=%d n_frames=%d)", n_mels,
                  (long long)T, (long long)content, (long long)seek0, (long long)stride, n_win, n_frames);
        return B2A_EINVAL;
    }
    if (out_fmt != B2A_FMT_F32 && out_fmt != B2A_FMT_F16) { set_error("mel_windows: output format must be F32 or F16"); return B2A_EUNSUPPORTED; }
    const i64 rows = (i64)n_win * n_mels;
    if (rows == 0) return B2A_OK;
    if (rows > 0x7fffffffLL) { set_error("mel_windows: too many rows"); return B2A_EINVAL; }
    if (out_fmt == B2A_FMT_F16) {
        auto k = mel_windows_kernel<true>;
        B2A_LAUNCH(k, (unsigned)rows, 256, 0, stream, d_mel, n_mels, T, content, seek0, stride, n_frames, d_out);
    } else {
        auto k = mel_windows_kernel<false>;
        B2A_LAUNCH(k, (unsigned)rows, 256, 0, stream, d_mel, n_mels, T, content, seek0, stride, n_frames, d_out);
    }
    B2A_CHECK_LAUNCH("mel_windows_kernel");
    return B2A_OK;
}

}  // namespace b2a

#if defined(B2A_PROFILE) && !defined(B2A_EMU)
// profiling builds only: the event trace of logmel_tc_kernel's CTA 0 (tools/probes/logmel_tc_trace.py)
extern "C" int b2a_debug_tc_trace(unsigned long long* h_out, int n, int reset) {
    if (h_out && n > 0) cudaMemcpyFromSymbol(h_out, b2a::g_tc_trace, sizeof(unsigned long long) * (size_t)(n < 16384 ? n : 16384));
    if (reset) { unsigned long long z = 0; cudaMemcpyToSymbol(b2a::g_tc_trace, &z, sizeof(z)); }
    return 0;
}
#endif

extern "C" {

size_t b2a_log_mel_workspace_bytes(int64_t batch, int64_t n, int64_t padding) {
    if (batch <= 0 || n < 0 || padding < 0) return 0;
    return b2a::logmel_workspace_bytes(batch, n, padding);
}

int b2a_log_mel(const void* d_audio, int fmt, int64_t batch, int64_t n, int64_t row_stride, const int64_t* d_n,
                int64_t padding, int n_mels, int norm_mode, float* d_out, int64_t* d_frames_out, void* d_ws,
                size_t ws_bytes, b2a_stream_t stream) {
    return b2a::logmel_launch(d_audio, fmt, batch, n, row_stride, (const b2a::i64*)d_n, padding, n_mels, norm_mode,
                              d_out, (b2a::i64*)d_frames_out, d_ws, ws_bytes, (cudaStream_t)stream);
}

int b2a_mel_windows(const float* d_mel, int n_mels, int64_t T, int64_t content_frames, int64_t seek0, int64_t stride,
                    int n_windows, int n_frames, int out_fmt, void* d_out, b2a_stream_t stream) {
    return b2a::mel_windows_launch(d_mel, n_mels, T, content_frames, seek0, stride, n_windows, n_frames, out_fmt, d_out,
                                   (cudaStream_t)stream);
}

}  // extern "C"
