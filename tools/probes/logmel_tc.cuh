// logmel_tc.cuh — Whisper log-mel of 16-bit PCM on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// PROBE, NOT IN THE RELEASE LIBRARY.  Round 2 built this kernel to replace the CUDA-core FFT (csrc/logmel.cu) for 16-bit
// input; it is parity-green on hardware (all GPU log-mel / pipeline tests at <= 1e-4) but NOT faster: 178 us against 177 us
// for the cfg2 clip's 288 000 frames, and 40 us slower inside b2a_pipeline (profiles/r02_logmel_tc.md has the timeline
// and the reasons: the work around the MMAs - window, folds, f16 splits, the 201-bin epilogue - is 7.5 k thread instructions
// per frame against 14 k for the FFT, but one CTA per SM of lockstep phases issues at half the FFT kernel's rate, and TMEM
// (448 of 512 columns are accumulators) leaves no room to overlap the phases).  BASELINE.json's rule applies: tensor cores only
// if the DFT-as-GEMM beats the FFT on measured evidence.  It is compiled into profiling builds (-DB2A_PROFILE,
// tools/probes/build_profile_lib.py) and into the TEST-ONLY emulation library, selected there with B2A_LM_IMPL=tc.
//
// Replaces whisper.audio.log_mel_spectrogram (openai-whisper whisper/audio.py), reached from model.transcribe at
// /root/reference/app/services/audio_processor.py:1076-1080, for every 16-bit input (the pipeline's case: Whisper reads the
// 16 kHz WAV back as int16 / 32768).  The 400-point DFT is a GEMM in Whisper's own order of operations:
//
//   xw[n] = w[n] x[n]                       periodic Hann in the time domain, f32 (x in the s16 / 4 domain)
//   ge[n] = xw[n] + xw[n+200], go[n] = xw[n] - xw[n+200]                 time aliasing: even / odd bins decouple
//   ae[n] = ge[n] + ge[200-n], ao[n] = ge[n] - ge[200-n], de[n] = go[n] - go[200-n], do[n] = go[n] + go[200-n]   (n = 0..100)
//   Re X[2j]   = sum_n ae[n] cos(pi j n / 100)         Im X[2j]   = sum_n ao[n] sin(pi j n / 100)
//   Re X[2j+1] = sum_n de[n] cos(pi (2j+1) n / 200)    Im X[2j+1] = sum_n do[n] sin(pi (2j+1) n / 200)
//
// i.e. four real products [128 frames x K 112] x [K 112 x N 112] per tile instead of a 400 x 402 DFT matrix: 100 kFLOP per
// frame and plane pair, and a basis of 8 planes x 21 KB that stays RESIDENT in shared memory.  Every operand is split in
// two f16 planes (v = hi + lo, basis c = f16(c) + f16(c - f16(c))) and all four plane products accumulate into one f32
// accumulator: 112 tcgen05.mma (M128 N112 K16, kind::f16, A from TMEM) per tile.  tools/studies/logmel_tc_windowed.py
// restates the arithmetic (incl. the accumulator's truncation): max |err| 2e-5 .. 7.5e-5 against the float64 oracle on
// tones over a noise floor, where torch.stft in f32 sits at 6e-5 .. 8e-5; the gate is 1e-4.
//
// One persistent CTA per SM, coupled only by mbarriers:
//   warps 0-15   workers, all alike: lane quadrant q = warp % 4 (thread = frame 32 q + lane of the tile), part r = warp / 4.
//                 fetch    the NEXT tile's 20 720 raw samples (128 hops + 240) are 2 590 16-byte chunks, five or six per thread,
//                          loaded into registers before the current tile is converted (gathered through the kept-range table
//                          in the pipeline = fused stream compaction; chunks that touch the clip's edges - reflect pad, zero
//                          pad, a zero-filled last millisecond - are assembled sample by sample) and dropped into shared
//                          memory once every worker is done with the current tile: 130 padded rows of one hop (pitch 328 B:
//                          conflict-free 8-byte reads by thread = frame).  The tile's own chunks also go out as the trimmed PCM.
//                          (A first version staged the rows with one 1-D bulk copy each: 130 small TMA requests per tile took
//                          8 us; a second one used eight dedicated loader warps and left eight converters: profiles/r02_logmel_tc.md.)
//                 convert  part r = n-values 16 s + 4 r .. + 3 of k-step s: LDS.64 of the four segments x[n], x[n+200], x[200-n],
//                          x[400-n], window + folds in f32, split into f16 planes, tcgen05.st into a ring of four 16-column
//                          operand slots in TMEM (slot = product), one full / free mbarrier pair per slot.  k-steps run from
//                          n = 96 down to 0 and only three plane products are accumulated (hi hi, hi lo, lo hi): with operands
//                          split as floats the lo lo term is below f32 rounding, and every accumulation less is one truncation
//                          of the tensor core's f32 accumulator less (tools/studies/logmel_tc_windowed.py: 8.1e-5 -> 5.2e-5 on
//                          the worst probed signal).
//                 epilogue part r = mel range r (mel_tc_tables_gen.inc): tcgen05.ld 16 bins of each accumulator at a time, power,
//                          slaney weights as FFMA immediates, log10, coalesced stores along T, running clip maximum and
//                          per-32-frame minimum for the floor pass (mel_floor_kernel).
//   warp 16      MMA issuer (elected lane) + TMEM allocation; tcgen05.commit releases operand slots and hands the tile's
//                accumulators (4 x 112 columns) to the epilogue.
//   warp 17      scout (pipeline only): looks up the kept ranges under the next tile (32-ary search in the offset table) and
//                publishes a window of 32 of them in shared memory, two tiles ahead of the converters.
// TMEM: accumulators [0, 448), operand ring [448, 512).  Shared memory: basis 171 KB + window 2 KB + raw tile 42 KB.
#pragma once
#include <utility>

#include "../../audio_processor_b200/csrc/fir_tc_common.cuh"
#include "../../audio_processor_b200/csrc/fir_tmem.cuh"

namespace b2a {

constexpr int kTcFrames = 128;                       // frames per tile = UMMA M
constexpr int kTcKS = 7;                             // k-steps of 16 (K = 112 >= 101)
constexpr int kTcNP = 4;                             // products: even-cos, even-sin, odd-cos, odd-sin
constexpr int kTcNB = 112;                           // N of every product
constexpr int kTcLbo = 13 * 128, kTcSbo = 128;       // basis plane: [k chunk of 8][n group of 8][8 rows][16 B], 13 x 13 stored
constexpr int kTcPlaneBytes = 13 * kTcLbo;           // 21 632 (chunk 13 / group 13 alias the next chunk / plane: finite values x zero operands)
constexpr int kTcBankBytes = 2 * kTcNP * kTcPlaneBytes + kTcLbo + 128;   // + zeros behind the last plane
constexpr int kTcWinFloats = 4 * 112;                // window tables: x[n], x[n+200], x[200-n], x[400-n], n = 0..111
constexpr int kTcBlobBytes = kTcBankBytes + kTcWinFloats * 4;
constexpr int kTcRows = kTcFrames + 2;               // hops staged per tile
constexpr int kTcRowBytes = 328;                     // 160 samples + 4 of padding: 82 words, conflict-free 8-byte reads by thread = frame
constexpr int kTcRawBytes = kTcRows * kTcRowBytes;
constexpr int kTcLastRowSamples = 88;                // samples of row 129 a tile needs (x[400] of its last frame is sample 80)
constexpr int kTcSegWin = 32;                        // kept ranges cached per tile (gather)
constexpr int kTcWorkers = 16, kTcIssuer = 16, kTcScout = 17;
constexpr int kTcThreads = 18 * 32;
constexpr int kTcWorkerThreads = kTcWorkers * 32;
constexpr int kTcChunks = (kTcFrames * kHop + 240) / 8;           // 16-byte chunks of a raw tile (2 590)
constexpr int kTcChunkRounds = (kTcChunks + kTcWorkerThreads - 1) / kTcWorkerThreads;   // 6
constexpr int kTcRingCol = kTcNP * kTcNB;            // 448
constexpr unsigned kTcIdesc = (1u << 4) | ((unsigned)(kTcNB >> 3) << 17) | ((unsigned)(kTcFrames >> 4) << 24);
// barriers
constexpr int kTcBarFull = 0, kTcBarFree = 4, kTcBarAccFull = 8, kTcBarAccFree = 9, kTcBarRawFull = 10, kTcBarRawFree = 11, kTcBarBank = 12,
              kTcBarWinFull = 13, kTcBarWinFree = 15, kTcNBars = 17;
constexpr int kTcSmemBank = 0;
constexpr int kTcSmemWin = kTcBankBytes;
constexpr int kTcSmemRaw = kTcBlobBytes;
constexpr int kTcSmemBars = kTcSmemRaw + kTcRawBytes + 32;
constexpr int kTcSmemSeg = kTcSmemBars + kTcNBars * kFmBarBytes;       // 2 x { i64 src[32]; int off[33] (relative to the tile's first raw sample); int pad }
constexpr int kTcSegBytes = kTcSegWin * 8 + (kTcSegWin + 1) * 4 + 4;
constexpr int kTcSmemMisc = kTcSmemSeg + 2 * kTcSegBytes;
constexpr int kTcSmemBytes = kTcSmemMisc + 64;
static_assert(kTcBankBytes % 16 == 0 && kTcSmemRaw % 16 == 0 && kTcSmemBars % 8 == 0 && kTcSmemSeg % 8 == 0 && kTcSegBytes % 8 == 0, "alignment");
static_assert(kTcSmemBytes <= 232448, "shared-memory budget (227 KB per CTA)");

#ifndef B2A_MEL_TC_TABLES_INCLUDED
#define B2A_MEL_TC_TABLES_INCLUDED
#include "_bin/mel_tc_tables_gen.inc"
#endif

const unsigned char* get_logmel_tc_blob();           // device copy of the basis bank + window tables (b2a_host.cu)

struct LogMelTcParams {
    const int16_t* audio;   // [batch] rows
    i64 row_stride;         // samples between rows
    i64 n;                  // samples per row (capacity when d_n != nullptr)
    const i64* d_n;         // optional device-side actual length (batch == 1)
    i64 padding;
    int batch;
    float* out;             // [batch][n_mels][T]
    i64* d_frames_out;
    int* gmax_key;          // [batch] (per-clip) or [1]
    int per_clip;
    int* tile_min_key;      // [batch][groups_cap]: minimum of every 32-frame group (float_to_key)
    i64 groups_cap;
    const unsigned char* blob;
    // fused stream compaction (pipeline): see LogMelParams in logmel.cu
    const int32_t* kept_ms;
    const i64* kept_off;
    const i64* info;
    int16_t* trim_out;
    i64 n_src;
};

// ---- primitives --------------------------------------------------------------------------------------------------
#ifndef B2A_EMU
__device__ __forceinline__ void tmem_ld16f(unsigned taddr, float (&r)[16]) {
    unsigned u[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned lds_u16(saddr_t a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
    return (unsigned)v;
}
__device__ __forceinline__ float4 lds_f4(saddr_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
// (lo, hi) -> packed f16x2, round to nearest even; and back
__device__ __forceinline__ unsigned pack_f16x2(float lo, float hi) {
    unsigned d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float2 unpack_f16x2(unsigned w) {
    float2 r;
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(r.x), "=f"(r.y) : "r"(w));
    return r;
}
// 1-D bulk copy shared -> global (TMA engine), tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_store(void* dst, saddr_t src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint2 lds64(saddr_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(saddr_t a, unsigned x, unsigned y) { asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(a), "r"(x), "r"(y) : "memory"); }
__device__ __forceinline__ void tmem_st2(unsigned taddr, unsigned r0, unsigned r1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u16(saddr_t a, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v) : "memory"); }
__device__ __forceinline__ void sts_zero16(saddr_t a) { asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(0u) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx_only(saddr_t bar, unsigned bytes) {
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
#else
static inline void tmem_ld16f(unsigned taddr, float (&r)[16]) {
    const int lane0 = (int)(taddr >> 16), col0 = (int)(taddr & 0xffffu);
    for (int j = 0; j < 16; j++) r[j] = g_emu_tmem[lane0 + emu_lane()][col0 + j];
}
static inline void tmem_ld_wait() {}
static inline unsigned lds_u16(saddr_t a) { return (unsigned)*(const unsigned short*)a; }
static inline float4 lds_f4(saddr_t a) { return *(const float4*)a; }
static inline unsigned pack_f16x2(float lo, float hi) { return (unsigned)b2a_f16::f32_to_f16(lo) | ((unsigned)b2a_f16::f32_to_f16(hi) << 16); }
static inline float2 unpack_f16x2(unsigned w) { return make_float2(b2a_f16::f16_to_f32((unsigned short)(w & 0xffffu)), b2a_f16::f16_to_f32((unsigned short)(w >> 16))); }
static inline void bulk_store(void* dst, saddr_t src, unsigned bytes) { memcpy(dst, (const void*)src, bytes); }
static inline void bulk_store_commit() {}
static inline void bulk_store_wait_read() {}
static inline void bulk_store_wait_all() {}
static inline uint2 lds64(saddr_t a) { return *(const uint2*)a; }
static inline void sts64(saddr_t a, unsigned x, unsigned y) { ((unsigned*)a)[0] = x; ((unsigned*)a)[1] = y; }
static inline void tmem_st2(unsigned taddr, unsigned r0, unsigned r1) {
    const int lane0 = (int)(taddr >> 16), col0 = (int)(taddr & 0xffffu);
    g_emu_tmem[lane0 + emu_lane()][col0] = __uint_as_float(r0);
    g_emu_tmem[lane0 + emu_lane()][col0 + 1] = __uint_as_float(r1);
}
static inline void prefetch_l2(const void*) {}
static inline uint4 ldg_nc16(const void* p) { return *(const uint4*)p; }
static inline void sts_u16(saddr_t a, unsigned v) { *(unsigned short*)a = (unsigned short)v; }
static inline void sts_zero16(saddr_t a) { memset((void*)a, 0, 16); }
static inline void mbar_expect_tx_only(saddr_t bar, unsigned bytes) { unsigned* b = (unsigned*)bar; b[3] += bytes; }
#endif

// ---- epilogue: one role = one mel range, straight-line code over its bins ---------------------------------------------
struct TcEpi {
    float a0, a1;            // open filters: even / odd mel
    float lmax, lmin;        // log2 domain
    float* out;              // &out[0][t]
    size_t T;
    bool valid;
};
template <int MEL>
__device__ __forceinline__ void tc_emit(TcEpi& e, float acc) {
    const float l2 = __log2f(fmaxf(acc, 1e-10f));
    e.lmax = fmaxf(e.lmax, l2);
    e.lmin = fminf(e.lmin, l2);
    if (e.valid) e.out[(size_t)MEL * e.T] = fmaf(l2 * 0.30102999566398120f, 0.25f, 1.0f);   // (log10 + 4) / 4, Whisper's two roundings
}
template <int NM, int R, int K>
__device__ __forceinline__ void tc_epi_bin(TcEpi& e, float re, float im) {
    using M = MelTc<NM>;
    if constexpr (K >= M::bin0[R] && K <= M::bin1[R]) {
        constexpr float w0 = M::w0[R][K], w1 = M::w1[R][K];
        constexpr int e0 = M::e0[R][K], e1 = M::e1[R][K];
        if constexpr (w0 != 0.0f || w1 != 0.0f) {
            const float p = fmaf(re, re, im * im);
            if constexpr (w0 != 0.0f) e.a0 = fmaf(w0, p, e.a0);
            if constexpr (w1 != 0.0f) e.a1 = fmaf(w1, p, e.a1);
        }
        if constexpr (e0 >= 0) { tc_emit<e0>(e, e.a0); e.a0 = 0.0f; }
        if constexpr (e1 >= 0) { tc_emit<e1>(e, e.a1); e.a1 = 0.0f; }
    }
}
template <int NM, int R, int C, int... JJ>
__device__ __forceinline__ void tc_epi_chunk_bins(TcEpi& e, const float (&re_e)[16], const float (&im_e)[16], const float (&re_o)[16],
                                                  const float (&im_o)[16], std::integer_sequence<int, JJ...>) {
    ((tc_epi_bin<NM, R, 32 * C + 2 * JJ>(e, re_e[JJ], im_e[JJ]), tc_epi_bin<NM, R, 32 * C + 2 * JJ + 1>(e, re_o[JJ], im_o[JJ])), ...);
}
// chunk C = bins 32 C .. 32 C + 31 = columns 16 C .. 16 C + 15 of the four accumulators
template <int NM, int R, int C>
__device__ __forceinline__ void tc_epi_chunk(TcEpi& e, unsigned tacc) {
    using M = MelTc<NM>;
    if constexpr (32 * C + 31 >= M::bin0[R] && 32 * C <= M::bin1[R]) {
        float re_e[16], im_e[16], re_o[16], im_o[16];
        tmem_ld16f(tacc + 0 * kTcNB + 16 * C, re_e);
        tmem_ld16f(tacc + 1 * kTcNB + 16 * C, im_e);
        tmem_ld16f(tacc + 2 * kTcNB + 16 * C, re_o);
        tmem_ld16f(tacc + 3 * kTcNB + 16 * C, im_o);
        tmem_ld_wait();
        tc_epi_chunk_bins<NM, R, C>(e, re_e, im_e, re_o, im_o, std::make_integer_sequence<int, 16>{});
    }
}
template <int NM, int R>
__device__ __forceinline__ void tc_epi_role(TcEpi& e, unsigned tacc) {
    tc_epi_chunk<NM, R, 0>(e, tacc); tc_epi_chunk<NM, R, 1>(e, tacc); tc_epi_chunk<NM, R, 2>(e, tacc); tc_epi_chunk<NM, R, 3>(e, tacc);
    tc_epi_chunk<NM, R, 4>(e, tacc); tc_epi_chunk<NM, R, 5>(e, tacc); tc_epi_chunk<NM, R, 6>(e, tacc);
}

// ---- profiling builds only (-DB2A_PROFILE, tools/probes/logmel_tc_trace.py): CTA 0 records clock64 at pipeline events ----
#if defined(B2A_PROFILE) && !defined(B2A_EMU)
__device__ unsigned long long g_tc_trace[16384];      // [0] = records written; record = event << 56 | warp << 48 | clock
#define TC_TRACE(ev)                                                                                              \
    do {                                                                                                          \
        if (blockIdx.x == 0 && lane == 0) {                                                                       \
            const unsigned long long i_ = atomicAdd(&g_tc_trace[0], 1ull) + 1ull;                                  \
            if (i_ < 16384ull) g_tc_trace[i_] = ((unsigned long long)(ev) << 56) | ((unsigned long long)warp << 48) | ((unsigned long long)clock64() & 0xffffffffffffull); \
        }                                                                                                         \
    } while (0)
#else
#define TC_TRACE(ev) do { } while (0)
#endif

// waits that usually last thousands of cycles (a whole phase of the tile) back off between polls, so that the waiting warps
// do not take issue slots from the working ones
#ifndef TC_WAIT_SLEEP
#define TC_WAIT_SLEEP 0
#endif
#if TC_WAIT_SLEEP && !defined(B2A_EMU)
__device__ __forceinline__ void mbar_wait_backoff(saddr_t bar, unsigned parity) {
    unsigned ok;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        __nanosleep(TC_WAIT_SLEEP);
    }
}
#define TC_WAIT_LONG(bar, par) mbar_wait_backoff(bar, par)
#else
#define TC_WAIT_LONG(bar, par) mbar_wait(bar, par)
#endif
#define TC_WAIT(bar, par) mbar_wait(bar, par)

// ---- the kernel ------------------------------------------------------------------------------------------------------
// one 16-byte chunk that touches an edge of the padded clip (reflect at both ends, zeros past n_act, zero-filled last millisecond):
// assembled sample by sample; out of line, it runs for a handful of chunks per clip
template <bool GATHER>
__device__ __noinline__ uint4 tc_edge_chunk(const LogMelTcParams& p, const int16_t* row, i64 qc, i64 n_act, i64 ltot, int n_seg) {
    unsigned w[4] = {0u, 0u, 0u, 0u};
    for (int j = 0; j < 8; j++) {
        i64 qq = qc + j;
        if (qq < 0) qq = -qq;
        if (qq >= ltot) qq = 2 * (ltot - 1) - qq;
        int sv = 0;
        if (qq >= 0 && qq < n_act) {
            if (!GATHER) sv = row[qq];
            else if (n_seg > 0) {
                int lo = 0, hi = n_seg - 1;                       // largest k with kept_off[k] <= qq
                while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (p.kept_off[mid] <= qq) lo = mid; else hi = mid - 1; }
                const i64 si = (i64)p.kept_ms[2 * lo] * 16 + (qq - p.kept_off[lo]);
                if (si < p.n_src) sv = p.audio[si];
            }
        }
        w[j >> 1] |= ((unsigned)sv & 0xffffu) << (16 * (j & 1));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <int NM, bool GATHER>
__global__ void __launch_bounds__(kTcThreads, 1) logmel_tc_kernel(const LogMelTcParams p) {
    B2A_DYN_SMEM(smem);
    desc_origin(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const saddr_t s_base = smem_addr(smem);
    const saddr_t bars = s_base + kTcSmemBars;
    auto BAR = [&](int i) -> saddr_t { return bars + (unsigned)(kFmBarBytes * i); };
    unsigned* tmem_slot = (unsigned*)(smem + kTcSmemMisc);

    i64 n_act = p.n;
    if (p.d_n) { n_act = *p.d_n; if (n_act > p.n) n_act = p.n; if (n_act < 0) n_act = 0; }
    const i64 ltot = n_act + p.padding;
    const i64 T = ltot / kHop;
    i64 tiles = (T + kTcFrames - 1) / kTcFrames;
    if (GATHER && tiles * (kTcFrames * kHop) < n_act) tiles++;        // a last partial hop still has trimmed samples to write
    const i64 n_work = tiles * p.batch;
    if (blockIdx.x == 0 && tid == 0 && p.d_frames_out) *p.d_frames_out = T;
    const int n_seg = GATHER ? (int)p.info[B2A_INFO_N_KEPT] : 0;

    if (tid == 0) {
        mbar_init(BAR(kTcBarFull), kTcWorkers / 2);                     // one group of eight warps per k-step
        mbar_init(BAR(kTcBarFree), 1);                                  // two "ring free" barriers, one per k-step parity: a converter
        mbar_init(BAR(kTcBarFree + 1), 1);                              // skips every other k-step, so it must never be two phases behind one barrier
        mbar_init(BAR(kTcBarAccFull), 1);
        mbar_init(BAR(kTcBarAccFree), kTcWorkers);
        mbar_init(BAR(kTcBarRawFull), kTcWorkers);
        mbar_init(BAR(kTcBarRawFree), kTcWorkers);
        mbar_init(BAR(kTcBarBank), 1);
        for (int i = 0; i < 2; i++) { mbar_init(BAR(kTcBarWinFull + i), 1); mbar_init(BAR(kTcBarWinFree + i), kTcWorkers); }
        mbar_fence_init();
    }
    if (warp == kTcIssuer) tmem_alloc(smem_addr(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = *tmem_slot;

    const bool one_clip = p.batch == 1;
    auto split_work = [&](i64 work, int& b, i64& tile) {
        if (one_clip) { b = 0; tile = work; }
        else { b = (int)(work / tiles); tile = work - (i64)b * tiles; }
    };

    // registers: the launch gives every thread 96 (18 warps: five on two of the four sub-partitions); the issuer / scout
    // warpgroup hands most of its share back and the workers take 112 (4 x 32 x 112 + 32 x 40 <= 16 384 per sub-partition)
    // (setmaxnreg rebalancing was tried: ptxas spilled more with it than without, see profiles/r02_logmel_tc.md)
    if (warp == kTcIssuer) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            mbar_expect_tx(BAR(kTcBarBank), (unsigned)kTcBlobBytes);
            for (int off = 0; off < kTcBlobBytes; off += 32768) {
                const int nb = kTcBlobBytes - off < 32768 ? kTcBlobBytes - off : 32768;
                bulk_load(s_base + kTcSmemBank + off, p.blob + off, (unsigned)nb, BAR(kTcBarBank));
            }
        }
        __syncwarp();
        mbar_wait(BAR(kTcBarBank), 0);
        const unsigned b_hi = (unsigned)(kTcSbo >> 4) | (1u << 14);
        const unsigned b_lo0 = desc_start(s_base + kTcSmemBank) | ((unsigned)(kTcLbo >> 4) << 16);
        unsigned g = 0, it = 0;
        for (i64 work = blockIdx.x; work < n_work; work += gridDim.x, it++) {
            if (it > 0) { TC_WAIT_LONG(BAR(kTcBarAccFree), (it - 1) & 1u); tc_fence_after(); }
            TC_TRACE(10);
#pragma unroll 1
            for (int s = kTcKS - 1; s >= 0; s--, g++) {                // n = 96 .. 111 first (see the header: accumulation order)
                mbar_wait(BAR(kTcBarFull), g & 1u);
                tc_fence_after();
                TC_TRACE(9);
#pragma unroll
                for (int pr = 0; pr < kTcNP; pr++) {
                    const unsigned d = tbase + (unsigned)(pr * kTcNB);
                    const unsigned ah = tbase + (unsigned)(kTcRingCol + 16 * pr), al = ah + 8u;
                    const unsigned bh = b_lo0 + (unsigned)(((2 * pr) * kTcPlaneBytes + 2 * s * kTcLbo) >> 4);
                    const unsigned bl = b_lo0 + (unsigned)(((2 * pr + 1) * kTcPlaneBytes + 2 * s * kTcLbo) >> 4);
                    umma_ts_warp(d, ah, bh, b_hi, kTcIdesc, s < kTcKS - 1 ? 1u : 0u);
                    umma_ts_warp(d, ah, bl, b_hi, kTcIdesc, 1u);
                    umma_ts_warp(d, al, bh, b_hi, kTcIdesc, 1u);
                }
                umma_commit_warp(BAR(kTcBarFree + (g & 1u)));
            }
            umma_commit_warp(BAR(kTcBarAccFull));
            TC_TRACE(11);
        }
    } else if (warp == kTcScout) {
        // ---------------- scout: kept ranges under the tiles this CTA will fetch, two fetches ahead ----------------
        if (GATHER && n_seg > 0) {
            unsigned j = 0;
            for (i64 work = blockIdx.x; work < n_work; work += gridDim.x, j++) {
                int b;
                i64 tile;
                split_work(work, b, tile);
                const i64 q0 = tile * (kTcFrames * kHop) - 200;
                const i64 qq = q0 > 0 ? q0 : 0;
                int lo = 0, cnt = n_seg;                               // the range holding qq lies in [lo, lo + cnt); kept_off[lo] <= qq
                while (cnt > 1) {                                      // 32-ary search: three rounds of parallel probes for 8 192 ranges
                    const int stride = (cnt + 31) / 32;
                    const int idx = lo + lane * stride;
                    const bool le = idx < lo + cnt && p.kept_off[idx] <= qq;
                    const int c = __popc(__ballot_sync(0xffffffffu, le));
                    const int nlo = lo + (c - 1) * stride;
                    cnt = nlo + stride > lo + cnt ? lo + cnt - nlo : stride;
                    lo = nlo;
                }
                const int k = lo + lane;
                const i64 off = k <= n_seg ? p.kept_off[k] - q0 : (i64)0x7fffffff;
                const i64 off32 = lo + kTcSegWin <= n_seg ? p.kept_off[lo + kTcSegWin] - q0 : (i64)0x7fffffff;
                const i64 src = k < n_seg ? (i64)p.kept_ms[2 * k] * 16 : 0;
                const unsigned buf = j & 1u, use = j >> 1;
                if (use > 0) mbar_wait(BAR(kTcBarWinFree + buf), (use - 1) & 1u);
                i64* w_src = (i64*)(smem + kTcSmemSeg + buf * kTcSegBytes);
                int* w_off = (int*)(w_src + kTcSegWin);
                w_src[lane] = src;
                w_off[lane] = off > 0x7fffffff ? 0x7fffffff : (int)off;
                if (lane == 0) w_off[kTcSegWin] = off32 > 0x7fffffff ? 0x7fffffff : (int)off32;
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(kTcBarWinFull + buf));
            }
        }
    } else {
        // ---------------- workers ----------------
        const int q = warp & 3, part = warp >> 2;                       // part: mel range of the epilogue
        const int grp = part & 1, half = part >> 1;                     // conversion: k-step group and n-half (see the k-step loop)
        const int f = 32 * q + lane;                                    // frame (row) of the tile
        const int wt = warp * 32 + lane;                                // worker thread 0..511
        const unsigned tlane = tbase + ((unsigned)(32 * q) << 16);
        const saddr_t rowp = s_base + kTcSmemRaw + (unsigned)(f * kTcRowBytes);
        const saddr_t winp = s_base + kTcSmemWin;
        float run_max = -3.0e38f;
        unsigned g = 0, it = 0, nfetch = 0;
        mbar_wait(BAR(kTcBarBank), 0);                                  // window tables
        // what the fetch of the next tile leaves behind for the store: per chunk the source (in units of 8 samples from
        // p.audio; the line is already on its way into L2) or a mark that it is an edge chunk
        unsigned pre_off[kTcChunkRounds];
        unsigned src_mask = 0, edge_mask = 0, own_mask = 0;              // per round: has a global source / edge chunk / also trimmed PCM (gather); neither source nor edge = zeros
        i64 pre_q0 = 0;
        const int16_t* pre_row = p.audio;

        // iteration -1 only primes the pipeline (fetch + store of the CTA's first tile)
#pragma unroll 1
        for (i64 work = (i64)blockIdx.x - (i64)gridDim.x; work < n_work; work += gridDim.x) {
            const bool live = work >= (i64)blockIdx.x;
            const i64 next = work + gridDim.x;
            const bool have_next = next < n_work;
            int b = 0;
            i64 tile = 0;
            if (live) split_work(work, b, tile);

            // ---- fetch: where the next tile's chunks come from; their lines start moving into L2 while this tile is converted ----
            TC_TRACE(1);
            if (have_next) {
                int bn;
                i64 tile_n;
                split_work(next, bn, tile_n);
                pre_row = p.audio + (size_t)bn * (size_t)p.row_stride;
                const i64 q0 = tile_n * (kTcFrames * kHop) - 200;        // padded-domain index of raw sample 0
                pre_q0 = q0;
                src_mask = edge_mask = own_mask = 0;
                const unsigned buf = nfetch & 1u;
                const i64* w_src = (const i64*)(smem + kTcSmemSeg + buf * kTcSegBytes);
                const int* w_off = (const int*)(w_src + kTcSegWin);
                if (GATHER && n_seg > 0) mbar_wait(BAR(kTcBarWinFull + buf), (nfetch >> 1) & 1u);
#pragma unroll
                for (int r = 0; r < kTcChunkRounds; r++) {
                    const int c = wt + kTcWorkerThreads * r;
                    pre_off[r] = 0u;
                    if (c < kTcChunks) {
                        const int qr = 8 * c;
                        const i64 qc = q0 + qr;
                        const int16_t* src = nullptr;
                        if (qc >= 0 && qc + 8 <= n_act) {
                            if (!GATHER) src = pre_row + qc;
                            else if (n_seg > 0 && qr < w_off[kTcSegWin]) {
                                int e = 0;                               // largest e with w_off[e] <= qr
#pragma unroll
                                for (int st = 16; st > 0; st >>= 1) if (w_off[e + st] <= qr) e += st;
                                const i64 si = w_src[e] + (qr - w_off[e]);
                                if (si + 8 <= p.n_src) src = p.audio + si;
                            }
                            if ((((uintptr_t)src) & 15) != 0 || ((src - p.audio) & 7) != 0) src = nullptr;
                            if (GATHER && c >= 25 && c < 25 + kTcFrames * kHop / 8) own_mask |= 1u << r;
                        }
                        if (src) {
                            prefetch_l2(src);
                            pre_off[r] = (unsigned)((src - p.audio) >> 3);
                            src_mask |= 1u << r;
                        } else if (!(qc >= n_act && qc + 8 <= ltot) && qc < ltot + 200) {
                            edge_mask |= 1u << r;                        // (else: zeros - right padding, or behind the last frame's window)
                        }
                    }
                }
                if (GATHER && n_seg > 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(kTcBarWinFree + buf));
                }
                nfetch++;
            }

            // ---- convert this tile ----
            TC_TRACE(2);
            if (live) {
                TC_WAIT_LONG(BAR(kTcBarRawFull), it & 1u);
                TC_TRACE(3);
                // Two groups of eight warps take the k-steps in turns (group = (warp / 4) % 2 converts the k-steps whose index g in
                // issue order has its parity): while one group's operands sit in the ring and the tensor core consumes them, the
                // other group is already computing the next k-step in registers, so the 4-slot ring no longer forces all sixteen
                // warps through every k-step in lockstep (profiles/r02_logmel_tc.md: 2.0 k cycles per k-step before).
                // Within a group a warp covers eight n-values: half = warp / 8 selects n = 16 s + 8 half .. + 7.
#pragma unroll 1
                for (int s = kTcKS - 1; s >= 0; s--, g++) {
                    if ((g & 1u) != (unsigned)grp) continue;            // the other group's k-step
#ifndef TC_UNROLL_H
#define TC_UNROLL_H 0
#endif
#if TC_UNROLL_H
#pragma unroll
#else
#pragma unroll 1
#endif
                    for (int h = 0; h < 2; h++) {
                        const int n0 = 16 * s + 8 * half + 4 * h;
                        const int a1 = 196 - n0, b1 = 200 - n0, a2 = 396 - n0, b2 = 400 - n0;
                        const uint2 f1 = lds64(rowp + (unsigned)(2 * n0));
                        const uint2 f2 = lds64(rowp + (unsigned)(kTcRowBytes + 80 + 2 * n0));
                        const uint2 r1 = lds64(rowp + (unsigned)(a1 >= 160 ? kTcRowBytes + 2 * (a1 - 160) : 2 * a1));
                        const unsigned r1x = lds_u16(rowp + (unsigned)(b1 >= 160 ? kTcRowBytes + 2 * (b1 - 160) : 2 * b1));
                        const uint2 r2 = lds64(rowp + (unsigned)(a2 >= 320 ? 2 * kTcRowBytes + 2 * (a2 - 320) : kTcRowBytes + 2 * (a2 - 160)));
                        const unsigned r2x = lds_u16(rowp + (unsigned)(b2 >= 320 ? 2 * kTcRowBytes + 2 * (b2 - 320) : kTcRowBytes + 2 * (b2 - 160)));
                        const float4 w1 = lds_f4(winp + (unsigned)(4 * (0 * 112 + n0)));
                        const float4 w2 = lds_f4(winp + (unsigned)(4 * (1 * 112 + n0)));
                        const float4 w3 = lds_f4(winp + (unsigned)(4 * (2 * 112 + n0)));
                        const float4 w4 = lds_f4(winp + (unsigned)(4 * (3 * 112 + n0)));
                        auto lo16 = [](unsigned v) -> float { return (float)(short)(v & 0xffffu); };
                        auto hi16 = [](unsigned v) -> float { return (float)((int)v >> 16); };
                        const float x1[4] = {lo16(f1.x), hi16(f1.x), lo16(f1.y), hi16(f1.y)};
                        const float x2[4] = {lo16(f2.x), hi16(f2.x), lo16(f2.y), hi16(f2.y)};
                        const float y1[4] = {(float)(short)r1x, hi16(r1.y), lo16(r1.y), hi16(r1.x)};     // x[200 - n0 - i]
                        const float y2[4] = {(float)(short)r2x, hi16(r2.y), lo16(r2.y), hi16(r2.x)};     // x[400 - n0 - i]
                        const float wa[4] = {w1.x, w1.y, w1.z, w1.w}, wb[4] = {w2.x, w2.y, w2.z, w2.w};
                        const float wc[4] = {w3.x, w3.y, w3.z, w3.w}, wd[4] = {w4.x, w4.y, w4.z, w4.w};
                        float v[kTcNP][4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const float t1 = wa[e] * x1[e], u1 = wc[e] * y1[e];
                            const float ge = fmaf(wb[e], x2[e], t1), go = fmaf(-wb[e], x2[e], t1);
                            const float he = fmaf(wd[e], y2[e], u1), ho = fmaf(-wd[e], y2[e], u1);
                            v[0][e] = ge + he; v[1][e] = ge - he; v[2][e] = go - ho; v[3][e] = go + ho;
                        }
                        unsigned hw[kTcNP][2], lw[kTcNP][2];
#pragma unroll
                        for (int pr = 0; pr < kTcNP; pr++)
#pragma unroll
                            for (int c = 0; c < 2; c++) {
                                hw[pr][c] = pack_f16x2(v[pr][2 * c], v[pr][2 * c + 1]);
                                const float2 hf = unpack_f16x2(hw[pr][c]);
                                lw[pr][c] = pack_f16x2(v[pr][2 * c] - hf.x, v[pr][2 * c + 1] - hf.y);
                            }
                        // the ring slots (one k-step: four products) are free once the tensor core has consumed the previous k-step
                        if (h == 0 && g > 0) { TC_WAIT(BAR(kTcBarFree + ((g - 1) & 1u)), ((g - 1) >> 1) & 1u); tc_fence_after(); }
#pragma unroll
                        for (int pr = 0; pr < kTcNP; pr++) {
                            const unsigned ts = tlane + (unsigned)(kTcRingCol + 16 * pr + 4 * half + 2 * h);
                            tmem_st2(ts, hw[pr][0], hw[pr][1]);
                            tmem_st2(ts + 8u, lw[pr][0], lw[pr][1]);
                        }
                    }
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(BAR(kTcBarFull));
                        if (s <= 1) mbar_arrive(BAR(kTcBarRawFree));     // this warp's last k-step of the tile (s = 0, or s = 1 for the other group)
                    }
                    TC_TRACE(4);
                }
            }

            // ---- store: the next tile's chunks (L2 hits by now) into shared memory once every worker is done with this tile ----
            if (have_next) {
                uint4 v[kTcChunkRounds];
#pragma unroll
                for (int r = 0; r < kTcChunkRounds; r++) {                     // all loads first (L2 hits), no call in between
                    v[r] = make_uint4(0u, 0u, 0u, 0u);
                    if ((src_mask >> r) & 1u) v[r] = ldg_nc16(p.audio + ((size_t)pre_off[r] << 3));
                }
                if (edge_mask) {
#pragma unroll
                    for (int r = 0; r < kTcChunkRounds; r++)
                        if ((edge_mask >> r) & 1u) v[r] = tc_edge_chunk<GATHER>(p, pre_row, pre_q0 + 8 * (wt + kTcWorkerThreads * r), n_act, ltot, n_seg);
                }
                TC_TRACE(5);
                if (live) TC_WAIT_LONG(BAR(kTcBarRawFree), it & 1u);
                TC_TRACE(6);
#pragma unroll
                for (int r = 0; r < kTcChunkRounds; r++) {
                    const int c = wt + kTcWorkerThreads * r;
                    if (c < kTcChunks) {
                        const saddr_t dst = s_base + kTcSmemRaw + (unsigned)((c / 20) * kTcRowBytes + (c % 20) * 16);
                        sts64(dst, v[r].x, v[r].y);
                        sts64(dst + 8u, v[r].z, v[r].w);
                        if (GATHER && ((own_mask >> r) & 1u)) *(uint4*)(p.trim_out + pre_q0 + 8 * c) = v[r];   // the tile's own samples = trimmed PCM
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(kTcBarRawFull));
            }

            // ---- epilogue of this tile ----
            TC_TRACE(7);
            if (live) {
                TC_WAIT_LONG(BAR(kTcBarAccFull), it & 1u);
                tc_fence_after();
                TC_TRACE(8);
                const i64 t = tile * kTcFrames + f;
                TcEpi e;
                e.a0 = 0.0f; e.a1 = 0.0f; e.lmax = -3.0e38f; e.lmin = 3.0e38f;
                e.valid = t < T;
                e.T = (size_t)T;
                e.out = p.out + (size_t)b * (size_t)NM * (size_t)T + (e.valid ? t : 0);
                if (part == 0) tc_epi_role<NM, 0>(e, tlane);
                else if (part == 1) tc_epi_role<NM, 1>(e, tlane);
                else if (part == 2) tc_epi_role<NM, 2>(e, tlane);
                else tc_epi_role<NM, 3>(e, tlane);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(kTcBarAccFree));
                TC_TRACE(12);
                // clip maximum and the minimum of this 32-frame group (lets the floor pass skip groups above the floor)
                const float gmin = warp_reduce_min_f(e.valid ? e.lmin : 3.0e38f);
                if (e.valid) run_max = fmaxf(run_max, e.lmax * 0.30102999566398120f);
                if (lane == 0 && tile * kTcFrames + 32 * q < T)
                    atomicMin(p.tile_min_key + (size_t)b * (size_t)p.groups_cap + (size_t)(tile * 4 + q), float_to_key(fmaf(gmin * 0.30102999566398120f, 0.25f, 1.0f)));
                if (p.per_clip) {
                    const float bm = warp_reduce_max_f(run_max);
                    if (lane == 0 && bm > -1.0e38f) atomicMax(p.gmax_key + b, float_to_key(bm));
                    run_max = -3.0e38f;
                }
                it++;
            }
        }
        if (!p.per_clip) {
            const float bm = warp_reduce_max_f(run_max);
            if (lane == 0 && bm > -1.0e38f) atomicMax(p.gmax_key, float_to_key(bm));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kTcIssuer) tmem_dealloc(tbase, 512);
}

}  // namespace b2a
