// logmel_tc.cuh — Whisper log-mel of 16-bit PCM on the 5th-generation tensor cores (tcgen05 + TMEM + bulk TMA), sm_100a.
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
// One persistent CTA per SM, 17 warps coupled only by mbarriers:
//   warps 8-15   loaders (idle while a tile is converted): fetch the NEXT tile's 20 720 raw samples (128 hops + 240) as 2 590
//                16-byte chunks into registers while the current tile is converted, and drop them into shared memory the
//                moment the converters release it: 130 padded rows of one hop (pitch 336 B: conflict-free 16-byte reads by
//                thread = frame).  In the pipeline the chunks are gathered through the kept-range table (fused stream
//                compaction) and the tile's own chunks go straight back out as the trimmed PCM; chunks that touch the clip's
//                edges (reflect pad, zero pad, a zero-filled last millisecond) are assembled sample by sample.
//                (A first version staged the rows with one 1-D bulk copy each: 130 small TMA requests per tile took 8 us,
//                profiles/r02_logmel_tc.md.)
//   warps 0-7    converters, thread = frame: role 0 / 1 = lower / upper 8 of a k-step's 16 n-values; LDS.128 of the four
//                segments x[n], x[n+200], x[200-n], x[400-n], window + folds in f32, split into f16 planes, tcgen05.st into a
//                ring of four 16-column operand slots in TMEM (slot = product), one full / free mbarrier pair per slot.
//   warp 16      MMA issuer (elected lane) + TMEM allocation; tcgen05.commit releases operand slots and hands the tile's
//                accumulators (4 x 112 columns) to the epilogue.
//   warps 0-15   epilogue, thread = frame, four roles per lane quadrant = four mel ranges (mel_tc_tables_gen.inc): tcgen05.ld
//                16 bins of each accumulator at a time, power, slaney weights as FFMA immediates, log10, coalesced stores
//                along T, running clip maximum and per-32-frame minimum for the floor pass (mel_floor_kernel).
// TMEM: accumulators [0, 448), operand ring [448, 512).  Shared memory: basis 171 KB + window 2 KB + raw tile 43 KB.
#pragma once
#include <utility>

#include "fir_tc_common.cuh"
#include "fir_tmem.cuh"

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
constexpr int kTcRowBytes = 336;                     // 160 samples + 8 of padding
constexpr int kTcRawBytes = kTcRows * kTcRowBytes;
constexpr int kTcLastRowSamples = 88;                // samples of row 129 a tile needs (x[400] of its last frame is sample 80)
constexpr int kTcSegWin = 32;                        // kept ranges cached per tile by every gathering loader warp
constexpr int kTcWorkers = 16, kTcIssuer = 16;
constexpr int kTcThreads = 17 * 32;
constexpr int kTcLoaders = 8 * 32;                   // loader threads (warps 8-15)
constexpr int kTcChunks = (kTcFrames * kHop + 240) / 8;           // 16-byte chunks of a raw tile (2 590)
constexpr int kTcChunkRounds = (kTcChunks + kTcLoaders - 1) / kTcLoaders;   // 11
constexpr int kTcRingCol = kTcNP * kTcNB;            // 448
constexpr unsigned kTcIdesc = (1u << 4) | ((unsigned)(kTcNB >> 3) << 17) | ((unsigned)(kTcFrames >> 4) << 24);
// barriers
constexpr int kTcBarFull = 0, kTcBarFree = 4, kTcBarAccFull = 8, kTcBarAccFree = 9, kTcBarRawFull = 10, kTcBarRawFree = 11, kTcBarBank = 12, kTcNBars = 13;
constexpr int kTcSmemBank = 0;
constexpr int kTcSmemWin = kTcBankBytes;
constexpr int kTcSmemRaw = kTcBlobBytes;
constexpr int kTcSmemBars = kTcSmemRaw + kTcRawBytes + 32;
constexpr int kTcSmemSeg = kTcSmemBars + kTcNBars * kFmBarBytes;       // [33] kept_off window, [32] source sample of each range
constexpr int kTcSmemMisc = kTcSmemSeg + 8 * (2 * kTcSegWin + 1) * 8;     // one window per loader warp
constexpr int kTcSmemBytes = kTcSmemMisc + 64;
static_assert(kTcBankBytes % 16 == 0 && kTcSmemRaw % 16 == 0 && kTcSmemBars % 8 == 0 && kTcSmemSeg % 8 == 0, "alignment");
static_assert(kTcSmemBytes <= 232448, "shared-memory budget (227 KB per CTA)");

#ifndef B2A_MEL_TC_TABLES_INCLUDED
#define B2A_MEL_TC_TABLES_INCLUDED
#include "mel_tc_tables_gen.inc"
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

// ---- the kernel ------------------------------------------------------------------------------------------------------
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

    if (tid == 0) {
        for (int i = 0; i < kTcNP; i++) { mbar_init(BAR(kTcBarFull + i), 8); mbar_init(BAR(kTcBarFree + i), 1); }
        mbar_init(BAR(kTcBarAccFull), 1);
        mbar_init(BAR(kTcBarAccFree), kTcWorkers);
        mbar_init(BAR(kTcBarRawFull), 8);
        mbar_init(BAR(kTcBarRawFree), 8);
        mbar_init(BAR(kTcBarBank), 1);
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
            if (it > 0) { mbar_wait(BAR(kTcBarAccFree), (it - 1) & 1u); tc_fence_after(); }
#pragma unroll 1
            for (int s = 0; s < kTcKS; s++, g++) {
#pragma unroll
                for (int pr = 0; pr < kTcNP; pr++) {
                    mbar_wait(BAR(kTcBarFull + pr), g & 1u);
                    tc_fence_after();
                    const unsigned d = tbase + (unsigned)(pr * kTcNB);
                    const unsigned ah = tbase + (unsigned)(kTcRingCol + 16 * pr), al = ah + 8u;
                    const unsigned bh = b_lo0 + (unsigned)(((2 * pr) * kTcPlaneBytes + 2 * s * kTcLbo) >> 4);
                    const unsigned bl = b_lo0 + (unsigned)(((2 * pr + 1) * kTcPlaneBytes + 2 * s * kTcLbo) >> 4);
                    umma_ts_warp(d, ah, bh, b_hi, kTcIdesc, s > 0 ? 1u : 0u);
                    umma_ts_warp(d, ah, bl, b_hi, kTcIdesc, 1u);
                    umma_ts_warp(d, al, bh, b_hi, kTcIdesc, 1u);
                    umma_ts_warp(d, al, bl, b_hi, kTcIdesc, 1u);
                    umma_commit_warp(BAR(kTcBarFree + pr));
                }
            }
            umma_commit_warp(BAR(kTcBarAccFull));
        }
    } else {
        // ---------------- workers: converters (roles 0, 1), loaders (roles 2, 3), epilogue (all four roles) ----------------
        const int q = warp & 3, role = warp >> 2;
        const int f = 32 * q + lane;                                    // frame (row) of the tile
        const unsigned tlane = tbase + ((unsigned)(32 * q) << 16);
        const saddr_t rowp = s_base + kTcSmemRaw + (unsigned)(f * kTcRowBytes);
        const saddr_t winp = s_base + kTcSmemWin;
        float run_max = -3.0e38f;
        unsigned g = 0, it = 0;
        if (role < 2) mbar_wait(BAR(kTcBarBank), 0);                    // window tables

        // ---- loader state (roles 2, 3): the next tile's chunks travel global -> registers -> shared memory ----
        const int lt = (warp - 8) * 32 + lane;                          // loader thread 0..255
        i64* s_off = (i64*)(smem + kTcSmemSeg) + (warp >= 8 ? warp - 8 : 0) * (2 * kTcSegWin + 1);   // this warp's kept_off[sg0 .. sg0 + 32]
        i64* s_src = s_off + kTcSegWin + 1;                              // 16 * kept_ms[2 (sg0 + l)]
        const int n_seg = GATHER ? (int)p.info[B2A_INFO_N_KEPT] : 0;
        uint4 pre[kTcChunkRounds];
        auto seg_of_global = [&](i64 qq) -> int {                        // largest k with kept_off[k] <= qq (n_seg > 0), per lane
            int lo = 0, hi = n_seg - 1;
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (p.kept_off[mid] <= qq) lo = mid; else hi = mid - 1; }
            return lo;
        };
        // source sample of trimmed index qq (0 <= qq < n_act); -1: zero (pydub zero-fills a rounded-up last millisecond)
        auto src_index = [&](i64 qq) -> i64 {
            if (!GATHER) return qq;
            if (n_seg <= 0) return -1;
            const int sg = seg_of_global(qq);
            const i64 si = (i64)p.kept_ms[2 * sg] * 16 + (qq - p.kept_off[sg]);
            return si < p.n_src ? si : -1;
        };
        auto fetch_tile = [&](i64 work_n) {
            int bn;
            i64 tile_n;
            split_work(work_n, bn, tile_n);
            const int16_t* row = p.audio + (size_t)bn * (size_t)p.row_stride;
            const i64 q0 = tile_n * (kTcFrames * kHop) - 200;            // padded-domain index of raw sample 0
            if (GATHER && n_seg > 0) {
                // kept range under the tile's first sample: 32-ary search (three rounds of parallel probes for 8 192 ranges),
                // then the window of the next 32 ranges into this warp's shared-memory slots
                const i64 qq = q0 > 0 ? q0 : 0;
                int lo = 0, cnt = n_seg;                                   // the answer lies in [lo, lo + cnt); kept_off[lo] <= qq
                while (cnt > 1) {
                    const int stride = (cnt + 31) / 32;
                    const int idx = lo + lane * stride;
                    const bool le = idx < lo + cnt && p.kept_off[idx] <= qq;
                    const int c = __popc(__ballot_sync(0xffffffffu, le));   // probes are monotone; lane 0 always holds
                    const int nlo = lo + (c - 1) * stride;
                    cnt = nlo + stride > lo + cnt ? lo + cnt - nlo : stride;
                    lo = nlo;
                }
                const int k = lo + lane;
                __syncwarp();
                s_off[lane] = k <= n_seg ? p.kept_off[k] : ((i64)1 << 62);
                s_src[lane] = k < n_seg ? (i64)p.kept_ms[2 * k] * 16 : 0;
                if (lane == 0) s_off[kTcSegWin] = lo + kTcSegWin <= n_seg ? p.kept_off[lo + kTcSegWin] : ((i64)1 << 62);
                __syncwarp();
            }
#pragma unroll
            for (int r = 0; r < kTcChunkRounds; r++) {
                const int c = lt + kTcLoaders * r;
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (c < kTcChunks) {
                    const i64 qc = q0 + 8 * (i64)c;
                    const int16_t* src = nullptr;
                    if (qc >= 0 && qc + 8 <= n_act) {
                        if (!GATHER) src = row + qc;
                        else if (n_seg > 0) {
                            if (qc < s_off[kTcSegWin]) {
                                int e = 0;                               // largest e with s_off[e] <= qc
#pragma unroll
                                for (int st = 16; st > 0; st >>= 1) if (s_off[e + st] <= qc) e += st;
                                src = p.audio + s_src[e] + (qc - s_off[e]);
                            } else {
                                const int sg = seg_of_global(qc);
                                src = p.audio + (i64)p.kept_ms[2 * sg] * 16 + (qc - p.kept_off[sg]);
                            }
                            if (src + 8 > p.audio + p.n_src) src = nullptr;   // zero-filled last millisecond: sample by sample
                        }
                        if ((((uintptr_t)src) & 15) != 0) src = nullptr;
                    }
                    if (src) {
                        v = *(const uint4*)src;
                    } else if (!(qc >= n_act && qc + 8 <= ltot) && qc < ltot + 200) {
                        // edge chunk: reflect at both ends of the padded clip, zeros past n_act
                        unsigned w[4] = {0u, 0u, 0u, 0u};
                        for (int j = 0; j < 8; j++) {
                            i64 qq = qc + j;
                            if (qq < 0) qq = -qq;
                            if (qq >= ltot) qq = 2 * (ltot - 1) - qq;
                            int sv = 0;
                            if (qq >= 0 && qq < n_act) { const i64 si = src_index(qq); if (si >= 0) sv = (GATHER ? p.audio : row)[si]; }
                            w[j >> 1] |= ((unsigned)sv & 0xffffu) << (16 * (j & 1));
                        }
                        v = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    // the tile's own 20 480 trimmed samples (raw samples 200 .. 20 679) are the trimmed PCM
                    if (GATHER && c >= 25 && c < 25 + kTcFrames * kHop / 8 && qc + 8 <= n_act) *(uint4*)(p.trim_out + qc) = v;
                }
                pre[r] = v;
            }
        };
        auto store_tile = [&]() {
#pragma unroll
            for (int r = 0; r < kTcChunkRounds; r++) {
                const int c = lt + kTcLoaders * r;
                if (c < kTcChunks) *(uint4*)(smem + kTcSmemRaw + (c / 20) * kTcRowBytes + (c % 20) * 16) = pre[r];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(kTcBarRawFull));
        };
        if (role >= 2 && (i64)blockIdx.x < n_work) { fetch_tile(blockIdx.x); store_tile(); }

        for (i64 work = blockIdx.x; work < n_work; work += gridDim.x, it++) {
            int b;
            i64 tile;
            split_work(work, b, tile);
            if (role >= 2) {
                // next tile: loads in flight while this one is converted, into shared memory once the converters are done with it
                if (work + gridDim.x < n_work) {
                    fetch_tile(work + gridDim.x);
                    mbar_wait(BAR(kTcBarRawFree), it & 1u);
                    store_tile();
                }
            } else {
                mbar_wait(BAR(kTcBarRawFull), it & 1u);
#pragma unroll 1
                for (int s = 0; s < kTcKS; s++, g++) {
                    const int n0 = 16 * s + 8 * role;
                    const int a1 = 192 - n0, b1 = 200 - n0, a2 = 392 - n0, b2 = 400 - n0;
                    const uint4 f1 = lds128(rowp + (unsigned)(2 * n0));
                    const uint4 f2 = lds128(rowp + (unsigned)(kTcRowBytes + 80 + 2 * n0));
                    const uint4 r1 = lds128(rowp + (unsigned)(a1 >= 160 ? kTcRowBytes + 2 * (a1 - 160) : 2 * a1));
                    const unsigned r1x = lds_u16(rowp + (unsigned)(b1 >= 160 ? kTcRowBytes + 2 * (b1 - 160) : 2 * b1));
                    const uint4 r2 = lds128(rowp + (unsigned)(a2 >= 320 ? 2 * kTcRowBytes + 2 * (a2 - 320) : kTcRowBytes + 2 * (a2 - 160)));
                    const unsigned r2x = lds_u16(rowp + (unsigned)(b2 >= 320 ? 2 * kTcRowBytes + 2 * (b2 - 320) : kTcRowBytes + 2 * (b2 - 160)));
                    if (s == kTcKS - 1) {                                // last read of the raw tile by this warp
                        __syncwarp();
                        if (lane == 0) mbar_arrive(BAR(kTcBarRawFree));
                    }
                    const unsigned fw1[4] = {f1.x, f1.y, f1.z, f1.w}, fw2[4] = {f2.x, f2.y, f2.z, f2.w};
                    const unsigned rw1[4] = {r1.x, r1.y, r1.z, r1.w}, rw2[4] = {r2.x, r2.y, r2.z, r2.w};
                    auto s16_at = [](const unsigned (&w)[4], int i) -> float {           // sample i of an 8-sample quad
                        const unsigned v = w[i >> 1];
                        return (i & 1) ? (float)((int)v >> 16) : (float)(short)(v & 0xffffu);
                    };
                    unsigned hw[kTcNP][4], lw[kTcNP][4];
#pragma unroll
                    for (int h = 0; h < 2; h++) {                                         // two halves of four n-values
                        const float4 w1 = lds_f4(winp + (unsigned)(4 * (0 * 112 + n0 + 4 * h)));
                        const float4 w2 = lds_f4(winp + (unsigned)(4 * (1 * 112 + n0 + 4 * h)));
                        const float4 w3 = lds_f4(winp + (unsigned)(4 * (2 * 112 + n0 + 4 * h)));
                        const float4 w4 = lds_f4(winp + (unsigned)(4 * (3 * 112 + n0 + 4 * h)));
                        const float wa[4] = {w1.x, w1.y, w1.z, w1.w}, wb[4] = {w2.x, w2.y, w2.z, w2.w};
                        const float wc[4] = {w3.x, w3.y, w3.z, w3.w}, wd[4] = {w4.x, w4.y, w4.z, w4.w};
                        float v[kTcNP][4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int i = 4 * h + e;
                            const float x1 = s16_at(fw1, i), x2 = s16_at(fw2, i);
                            const float y1 = i == 0 ? (float)(short)r1x : s16_at(rw1, 8 - i);
                            const float y2 = i == 0 ? (float)(short)r2x : s16_at(rw2, 8 - i);
                            const float t1 = wa[e] * x1, u1 = wc[e] * y1;
                            const float ge = fmaf(wb[e], x2, t1), go = fmaf(-wb[e], x2, t1);
                            const float he = fmaf(wd[e], y2, u1), ho = fmaf(-wd[e], y2, u1);
                            v[0][e] = ge + he; v[1][e] = ge - he; v[2][e] = go - ho; v[3][e] = go + ho;
                        }
#pragma unroll
                        for (int pr = 0; pr < kTcNP; pr++)
#pragma unroll
                            for (int c = 0; c < 2; c++) {
                                const unsigned hi = pack_f16x2(v[pr][2 * c], v[pr][2 * c + 1]);
                                const float2 hf = unpack_f16x2(hi);
                                hw[pr][2 * h + c] = hi;
                                lw[pr][2 * h + c] = pack_f16x2(v[pr][2 * c] - hf.x, v[pr][2 * c + 1] - hf.y);
                            }
                    }
#pragma unroll
                    for (int pr = 0; pr < kTcNP; pr++) {
                        if (g > 0) { mbar_wait(BAR(kTcBarFree + pr), (g - 1) & 1u); tc_fence_after(); }
                        const unsigned ts = tlane + (unsigned)(kTcRingCol + 16 * pr + 4 * role);
                        tmem_st4(ts, hw[pr][0], hw[pr][1], hw[pr][2], hw[pr][3]);
                        tmem_st4(ts + 8u, lw[pr][0], lw[pr][1], lw[pr][2], lw[pr][3]);
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(BAR(kTcBarFull + pr));
                    }
                }
            }
            // ---- epilogue of this tile ----
            mbar_wait(BAR(kTcBarAccFull), it & 1u);
            tc_fence_after();
            const i64 t = tile * kTcFrames + f;
            TcEpi e;
            e.a0 = 0.0f; e.a1 = 0.0f; e.lmax = -3.0e38f; e.lmin = 3.0e38f;
            e.valid = t < T;
            e.T = (size_t)T;
            e.out = p.out + (size_t)b * (size_t)NM * (size_t)T + (e.valid ? t : 0);
            if (role == 0) tc_epi_role<NM, 0>(e, tlane);
            else if (role == 1) tc_epi_role<NM, 1>(e, tlane);
            else if (role == 2) tc_epi_role<NM, 2>(e, tlane);
            else tc_epi_role<NM, 3>(e, tlane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR(kTcBarAccFree));
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
