// fir_mma.cuh — fused s16 decode (stereo or mono) + downmix + polyphase FIR as a banded-Toeplitz product on the tensor
// cores (sm_100a), for the two named rate pairs (44.1 kHz and 48 kHz -> 16 kHz).
//
// Replaces libswresample's swr_convert inner loop behind the reference's
//   ffmpeg -ar 16000 -ac 1 -c:a pcm_s16le   (app/services/audio_processor.py:912-923).
//
// Why tensor cores: the polyphase FIR needs 92 (100) multiply-adds per output with 160 distinct phases.  With
// CUDA-core FFMAs every (phase, tap) pair is its own instruction; the fully unrolled immediate-operand form is a
// ~280 KB instruction stream that ncu shows stalled on instruction fetch (profiles/r01_fir_history.md:
// smsp no_instruction stall 12.4 per issue, 18 % issue utilisation), and register-operand FFMAs issue at half
// rate.  As a matrix product the filter bank is a constant operand that lives in registers and the code is a
// 200-instruction loop.
//
// Formulation.  A run = 160 consecutive outputs (one phase period at 44.1 kHz, L = 160) = 10 blocks of 16.
// For block b the 16 outputs read a window of <= 16*KS input frames starting at frame kb = (16 b M) / L of the
// run, so for 16 runs at once
//     D[16 runs x 16 outputs] = X[16 runs x 16 KS frames] * T_b[16 KS frames x 16 outputs]
// with T_b the (constant, zero padded) taps of the block.  mma.sync.m16n8k16 (f16 in, f32 accumulate) computes it:
// A = X from shared memory, B = T_b from registers (loaded once per persistent warp; MMA warp b owns block b).
// Within a k-step the 16 k-slots are permuted so that lane t owns frames 4t..4t+3: an A fragment row is one
// 8-byte shared-memory load (the filter table is built with the same permutation).
//
// Exactness.  The window value is v = L + R (stereo) or 2 x (mono): a 17-bit integer; the output is sum(t * v) / 2.  v = 128 * hv + lo
// with hv in [-512, 511] and lo in [0, 127]: both exact in f16.  Taps are scaled by 2^12 and split T = T_hi + T_lo
// (f16 each, |T - T_hi - T_lo| <= 2^-22 |T|).  All four cross terms are accumulated in f32:
//     out = 2^-6 * (T_hi + T_lo) . hv  +  2^-13 * (T_hi + T_lo) . lo
// f16 x f16 products are exact in f32, so the only differences from the f32 FMA chain of the CPU engine are the
// 2^-22 tap representation and the summation order — the same size as f32 rounding itself (tests: <= 1 LSB from
// libswresample, >= 99.8 % of samples identical, <= 1e-5 from the float64 restatement).
//
// Kernel: persistent CTA per SM, 16 warps in two roles that run concurrently on different pipes, coupled only by
// mbarriers (no block-wide barrier in the loop); tile = 16 runs, every buffer double buffered:
//   converter warps (6): wait for the tile's raw frames (one contiguous span fetched by 1-D bulk TMA, issued four
//       tiles ahead), split every frame into (hv, lo) f16 and write the two planes (run-major rows, pitch P with
//       P/2 = 8 or 24 mod 32 words: conflict-free 8-byte fragment loads).
//   MMA warps (10): wait for the planes, run the k-steps (all-zero tap k-steps skipped at compile time), round
//       half-to-even + clip to s16 (swr audioconvert) and store straight from the accumulator fragments: a lane
//       holds 2 consecutive outputs of a run, the two 8-output halves of a block fill one 32-byte sector, and a
//       block of 16 outputs is exactly one millisecond, so the per-millisecond sum of squares (uint64) that the
//       silence detector consumes is two shuffles away.
#pragma once
#include <cstdlib>

#include "b2a_common.cuh"

namespace b2a {

template <int IN_RATE> struct FirMmaTraits;
template <> struct FirMmaTraits<44100> { static constexpr int L = 160, M = 441, TAPS = 92, KS = 9, P = 560; };
template <> struct FirMmaTraits<48000> { static constexpr int L = 1, M = 3, TAPS = 100, KS = 10, P = 592; };

constexpr int kFmNout = 160;            // outputs per run (10 ms)
constexpr int kFmBlocks = 10;           // 16-output blocks per run
constexpr int kFmRT = 16;               // runs per tile (= MMA M)
constexpr int kFmMmaWarps = 10;         // MMA warp b owns block b
constexpr int kFmCvtWarps = 6;          // converter warps
constexpr int kFmCvtThreads = kFmCvtWarps * 32;
constexpr int kFmThreads = (kFmMmaWarps + kFmCvtWarps) * 32;
constexpr int kFmTapShift = 12;         // taps are scaled by 2^12 before the f16 split
constexpr int kFmRawBufs = 4;           // raw-tile ring: TMA loads are issued this many tiles ahead
constexpr int kCvtIlp = 8;              // frame pairs in flight per converter thread
constexpr int kFmBarBytes = 16;         // one mbarrier slot (8 bytes on the GPU; the emulation keeps counters in 16)

template <int IN_RATE, int CH>
struct FirMmaGeom {
    using TR = FirMmaTraits<IN_RATE>;
    static constexpr int L = TR::L, M = TR::M, TAPS = TR::TAPS, KS = TR::KS, P = TR::P;
    static constexpr int CENTER = (TAPS - 1) / 2;
    static constexpr int S = kFmNout * M / L;                       // input frames per run (441 / 480)
    static constexpr int kb(int b) { return (16 * b * M) / L; }     // window start of block b (frames from the row origin)
    static constexpr int PC = kb(kFmBlocks - 1) + 16 * KS;          // columns actually read by the MMAs (540 / 592)
    static constexpr int PW = P / 2;                                // words per plane row
    static constexpr int FB = 2 * CH;                               // bytes per input frame (s16 samples)
    static constexpr int CENTER_BYTES16 = (CENTER * FB + 15) / 16 * 16;
    static constexpr int AL = (CENTER_BYTES16 - CENTER * FB) / FB;  // frames the raw tile starts early (16-byte alignment)
    static constexpr int RAWF = (kFmRT - 1) * S + PC;               // frames a tile's rows touch
    static constexpr int RAW_BYTES = ((AL + RAWF) * FB + 15) / 16 * 16;
    static constexpr int PLANE_WORDS = kFmRT * PW;                  // one plane (hv or lo) of one tile
    static constexpr int PLANES_BYTES = 2 * PLANE_WORDS * 4;        // hv + lo
    static constexpr int NBARS = 2 * kFmRawBufs + 4;
    static constexpr int SMEM_BYTES = kFmRawBufs * RAW_BYTES + 2 * PLANES_BYTES + NBARS * kFmBarBytes;
    static constexpr int PAIRS_ROW = PC / 2;                        // frame pairs converted per row
    static constexpr int PAIRS = kFmRT * PAIRS_ROW;
    static constexpr int CVT_TRIPS = (PAIRS + kFmCvtThreads - 1) / kFmCvtThreads;
    // k-step s contributes to the 8-output half nt of SOME block iff [16 s, 16 s + 16) meets a window of that half
    static constexpr bool needed(int s, int nt) {
        int lo = 1 << 30, hi = 0;
        for (int b = 0; b < kFmBlocks; b++)
            for (int j = 8 * nt; j < 8 * nt + 8; j++) {
                const int rel = ((16 * b + j) * M) / L - kb(b);
                lo = rel < lo ? rel : lo;
                hi = rel + TAPS > hi ? rel + TAPS : hi;
            }
        return 16 * s < hi && 16 * s + 16 > lo;
    }
    static_assert((kFmRT * S * FB) % 16 == 0 && (CENTER_BYTES16 - CENTER * FB) % FB == 0, "tile pitch / lead-in must keep 16-byte alignment");
    static_assert(P % 4 == 0 && (PW % 32 == 8 || PW % 32 == 24), "plane pitch must make 8-byte fragment loads conflict free");
    static_assert(P >= PC && PC % 2 == 0, "row must hold every column the MMAs read");
    static_assert(16 * KS >= ((15 * M) / L + 1) + TAPS, "K must cover the widest window of a block");
    static_assert(kb(1) % 4 == 0 && kb(3) % 4 == 0 && kb(9) % 4 == 0, "block windows must start on 8-byte plane boundaries");
};

struct FirMmaArgs {
    const unsigned char* in;     // raw interleaved s16 frames (stereo or mono)
    int16_t* out_s16;            // nullable
    u64* energy;                 // nullable
    const uint2* btab;           // [block][KS][nt][term][lane] B fragments (taps), see build_fir_mma_table
    i64 tile_lo, tile_hi;        // tiles [tile_lo, tile_hi) are produced (tile t = runs [t*RT, (t+1)*RT))
    int phases;                  // profiling aid (env B2A_FIR_PHASES, tools/fir_phase_probe.py): bit 0 convert, bit 1 k-steps,
                                 // bit 2 epilogue; 7 = the product, anything else gives wrong results on purpose
};

// ---- primitives (GPU: PTX; TEST-ONLY emulation: tests/emu) -------------------------------------------------
#ifndef B2A_EMU
typedef unsigned saddr_t;
__device__ __forceinline__ saddr_t smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(saddr_t bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(saddr_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(saddr_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(saddr_t bar, unsigned parity) {
    unsigned ok;
    do {
        // the suspend-time hint parks the warp in hardware until the phase completes (or the hint expires) instead of
        // re-issuing the poll through the MIO queue, which the shared-memory stores of the working warps need
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    } while (!ok);
}
// 1-D bulk copy global -> shared (TMA engine), completion counted on the mbarrier
__device__ __forceinline__ void bulk_load(saddr_t dst, const void* src, unsigned bytes, saddr_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// packed f16 subtraction on raw bits
__device__ __forceinline__ unsigned hsub2_bits(unsigned a, unsigned b) {
    unsigned r;
    asm("sub.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
#else
typedef uintptr_t saddr_t;
static inline saddr_t smem_addr(const void* p) { return (uintptr_t)p; }
// mbarrier stand-in: [0] completed phases, [1] arrivals still expected, [2] arrival count, [3] bytes still expected
static inline void emu_mbar_try_complete(unsigned* b) {
    if (b[1] == 0 && (int)b[3] <= 0) { b[0]++; b[1] = b[2]; b[3] = 0; }
}
static inline void mbar_init(saddr_t bar, unsigned count) { unsigned* b = (unsigned*)bar; b[0] = 0; b[1] = count; b[2] = count; b[3] = 0; }
static inline void mbar_fence_init() {}
static inline void mbar_expect_tx(saddr_t bar, unsigned bytes) { unsigned* b = (unsigned*)bar; b[3] += bytes; b[1]--; emu_mbar_try_complete(b); }
static inline void mbar_arrive(saddr_t bar) { unsigned* b = (unsigned*)bar; b[1]--; emu_mbar_try_complete(b); }
static inline void mbar_wait(saddr_t bar, unsigned parity) { while ((((volatile unsigned*)bar)[0] & 1u) == parity) emu::yield(); }
static inline void bulk_load(saddr_t dst, const void* src, unsigned bytes, saddr_t bar) {
    memcpy((void*)dst, src, bytes);
    unsigned* b = (unsigned*)bar;
    b[3] -= bytes;
    emu_mbar_try_complete(b);
}
static inline unsigned hsub2_bits(unsigned a, unsigned b) {
    unsigned lo = emu::f32_to_f16(emu::f16_to_f32((unsigned short)(a & 0xffff)) - emu::f16_to_f32((unsigned short)(b & 0xffff)));
    unsigned hi = emu::f32_to_f16(emu::f16_to_f32((unsigned short)(a >> 16)) - emu::f16_to_f32((unsigned short)(b >> 16)));
    return lo | (hi << 16);
}
static inline void mma_16816(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) { emu::mma_m16n8k16_f16(d, a, b0, b1); }
#endif

// one block of one 16-run tile: D12 = (T_hi + T_lo) . hv,  D34 = (T_hi + T_lo) . lo
// phv / plo: this lane's row pointers (row g) into the hv / lo planes at column kb + 4t; row g+8 is 8*PW words further
template <int IN_RATE>
__device__ __forceinline__ void fir_mma_block(const unsigned* __restrict__ phv, const unsigned* __restrict__ plo,
                                              const uint2 (&breg)[FirMmaTraits<IN_RATE>::KS][2][2],
                                              float (&d12)[2][4], float (&d34)[2][4]) {
    using G = FirMmaGeom<IN_RATE, 2>;      // plane geometry does not depend on the channel count
#pragma unroll
    for (int s = 0; s < G::KS; s++) {
        // lane t owns frames 4t..4t+3 of the k-step: word 0 = k-slots (2t, 2t+1), word 1 = k-slots (2t+8, 2t+9)
        const uint2 h0 = *(const uint2*)(phv + 8 * s);
        const uint2 h1 = *(const uint2*)(phv + 8 * s + 8 * G::PW);
        const uint2 l0 = *(const uint2*)(plo + 8 * s);
        const uint2 l1 = *(const uint2*)(plo + 8 * s + 8 * G::PW);
        const unsigned ahv[4] = {h0.x, h1.x, h0.y, h1.y};
        const unsigned alo[4] = {l0.x, l1.x, l0.y, l1.y};
        // issue order: both 8-output halves of the T_hi term, then of the T_lo term, so that two HMMAs on the same
        // accumulator are four instructions apart (the legacy HMMA result latency is several issue slots)
#pragma unroll
        for (int term = 0; term < 2; term++)
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                if (!G::needed(s, nt)) continue;      // all-zero taps for every block: compile-time skip
                mma_16816(d12[nt], ahv, breg[s][nt][term].x, breg[s][nt][term].y);
                mma_16816(d34[nt], alo, breg[s][nt][term].x, breg[s][nt][term].y);
            }
    }
}

template <int IN_RATE, int CH>
__global__ void __launch_bounds__(kFmThreads, 1) fir_mma_kernel(const FirMmaArgs a) {
    using G = FirMmaGeom<IN_RATE, CH>;
    B2A_DYN_SMEM(smem);
    unsigned char* raw0 = smem;                                                  // [kFmRawBufs][RAW_BYTES]
    unsigned* planes0 = (unsigned*)(smem + kFmRawBufs * G::RAW_BYTES);           // [2][hv plane | lo plane]
    const saddr_t bars = smem_addr(smem + kFmRawBufs * G::RAW_BYTES + 2 * G::PLANES_BYTES);
    // barrier slots: RF raw full (TMA), RE raw empty, PF planes full, PE planes empty
    auto RF = [&](int b) { return bars + (unsigned)((0 + b) * kFmBarBytes); };
    auto RE = [&](int b) { return bars + (unsigned)((kFmRawBufs + b) * kFmBarBytes); };
    auto PF = [&](int b) { return bars + (unsigned)((2 * kFmRawBufs + b) * kFmBarBytes); };
    auto PE = [&](int b) { return bars + (unsigned)((2 * kFmRawBufs + 2 + b) * kFmBarBytes); };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int b = 0; b < kFmRawBufs; b++) {
            mbar_init(RF(b), 1);
            mbar_init(RE(b), kFmCvtWarps);
        }
        for (int b = 0; b < 2; b++) {
            mbar_init(PF(b), kFmCvtWarps);
            mbar_init(PE(b), kFmMmaWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const i64 tile0 = a.tile_lo + blockIdx.x;
    const i64 stride = gridDim.x;

    if (warp < kFmMmaWarps) {
        // =================================== MMA warps: block b = warp ===================================
        const int g = lane >> 2, t = lane & 3;
        // constant operand: this warp's block of the filter bank, as B fragments (hi, lo) per k-step and 8-output half
        uint2 breg[G::KS][2][2];
        {
            const uint2* bt = a.btab + (size_t)warp * (G::KS * 2 * 2 * 32) + lane;
#pragma unroll
            for (int s = 0; s < G::KS; s++)
#pragma unroll
                for (int nt = 0; nt < 2; nt++)
#pragma unroll
                    for (int term = 0; term < 2; term++)
                        breg[s][nt][term] = G::needed(s, nt) ? bt[((s * 2 + nt) * 2 + term) * 32] : make_uint2(0u, 0u);
        }
        const int kbw = ((16 * warp * G::M) / G::L) / 2;            // window start of this block, in plane words
        const int frag_off = g * G::PW + kbw + 2 * t;               // row g, frames kb + 4t ..
        const int col = 16 * warp + 2 * t;                          // first output column this lane owns
        int it = 0;
        for (i64 tile = tile0; tile < a.tile_hi; tile += stride, it++) {
            const int b = it & 1;
            const unsigned ph = (unsigned)((it >> 1) & 1);
            const unsigned* phv = planes0 + (size_t)b * (2 * G::PLANE_WORDS) + frag_off;
            const unsigned* plo = phv + G::PLANE_WORDS;
            float d12[2][4], d34[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; nt++)
#pragma unroll
                for (int e = 0; e < 4; e++) { d12[nt][e] = 0.f; d34[nt][e] = 0.f; }
            mbar_wait(PF(b), ph);
            if (a.phases & 2) fir_mma_block<IN_RATE>(phv, plo, breg, d12, d34);
            __syncwarp();
            if (lane == 0) mbar_arrive(PE(b));                      // planes[b] may be refilled
            // ---- epilogue: d[0],d[1] = run g, outputs 2t, 2t+1 of an 8-output half; d[2],d[3] = run g + 8 ----
            if (!(a.phases & 4)) continue;
            const i64 m_run = ((i64)tile * kFmRT + g) * kFmNout + col;          // first output this lane owns in run g
            u64 e_lo = 0, e_hi = 0;                                              // sum of squares of this lane's outputs (runs g, g+8)
#pragma unroll
            for (int nt = 0; nt < 2; nt++) {
                const float sc12 = 1.0f / 64.0f, sc34 = 1.0f / 8192.0f;
                const int q0 = quant_s16(fmaf(d12[nt][0], sc12, d34[nt][0] * sc34));
                const int q1 = quant_s16(fmaf(d12[nt][1], sc12, d34[nt][1] * sc34));
                const int q2 = quant_s16(fmaf(d12[nt][2], sc12, d34[nt][2] * sc34));
                const int q3 = quant_s16(fmaf(d12[nt][3], sc12, d34[nt][3] * sc34));
                if (a.out_s16) {
                    *(unsigned*)(a.out_s16 + m_run + 8 * nt) = (unsigned)(q0 & 0xffff) | ((unsigned)q1 << 16);
                    *(unsigned*)(a.out_s16 + m_run + 8 * kFmNout + 8 * nt) = (unsigned)(q2 & 0xffff) | ((unsigned)q3 << 16);
                }
                e_lo += (u64)(unsigned)(q0 * q0) + (u64)(unsigned)(q1 * q1);
                e_hi += (u64)(unsigned)(q2 * q2) + (u64)(unsigned)(q3 * q3);
            }
            // the 16 outputs of (run, block) are one millisecond: reduce over the 4 lanes t = 0..3 that share a run
            e_lo += __shfl_xor_sync(0xffffffffu, e_lo, 1);
            e_hi += __shfl_xor_sync(0xffffffffu, e_hi, 1);
            e_lo += __shfl_xor_sync(0xffffffffu, e_lo, 2);
            e_hi += __shfl_xor_sync(0xffffffffu, e_hi, 2);
            if (a.energy && t == 0) {
                const i64 ms = ((i64)tile * kFmRT + g) * kFmBlocks + warp;
                a.energy[ms] = e_lo;
                a.energy[ms + 8 * kFmBlocks] = e_hi;
            }
        }
    } else {
        // ============================ converter warps (and the TMA issuer) ============================
        const int ctid = tid - kFmMmaWarps * 32;
        auto issue = [&](i64 tile, int b) {
            // raw span of the tile: frames [tile*RT*S - CENTER - AL, +RAW_BYTES/4), 16-byte aligned at both ends
            const unsigned char* src = a.in + ((i64)tile * kFmRT * G::S - G::CENTER - G::AL) * G::FB;
            mbar_expect_tx(RF(b), (unsigned)G::RAW_BYTES);
            constexpr int PIECE = 16384;
#pragma unroll 1
            for (int off = 0; off < G::RAW_BYTES; off += PIECE) {
                const int nb = (G::RAW_BYTES - off) < PIECE ? (G::RAW_BYTES - off) : PIECE;
                bulk_load(smem_addr(raw0 + (size_t)b * G::RAW_BYTES + off), src + off, (unsigned)nb, RF(b));
            }
        };
        if (ctid == 0) {
            for (int r = 0; r < kFmRawBufs; r++)
                if (tile0 + r * stride < a.tile_hi) issue(tile0 + r * stride, r);
        }
        int rb = 0;                 // raw ring slot of this tile
        unsigned rph = 0;           // its phase parity
        int it = 0;
        i64 tile = tile0;
        for (; tile < a.tile_hi; tile += stride, it++) {
            const int b = it & 1;
            const unsigned ph = (unsigned)((it >> 1) & 1);
            mbar_wait(RF(rb), rph);                                 // raw frames landed
            mbar_wait(PE(b), ph ^ 1u);                              // planes[b] released by the MMA warps (passes on first use)
            // ---- raw s16 stereo -> hv / lo f16 planes.  pair q = ctid + 192 j of the tile <-> row n = q / PAIRS_ROW,
            //      columns 2kp, 2kp+1 (kp = q % PAIRS_ROW); for a fixed trip j the row is n0(j) or n0(j)+1.
            if (a.phases & 1) {
                const unsigned* rawt = (const unsigned*)(raw0 + (size_t)rb * G::RAW_BYTES) + G::AL + 2 * ctid;             // stereo: one word per frame
                const unsigned short* rawm = (const unsigned short*)(raw0 + (size_t)rb * G::RAW_BYTES) + G::AL + 2 * ctid;  // mono
                unsigned* plt = planes0 + (size_t)b * (2 * G::PLANE_WORDS) + ctid;
#pragma unroll
                for (int j0 = 0; j0 < G::CVT_TRIPS; j0 += kCvtIlp) {
                    unsigned r0[kCvtIlp], r1[kCvtIlp];
                    int po[kCvtIlp];
                    bool ok[kCvtIlp];
#pragma unroll
                    for (int e = 0; e < kCvtIlp; e++) {
                        const int j = j0 + e;
                        const int qbase = kFmCvtThreads * j;                           // q = qbase + ctid
                        const int n0 = qbase / G::PAIRS_ROW;
                        const int thr = (n0 + 1) * G::PAIRS_ROW - qbase;               // ctid >= thr  <=>  row n0 + 1
                        const bool up = ctid >= thr;
                        ok[e] = j < G::CVT_TRIPS && (qbase + ctid) < G::PAIRS;
                        // raw frame index = n*S + 2*kp = 2q + n*(S - 2*PAIRS_ROW);  plane word = n*PW + kp = q + n*(PW - PAIRS_ROW)
                        const int ro = 2 * qbase + (n0 + (up ? 1 : 0)) * (G::S - 2 * G::PAIRS_ROW);
                        po[e] = qbase + (n0 + (up ? 1 : 0)) * (G::PW - G::PAIRS_ROW);
                        if (CH == 2) {
                            r0[e] = ok[e] ? rawt[ro] : 0u;
                            r1[e] = ok[e] ? rawt[ro + 1] : 0u;
                        } else {
                            r0[e] = ok[e] ? (unsigned)rawm[ro] : 0u;
                            r1[e] = ok[e] ? (unsigned)rawm[ro + 1] : 0u;
                        }
                    }
#pragma unroll
                    for (int e = 0; e < kCvtIlp; e++) {
                        // u = v + 65536 in [0, 131070] with v = L + R (stereo) or 2 x (mono): u >> 7 = hv + 512, u & 127 = lo
                        const unsigned u0 = CH == 2 ? (unsigned)__dp2a_lo((int)r0[e], 0x0101, 65536) : (unsigned)(2 * (int)(short)r0[e] + 65536);
                        const unsigned u1 = CH == 2 ? (unsigned)__dp2a_lo((int)r1[e], 0x0101, 65536) : (unsigned)(2 * (int)(short)r1[e] + 65536);
                        const unsigned whv = ((u1 >> 7) << 16) + ((u0 >> 7) + 0x64006400u);       // f16 bits of 1024 + hv + 512
                        const unsigned wlo = ((u1 & 127u) << 16) + ((u0 & 127u) + 0x64006400u);   // f16 bits of 1024 + lo
                        if (ok[e]) {
                            plt[po[e]] = hsub2_bits(whv, 0x66006600u);                            // minus 1536: exact integers
                            plt[po[e] + G::PLANE_WORDS] = hsub2_bits(wlo, 0x64006400u);           // minus 1024
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(PF(b)); mbar_arrive(RE(rb)); }
            // refill this ring slot with the tile kFmRawBufs steps ahead once every converter warp has released it
            if (ctid == 0 && tile + kFmRawBufs * stride < a.tile_hi) {
                mbar_wait(RE(rb), rph);
                issue(tile + kFmRawBufs * stride, rb);
            }
            if (++rb == kFmRawBufs) { rb = 0; rph ^= 1u; }
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------
struct FirMmaPlan {
    i64 out_lo, out_hi;    // outputs [out_lo, out_hi) come from the tensor-core kernel
};

const uint2* get_fir_mma_table(int in_rate);   // device table for the current device (b2a_host.cu); nullptr + error on failure

template <int IN_RATE, int CH>
static inline int fir_mma_launch(const void* d_in, i64 n_in, int16_t* d_out_s16, u64* d_energy, FirMmaPlan* plan, cudaStream_t stream, i64 first_tile = 1) {
    using G = FirMmaGeom<IN_RATE, CH>;
    plan->out_lo = plan->out_hi = 0;
    // tile t reads frames [t*RT*S - CENTER - AL, that + RAW_BYTES/FB), FB = bytes per frame: t >= 1 keeps the start inside the clip; the first
    // 16 runs (reflect head) and the tail go to the table-driven kernel
    const i64 tile_lo = first_tile < 1 ? 1 : first_tile;       // (the tcgen05 kernel hands over the 16-run tiles behind its last span)
    const i64 span_end = (i64)G::RAW_BYTES / G::FB - G::CENTER - G::AL;      // relative to the tile's first run start (round 1 divided by 4 for mono too: the last tile read past the clip on some lengths)
    const i64 tile_hi = (n_in - span_end) >= 0 ? (n_in - span_end) / ((i64)kFmRT * G::S) + 1 : 0;   // exclusive
    if (tile_hi <= tile_lo) return 0;
    const uint2* tab = get_fir_mma_table(IN_RATE);
    if (!tab) return B2A_ECUDA;
    auto k = fir_mma_kernel<IN_RATE, CH>;
    // the dynamic shared-memory opt-in is a per-device function attribute: remember which devices have it
    // (idempotent; a benign race only repeats the call)
    static unsigned long long attr_mask = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !((attr_mask >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fir_mma_kernel)");
        if (dev >= 0 && dev < 64) attr_mask |= 1ull << dev;
    }
    FirMmaArgs a;
    a.in = (const unsigned char*)d_in; a.out_s16 = d_out_s16; a.energy = d_energy; a.btab = tab;
    a.tile_lo = tile_lo; a.tile_hi = tile_hi;
    a.phases = 7;
#ifdef B2A_PROFILE
    if (const char* ph = getenv("B2A_FIR_PHASES")) a.phases = atoi(ph);      // profiling builds only (-DB2A_PROFILE): masks kernel phases, WRONG output
#endif
    const i64 tiles = tile_hi - tile_lo;
    unsigned grid = (unsigned)(tiles < 148 ? tiles : 148);                   // persistent: one CTA per SM
#if defined(B2A_PROFILE) || defined(B2A_EMU)
    if (const char* gs = getenv("B2A_FIR_GRID")) {                           // emulation / profiling builds only: few CTAs => many tiles per CTA
        const int gv = atoi(gs);
        if (gv > 0 && (unsigned)gv < grid) grid = (unsigned)gv;
    }
#endif
    B2A_LAUNCH(k, grid, kFmThreads, G::SMEM_BYTES, stream, a);
    B2A_CHECK_LAUNCH("fir_mma_kernel");
    plan->out_lo = tile_lo * kFmRT * kFmNout;
    plan->out_hi = tile_hi * kFmRT * kFmNout;
    return 1;
}

}  // namespace b2a
