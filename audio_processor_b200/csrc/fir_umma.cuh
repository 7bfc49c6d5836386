// fir_umma.cuh — fused s16 stereo decode + downmix + polyphase FIR on the 5th-generation tensor cores
// (tcgen05.mma, accumulators in TMEM), sm_100a, for the two named rate pairs (44.1 kHz and 48 kHz -> 16 kHz).
//
// Replaces libswresample's swr_convert inner loop behind the reference's
//   ffmpeg -ar 16000 -ac 1 -c:a pcm_s16le   (app/services/audio_processor.py:912-923).
//
// Same banded-Toeplitz formulation and the same exact f16 splits as fir_mma.cuh (the legacy mma.sync kernel, still
// used for mono input), re-laid for tcgen05:
//   run   = 160 consecutive outputs = 10 blocks of 16 (one block = one millisecond at 16 kHz); a run consumes S input
//           frames (441 / 480), so 128 consecutive runs are 128 rows that share one filter matrix per block;
//   block b of 128 runs:  D[128 x 32] = X[128 x 16 KS] * [T_hi | T_lo][16 KS x 32]
//           X = (hv + lo / 128) with v = L + R = 128 hv + lo  (two f16 planes, both exact, accumulated into the same D),
//           T = 2^12 taps split T_hi + T_lo (f16 each); columns 0-15 hold the T_hi partial sums of the 16 outputs,
//           columns 16-31 the T_lo ones; the epilogue adds the two and scales by 2^-6.
//   => 2 KS tcgen05.mma (M128 N32 K16, kind::f16, f32 accumulate) per block, issued by one thread.
//
// Shared memory cannot hold a tile's planes (128 rows x 544 frames x 4 B = 278 KB) next to the 92 KB filter bank, and a
// raw tile (226 KB) does not fit either, so the planes are a COLUMN RING: 4 pieces of 64 columns for all 128 rows
// (+ one mirrored chunk so that a K=16 operand never straddles the wrap), stored in the canonical no-swizzle K-major
// layout the MMA reads: [chunk of 8 columns][row][16 B], chunk pitch LBO = 2064 B (128 rows x 16 B + 16 B of padding
// that spreads the stores over the banks), 8-row groups SBO = 128 B apart.  Layout, descriptor fields and the TMEM
// accumulator layout (lane = row, column = n) are pinned on hardware by tools/probes/umma_probe.cu.
//
// Alignment.  A run is 441 frames = 1764 bytes, so consecutive runs sit at four different offsets against the 16-byte
// grid of the input.  A tile therefore takes every FOURTH run (class c = run % 4, rows 7056 bytes apart: 16-byte
// aligned), and its plane columns start SH_c frames before the filter origin so that every row of the tile begins on a
// quad; the shift is folded into the class's own copy of the filter bank (4 x 92 KB in global memory, one resident per
// CTA).  Converter warps then read aligned 16-byte quads straight from global memory (a half-warp covers 256 contiguous
// bytes of one row; no shared-memory staging ring), split them into the two planes and store 8 bytes per plane.  CTA b
// works on class b % 4, and CTAs b..b+3 walk the same 512-run spans together, so the rows' overlapping halos meet in L2.
//
// Roles (768 threads = 6 warpgroups, one persistent CTA per SM), coupled only by mbarriers; registers are moved between
// the warpgroups with setmaxnreg:
//   warps 0-3   epilogue: warp q owns TMEM lanes 32q..32q+31 (thread = run): tcgen05.ld 32 columns, add the two
//               halves, round half-to-even + saturate to s16 (swr audioconvert), two 16-byte stores, and the exact uint64
//               sum of squares of the millisecond for the silence detector;
//   warp  4     MMA issuer: runs the compile-time k-step schedule (fir_umma_schedule, column order) fully unrolled in
//               uniform control flow, one elected lane issuing the tcgen05.mma and the tcgen05.commit arrivals that
//               release ring pieces back to the converters and hand accumulators to the epilogue (ring of 8
//               accumulators); owns the TMEM allocation; warps 5-7 only pad the warpgroup;
//   warps 8-23  converters (global -> registers -> planes): two register sets alternate and every register quad is
//               refilled with the piece two ahead the moment it is consumed.
//
// Status (profiles/r01_fir_umma.md): bit-for-bit the same results as the mma.sync kernel's gates on B200, 306-350 us for a
// one-hour clip against 246 us: the register-staged loads stream at ~2 TB/s (L1 miss tracking, not registers, bounds the
// bytes in flight), so this kernel is opt-in (B2A_FIR_IMPL=umma) until the planes move to TMEM and TMA feeds a raw ring.
#pragma once
#include <utility>

#include "fir_mma.cuh"

namespace b2a {

constexpr int kFuRT = 128;                 // runs per tile = UMMA M
constexpr int kFuClasses = 4;              // a tile takes every 4th run (class = run % 4)
constexpr int kFuSpan = kFuRT * kFuClasses;   // runs covered by the 4 tiles of a span
constexpr int kFuRun0 = 16;                // first run of span 0 (multiple of 16: the head is the legacy kernel's first tile)
constexpr int kFuPiece = 64;               // columns per ring piece
constexpr int kFuLbo = kFuRT * 16 + 16;    // bytes between consecutive chunks of a plane
constexpr int kFuSbo = 128;                // bytes between 8-row groups inside a chunk
constexpr int kFuEpiWarps = 4;
constexpr int kFuCvtWarps = 16;
constexpr int kFuMmaWarp = 4;              // warp 4 issues the MMAs and owns the TMEM allocation; warps 5-7 only pad its warpgroup (setmaxnreg is per warpgroup)
constexpr int kFuCvtWarp0 = 8;             // converters = warpgroups 2 and 3
constexpr int kFuThreads = (kFuCvtWarp0 + kFuCvtWarps) * 32;
constexpr int kFuRegsEpi = 72, kFuRegsMma = 56, kFuRegsCvt = 88;    // 128 (72 + 56 + 4 x 88) = 61440 <= 768 x 80 launch registers: setmaxnreg.inc only draws on what the CTA released
constexpr int kFuRowPairs = kFuRT / 2 / kFuCvtWarps;            // row pairs per converter warp and piece
static_assert(4 * 32 * (kFuRegsEpi + kFuRegsMma) + kFuCvtWarps * 32 * kFuRegsCvt <= kFuThreads * (65536 / kFuThreads / 8 * 8),
              "setmaxnreg.inc can only draw on registers the CTA released: the role budgets must fit the launch allocation");
constexpr int kFuDSlots = 8;               // accumulator ring (blocks overlap in time: k-steps are issued in column order)
constexpr int kFuDCols = 32;               // TMEM columns per accumulator: [T_hi sums | T_lo sums]
constexpr int kFuTmemCols = kFuDSlots * kFuDCols;               // 256 (power of two >= 32)
constexpr int kFuBTile = 32 * 16 * 2;      // one [N = 32][K = 16] f16 operand tile
constexpr int kFuBLbo = 128, kFuBSbo = 256;                     // B tile: [n group of 8][k chunk][8 rows][16 B]
constexpr int kFuSchedMax = 112;           // k-steps of one tile (90 / 100)
// schedule item: bits 0-6 chunk of the k-step's operand (kbp(b) / 8 + 2 s), 7-13 filter tile (b KS + s), 14-17 block b,
// 18 first k-step of the block, 19 last, 20-23 pieces of the tile that must be full before it, 24-27 pieces to release after it
constexpr unsigned kFuItFirst = 1u << 18, kFuItLast = 1u << 19;
struct FirUmmaSched { int n; unsigned w[kFuSchedMax]; };
constexpr unsigned kFuIdesc = (1u << 4) | ((unsigned)(kFuDCols >> 3) << 17) | ((unsigned)(kFuRT >> 4) << 24);   // f16 x f16 -> f32, K-major, N = 32, M = 128

template <int IN_RATE>
struct FirUmmaGeom {
    using TR = FirMmaTraits<IN_RATE>;
    static constexpr int L = TR::L, DEC = TR::M, TAPS = TR::TAPS;
    static constexpr int CENTER = (TAPS - 1) / 2;
    static constexpr int S = kFmNout * DEC / L;                          // input frames per run (441 / 480)
    static constexpr int kb(int b) { return (16 * b * DEC) / L; }        // exact window start of block b (column of the row)
    static constexpr int kbp(int b) { return kb(b) & ~7; }               // rounded down to a chunk
    static constexpr int window() {                                      // columns a block's window must span from kbp(b)
        int w = 0;
        for (int b = 0; b < kFmBlocks; b++)
            for (int j = 0; j < 16; j++) {
                const int e = ((16 * b + j) * DEC) / L - kbp(b) + TAPS + 3;     // + the largest class shift
                w = e > w ? e : w;
            }
        return w;
    }
    // class c = rows run0 + c + 4 r: frames by which the rows' first quad precedes the filter origin S run - CENTER
    static constexpr int shift(int run0, int c) { return (int)((((long long)(run0 + c)) * S - CENTER) & 3); }
    static constexpr int KS = (window() + 15) / 16;                      // k-steps per block (9 / 10)
    static constexpr int COLS = kbp(kFmBlocks - 1) + 16 * KS;            // columns of a row the MMAs read
    static constexpr int PIECES = (COLS + kFuPiece - 1) / kFuPiece;      // ring pieces per tile (9 / 10)
    // one phase (48 kHz: L = 1) and chunk-aligned block windows => every block has the same filter matrix
    static constexpr bool SHARED_B = (L == 1) && (kb(1) % 8 == 0);
    static constexpr int BBLOCKS = SHARED_B ? 1 : kFmBlocks;
    static constexpr int B_BYTES = BBLOCKS * KS * kFuBTile;
    // ring: 4 pieces next to the 92 KB filter bank of 44.1 kHz; 6 where the bank is one shared 10 KB matrix (48 kHz,
    // whose 160-column windows span 4 pieces)
    static constexpr int RING_PIECES = SHARED_B ? 6 : 4;
    static constexpr int RING_CHUNKS = RING_PIECES * kFuPiece / 8;       // chunks of 8 columns (+ 1 mirror of chunk 0)
    static constexpr int PLANE_BYTES = (RING_CHUNKS + 1) * kFuLbo;
    static constexpr int NBARS = 2 * RING_PIECES + 2 * kFuDSlots;
    static constexpr int SMEM_BYTES = 2 * PLANE_BYTES + B_BYTES + NBARS * kFmBarBytes + 16;
    static constexpr int piece_last(int b) { return (kbp(b) + 16 * KS - 1) / kFuPiece; }                       // last piece block b reads
    static constexpr int pieces_free_after(int b) { return b + 1 < kFmBlocks ? kbp(b + 1) / kFuPiece : PIECES; }  // pieces no later block reads
    static constexpr int max_span() {
        int m = 0;
        for (int b = 0; b < kFmBlocks; b++) { const int s = piece_last(b) - kbp(b) / kFuPiece + 1; m = s > m ? s : m; }
        return m;
    }
    static_assert(max_span() < RING_PIECES, "a block window must leave one ring piece for the converters to run ahead");
    static_assert(SMEM_BYTES <= 232448, "shared-memory budget (227 KB per CTA)");
    static_assert(kFmBlocks * KS <= kFuSchedMax && PIECES < 16 && kFmBlocks * KS < 128, "schedule item fields");
    static constexpr int ROWQ = S;                                       // quads between consecutive rows of a tile (4 runs)
    static constexpr int SPANQ = kFuSpan * S / 4;                        // quads between consecutive spans
};

struct FirUmmaArgs {
    const unsigned char* in;     // interleaved s16 stereo frames
    int16_t* out_s16;            // nullable
    u64* energy;                 // nullable
    const uint4* btab;           // [class][B_BYTES] filter banks as UMMA B tiles, see build_fir_umma_table
    unsigned long long* trace;   // profiling aid (env B2A_FIR_TRACE, tools/fir_trace.py): CTA 0 records clock64 at pipeline events; nullptr otherwise
    int phases;                  // profiling aid (env B2A_FIR_PHASES): bit 0 plane stores, 2 epilogue, 3 global loads, 4 L2 prefetch of the next span; 31 = the product
    int spans;                   // spans [0, spans): span t = runs [kFuRun0 + 512 t, +512); CTA b converts class b % 4 of spans b / 4, b / 4 + gridDim / 4, ..
};

// ---- primitives (GPU: PTX; TEST-ONLY emulation: tests/emu) -------------------------------------------------
#ifndef B2A_EMU
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(saddr_t slot, unsigned cols) {      // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned base, unsigned cols) {   // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
// shared-memory matrix descriptor, SWIZZLE_NONE, K-major: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ unsigned long long umma_desc(saddr_t addr, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((addr >> 4) & 0x3fffu) | ((unsigned long long)((lbo >> 4) & 0x3fffu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// D[tmem] (+)= A[smem] * B[smem], M = 128, N = 32, K = 16, one thread issues
__device__ __forceinline__ void umma_f16(unsigned d_tmem, saddr_t a, unsigned a_lbo, unsigned a_sbo, saddr_t b, unsigned b_lbo, unsigned b_sbo,
                                         unsigned idesc, unsigned accumulate) {
    const unsigned long long da = umma_desc(a, a_lbo, a_sbo), db = umma_desc(b, b_lbo, b_sbo);
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ unsigned desc_start(saddr_t addr) { return (addr >> 4) & 0x3fffu; }
__device__ __forceinline__ void desc_origin(const void*) {}
// same with the two 32-bit halves of each descriptor
__device__ __forceinline__ void umma_f16_desc(unsigned d_tmem, unsigned a_lo, unsigned a_hi, unsigned b_lo, unsigned b_hi, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
// the same, executed by a whole (converged) warp: one elected lane issues.  Issuing from warp-uniform code lets the
// compiler keep descriptors and addresses in uniform registers instead of wrapping every tcgen05 instruction of a
// single-lane branch in an election loop.
__device__ __forceinline__ void umma_f16_desc_warp(unsigned d_tmem, unsigned a_lo, unsigned a_hi, unsigned b_lo, unsigned b_hi, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\telect.sync _|e, 0xffffffff;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_warp(saddr_t bar) {
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
// arrive on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(saddr_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 TMEM columns of this warp's 32 lanes (thread = lane) -> registers; waits for the data
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned hfma2_bits(unsigned a, unsigned b, unsigned c) {
    unsigned r;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// round half-to-even and saturate to int16 in one conversion (libswresample: lrintf + av_clip_int16)
__device__ __forceinline__ int quant_s16_sat(float v) {
    short q;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(q) : "f"(v));
    return (int)q;
}
__device__ __forceinline__ unsigned long long clk64() { return (unsigned long long)clock64(); }
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
// contiguous L2 prefetch by the TMA unit (no shared-memory destination)
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
#else
static float g_emu_tmem[128][512];
static inline void fence_proxy_async() {}
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}
static inline void tmem_alloc(saddr_t slot, unsigned) { *(unsigned*)slot = 0u; }
static inline void tmem_dealloc(unsigned, unsigned) {}
static inline void umma_f16(unsigned d_tmem, saddr_t a, unsigned a_lbo, unsigned a_sbo, saddr_t b, unsigned b_lbo, unsigned b_sbo,
                            unsigned idesc, unsigned accumulate) {
    const int n_dim = (int)((idesc >> 17) & 0x3fu) * 8, col0 = (int)(d_tmem & 0xffffu);
    for (int row = 0; row < 128; row++)
        for (int n = 0; n < n_dim; n++) {
            float acc = accumulate ? g_emu_tmem[row][col0 + n] : 0.0f;
            for (int k = 0; k < 16; k++) {
                const unsigned short av = *(const unsigned short*)(a + (size_t)(k / 8) * a_lbo + (size_t)(row / 8) * a_sbo + (row % 8) * 16 + (k % 8) * 2);
                const unsigned short bv = *(const unsigned short*)(b + (size_t)(k / 8) * b_lbo + (size_t)(n / 8) * b_sbo + (n % 8) * 16 + (k % 8) * 2);
                acc += emu::f16_to_f32(av) * emu::f16_to_f32(bv);
            }
            g_emu_tmem[row][col0 + n] = acc;
        }
}
// emulation: the descriptor's 14-bit start field cannot hold a host pointer, so the address travels in full in a side table
static saddr_t g_emu_desc_base = 0;
static inline void desc_origin(const void* smem_base) { g_emu_desc_base = (saddr_t)smem_base; }
static inline unsigned desc_start(saddr_t addr) { return (unsigned)(((addr - g_emu_desc_base) >> 4) & 0x3fffu); }
static inline void umma_f16_desc(unsigned d_tmem, unsigned a_lo, unsigned a_hi, unsigned b_lo, unsigned b_hi, unsigned idesc, unsigned accumulate) {
    const saddr_t a = g_emu_desc_base + ((saddr_t)(a_lo & 0x3fffu) << 4), b = g_emu_desc_base + ((saddr_t)(b_lo & 0x3fffu) << 4);
    umma_f16(d_tmem, a, ((a_lo >> 16) & 0x3fffu) << 4, (a_hi & 0x3fffu) << 4, b, ((b_lo >> 16) & 0x3fffu) << 4, (b_hi & 0x3fffu) << 4, idesc, accumulate);
}
static inline void umma_commit(saddr_t bar) { mbar_arrive(bar); }
static inline void umma_f16_desc_warp(unsigned d_tmem, unsigned a_lo, unsigned a_hi, unsigned b_lo, unsigned b_hi, unsigned idesc, unsigned accumulate) {
    if (emu_lane() == 0) umma_f16_desc(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
    __syncwarp();
}
static inline void umma_commit_warp(saddr_t bar) { if (emu_lane() == 0) mbar_arrive(bar); __syncwarp(); }
static inline void tmem_ld32(unsigned taddr, unsigned (&r)[32]) {
    const int lane0 = (int)(taddr >> 16), col0 = (int)(taddr & 0xffffu);
    for (int j = 0; j < 32; j++) r[j] = __float_as_uint(g_emu_tmem[lane0 + emu_lane()][col0 + j]);
}
static inline unsigned hfma2_bits(unsigned a, unsigned b, unsigned c) {
    auto one = [](unsigned short x, unsigned short y, unsigned short z) {
        return (unsigned)emu::f32_to_f16((float)((double)emu::f16_to_f32(x) * (double)emu::f16_to_f32(y) + (double)emu::f16_to_f32(z)));
    };
    return one((unsigned short)(a & 0xffff), (unsigned short)(b & 0xffff), (unsigned short)(c & 0xffff)) |
           (one((unsigned short)(a >> 16), (unsigned short)(b >> 16), (unsigned short)(c >> 16)) << 16);
}
static inline int quant_s16_sat(float v) { return quant_s16(v); }
static inline unsigned long long clk64() { return 0ull; }
template <int N> static inline void reg_dealloc() {}
template <int N> static inline void reg_alloc() {}
static inline void l2_prefetch(const void*, unsigned) {}
static inline uint4 ldg_stream(const uint4* p) { return *p; }
#endif

// ---- MMA issue schedule (compile time) ---------------------------------------------------------------------------
// The k-steps of a tile are issued in the order their columns arrive, not block by block: k-step (b, s) reads the chunk
// pair kbp(b) / 8 + 2 s, +1 and can go as soon as the piece holding its second chunk is converted.  A ring piece is
// therefore released right after the last k-step that reads it, and the converters run ahead of the tensor core instead
// of waiting for whole 144-column block windows.  The schedule is a constant expression and the issuer's loop over it is
// fully unrolled: a table-driven issuer needs ~150 instructions per k-step (decode, descriptor arithmetic, moves to
// uniform registers) and, sharing its scheduler with five other warps, ~1100 cycles per k-step — twice the HBM time of
// a tile for the 90 k-steps (profiles/r01_fir_umma.md).  Unrolled, a k-step is its two tcgen05.mma and three
// uniform-datapath instructions.
template <int IN_RATE>
constexpr FirUmmaSched fir_umma_schedule() {
    using G = FirUmmaGeom<IN_RATE>;
    FirUmmaSched out{};
    int ib[kFuSchedMax] = {}, is[kFuSchedMax] = {}, ineed[kFuSchedMax] = {};
    int n = 0;
    for (int p = 0; p < G::PIECES; p++)
        for (int pass = 0; pass < 2; pass++)                       // k-steps that still read the previous piece first
            for (int b = 0; b < kFmBlocks; b++)
                for (int s = 0; s < G::KS; s++) {
                    const int c0 = G::kbp(b) / 8 + 2 * s, p0 = c0 / 8, p1 = (c0 + 1) / 8;
                    if (p1 != p || (pass == 0) != (p0 < p1)) continue;
                    ib[n] = b; is[n] = s; ineed[n] = p1;
                    n++;
                }
    int last_reader[16] = {};
    for (int q = 0; q < 16; q++) last_reader[q] = -1;
    for (int j = 0; j < n; j++) {
        const int c0 = G::kbp(ib[j]) / 8 + 2 * is[j];
        last_reader[c0 / 8] = j;
        last_reader[(c0 + 1) / 8] = j;
    }
    out.n = n;
    int freed = 0;
    for (int j = 0; j < n; j++) {
        unsigned word = (unsigned)(G::kbp(ib[j]) / 8 + 2 * is[j]) | ((unsigned)((G::SHARED_B ? 0 : ib[j]) * G::KS + is[j]) << 7) | ((unsigned)ib[j] << 14);
        if (is[j] == 0) word |= kFuItFirst;
        if (is[j] == G::KS - 1) word |= kFuItLast;
        unsigned frees = 0;
        // pieces are released in order; the last item releases whatever is left of the tile
        while (freed < G::PIECES && (last_reader[freed] <= j || j == n - 1)) { freed++; frees++; }
        word |= (unsigned)(ineed[j] + 1) << 20;
        word |= frees << 24;
        out.w[j] = word;
    }
    return out;
}
template <int IN_RATE> struct FirUmmaSchedOf { static constexpr FirUmmaSched value = fir_umma_schedule<IN_RATE>(); };

// issuer state (all warp-uniform)
struct FirUmmaIssue {
    saddr_t pf0, pe0, df0, de0;       // barrier arrays
    unsigned a_lo0, a_hi, b_lo0, b_hi, tmem;
    unsigned piece0, chunk_base, blk_base;
    int phases;
};

// k-step J of the schedule: every field is a compile-time constant, only the ring rotation and the barrier parities
// depend on the tile counter
template <int IN_RATE, int J>
__device__ __forceinline__ void fir_umma_issue_item(const FirUmmaIssue& c) {
    using G = FirUmmaGeom<IN_RATE>;
    constexpr unsigned item = FirUmmaSchedOf<IN_RATE>::value.w[J];
    constexpr unsigned need = (item >> 20) & 15u, prev_need = J == 0 ? 0u : ((FirUmmaSchedOf<IN_RATE>::value.w[J == 0 ? 0 : J - 1] >> 20) & 15u);
    constexpr unsigned b = (item >> 14) & 15u, cidx = item & 127u, bidx = (item >> 7) & 127u, frees = (item >> 24) & 15u;
    constexpr bool first = (item & kFuItFirst) != 0, last = (item & kFuItLast) != 0;
    // pieces released before this item (compile time): sum of the frees of the items before it
    if constexpr (need > prev_need) {
#pragma unroll
        for (unsigned p = prev_need; p < need; p++) {
            const unsigned P = c.piece0 + p;
            mbar_wait(c.pf0 + (P % G::RING_PIECES) * kFmBarBytes, (P / G::RING_PIECES) & 1u);
        }
        fence_proxy_async();              // the converters' generic-proxy stores -> visible to the tensor core's async-proxy reads
        tc_fence_after();
    }
    const unsigned n = c.blk_base + b;                                  // accumulator counter of the block
    const unsigned ds = n % kFuDSlots;
    if constexpr (first) { mbar_wait(c.de0 + ds * kFmBarBytes, ((n / kFuDSlots) & 1u) ^ 1u); tc_fence_after(); }   // passes on first use
    const unsigned d_tmem = c.tmem + ds * kFuDCols;
    // the last ring chunk pairs with the mirror of chunk 0 stored right behind it
    const unsigned a_lo = c.a_lo0 + ((c.chunk_base + cidx) % G::RING_CHUNKS) * (unsigned)(kFuLbo >> 4);
    const unsigned b_lo = c.b_lo0 + bidx * (unsigned)(kFuBTile >> 4);
    umma_f16_desc_warp(d_tmem, a_lo, c.a_hi, b_lo, c.b_hi, kFuIdesc, first ? 0u : 1u);
    umma_f16_desc_warp(d_tmem, a_lo + (unsigned)(G::PLANE_BYTES >> 4), c.a_hi, b_lo, c.b_hi, kFuIdesc, 1u);
    if constexpr (last) umma_commit_warp(c.df0 + ds * kFmBarBytes);
    if constexpr (frees > 0) {
        // pieces released so far = pieces released by the items before this one
        constexpr unsigned freed0 = [] { unsigned f = 0; for (int k = 0; k < J; k++) f += (FirUmmaSchedOf<IN_RATE>::value.w[k] >> 24) & 15u; return f; }();
        static_assert(freed0 + frees <= need || J == FirUmmaSchedOf<IN_RATE>::value.n - 1, "a piece is awaited before it is released");
#pragma unroll
        for (unsigned f = 0; f < frees; f++) {
            const unsigned P = c.piece0 + freed0 + f;
            if (freed0 + f >= need) mbar_wait(c.pf0 + (P % G::RING_PIECES) * kFmBarBytes, (P / G::RING_PIECES) & 1u);   // (last item only: junk pieces nobody reads)
            umma_commit_warp(c.pe0 + (P % G::RING_PIECES) * kFmBarBytes);
        }
    }
}
template <int IN_RATE, int... Js>
__device__ __forceinline__ void fir_umma_issue_tile(const FirUmmaIssue& c, std::integer_sequence<int, Js...>) {
    (fir_umma_issue_item<IN_RATE, Js>(c), ...);
}

// ---- converter role ------------------------------------------------------------------------------------------
// Warp cw, row pair i, half-warp h: row 2 (kFuCvtWarps i + cw) + h; lane q (of 16) converts the piece's columns 4q..4q+3 = one
// aligned quad of the input.  Two register sets alternate; every register quad is refilled with the piece two ahead as
// soon as it is consumed.
template <int IN_RATE>
__device__ __forceinline__ void fir_umma_convert(const uint4* __restrict__ in_q, unsigned char* ring, saddr_t pf0, saddr_t pe0, int cw, int lane,
                                                 i64 span_stride_q, int total, unsigned long long* trace, int phases, const unsigned char* pf_base) {
    using G = FirUmmaGeom<IN_RATE>;
    constexpr int ROWQ = 2 * kFuCvtWarps * G::ROWQ;           // quads between the rows of consecutive row pairs
    constexpr unsigned KHV = 65536u + (0x6400u << 7);         // dp2a bias: (u >> 7) = f16 bits of 1024 + (hv + 512), u & 127 = lo
    if (total == 0) return;
    const int h = lane >> 4, q = lane & 15;
    const int row0 = 2 * cw + h;
    const uint4* tbase = in_q + (i64)row0 * G::ROWQ + q;      // this lane's quad of piece 0 of the CTA's first tile
    const uint4* lp = tbase;                                  // load cursor: piece l_P (= piece being converted + 2 in the steady state)
    int l_p = 0, l_P = 0;
    auto advance_load = [&]() {
        l_P++;
        lp += kFuPiece / 4;
        if (++l_p == G::PIECES) { l_p = 0; tbase += span_stride_q; lp = tbase; }
    };
    int c_P = 0, c_p = 0, slot = 0;
    unsigned pe_parity = 1;                                   // passes on first use of a slot
    // DRAM sees the column pieces as 256-byte reads 7 KB apart (a few percent of each DRAM page per visit: ~2 TB/s at
    // best).  One lane per CTA therefore asks the TMA unit to pull the CTA's contiguous quarter of the NEXT span into L2
    // while the current tile is converted; the piece loads then hit L2.
    constexpr unsigned kQuarter = (unsigned)(kFuSpan * G::S * 4 / kFuClasses);
    auto prefetch_quarter = [&](const unsigned char* q0) {
#pragma unroll 1
        for (unsigned off = 0; off < kQuarter; off += 32768u) l2_prefetch(q0 + off, kQuarter - off < 32768u ? kQuarter - off : 32768u);
    };
    if (pf_base) prefetch_quarter(pf_base);
    uint4 set[2][kFuRowPairs];
    auto step = [&](uint4 (&cur)[kFuRowPairs]) {
        if (pf_base && c_p == 0 && c_P + G::PIECES < total) prefetch_quarter(pf_base + (size_t)(c_P / G::PIECES + 1) * (size_t)span_stride_q * 16);
        if (trace && c_P < 320) trace[3 * c_P] = clk64();
        mbar_wait(pe0 + (unsigned)(slot * kFmBarBytes), pe_parity);      // MMAs done with the old contents of the slot
        if (trace && c_P < 320) trace[3 * c_P + 1] = clk64();
        unsigned char* dst0 = ring + (size_t)(slot * (kFuPiece / 8) + (q >> 1)) * kFuLbo + (q & 1) * 8 + row0 * 16;
        const bool refill = l_P < total;
#pragma unroll
        for (int i = 0; i < kFuRowPairs; i++) {
            const uint4 v = cur[i];
            const unsigned u0 = (unsigned)__dp2a_lo((int)v.x, 0x0101, (int)KHV), u1 = (unsigned)__dp2a_lo((int)v.y, 0x0101, (int)KHV);
            const unsigned u2 = (unsigned)__dp2a_lo((int)v.z, 0x0101, (int)KHV), u3 = (unsigned)__dp2a_lo((int)v.w, 0x0101, (int)KHV);
            const unsigned hva = hsub2_bits(((u1 >> 7) << 16) + (u0 >> 7), 0x66006600u);       // (1024 + hv + 512) - 1536
            const unsigned hvb = hsub2_bits(((u3 >> 7) << 16) + (u2 >> 7), 0x66006600u);
            const unsigned loa = hfma2_bits((((u1 & 127u) << 16) | (u0 & 127u)) | 0x64006400u, 0x20002000u, 0xC800C800u);   // (1024 + lo) / 128 - 8
            const unsigned lob = hfma2_bits((((u3 & 127u) << 16) | (u2 & 127u)) | 0x64006400u, 0x20002000u, 0xC800C800u);
            unsigned char* d = dst0 + i * (2 * kFuCvtWarps * 16);
            if (phases & 1) {
                *(uint2*)d = make_uint2(hva, hvb);
                *(uint2*)(d + G::PLANE_BYTES) = make_uint2(loa, lob);
            } else if (hva + hvb + loa + lob == 0x12345u) *(uint2*)d = make_uint2(hva, hvb);
            if (refill && (phases & 8)) cur[i] = ldg_stream(lp + i * ROWQ);                   // (the asm's memory clobber keeps it behind the stores)
        }
        if (slot == 0 && q < 2) {                                             // mirror of ring chunk 0 behind the last chunk
#pragma unroll
            for (int i = 0; i < kFuRowPairs; i++) {
                unsigned char* d = dst0 + i * (2 * kFuCvtWarps * 16);
                *(uint2*)(d + (size_t)G::RING_CHUNKS * kFuLbo) = *(const uint2*)d;
                *(uint2*)(d + (size_t)G::RING_CHUNKS * kFuLbo + G::PLANE_BYTES) = *(const uint2*)(d + G::PLANE_BYTES);
            }
        }
        advance_load();
        // no fence.proxy.async here: it compiles to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the membar waits for this
        // thread's outstanding global loads, i.e. it exposes the full latency of the prefetch every piece.  The
        // generic-proxy stores are released by the mbarrier arrive; the MMA issuer acquires it and runs the proxy fence
        // on its side, before its first tcgen05.mma on the piece.
        __syncwarp();
        if (lane == 0) mbar_arrive(pf0 + (unsigned)(slot * kFmBarBytes));
        if (trace && c_P < 320) trace[3 * c_P + 2] = clk64();
        c_P++;
        if (++slot == G::RING_PIECES) { slot = 0; pe_parity ^= 1u; }
        if (++c_p == G::PIECES) c_p = 0;
    };
#pragma unroll
    for (int k = 0; k < 2; k++) {
        if (l_P < total) {
#pragma unroll
            for (int i = 0; i < kFuRowPairs; i++) set[k][i] = ldg_stream(lp + i * ROWQ);
        }
        advance_load();
    }
    while (true) {
        step(set[0]);
        if (c_P >= total) break;
        step(set[1]);
        if (c_P >= total) break;
    }
}

template <int IN_RATE>
__global__ void __launch_bounds__(kFuThreads, 1) fir_umma_kernel(const FirUmmaArgs a) {
    using G = FirUmmaGeom<IN_RATE>;
    B2A_DYN_SMEM(smem);
    desc_origin(smem);
    unsigned char* ring = smem;                                   // [plane hv | plane lo][RING_CHUNKS + 1][kFuLbo]
    unsigned char* btab = smem + 2 * G::PLANE_BYTES;               // [BBLOCKS][KS][kFuBTile]
    const saddr_t bars = smem_addr(btab + G::B_BYTES);
    unsigned* tmem_slot = (unsigned*)(btab + G::B_BYTES + G::NBARS * kFmBarBytes);
    // barrier slots: PF piece full (converters -> MMA), PE piece free (MMA commit -> converters),
    //                DF accumulator full (MMA commit -> epilogue), DE accumulator drained (epilogue -> MMA)
    auto PF = [&](int s) { return bars + (unsigned)((0 + s) * kFmBarBytes); };
    auto PE = [&](int s) { return bars + (unsigned)((G::RING_PIECES + s) * kFmBarBytes); };
    auto DF = [&](int s) { return bars + (unsigned)((2 * G::RING_PIECES + s) * kFmBarBytes); };
    auto DE = [&](int s) { return bars + (unsigned)((2 * G::RING_PIECES + kFuDSlots + s) * kFmBarBytes); };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int cls = blockIdx.x & (kFuClasses - 1);              // this CTA's run class
    for (int i = tid; i < G::B_BYTES / 16; i += kFuThreads) ((uint4*)btab)[i] = a.btab[(size_t)cls * (G::B_BYTES / 16) + i];
    fence_proxy_async();                                          // the tensor core reads shared memory through the async proxy
    if (tid == 0) {
        for (int s = 0; s < G::RING_PIECES; s++) { mbar_init(PF(s), kFuCvtWarps); mbar_init(PE(s), 1); }
        for (int s = 0; s < kFuDSlots; s++) { mbar_init(DF(s), 1); mbar_init(DE(s), kFuEpiWarps); }
        mbar_fence_init();
    }
    if (warp == kFuMmaWarp) tmem_alloc(smem_addr(tmem_slot), kFuTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    const int span0 = blockIdx.x / kFuClasses, span_stride = gridDim.x / kFuClasses;     // gridDim.x is a multiple of 4
    const int n_tiles = span0 < a.spans ? (a.spans - span0 + span_stride - 1) / span_stride : 0;
    const i64 run_base = kFuRun0 + (i64)span0 * kFuSpan + cls;   // row r of the CTA's tile `it` is run run_base + it * span_stride * 512 + 4 r

    if (warp < kFuEpiWarps) {
        // ===================================== epilogue: thread = run =====================================
        reg_dealloc<kFuRegsEpi>();
        const int row = warp * 32 + lane;
        int n = 0;                                                // accumulator counter
        for (int it = 0; it < n_tiles; it++) {
            const i64 run = run_base + (i64)it * span_stride * kFuSpan + kFuClasses * row;
#pragma unroll 1
            for (int b = 0; b < kFmBlocks; b++, n++) {
                const int ds = n % kFuDSlots;
                mbar_wait(DF(ds), (unsigned)((n / kFuDSlots) & 1));
                if (a.trace && blockIdx.x == 0 && tid == 0 && n < 1000) a.trace[2048 + 4 * 2048 + n] = clk64();
                tc_fence_after();
                unsigned r[32];
                tmem_ld32(tmem + ((unsigned)(warp * 32) << 16) + (unsigned)(ds * kFuDCols), r);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(DE(ds));               // the accumulator may be overwritten
                if (!(a.phases & 4)) continue;
                unsigned w[8];
                u64 e = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int q0 = quant_s16_sat((__uint_as_float(r[2 * j]) + __uint_as_float(r[16 + 2 * j])) * (1.0f / 64.0f));
                    const int q1 = quant_s16_sat((__uint_as_float(r[2 * j + 1]) + __uint_as_float(r[16 + 2 * j + 1])) * (1.0f / 64.0f));
                    w[j] = (unsigned)(q0 & 0xffff) | ((unsigned)q1 << 16);
                    e += (u64)((unsigned)(q0 * q0) + (unsigned)(q1 * q1));    // two squares fit 32 bits (<= 2^31)
                }
                if (a.out_s16) {
                    uint4* dst = (uint4*)(a.out_s16 + run * kFmNout + 16 * b);
                    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
                    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
                }
                if (a.energy) a.energy[run * kFmBlocks + b] = e;  // the block's 16 outputs are exactly one millisecond
            }
        }
    } else if (warp < kFuCvtWarp0) {
        // ===================================== MMA issuer (one thread) =====================================
        reg_dealloc<kFuRegsMma>();
        if (warp == kFuMmaWarp) {
            // the whole warp runs the (fully unrolled) schedule in uniform control flow; one elected lane issues
            FirUmmaIssue c;
            c.pf0 = PF(0); c.pe0 = PE(0); c.df0 = DF(0); c.de0 = DE(0);
            // descriptor halves: start >> 4 | LBO >> 4 << 16 (low word), SBO >> 4 | version 1 << 14 (high word)
            c.a_lo0 = desc_start(smem_addr(ring)) | ((unsigned)(kFuLbo >> 4) << 16); c.a_hi = (unsigned)(kFuSbo >> 4) | (1u << 14);
            c.b_lo0 = desc_start(smem_addr(btab)) | ((unsigned)(kFuBLbo >> 4) << 16); c.b_hi = (unsigned)(kFuBSbo >> 4) | (1u << 14);
            c.tmem = tmem; c.phases = a.phases;
            c.piece0 = 0; c.blk_base = 0;
#pragma unroll 1
            for (int it = 0; it < n_tiles; it++, c.piece0 += G::PIECES, c.blk_base += kFmBlocks) {
                c.chunk_base = c.piece0 * (kFuPiece / 8);          // global chunk index of the tile's column 0
                fir_umma_issue_tile<IN_RATE>(c, std::make_integer_sequence<int, FirUmmaSchedOf<IN_RATE>::value.n>{});
            }
        }
        __syncwarp();
    } else {
        // ============================ converters: global -> registers -> planes ============================
        reg_alloc<kFuRegsCvt>();
        const int cw = warp - kFuCvtWarp0;
        // first quad of row 0 of the CTA's first tile: the class shift makes it land on the 16-byte grid
        const i64 f0 = run_base * G::S - G::CENTER - G::shift(kFuRun0, cls);
        fir_umma_convert<IN_RATE>((const uint4*)a.in + (f0 >> 2), ring, PF(0), PE(0), cw, lane, (i64)span_stride * G::SPANQ, n_tiles * G::PIECES,
                                  (a.trace && blockIdx.x == 0 && lane == 0 && (cw == 0 || cw == kFuCvtWarps - 1)) ? a.trace + (cw == 0 ? 0 : 1024) : nullptr, a.phases,
                                  // the CTA's quarter of its first span (16-byte aligned: the span starts on a quad of class 0 or just before)
                                  (cw == 0 && lane == 0 && (a.phases & 16)) ? a.in + (((run_base - cls) * G::S - G::CENTER - 3) >> 2 << 2) * 4 + (size_t)cls * (kFuSpan * G::S * 4 / kFuClasses) : nullptr);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kFuMmaWarp) tmem_dealloc(tmem, kFuTmemCols);
}

// ---- host side ---------------------------------------------------------------------------------------------
const uint4* get_fir_umma_table(int in_rate);   // device table for the current device (b2a_host.cu); nullptr + error on failure

template <int IN_RATE>
static inline int fir_umma_launch(const void* d_in, i64 n_in, int16_t* d_out_s16, u64* d_energy, FirMmaPlan* plan, cudaStream_t stream) {
    using G = FirUmmaGeom<IN_RATE>;
    plan->out_lo = plan->out_hi = 0;
    // span t = runs [16 + 512 t, +512).  A row reads whole pieces from its (shifted) first quad: frames up to
    // S run - CENTER + 64 PIECES - 1 of the span's last run must exist.
    const i64 last_need = -(i64)G::CENTER + (i64)kFuPiece * G::PIECES;                 // relative to S * run of the last row
    const i64 room = n_in - last_need - (i64)G::S * (kFuRun0 - 1);
    const i64 spans = room >= (i64)G::S * kFuSpan ? room / ((i64)G::S * kFuSpan) : 0;
    if (spans <= 0) return 0;
    const uint4* tab = get_fir_umma_table(IN_RATE);
    if (!tab) return B2A_ECUDA;
    auto k = fir_umma_kernel<IN_RATE>;
    static unsigned long long attr_mask = 0;                     // per-device opt-in, see fir_mma_launch
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !((attr_mask >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fir_umma_kernel)");
        if (dev >= 0 && dev < 64) attr_mask |= 1ull << dev;
    }
    FirUmmaArgs a;
    a.in = (const unsigned char*)d_in; a.out_s16 = d_out_s16; a.energy = d_energy; a.btab = tab;
    a.spans = (int)spans;
    a.phases = 31;
    if (const char* ph = getenv("B2A_FIR_PHASES")) a.phases = atoi(ph);          // profiling only
    a.trace = nullptr;
    if (const char* tr = getenv("B2A_FIR_TRACE")) a.trace = (unsigned long long*)strtoull(tr, nullptr, 0);   // device pointer of >= 96 KB, profiling only
    i64 lanes = spans < 37 ? spans : 37;                         // persistent: 4 CTAs (one per class) per span lane, 148 SMs
    if (const char* gs = getenv("B2A_FIR_GRID")) {               // test knob: few CTAs => many tiles per CTA
        const int gv = (atoi(gs) + kFuClasses - 1) / kFuClasses;
        if (gv > 0 && gv < lanes) lanes = gv;
    }
    B2A_LAUNCH(k, (unsigned)(lanes * kFuClasses), kFuThreads, G::SMEM_BYTES, stream, a);
    B2A_CHECK_LAUNCH("fir_umma_kernel");
    plan->out_lo = (i64)kFuRun0 * kFmNout;
    plan->out_hi = ((i64)kFuRun0 + spans * kFuSpan) * kFmNout;
    return 1;
}

}  // namespace b2a
