// fir_tc_common.cuh — what the tcgen05 FIR kernels share: class-tile geometry, filter-bank layout, the compile-time k-step
// schedule and the PTX wrappers (tcgen05 / TMEM / mbarrier helpers).  Used by fir_tmem.cuh (the product kernel) and by the
// superseded register-staged variant kept under tools/probes/fir_umma_kernel.cuh.
//
// Replaces libswresample's swr_convert inner loop behind the reference's
//   ffmpeg -ar 16000 -ac 1 -c:a pcm_s16le   (app/services/audio_processor.py:912-923).
//
// Formulation (same banded-Toeplitz product and the same exact f16 splits as fir_mma.cuh, re-laid for tcgen05):
//   run   = 160 consecutive outputs = 10 blocks of 16 (one block = one millisecond at 16 kHz); a run consumes S input
//           frames (441 / 480), so 128 consecutive runs are 128 rows that share one filter matrix per block;
//   block b of 128 runs:  D[128 x 32] = X[128 x 16 KS] * [T_hi | T_lo][16 KS x 32]
//           X = (hv + lo / 128) with v = L + R = 128 hv + lo  (two f16 planes, both exact, accumulated into the same D),
//           T = 2^12 taps split T_hi + T_lo (f16 each); columns 0-15 hold the T_hi partial sums of the 16 outputs,
//           columns 16-31 the T_lo ones; the epilogue adds the two and scales by 2^-6.
//   => 2 KS tcgen05.mma (M128 N32 K16, kind::f16, f32 accumulate) per block, issued by one thread.
// Alignment.  A run is 441 frames = 1764 bytes, so consecutive runs sit at four different offsets against the 16-byte
// grid of the input.  A tile therefore takes every FOURTH run (class c = run % 4, rows 7056 bytes apart: 16-byte
// aligned), and its plane columns start SH_c frames before the filter origin so that every row of the tile begins on a
// quad; the shift is folded into the class's own copy of the filter bank (4 x 92 KB in global memory, one resident per
// CTA).  CTA b works on class b % 4, and CTAs b..b+3 walk the same 512-run spans together, so the rows' overlapping
// halos meet in L2.  Layout, descriptor fields and the TMEM accumulator layout (lane = row, column = n) are pinned on
// hardware by tools/probes/umma_probe.cu.
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
// the tensor core's f32 accumulator as the hardware probe saw it (tools/probes/umma_accum_probe.cu): the 16 products of a
// k-step are summed exactly, the running sum is rounded TOWARD ZERO
static inline float emu_mma_accumulate(float acc, double sum16) {
    const double v = (double)acc + sum16;
    float f = (float)v;
    if (std::fabs((double)f) > std::fabs(v)) f = std::nextafterf(f, 0.0f);
    return f;
}
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
            double sum = 0.0;
            for (int k = 0; k < 16; k++) {
                const unsigned short av = *(const unsigned short*)(a + (size_t)(k / 8) * a_lbo + (size_t)(row / 8) * a_sbo + (row % 8) * 16 + (k % 8) * 2);
                const unsigned short bv = *(const unsigned short*)(b + (size_t)(k / 8) * b_lbo + (size_t)(n / 8) * b_sbo + (n % 8) * 16 + (k % 8) * 2);
                sum += (double)emu::f16_to_f32(av) * (double)emu::f16_to_f32(bv);
            }
            g_emu_tmem[row][col0 + n] = emu_mma_accumulate(accumulate ? g_emu_tmem[row][col0 + n] : 0.0f, sum);
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

const uint4* get_fir_umma_table(int in_rate);   // device filter banks for the current device (b2a_host.cu); nullptr + error on failure

}  // namespace b2a
