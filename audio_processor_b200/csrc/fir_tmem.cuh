// fir_tmem.cuh — the tcgen05 FIR fed by TMA, with the f16 operand planes in TENSOR MEMORY (sm_100a).
//
// Mathematics, class tiles, filter banks and the compile-time k-step schedule are in fir_tc_common.cuh (read that header
// first).  The data path is TMA + TMEM because profiles/r01_fir_umma.md showed that register-staged global loads cannot
// feed the kernel (L1 miss tracking caps the bytes in flight at ~14 KB per SM, ~2 TB/s):
//
//   global --2-D TMA (SWIZZLE_128B boxes of 128 rows x 32 frames)--> raw ring in shared memory (4 x 32 KB)
//          --converter warps, thread = row: LDS.128 -> (hv, lo / 128) f16 pairs -> tcgen05.st--> plane ring in TMEM
//          --tcgen05.mma, A operand from TMEM, B = filter bank in shared memory--> accumulators in TMEM --> epilogue
//
// The class tiles are what makes the TMA path possible: rows of a tile are 4 runs = 7056 bytes apart
// (a legal tensor-map stride) and start on a 16-byte quad.  With the planes in TMEM the shared memory holds only the raw
// ring and the 92 KB filter bank; there is no generic-proxy store the tensor core has to see (no proxy fence), and 128 KB
// of input is in flight per SM without a single register.
//
// TMEM map (512 columns allocated): accumulators [0, 32 DSLOTS); plane hv: 4 columns per 8-frame chunk, RING_CHUNKS
// chunks + a mirror of chunk 0 (so that a K = 16 operand never straddles the wrap); plane lo behind it.
// A-operand layout (pinned by tools/probes/umma_probe.cu, TS mode): lane = row, one column = two consecutive K elements
// (low half first); an operand may start on any 4-column boundary.
//
// Roles (512 threads), coupled only by mbarriers:
//   warps 0-3   epilogue: warp q owns TMEM lanes 32q..32q+31 (thread = run): tcgen05.ld 32 columns, add the two halves,
//               round half-to-even + saturate to s16 (swr audioconvert), two 16-byte stores, and the exact uint64 sum of
//               squares of the millisecond for the silence detector;
//   warp  4     MMA issuer (compile-time schedule, elected issue) + TMEM allocation;
//   warp  5     TMA producer: lane 0 keeps the raw ring full (two boxes per 64-column piece);
//   warps 6-7   the clip's edge outputs (reflect / symmetric extension at the head, the tail behind the last span),
//               table-driven, one thread per output, spread over all CTAs: no separate edge kernel on the stream;
//   warps 8-15  converters: warp w converts lane quadrant w % 4 (rows 32 (w % 4) + lane), frames 32 h .. 32 h + 31 of the
//               piece with h = (w - 8) / 4: eight swizzle-aware LDS.128, 16 + 16 packed words (dp2a + PRMT), two tcgen05.st.
//
// Measured (profiles/r01_ncu_fir_tmem.txt, one-hour 44.1 kHz stereo clip): 147-153 us, DRAM 636 MB read + 127 MB written =
// the algorithmic bytes, 5.1-5.3 TB/s = 78-81 % of the measured HBM copy peak (the mma.sync kernel: 236 us).
#pragma once
#include "fir_tc_common.cuh"
#include "resample_generic.cuh"

#ifndef B2A_EMU
#include <cuda.h>
#endif

namespace b2a {

constexpr int kFtRawSlots = 4;             // raw ring depth (pieces of 128 rows x 64 frames x 4 B = 32 KB)
constexpr int kFtBoxFrames = 32;           // frames per TMA box row (128 bytes: one SWIZZLE_128B atom row)
constexpr int kFtBoxBytes = kFuRT * kFtBoxFrames * 4;            // 16 KB
constexpr int kFtPieceBytes = 2 * kFtBoxBytes;                   // 32 KB
constexpr int kFtCvtWarps = 8;
constexpr int kFtThreads = (8 + kFtCvtWarps) * 32;
constexpr int kFtMmaWarp = 4, kFtTmaWarp = 5, kFtCvtWarp0 = 8;
constexpr int kFtTensorD0 = 2304;          // declared inner extent of the (overlapping-row) tensor, frames

template <int IN_RATE>
struct FirTmemGeom : FirUmmaGeom<IN_RATE> {
    using G = FirUmmaGeom<IN_RATE>;
    // plane ring in TMEM: a block window must leave one piece for the converters (4 pieces at 44.1 kHz, 5 at 48 kHz);
    // a 5-piece ring at 44.1 kHz (5 accumulator slots instead of 7) was measured: 160 -> 164 us, no gain
    static constexpr int A_PIECES = G::max_span() + 1 > 4 ? G::max_span() + 1 : 4;
    static constexpr int A_CHUNKS = A_PIECES * kFuPiece / 8;
    static constexpr int PLANE_COLS = A_CHUNKS * 4 + 4;                  // + mirror of chunk 0
    // accumulator slots: a power of two (the issuer computes the slot of every block; 4 measured 2 us faster than the 7 that fit)
    static constexpr int DSLOTS = 4;
    static constexpr int COL_HV = DSLOTS * kFuDCols, COL_LO = COL_HV + PLANE_COLS;
    static constexpr int NBARS = 2 * kFtRawSlots + 2 * A_PIECES + 2 * DSLOTS;
    static constexpr int SMEM_BYTES = 1024 + kFtRawSlots * kFtPieceBytes + G::B_BYTES + NBARS * kFmBarBytes + 16;
    // first frame of class c's rows relative to the row stride grid: row Rg (global 4-run row) of class c starts at frame
    // 4 S Rg + X(c) with X(c) = S (kFuRun0 + c) - CENTER - shift(c), a multiple of 4
    static constexpr int X(int c) { return G::S * (kFuRun0 + c) - G::CENTER - G::shift(kFuRun0, c); }
    static constexpr int XMIN = X(0);
    static_assert(DSLOTS >= 4 && COL_LO + PLANE_COLS <= 512, "TMEM budget");
    static_assert(SMEM_BYTES <= 232448, "shared-memory budget (227 KB per CTA)");
    static_assert(X(0) % 4 == 0 && X(1) % 4 == 0 && X(2) % 4 == 0 && X(3) % 4 == 0, "class rows start on a quad");
    static_assert(X(3) - XMIN + kFuPiece * G::PIECES <= kFtTensorD0, "declared tensor extent covers every box");
};

#ifdef B2A_EMU
struct CUtensorMap { const unsigned char* base; long long row_stride_bytes; };
#endif

struct FirTmemArgs {
    alignas(64) CUtensorMap tmap;   // uint32 frames, dims {kFtTensorD0, rows}, row stride 4 S frames, box {32, 128}, SWIZZLE_128B
    const unsigned char* in;        // the clip (diagnostics)
    int16_t* out_s16;               // nullable
    u64* energy;                    // nullable
    const uint4* btab;              // [class][B_BYTES] filter banks (build_fir_umma_table)
    int spans;
    int phases;                     // profiling aid (env B2A_FIR_PHASES): bit 0 conversion (LDS + split + tcgen05.st), bit 2 epilogue stores; 31 = the product
    GenericParams edge;             // the clip's head and tail outputs (table-driven, one thread per output): warps 6-7
    i64 edge_total;                 // linear indices of `edge` (resample_generic_total), 0 = none
};

// ---- primitives ---------------------------------------------------------------------------------------------------
#ifndef B2A_EMU
// one SWIZZLE_128B box [128 rows][32 frames] global -> shared, completion counted on the mbarrier
__device__ __forceinline__ void tma_load_box(saddr_t dst, const CUtensorMap* tm, int x, int y, saddr_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) { asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory"); }
__device__ __forceinline__ uint4 lds128(saddr_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
// 16 registers -> 16 TMEM columns of this warp's 32 lanes (thread = lane)
__device__ __forceinline__ void tmem_st16(unsigned taddr, const unsigned (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st4(unsigned taddr, unsigned r0, unsigned r1, unsigned r2, unsigned r3) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem], M = 128, K = 16; executed by a converged warp, one elected lane issues
__device__ __forceinline__ void umma_ts_warp(unsigned d_tmem, unsigned a_tmem, unsigned b_lo, unsigned b_hi, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\telect.sync _|e, 0xffffffff;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
#else
static inline void tma_load_box(saddr_t dst, const CUtensorMap* tm, int x, int y, saddr_t bar) {
    for (int r = 0; r < kFuRT; r++)
        for (int j = 0; j < 8; j++)
            memcpy((void*)(dst + (saddr_t)r * 128 + (saddr_t)((j ^ (r & 7)) * 16)), tm->base + (long long)(y + r) * tm->row_stride_bytes + (long long)x * 4 + j * 16, 16);
    unsigned* b = (unsigned*)bar;
    b[3] -= kFtBoxBytes;
    emu_mbar_try_complete(b);
}
static inline void tma_prefetch_desc(const CUtensorMap*) {}
static inline uint4 lds128(saddr_t a) { return *(const uint4*)a; }
static inline void tmem_st16(unsigned taddr, const unsigned (&r)[16]) {
    const int lane0 = (int)(taddr >> 16), col0 = (int)(taddr & 0xffffu);
    for (int j = 0; j < 16; j++) g_emu_tmem[lane0 + emu_lane()][col0 + j] = __uint_as_float(r[j]);
}
static inline void tmem_st4(unsigned taddr, unsigned r0, unsigned r1, unsigned r2, unsigned r3) {
    const int lane0 = (int)(taddr >> 16), col0 = (int)(taddr & 0xffffu);
    const unsigned r[4] = {r0, r1, r2, r3};
    for (int j = 0; j < 4; j++) g_emu_tmem[lane0 + emu_lane()][col0 + j] = __uint_as_float(r[j]);
}
static inline void tmem_st_wait() {}
static inline void umma_ts_warp(unsigned d_tmem, unsigned a_tmem, unsigned b_lo, unsigned b_hi, unsigned idesc, unsigned accumulate) {
    if (emu_lane() == 0) {
        const saddr_t b = g_emu_desc_base + ((saddr_t)(b_lo & 0x3fffu) << 4);
        const unsigned b_lbo = ((b_lo >> 16) & 0x3fffu) << 4, b_sbo = (b_hi & 0x3fffu) << 4;
        const int n_dim = (int)((idesc >> 17) & 0x3fu) * 8, col0 = (int)(d_tmem & 0xffffu), acol = (int)(a_tmem & 0xffffu);
        for (int row = 0; row < 128; row++)
            for (int n = 0; n < n_dim; n++) {
                double sum = 0.0;
                for (int k = 0; k < 16; k++) {
                    const unsigned aw = __float_as_uint(g_emu_tmem[row][acol + k / 2]);
                    const unsigned short av = (unsigned short)((k & 1) ? (aw >> 16) : (aw & 0xffffu));
                    const unsigned short bv = *(const unsigned short*)(b + (size_t)(k / 8) * b_lbo + (size_t)(n / 8) * b_sbo + (n % 8) * 16 + (k % 8) * 2);
                    sum += (double)emu::f16_to_f32(av) * (double)emu::f16_to_f32(bv);
                }
                g_emu_tmem[row][col0 + n] = emu_mma_accumulate(accumulate ? g_emu_tmem[row][col0 + n] : 0.0f, sum);
            }
    }
    __syncwarp();
}
#endif

// issuer state (warp-uniform)
struct FirTmemIssue {
    saddr_t pf0, pe0, df0, de0;
    unsigned b_lo0, b_hi, tmem;
    unsigned piece0, chunk_base, blk_base;
};

template <int IN_RATE, int J>
__device__ __forceinline__ void fir_tmem_issue_item(const FirTmemIssue& c) {
    using G = FirTmemGeom<IN_RATE>;
    constexpr unsigned item = FirUmmaSchedOf<IN_RATE>::value.w[J];
    constexpr unsigned need = (item >> 20) & 15u, prev_need = J == 0 ? 0u : ((FirUmmaSchedOf<IN_RATE>::value.w[J == 0 ? 0 : J - 1] >> 20) & 15u);
    constexpr unsigned b = (item >> 14) & 15u, cidx = item & 127u, bidx = (item >> 7) & 127u, frees = (item >> 24) & 15u;
    constexpr bool first = (item & kFuItFirst) != 0, last = (item & kFuItLast) != 0;
    if constexpr (need > prev_need) {
#pragma unroll
        for (unsigned p = prev_need; p < need; p++) {
            const unsigned P = c.piece0 + p;
            mbar_wait(c.pf0 + (P % G::A_PIECES) * kFmBarBytes, (P / G::A_PIECES) & 1u);
        }
        tc_fence_after();
    }
    const unsigned n = c.blk_base + b;                                  // accumulator counter of the block
    const unsigned ds = n % G::DSLOTS;
    if constexpr (first) { mbar_wait(c.de0 + ds * kFmBarBytes, ((n / G::DSLOTS) & 1u) ^ 1u); tc_fence_after(); }   // passes on first use
    const unsigned d_tmem = c.tmem + ds * kFuDCols;
    // the last ring chunk pairs with the mirror of chunk 0 stored right behind it
    const unsigned a_col = ((c.chunk_base + cidx) % G::A_CHUNKS) * 4u;
    const unsigned b_lo = c.b_lo0 + bidx * (unsigned)(kFuBTile >> 4);
    // (no profiling knob around these two: the issuer is the kernel's critical path — one extra runtime branch per k-step
    // cost 36 us, profiles/r01_probes.md)
    umma_ts_warp(d_tmem, c.tmem + G::COL_HV + a_col, b_lo, c.b_hi, kFuIdesc, first ? 0u : 1u);
    umma_ts_warp(d_tmem, c.tmem + G::COL_LO + a_col, b_lo, c.b_hi, kFuIdesc, 1u);
    if constexpr (last) umma_commit_warp(c.df0 + ds * kFmBarBytes);
    if constexpr (frees > 0) {
        constexpr unsigned freed0 = [] { unsigned f = 0; for (int k = 0; k < J; k++) f += (FirUmmaSchedOf<IN_RATE>::value.w[k] >> 24) & 15u; return f; }();
#pragma unroll
        for (unsigned f = 0; f < frees; f++) {
            const unsigned P = c.piece0 + freed0 + f;
            if (freed0 + f >= need) mbar_wait(c.pf0 + (P % G::A_PIECES) * kFmBarBytes, (P / G::A_PIECES) & 1u);   // (last item only: pieces nobody reads)
            umma_commit_warp(c.pe0 + (P % G::A_PIECES) * kFmBarBytes);
        }
    }
}
template <int IN_RATE, int... Js>
__device__ __forceinline__ void fir_tmem_issue_tile(const FirTmemIssue& c, std::integer_sequence<int, Js...>) {
    (fir_tmem_issue_item<IN_RATE, Js>(c), ...);
}

template <int IN_RATE>
__global__ void __launch_bounds__(kFtThreads, 1) fir_tmem_kernel(const __grid_constant__ FirTmemArgs a) {
    using G = FirTmemGeom<IN_RATE>;
    B2A_DYN_SMEM(smem_raw);
    // SWIZZLE_128B boxes want 1024-byte aligned shared memory
    unsigned char* smem = smem_raw + ((1024u - (unsigned)(smem_addr(smem_raw) & 1023u)) & 1023u);
    desc_origin(smem);
    unsigned char* raw = smem;                                            // [kFtRawSlots][2 boxes][128 rows][128 B]
    unsigned char* btab = smem + kFtRawSlots * kFtPieceBytes;             // [BBLOCKS][KS][kFuBTile]
    const saddr_t bars = smem_addr(btab + G::B_BYTES);
    unsigned* tmem_slot = (unsigned*)(btab + G::B_BYTES + G::NBARS * kFmBarBytes);
    // RF raw full (TMA), RE raw free (converters), PF plane piece full (converters), PE plane piece free (MMA commit),
    // DF accumulator full (MMA commit), DE accumulator drained (epilogue)
    auto RF = [&](int s) { return bars + (unsigned)(s * kFmBarBytes); };
    auto RE = [&](int s) { return bars + (unsigned)((kFtRawSlots + s) * kFmBarBytes); };
    auto PF = [&](int s) { return bars + (unsigned)((2 * kFtRawSlots + s) * kFmBarBytes); };
    auto PE = [&](int s) { return bars + (unsigned)((2 * kFtRawSlots + G::A_PIECES + s) * kFmBarBytes); };
    auto DF = [&](int s) { return bars + (unsigned)((2 * kFtRawSlots + 2 * G::A_PIECES + s) * kFmBarBytes); };
    auto DE = [&](int s) { return bars + (unsigned)((2 * kFtRawSlots + 2 * G::A_PIECES + G::DSLOTS + s) * kFmBarBytes); };
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    const int cls = blockIdx.x & (kFuClasses - 1);              // this CTA's run class
    for (int i = tid; i < G::B_BYTES / 16; i += kFtThreads) ((uint4*)btab)[i] = a.btab[(size_t)cls * (G::B_BYTES / 16) + i];
    fence_proxy_async();                                          // the tensor core reads the filter bank through the async proxy
    if (tid == 0) {
        for (int s = 0; s < kFtRawSlots; s++) { mbar_init(RF(s), 1); mbar_init(RE(s), kFtCvtWarps); }
        for (int s = 0; s < G::A_PIECES; s++) { mbar_init(PF(s), kFtCvtWarps); mbar_init(PE(s), 1); }
        for (int s = 0; s < G::DSLOTS; s++) { mbar_init(DF(s), 1); mbar_init(DE(s), kFuEpiWarps); }
        mbar_fence_init();
    }
    if (warp == kFtMmaWarp) tmem_alloc(smem_addr(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;

    const int span0 = blockIdx.x / kFuClasses, span_stride = gridDim.x / kFuClasses;     // gridDim.x is a multiple of 4
    const int n_tiles = span0 < a.spans ? (a.spans - span0 + span_stride - 1) / span_stride : 0;
    const i64 run_base = kFuRun0 + (i64)span0 * kFuSpan + cls;   // row r of the CTA's tile `it` is run run_base + it * span_stride * 512 + 4 r
    const int total = n_tiles * G::PIECES;

    if (warp < kFuEpiWarps) {
        // ===================================== epilogue: thread = run =====================================
        const int row = warp * 32 + lane;
        unsigned n = 0;                                           // accumulator counter
        for (int it = 0; it < n_tiles; it++) {
            const i64 run = run_base + (i64)it * span_stride * kFuSpan + kFuClasses * row;
#pragma unroll 1
            for (int b = 0; b < kFmBlocks; b++, n++) {
                const unsigned ds = n % G::DSLOTS;
                mbar_wait(DF(ds), (n / G::DSLOTS) & 1u);
                tc_fence_after();
                unsigned r[32];
                tmem_ld32(tmem + ((unsigned)(warp * 32) << 16) + ds * kFuDCols, r);
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
    } else if (warp == kFtMmaWarp) {
        // ===================================== MMA issuer (elected lane, uniform control flow) =====================================
        FirTmemIssue c;
        c.pf0 = PF(0); c.pe0 = PE(0); c.df0 = DF(0); c.de0 = DE(0);
        c.b_lo0 = desc_start(smem_addr(btab)) | ((unsigned)(kFuBLbo >> 4) << 16); c.b_hi = (unsigned)(kFuBSbo >> 4) | (1u << 14);
        c.tmem = tmem;
        c.piece0 = 0; c.blk_base = 0;
#pragma unroll 1
        for (int it = 0; it < n_tiles; it++, c.piece0 += G::PIECES, c.blk_base += kFmBlocks) {
            c.chunk_base = c.piece0 * (kFuPiece / 8);              // global chunk index of the tile's column 0
            fir_tmem_issue_tile<IN_RATE>(c, std::make_integer_sequence<int, FirUmmaSchedOf<IN_RATE>::value.n>{});
        }
    } else if (warp == kFtTmaWarp) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            tma_prefetch_desc(&a.tmap);
            const int x_cls = G::X(cls) - G::XMIN;                // class column offset inside the declared tensor rows
            int slot = 0, p = 0, y = span0 * kFuRT;               // tensor row of the tile's row 0
            unsigned parity = 1;                                  // passes on first use
            // (an L2 prefetch of the CTA's contiguous quarter of the next span, cp.async.bulk.prefetch.L2, was measured
            // here: 158 -> 199 us for the call — the boxes already stream at 5.3 TB/s and the prefetch only adds traffic)
            for (int P = 0; P < total; P++) {
                mbar_wait(RE(slot), parity);
                mbar_expect_tx(RF(slot), (unsigned)kFtPieceBytes);
                const saddr_t dst = smem_addr(raw) + (unsigned)(slot * kFtPieceBytes);
                tma_load_box(dst, &a.tmap, x_cls + kFuPiece * p, y, RF(slot));
                tma_load_box(dst + kFtBoxBytes, &a.tmap, x_cls + kFuPiece * p + kFtBoxFrames, y, RF(slot));
                if (++slot == kFtRawSlots) { slot = 0; parity ^= 1u; }
                if (++p == G::PIECES) { p = 0; y += span_stride * kFuRT; }
            }
        }
        __syncwarp();
    } else if (warp < kFtCvtWarp0) {
        // ===================================== edge outputs (warps 6-7 of every CTA) =====================================
        const i64 groups = (a.edge_total + 31) / 32;
        for (i64 g = (i64)blockIdx.x * 2 + (warp - 6); g < groups; g += (i64)gridDim.x * 2) resample_generic_output(a.edge, g * 32 + lane);
    } else {
        // ===================================== converters: thread = row =====================================
        // dp2a bias.  u = 2 (L + R) + KHV2 puts the f16 bits of 1024 + (hv + 512) into bytes 1-2 (hv = (L + R) >> 7), 2 lo into
        // byte 0 (lo = (L + R) & 127) and 0x64 into byte 3, so one PRMT per plane packs two frames: bytes (1, 2) of both words
        // are the hv pair, bytes (0, 3) are the f16 pair 1024 + 2 lo.  (With the fields at bit 7 the packing took shifts and
        // masks: 11 instead of 6 instructions per pair of frames, and the converters are 55 % of the kernel's instructions.)
        constexpr unsigned KHV2 = 0x64000000u + 2u * (65536u + (0x6400u << 7));
        const int cw = warp - kFtCvtWarp0, q = cw & 3, h = cw >> 2;
        const int row = 32 * q + lane;
        const unsigned t_lane = (unsigned)(32 * q) << 16;
        const saddr_t row_s = smem_addr(raw) + (unsigned)(h * kFtBoxBytes + row * 128);
        const unsigned sw = (unsigned)(row & 7);
        int rslot = 0, aslot = 0;
        unsigned rparity = 0, aparity = 1;                        // plane slots pass on first use
        for (int P = 0; P < total; P++) {
            mbar_wait(RF(rslot), rparity);                        // the piece's raw frames landed
            if (!(a.phases & 1)) {                                // profiling: barrier traffic only
                __syncwarp();
                if (lane == 0) mbar_arrive(RE(rslot));
                mbar_wait(PE(aslot), aparity);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(PF(aslot));
                if (++rslot == kFtRawSlots) { rslot = 0; rparity ^= 1u; }
                if (++aslot == G::A_PIECES) { aslot = 0; aparity ^= 1u; }
                continue;
            }
            uint4 v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = lds128(row_s + (unsigned)(rslot * kFtPieceBytes) + (((unsigned)j ^ sw) << 4));
            unsigned hv[16], lo[16];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const unsigned u0 = (unsigned)__dp2a_lo((int)v[j].x, 0x0202, (int)KHV2), u1 = (unsigned)__dp2a_lo((int)v[j].y, 0x0202, (int)KHV2);
                const unsigned u2 = (unsigned)__dp2a_lo((int)v[j].z, 0x0202, (int)KHV2), u3 = (unsigned)__dp2a_lo((int)v[j].w, 0x0202, (int)KHV2);
                hv[2 * j] = hsub2_bits(__byte_perm(u0, u1, 0x6521), 0x66006600u);             // (1024 + hv + 512) - 1536
                hv[2 * j + 1] = hsub2_bits(__byte_perm(u2, u3, 0x6521), 0x66006600u);
                lo[2 * j] = hfma2_bits(__byte_perm(u0, u1, 0x7430), 0x1C001C00u, 0xC400C400u);   // (1024 + 2 lo) / 256 - 4 = lo / 128
                lo[2 * j + 1] = hfma2_bits(__byte_perm(u2, u3, 0x7430), 0x1C001C00u, 0xC400C400u);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(RE(rslot));                // the raw slot may be refilled (its frames are in registers)
            mbar_wait(PE(aslot), aparity);                        // MMAs done with the old contents of the plane slot
            tc_fence_after();
            const unsigned col = (unsigned)(aslot * (kFuPiece / 2) + h * 16);
            tmem_st16(tmem + t_lane + G::COL_HV + col, hv);
            tmem_st16(tmem + t_lane + G::COL_LO + col, lo);
            if (aslot == 0 && h == 0) {                           // mirror of ring chunk 0 behind the last chunk
                tmem_st4(tmem + t_lane + G::COL_HV + G::A_CHUNKS * 4, hv[0], hv[1], hv[2], hv[3]);
                tmem_st4(tmem + t_lane + G::COL_LO + G::A_CHUNKS * 4, lo[0], lo[1], lo[2], lo[3]);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(PF(aslot));
            if (++rslot == kFtRawSlots) { rslot = 0; rparity ^= 1u; }
            if (++aslot == G::A_PIECES) { aslot = 0; aparity ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kFtMmaWarp) tmem_dealloc(tmem, 512);
}

// ---- host side ---------------------------------------------------------------------------------------------
#ifndef B2A_EMU
typedef CUresult (*b2a_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline b2a_tmap_encode_fn fir_tmem_encoder() {
    static b2a_tmap_encode_fn fn = [] {
        b2a_tmap_encode_fn f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return f;
    }();
    return fn;
}
#endif

// outputs the kernel would produce for a clip of n_in frames (no launch); returns the number of spans
template <int IN_RATE>
static inline i64 fir_tmem_plan(i64 n_in, FirMmaPlan* plan) {
    using G = FirTmemGeom<IN_RATE>;
    plan->out_lo = plan->out_hi = 0;
    // tensor rows are 4 runs apart and declared kFtTensorD0 frames long from frame XMIN on: the last declared row must end
    // inside the clip (which also covers every box: X(3) - XMIN + 64 PIECES <= kFtTensorD0)
    const i64 rows_fit = n_in >= (i64)G::XMIN + kFtTensorD0 ? (n_in - G::XMIN - kFtTensorD0) / ((i64)4 * G::S) + 1 : 0;
    const i64 spans = rows_fit / kFuRT;
    if (spans <= 0) return 0;
    plan->out_lo = (i64)kFuRun0 * kFmNout;
    plan->out_hi = ((i64)kFuRun0 + spans * kFuSpan) * kFmNout;
    return spans;
}

// edge: the outputs outside [plan->out_lo, plan->out_hi) (and outside whatever another kernel covers) that warps 6-7 compute
template <int IN_RATE>
static inline int fir_tmem_launch(const void* d_in, i64 n_in, int16_t* d_out_s16, u64* d_energy, FirMmaPlan* plan, const GenericParams* edge,
                                  cudaStream_t stream) {
    using G = FirTmemGeom<IN_RATE>;
    const i64 spans = fir_tmem_plan<IN_RATE>(n_in, plan);
    if (spans <= 0) return 0;
    const uint4* tab = get_fir_umma_table(IN_RATE);
    if (!tab) return B2A_ECUDA;
    auto k = fir_tmem_kernel<IN_RATE>;
    static unsigned long long attr_mask = 0;                     // per-device opt-in, see fir_mma_launch
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !((attr_mask >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_BYTES);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fir_tmem_kernel)");
        if (dev >= 0 && dev < 64) attr_mask |= 1ull << dev;
    }
    FirTmemArgs a;
#ifndef B2A_EMU
    b2a_tmap_encode_fn enc = fir_tmem_encoder();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return B2A_ECUDA; }
    const cuuint64_t dims[2] = {(cuuint64_t)kFtTensorD0, (cuuint64_t)(spans * kFuRT)};
    const cuuint64_t strides[1] = {(cuuint64_t)4 * G::S * 4};                       // bytes between tensor rows (4 runs)
    const cuuint32_t box[2] = {(cuuint32_t)kFtBoxFrames, (cuuint32_t)kFuRT}, estr[2] = {1, 1};
    CUresult r = enc(&a.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, (void*)((const unsigned char*)d_in + (size_t)G::XMIN * 4), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return B2A_ECUDA; }
#else
    a.tmap.base = (const unsigned char*)d_in + (size_t)G::XMIN * 4;
    a.tmap.row_stride_bytes = (long long)4 * G::S * 4;
#endif
    a.in = (const unsigned char*)d_in; a.out_s16 = d_out_s16; a.energy = d_energy; a.btab = tab;
    a.spans = (int)spans;
    a.phases = 31;
#ifdef B2A_PROFILE
    if (const char* ph = getenv("B2A_FIR_PHASES")) a.phases = atoi(ph);          // profiling builds only (-DB2A_PROFILE): masks kernel phases, WRONG output
#endif
    a.edge_total = 0;
    if (edge) { a.edge = *edge; a.edge_total = resample_generic_total(*edge); }
    else memset(&a.edge, 0, sizeof(a.edge));
    i64 lanes = spans < 37 ? spans : 37;                         // persistent: 4 CTAs (one per class) per span lane, 148 SMs
#if defined(B2A_PROFILE) || defined(B2A_EMU)
    if (const char* gs = getenv("B2A_FIR_GRID")) {               // emulation / profiling builds only: few CTAs => many tiles per CTA
        const int gv = (atoi(gs) + kFuClasses - 1) / kFuClasses;
        if (gv > 0 && gv < lanes) lanes = gv;
    }
#endif
    B2A_LAUNCH(k, (unsigned)(lanes * kFuClasses), kFtThreads, G::SMEM_BYTES, stream, a);
    B2A_CHECK_LAUNCH("fir_tmem_kernel");
    return 1;
}

}  // namespace b2a
