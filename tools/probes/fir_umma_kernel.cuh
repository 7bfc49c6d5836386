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
//
// SUPERSEDED, NOT BUILT INTO libb2a.so: kept as the record behind profiles/r01_fir_umma.md.  The product kernel is
// audio_processor_b200/csrc/fir_tmem.cuh; the shared geometry / schedule / PTX wrappers live in csrc/fir_tc_common.cuh.
#pragma once
#include "../../audio_processor_b200/csrc/fir_tc_common.cuh"

namespace b2a {

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
