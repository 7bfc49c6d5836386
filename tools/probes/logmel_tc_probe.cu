// logmel_tc_probe.cu — PROBE (not on the product path): the Whisper STFT + mel projection of s16 PCM on tcgen05, in the
// form DESIGN.md §7.1 plans for round 2.  Driven by tools/probes/logmel_tc_probe.py (bank / mel table construction, the
// float64 check and the timing live there).
//
//   frame x[0..399] (hop 160, no window)  ->  four folded integer sequences, n = 0..111 (x[400] := x[0]):
//       ae = (x[n] + x[n+200]) + (x[200-n] + x[400-n])      even bins, cos        ao = (..) - (..)   even bins, sin
//       de = (x[n] - x[n+200]) - (x[200-n] - x[400-n])      odd bins, cos         do = (..) + (..)   odd bins, sin
//   planes hi = (v + 64) >> 7, lo = (v - 128 hi) / 128 (exact f16), written by thread = frame converter warps into a ring of
//   TMEM columns (TS-mode A operand); basis = f16(c) + f16(c - f16(c)) resident in shared memory (8 planes x 25 088 B);
//   D[128 frames x 112] (+)= A[128 x 16] B[16 x 112], 4 products x 7 k-steps x 4 plane pairs = 112 MMAs per tile into
//   4 x 112 accumulator columns; epilogue thread = frame: Hann as the 3-tap X[k]/2 - (X[k-1] + X[k+1])/4, power, two running
//   mel sums (a bin feeds at most two neighbouring slaney filters), log10, coalesced stores along T.
//
// Build: nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC -o _bin/liblogmel_tc_probe.so logmel_tc_probe.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <utility>

typedef unsigned saddr_t;
typedef long long i64;

namespace {

constexpr int kNP = 4;                         // products: even-cos, even-sin, odd-cos, odd-sin
constexpr int kKS = 7;                         // k-steps of 16 (K = 112 >= 101)
constexpr int kNB = 112;                       // N of every product (>= 101 bins)
constexpr int kPlaneBytes = 14 * 14 * 128;     // [k chunk 14][n group 14][8 rows][16 B]
constexpr int kBankBytes = kNP * 2 * kPlaneBytes;
constexpr unsigned kLbo = 14 * 128, kSbo = 128;
constexpr int kRingCol0 = kNP * kNB;           // 448: four slots of 16 columns (hi 8 + lo 8), slot = product
constexpr unsigned kIdesc = (1u << 4) | ((unsigned)(kNB >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
constexpr int kThreads = 16 * 32;              // warpgroups: 0-3 epilogue, 4-7 and 8-11 converters, 12 issuer (13-15 idle)
constexpr int kRegEpi = 104, kRegConv = 184, kRegIssue = 40;   // setmaxnreg budgets: 4 x 32 x (104 + 2 x 184 + 40) = 65536
static_assert(128 * (kRegEpi + 2 * kRegConv + kRegIssue) <= 65536, "role budgets exceed the launch allocation: setmaxnreg.inc would wait forever");
constexpr int kMels = 80;

#include "logmel_tc_mel80.inc"   // kW0 / kW1 / kEmit / kMEmit[201], kMelLast: compile-time, so the epilogue is branch-free straight-line code

__device__ __forceinline__ saddr_t smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(saddr_t bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(saddr_t bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(saddr_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(saddr_t bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_load(saddr_t dst, const void* src, unsigned bytes, saddr_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_ts_warp(unsigned d_tmem, unsigned a_tmem, unsigned b_lo, unsigned b_hi, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p, e;\n\t.reg .b64 db;\n\telect.sync _|e, 0xffffffff;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_warp(saddr_t bar) {
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, const unsigned (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&r)[16]) {
    unsigned u[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                   "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned pack_f16x2(float lo, float hi) {
    unsigned d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// packed f16 arithmetic on raw bits
__device__ __forceinline__ unsigned hadd2(unsigned a, unsigned b) { unsigned d; asm("add.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned hsub2(unsigned a, unsigned b) { unsigned d; asm("sub.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned hfma2(unsigned a, unsigned b, unsigned c) { unsigned d; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
// a pair of s16 samples x = 128 xh + xl -> (xh, xh') and (xl / 128, xl' / 128) as f16x2, exact: 0x6400 | n is the f16 1024 + n
__device__ __forceinline__ unsigned split_hi(unsigned w) { return hsub2((((w ^ 0x80008000u) >> 7) & 0x01ff01ffu) | 0x64006400u, 0x65006500u); }
__device__ __forceinline__ unsigned split_lo(unsigned w) { return hfma2((w & 0x007f007fu) | 0x64006400u, 0x20002000u, 0xC800C800u); }
template <int N> __device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ int s16_lo(unsigned w) { return (int)(short)(w & 0xffffu); }
__device__ __forceinline__ int s16_hi(unsigned w) { return (int)w >> 16; }

// ---- epilogue pieces: everything indexed by the bin is a template parameter, so the 201 bins become straight-line code ----
struct Epi {
    float pr, pi, cr, ci, acc0, acc1;     // X[k-1], X[k]; the open mel filter and the next one
    float* out;                           // &out[0][f]
    i64 T;
    bool live;
    float* dbg;                           // &dbg_power[row][0] or nullptr
};
__device__ __forceinline__ float log10_mel(float acc) { return 0.30102999566f * __log2f(fmaxf(acc * 1.52587890625e-5f, 1e-10f)); }
// bin K: cur = X[K], prev = X[K-1], next = (nr, ni)
template <int K>
__device__ __forceinline__ void epi_step(Epi& e, float nr, float ni) {
    const float wr = 0.5f * e.cr - 0.25f * (e.pr + nr), wi = 0.5f * e.ci - 0.25f * (e.pi + ni);
    const float pw = wr * wr + wi * wi;
    if (e.dbg) e.dbg[K] = pw;
    constexpr int n_emit = kEmit[K], m_emit = kMEmit[K];
    constexpr float w0 = kW0[K], w1 = kW1[K];
    static_assert(n_emit <= 2, "a bin closes at most two filters");
    if constexpr (n_emit >= 1) { if (e.live) e.out[(size_t)m_emit * (size_t)e.T] = log10_mel(e.acc0); }
    if constexpr (n_emit >= 2) { if (e.live) e.out[(size_t)(m_emit + 1) * (size_t)e.T] = log10_mel(e.acc1); }
    if constexpr (n_emit == 1) { e.acc0 = e.acc1; e.acc1 = 0.f; }
    if constexpr (n_emit == 2) { e.acc0 = 0.f; e.acc1 = 0.f; }
    if constexpr (w0 != 0.0f) e.acc0 = fmaf(w0, pw, e.acc0);
    if constexpr (w1 != 0.0f) e.acc1 = fmaf(w1, pw, e.acc1);
    e.pr = e.cr; e.pi = e.ci; e.cr = nr; e.ci = ni;
}
template <int C, int JJ>
__device__ __forceinline__ void epi_pair(Epi& e, const float (&re_e)[16], const float (&im_e)[16], const float (&re_o)[16], const float (&im_o)[16]) {
    constexpr int j = 16 * C + JJ;
    if constexpr (j == 0) { e.cr = re_e[0]; e.ci = im_e[0]; e.pr = re_o[0]; e.pi = -im_o[0]; }       // X[0]; X[-1] = conj X[1]
    else if constexpr (2 * j <= 200) epi_step<2 * j - 1>(e, re_e[JJ], im_e[JJ]);
    if constexpr (2 * j + 1 <= 199) epi_step<2 * j>(e, re_o[JJ], im_o[JJ]);
}
template <int C, int... JJ>
__device__ __forceinline__ void epi_round(Epi& e, const float (&re_e)[16], const float (&im_e)[16], const float (&re_o)[16], const float (&im_o)[16],
                                          std::integer_sequence<int, JJ...>) {
    (epi_pair<C, JJ>(e, re_e, im_e, re_o, im_o), ...);
}

struct Args {
    const int16_t* pcm;      // frame f = pcm[160 f .. 160 f + 399] (already reflect-padded by the caller), 32-byte aligned
    i64 n_frames;
    const unsigned char* bank;
    float* out;              // [80][n_frames] log10(mel power)
    float* dbg_power;        // optional [128][201] windowed power of tile 0
    int phases;              // bit 0: converter loads + ALU, bit 1: MMAs, bit 2: epilogue math + stores (timing knobs; all = 7);
                             // bit 3: the epilogue does only 3 of its 7 rounds (WRONG results: sizes a 12-warp epilogue)
};

__global__ void __launch_bounds__(kThreads, 1) logmel_tc_kernel(const Args a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* bank = smem;
    unsigned long long* bars = (unsigned long long*)(smem + kBankBytes);   // [0..3] full, [4..7] free, 8 acc_full, 9 acc_free, 10 bank
    unsigned* tmem_slot = (unsigned*)(bars + 12);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const saddr_t bar0 = smem_addr(bars);
    auto BAR = [&](int i) { return bar0 + 8u * (unsigned)i; };
    if (tid == 0) {
        for (int i = 0; i < 4; i++) { mbar_init(BAR(i), 4); mbar_init(BAR(4 + i), 1); }
        mbar_init(BAR(8), 1); mbar_init(BAR(9), 4); mbar_init(BAR(10), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 12) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tbase = *tmem_slot;
    const i64 n_tiles = (a.n_frames + 127) / 128;

    if (warp >= 12) {
        // ---------------- MMA issuer (warp 12; the rest of its warpgroup only gives its registers away) ----------------
        reg_dealloc<kRegIssue>();
        if (warp == 12) {
        if (lane == 0) {
            mbar_expect_tx(BAR(10), kBankBytes);
            for (int i = 0; i < kNP * 2; i++) bulk_load(smem_addr(bank) + i * kPlaneBytes, a.bank + (size_t)i * kPlaneBytes, kPlaneBytes, BAR(10));
        }
        mbar_wait(BAR(10), 0);
        const unsigned b_hi_word = (kSbo >> 4) | (1u << 14);
        unsigned g = 0, it = 0;
        for (i64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            if (it > 0) { mbar_wait(BAR(9), (it - 1) & 1); tc_fence_after(); }
#pragma unroll 1
            for (int s = 0; s < kKS; s++, g++) {
#pragma unroll
                for (int p = 0; p < kNP; p++) {
                    mbar_wait(BAR(p), g & 1);
                    tc_fence_after();
                    if (a.phases & 2) {
                        const unsigned d = tbase + p * kNB, ah = tbase + kRingCol0 + p * 16, al = ah + 8;
                        const unsigned bh = (((smem_addr(bank) + (p * 2 + 0) * kPlaneBytes + 2 * s * kLbo) >> 4) & 0x3fffu) | ((kLbo >> 4) << 16);
                        const unsigned bl = (((smem_addr(bank) + (p * 2 + 1) * kPlaneBytes + 2 * s * kLbo) >> 4) & 0x3fffu) | ((kLbo >> 4) << 16);
                        umma_ts_warp(d, ah, bh, b_hi_word, kIdesc, s > 0);
                        umma_ts_warp(d, ah, bl, b_hi_word, kIdesc, 1);
                        umma_ts_warp(d, al, bh, b_hi_word, kIdesc, 1);
                        umma_ts_warp(d, al, bl, b_hi_word, kIdesc, 1);
                    }
                    umma_commit_warp(BAR(4 + p));
                }
            }
            umma_commit_warp(BAR(8));
        }
        }
    } else if (warp >= 4) {
        reg_alloc<kRegConv>();
        // ---------------- converters: thread = frame; of the two warps of a lane quadrant one builds the even-bin products
        // (sums), the other the odd-bin products (differences); every warp touches its two ring slots at every k-step ----------------
        const int cw = warp - 4, quad = warp & 3, odd = cw >> 2;
        const int row = quad * 32 + lane;
        struct Raw { uint4 f1a, f1b, f2a, f2b, r1a, r1b, r2a, r2b; int r1x, r2x; };
        auto load_raw = [&](i64 tile, int s) {
            i64 f = tile * 128 + row;
            if (f >= a.n_frames) f = a.n_frames - 1;
            const int16_t* x = a.pcm + f * 160;
            const uint4* q = (const uint4*)x;
            // x[16s .. 16s+15], x[16s+200 ..], x[184-16s .. 199-16s], x[384-16s .. 399-16s]: all 16-byte aligned
            Raw r;
            r.f1a = q[2 * s]; r.f1b = q[2 * s + 1]; r.f2a = q[25 + 2 * s]; r.f2b = q[26 + 2 * s];
            r.r1a = q[23 - 2 * s]; r.r1b = q[24 - 2 * s]; r.r2a = q[48 - 2 * s]; r.r2b = q[49 - 2 * s];
            r.r1x = (int)x[200 - 16 * s]; r.r2x = (int)x[s == 0 ? 0 : 400 - 16 * s];
            return r;
        };
        unsigned g = 0;
        i64 tile = blockIdx.x;
        int s = 0;
        Raw cur;
        Raw zero{};
        if (tile < n_tiles) cur = (a.phases & 1) ? load_raw(tile, 0) : zero;
        while (tile < n_tiles) {
            i64 ntile = tile; int ns = s + 1;
            if (ns == kKS) { ns = 0; ntile += gridDim.x; }
            Raw nxt = cur;
            if (ntile < n_tiles && (a.phases & 1)) nxt = load_raw(ntile, ns);   // the next k-step's samples are in flight while this one converts
            const unsigned fw1[8] = {cur.f1a.x, cur.f1a.y, cur.f1a.z, cur.f1a.w, cur.f1b.x, cur.f1b.y, cur.f1b.z, cur.f1b.w};
            const unsigned fw2[8] = {cur.f2a.x, cur.f2a.y, cur.f2a.z, cur.f2a.w, cur.f2b.x, cur.f2b.y, cur.f2b.z, cur.f2b.w};
            const unsigned rw1[8] = {cur.r1a.x, cur.r1a.y, cur.r1a.z, cur.r1a.w, cur.r1b.x, cur.r1b.y, cur.r1b.z, cur.r1b.w};
            const unsigned rw2[8] = {cur.r2a.x, cur.r2a.y, cur.r2a.z, cur.r2a.w, cur.r2b.x, cur.r2b.y, cur.r2b.z, cur.r2b.w};
            // The planes are linear in the samples: split every sample once (x = 128 xh + xl), fold in packed f16.
            // sum = (x[n] +- x[n+200]) + (x[200-n] +- x[400-n]) is ae (even warp) or do (odd warp), diff is ao or de.
            unsigned ws_h[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ws_l[8] = {0, 0, 0, 0, 0, 0, 0, 0}, wd_h[8] = {0, 0, 0, 0, 0, 0, 0, 0}, wd_l[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (a.phases & 1) {
                const unsigned sgn = odd ? 0xBC00BC00u : 0x3C003C00u;        // (-1, -1) or (1, 1)
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    // reversed pair c = (block[16 - 2c], block[15 - 2c]); block[16] is the extra scalar
                    const unsigned q1 = __byte_perm(c == 0 ? (unsigned)cur.r1x : rw1[8 - (c == 0 ? 1 : c)], rw1[7 - c], 0x7610);
                    const unsigned q2 = __byte_perm(c == 0 ? (unsigned)cur.r2x : rw2[8 - (c == 0 ? 1 : c)], rw2[7 - c], 0x7610);
                    const unsigned uh = hfma2(split_hi(fw2[c]), sgn, split_hi(fw1[c])), ul = hfma2(split_lo(fw2[c]), sgn, split_lo(fw1[c]));
                    const unsigned vh = hfma2(split_hi(q2), sgn, split_hi(q1)), vl = hfma2(split_lo(q2), sgn, split_lo(q1));
                    ws_h[c] = hadd2(uh, vh); ws_l[c] = hadd2(ul, vl);
                    wd_h[c] = hsub2(uh, vh); wd_l[c] = hsub2(ul, vl);
                }
            }
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int p = 2 * odd + e;
                const bool use_sum = (e == 0) == (odd == 0);               // p0 = ae (sum), p1 = ao (diff), p2 = de (diff), p3 = do (sum)
                unsigned wh[8], wl[8];
#pragma unroll
                for (int c = 0; c < 8; c++) { wh[c] = use_sum ? ws_h[c] : wd_h[c]; wl[c] = use_sum ? ws_l[c] : wd_l[c]; }
                if (g > 0) { mbar_wait(BAR(4 + p), (g - 1) & 1); tc_fence_after(); }
                const unsigned t = tbase + ((unsigned)(quad * 32) << 16) + kRingCol0 + p * 16;
                tmem_st8(t, wh);
                tmem_st8(t + 8, wl);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(BAR(p));
            }
            g++;
            cur = nxt; tile = ntile; s = ns;
        }
    } else {
        // ---------------- epilogue: thread = frame ----------------
        reg_dealloc<kRegEpi>();
        const int row = warp * 32 + lane;
        unsigned it = 0;
        for (i64 tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, it++) {
            const i64 f = tile * 128 + row;
            const bool live = f < a.n_frames;
            mbar_wait(BAR(8), it & 1);
            tc_fence_after();
            Epi e{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, a.out + (live ? f : 0), a.n_frames, live && (a.phases & 4) != 0,
                  (a.dbg_power && tile == 0) ? a.dbg_power + row * 201 : nullptr};
            const unsigned t0 = tbase + ((unsigned)(warp * 32) << 16);
            auto round = [&](auto cc) {
                constexpr int c = decltype(cc)::value;
                float re_e[16], im_e[16], re_o[16], im_o[16];
                if ((a.phases & 8) && c >= 2 && c < kKS - 1) return;       // timing only: this warp's share if three warps per quadrant split the bins
                tmem_ld16(t0 + 0 * kNB + 16 * c, re_e);
                tmem_ld16(t0 + 1 * kNB + 16 * c, im_e);
                tmem_ld16(t0 + 2 * kNB + 16 * c, re_o);
                tmem_ld16(t0 + 3 * kNB + 16 * c, im_o);
                tmem_ld_wait();
                if (c == kKS - 1) {            // every accumulator column has been read: the next tile's MMAs may start
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR(9));
                }
                if (a.phases & 4) epi_round<c>(e, re_e, im_e, re_o, im_o, std::make_integer_sequence<int, 16>{});
            };
            round(std::integral_constant<int, 0>{}); round(std::integral_constant<int, 1>{}); round(std::integral_constant<int, 2>{});
            round(std::integral_constant<int, 3>{}); round(std::integral_constant<int, 4>{}); round(std::integral_constant<int, 5>{});
            round(std::integral_constant<int, 6>{});
            if (a.phases & 4) {
                epi_step<200>(e, e.pr, -e.pi);                                                          // X[201] = conj X[199]
                if (e.live) {
                    e.out[(size_t)kMelLast * (size_t)e.T] = log10_mel(e.acc0);
                    if (kMelLast + 1 < kMels) e.out[(size_t)(kMelLast + 1) * (size_t)e.T] = log10_mel(e.acc1);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

}  // namespace

extern "C" int lmtc_bank_bytes() { return kBankBytes; }

extern "C" int lmtc_run(const int16_t* d_pcm, long long n_frames, const unsigned char* d_bank, float* d_out, float* d_dbg_power, int phases,
                        int grid, void* stream) {
    Args a{d_pcm, n_frames, d_bank, d_out, d_dbg_power, phases};
    const size_t smem = kBankBytes + 256;
    cudaError_t e = cudaFuncSetAttribute(logmel_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { printf("cudaFuncSetAttribute: %s\n", cudaGetErrorString(e)); return -1; }
    logmel_tc_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(a);
    e = cudaGetLastError();
    if (e != cudaSuccess) { printf("launch: %s\n", cudaGetErrorString(e)); return -2; }
    return 0;
}
