// silence.cu — pydub-exact silence detection and stream compaction on sm_100a.
//
// Implements the step the reference intends at app/services/audio_processor.py:1046-1047
// ("音頻預處理 (移除靜音)"; preprocess_audio :305-314) with the semantics of pydub 0.25.1
// pydub/silence.py (detect_silence / detect_nonsilent / split_on_silence) — see
// oracle/pydub_silence.py for the literal restatement these kernels are tested against.
//
// Exact-integer formulation:
//   e[t]      = sum of x^2 over millisecond t (uint64; produced by the resampler epilogue or energy_ms_kernel)
//   silent(i) = sum_{t=i}^{i+W-1} e[t] < n_win * (floor(thr)+1)^2        (<=> audioop.rms(window) <= thr)
//   starts    = { i : i % step == 0 or i == last } with silent(i),  last = len_ms - W
//   ranges    = runs of starts with gaps <= W (or == step) -> [first, last_start + W)
//   nonsilent = complement; kept = nonsilent +- keep_silence, overlapping neighbours meet at the midpoint
//
// Kernels: silence_kernel (ONE launch for the windowed energy test, the coverage of the silent starts, the ordered range
// tables and split_on_silence's kept ranges; see the comment above it), compact_kernel (standalone stream compaction:
// gather kept milliseconds; binary search in the offset table; inside b2a_pipeline the log-mel tile loader gathers itself).
#include "b2a_common.cuh"

#include <cmath>

namespace b2a {

constexpr int SIL_THREADS = 1024;           // one block per SM
constexpr int SIL_WARPS = SIL_THREADS / 32;
constexpr int SIL_MIN_CHUNK = 4096;         // milliseconds per block, at least (short clips use fewer blocks)
constexpr int SIL_MAX_BLOCKS = 148;         // every block resident at once (the run-count exchange spins on its predecessors)
constexpr int SIL_MAX_CHUNKS = 8192;        // status words in the workspace (148 chunks of 24 576 ms cover an hour)
constexpr int SIL_SMEM_BUDGET = 231000;     // dynamic shared memory per block (227 KB less the static part)

struct SilenceCfg {
    i64 n_samples;      // F
    int spm;            // samples per millisecond
    i64 len_ms;         // pydub len(segment)
    i64 n_energy;       // entries in e[]
    int W;              // min_silence_len
    int step;           // seek_step
    i64 keep;           // keep_silence in ms (already resolved for keep_silence=True)
    i64 last;           // len_ms - W  (< 0: clip shorter than the window -> nothing is silent)
    u64 limit;          // n_win * (floor(thr)+1)^2
    int cap;
    int chunk;          // cover positions per chunk (multiple of 32)
    int n_chunks;
    int n_blocks;       // grid: min(n_chunks, SIL_MAX_BLOCKS)
};

// ---- per-ms energy of an existing s16 mono buffer ------------------------------------------
__global__ void __launch_bounds__(256) energy_ms_kernel(const int16_t* __restrict__ pcm, i64 n, int spm, i64 n_energy,
                                                        u64* __restrict__ e) {
    i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_energy) return;
    i64 base = t * spm;
    u64 acc = 0;
    for (int j = 0; j < spm; j++) {
        i64 idx = base + j;
        int v = idx < n ? (int)pcm[idx] : 0;
        acc += (u64)(unsigned)(v * v);
    }
    e[t] = acc;
}

// block-wide exclusive prefix over per-thread values (SIL_THREADS threads); returns exclusive prefix, total in *total
template <class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* s_warp /*[SIL_WARPS]*/, T* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    __syncthreads();             // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    // the 32 warp totals are scanned by every warp with shuffles (one shared-memory load per thread)
    T winc = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T y = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += y;
    }
    static_assert(SIL_WARPS == 32, "one warp total per lane");
    const T wprev = __shfl_sync(0xffffffffu, winc, (warp + 31) & 31);
    const T wpre = warp ? wprev : (T)0;
    *total = __shfl_sync(0xffffffffu, winc, 31);
    return wpre + inc - v;
}
// the same with max (values >= -2^30)
__device__ __forceinline__ int block_exclusive_max(int v, int* s_warp /*[SIL_WARPS]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kNone = -(1 << 30);
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = max(inc, y);
    }
    __syncthreads();
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int winc = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc = max(winc, y);
    }
    const int wprev = __shfl_sync(0xffffffffu, winc, (warp + 31) & 31);
    const int wpre = warp ? wprev : kNone;
    int ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = kNone;
    return max(wpre, ex);
}

// split_on_silence's range arithmetic + exclusive scan of kept lengths (in samples); one block
__device__ void silence_kept_pass(const int32_t* __restrict__ nonsilent_ms, const SilenceCfg& c, int32_t* __restrict__ kept_ms,
                                  i64* __restrict__ kept_off, i64* __restrict__ info, i64* s_w64, i64* s_carry_p) {
    const int tid = threadIdx.x;
    const int n = (int)info[B2A_INFO_N_NONSILENT];
    if (tid == 0) *s_carry_p = 0;
    __syncthreads();
    for (int base = 0; base < n || base == 0; base += SIL_THREADS) {
        const int k = base + tid;
        i64 len = 0;
        i64 s = 0, en = 0;
        if (k < n) {
            const i64 ns_s = nonsilent_ms[2 * k], ns_e = nonsilent_ms[2 * k + 1];
            s = ns_s - c.keep;
            en = ns_e + c.keep;
            if (k > 0) {               // boundary with the previous range
                const i64 last_end = (i64)nonsilent_ms[2 * (k - 1) + 1] + c.keep;
                if (s < last_end) s = (last_end + s) / 2;       // both operands non-negative in sum: floor == trunc
            }
            if (k + 1 < n) {           // boundary with the next range
                const i64 next_start = (i64)nonsilent_ms[2 * (k + 1)] - c.keep;
                if (next_start < en) en = (en + next_start) / 2;
            }
            if (s < 0) s = 0;
            if (en > c.len_ms) en = c.len_ms;
            if (en < s) en = s;
            len = (en - s) * c.spm;
            if (kept_ms) { kept_ms[2 * k] = (int32_t)s; kept_ms[2 * k + 1] = (int32_t)en; }
        }
        i64 tot;
        i64 pre = block_exclusive_scan<i64>(len, s_w64, &tot);
        const i64 carry = *s_carry_p;
        if (k < n) kept_off[k] = carry + pre;
        __syncthreads();
        if (tid == 0) *s_carry_p = carry + tot;
        __syncthreads();
        if (n == 0) break;
    }
    if (tid == 0) {
        kept_off[n] = *s_carry_p;
        info[B2A_INFO_N_KEPT] = n;
        info[B2A_INFO_N_KEEP] = *s_carry_p;
    }
}

// ---- the whole of pydub's detect_silence / detect_nonsilent / split_on_silence range logic in ONE launch -----------------
// The clip is cut into chunks of `chunk` cover positions (a multiple of 32 ms; one chunk per block for clips up to about
// an hour, otherwise block b takes chunks b, b + grid, ...).  For the chunk [c0, c0 + chunk) a block evaluates the window
// starts i in [c0 - W, c0 + chunk) it needs:
//   1. the energies e[c0 - Wr .. c0 + chunk + W) are staged in shared memory (coalesced 16-byte loads) and turned, in place,
//      into their exclusive prefix sums P (every thread owns an odd number of consecutive entries: conflict-free; one
//      block-wide scan of the per-thread totals);
//   2. E(i) = sum e[i .. i+W) = P[i + W] - P[i]: two shared-memory loads, a subtraction and a compare per start; the
//      silent-start flags of 32 starts are one ballot word in shared memory;
//   3. a start covers the W ms behind it: cover(t) <=> t - (last flagged start <= t) < W, i.e. a block-wide prefix MAXIMUM of
//      flag positions and a few bit operations per word of 32 ms;
//   4. run starts (covered <-> uncovered transitions) are counted per word; the chunks exchange their counts through a status
//      word each (blocks take chunks in increasing order and wait only for chunks of lower index, which are resident or done)
//      and scatter their run boundaries into the ordered range tables: a silent run's start is the end of the nonsilent
//      run before it, and vice versa;
//   5. the block that finishes last (ticket counter) does split_on_silence's keep_silence / midpoint / clamp arithmetic and
//      the exclusive scan of the kept lengths.
// Exact-integer throughout (uint64 energies; prefix sums wrap consistently mod 2^64).
struct SilenceWs {            // global workspace, zeroed on the stream before the launch
    unsigned int ticket;
    unsigned int pad;
    unsigned long long status[1];    // [n_chunks]: bit 63 valid | nonsilent starts << 31 | silent starts
};

__global__ void __launch_bounds__(SIL_THREADS, 1) silence_kernel(const u64* __restrict__ e, SilenceCfg c, int32_t* __restrict__ silent_ms,
                                                                 int32_t* __restrict__ nonsilent_ms, int32_t* __restrict__ kept_ms,
                                                                 i64* __restrict__ kept_off, i64* __restrict__ info, SilenceWs* __restrict__ ws) {
    B2A_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = c.W, Wr = (W + 31) & ~31;
    const int chunk = c.chunk;
    // every position of the clip fits 32 bits: index arithmetic is int, the energies are uint64
    const int len_ms = (int)c.len_ms, last = (int)c.last;
    const int n_valid = (int)(c.n_energy < c.len_ms ? c.n_energy : c.len_ms);
    const int n_fw = (chunk + Wr) / 32;                 // flag words: starts [f_lo, c0 + chunk)
    const int n_cw = chunk / 32;                        // cover words: positions [c0, c0 + chunk)
    const int n_P = chunk + Wr + W;                     // prefix sums over energies [f_lo, c0 + chunk + W): P[k] = sum of the first k
    u64* s_P = (u64*)smem_raw;                          // [n_P] (16-byte aligned)
    unsigned* s_flag = (unsigned*)(s_P + ((n_P + 1) & ~1));   // [n_fw]
    unsigned* s_cov = s_flag + n_fw;                    // [n_cw]
    int* s_last = (int*)(s_cov + n_cw);                 // [n_fw]: last flagged start (relative to f_lo) before word fw
    __shared__ i64 s_w64[SIL_WARPS];
    __shared__ int s_w32[SIL_WARPS];
    __shared__ i64 s_carry;
    __shared__ unsigned long long s_acc;                // packed run counts of the chunks this block had to wait for
    __shared__ int s_is_last;
    __shared__ unsigned s_prev_cov;                     // cover(c0 - 1), by the same rule (the word before the chunk)

    auto energy = [&](int t) -> u64 { return (unsigned)t < (unsigned)n_valid ? e[t] : 0ull; };
    const bool e_vec = (((uintptr_t)e) & 15) == 0;      // 16-byte loads of energy pairs

    i64 base_run = 0;                                   // packed run counts (nonsilent << 32 | silent) of all chunks before the current one
  for (int ck = blockIdx.x; ck < c.n_chunks; ck += gridDim.x) {
    if (tid == 0) s_acc = 0ull;                         // read after several barriers
    const int c0 = ck * chunk;                          // first cover position of this chunk
    const int f_lo = c0 - Wr;                           // first window start evaluated (32-aligned, <= c0 - W; may be negative)

    // ---- 1. energies -> shared memory -> exclusive prefix sums, in place ----
    for (int k0 = 2 * tid; k0 < n_P; k0 += 16 * SIL_THREADS) {
        ulonglong2 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int k = k0 + 2 * u * SIL_THREADS, t = f_lo + k;
            if (k < n_P && t >= 0 && t + 1 < n_valid && e_vec) v[u] = *(const ulonglong2*)(e + t);     // f_lo and k are even
            else { v[u].x = energy(t); v[u].y = energy(t + 1); }
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int k = k0 + 2 * u * SIL_THREADS;
            if (k < n_P) *(ulonglong2*)(s_P + k) = v[u];               // the array is padded to an even length
        }
    }
    __syncthreads();
    {
        const int per = ((n_P + SIL_THREADS - 1) / SIL_THREADS) | 1;   // odd: the 64-bit accesses of a half-warp hit distinct banks
        const int lo = min(tid * per, n_P), hi = min(lo + per, n_P);
        u64 sum = 0;
        for (int k = lo; k < hi; k++) sum += s_P[k];
        u64 tot;
        u64 run = block_exclusive_scan<u64>(sum, (u64*)s_w64, &tot);
        for (int k = lo; k < hi; k++) {
            const u64 x = s_P[k];
            s_P[k] = run;
            run += x;
        }
    }
    __syncthreads();

    // ---- 2. silent-start flags: E(i) = P[i + W] - P[i] ----
    for (int fw0 = warp; fw0 < n_fw; fw0 += 4 * SIL_WARPS) {
        u64 E[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int fw = fw0 + u * SIL_WARPS;
            const int k = 32 * (fw < n_fw ? fw : fw0) + lane;
            E[u] = s_P[k + W] - s_P[k];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int fw = fw0 + u * SIL_WARPS;
            if (fw >= n_fw) break;
            const int i = f_lo + 32 * fw + lane;                       // this lane's window start
            const bool cand = last >= 0 && (unsigned)i <= (unsigned)last && (c.step == 1 || (i % c.step) == 0 || i == last);
            const unsigned word = __ballot_sync(0xffffffffu, cand && E[u] < c.limit);
            if (lane == 0) s_flag[fw] = word;
        }
    }
    __syncthreads();

    // ---- 3. last flagged start before every flag word (exclusive prefix maximum), then the cover words ----
    const int kNone = -(1 << 30);
    {
        const int per = (n_fw + SIL_THREADS - 1) / SIL_THREADS;
        const int lo = min(tid * per, n_fw), hi = min(lo + per, n_fw);
        int mx = kNone;
        for (int fw = lo; fw < hi; fw++) {
            const unsigned f = s_flag[fw];
            if (f) mx = 32 * fw + 31 - __clz((int)f);
        }
        int run = block_exclusive_max(mx, s_w32);
        for (int fw = lo; fw < hi; fw++) {
            s_last[fw] = run;
            const unsigned f = s_flag[fw];
            if (f) run = 32 * fw + 31 - __clz((int)f);
        }
    }
    __syncthreads();
    auto window_silent = [&](int i) -> bool {            // direct evaluation (only for the seek_step > W continuity rule)
        u64 E = 0;
        for (int k = 0; k < W; k++) E += energy(i + k);
        return E < c.limit;
    };
    auto cover_word = [&](int cw) -> unsigned {          // cw >= -1: positions [c0 + 32 cw, +32)
        const int fw = cw + Wr / 32;
        const unsigned f = s_flag[fw];
        const int nb = s_last[fw] + W - 32 * fw;          // leading positions of the word still covered by an earlier start
        unsigned cov = nb >= 32 ? 0xffffffffu : nb <= 0 ? 0u : ((1u << nb) - 1u);
        if (W >= 32) {
            if (f) cov |= 0xffffffffu << (__ffs((int)f) - 1);   // a start covers the rest of its word
        } else {
            for (int k = 0; k < W; k++) cov |= f << k;
        }
        const int t0 = c0 + 32 * cw;
        if (c.step > W && cov != 0xffffffffu) {
            // pydub merges consecutive candidates p, p + step that are both silent even though no window covers the gap
            for (int b = 0; b < 32; b++) {
                const int t = t0 + b;
                if (((cov >> b) & 1u) || t >= len_ms || t < 0) continue;
                const i64 pc = ((i64)t / c.step) * c.step, nc = pc + c.step;
                if (nc <= c.last && window_silent((int)pc) && window_silent((int)nc)) cov |= 1u << b;
            }
        }
        const int left = len_ms - t0;                     // positions of this word inside the clip
        if (left < 32) cov &= left <= 0 ? 0u : ((1u << left) - 1u);
        return cov;
    };
    for (int cw = tid; cw < n_cw; cw += SIL_THREADS) s_cov[cw] = cover_word(cw);
    if (tid == 0) s_prev_cov = c0 > 0 ? (cover_word(-1) >> 31) : 0u;
    __syncthreads();

    // ---- 4. run starts per word, counts, exchange between blocks ----
    const bool prev_block_cov = s_prev_cov != 0u;
    const int per_c = (n_cw + SIL_THREADS - 1) / SIL_THREADS;
    const int lo_c = min(tid * per_c, n_cw), hi_c = min(lo_c + per_c, n_cw);
    auto start_masks = [&](int cw, unsigned& ns, unsigned& ss) {
        const unsigned cov = s_cov[cw];
        const int t0 = c0 + 32 * cw;
        const int left = len_ms - t0;
        const unsigned valid = left >= 32 ? 0xffffffffu : left <= 0 ? 0u : ((1u << left) - 1u);
        unsigned pbit;                                     // cover(t0 - 1)
        if (cw > 0) pbit = s_cov[cw - 1] >> 31;
        else pbit = prev_block_cov ? 1u : 0u;
        unsigned prev_ns = (cov << 1) | pbit, prev_ss = prev_ns;
        if (t0 == 0) { prev_ns |= 1u; prev_ss &= ~1u; }    // before the clip: a nonsilent run starts if t = 0 is uncovered, a silent one if covered
        ns = ~cov & prev_ns & valid;
        ss = cov & ~prev_ss & valid;
    };
    i64 cnt = 0;                                           // nonsilent starts << 32 | silent starts
    for (int cw = lo_c; cw < hi_c; cw++) {
        unsigned ns, ss;
        start_masks(cw, ns, ss);
        cnt += ((i64)__popc(ns) << 32) | (i64)__popc(ss);
    }
    i64 tot;
    const i64 pre = block_exclusive_scan<i64>(cnt, s_w64, &tot);
    {
        if (tid == 0) {
            const unsigned long long st = (1ull << 63) | ((unsigned long long)(tot >> 32) << 31) | (unsigned long long)(tot & 0x7fffffffLL);
            atomicExch(&ws->status[ck], st);
        }
        // counts of the earlier chunks, one per thread: on the block's first chunk every one of them (< SIL_MAX_BLOCKS),
        // afterwards the chunks since its previous one (gridDim.x of them)
        const bool first = ck == (int)blockIdx.x;
        const int p_lo = first ? 0 : ck - (int)gridDim.x;
        i64 part = 0;
        if (p_lo + tid < ck) {
            unsigned long long st;
            do { st = *(volatile unsigned long long*)&ws->status[p_lo + tid]; } while (!(st >> 63));
            part = (i64)(((st >> 31) & 0xffffffffull) << 32) | (i64)(st & 0x7fffffffull);
        }
        if (warp * 32 < ck - p_lo) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0 && part) atomicAdd(&s_acc, (unsigned long long)part);
        }
        __syncthreads();
        base_run = (first ? 0 : base_run) + (i64)s_acc;
    }

    // ---- 5. scatter: every run start is also the end of the run of the other kind before it ----
    {
        i64 at = base_run + pre;
        int i0 = (int)(at >> 32), i1 = (int)(at & 0xffffffffLL);     // nonsilent / silent starts before this thread's first word
        for (int cw = lo_c; cw < hi_c; cw++) {
            unsigned ns, ss;
            start_masks(cw, ns, ss);
            const int t0 = c0 + 32 * cw;
            unsigned both = ns | ss;                                   // never both at one position
            while (both) {
                const int b = __ffs((int)both) - 1;
                both &= both - 1;
                const int32_t t = (int32_t)(t0 + b);
                if ((ns >> b) & 1u) {
                    if (i0 < c.cap && nonsilent_ms) nonsilent_ms[2 * i0] = t;
                    if (t > 0 && i1 - 1 < c.cap && silent_ms) silent_ms[2 * (i1 - 1) + 1] = t;      // ends the silent run before it
                    i0++;
                } else {
                    if (i1 < c.cap && silent_ms) silent_ms[2 * i1] = t;
                    if (t > 0 && i0 - 1 < c.cap && nonsilent_ms) nonsilent_ms[2 * (i0 - 1) + 1] = t;
                    i1++;
                }
            }
        }
    }
    if (ck == c.n_chunks - 1 && tid == 0) {
        // totals, and the end of the run that is open at the end of the clip
        const i64 all = base_run + tot;
        const int g0 = (int)(all >> 32), g1 = (int)(all & 0xffffffffLL);
        const int tl = len_ms - 1 - c0;                                // last position of the clip, relative to this block
        const bool last_cov = tl >= 0 && ((s_cov[tl / 32] >> (tl % 32)) & 1u);
        if (last_cov) { if (g1 >= 1 && g1 - 1 < c.cap && silent_ms) silent_ms[2 * (g1 - 1) + 1] = (int32_t)c.len_ms; }
        else { if (g0 >= 1 && g0 - 1 < c.cap && nonsilent_ms) nonsilent_ms[2 * (g0 - 1) + 1] = (int32_t)c.len_ms; }
        info[B2A_INFO_N_NONSILENT] = g0 < c.cap ? g0 : c.cap;
        info[B2A_INFO_N_SILENT] = g1 < c.cap ? g1 : c.cap;
        info[B2A_INFO_OVERFLOW] = (g0 > c.cap || g1 > c.cap) ? 1 : 0;
        info[B2A_INFO_LEN_MS] = c.len_ms;
    }

    __syncthreads();                                    // shared memory is reused by the block's next chunk
  }

    // ---- 6. the block that finishes last derives the kept ranges ----
    __threadfence();
    __syncthreads();
    if (tid == 0) s_is_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    silence_kept_pass(nonsilent_ms, c, kept_ms, kept_off, info, s_w64, &s_carry);
}

// gather: one thread per kept millisecond.  Segments start on ms boundaries in both source and destination,
// so for 16 samples/ms every copy is two aligned 16-byte vectors.
__global__ void __launch_bounds__(256) compact_kernel(const int16_t* __restrict__ pcm, i64 n_samples, int spm,
                                                      const int32_t* __restrict__ kept_ms, const i64* __restrict__ kept_off,
                                                      const i64* __restrict__ info, int16_t* __restrict__ out, i64 out_cap) {
    const i64 n_keep = info[B2A_INFO_N_KEEP];
    const int n_seg = (int)info[B2A_INFO_N_KEPT];
    const i64 d_ms = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 d0 = d_ms * spm;
    if (d0 >= n_keep || n_seg <= 0) return;
    // largest k with kept_off[k] <= d0
    int lo = 0, hi = n_seg - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (kept_off[mid] <= d0) lo = mid; else hi = mid - 1;
    }
    const i64 src0 = (i64)kept_ms[2 * lo] * spm + (d0 - kept_off[lo]);
    if (d0 + spm > out_cap) return;
    const bool vec = (spm % 8 == 0) && (src0 + spm <= n_samples) &&
                     ((((uintptr_t)(pcm + src0)) | ((uintptr_t)(out + d0))) & 15) == 0;
    if (vec) {
        const uint4* s4 = (const uint4*)(pcm + src0);
        uint4* d4 = (uint4*)(out + d0);
        for (int j = 0; j < spm / 8; j++) d4[j] = s4[j];
    } else {
        for (int j = 0; j < spm; j++) {
            i64 si = src0 + j;
            out[d0 + j] = si < n_samples ? pcm[si] : (int16_t)0;     // pydub zero-fills a rounded-up last ms
        }
    }
}

// degenerate clip (len_ms == 0): detect_nonsilent -> [[0, 0]], nothing kept
__global__ void silence_empty_kernel(int32_t* nonsilent_ms, int32_t* kept_ms, i64* kept_off, i64* info, int cap) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int i = 0; i < B2A_INFO_LEN; i++) info[i] = 0;
        if (cap > 0) {
            if (nonsilent_ms) { nonsilent_ms[0] = 0; nonsilent_ms[1] = 0; }
            if (kept_ms) { kept_ms[0] = 0; kept_ms[1] = 0; }
            info[B2A_INFO_N_NONSILENT] = 1;
            info[B2A_INFO_N_KEPT] = 1;
            kept_off[0] = 0;
            kept_off[1] = 0;
        }
    }
}

static i64 pydub_len_ms(i64 n_frames, int sample_rate) {
    // pydub AudioSegment.__len__: round(1000 * (frame_count / frame_rate)), Python round = half-to-even
    double v = 1000.0 * ((double)n_frames / (double)sample_rate);
    return (i64)std::nearbyint(v);
}

// workspace: the ticket counter and the chunks' status words
static size_t silence_ws_bytes(i64, int) { return align_up(sizeof(SilenceWs) + 8 * (size_t)SIL_MAX_CHUNKS, 256) + 256; }

// dynamic shared memory of silence_kernel: prefix sums (padded to an even count) + flag, cover and last-start words
static size_t silence_smem_bytes(int chunk, int W) {
    const int Wr = (W + 31) & ~31;
    const size_t n_P = (size_t)chunk + Wr + W;
    return ((n_P + 1) & ~(size_t)1) * 8 + (size_t)((chunk + Wr) / 32) * 8 + (size_t)(chunk / 32) * 4 + 64;
}

int silence_build_cfg(i64 n_samples, int sample_rate, const b2a_silence_params* prm, int cap, SilenceCfg* c) {
    if (!prm) { set_error("silence: null params"); return B2A_EINVAL; }
    if (n_samples < 0 || sample_rate <= 0) { set_error("silence: bad length/rate"); return B2A_EINVAL; }
    if (sample_rate % 1000) { set_error("silence: sample_rate %d is not a whole number of samples per ms", sample_rate); return B2A_EUNSUPPORTED; }
    if (prm->min_silence_len <= 0 || prm->seek_step <= 0 || prm->keep_silence < -1 || cap <= 0) {
        set_error("silence: bad parameters (min_silence_len=%d seek_step=%d keep_silence=%d cap=%d)", prm->min_silence_len,
                  prm->seek_step, prm->keep_silence, cap);
        return B2A_EINVAL;
    }
    if (prm->min_silence_len > 10000) { set_error("silence: min_silence_len > 10000 ms unsupported"); return B2A_EUNSUPPORTED; }
    c->n_samples = n_samples;
    c->spm = sample_rate / 1000;
    c->len_ms = pydub_len_ms(n_samples, sample_rate);
    c->n_energy = (n_samples + c->spm - 1) / c->spm;
    c->W = prm->min_silence_len;
    c->step = prm->seek_step;
    c->keep = prm->keep_silence < 0 ? c->len_ms : prm->keep_silence;
    c->last = c->len_ms - c->W;
    // pydub: silence_thresh = db_to_float(dB) * max_possible_amplitude;  rms <= thresh
    double thr = std::pow(10.0, prm->silence_thresh / 20.0) * 32768.0;
    double kf = std::floor(thr) + 1.0;
    if (kf < 1.0) kf = 1.0;
    if (kf > 4.0e6) kf = 4.0e6;            // far above any 16-bit rms; keeps the product inside uint64
    u64 k = (u64)kf;
    c->limit = (u64)c->W * (u64)c->spm * k * k;
    c->cap = cap;
    // chunks: as many as there are SMs when the clip is long enough (at least SIL_MIN_CHUNK ms each); the chunk's energies
    // and their prefix sums must fit in shared memory, so very long clips take several rounds of SIL_MAX_BLOCKS chunks
    const i64 Wr = (c->W + 31) / 32 * 32;
    i64 cap_chunk = (SIL_SMEM_BUDGET - 64 - 8 * (Wr + c->W + 2)) / 9 / 32 * 32;
    while (cap_chunk >= 0 && (i64)silence_smem_bytes((int)cap_chunk + 32, c->W) <= SIL_SMEM_BUDGET) cap_chunk += 32;
    if (cap_chunk < 32) { set_error("silence: min_silence_len %d does not fit the shared-memory window", c->W); return B2A_EUNSUPPORTED; }
    i64 blocks = (c->len_ms + SIL_MIN_CHUNK - 1) / SIL_MIN_CHUNK;
    if (blocks > SIL_MAX_BLOCKS) blocks = SIL_MAX_BLOCKS;
    if (blocks < 1) blocks = 1;
    i64 rounds = (c->len_ms + blocks * cap_chunk - 1) / (blocks * cap_chunk);
    if (rounds < 1) rounds = 1;
    i64 chunk = ((c->len_ms + blocks * rounds - 1) / (blocks * rounds) + 31) / 32 * 32;
    if (chunk < 32) chunk = 32;
    if (chunk > cap_chunk) chunk = cap_chunk;
    i64 n_chunks = (c->len_ms + chunk - 1) / chunk;
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > SIL_MAX_CHUNKS || c->len_ms + chunk + 2 * Wr >= (i64)1 << 31) {
        set_error("silence: clip of %lld ms is too long (at most %d chunks of %lld ms)", (long long)c->len_ms, SIL_MAX_CHUNKS, (long long)chunk);
        return B2A_EUNSUPPORTED;
    }
    c->chunk = (int)chunk;
    c->n_chunks = (int)n_chunks;
    c->n_blocks = (int)(n_chunks < SIL_MAX_BLOCKS ? n_chunks : SIL_MAX_BLOCKS);
    return B2A_OK;
}

int silence_launch(const u64* d_energy, i64 n_samples, int sample_rate, const b2a_silence_params* prm, int cap,
                   int32_t* d_silent, int32_t* d_nonsilent, int32_t* d_kept, i64* d_kept_off, i64* d_info, void* d_ws,
                   size_t ws_bytes, cudaStream_t stream) {
    if (!d_energy || !d_kept_off || !d_info || !d_ws) { set_error("silence: null pointer"); return B2A_EINVAL; }
    if (!d_nonsilent) { set_error("silence: d_nonsilent_ms is required (kept ranges derive from it)"); return B2A_EINVAL; }
    SilenceCfg c;
    int rc = silence_build_cfg(n_samples, sample_rate, prm, cap, &c);
    if (rc) return rc;
    if (ws_bytes < silence_ws_bytes(n_samples, sample_rate)) { set_error("silence: workspace too small"); return B2A_EWORKSPACE; }
    if (c.len_ms <= 0) {
        auto k0 = silence_empty_kernel;
        B2A_LAUNCH(k0, 1, 32, 0, stream, d_nonsilent, d_kept, d_kept_off, d_info, cap);
        B2A_CHECK_LAUNCH("silence_empty_kernel");
        return B2A_OK;
    }
    if (((uintptr_t)d_ws) & 7) { set_error("silence: workspace must be 8-byte aligned"); return B2A_EINVAL; }
    SilenceWs* wsp = (SilenceWs*)d_ws;
    cudaError_t e = cudaMemsetAsync(wsp, 0, sizeof(SilenceWs) + 8 * (size_t)c.n_chunks, stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(silence workspace)");
    const size_t smem = silence_smem_bytes(c.chunk, c.W);
    static const size_t kSmemMax = SIL_SMEM_BUDGET;
    auto k1 = silence_kernel;
    {
        // the opt-in to large dynamic shared memory is set ONCE per device to the largest size any call can need
        // (concurrent callers with different min_silence_len must not lower each other's limit)
        static unsigned long long attr_mask = 0;
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64 || !((attr_mask >> dev) & 1ull)) {
            e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax);
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(silence_kernel)");
            if (dev >= 0 && dev < 64) attr_mask |= 1ull << dev;
        }
    }
    B2A_LAUNCH(k1, c.n_blocks, SIL_THREADS, smem, stream, d_energy, c, d_silent, d_nonsilent, d_kept, d_kept_off, d_info, wsp);
    B2A_CHECK_LAUNCH("silence_kernel");
    return B2A_OK;
}

int compact_launch(const int16_t* d_pcm, i64 n_samples, int sample_rate, const int32_t* d_kept, const i64* d_kept_off,
                   const i64* d_info, int16_t* d_out, i64 out_cap, cudaStream_t stream) {
    if (!d_pcm || !d_kept || !d_kept_off || !d_info || !d_out) { set_error("compact: null pointer"); return B2A_EINVAL; }
    if (sample_rate <= 0 || sample_rate % 1000) { set_error("compact: bad sample rate"); return B2A_EUNSUPPORTED; }
    const int spm = sample_rate / 1000;
    i64 len_ms = pydub_len_ms(n_samples, sample_rate);
    if (len_ms <= 0) return B2A_OK;
    if (out_cap < len_ms * spm) { set_error("compact: output capacity %lld < %lld", (long long)out_cap, (long long)(len_ms * spm)); return B2A_EINVAL; }
    auto k = compact_kernel;
    i64 blocks = (len_ms + 255) / 256;
    B2A_LAUNCH(k, (unsigned)blocks, 256, 0, stream, d_pcm, n_samples, spm, d_kept, d_kept_off, d_info, d_out, out_cap);
    B2A_CHECK_LAUNCH("compact_kernel");
    return B2A_OK;
}

int energy_launch(const int16_t* d_pcm, i64 n, int sample_rate, u64* d_energy, cudaStream_t stream) {
    if (!d_pcm || !d_energy) { set_error("energy: null pointer"); return B2A_EINVAL; }
    if (sample_rate <= 0 || sample_rate % 1000) { set_error("energy: bad sample rate"); return B2A_EUNSUPPORTED; }
    const int spm = sample_rate / 1000;
    i64 ne = (n + spm - 1) / spm;
    if (ne <= 0) return B2A_OK;
    auto k = energy_ms_kernel;
    B2A_LAUNCH(k, (unsigned)((ne + 255) / 256), 256, 0, stream, d_pcm, n, spm, ne, d_energy);
    B2A_CHECK_LAUNCH("energy_ms_kernel");
    return B2A_OK;
}

size_t silence_workspace_bytes(i64 n_samples, int sample_rate) { return silence_ws_bytes(n_samples, sample_rate); }

}  // namespace b2a

extern "C" {

size_t b2a_silence_workspace_bytes(int64_t n_samples, int sample_rate) {
    if (n_samples < 0 || sample_rate <= 0) return 0;
    return b2a::silence_ws_bytes(n_samples, sample_rate);
}

int b2a_energy_ms(const int16_t* d_pcm, int64_t n, int sample_rate, uint64_t* d_energy_ms, b2a_stream_t stream) {
    return b2a::energy_launch(d_pcm, n, sample_rate, (b2a::u64*)d_energy_ms, (cudaStream_t)stream);
}

int b2a_detect_silence(const uint64_t* d_energy_ms, int64_t n_samples, int sample_rate, const b2a_silence_params* params,
                       int32_t cap, int32_t* d_silent_ms, int32_t* d_nonsilent_ms, int32_t* d_kept_ms, int64_t* d_kept_off,
                       int64_t* d_info, void* d_ws, size_t ws_bytes, b2a_stream_t stream) {
    return b2a::silence_launch((const b2a::u64*)d_energy_ms, n_samples, sample_rate, params, cap, d_silent_ms, d_nonsilent_ms,
                               d_kept_ms, (b2a::i64*)d_kept_off, (b2a::i64*)d_info, d_ws, ws_bytes, (cudaStream_t)stream);
}

int b2a_compact(const int16_t* d_pcm, int64_t n_samples, int sample_rate, const int32_t* d_kept_ms, const int64_t* d_kept_off,
                const int64_t* d_info, int16_t* d_out, int64_t out_capacity, b2a_stream_t stream) {
    return b2a::compact_launch(d_pcm, n_samples, sample_rate, d_kept_ms, (const b2a::i64*)d_kept_off, (const b2a::i64*)d_info,
                               d_out, out_capacity, (cudaStream_t)stream);
}

}  // extern "C"
