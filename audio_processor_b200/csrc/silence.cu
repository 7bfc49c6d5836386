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
// Kernels: cover_kernel (windowed energy test + coverage mask + per-tile run counts, all from
// shared-memory prefix sums over a tile with a W-ms halo), ranges_kernel (ordered scatter of the
// run boundaries), kept_kernel (one block: keep_silence padding, midpoints, clamps, exclusive scan of
// lengths), compact_kernel (gather kept milliseconds; warp-per-32-ms binary search in the offset table).
#include "b2a_common.cuh"

#include <cmath>

namespace b2a {

constexpr int SIL_TS = 2048;        // milliseconds per tile (4096 was measured: fewer halo re-reads but 3 instead of 5 blocks per SM, 38 vs 33.5 us)
constexpr int SIL_THREADS = 256;
constexpr int SIL_PER_THREAD = SIL_TS / SIL_THREADS;   // 8 consecutive ms per thread

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
    int n_tiles;
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

// block-wide exclusive prefix over per-thread values (256 threads); returns exclusive prefix, total in *total
template <class T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* s_warp /*[8]*/, T* total) {
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
    T wpre = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SIL_THREADS / 32; w++) {
        T x = s_warp[w];
        if (w < warp) wpre += x;
        tot += x;
    }
    *total = tot;
    return wpre + inc - v;
}

// cover[t] for t in [0, len_ms): 1 if millisecond t lies inside a (merged) silent range.
// counts[tile*2+0] = nonsilent runs starting in the tile, counts[tile*2+1] = silent runs starting in the tile.
__global__ void __launch_bounds__(SIL_THREADS) cover_kernel(const u64* __restrict__ e, SilenceCfg c,
                                                            unsigned char* __restrict__ cover, int* __restrict__ counts) {
    B2A_DYN_SMEM(smem_raw);
    const int W = c.W;
    const int NE = SIL_TS + 2 * W + 1;          // energies e[t0-W-1 .. t0+TS+W-1]  (+1 slot for the prefix)
    u64* s_p = (u64*)smem_raw;                  // NE+1 : exclusive prefix of energies
    int* s_c = (int*)(s_p + NE + 1);            // SIL_TS + W + 2 : exclusive prefix of start flags f[t0-W-1 ..]
    __shared__ u64 s_w64[SIL_THREADS / 32];
    __shared__ int s_w32[SIL_THREADS / 32];
    __shared__ int s_cnt[2];

    const int tid = threadIdx.x;
    const i64 t0 = (i64)blockIdx.x * SIL_TS;
    const i64 ebase = t0 - W - 1;               // energy index of local slot 0
    if (tid < 2) s_cnt[tid] = 0;

    // ---- exclusive prefix of energies over the tile + halos ----
    // coalesced load of the raw energies (slot j+1), then each thread scans a contiguous chunk; the chunk length is odd
    // so that the 8-byte shared-memory accesses of a warp (stride = chunk) fall in distinct banks
    for (int j = tid; j < NE; j += SIL_THREADS) {
        const i64 t = ebase + j;
        s_p[j + 1] = (t >= 0 && t < c.len_ms && t < c.n_energy) ? e[t] : 0ull;
    }
    __syncthreads();
    {
        const int per = ((NE + SIL_THREADS - 1) / SIL_THREADS) | 1;
        const int lo = min(tid * per, NE), hi = min(lo + per, NE);
        u64 sum = 0;
        for (int j = lo; j < hi; j++) sum += s_p[j + 1];
        u64 tot;
        u64 pre = block_exclusive_scan<u64>(sum, s_w64, &tot);
        // turn the raw values into an exclusive prefix: s_p[j] = sum of slots < j
        u64 run = pre;
        for (int j = lo; j < hi; j++) {
            u64 v = s_p[j + 1];
            s_p[j + 1] = run + v;                // inclusive at j -> exclusive at j+1
            run += v;
        }
        if (tid == 0) s_p[0] = 0;
    }
    __syncthreads();

    // ---- start flags f[i] for i in [t0-W-1, t0+TS), then their exclusive prefix ----
    {
        const int NF = SIL_TS + W + 1;
        const int per = ((NF + SIL_THREADS - 1) / SIL_THREADS) | 1;      // odd: conflict-free strided shared-memory access
        const int lo = min(tid * per, NF), hi = min(lo + per, NF);
        int sum = 0;
        for (int j = lo; j < hi; j++) {
            i64 i = ebase + j;                   // flag slot j <-> start ms i (same origin as energies)
            int f = 0;
            if (i >= 0 && i <= c.last && (c.step == 1 || (i % c.step) == 0 || i == c.last)) {
                u64 E = s_p[j + W] - s_p[j];     // sum e[i .. i+W-1]
                f = E < c.limit;
            }
            s_c[j + 1] = f;
            sum += f;
        }
        int tot;
        int pre = block_exclusive_scan<int>(sum, s_w32, &tot);
        int run = pre;
        for (int j = lo; j < hi; j++) {
            int v = s_c[j + 1];
            s_c[j + 1] = run + v;
            run += v;
        }
        if (tid == 0) s_c[0] = 0;
    }
    __syncthreads();

    // ---- coverage for t in [t0-1, t0+TS): any start in (t-W, t]  (+ the seek_step > W continuity rule) ----
    // local flag slot of start i is j = i - ebase; covered(t) <=> C[j(t)+1] - C[j(t-W+1)] > 0
    auto covered = [&](i64 t) -> int {
        if (t < 0 || t >= c.len_ms) return 0;
        int jt = (int)(t - ebase);
        int jl = jt - W + 1;
        if (jl < 0) jl = 0;
        int cov = (s_c[jt + 1] - s_c[jl]) > 0;
        if (!cov && c.step > W) {
            // consecutive candidates p, p+step both silent are "continuous" in pydub and merge
            i64 pc = (t / c.step) * c.step;
            i64 nc = pc + c.step;
            if (nc <= c.last) {
                u64 E0 = 0, E1 = 0;
                for (int k = 0; k < W; k++) { E0 += e[pc + k]; E1 += e[nc + k]; }
                cov = (E0 < c.limit) && (E1 < c.limit);
            }
        }
        return cov;
    };
    int ns_starts = 0, s_starts = 0;
    {
        const i64 tb = t0 + (i64)tid * SIL_PER_THREAD;
        int prev = (tb == 0) ? -1 : covered(tb - 1);     // -1 = before the clip
#pragma unroll
        for (int k = 0; k < SIL_PER_THREAD; k++) {
            i64 t = tb + k;
            if (t < c.len_ms) {
                int cv = covered(t);
                cover[t] = (unsigned char)cv;
                if (!cv && (prev != 0)) ns_starts++;  // nonsilent run starts: uncovered and (t==0 or previous covered)
                if (cv && (prev != 1)) s_starts++;    // silent run starts: covered and (t==0 or previous uncovered)
                prev = cv;
            }
        }
    }
    ns_starts = warp_reduce_sum_i(ns_starts);
    s_starts = warp_reduce_sum_i(s_starts);
    if ((tid & 31) == 0) { atomicAdd(&s_cnt[0], ns_starts); atomicAdd(&s_cnt[1], s_starts); }
    __syncthreads();
    if (tid < 2) counts[blockIdx.x * 2 + tid] = s_cnt[tid];
}

// ordered scatter of run boundaries.  A run containing ms t has index (#run starts at positions <= t) - 1.
__global__ void __launch_bounds__(SIL_THREADS) ranges_kernel(const unsigned char* __restrict__ cover, const int* __restrict__ counts,
                                                             SilenceCfg c, int32_t* __restrict__ silent_ms,
                                                             int32_t* __restrict__ nonsilent_ms, i64* __restrict__ info) {
    __shared__ int s_w32[SIL_THREADS / 32];
    __shared__ int s_base[2];
    const int tid = threadIdx.x;
    const i64 t0 = (i64)blockIdx.x * SIL_TS;

    // runs that started in earlier tiles
    int a0 = 0, a1 = 0, g0 = 0, g1 = 0;
    for (int i = tid; i < c.n_tiles; i += SIL_THREADS) {
        int x0 = counts[i * 2], x1 = counts[i * 2 + 1];
        g0 += x0; g1 += x1;
        if (i < (int)blockIdx.x) { a0 += x0; a1 += x1; }
    }
    int tot;
    block_exclusive_scan<int>(a0, s_w32, &tot); if (tid == 0) s_base[0] = tot;
    block_exclusive_scan<int>(a1, s_w32, &tot); if (tid == 0) s_base[1] = tot;
    int gt0, gt1;
    block_exclusive_scan<int>(g0, s_w32, &gt0);
    block_exclusive_scan<int>(g1, s_w32, &gt1);
    __syncthreads();
    if (blockIdx.x == 0 && tid == 0) {
        info[B2A_INFO_N_NONSILENT] = gt0 < c.cap ? gt0 : c.cap;
        info[B2A_INFO_N_SILENT] = gt1 < c.cap ? gt1 : c.cap;
        info[B2A_INFO_OVERFLOW] = (gt0 > c.cap || gt1 > c.cap) ? 1 : 0;
        info[B2A_INFO_LEN_MS] = c.len_ms;
    }

    const i64 tb = t0 + (i64)tid * SIL_PER_THREAD;
    int cv[SIL_PER_THREAD + 2];
#pragma unroll
    for (int k = 0; k < SIL_PER_THREAD + 2; k++) {
        i64 t = tb - 1 + k;
        cv[k] = (t >= 0 && t < c.len_ms) ? (int)cover[t] : -1;   // -1 = outside the clip
    }
    int n0 = 0, n1 = 0;
#pragma unroll
    for (int k = 1; k <= SIL_PER_THREAD; k++) {
        if (cv[k] == 0 && cv[k - 1] != 0) n0++;
        if (cv[k] == 1 && cv[k - 1] != 1) n1++;
    }
    int e0 = block_exclusive_scan<int>(n0, s_w32, &tot) + s_base[0];
    int e1 = block_exclusive_scan<int>(n1, s_w32, &tot) + s_base[1];
    // e0/e1 = number of nonsilent/silent run starts strictly before this thread's first ms
#pragma unroll
    for (int k = 1; k <= SIL_PER_THREAD; k++) {
        const i64 t = tb - 1 + k;
        if (cv[k] < 0) continue;
        if (cv[k] == 0) {
            if (cv[k - 1] != 0) { if (e0 < c.cap && nonsilent_ms) nonsilent_ms[2 * e0] = (int32_t)t; e0++; }
            if (cv[k + 1] != 0) { int r = e0 - 1; if (r < c.cap && nonsilent_ms) nonsilent_ms[2 * r + 1] = (int32_t)(t + 1); }
        } else {
            if (cv[k - 1] != 1) { if (e1 < c.cap && silent_ms) silent_ms[2 * e1] = (int32_t)t; e1++; }
            if (cv[k + 1] != 1) { int r = e1 - 1; if (r < c.cap && silent_ms) silent_ms[2 * r + 1] = (int32_t)(t + 1); }
        }
    }
}

// one block: split_on_silence's range arithmetic + exclusive scan of kept lengths (in samples)
__global__ void __launch_bounds__(SIL_THREADS) kept_kernel(const int32_t* __restrict__ nonsilent_ms, SilenceCfg c,
                                                           int32_t* __restrict__ kept_ms, i64* __restrict__ kept_off,
                                                           i64* __restrict__ info) {
    __shared__ i64 s_w64[SIL_THREADS / 32];
    __shared__ i64 s_carry;
    const int tid = threadIdx.x;
    const int n = (int)info[B2A_INFO_N_NONSILENT];
    if (tid == 0) s_carry = 0;
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
        const i64 carry = s_carry;
        if (k < n) kept_off[k] = carry + pre;
        __syncthreads();
        if (tid == 0) s_carry = carry + tot;
        __syncthreads();
        if (n == 0) break;
    }
    if (tid == 0) {
        kept_off[n] = s_carry;
        info[B2A_INFO_N_KEPT] = n;
        info[B2A_INFO_N_KEEP] = s_carry;
    }
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

// workspace: [cover: len_ms bytes][counts: 2*n_tiles ints][nonsilent scratch when the caller passes NULL]
static size_t silence_ws_bytes(i64 n_samples, int sample_rate) {
    i64 len_ms = pydub_len_ms(n_samples, sample_rate) + 2;
    i64 n_tiles = (len_ms + SIL_TS - 1) / SIL_TS + 1;
    return align_up((size_t)len_ms, 256) + align_up((size_t)n_tiles * 8, 256) + 256;
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
    c->n_tiles = (int)((c->len_ms + SIL_TS - 1) / SIL_TS);
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
    unsigned char* cover = (unsigned char*)d_ws;
    int* counts = (int*)((char*)d_ws + align_up((size_t)c.len_ms + 2, 256));
    size_t smem = (size_t)(SIL_TS + 2 * c.W + 2) * 8 + (size_t)(SIL_TS + c.W + 2) * 4 + 64;
    auto k1 = cover_kernel;
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(cover_kernel)");
    B2A_LAUNCH(k1, c.n_tiles, SIL_THREADS, smem, stream, d_energy, c, cover, counts);
    B2A_CHECK_LAUNCH("cover_kernel");
    auto k2 = ranges_kernel;
    B2A_LAUNCH(k2, c.n_tiles, SIL_THREADS, 0, stream, (const unsigned char*)cover, (const int*)counts, c, d_silent, d_nonsilent, d_info);
    B2A_CHECK_LAUNCH("ranges_kernel");
    auto k3 = kept_kernel;
    B2A_LAUNCH(k3, 1, SIL_THREADS, 0, stream, (const int32_t*)d_nonsilent, c, d_kept, d_kept_off, d_info);
    B2A_CHECK_LAUNCH("kept_kernel");
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
