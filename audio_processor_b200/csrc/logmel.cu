// logmel.cu — Whisper log-mel spectrogram on sm_100a.
//
// Replaces whisper.audio.log_mel_spectrogram (openai-whisper whisper/audio.py), which the
// reference reaches through model.transcribe at app/services/audio_processor.py:1076-1080:
//   reflect-pad 200 | 400-sample frames at hop 160 | periodic Hann | rFFT-400 | |.|^2 |
//   drop last frame | slaney mel [n_mels x 201] | log10(max(.,1e-10)) | max(., gmax-8) | (x+4)/4
//
// Kernel design (K4 "stft_mel", K5 "mel_floor"):
//  * persistent blocks of 128 threads; each iteration handles a tile of 24 frames of one clip.
//  * the tile's 4080 samples are staged once in shared memory as f32 (reflect / zero-pad resolved
//    at load), skewed by 10 words per hop so the 6 frames a warp works on hit disjoint banks.
//  * one real 400-point FFT per frame = one complex 200-point FFT of (even,odd) samples, done by
//    5 threads: radix-10 butterflies (2x5 prime-factor) -> twiddle -> shared-memory transpose ->
//    radix-20 butterflies (4x5 prime-factor).  Each thread owns output residues {u, 10-u} mod 10,
//    so the real-FFT unpack pairs (k, 200-k) stay inside one thread: no second exchange.
//  * power spectrum goes back to shared memory; the sparse mel projection (<=14 bins per filter),
//    log10, running max/min and the [n_mels][T] store are done per (mel, 8-frame group) so each
//    filter row is loaded once per 8 frames and stores are 32-byte runs.
//  * K5 applies Whisper's global floor in place and skips tiles whose minimum is already above it.
#include "b2a_tables.cuh"

namespace b2a {

constexpr int LM_FRAMES = 24;                       // frames per tile
constexpr int LM_THREADS = 128;                     // 4 warps x (6 frames x 5 threads)
constexpr int LM_TILE = LM_FRAMES * kHop + 240;     // 4080 samples
constexpr int LM_SKEW = 10;                         // extra words per hop (bank de-phasing)
constexpr int LM_TILE_WORDS = LM_TILE + LM_SKEW * (LM_TILE / kHop + 1);
constexpr int LM_EF = 426;                          // exchange words per frame (>= 400, = 10 mod 32)
constexpr int LM_MW = kMelMaxWidth + 1;             // padded mel weight row

struct cpx { float r, i; };
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return {a.r + b.r, a.i + b.i}; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return {a.r - b.r, a.i - b.i}; }
__device__ __forceinline__ cpx cmul(cpx a, cpx b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
__device__ __forceinline__ cpx cmuli_neg(cpx a) { return {a.i, -a.r}; }  // -i * a

// 5-point DFT (forward), in place
__device__ __forceinline__ void dft5(cpx& x0, cpx& x1, cpx& x2, cpx& x3, cpx& x4) {
    const float kS1 = 0.95105651629515357f, kS2 = 0.58778525229247313f, kC = 0.55901699437494742f;
    cpx t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
    cpx t5 = cadd(t1, t2);
    cpx m1 = {x0.r - 0.25f * t5.r, x0.i - 0.25f * t5.i};
    cpx m2 = {(t1.r - t2.r) * kC, (t1.i - t2.i) * kC};
    cpx s = cadd(m1, m2), d = csub(m1, m2);
    cpx u = {kS1 * t3.r + kS2 * t4.r, kS1 * t3.i + kS2 * t4.i};
    cpx v = {kS2 * t3.r - kS1 * t4.r, kS2 * t3.i - kS1 * t4.i};
    cpx mu = cmuli_neg(u), mv = cmuli_neg(v);
    x0 = cadd(x0, t5);
    x1 = cadd(s, mu);
    x4 = csub(s, mu);
    x2 = cadd(d, mv);
    x3 = csub(d, mv);
}

__device__ __forceinline__ void dft4(cpx& x0, cpx& x1, cpx& x2, cpx& x3) {
    cpx a = cadd(x0, x2), b = csub(x0, x2), c = cadd(x1, x3), d = csub(x1, x3);
    cpx md = cmuli_neg(d);
    x0 = cadd(a, c);
    x1 = cadd(b, md);
    x2 = csub(a, c);
    x3 = csub(b, md);
}

// 10-point DFT via 2x5 prime-factor map: n = (5a+2b)%10, k = (5c+6d)%10.  in x[10] -> out y[10]
__device__ __forceinline__ void dft10(const cpx (&x)[10], cpx (&y)[10]) {
    cpx s[5], d[5];
#pragma unroll
    for (int b = 0; b < 5; b++) {
        cpx p = x[(2 * b) % 10], q = x[(5 + 2 * b) % 10];
        s[b] = cadd(p, q);
        d[b] = csub(p, q);
    }
    dft5(s[0], s[1], s[2], s[3], s[4]);
    dft5(d[0], d[1], d[2], d[3], d[4]);
#pragma unroll
    for (int e = 0; e < 5; e++) {
        y[(6 * e) % 10] = s[e];
        y[(5 + 6 * e) % 10] = d[e];
    }
}

// 20-point DFT via 4x5 prime-factor map: n = (5a+4b)%20, k = (5c+16d)%20.
__device__ __forceinline__ void dft20(const cpx (&x)[20], cpx (&y)[20]) {
    cpx r[4][5];
#pragma unroll
    for (int b = 0; b < 5; b++) {
        cpx a0 = x[(4 * b) % 20], a1 = x[(5 + 4 * b) % 20], a2 = x[(10 + 4 * b) % 20], a3 = x[(15 + 4 * b) % 20];
        dft4(a0, a1, a2, a3);
        r[0][b] = a0; r[1][b] = a1; r[2][b] = a2; r[3][b] = a3;
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
        dft5(r[c][0], r[c][1], r[c][2], r[c][3], r[c][4]);
#pragma unroll
        for (int d = 0; d < 5; d++) y[(5 * c + 16 * d) % 20] = r[c][d];
    }
}

// real-FFT unpack of one conjugate pair + power:  Z = FFT200(x_even + i x_odd), m = 200-k.
__device__ __forceinline__ void unpack_pair(cpx zk, cpx zm, float2 tw, float& pk, float& pm) {
    float er = 0.5f * (zk.r + zm.r), ei = 0.5f * (zk.i - zm.i);
    float orr = 0.5f * (zk.r - zm.r), oi = 0.5f * (zk.i + zm.i);
    float tr = tw.x * oi - tw.y * orr;
    float ti = tw.x * orr + tw.y * oi;
    float ar = er + tr, ai = ei - ti;
    float br = er - tr, bi = ei + ti;
    pk = ar * ar + ai * ai;
    pm = br * br + bi * bi;
}

struct LogMelParams {
    const void* audio;     // [batch] rows
    int fmt;               // B2A_FMT_*
    i64 row_stride;        // samples between rows
    i64 n;                 // samples per row (capacity when d_n != nullptr)
    const i64* d_n;        // optional device-side actual length (batch == 1)
    i64 padding;           // zeros appended on the right
    int n_mels;
    int batch;
    float* out;            // [batch][n_mels][T]
    i64* d_frames_out;     // optional
    int* gmax_key;         // [batch] (per-clip) or [1]
    int per_clip;
    float* tile_min;       // [batch][tiles_cap]
    i64 tiles_cap;         // tiles per clip at capacity
    const LogMelTables* tab;
};

__device__ __forceinline__ float load_sample(const void* audio, int fmt, i64 idx) {
    if (fmt == B2A_FMT_S16) return (float)((const int16_t*)audio)[idx] * (1.0f / 32768.0f);
    return ((const float*)audio)[idx];
}

__global__ void __launch_bounds__(LM_THREADS, 3) stft_mel_kernel(LogMelParams p) {
    B2A_DYN_SMEM(smem_raw);
    float* s_tile = (float*)smem_raw;                       // LM_TILE_WORDS
    float* s_ex = s_tile + LM_TILE_WORDS;                   // LM_FRAMES * LM_EF
    float* s_win = s_ex + LM_FRAMES * LM_EF;                // 400
    float2* s_tw200 = (float2*)(s_win + kNFFT);             // 200
    float2* s_tw400 = s_tw200 + 200;                        // 201 (+1 pad)
    float* s_melw = (float*)(s_tw400 + 202);                // n_mels * LM_MW
    int* s_mstart = (int*)(s_melw + kMelMaxMels * LM_MW);   // 128
    int* s_mlen = s_mstart + kMelMaxMels;                   // 128
    float* s_red = (float*)(s_mlen + kMelMaxMels);          // 8

    const int tid = threadIdx.x;
    const int n_mels = p.n_mels;
    const LogMelTables* tab = p.tab;

    // ---- stage the tables once per block ----
    for (int i = tid; i < kNFFT; i += LM_THREADS) s_win[i] = tab->win[i];
    for (int i = tid; i < 200; i += LM_THREADS) s_tw200[i] = tab->tw200[i];
    for (int i = tid; i < kNBins; i += LM_THREADS) s_tw400[i] = tab->tw400[i];
    for (int i = tid; i < n_mels * kMelMaxWidth; i += LM_THREADS) {
        int m = i / kMelMaxWidth, j = i % kMelMaxWidth;
        s_melw[m * LM_MW + j] = tab->mel_w[i];
    }
    for (int i = tid; i < n_mels; i += LM_THREADS) { s_mstart[i] = tab->mel_start[i]; s_mlen[i] = tab->mel_len[i]; }

    i64 n_act = p.n;
    if (p.d_n) { n_act = *p.d_n; if (n_act > p.n) n_act = p.n; if (n_act < 0) n_act = 0; }
    const i64 ltot = n_act + p.padding;          // padded length
    const i64 T = ltot / kHop;                   // frames
    const i64 tiles = (T + LM_FRAMES - 1) / LM_FRAMES;
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0 && p.d_frames_out) *p.d_frames_out = T;

    const int warp = tid >> 5, lane = tid & 31;
    const int fl_w = lane / 5, u = lane % 5;     // frame within warp, role within frame
    const bool fft_lane = lane < 30;
    const int fl = warp * 6 + fl_w;              // frame within tile
    const int k1a = u, k1b = (u == 0) ? 5 : 10 - u;

    float run_max = -3.0e38f;

    for (i64 work = blockIdx.x; work < tiles * p.batch; work += gridDim.x) {
        const int b = (int)(work / tiles);
        const i64 tile = work % tiles;
        const i64 t0 = tile * LM_FRAMES;
        const char* row = (const char*)p.audio + (size_t)b * (size_t)p.row_stride * (p.fmt == B2A_FMT_S16 ? 2 : 4);

        __syncthreads();   // previous iteration's readers are done with s_tile / s_ex (and tables are staged)
        // ---- tile load: padded-domain index q = 160*t0 - 200 + i, reflect at both ends, zeros past n_act ----
        for (int i = tid; i < LM_TILE; i += LM_THREADS) {
            i64 q = t0 * kHop - 200 + i;
            if (q < 0) q = -q;
            if (q >= ltot) q = 2 * (ltot - 1) - q;
            float v = 0.0f;
            if (q >= 0 && q < n_act) v = load_sample(row, p.fmt, q);
            s_tile[i + LM_SKEW * (i / kHop)] = v;
        }
        __syncthreads();

        // ---- stage 1: radix-10 butterflies over n1 for n2 = u + 5j, twiddle, transpose into s_ex ----
        float* ex = s_ex + fl * LM_EF;
        if (fft_lane) {
            const float* fr = s_tile + fl * (kHop + LM_SKEW);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int n2 = u + 5 * j;
                cpx x[10], y[10];
#pragma unroll
                for (int n1 = 0; n1 < 10; n1++) {
                    const int n = 20 * n1 + n2;
                    float2 xv = *(const float2*)(fr + 2 * n + LM_SKEW * (n1 / 4));
                    float2 wv = *(const float2*)(s_win + 2 * n);
                    x[n1].r = xv.x * wv.x;
                    x[n1].i = xv.y * wv.y;
                }
                dft10(x, y);
#pragma unroll
                for (int k1 = 0; k1 < 10; k1++) {
                    float2 tw = s_tw200[n2 * 10 + k1];
                    cpx w = {tw.x, tw.y};
                    cpx v = (k1 == 0) ? y[0] : cmul(y[k1], w);
                    *(float2*)(ex + k1 * 40 + 2 * n2) = make_float2(v.r, v.i);
                }
            }
        }
        __syncwarp();

        // ---- stage 2: radix-20 butterflies over n2 for residues k1a, k1b; unpack conjugate pairs in registers ----
        cpx za[20], zb[20];
        if (fft_lane) {
            cpx x[20];
#pragma unroll
            for (int n2 = 0; n2 < 20; n2++) { float2 v = *(const float2*)(ex + k1a * 40 + 2 * n2); x[n2] = {v.x, v.y}; }
            dft20(x, za);
#pragma unroll
            for (int n2 = 0; n2 < 20; n2++) { float2 v = *(const float2*)(ex + k1b * 40 + 2 * n2); x[n2] = {v.x, v.y}; }
            dft20(x, zb);
        }
        __syncwarp();   // every lane has read its exchange rows: the power spectrum may overwrite them
        if (fft_lane) {
            float* pw = ex;   // P[0..200] aliases the frame's exchange area
            if (u != 0) {
                // Z[k] = za[j] (k = u+10j),  Z[200-k] = zb[19-j]
#pragma unroll
                for (int j = 0; j < 20; j++) {
                    const int k = u + 10 * j;
                    float pk, pm;
                    unpack_pair(za[j], zb[19 - j], s_tw400[k], pk, pm);
                    pw[k] = pk;
                    pw[200 - k] = pm;
                }
            } else {
                // residue 0 pairs with itself: (10j, 200-10j); residue 5: (5+10j, 195-10j)
#pragma unroll
                for (int j = 0; j <= 10; j++) {
                    const int k = 10 * j;
                    float pk, pm;
                    unpack_pair(za[j], za[(20 - j) % 20], s_tw400[k], pk, pm);
                    pw[k] = pk;
                    pw[200 - k] = pm;
                }
#pragma unroll
                for (int j = 0; j < 10; j++) {
                    const int k = 5 + 10 * j;
                    float pk, pm;
                    unpack_pair(zb[j], zb[19 - j], s_tw400[k], pk, pm);
                    pw[k] = pk;
                    pw[200 - k] = pm;
                }
            }
        }
        __syncthreads();

        // ---- mel projection + log10 + store, one (mel, 8-frame group) per task ----
        float tmin = 3.0e38f;
        float* outb = p.out + (size_t)b * (size_t)n_mels * (size_t)T;
        for (int task = tid; task < n_mels * (LM_FRAMES / 8); task += LM_THREADS) {
            const int g = task / n_mels, m = task % n_mels;
            const int ks = s_mstart[m], kl = s_mlen[m];
            float acc[8];
#pragma unroll
            for (int f = 0; f < 8; f++) acc[f] = 0.0f;
            const float* pp = s_ex + (g * 8) * LM_EF + ks;
            for (int j = 0; j < kl; j++) {
                const float w = s_melw[m * LM_MW + j];
#pragma unroll
                for (int f = 0; f < 8; f++) acc[f] = fmaf(w, pp[f * LM_EF + j], acc[f]);
            }
            const i64 tbase = t0 + g * 8;
            float* orow = outb + (size_t)m * (size_t)T + tbase;
#pragma unroll
            for (int f = 0; f < 8; f++) {
                if (tbase + f < T) {
                    float lg = __log2f(fmaxf(acc[f], 1e-10f)) * 0.30102999566398120f;
                    run_max = fmaxf(run_max, lg);
                    float sv = (lg + 4.0f) * 0.25f;
                    tmin = fminf(tmin, sv);
                    orow[f] = sv;
                }
            }
        }
        // per-tile minimum (lets mel_floor skip tiles that need no clamping)
        tmin = warp_reduce_min_f(tmin);
        if (lane == 0) s_red[warp] = tmin;
        __syncthreads();
        if (tid == 0) {
            float m0 = fminf(fminf(s_red[0], s_red[1]), fminf(s_red[2], s_red[3]));
            p.tile_min[(size_t)b * (size_t)p.tiles_cap + (size_t)tile] = m0;
        }
        if (p.per_clip) {
            float bm = warp_reduce_max_f(run_max);
            if (lane == 0 && bm > -1.0e38f) atomicMax(p.gmax_key + b, float_to_key(bm));
            run_max = -3.0e38f;
        }
    }
    if (!p.per_clip) {
        float bm = warp_reduce_max_f(run_max);
        if (lane == 0 && bm > -1.0e38f) atomicMax(p.gmax_key, float_to_key(bm));
    }
}

// K5: out = max(out, ((gmax - 8) + 4) / 4) in place; tiles whose minimum already clears the floor are skipped.
__global__ void __launch_bounds__(256) mel_floor_kernel(LogMelParams p) {
    i64 n_act = p.n;
    if (p.d_n) { n_act = *p.d_n; if (n_act > p.n) n_act = p.n; if (n_act < 0) n_act = 0; }
    const i64 T = (n_act + p.padding) / kHop;
    const i64 tiles = (T + LM_FRAMES - 1) / LM_FRAMES;
    const int n_mels = p.n_mels;
    for (i64 work = blockIdx.x; work < tiles * p.batch; work += gridDim.x) {
        const int b = (int)(work / tiles);
        const i64 tile = work % tiles;
        const float gmax = key_to_float(p.gmax_key[p.per_clip ? b : 0]);
        const float floor_v = ((gmax - 8.0f) + 4.0f) * 0.25f;
        if (p.tile_min[(size_t)b * (size_t)p.tiles_cap + (size_t)tile] >= floor_v) continue;
        const i64 t0 = tile * LM_FRAMES;
        float* outb = p.out + (size_t)b * (size_t)n_mels * (size_t)T;
        for (int e = threadIdx.x; e < n_mels * LM_FRAMES; e += blockDim.x) {
            const int m = e / LM_FRAMES, f = e % LM_FRAMES;
            if (t0 + f < T) {
                float* q = outb + (size_t)m * (size_t)T + t0 + f;
                float v = *q;
                if (v < floor_v) *q = floor_v;
            }
        }
    }
}

__global__ void logmel_init_kernel(int* gmax_key, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) gmax_key[i] = float_to_key(-3.0e38f);
}

static size_t logmel_smem_bytes() {
    size_t words = (size_t)LM_TILE_WORDS + (size_t)LM_FRAMES * LM_EF + kNFFT + 2 * 200 + 2 * 202 +
                   (size_t)kMelMaxMels * LM_MW + 2 * kMelMaxMels + 8;
    return words * 4;
}

static i64 logmel_tiles_cap(i64 n, i64 padding) {
    i64 T = (n + padding) / kHop;
    return (T + LM_FRAMES - 1) / LM_FRAMES;
}

// workspace layout: [gmax keys: batch ints, padded to 256 B][tile_min: batch * tiles_cap floats]
size_t logmel_workspace_bytes(i64 batch, i64 n, i64 padding) {
    return align_up((size_t)batch * 4, 256) + align_up((size_t)batch * (size_t)logmel_tiles_cap(n, padding) * 4, 256) + 256;
}

int logmel_launch(const void* d_audio, int fmt, i64 batch, i64 n, i64 row_stride, const i64* d_n, i64 padding,
                  int n_mels, int norm_mode, float* d_out, i64* d_frames_out, void* d_ws, size_t ws_bytes,
                  cudaStream_t stream) {
    if (!d_audio || !d_out || !d_ws) { set_error("log_mel: null pointer"); return B2A_EINVAL; }
    if (fmt != B2A_FMT_S16 && fmt != B2A_FMT_F32) { set_error("log_mel: bad fmt %d", fmt); return B2A_EINVAL; }
    if (batch <= 0 || n < 0 || padding < 0 || row_stride < n) { set_error("log_mel: bad shape"); return B2A_EINVAL; }
    if (d_n && batch != 1) { set_error("log_mel: device-side length needs batch == 1"); return B2A_EINVAL; }
    if (!d_n && n + padding <= 200) { set_error("log_mel: need more than 200 samples (reflect pad), got %lld", (long long)(n + padding)); return B2A_EINVAL; }
    if (norm_mode != B2A_NORM_WHISPER && norm_mode != B2A_NORM_PER_CLIP) { set_error("log_mel: bad norm_mode"); return B2A_EINVAL; }
    if (ws_bytes < logmel_workspace_bytes(batch, n, padding)) { set_error("log_mel: workspace too small"); return B2A_EWORKSPACE; }
    const LogMelTables* tab = get_logmel_tables(n_mels);
    if (!tab) return B2A_EINVAL;

    LogMelParams p;
    p.audio = d_audio; p.fmt = fmt; p.row_stride = row_stride; p.n = n; p.d_n = d_n; p.padding = padding;
    p.n_mels = n_mels; p.batch = (int)batch; p.out = d_out; p.d_frames_out = d_frames_out;
    p.gmax_key = (int*)d_ws;
    p.per_clip = (norm_mode == B2A_NORM_PER_CLIP);
    p.tile_min = (float*)((char*)d_ws + align_up((size_t)batch * 4, 256));
    p.tiles_cap = logmel_tiles_cap(n, padding);
    p.tab = tab;

    i64 work = p.tiles_cap * batch;
    if (work <= 0) { if (d_frames_out) cudaMemsetAsync(d_frames_out, 0, 8, stream); return B2A_OK; }
    int nkeys = (int)batch;
    auto kinit = logmel_init_kernel;
    B2A_LAUNCH(kinit, (nkeys + 255) / 256, 256, 0, stream, p.gmax_key, nkeys);
    size_t smem = logmel_smem_bytes();
    auto k4 = stft_mel_kernel;
    static thread_local bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(stft_mel)");
        attr_set = true;
    }
    i64 grid = work < 148 * 3 * 2 ? work : 148 * 3 * 2;   // persistent: 3 CTAs/SM x 2 rounds of 148 SMs
    B2A_LAUNCH(k4, (unsigned)grid, LM_THREADS, smem, stream, p);
    B2A_CHECK_LAUNCH("stft_mel_kernel");
    i64 grid5 = work < 148 * 8 ? work : 148 * 8;
    auto k5 = mel_floor_kernel;
    B2A_LAUNCH(k5, (unsigned)grid5, 256, 0, stream, p);
    B2A_CHECK_LAUNCH("mel_floor_kernel");
    return B2A_OK;
}

}  // namespace b2a

extern "C" {

size_t b2a_log_mel_workspace_bytes(int64_t batch, int64_t n, int64_t padding) {
    if (batch <= 0 || n < 0 || padding < 0) return 0;
    return b2a::logmel_workspace_bytes(batch, n, padding);
}

int b2a_log_mel(const void* d_audio, int fmt, int64_t batch, int64_t n, int64_t row_stride, const int64_t* d_n,
                int64_t padding, int n_mels, int norm_mode, float* d_out, int64_t* d_frames_out, void* d_ws,
                size_t ws_bytes, b2a_stream_t stream) {
    return b2a::logmel_launch(d_audio, fmt, batch, n, row_stride, (const b2a::i64*)d_n, padding, n_mels, norm_mode,
                              d_out, (b2a::i64*)d_frames_out, d_ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"
