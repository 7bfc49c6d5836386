// b2a_host.cu — host side of libb2a: error slot, table designs (double precision), device table cache.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "b2a_tables.cuh"
#include "f16_bits.h"
#include "fir_mma.cuh"
#include "fir_tc_common.cuh"
#include "fir_design.h"
#include "mel_design.h"

namespace b2a {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return B2A_ECUDA;
}

// resampler design lives in fir_design.h
static long long gcd_ll(long long a, long long b) { return b2a_design::gcd_ll(a, b); }

void design_resampler_host(int in_rate, int out_rate, int* L_out, int* M_out, int* taps_out, float** h_taps_out) {
    b2a_design::design_resampler(in_rate, out_rate, L_out, M_out, taps_out, h_taps_out);
}

// ---- slaney mel filterbank (whisper mel_filters.npz): design in mel_design.h, shared with tools/gen_mel_tables.cpp
void design_mel_host(int n_mels, float* filters) { b2a_design::design_mel(n_mels, kNBins, (double)kSampleRate, filters); }

// ---- device table cache ------------------------------------------------------------------------
static std::mutex g_mu;
static std::map<std::pair<int, int>, LogMelTables*> g_logmel;                       // (device, n_mels)
static std::map<std::pair<int, std::pair<int, int>>, ResampleDesign*> g_resample;   // (device, (in, out))

const LogMelTables* get_logmel_tables(int n_mels) {
    if (n_mels != 80 && n_mels != 128) { set_error("n_mels must be 80 or 128 (got %d)", n_mels); return nullptr; }
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return nullptr; }
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_pair(dev, n_mels);
    auto it = g_logmel.find(key);
    if (it != g_logmel.end()) return it->second;

    LogMelTables* h = (LogMelTables*)calloc(1, sizeof(LogMelTables));
    const double kPi = 3.14159265358979323846;
    for (int n = 0; n < kNFFT; n++) h->win[n] = (float)(0.5 - 0.5 * std::cos(2.0 * kPi * n / kNFFT));
    for (int n2 = 0; n2 < 20; n2++)
        for (int k1 = 0; k1 < 10; k1++) {
            double a = -2.0 * kPi * (double)(n2 * k1) / 200.0;
            h->tw200[n2 * 10 + k1] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
    for (int k = 0; k < kNBins; k++) {
        double a = 2.0 * kPi * (double)k / 400.0;
        h->tw400[k] = make_float2((float)std::cos(a), (float)std::sin(a));
    }
    std::vector<float> filt((size_t)n_mels * kNBins);
    design_mel_host(n_mels, filt.data());
    h->n_mels = n_mels;
    for (int m = 0; m < n_mels; m++) {
        int first = -1, last = -1;
        for (int k = 0; k < kNBins; k++)
            if (filt[(size_t)m * kNBins + k] != 0.0f) { if (first < 0) first = k; last = k; }
        if (first < 0) { first = 0; last = -1; }
        int len = last - first + 1;
        if (len > kMelMaxWidth) { free(h); set_error("mel filter %d wider than %d bins", m, kMelMaxWidth); return nullptr; }
        h->mel_start[m] = first;
        h->mel_len[m] = len;
        for (int j = 0; j < len; j++) h->mel_w[m * kMelMaxWidth + j] = filt[(size_t)m * kNBins + first + j];
    }
    LogMelTables* d = nullptr;
    e = cudaMalloc((void**)&d, sizeof(LogMelTables));
    if (e == cudaSuccess) e = cudaMemcpy(d, h, sizeof(LogMelTables), cudaMemcpyHostToDevice);
    free(h);
    if (e != cudaSuccess) { cuda_fail(e, "log-mel table upload"); return nullptr; }
    g_logmel[key] = d;
    return d;
}

#if defined(B2A_PROFILE) || defined(B2A_EMU)
// ---- basis bank + window tables of the tensor-core log-mel PROBE (tools/probes/logmel_tc.cuh; not in the release library) ----
// Eight f16 planes (product p = even-cos, even-sin, odd-cos, odd-sin; hi = f16(c), lo = f16(c - hi)) in the canonical
// no-swizzle K-major UMMA layout: element (k = n, row = j) of a plane at (n / 8) * kTcLbo + (j / 8) * 128 + (j % 8) * 16 +
// (n % 8) * 2 bytes, 13 x 13 blocks of 8 x 8 stored (n, j <= 103; the MMAs' K = N = 112 read one block further, which aliases
// the next chunk / plane: finite values against zero operands).  Rows n = 100 of the even-cos and odd-sin products carry
// 1/2 because the fold counts x[100] and x[300] twice.  Behind the planes: 1 792 zero bytes, then the four window
// tables w[n], w[n+200], w[200-n], w[400-n] (n = 0..111, zero for n > 100 and for the partners of n = 0) times 1/4:
// the kernel's spectra live in the (s16 / 4) domain so that every folded value fits the f16 range.
static std::map<int, const unsigned char*> g_logmel_tc;      // device -> blob

const unsigned char* get_logmel_tc_blob() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return nullptr; }
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_logmel_tc.find(dev);
    if (it != g_logmel_tc.end()) return it->second;
    const int kLbo = 13 * 128, kPlane = 13 * kLbo, kBank = 8 * kPlane + kLbo + 128, kBlob = kBank + 4 * 112 * 4;
    std::vector<unsigned char> h((size_t)kBlob, 0);
    const double kPi = 3.14159265358979323846;
    for (int pr = 0; pr < 4; pr++)
        for (int n = 0; n < 104; n++)
            for (int j = 0; j < 104; j++) {
                double c = 0.0;
                // angles as exact multiples of 2 pi / 400
                const int m_even = (int)(((long long)2 * j * n) % 400), m_odd = (int)(((long long)(2 * j + 1) * n) % 400);
                if (pr == 0 && n <= 100 && j <= 100) c = std::cos(2.0 * kPi * m_even / 400.0) * (n == 100 ? 0.5 : 1.0);
                if (pr == 1 && n >= 1 && n <= 99 && j >= 1 && j <= 99) c = std::sin(2.0 * kPi * m_even / 400.0);
                if (pr == 2 && n <= 99 && j <= 99) c = std::cos(2.0 * kPi * m_odd / 400.0);
                if (pr == 3 && n >= 1 && n <= 100 && j <= 99) c = std::sin(2.0 * kPi * m_odd / 400.0) * (n == 100 ? 0.5 : 1.0);
                const uint16_t hi = b2a_f16::f32_to_f16((float)c);
                const uint16_t lo = b2a_f16::f32_to_f16((float)(c - (double)b2a_f16::f16_to_f32(hi)));
                const size_t off = (size_t)(n / 8) * kLbo + (size_t)(j / 8) * 128 + (size_t)(j % 8) * 16 + (size_t)(n % 8) * 2;
                memcpy(&h[(size_t)(2 * pr) * kPlane + off], &hi, 2);
                memcpy(&h[(size_t)(2 * pr + 1) * kPlane + off], &lo, 2);
            }
    float* win = (float*)&h[(size_t)kBank];
    auto hann = [&](int n) { return (float)(0.5 - 0.5 * std::cos(2.0 * kPi * (double)(n % 400) / 400.0)); };   // torch.hann_window(400), periodic, f32
    for (int n = 0; n <= 100; n++) {
        win[0 * 112 + n] = 0.25f * hann(n);
        win[1 * 112 + n] = 0.25f * hann(n + 200);
        win[2 * 112 + n] = n == 0 ? 0.0f : 0.25f * hann(200 - n);
        win[3 * 112 + n] = n == 0 ? 0.0f : 0.25f * hann(400 - n);
    }
    unsigned char* d = nullptr;
    e = cudaMalloc((void**)&d, h.size());
    if (e == cudaSuccess) e = cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cuda_fail(e, "log-mel basis upload"); return nullptr; }
    g_logmel_tc[dev] = d;
    return d;
}
#endif

const ResampleDesign* get_resample_design(int in_rate, int out_rate) {
    if (in_rate <= 0 || out_rate <= 0) { set_error("bad sample rate %d -> %d", in_rate, out_rate); return nullptr; }
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return nullptr; }
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_pair(dev, std::make_pair(in_rate, out_rate));
    auto it = g_resample.find(key);
    if (it != g_resample.end()) return it->second;
    ResampleDesign* r = (ResampleDesign*)calloc(1, sizeof(ResampleDesign));
    float* h = nullptr;
    design_resampler_host(in_rate, out_rate, &r->L, &r->M, &r->taps, &h);
    if ((size_t)r->L * r->taps > (size_t)(1 << 22)) {
        free(h); free(r);
        set_error("rate pair %d -> %d needs %d phases: unsupported", in_rate, out_rate, r->L);
        return nullptr;
    }
    r->in_rate = in_rate; r->out_rate = out_rate;
    r->center = (r->taps - 1) / 2;
    r->h_taps = h;
    float* d = nullptr;
    size_t bytes = sizeof(float) * (size_t)r->L * r->taps;
    e = cudaMalloc((void**)&d, bytes);
    if (e == cudaSuccess) e = cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { free(h); free(r); cuda_fail(e, "resampler tap upload"); return nullptr; }
    r->d_taps = d;
    g_resample[key] = r;
    return r;
}

// ---- filter bank of the tensor-core FIR (fir_mma.cuh) as mma.m16n8k16 B fragments ------------------------------
// layout [block b][k-step s][8-output half nt][term: 0 = T_hi, 1 = T_lo][lane] -> uint2 {b0, b1};
// lane (g = lane >> 2, t = lane & 3): b0 = k-slots (2t, 2t+1), b1 = k-slots (2t+8, 2t+9) of column n = g.  The
// kernel permutes the k-slots of a k-step so that lane t owns input frames 4t..4t+3 (one 8-byte A load per row):
// slot 2t, 2t+1, 2t+8, 2t+9 <-> frame offsets f = 4t, 4t+1, 4t+2, 4t+3, and
// B[f][n] = 2^12 * tap[phase(J)][kb + 16 s + f - (J M)/L],  J = 16 b + 8 nt + n  (zero outside the filter).
template <int IN_RATE>
static void build_fir_mma_table(const float* taps /*[L][TAPS]*/, std::vector<uint2>& out) {
    using G = FirMmaGeom<IN_RATE, 2>;       // tap geometry does not depend on the channel count
    out.assign((size_t)kFmBlocks * G::KS * 2 * 2 * 32, make_uint2(0u, 0u));
    for (int b = 0; b < kFmBlocks; b++)
        for (int s = 0; s < G::KS; s++)
            for (int nt = 0; nt < 2; nt++)
                for (int lane = 0; lane < 32; lane++) {
                    const int g = lane >> 2, t = lane & 3;
                    const int J = 16 * b + 8 * nt + g;
                    const int BJ = (J * G::M) / G::L, ph = (J * G::M) % G::L;
                    uint16_t hi[4], lo[4];
                    for (int e = 0; e < 4; e++) {
                        const int f = 4 * t + e;
                        const int i = G::kb(b) + 16 * s + f - BJ;
                        float T = 0.0f;
                        if (i >= 0 && i < G::TAPS) T = taps[(size_t)ph * G::TAPS + i] * (float)(1 << kFmTapShift);
                        hi[e] = b2a_f16::f32_to_f16(T);
                        lo[e] = b2a_f16::f32_to_f16(T - b2a_f16::f16_to_f32(hi[e]));
                    }
                    const size_t base = ((((size_t)b * G::KS + s) * 2 + nt) * 2) * 32 + lane;
                    out[base] = make_uint2((unsigned)hi[0] | ((unsigned)hi[1] << 16), (unsigned)hi[2] | ((unsigned)hi[3] << 16));
                    out[base + 32] = make_uint2((unsigned)lo[0] | ((unsigned)lo[1] << 16), (unsigned)lo[2] | ((unsigned)lo[3] << 16));
                }
}

static std::map<std::pair<int, int>, const uint2*> g_fir_mma;   // (device, in_rate)

const uint2* get_fir_mma_table(int in_rate) {
    const ResampleDesign* des = get_resample_design(in_rate, kSampleRate);
    if (!des) return nullptr;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return nullptr; }
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_pair(dev, in_rate);
    auto it = g_fir_mma.find(key);
    if (it != g_fir_mma.end()) return it->second;
    std::vector<uint2> h;
    if (in_rate == 44100) build_fir_mma_table<44100>(des->h_taps, h);
    else if (in_rate == 48000) build_fir_mma_table<48000>(des->h_taps, h);
    else { set_error("no tensor-core FIR table for %d Hz", in_rate); return nullptr; }
    uint2* d = nullptr;
    e = cudaMalloc((void**)&d, h.size() * sizeof(uint2));
    if (e == cudaSuccess) e = cudaMemcpy(d, h.data(), h.size() * sizeof(uint2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cuda_fail(e, "FIR table upload"); return nullptr; }
    g_fir_mma[key] = d;
    return d;
}


// ---- filter bank of the tcgen05 FIR (fir_tc_common.cuh, fir_tmem.cuh) as UMMA B operand tiles ---------------------------------------
// [class c][block b][k-step s] -> one [N = 32][K = 16] f16 tile in the canonical no-swizzle K-major layout
// (element (n, k) at (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2 bytes): rows 0-15 = T_hi, rows 16-31 =
// T_lo of output J = 16 b + n % 16.  Plane column j of a class-c row is input frame S run - CENTER - shift(c) + j, so
// column k of the tile (j = kbp(b) + 16 s + k) carries
// T = 2^12 * tap[phase(J)][j - shift(c) - (J DEC) / L]  (zero outside the filter).
template <int IN_RATE>
static void build_fir_umma_table(const float* taps /*[L][TAPS]*/, std::vector<unsigned char>& out) {
    using G = FirUmmaGeom<IN_RATE>;
    out.assign((size_t)kFuClasses * G::B_BYTES, 0);
    for (int c = 0; c < kFuClasses; c++)
        for (int b = 0; b < G::BBLOCKS; b++)
            for (int s = 0; s < G::KS; s++)
                for (int n = 0; n < 32; n++)
                    for (int k = 0; k < 16; k++) {
                        const int J = 16 * b + (n & 15);
                        const int BJ = (J * G::DEC) / G::L, ph = (J * G::DEC) % G::L;
                        const int i = G::kbp(b) + 16 * s + k - G::shift(kFuRun0, c) - BJ;
                        float T = 0.0f;
                        if (i >= 0 && i < G::TAPS) T = taps[(size_t)ph * G::TAPS + i] * (float)(1 << kFmTapShift);
                        const uint16_t hi = b2a_f16::f32_to_f16(T);
                        const uint16_t lo = b2a_f16::f32_to_f16(T - b2a_f16::f16_to_f32(hi));
                        const uint16_t v = n < 16 ? hi : lo;
                        const size_t off = (size_t)c * G::B_BYTES + ((size_t)b * G::KS + s) * kFuBTile + (n / 8) * kFuBSbo + (k / 8) * kFuBLbo +
                                           (n % 8) * 16 + (k % 8) * 2;
                        memcpy(&out[off], &v, 2);
                    }
}

static std::map<std::pair<int, int>, const uint4*> g_fir_umma;   // (device, in_rate)

const uint4* get_fir_umma_table(int in_rate) {
    const ResampleDesign* des = get_resample_design(in_rate, kSampleRate);
    if (!des) return nullptr;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return nullptr; }
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_pair(dev, in_rate);
    auto it = g_fir_umma.find(key);
    if (it != g_fir_umma.end()) return it->second;
    std::vector<unsigned char> h;
    if (in_rate == 44100) build_fir_umma_table<44100>(des->h_taps, h);
    else if (in_rate == 48000) build_fir_umma_table<48000>(des->h_taps, h);
    else { set_error("no tcgen05 FIR table for %d Hz", in_rate); return nullptr; }
    uint4* d = nullptr;
    e = cudaMalloc((void**)&d, h.size());
    if (e == cudaSuccess) e = cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cuda_fail(e, "FIR table upload"); return nullptr; }
    g_fir_umma[key] = d;
    return d;
}

}  // namespace b2a

template <int IN_RATE>
static int fir_schedule_dump(uint32_t* out, int cap) {
    using G = b2a::FirUmmaGeom<IN_RATE>;
    constexpr b2a::FirUmmaSched S = b2a::FirUmmaSchedOf<IN_RATE>::value;
    if (cap < 16 + S.n) { b2a::set_error("capacity %d < %d", cap, 16 + S.n); return B2A_EINVAL; }
    for (int i = 0; i < 16; i++) out[i] = 0;
    out[0] = (uint32_t)S.n; out[1] = (uint32_t)G::PIECES; out[2] = (uint32_t)G::KS;
    for (int b = 0; b < b2a::kFmBlocks; b++) out[3 + b] = (uint32_t)(G::kbp(b) / 8);
    for (int j = 0; j < S.n; j++) out[16 + j] = S.w[j];
    return 16 + S.n;
}

extern "C" {

int b2a_version(void) { return 100; }  // 0.1.0

const char* b2a_last_error(void) { return b2a::g_err; }

int64_t b2a_launch_count(void) { return (int64_t)b2a::g_launches.load(std::memory_order_relaxed); }

int b2a_fir_schedule(int in_rate, uint32_t* out, int capacity_words) {
    if (!out) { b2a::set_error("bad argument"); return B2A_EINVAL; }
    if (in_rate == 44100) return fir_schedule_dump<44100>(out, capacity_words);
    if (in_rate == 48000) return fir_schedule_dump<48000>(out, capacity_words);
    b2a::set_error("no tcgen05 FIR for %d Hz", in_rate);
    return B2A_EUNSUPPORTED;
}

int b2a_resample_ntaps(int in_rate, int out_rate, int* phases) {
    if (in_rate <= 0 || out_rate <= 0) { b2a::set_error("bad sample rate"); return B2A_EINVAL; }
    int L, M, taps; float* h = nullptr;
    b2a::design_resampler_host(in_rate, out_rate, &L, &M, &taps, &h);
    free(h);
    if (phases) *phases = L;
    return taps;
}

int b2a_resample_taps(int in_rate, int out_rate, float* h_taps, size_t capacity_floats) {
    if (in_rate <= 0 || out_rate <= 0 || !h_taps) { b2a::set_error("bad argument"); return B2A_EINVAL; }
    int L, M, taps; float* h = nullptr;
    b2a::design_resampler_host(in_rate, out_rate, &L, &M, &taps, &h);
    size_t n = (size_t)L * taps;
    if (capacity_floats < n) { free(h); b2a::set_error("capacity %zu < %zu", capacity_floats, n); return B2A_EINVAL; }
    memcpy(h_taps, h, n * sizeof(float));
    free(h);
    return B2A_OK;
}

int b2a_mel_filters(int n_mels, float* h_filters, size_t capacity_floats) {
    if ((n_mels != 80 && n_mels != 128) || !h_filters || capacity_floats < (size_t)n_mels * b2a::kNBins) {
        b2a::set_error("bad argument");
        return B2A_EINVAL;
    }
    b2a::design_mel_host(n_mels, h_filters);
    return B2A_OK;
}

int64_t b2a_resample_out_len(int64_t n_in, int in_rate, int out_rate) {
    return (int64_t)b2a_design::one_shot_out_len(n_in, in_rate, out_rate);   // exactly what one-shot swr_convert + flush returns
}

int64_t b2a_energy_len(int64_t n_out, int out_rate) {
    if (n_out <= 0 || out_rate <= 0 || out_rate % 1000) return 0;
    int spm = out_rate / 1000;
    return (n_out + spm - 1) / spm;
}

int64_t b2a_log_mel_frames(int64_t n, int64_t padding) { return (n + padding) / b2a::kHop; }

}  // extern "C"
