// fir_design.h — libswresample's default resampling filter, restated (host, double precision).
//
// Reference behaviour: FFmpeg libswresample (resample.c build_filter) as driven by
//   ffmpeg -ar 16000 -ac 1 -c:a pcm_s16le      (/root/reference/app/services/audio_processor.py:912-920)
// with no resampler options => filter_size 32, phase_shift 10, exact_rational, cutoff 0.97,
// Kaiser window beta 9.  See SURVEY.md A.1 and oracle/resample_oracle.py (float64 restatement,
// pinned against the real library).  Header-only: the runtime library builds the float taps (and from them the tensor-core filter tables) once per
// rate pair.
#pragma once
#include <cmath>
#include <cstdlib>
#include <vector>

namespace b2a_design {

static inline long long gcd_ll(long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; }

static inline double bessel_i0(double x) {
    // power series sum_k ((x/2)^(2k) / (k!)^2); x <= 9 here, converges in < 40 terms
    double q = x * x * 0.25, term = 1.0, sum = 1.0;
    for (int k = 1; k < 200; k++) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < sum * 1e-18) break;
    }
    return sum;
}

// h_taps_out: malloc'ed float[L][taps]; every phase normalised to unit DC gain.
static inline void design_resampler(int in_rate, int out_rate, int* L_out, int* M_out, int* taps_out, float** h_taps_out) {
    const double kCutoff = 0.97, kBeta = 9.0;
    const int kFilterSize = 32;
    long long g = gcd_ll(in_rate, out_rate);
    int L = (int)(out_rate / g), M = (int)(in_rate / g);
    double factor = std::fmin((double)out_rate * kCutoff / (double)in_rate, 1.0);
    int taps = (int)std::ceil(kFilterSize / factor);
    taps = (taps + 1) & ~1;                                   // == resampler_taps(in_rate, out_rate)
    int center = (taps - 1) / 2;
    float* h = (float*)malloc(sizeof(float) * (size_t)L * taps);
    std::vector<double> row(taps);
    const double kPi = 3.14159265358979323846;
    for (int ph = 0; ph < L; ph++) {
        double norm = 0.0;
        for (int i = 0; i < taps; i++) {
            double x = kPi * ((double)(i - center) - (double)ph / (double)L) * factor;
            double y = (x == 0.0) ? 1.0 : std::sin(x) / x;
            double w = 2.0 * x / (factor * taps * kPi);
            double t = 1.0 - w * w;
            y *= bessel_i0(kBeta * std::sqrt(t > 0.0 ? t : 0.0));
            row[i] = y;
            norm += y;
        }
        for (int i = 0; i < taps; i++) h[(size_t)ph * taps + i] = (float)(row[i] / norm);
    }
    *L_out = L; *M_out = M; *taps_out = taps; *h_taps_out = h;
}

// taps per phase of the default design for a rate pair (ceil(32 / factor) rounded up to even)
static inline int resampler_taps(int in_rate, int out_rate) {
    double factor = std::fmin((double)out_rate * 0.97 / (double)in_rate, 1.0);
    int taps = (int)std::ceil(32 / factor);
    return (taps + 1) & ~1;
}

// Number of samples a one-shot swr_convert(all input) followed by a flush returns — what `ffmpeg -i IN -ar out_rate`
// writes.  Restates libswresample's buffering (swresample.c resample(), resample.c invert_initial_buffer / swri_resample /
// resample_flush): the stream opens with `center` reflected samples; the first call emits every output whose taps-long
// window fits, m*M < (n_in - taps/2)*L; the flush appends R = (min(buffered, taps) + 1) / 2 mirrored samples, `buffered`
// being what the first call left unconsumed, and emits what fits then: ceil((n_in - taps/2 + R) * L / M).  R is taps/2 or
// one less, which is why ceil(n_in*L/M) overshoots by one on about a third of the lengths.  Inputs shorter than taps + 1
// wait in the initial buffer until the flush (buffered = n_in).  oracle/resample_oracle.py:out_len is the same model,
// checked against the real library on 25 030 cases.
static inline long long one_shot_out_len(long long n_in, int in_rate, int out_rate) {
    if (n_in <= 0 || in_rate <= 0 || out_rate <= 0) return 0;
    const long long g = gcd_ll(in_rate, out_rate);
    const long long L = out_rate / g, M = in_rate / g;
    if (L == 1 && M == 1) return n_in;
    const long long taps = resampler_taps(in_rate, out_rate), center = (taps - 1) / 2;
    auto cdiv = [](__int128 a, long long b) -> long long { return (long long)(a >= 0 ? (a + b - 1) / b : -((-a) / b)); };
    long long buffered;
    if (n_in < taps + 1) buffered = n_in;
    else {
        long long n1 = cdiv((__int128)(n_in - taps / 2) * L, M);
        if (n1 < 0) n1 = 0;
        buffered = center + n_in - (long long)(((__int128)n1 * M) / L);
    }
    const long long refl = ((buffered < taps ? buffered : taps) + 1) / 2;
    if (n_in < taps + 1 && n_in + refl < taps + 1) return 0;
    const long long n = cdiv((__int128)(n_in - taps / 2 + refl) * L, M);
    return n > 0 ? n : 0;
}

}  // namespace b2a_design
