// b2a_tables.cuh — immutable lookup tables (built once per process+device, host double -> f32)
#pragma once
#include "b2a_common.cuh"

namespace b2a {

constexpr int kMelMaxWidth = 16;   // >= widest slaney triangle in FFT bins (14 @80 mels, 9 @128)
constexpr int kMelMaxMels = 128;

// Everything the log-mel kernel needs besides the audio.  One blob in global memory;
// each (persistent) block stages it into shared memory once.
struct LogMelTables {
    float win[kNFFT];                     // periodic Hann, torch.hann_window(400)
    float2 tw200[200];                    // [n2*10 + k1] = exp(-2*pi*i*n2*k1/200)   (stage-1 twiddles)
    float2 tw400[kNBins];                 // [k] = (cos, sin)(2*pi*k/400)            (real-FFT unpack)
    int mel_start[kMelMaxMels];           // first FFT bin with non-zero weight
    int mel_len[kMelMaxMels];             // number of consecutive non-zero bins
    float mel_w[kMelMaxMels * kMelMaxWidth];
    int n_mels;
    int pad_[3];
};

// Polyphase filter bank for one (in_rate, out_rate) pair, libswresample default design.
struct ResampleDesign {
    int in_rate, out_rate;
    int L, M;        // out = in * L / M, L phases
    int taps, center;
    const float* d_taps;   // device [L][taps]
    const float* h_taps;   // host copy
};

// host-side designs (double precision), see b2a_host.cu
void design_resampler_host(int in_rate, int out_rate, int* L, int* M, int* taps, float** h_taps_out);
void design_mel_host(int n_mels, float* filters /*[n_mels][201]*/);

// cached device tables for the current device (thread-safe); nullptr + error set on failure
const LogMelTables* get_logmel_tables(int n_mels);
const ResampleDesign* get_resample_design(int in_rate, int out_rate);

}  // namespace b2a
