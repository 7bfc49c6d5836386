// mel_design.h — the slaney mel filterbank Whisper ships as assets/mel_filters.npz, regenerated (host, double).
//
// Reference: whisper.audio.mel_filters -> librosa.filters.mel(sr=16000, n_fft=400, n_mels) (slaney scale, slaney
// norm), reached from model.transcribe at /root/reference/app/services/audio_processor.py:1076-1080.  Restated from
// the published definition (see oracle/whisper_logmel.py, cross-checked there against transformers' mel_filter_bank).
// Header-only so that the build-time table generator (tools/gen_mel_tables.cpp) and the runtime library
// (b2a_mel_filters) produce identical float weights.
#pragma once
#include <cmath>
#include <vector>

namespace b2a_design {

static inline double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static inline double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// filters: float [n_mels][n_bins], n_bins = 1 + n_fft/2
static inline void design_mel(int n_mels, int n_bins, double sample_rate, float* filters) {
    std::vector<double> fftf(n_bins), melf(n_mels + 2);
    for (int k = 0; k < n_bins; k++) fftf[k] = (sample_rate / 2.0) * (double)k / (double)(n_bins - 1);
    const double m0 = hz_to_mel(0.0), m1 = hz_to_mel(sample_rate / 2.0);
    for (int i = 0; i < n_mels + 2; i++) melf[i] = mel_to_hz(m0 + (m1 - m0) * (double)i / (double)(n_mels + 1));
    for (int i = 0; i < n_mels; i++) {
        const double fd0 = melf[i + 1] - melf[i], fd1 = melf[i + 2] - melf[i + 1];
        const double enorm = 2.0 / (melf[i + 2] - melf[i]);
        for (int k = 0; k < n_bins; k++) {
            const double lower = (fftf[k] - melf[i]) / fd0;
            const double upper = (melf[i + 2] - fftf[k]) / fd1;
            const double w = std::fmax(0.0, std::fmin(lower, upper));
            filters[(size_t)i * n_bins + k] = (float)(w * enorm);
        }
    }
}

}  // namespace b2a_design
