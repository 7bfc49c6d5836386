// fir_fast_44100_s16x1.cu — fast FIR instantiation: 44100 Hz -> 16 kHz, s16x1 input (see resample_fast.cuh)
#include "resample_fast.cuh"
B2A_DEFINE_FIR_FAST(fir_fast_run_44100_s16x1, 44100, B2A_FMT_S16, 1)
