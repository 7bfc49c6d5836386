// fir_fast_48000_s16x2.cu — fast FIR instantiation: 48000 Hz -> 16 kHz, s16x2 input (see resample_fast.cuh)
#include "resample_fast.cuh"
B2A_DEFINE_FIR_FAST(fir_fast_run_48000_s16x2, 48000, B2A_FMT_S16, 2)
