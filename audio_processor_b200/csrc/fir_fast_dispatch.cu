// fir_fast_dispatch.cu — picks the compile-time-tap FIR instantiation for a (rate, format, channels) triple.
#include "resample_fast.cuh"
#include "fir_mma.cuh"

namespace b2a {

#define B2A_DECL(NAME) int NAME(const void*, i64, i64, int16_t*, float*, u64*, FirFastPlan*, cudaStream_t)
B2A_DECL(fir_fast_run_44100_s16x1);
B2A_DECL(fir_fast_run_48000_s16x1);
#undef B2A_DECL
// stereo s16 at the two named rates: tensor-core kernel (fir_mma.cuh); s16 + energy outputs only
int fir_mma_run_44100(const void*, i64, int16_t*, u64*, FirMmaPlan*, cudaStream_t);
int fir_mma_run_48000(const void*, i64, int16_t*, u64*, FirMmaPlan*, cudaStream_t);

// returns 1 if a fast kernel was launched (plan filled), 0 if this input has no fast path, <0 on error
int fir_fast_dispatch(int in_rate, int fmt, int channels, const void* d_in, i64 n_in, i64 n_out, int16_t* d_out_s16,
                      float* d_out_f32, u64* d_energy, FirFastPlan* plan, cudaStream_t stream) {
    plan->set_first = plan->set_count = 0;
    plan->out_lo = plan->out_hi = 0;
    plan->energy_atomic = false;
    if (fmt != B2A_FMT_S16) return 0;       // float input goes through the table-driven kernel
    if (channels == 2 && (in_rate == 44100 || in_rate == 48000)) {
        if (d_out_f32) return 0;              // pre-quantisation float output: table-driven kernel
        FirMmaPlan tp;
        int rc = in_rate == 44100 ? fir_mma_run_44100(d_in, n_in, d_out_s16, d_energy, &tp, stream)
                                  : fir_mma_run_48000(d_in, n_in, d_out_s16, d_energy, &tp, stream);
        if (rc <= 0) return rc;
        plan->set_first = 1; plan->set_count = 1;         // "something was launched"
        plan->out_lo = tp.out_lo; plan->out_hi = tp.out_hi;
        return 1;
    }
    if (in_rate == 44100 && channels == 1) return fir_fast_run_44100_s16x1(d_in, n_in, n_out, d_out_s16, d_out_f32, d_energy, plan, stream);
    if (in_rate == 48000 && channels == 1) return fir_fast_run_48000_s16x1(d_in, n_in, n_out, d_out_s16, d_out_f32, d_energy, plan, stream);
    return 0;
}

}  // namespace b2a
