// fir_tmem_48000.cu — TMA-fed tcgen05 FIR instantiation: 48 kHz s16 stereo -> 16 kHz mono (see fir_tmem.cuh)
#include "fir_tmem.cuh"
namespace b2a {
i64 fir_tmem_plan_48000(i64 n_in, FirMmaPlan* plan) { return fir_tmem_plan<48000>(n_in, plan); }
int fir_tmem_run_48000(const void* d_in, i64 n_in, int16_t* d_out_s16, u64* d_energy, FirMmaPlan* plan, const GenericParams* edge, cudaStream_t stream) {
    return fir_tmem_launch<48000>(d_in, n_in, d_out_s16, d_energy, plan, edge, stream);
}
}  // namespace b2a
