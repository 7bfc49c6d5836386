// fir_umma_48000.cu — tcgen05 FIR instantiation: 48 kHz s16 stereo -> 16 kHz mono (see fir_umma.cuh)
#include "fir_umma.cuh"
namespace b2a {
int fir_umma_run_48000(const void* d_in, i64 n_in, int16_t* d_out_s16, u64* d_energy, FirMmaPlan* plan, cudaStream_t stream) {
    return fir_umma_launch<48000>(d_in, n_in, d_out_s16, d_energy, plan, stream);
}
}  // namespace b2a
