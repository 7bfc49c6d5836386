// fir_mma_44100.cu — tensor-core FIR instantiations: 44.1 kHz s16 (stereo, mono) -> 16 kHz mono (see fir_mma.cuh)
#include "fir_mma.cuh"
namespace b2a {
int fir_mma_run_44100(int channels, const void* d_in, i64 n_in, int16_t* d_out_s16, u64* d_energy, FirMmaPlan* plan, cudaStream_t stream, i64 first_tile) {
    return channels == 2 ? fir_mma_launch<44100, 2>(d_in, n_in, d_out_s16, d_energy, plan, stream, first_tile)
                         : fir_mma_launch<44100, 1>(d_in, n_in, d_out_s16, d_energy, plan, stream, first_tile);
}
}  // namespace b2a
