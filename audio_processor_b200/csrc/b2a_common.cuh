// b2a_common.cuh — shared definitions for the sm_100a kernels and the C ABI.
//
// The same sources compile two ways:
//   * nvcc -gencode arch=compute_100a,code=sm_100a  -> libb2a.so (the product; CUDA only)
//   * g++  -DB2A_EMU -include tests/emu/cuda_emu.h  -> tests/emu/libb2a_emu.so (TEST-ONLY fiber
//     emulation used to check index math / integer exactness on a GPU-less box; never loaded
//     by the product path).
#pragma once

#include <cstddef>
#include <cstdint>

#ifdef B2A_EMU
#include "cuda_emu.h"
#define B2A_LAUNCH(kern, grid, block, smem, stream, ...) (b2a::count_launch(), emu::launch(dim3(grid), dim3(block), (size_t)(smem), kern, __VA_ARGS__))
#define B2A_DYN_SMEM(name) unsigned char* name = emu::g_dyn_smem
#else
#include <cuda_runtime.h>
#define B2A_LAUNCH(kern, grid, block, smem, stream, ...) (b2a::count_launch(), kern<<<dim3(grid), dim3(block), (size_t)(smem), stream>>>(__VA_ARGS__))
#define B2A_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#include "../../include/b2a.h"

namespace b2a {

void count_launch();   // kernel-launch counter behind b2a_launch_count()

typedef unsigned long long u64;
typedef long long i64;

constexpr int kSampleRate = 16000;   // whisper.audio.SAMPLE_RATE
constexpr int kNFFT = 400;           // whisper.audio.N_FFT
constexpr int kHop = 160;            // whisper.audio.HOP_LENGTH
constexpr int kNBins = 201;          // 1 + N_FFT/2

// thread-local error message (b2a_last_error)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define B2A_CHECK_LAUNCH(what)                                  \
    do {                                                        \
        cudaError_t e__ = cudaGetLastError();                   \
        if (e__ != cudaSuccess) return b2a::cuda_fail(e__, what); \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- small device helpers --------------------------------------------------------------
__device__ __forceinline__ int warp_reduce_sum_i(int v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_reduce_max_f(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_reduce_min_f(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// monotone float <-> int key so that integer atomicMax/Min order floats (incl. negatives)
__device__ __forceinline__ int float_to_key(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : (i ^ 0x7fffffff);
}
__device__ __forceinline__ float key_to_float(int k) {
    return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff));
}

// clip(rint(v)) to int16 exactly like libswresample's lrintf + av_clip_int16
__device__ __forceinline__ int quant_s16(float v) {
    int q = __float2int_rn(v);
    return max(-32768, min(32767, q));
}

}  // namespace b2a
