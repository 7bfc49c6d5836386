// f32x2_probe.cu — issue cost of packed FP32 (FFMA2 / FADD2) against scalar FFMA / FADD on sm_100a.
// Each thread runs ITER iterations over 8 independent complex accumulators; variant 0: 2 scalar ops per complex value,
// variant 1: 1 packed op.  Prints cycles per iteration per warp (4 warps per SM sub-partition resident, 148 x 4 blocks).
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { return (u64)__float_as_uint(a) | ((u64)__float_as_uint(b) << 32); }
template <int V, int MIX> __global__ void __launch_bounds__(512) k(float* out, int iters, float s, long long* cyc) {
    float ar[8], ai[8];
    u64 ap[8];
    unsigned ix[8] = {1, 2, 3, 4, 5, 6, 7, 8};
    for (int j = 0; j < 8; j++) { ar[j] = threadIdx.x + j; ai[j] = j - 3.f; ap[j] = pk(ar[j], ai[j]); }
    const u64 sp = pk(s, s), cp = pk(0.5f, -0.25f);
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (V == 0) { ar[j] = fmaf(ar[j], s, 0.5f); ai[j] = fmaf(ai[j], s, -0.25f); ar[j] += 1.0f; ai[j] += 2.0f; }
            else {
                asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ap[j]) : "l"(ap[j]), "l"(sp), "l"(cp));
                asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(ap[j]) : "l"(ap[j]), "l"(cp));
            }
        }
        if (MIX) {
            // 32 integer ALU instructions per iteration next to the FP work (do they fit the slots packed ops leave free?)
#pragma unroll
            for (int j = 0; j < 16; j++) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ix[j & 7]) : "r"(i), "r"(j)); asm volatile("add.u32 %0, %0, %1;" : "+r"(ix[(j + 3) & 7]) : "r"(i)); }
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int j = 0; j < 8; j++) acc += V == 0 ? ar[j] + ai[j] : __uint_as_float((unsigned)ap[j]) + __uint_as_float((unsigned)(ap[j] >> 32));
    for (int j = 0; j < 8; j++) acc += (float)ix[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[V + 2 * MIX] = t1 - t0;
}
int main() {
    float* o; long long* c; cudaMalloc(&o, 148 * 4 * 512 * 4); cudaMalloc(&c, 32);
    const int iters = 4096;
    for (int rep = 0; rep < 2; rep++) {
        k<0, 0><<<148, 512>>>(o, iters, 0.999f, c);
        k<1, 0><<<148, 512>>>(o, iters, 0.999f, c);
        k<0, 1><<<148, 512>>>(o, iters, 0.999f, c);
        k<1, 1><<<148, 512>>>(o, iters, 0.999f, c);
    }
    cudaDeviceSynchronize();
    long long h[4]; cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
    // per iteration a warp issues 32 scalar (16 FFMA + 16 FADD) or 16 packed instructions; 4 warps share a sub-partition
    printf("scalar: %.2f cycles / iteration / warp (32 instr)  packed: %.2f cycles / iteration / warp (16 instr)\n", (double)h[0] / iters, (double)h[1] / iters);
    printf("=> per sub-partition (4 warps): scalar %.2f cycles per instr, packed %.2f cycles per instr\n", (double)h[0] / iters / 32 / 4, (double)h[1] / iters / 16 / 4);
    printf("with 32 integer ALU instr per iteration: scalar %.2f cycles / iteration / warp (64 instr), packed %.2f (48 instr)\n", (double)h[2] / iters, (double)h[3] / iters);
    return 0;
}
