// probe: overlapping-row 2D tensor map (row stride 7056 B < dim0*4) + 128B swizzle + arbitrary x0
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap tm, int x0, int y0, uint32_t* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar;
    unsigned sb = (unsigned)__cvta_generic_to_shared(&bar);
    unsigned sd = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(32 * 64 * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(sd), "l"(&tm), "r"(x0), "r"(y0), "r"(sb) : "memory");
    }
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(sb), "r"(0) : "memory");
    }
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) out[i] = ((uint32_t*)smem)[i];
}
int main(int argc, char** argv) {
    const int P4 = 1764, ROWS = 300; const int EXTRA = argc > 1 ? atoi(argv[1]) : 96; const int SW = argc > 2 ? atoi(argv[2]) : 1;
    size_t n = (size_t)P4 * ROWS + 4096;
    std::vector<uint32_t> h(n);
    for (size_t i = 0; i < n; i++) h[i] = (uint32_t)i;
    uint32_t *d, *o;
    cudaMalloc(&d, n * 4); cudaMalloc(&o, 32 * 64 * 4);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    printf("entry point: %d %d %p\n", (int)e, (int)q, (void*)enc);
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)(P4 + EXTRA), (cuuint64_t)ROWS};
    cuuint64_t strides[1] = {(cuuint64_t)P4 * 4};
    cuuint32_t box[2] = {32, 64};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     SW ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("EXTRA=%d SW=%d encode: %d\n", EXTRA, SW, (int)r);
    if (r != CUDA_SUCCESS) return 1;
    int x0s[3] = {1755, -48, 7}, y0s[3] = {5, 0, 280};
    for (int t = 0; t < 3; t++) {
        k<<<1, 128, 32 * 64 * 4 + 1024>>>(tm, x0s[t], y0s[t], o);
        e = cudaDeviceSynchronize(); if (e) { printf("test %d: error %d %s\n", t, (int)e, cudaGetErrorString(e)); return 2; }
        std::vector<uint32_t> res(32 * 64);
        cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, badsw = 0;
        for (int row = 0; row < 64; row++)
            for (int c = 0; c < 32; c++) {
                long long x = x0s[t] + c, y = y0s[t] + row;
                uint32_t want = (x < 0 || x >= P4 + EXTRA || y >= ROWS) ? 0u : (uint32_t)(y * P4 + x);
                int chunk = c / 4, w = c % 4;
                uint32_t got_sw = res[row * 32 + ((chunk ^ (row & 7)) * 4 + w)];
                uint32_t got_lin = res[row * 32 + c];
                if (got_sw != want) badsw++;
                if (got_lin != want) bad++;
            }
        printf("test %d (x0=%d,y0=%d): sync=%d mismatches linear=%d swizzled(chunk^(row&7))=%d  sample row1: %u %u %u %u | %u\n", t, x0s[t], y0s[t],
               (int)e, bad, badsw, res[32], res[33], res[34], res[35], res[36]);
    }
    return 0;
}
