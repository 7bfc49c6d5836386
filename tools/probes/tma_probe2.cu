// bisect probe for the TMA path: A = mbarrier only, B = 1D bulk copy, C = 2D tensor via libcu++, D = 2D tensor raw PTX
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
namespace cde = cuda::device::experimental;
using barrier = cuda::barrier<cuda::thread_scope_block>;
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void kA(uint32_t* out) {
    __shared__ __align__(8) unsigned long long bar;
    unsigned sb = (unsigned)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sb) : "memory");
    unsigned ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(sb), "r"(0) : "memory");
    out[threadIdx.x] = 7;
}
__global__ void kB(const uint32_t* in, uint32_t* out) {
    __shared__ __align__(128) uint32_t buf[2048];
    __shared__ __align__(8) unsigned long long bar;
    unsigned sb = (unsigned)__cvta_generic_to_shared(&bar), sd = (unsigned)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(8192) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sd), "l"(in), "r"(8192), "r"(sb) : "memory");
    }
    unsigned ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(sb), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = buf[i];
}
__global__ void kC(const __grid_constant__ CUtensorMap tm, int x0, int y0, uint32_t* out) {
    __shared__ alignas(1024) uint32_t buf[64][32];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&buf, &tm, x0, y0, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(buf));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = (&buf[0][0])[i];
}
__global__ void kD(const __grid_constant__ CUtensorMap tm, int x0, int y0, uint32_t* out) {
    __shared__ alignas(1024) uint32_t buf[64][32];
    __shared__ __align__(8) unsigned long long bar;
    unsigned sb = (unsigned)__cvta_generic_to_shared(&bar), sd = (unsigned)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(8192) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(sd), "l"(&tm), "r"(x0), "r"(y0), "r"(sb) : "memory");
    }
    unsigned ok = 0;
    while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(sb), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = (&buf[0][0])[i];
}
int main(int argc, char** argv) {
    const char mode = argc > 1 ? argv[1][0] : 'A';
    const int EXTRA = argc > 2 ? atoi(argv[2]) : 0, SW = argc > 3 ? atoi(argv[3]) : 0;
    const int P4 = 1764, ROWS = 300;
    size_t n = (size_t)P4 * ROWS + 4096;
    std::vector<uint32_t> h(n);
    for (size_t i = 0; i < n; i++) h[i] = (uint32_t)i;
    uint32_t *d, *o;
    cudaMalloc(&d, n * 4); cudaMalloc(&o, 2048 * 4); cudaMemset(o, 0, 8192);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    CUtensorMap tm;
    if (mode == 'C' || mode == 'D') {
        EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
        cuuint64_t dims[2] = {(cuuint64_t)(P4 + EXTRA), (cuuint64_t)ROWS};
        cuuint64_t strides[1] = {(cuuint64_t)P4 * 4};
        cuuint32_t box[2] = {32, 64}, es[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         SW ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("mode %c EXTRA=%d SW=%d encode=%d\n", mode, EXTRA, SW, (int)r);
    }
    int x0 = argc > 4 ? atoi(argv[4]) : 1755, y0 = argc > 5 ? atoi(argv[5]) : 5;
    if (mode == 'A') kA<<<1, 128>>>(o);
    if (mode == 'B') kB<<<1, 128>>>(d, o);
    if (mode == 'C') kC<<<1, 128>>>(tm, x0, y0, o);
    if (mode == 'D') kD<<<1, 128>>>(tm, x0, y0, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("mode %c: sync=%d (%s)\n", mode, (int)e, cudaGetErrorString(e));
    if (e) return 2;
    std::vector<uint32_t> res(2048);
    cudaMemcpy(res.data(), o, 8192, cudaMemcpyDeviceToHost);
    if (mode == 'C' || mode == 'D') {
        int bad = 0, badsw = 0;
        for (int row = 0; row < 64; row++)
            for (int c = 0; c < 32; c++) {
                long long x = x0 + c, y = y0 + row;
                uint32_t want = (x < 0 || x >= P4 + EXTRA || y >= ROWS) ? 0u : (uint32_t)(y * P4 + x);
                if (res[row * 32 + (((c / 4) ^ (row & 7)) * 4 + c % 4)] != want) badsw++;
                if (res[row * 32 + c] != want) bad++;
            }
        printf("  x0=%d y0=%d mismatches: linear=%d swizzled=%d ; row1: %u %u %u %u | %u (want %u)\n", x0, y0, bad, badsw, res[32], res[33], res[34], res[35], res[36], (uint32_t)((y0 + 1) * P4 + x0));
    } else printf("  out[0..3]=%u %u %u %u\n", res[0], res[1], res[2], res[3]);
    return 0;
}
