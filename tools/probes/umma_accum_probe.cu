// umma_accum_probe.cu — precision of a tcgen05 f16 x f16 -> f32 accumulate chain (K = 112 in 7 steps) for operands shaped like the
// log-mel DFT study (tools/studies/logmel_tc_numerics.py): is the fp32 accumulation as good as a round-to-nearest fmaf chain?
// Kernel and descriptor layout are those of umma_probe.cu.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o _bin/umma_accum_probe umma_accum_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint64_t make_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;     // descriptor version 1 (Blackwell)
    return d;                   // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
__device__ __forceinline__ void umma_f16(unsigned d_tmem, uint64_t da, uint64_t db, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

struct Step { int a_chunk; int b_tile; int d_col; int acc; };
constexpr int kALbo = 2064;          // bytes between K chunks of A (128 rows x 16 B + 16 B pad)
constexpr int kAChunks = 14;         // A holds K = 112 (7 k-steps of 16)
constexpr int kBTile = 512;          // one [N=16][K=16] B tile: [ngroup 2][kchunk 2][8 rows][16 B]

__global__ void __launch_bounds__(160, 1) probe_kernel(const unsigned char* a_img, const unsigned char* b_img, int n_btiles, const Step* steps, int n_steps,
                                                       int swap_fields, unsigned idesc, float* out /*[128][64]*/) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sa = smem;                                  // kAChunks * kALbo
    unsigned char* sb = smem + kAChunks * kALbo;               // n_btiles * 512
    __shared__ __align__(8) unsigned long long bar_mem;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kAChunks * kALbo; i += blockDim.x) sa[i] = a_img[i];
    for (int i = tid; i < n_btiles * kBTile; i += blockDim.x) sb[i] = b_img[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const unsigned bar = smem_u32(&bar_mem);
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = tmem_base_s;
    if (tid == 128) {
        for (int s = 0; s < n_steps; s++) {
            const Step st = steps[s];
            const unsigned a_addr = smem_u32(sa) + st.a_chunk * kALbo;
            const unsigned b_addr = smem_u32(sb) + st.b_tile * kBTile;
            const uint64_t da = swap_fields ? make_desc(a_addr, 128, kALbo) : make_desc(a_addr, kALbo, 128);
            const uint64_t db = swap_fields ? make_desc(b_addr, 256, 128) : make_desc(b_addr, 128, 256);
            umma_f16(tbase + st.d_col, da, db, idesc, (unsigned)st.acc);
        }
        umma_commit(bar);
    }
    if (warp < 4) {
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < 64; c0 += 16) {
            unsigned r[16];
            const unsigned taddr = tbase + ((unsigned)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; j++) out[tid * 64 + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(64u) : "memory");
}

static unsigned make_idesc(int m, int n) { return (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24); }

// accumulate-chain precision: D[128][16] = sum over 7 k-steps of A[:, 16s:16s+16] * B_s, operands shaped like the log-mel
// DFT study (A = integer planes up to +-1024 or +-64, B = f16 cos/sin values), against the exact (float64) sum and against
// a sequential fp32 (round-to-nearest) accumulation on the host.
int main() {
    const int M = 128, KA = kAChunks * 8, NB = 7;
    printf("header of the probe this kernel came from: umma_probe.cu (descriptor layout pinned there)\n");
    for (int variant = 0; variant < 3; variant++) {
        std::vector<float> A(M * KA), B(NB * 16 * 16);
        srand(11 + variant);
        const int amax = variant == 1 ? 64 : 1024;
        for (auto& v : A) v = variant == 2 ? 1000.0f + (float)(rand() % 25) : (float)((rand() % (2 * amax + 1)) - amax);   // variant 2: same-sign large sums
        for (int t = 0; t < NB; t++)
            for (int n = 0; n < 16; n++)
                for (int k = 0; k < 16; k++) {
                    float c = variant == 2 ? 0.75f + 0.2f * cosf(0.37f * (t * 16 + k) * (n + 1)) : cosf(6.2831853f * (float)((t * 16 + k) * (n + 3)) / 200.0f);
                    B[(t * 16 + n) * 16 + k] = __half2float(__float2half(c));
                }
        std::vector<unsigned char> a_img(kAChunks * kALbo, 0), b_img(NB * kBTile, 0);
        for (int r = 0; r < M; r++)
            for (int k = 0; k < KA; k++) {
                __half h = __float2half(A[r * KA + k]);
                size_t off = (size_t)(k / 8) * kALbo + (size_t)(r / 8) * 128 + (r % 8) * 16 + (k % 8) * 2;
                memcpy(&a_img[off], &h, 2);
            }
        for (int t = 0; t < NB; t++)
            for (int n = 0; n < 16; n++)
                for (int k = 0; k < 16; k++) {
                    __half h = __float2half(B[(t * 16 + n) * 16 + k]);
                    size_t off = (size_t)t * kBTile + (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
                    memcpy(&b_img[off], &h, 2);
                }
        std::vector<Step> steps;
        for (int s = 0; s < NB; s++) steps.push_back({2 * s, s, 0, s > 0});
        unsigned char *d_a, *d_b; Step* d_steps; float* d_out;
        CK(cudaMalloc(&d_a, a_img.size())); CK(cudaMalloc(&d_b, b_img.size())); CK(cudaMalloc(&d_steps, steps.size() * sizeof(Step))); CK(cudaMalloc(&d_out, M * 64 * 4));
        CK(cudaMemcpy(d_a, a_img.data(), a_img.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_b, b_img.data(), b_img.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_steps, steps.data(), steps.size() * sizeof(Step), cudaMemcpyHostToDevice));
        const size_t smem = kAChunks * kALbo + NB * kBTile;
        CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaMemset(d_out, 0xff, M * 64 * 4));
        probe_kernel<<<1, 160, smem>>>(d_a, d_b, NB, d_steps, (int)steps.size(), 0, make_idesc(128, 16), d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> out(M * 64);
        CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
        double max_tc = 0, max_f32 = 0, sum_tc = 0, sum_f32 = 0, sgn_tc = 0, maxmag = 0;
        int exact_tc = 0;
        for (int r = 0; r < M; r++)
            for (int n = 0; n < 16; n++) {
                double ex = 0; float f = 0;
                for (int k = 0; k < NB * 16; k++) {
                    const float a = A[r * KA + k], b = B[((k / 16) * 16 + n) * 16 + (k % 16)];
                    ex += (double)a * (double)b;
                    f = fmaf(a, b, f);
                }
                const double dt = (double)out[r * 64 + n] - ex, df = (double)f - ex;
                max_tc = fmax(max_tc, fabs(dt)); max_f32 = fmax(max_f32, fabs(df));
                sum_tc += fabs(dt); sum_f32 += fabs(df); sgn_tc += dt * (ex >= 0 ? 1 : -1);
                maxmag = fmax(maxmag, fabs(ex));
                if ((float)ex == out[r * 64 + n]) exact_tc++;
            }
        const int cnt = M * 16;
        printf("variant %d (|A| <= %d, K = 112 in 7 accumulate steps, max |sum| %.0f): tcgen05 max err %.3g mean %.3g signed-toward-zero mean %.3g, "
               "correctly rounded %d / %d | host fmaf chain max err %.3g mean %.3g | fp32 ulp at max |sum| = %.3g\n",
               variant, variant == 2 ? 1024 : amax, maxmag, max_tc, sum_tc / cnt, -sgn_tc / cnt, exact_tc, cnt, max_f32, sum_f32 / cnt, maxmag * 1.19e-7);
    }
    return 0;
}
