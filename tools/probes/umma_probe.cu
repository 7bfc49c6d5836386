// umma_probe.cu — pins the tcgen05 facts the FIR kernel relies on (run on a B200; prints PASS/FAIL lines).
//   1. shared-memory matrix descriptor, SWIZZLE_NONE, K-major: core matrix = 8 rows x 16 bytes, the two K chunks of one
//      K=16 MMA are LBO apart, 8-row groups are SBO apart (hypothesis V0; V1 = fields swapped)
//   2. instruction descriptor for kind::f16, f16 x f16 -> f32, M=128, N=16
//   3. accumulate predicate, descriptor advance along K (even and odd chunk starts), second accumulator at a column offset
//   4. D layout for M=128 / cta_group::1: TMEM lane = row, column = n; tcgen05.ld 32x32b by the warp of each lane quadrant
//   5. issue rate of back-to-back M128 N16 K16 (and N32) MMAs
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o umma_probe umma_probe.cu
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
constexpr int kAChunks = 8;          // A holds K = 64
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

// issue-rate probe: reps MMAs back to back (unrolled x8), alternating over `naccs` accumulators.
// mode 0: A from smem SWIZZLE_NONE; 1: A from smem SWIZZLE_128B; 2: A from TMEM
__device__ __forceinline__ void umma_f16_ts(unsigned d_tmem, unsigned a_tmem, uint64_t db, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__global__ void __launch_bounds__(160, 1) rate_kernel(int reps, int naccs, int n, int mode, unsigned idesc, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar_mem;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * 1024 / 4; i += blockDim.x) ((unsigned*)smem)[i] = 0x3c003c00u;   // f16 1.0
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const unsigned bar = smem_u32(&bar_mem);
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = tmem_base_s;
    if (tid == 128) {
        const unsigned a_addr = smem_u32(smem), b_addr = smem_u32(smem) + 32768;
        uint64_t da = make_desc(a_addr, 2048, 128);
        if (mode == 1) da = make_desc(a_addr, 16, 1024) | ((uint64_t)2 << 61);
        const uint64_t db = make_desc(b_addr, 128, 256);
        const unsigned d0 = tbase, d1 = tbase + (naccs > 1 ? (unsigned)n : 0u);
        const unsigned a_tm = tbase + 256 + 64;     // TS mode: A tile lives in TMEM columns (8 columns per K=16)
        const long long t0 = clock64();
        if (mode == 2) {
            for (int s = 0; s < reps; s += 8) {
#pragma unroll
                for (int u = 0; u < 8; u++) umma_f16_ts((u & 1) ? d1 : d0, a_tm + 8 * u, db, idesc, 1u);
            }
        } else {
            for (int s = 0; s < reps; s += 8) {
#pragma unroll
                for (int u = 0; u < 8; u++) umma_f16((u & 1) ? d1 : d0, da + (uint64_t)(u * (mode == 1 ? 2 : 256)), db, idesc, 1u);
            }
        }
        const long long t1 = clock64();
        umma_commit(bar);
        mbar_wait(bar, 0);
        const long long t2 = clock64();
        cycles[0] = t1 - t0;
        cycles[1] = t2 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}


// TS-mode correctness: A[128 x 32] f16 is written to TMEM by the four quadrant warps (thread = row, register j = column
// j = K elements 2j (low half), 2j+1 (high half)) starting at column a_col0; D = A[:, k0:k0+16] * B0 for two k0 values
// (k0 = 0 -> A column a_col0, k0 = 8 -> A column a_col0 + 4: a 4-column-aligned operand address)
__global__ void __launch_bounds__(160, 1) ts_kernel(const unsigned* a_words /*[128][16]*/, const unsigned char* b_img, unsigned idesc, int a_col0, float* out /*[128][32]*/) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bar_mem;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kBTile; i += blockDim.x) smem[i] = b_img[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const unsigned bar = smem_u32(&bar_mem);
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tbase = tmem_base_s;
    if (warp < 4) {
        unsigned r[16];
        for (int j = 0; j < 16; j++) r[j] = a_words[tid * 16 + j];
        const unsigned taddr = tbase + ((unsigned)(warp * 32) << 16) + 64 + a_col0;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                       "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 128) {
        const uint64_t db = make_desc(smem_u32(smem), 128, 256);
        for (int v = 0; v < 2; v++) {
            const unsigned a_tm = tbase + 64 + a_col0 + 4 * v, d_tm = tbase + 16 * v;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                         ::"r"(d_tm), "r"(a_tm), "l"(db), "r"(idesc), "r"(0u) : "memory");
        }
        umma_commit(bar);
    }
    if (warp < 4) {
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < 32; c0 += 16) {
            unsigned r[16];
            const unsigned taddr = tbase + ((unsigned)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; j++) out[tid * 32 + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(128u) : "memory");
}

static unsigned make_idesc(int m, int n) {
    // c_format f32 = 1 at [4,6); a/b format f16 = 0; K-major both; n>>3 at [17,23); m>>4 at [24,29)
    return (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}

int main() {
    const int M = 128, KA = kAChunks * 8, NB = 4;   // A[128][64], 4 B tiles [16][16]
    std::vector<float> A(M * KA), B(NB * 16 * 16);
    srand(7);
    for (auto& v : A) v = (float)((rand() % 257) - 128);            // exact in f16
    for (auto& v : B) v = (float)((rand() % 33) - 16) / 8.0f;
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
                __half h = __float2half(B[(t * 16 + n) * 16 + k]);      // B[t][n][k]
                size_t off = (size_t)t * kBTile + (n / 8) * 256 + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2;
                memcpy(&b_img[off], &h, 2);
            }
    // program: D0 (cols 0..15) = A[:,0:16] B0 + A[:,16:32] B1 + A[:,32:48] B2 ; D1 (cols 16..31) = A[:,8:24] B3 ; D2 (cols 32..47) = A[:,40:56] B0 + A[:,24:40] B1
    std::vector<Step> steps = {{0, 0, 0, 0}, {2, 1, 0, 1}, {4, 2, 0, 1}, {1, 3, 16, 0}, {5, 0, 32, 0}, {3, 1, 32, 1}};
    std::vector<float> ref(M * 64, 0.f);
    for (auto& st : steps)
        for (int r = 0; r < M; r++)
            for (int n = 0; n < 16; n++) {
                float acc = st.acc ? ref[r * 64 + st.d_col + n] : 0.f;
                for (int k = 0; k < 16; k++) acc += A[r * KA + st.a_chunk * 8 + k] * B[(st.b_tile * 16 + n) * 16 + k];
                ref[r * 64 + st.d_col + n] = acc;
            }
    unsigned char *d_a, *d_b; Step* d_steps; float* d_out;
    CK(cudaMalloc(&d_a, a_img.size())); CK(cudaMalloc(&d_b, b_img.size())); CK(cudaMalloc(&d_steps, steps.size() * sizeof(Step))); CK(cudaMalloc(&d_out, M * 64 * 4));
    CK(cudaMemcpy(d_a, a_img.data(), a_img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_b, b_img.data(), b_img.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_steps, steps.data(), steps.size() * sizeof(Step), cudaMemcpyHostToDevice));
    const size_t smem = kAChunks * kALbo + NB * kBTile;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int variant = 0; variant < 1; variant++) {   // variant 1 (fields swapped) faults on hardware: V0 is the layout
        CK(cudaMemset(d_out, 0xff, M * 64 * 4));
        probe_kernel<<<1, 160, smem>>>(d_a, d_b, NB, d_steps, (int)steps.size(), variant, make_idesc(128, 16), d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("variant %d: kernel failed: %s\n", variant, cudaGetErrorString(e)); return 1; }
        std::vector<float> out(M * 64);
        CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0; int bad = 0;
        for (int r = 0; r < M; r++)
            for (int c = 0; c < 48; c++) {
                double d = fabs((double)out[r * 64 + c] - ref[r * 64 + c]);
                if (!(d <= 1e-3)) bad++;
                if (d > maxerr || d != d) maxerr = d;
            }
        printf("variant %d (%s): max err %.4g, bad %d / %d -> %s\n", variant, variant ? "LBO/SBO swapped" : "LBO = K-chunk stride, SBO = 8-row-group stride", maxerr, bad, M * 48, bad == 0 ? "PASS" : "FAIL");
        if (bad && variant == 0) {
            printf("  sample row 0: got"); for (int c = 0; c < 8; c++) printf(" %g", out[c]); printf(" | want"); for (int c = 0; c < 8; c++) printf(" %g", ref[c]); printf("\n");
            printf("  sample row 9: got"); for (int c = 0; c < 8; c++) printf(" %g", out[9 * 64 + c]); printf(" | want"); for (int c = 0; c < 8; c++) printf(" %g", ref[9 * 64 + c]); printf("\n");
        }
    }
    {   // TS mode: A from TMEM
        std::vector<unsigned> aw(128 * 16);
        for (int r = 0; r < 128; r++)
            for (int j = 0; j < 16; j++) {
                __half lo = __float2half(A[r * KA + 2 * j]), hi = __float2half(A[r * KA + 2 * j + 1]);
                unsigned short l, h; memcpy(&l, &lo, 2); memcpy(&h, &hi, 2);
                aw[r * 16 + j] = (unsigned)l | ((unsigned)h << 16);
            }
        unsigned* d_aw; float* d_o2; CK(cudaMalloc(&d_aw, aw.size() * 4)); CK(cudaMalloc(&d_o2, 128 * 32 * 4));
        CK(cudaMemcpy(d_aw, aw.data(), aw.size() * 4, cudaMemcpyHostToDevice));
        for (int a_col0 : {0, 4, 20}) {
            CK(cudaMemset(d_o2, 0xff, 128 * 32 * 4));
            ts_kernel<<<1, 160, kBTile>>>(d_aw, d_b, make_idesc(128, 16), a_col0, d_o2);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("TS a_col0=%d: kernel failed: %s\n", a_col0, cudaGetErrorString(e)); return 1; }
            std::vector<float> o2(128 * 32);
            CK(cudaMemcpy(o2.data(), d_o2, o2.size() * 4, cudaMemcpyDeviceToHost));
            for (int v = 0; v < 2; v++) {
                int bad = 0; double me = 0;
                for (int r = 0; r < 128; r++)
                    for (int n = 0; n < 16; n++) {
                        float acc = 0;
                        for (int k = 0; k < 16; k++) acc += A[r * KA + 8 * v + k] * B[(0 * 16 + n) * 16 + k];
                        double d = fabs((double)o2[r * 32 + 16 * v + n] - acc);
                        if (!(d <= 1e-3)) bad++;
                        if (d > me || d != d) me = d;
                    }
                printf("TS mode (A in TMEM, thread = row, 2 f16 per column) a_col0=%d operand column +%d: max err %.4g bad %d -> %s\n", a_col0, 4 * v, me, bad, bad ? "FAIL" : "PASS");
            }
        }
    }
    long long* d_cyc; CK(cudaMalloc(&d_cyc, 16));
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    for (int mode = 0; mode < 3; mode++)
        for (int n : {16, 32, 64, 128, 256})
            for (int naccs : {1, 2}) {
                if (n * naccs > 256) continue;
                const int reps = 4096;
                const int m = 128;
                if (n % 16 != 0) continue;
                rate_kernel<<<1, 160, 65536>>>(reps, naccs, n, mode, make_idesc(m, n), d_cyc);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("rate mode %d N=%d: kernel failed: %s\n", mode, n, cudaGetErrorString(e)); return 1; }
                long long cyc[2]; CK(cudaMemcpy(cyc, d_cyc, 16, cudaMemcpyDeviceToHost));
                printf("rate mode %d (%s) M128 N%-3d K16, %d accumulator(s): issue %.1f cycles/MMA, complete %.1f cycles/MMA\n", mode,
                       mode == 0 ? "A smem no-swizzle" : mode == 1 ? "A smem SW128" : "A in TMEM", n, naccs, (double)cyc[0] / reps, (double)cyc[1] / reps);
            }
    return 0;
}
