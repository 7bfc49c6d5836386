// cuda_emu.h — TEST-ONLY CPU emulation of the CUDA execution model.
//
// This container has no GPU, and GPU minutes are rationed.  To check kernel
// index math, barriers and integer bit-exactness before going to a B200, the
// kernels under audio_processor_b200/csrc are ALSO compiled with g++ against
// this shim (see tests/emu/build_emu.py) into tests/emu/libb2a_emu.so, which
// only the `-m "not gpu"` tests load.  It is never built into, linked with or
// loaded by the product library (libb2a.so) — the product path has no CPU
// fallback and fails loudly without CUDA.
//
// Model: one OS thread; each CUDA thread of a block is a ucontext fiber;
// blocks run one after another.  __syncthreads/__syncwarp/shuffles/ballots
// are cooperative barriers between fibers.  Atomics are plain (serial).
#pragma once
#ifndef B2A_EMU
#error "cuda_emu.h is only for the emulation build"
#endif

#include <ucontext.h>
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../audio_processor_b200/csrc/f16_bits.h"

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))
#define __grid_constant__

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct short2 { short x, y; };
struct ushort2 { unsigned short x, y; };
struct __attribute__((aligned(16))) longlong2 { long long x, y; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
static inline short2 make_short2(short x, short y) { return {x, y}; }

// ---- runtime API subset ----------------------------------------------------------
typedef int cudaError_t;
typedef struct CUstream_st* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2,
                      cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? 0 : 1; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
typedef struct CUevent_st* cudaEvent_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }

namespace emu {

struct Fiber {
    ucontext_t ctx;
    std::vector<char> stack;
    bool done = false;
};

struct Block {
    std::vector<Fiber> fib;
    ucontext_t sched;
    int cur = -1;
    int nthreads = 0;
    int alive = 0;
    // block barrier
    int bar_count = 0;
    unsigned bar_gen = 0;
    // warp barriers / exchange
    std::vector<int> wbar_count;
    std::vector<unsigned> wbar_gen;
    std::vector<int> walive;
    std::vector<uint64_t> xchg;  // [nwarps*32]
    std::vector<uint32_t> wscr;  // [nwarps*32*8] per-lane scratch for warp-collective emulations (mma)
    void (*entry)(void*) = nullptr;
    void* entry_arg = nullptr;
};

extern Block* g_blk;
extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;

inline void set_tid(int t) {
    g_threadIdx.x = t % g_blockDim.x;
    g_threadIdx.y = (t / g_blockDim.x) % g_blockDim.y;
    g_threadIdx.z = t / (g_blockDim.x * g_blockDim.y);
}
inline void yield() {
    Block* b = g_blk;
    int me = b->cur;
    swapcontext(&b->fib[me].ctx, &b->sched);
    set_tid(me);
}
inline void syncthreads() {
    Block* b = g_blk;
    unsigned gen = b->bar_gen;
    if (++b->bar_count >= b->alive) { b->bar_count = 0; b->bar_gen++; return; }
    while (b->bar_gen == gen) yield();
}
inline void syncwarp() {
    Block* b = g_blk;
    int w = b->cur / 32;
    unsigned gen = b->wbar_gen[w];
    if (++b->wbar_count[w] >= b->walive[w]) { b->wbar_count[w] = 0; b->wbar_gen[w]++; return; }
    while (b->wbar_gen[w] == gen) yield();
}
template <class T> inline uint64_t to_bits(T v) { uint64_t u = 0; static_assert(sizeof(T) <= 8, ""); memcpy(&u, &v, sizeof(T)); return u; }
template <class T> inline T from_bits(uint64_t u) { T v; memcpy(&v, &u, sizeof(T)); return v; }
template <class T> inline T exchange(T v, int src_lane) {
    Block* b = g_blk;
    int w = b->cur / 32, lane = b->cur % 32;
    b->xchg[w * 32 + lane] = to_bits(v);
    syncwarp();
    int s = src_lane;
    T r = (s >= 0 && s < 32 && w * 32 + s < b->nthreads) ? from_bits<T>(b->xchg[w * 32 + s]) : v;
    syncwarp();
    return r;
}

void run_block(void (*entry)(void*), void* arg, int nthreads);

static inline float f16_to_f32(unsigned short h) { return b2a_f16::f16_to_f32(h); }
static inline unsigned short f32_to_f16(float f) { return b2a_f16::f32_to_f16(f); }

// mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32, warp collective: every lane publishes its fragments, then
// computes its four D elements in f32 (k ascending).  Fragment layouts per the PTX ISA.
inline void mma_m16n8k16_f16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    Block* b = g_blk;
    const int w = b->cur / 32, lane = b->cur % 32;
    uint32_t* scr = b->wscr.data() + (size_t)w * 32 * 8;
    for (int i = 0; i < 4; i++) scr[lane * 8 + i] = a[i];
    scr[lane * 8 + 4] = b0;
    scr[lane * 8 + 5] = b1;
    syncwarp();
    const int g = lane >> 2, t = lane & 3;
    auto A = [&](int row, int k) -> float {
        const int l = (row & 7) * 4 + ((k & 7) >> 1);
        const int reg = (row >> 3) + 2 * (k >> 3);
        const uint32_t v = scr[l * 8 + reg];
        return f16_to_f32((unsigned short)((k & 1) ? (v >> 16) : (v & 0xffff)));
    };
    auto B = [&](int k, int col) -> float {
        const int l = col * 4 + ((k & 7) >> 1);
        const uint32_t v = scr[l * 8 + 4 + (k >> 3)];
        return f16_to_f32((unsigned short)((k & 1) ? (v >> 16) : (v & 0xffff)));
    };
    float r[4];
    for (int e = 0; e < 4; e++) {
        const int row = g + 8 * (e >> 1), col = 2 * t + (e & 1);
        float acc = d[e];
        for (int k = 0; k < 16; k++) acc += A(row, k) * B(k, col);
        r[e] = acc;
    }
    syncwarp();
    for (int e = 0; e < 4; e++) d[e] = r[e];
}

template <class... KArgs, class... Args>
inline void launch(dim3 grid, dim3 block, size_t smem, void (*k)(KArgs...), Args... args) {
    g_gridDim = grid; g_blockDim = block;
    std::vector<unsigned char> dyn(smem + 1024);
    g_dyn_smem = (unsigned char*)(((uintptr_t)dyn.data() + 1023) & ~(uintptr_t)1023);   // SWIZZLE_128B boxes need 1 KiB alignment
    auto thunk = [&]() { k(static_cast<KArgs>(args)...); };
    using Th = decltype(thunk);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                g_blockIdx = {bx, by, bz};
                run_block([](void* p) { (*(Th*)p)(); }, &thunk, (int)(block.x * block.y * block.z));
            }
    g_dyn_smem = nullptr;
}
}  // namespace emu

#define threadIdx (emu::g_threadIdx)
#define blockIdx (emu::g_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::syncwarp(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline int emu_lane() { return emu::g_blk->cur % 32; }
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int lane = emu_lane();
    int base = lane / width * width;
    return emu::exchange(v, base + (src % width));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    (void)width;
    return emu::exchange(v, emu_lane() ^ m);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = emu_lane();
    int src = lane - (int)d;
    if (src < lane / width * width) src = lane;
    return emu::exchange(v, src);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = emu_lane();
    int src = lane + (int)d;
    if (src >= lane / width * width + width) src = lane;
    return emu::exchange(v, src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Block* b = emu::g_blk;
    int w = b->cur / 32, lane = b->cur % 32;
    b->xchg[w * 32 + lane] = pred ? 1 : 0;
    emu::syncwarp();
    unsigned r = 0;
    for (int i = 0; i < 32 && w * 32 + i < b->nthreads; i++)
        if (!b->fib[w * 32 + i].done && b->xchg[w * 32 + i]) r |= 1u << i;
    emu::syncwarp();
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) {
    emu::Block* b = emu::g_blk;
    int w = b->cur / 32;
    unsigned act = 0;
    for (int i = 0; i < 32 && w * 32 + i < b->nthreads; i++) if (!b->fib[w * 32 + i].done) act |= 1u << i;
    return __ballot_sync(m, p) == act;
}
static inline unsigned __activemask() { return 0xffffffffu; }

// ---- atomics (serial execution => plain ops) ---------------------------------------------
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { auto o = *p; *p = o + v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }

// ---- intrinsics --------------------------------------------------------------------------
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float2int_rn(float f) {
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)lrintf(f);
}
static inline float __int2float_rn(int i) { return (float)i; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float emu_log2f(float x) { return log2f(x); }
static inline float emu_log10f(float x) { return log10f(x); }
#define __log2f emu_log2f
#define __log10f emu_log10f
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; i++) if (x & (1u << i)) r |= 1u << (31 - i); return r; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const unsigned long long v = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned sel = (s >> (4 * i)) & 0xFu;
        unsigned b = (unsigned)(v >> (8 * (sel & 7u))) & 0xFFu;
        if (sel & 8u) b = (b & 0x80u) ? 0xFFu : 0u;
        r |= b << (8 * i);
    }
    return r;
}
static inline int __dp2a_lo(int a, int b, int c) {
    short a0 = (short)(a & 0xffff), a1 = (short)((unsigned)a >> 16);
    signed char b0 = (signed char)(b & 0xff), b1 = (signed char)((b >> 8) & 0xff);
    return c + a0 * b0 + a1 * b1;
}
static inline long long __mul64hi(long long a, long long b) { return (long long)(((__int128)a * b) >> 64); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
using std::max;
using std::min;
static inline long long max(long long a, int b) { return a > b ? a : (long long)b; }
static inline long long min(long long a, int b) { return a < b ? a : (long long)b; }
static inline long long max(int a, long long b) { return a > b ? (long long)a : b; }
static inline long long min(int a, long long b) { return a < b ? (long long)a : b; }
