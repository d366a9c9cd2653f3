// Host emulation of the CUDA execution model — DEVELOPMENT / TEST TOOL ONLY.
//
// There is no GPU in the build container, so the SIMT kernels of this package can also be compiled as plain
// C++ (g++ -DVANERF_HOST_EMUL -ffp-contract=off) into tests/_emul/libvanerf_emul.so, where a thread block is a
// group of fibers, __syncthreads() is a barrier and warp shuffles go through a per-warp exchange buffer.
// Only tests/ load that library, to check kernel logic against the oracle before GPU time is spent.  The product
// (vanerf_b200/_lib.py) loads the nvcc-built libvanerf_b200.so only and raises if it is missing: this header is
// never part of that build.  tcgen05 / TMA kernels are not emulated.
#pragma once
#ifdef VANERF_HOST_EMUL

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_ { unsigned x, y, z; };
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

using std::max;
using std::min;
typedef void* cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0

namespace emul {
struct Barrier {                  // members of a block are fibers on one OS thread: no locking needed
    int count = 0, gen = 0, n = 0;
    void init(int n_) { n = n_; count = 0; gen = 0; }
    void wait();
};
struct Block {
    Barrier bar;
    std::vector<Barrier> warp_bar;
    std::vector<uint64_t> xchg;   // 32 slots per warp
    char* dyn_smem = nullptr;
};
extern thread_local uint3_ t_threadIdx, t_blockIdx;
extern thread_local Block* t_block;
extern dim3 g_blockDim, g_gridDim;
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
}  // namespace emul

#define threadIdx (emul::t_threadIdx)
#define blockIdx (emul::t_blockIdx)
#define blockDim (emul::g_blockDim)
#define gridDim (emul::g_gridDim)
#define warpSize 32

// __shared__ variables: one block at a time per OS thread, so thread-local statics are per-block storage.
#define __shared__ static thread_local
#define EMUL_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emul::t_block->dyn_smem)

static inline void __syncthreads() { emul::t_block->bar.wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::t_block->warp_bar[threadIdx.x / 32].wait(); }

template <typename T>
static inline T emul_shfl(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    emul::Block* b = emul::t_block;
    int w = threadIdx.x / 32, l = threadIdx.x % 32;
    uint64_t bits = 0;
    std::memcpy(&bits, &v, sizeof(T));
    b->xchg[w * 32 + l] = bits;
    b->warp_bar[w].wait();
    uint64_t o = b->xchg[w * 32 + (src_lane & 31)];
    b->warp_bar[w].wait();
    T r;
    std::memcpy(&r, &o, sizeof(T));
    return r;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int lane) { return emul_shfl(v, lane); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emul_shfl(v, (int)(threadIdx.x % 32) ^ m); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    int l = threadIdx.x % 32;
    T o = emul_shfl(v, l - (int)d < 0 ? l : l - (int)d);
    return o;
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    int l = threadIdx.x % 32;
    return emul_shfl(v, l + (int)d > 31 ? l : l + (int)d);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (emul_shfl(pred ? 1 : 0, i) ? 1u : 0u) << i;
    return r;
}

// ---- exact / fast math intrinsics (build with -ffp-contract=off)
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; std::memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float fminf_(float a, float b) { return fminf(a, b); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }

static inline unsigned long long atomicMin(unsigned long long* a, unsigned long long v) {
    unsigned long long old = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(a, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_RELAXED); }
static inline float atomicAdd(float* a, float v) {
    unsigned* p = reinterpret_cast<unsigned*>(a);
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED), nw;
    float f;
    do { std::memcpy(&f, &old, 4); f += v; std::memcpy(&nw, &f, 4); } while (!__atomic_compare_exchange_n(p, &old, nw, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    return f - v;
}
static inline int atomicMin(int* a, int v) {
    int old = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (v < old && !__atomic_compare_exchange_n(a, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int atomicMax(int* a, int v) {
    int old = __atomic_load_n(a, __ATOMIC_RELAXED);
    while (v > old && !__atomic_compare_exchange_n(a, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline unsigned atomicOr(unsigned* a, unsigned v) { return __atomic_fetch_or(a, v, __ATOMIC_RELAXED); }
static inline int atomicExch(int* a, int v) { return __atomic_exchange_n(a, v, __ATOMIC_RELAXED); }

// ---- just enough of the CUDA runtime API for api.cu (host memory stands in for device memory)
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
typedef int cudaEvent_t;
#define cudaEventDisableTiming 2
#define cudaHostAllocDefault 0
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = 1; return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
static inline cudaError_t cudaFreeHost(void* p) { std::free(p); return 0; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return 0; }
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }

#define VANERF_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emul::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); })

#endif  // VANERF_HOST_EMUL
