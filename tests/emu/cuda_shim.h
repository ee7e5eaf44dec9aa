// cuda_shim.h -- TEST INFRASTRUCTURE. Lets g++ compile mcts.jl_b200/csrc/*.cu as a single-threaded host
// program (libmctsb200_emu.so) so that the CPU-only test suite can exercise the library's HOST logic
// (option validation, table sizing, widening thresholds, export, sharding) and the G = 1 kernel logic
// without a GPU. The shipped package never loads this build: mcts.jl_b200/_lib.py only ever opens
// csrc/libmctsb200.so and raises if it (or a CUDA device) is missing.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define CUDART_INF (std::numeric_limits<double>::infinity())

struct EmuDim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local EmuDim3 threadIdx, blockIdx, blockDim;

using std::max;
using std::min;

static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xffffffffu) {}
template <class T> static inline T __shfl_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int, int = 32) { return v; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct double2 { double x, y; };
struct ulonglong2 { unsigned long long x, y; };
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d; std::memcpy(&d, &u, 8); return d;
}
static inline int __double2loint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(uint32_t)u; }
static inline int __double2hiint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(uint32_t)(u >> 32); }
static inline int __any_sync(unsigned, int p) { return p; }
static inline long long __double_as_longlong(double x) { long long u; std::memcpy(&u, &x, 8); return u; }
static inline double __longlong_as_double(long long u) { double x; std::memcpy(&x, &u, 8); return x; }
template <class T, class U> static inline T atomicAdd(T* p, U v) { T old = *p; *p += (T)v; return old; }
template <class T, class U> static inline T atomicMax(T* p, U v) { T old = *p; if ((T)v > old) *p = (T)v; return old; }
template <class T, class U> static inline T atomicOr(T* p, U v) { T old = *p; *p |= (T)v; return old; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {  // PRMT, default mode
    const uint64_t src = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t n = (s >> (4 * i)) & 0xfu;
        uint32_t b = (uint32_t)(src >> (8 * (n & 7u))) & 0xffu;
        if (n & 8u) b = (b & 0x80u) ? 0xffu : 0x00u;
        r |= b << (8 * i);
    }
    return r;
}
struct uint2 { unsigned x, y; };

namespace mctsb200_emu {
template <class F>
static inline void run(unsigned grid, unsigned block, F&& f) {
    blockDim.x = block;
    for (unsigned b = 0; b < grid; ++b)
        for (unsigned t = 0; t < block; ++t) {
            blockIdx.x = b;
            threadIdx.x = t;
            f();
        }
}
}  // namespace mctsb200_emu

// ---- a pretend CUDA runtime: device memory is host memory, everything is synchronous ----
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef std::chrono::steady_clock::time_point* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { int major = 10, minor = 0, multiProcessorCount = 148; };
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }
// (cudaMalloc aligns to 256 bytes; the kernels' 32-byte aligned row types rely on it)
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n ? n : 1) + 255 & ~(size_t)255); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new std::chrono::steady_clock::time_point(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { *e = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(*b - *a).count(); return cudaSuccess; }
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) { *n = 4; return cudaSuccess; }
