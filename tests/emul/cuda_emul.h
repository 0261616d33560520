// cuda_emul.h -- TEST-ONLY single-threaded SIMT emulator for the kernel sources in tap-stark_b200/csrc.
//
// This is NOT a CPU fallback of the product.  The shipped library (libtapstark_b200.so) is compiled by
// nvcc for sm_100a only and has no host compute path.  This header exists so that the *same kernel source
// files* can be compiled by g++ into tests/emul/_build/libtapstark_emul.so and checked against the oracle
// in the GPU-less build container (index math, digit decomposition, twiddle exponents, Blake3 state
// machine).  The product package refuses to load the emulated library.
//
// Model: one CUDA block at a time; every CUDA thread is a ucontext fiber; __syncthreads() yields to a
// round-robin scheduler, so barrier semantics are exact and execution is deterministic.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

namespace ts_emul {
extern dim3 g_threadIdx, g_blockIdx, g_blockDim, g_gridDim;
extern unsigned char *g_smem;
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body);
void syncthreads();
extern unsigned long long g_launches;
}  // namespace ts_emul

#define threadIdx (ts_emul::g_threadIdx)
#define blockIdx (ts_emul::g_blockIdx)
#define blockDim (ts_emul::g_blockDim)
#define gridDim (ts_emul::g_gridDim)

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
static inline void __syncthreads() { ts_emul::syncthreads(); }
// warp-level barrier: with one-fiber-per-thread round-robin scheduling a yield gives exactly the needed
// guarantee (when a thread resumes, every other live thread has passed its previous sync point)
static inline void __syncwarp() { ts_emul::syncthreads(); }
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint32_t __brev(uint32_t x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    x = ((x >> 8) & 0x00FF00FFu) | ((x & 0x00FF00FFu) << 8);
    return (x >> 16) | (x << 16);
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
template <class T> static inline T __ldg(const T *p) { return *p; }
using std::max;
using std::min;

// ---- minimal runtime API ---------------------------------------------------------------------
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaMalloc(void **p, size_t n) {
    *p = nullptr;
    if (posix_memalign(p, 256, n ? n : 256)) return 1;
    memset(*p, 0xCD, n);  // poison: uninitialised reads show up
    return 0;
}
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) {
    memmove(d, s, n);
    return 0;
}
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k) {
    return cudaMemcpyAsync(d, s, n, k, 0);
}
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t wbytes, size_t h,
                                            cudaMemcpyKind, cudaStream_t = 0) {
    for (size_t i = 0; i < h; i++) memmove((char *)d + i * dp, (const char *)s + i * sp, wbytes);
    return 0;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return 0; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
struct cudaDeviceProp { int multiProcessorCount; char name[64]; };
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    p->multiProcessorCount = 4;
    strcpy(p->name, "SIMT-emulator (test only)");
    return 0;
}
