// ts_platform.h -- build-mode shim.  Product build: nvcc, sm_100a, real CUDA runtime.
// TS_EMULATE (tests/emul only): the same sources compiled by g++ against a fiber-based SIMT emulator so
// kernel logic can be checked where no GPU exists.  The emulated build is never shipped or loaded by the
// product package.
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef TS_EMULATE
#include "cuda_emul.h"
#define TS_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(ts_emul::g_smem)
#define TS_LAUNCH(kfn, grid, block, smem, stream, ...) \
    ts_emul::launch(dim3(grid), dim3(block), (smem), [=]() { kfn(__VA_ARGS__); })
#define TS_UNROLL
#define TS_UNROLL2
#define TS_NOUNROLL
#define TS_UNROLL16
#else
#include <cuda_runtime.h>
#define TS_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char name##_raw_[];  \
    type *name = reinterpret_cast<type *>(name##_raw_)
#define TS_LAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define TS_UNROLL _Pragma("unroll")
#define TS_NOUNROLL _Pragma("unroll 1")
#define TS_UNROLL16 _Pragma("unroll 16")
#ifndef TS_NO_LANE_UNROLL2
#define TS_UNROLL2 _Pragma("unroll 2")
#else
#define TS_UNROLL2 _Pragma("unroll 1")
#endif
#endif

#define TS_HD __host__ __device__ __forceinline__
#define TS_D __device__ __forceinline__
