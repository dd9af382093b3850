// host_shim.h -- lets g++ parse the device headers of splashsurf_b200/csrc so that the per-thread (non-collective) kernels
// can be stepped on the CPU by tests/test_post_emulation.py.  TEST INFRASTRUCTURE ONLY: warp/block collectives are stubs
// that abort when executed, so only kernels without them may be run.
#pragma once
#include <cuda_runtime.h>      // host side: float4, int2, make_float4, ...
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#undef __shared__
#undef __restrict__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __restrict__

struct EmulIdx { unsigned x = 0, y = 0, z = 0; };
static thread_local EmulIdx blockIdx, threadIdx, blockDim, gridDim;

// IEEE single operations (the TU is built with -ffp-contract=off -fno-fast-math)
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
using std::max;
using std::min;
static inline int min(int a, unsigned b) { return a < (int)b ? a : (int)b; }

[[noreturn]] static inline void emul_collective(const char *what) { fprintf(stderr, "host emulation: collective %s executed\n", what); abort(); }
static inline void __syncthreads() { emul_collective("__syncthreads"); }
static inline int __syncthreads_or(int) { emul_collective("__syncthreads_or"); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul_collective("__syncwarp"); }
static inline unsigned __ballot_sync(unsigned, int) { emul_collective("__ballot_sync"); }
static inline int __all_sync(unsigned, int) { emul_collective("__all_sync"); }
static inline int __any_sync(unsigned, int) { emul_collective("__any_sync"); }
template <typename T> static inline T __shfl_sync(unsigned, T, int) { emul_collective("__shfl_sync"); }
template <typename T> static inline T __shfl_up_sync(unsigned, T, int) { emul_collective("__shfl_up_sync"); }
template <typename T> static inline T __shfl_down_sync(unsigned, T, int) { emul_collective("__shfl_down_sync"); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T, int) { emul_collective("__shfl_xor_sync"); }
// single-threaded emulation: atomics are plain read-modify-writes
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
