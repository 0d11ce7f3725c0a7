// cpb_rt.h -- thin runtime layer under the kernels.
//
// Normal build (nvcc, sm_100a): the real CUDA runtime; LAUNCH() is <<<>>>.
//
// CPB_EMU build (g++ -x c++ -DCPB_EMU, used ONLY by tools/emu, a developer tool for the
// GPU-less build container): kernels are executed as plain loops over (block, thread) so
// the per-thread logic can be debugged against the oracle before spending GPU time.  The
// emulated library is never built by build(), never loaded by the package, and is not a
// fallback: libcpb200.so contains only the CUDA path.
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#ifndef CPB_EMU
// ------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#include <cooperative_groups.h>

extern unsigned long long g_cpb_launches;   // kernels launched by this library (bench.py reports it)
#define LAUNCH(kernel, grid, block, stream, ...) do { g_cpb_launches++; kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__); } while(0)
#define LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) do { g_cpb_launches++; kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); } while(0)
#define CPB_DEVICE __device__ __forceinline__
#define CPB_HD __host__ __device__ __forceinline__

#else
// ------------------------------------------------------------------ emulation
#include <algorithm>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define CPB_DEVICE static inline
#define CPB_HD static inline

struct emu_dim3 { unsigned x, y, z; emu_dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
typedef emu_dim3 dim3;
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

struct double2 { double x, y; };
struct double4 { double x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y){ ulonglong2 v; v.x = x; v.y = y; return v; }
static inline double2 make_double2(double x, double y){ double2 r = {x, y}; return r; }
static inline double4 make_double4(double x, double y, double z, double w){ double4 r = {x, y, z, w}; return r; }
static inline int2 make_int2(int x, int y){ int2 r = {x, y}; return r; }
static inline int4 make_int4(int x, int y, int z, int w){ int4 r = {x, y, z, w}; return r; }

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
typedef void *cudaGraphExec_t;
#define cudaSuccess 0
enum { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline cudaError_t cudaMalloc(void **p, size_t n){ *p = calloc(1, n ? n : 1); return 0; }
static inline cudaError_t cudaFree(void *p){ free(p); return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n){ *p = calloc(1, n ? n : 1); return 0; }
static inline cudaError_t cudaFreeHost(void *p){ free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t){ memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, int){ memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t){ memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t){ return 0; }
static inline cudaError_t cudaDeviceSynchronize(){ return 0; }
static inline cudaError_t cudaGetLastError(){ return 0; }
static inline const char *cudaGetErrorString(cudaError_t){ return "emu"; }
static inline cudaError_t cudaSetDevice(int){ return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s){ *s = 0; return 0; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t){ return 0; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e){ *e = 0; return 0; }
#define cudaEventDisableTiming 2
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned){ *e = 0; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned){ return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t){ return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t){ return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t){ return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t){ *ms = 0; return 0; }

template <typename K, typename... Args>
static inline void emu_launch(K kernel, unsigned grid, unsigned block, Args... args)
{
	gridDim = emu_dim3(grid); blockDim = emu_dim3(block);
	for(unsigned b = 0; b < grid; b++){
		blockIdx = emu_dim3(b);
		for(unsigned t = 0; t < block; t++){
			threadIdx = emu_dim3(t);
			kernel(args...);
		}
	}
}
#define LAUNCH(kernel, grid, block, stream, ...) emu_launch(kernel, (unsigned)(grid), (unsigned)(block), __VA_ARGS__)

// atomics: sequential semantics
template <typename T> static inline T atomicAdd(T *p, T v){ T o = *p; *p = o + v; return o; }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v){ unsigned long long o = *p; if(v > o) *p = v; return o; }
static inline unsigned atomicMax(unsigned *p, unsigned v){ unsigned o = *p; if(v > o) *p = v; return o; }
static inline int atomicMax(int *p, int v){ int o = *p; if(v > o) *p = v; return o; }
static inline int atomicMin(int *p, int v){ int o = *p; if(v < o) *p = v; return o; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v){ unsigned long long o = *p; if(v < o) *p = v; return o; }
static inline unsigned atomicMin(unsigned *p, unsigned v){ unsigned o = *p; if(v < o) *p = v; return o; }
static inline unsigned long long atomicCAS(unsigned long long *p, unsigned long long c, unsigned long long v){ unsigned long long o = *p; if(o == c) *p = v; return o; }
static inline unsigned atomicCAS(unsigned *p, unsigned c, unsigned v){ unsigned o = *p; if(o == c) *p = v; return o; }
static inline int atomicCAS(int *p, int c, int v){ int o = *p; if(o == c) *p = v; return o; }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v){ unsigned long long o = *p; *p = o | v; return o; }
static inline unsigned atomicOr(unsigned *p, unsigned v){ unsigned o = *p; *p = o | v; return o; }
static inline int atomicExch(int *p, int v){ int o = *p; *p = v; return o; }
static inline unsigned atomicExch(unsigned *p, unsigned v){ unsigned o = *p; *p = v; return o; }
static inline void __threadfence(){}
static inline int __clz(int x){ return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x){ return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffsll(long long x){ return __builtin_ffsll(x); }
static inline int __popc(unsigned x){ return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x){ return __builtin_popcountll(x); }
template <typename T> static inline T __ldg(const T *p){ return *p; }
static inline double __longlong_as_double(long long v){ double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d){ long long v; memcpy(&v, &d, 8); return v; }
#endif

// ------------------------------------------------------------------ common helpers
#define CPB_CHECK(expr) do { cudaError_t e_ = (expr); if(e_ != cudaSuccess){ cpb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while(0)

void cpb_set_error(const char *fmt, ...);

CPB_HD int cpb_div_up(int a, int b){ return (a + b - 1)/b; }

// global thread id / grid size for grid-stride loops
#define CPB_TID ((int)(blockIdx.x*blockDim.x + threadIdx.x))
#define CPB_NTHREADS ((int)(gridDim.x*blockDim.x))
