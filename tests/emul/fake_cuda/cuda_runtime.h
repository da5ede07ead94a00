// Stand-in for <cuda_runtime.h> when a .cu file of gr_dvbt_b200/csrc is compiled WHOLE (host orchestration included) for
// the host by tests/emul/build_vit_emul.py: "device" memory is host memory, streams and events do nothing, a kernel
// launch (rewritten from <<< >>> by the builder) runs the kernel on host threads (cuda_host_emul.h).  Test infrastructure.
#pragma once
#include "../cuda_host_emul.h"

#include <math.h>

struct float2 { float x, y; };
struct double2 { double x, y; };
struct int2 { int x, y; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uchar4 { unsigned char x, y, z, w; };
struct char4 { signed char x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }

// single / double operations with one rounding each (the TU is built with -ffp-contract=off)
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline unsigned __float_as_uint(float f) { unsigned v; memcpy(&v, &f, 4); return v; }
static inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float2int_rn(float f) { return (int)lrintf(f); }
static inline float __double2float_rn(double d) { return (float)d; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline long long clock64() { return 0; }
template <class T> static inline size_t __cvta_generic_to_shared(T *) { return 0; }
#define __grid_constant__

// ---- runtime API
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmul = 1, cudaErrorNotReady = 600 };
typedef struct emul_stream *cudaStream_t;
typedef struct emul_event *cudaEvent_t;
#define cudaStreamLegacy ((cudaStream_t)0)
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyHostToHost = 0 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventBlockingSync = 1 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaDevAttrMultiProcessorCount = 16, cudaDevAttrClockRate = 13 };
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated CUDA runtime"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int *v, int attr, int) { *v = attr == cudaDevAttrMultiProcessorCount ? 148 : 1965000; return cudaSuccess; }
#ifdef DVBT_B200_EXACT_ALLOC
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorEmul; }
#else
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n + 256, 1); return *p ? cudaSuccess : cudaErrorEmul; }
#endif
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, int) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dpitch, const void *s, size_t spitch, size_t width, size_t height, int, cudaStream_t = nullptr) {
  for (size_t r = 0; r < height; r++) memmove((char *)d + r * dpitch, (const char *)s + r * spitch, width);
  return cudaSuccess;
}
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void *)&(sym), (src), (n)), cudaSuccess)
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = (cudaStream_t)1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest) { *least = 0; *greatest = -5; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = (cudaStream_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; };
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (cudaEvent_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = (cudaEvent_t)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }

// two-dimensional grids: one pass over the x dimension per y
template <class Kernel, class... Args>
static void emul_launch(Kernel kern, dim3 grid, unsigned block, Args... args) {
  gridDim.y = grid.y;
  for (unsigned y = 0; y < grid.y; y++) {
    emul_cur_block_y = y;
    emul_launch(kern, grid.x, block, args...);
  }
  emul_cur_block_y = 0;
  gridDim.y = 1;
}
