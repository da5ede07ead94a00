// Internal helpers shared by the CUDA translation units of libdvbt_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <sched.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dvbt_b200.h"

namespace dvbt {

void set_error(const char *fmt, ...);
void count_launch(unsigned n = 1);
int ensure_device();  // 0 or a negative DVBT_B200_E* code (sets the error text)

#define DVBT_CUDA_TRY(expr)                                                              \
  do {                                                                                   \
    cudaError_t err__ = (expr);                                                          \
    if (err__ != cudaSuccess) {                                                          \
      dvbt::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,               \
                      cudaGetErrorString(err__));                                        \
      return DVBT_B200_ECUDA;                                                            \
    }                                                                                    \
  } while (0)

// A handle lives on the device that was current when it was created (dvbt_b200_set_device).  CUDA's
// current device is per host thread, so every entry point that takes a handle enters this scope: a
// handle may be driven from any thread (GNU Radio runs one thread per block) on a multi-GPU host.
struct DeviceScope {
  int prev = -1, dev = -1;
  explicit DeviceScope(int device) : dev(device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (dev >= 0 && prev != dev) cudaSetDevice(dev);
  }
  ~DeviceScope() {
    if (prev >= 0 && dev >= 0 && prev != dev) cudaSetDevice(prev);
  }
};
inline int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) d = 0;
  return d;
}

// Entry points that take DEVICE pointers run on the handle's own non-blocking stream, which does not order itself
// against the legacy default stream - where cudaMemset/cudaMemcpy and PyTorch's allocations, fills and copies run.
// Work the caller queued there before the call (e.g. zero-filling the output buffer) must not overtake or trail the
// handle's kernels, so the handle's stream first waits for everything already queued on the default stream.
inline int join_default_stream(cudaStream_t st) {
  cudaEvent_t ev;
  DVBT_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ev, cudaStreamLegacy);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev, 0);
  cudaEventDestroy(ev);   // released once the wait has been satisfied
  DVBT_CUDA_TRY(e);
  return 0;
}

// Waits for a stream the way cudaStreamSynchronize does, but gives the core away between polls.  The library is driven by
// one host thread per handle (GNU Radio: one per block); with as many waiting threads as cores - 8 ranks x 4 captures in
// flight on a 32-core host - a spinning synchronise that the kernel deschedules for a time slice stalls its capture for
// milliseconds (measured: 14 % at 8 GPUs).  sched_yield returns at once when nothing else wants the core.
inline cudaError_t stream_wait(cudaStream_t st) {
  for (;;) {
    cudaError_t e = cudaStreamQuery(st);
    if (e != cudaErrorNotReady) return e;
    sched_yield();
  }
}

// A growable device (or pinned-host) buffer; never shrinks.
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool host = false;
  int reserve(size_t bytes) {
    if (bytes <= cap) return 0;
    release();
#ifdef DVBT_B200_EXACT_ALLOC   // memcheck builds (tests/emul under AddressSanitizer): only the 16 bytes that the
    size_t want = bytes + 16;  // 16-byte staging helpers may read past a range by contract, no slack that would hide more
#else
    size_t want = bytes + bytes / 8 + 256;
#endif
    cudaError_t e = host ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      p = nullptr;
      cap = 0;
      set_error("%s of %zu bytes failed: %s", host ? "cudaMallocHost" : "cudaMalloc", want,
                cudaGetErrorString(e));
      return DVBT_B200_ENOMEM;
    }
    cap = want;
    return 0;
  }
  void release() {
    if (p) {
      if (host) cudaFreeHost(p); else cudaFree(p);
    }
    p = nullptr;
    cap = 0;
  }
  template <class T> T *as() const { return (T *)p; }
};

}  // namespace dvbt
