// Internal helpers shared by the CUDA translation units of libdvbt_b200.so.
#pragma once
#include <cuda_runtime.h>
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

// Waits for a stream.  Default: cudaStreamSynchronize (the driver spins - lowest latency, right for one thread per GPU).
// DVBT_B200_BLOCKING_WAIT=1: record an event created with cudaEventBlockingSync and sleep on it - for hosts that drive more
// waiting threads than they have cores to spin on (8 ranks x 4 captures in flight on 32 cores lost 14 % to descheduled
// spinners); the wake-up latency is hidden when other captures keep the GPU busy.  (A polling wait that yields the core
// between cudaStreamQuery calls was tried too: four polling threads contend with each other's kernel launches inside the
// driver, 12.7 ms per batch of four captures instead of 7.1.)
cudaError_t stream_wait(cudaStream_t st);
void set_blocking_wait(int on);

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

// Host <-> device copies of the block-level entry points (`*_work`): the scheduler's buffers are pageable, and a
// cudaMemcpyAsync from pageable memory is staged by the driver in small pieces (measured 2-4 GB/s on the B200 boxes,
// the whole cost of a drop-in block call).  A handle therefore stages through two pinned buffers of its own: the host
// memcpy of one chunk runs while the DMA of the previous one is in flight.  Pinned caller memory (cudaHostAlloc /
// cudaHostRegister) is recognised and copied directly.
struct Staging {
  static constexpr size_t kChunk = 2u << 20;
  void *pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int ensure();
  void release();
  // enqueue the copy of `bytes` from host memory to device memory on `st` (returns when the host buffer has been read)
  int h2d(void *d_dst, const void *h_src, size_t bytes, cudaStream_t st);
  // copy device memory to host memory; returns when the host buffer holds the data (the stream's earlier work is waited for)
  int d2h(void *h_dst, const void *d_src, size_t bytes, cudaStream_t st);
};

}  // namespace dvbt
