// libdvbt_b200.so: error text, device selection, launch counter.
#include "common.cuh"
#include "chain_internal.cuh"

#include <atomic>
#include <stdlib.h>
#include <string.h>

namespace dvbt {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int ensure_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error("no CUDA device available (%s); libdvbt_b200 has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return DVBT_B200_ENODEV;
  }
  return 0;
}

static std::atomic<int> g_blocking_wait{-1};   // -1: not decided yet (DVBT_B200_BLOCKING_WAIT), 0 spin, 1 block
void set_blocking_wait(int on) { g_blocking_wait.store(on ? 1 : 0); }

cudaError_t stream_wait(cudaStream_t st) {
  int blocking = g_blocking_wait.load(std::memory_order_relaxed);
  if (blocking < 0) {
    blocking = (getenv("DVBT_B200_BLOCKING_WAIT") && atoi(getenv("DVBT_B200_BLOCKING_WAIT")) != 0) ? 1 : 0;
    g_blocking_wait.store(blocking);
  }
  if (!blocking) return cudaStreamSynchronize(st);
  thread_local cudaEvent_t ev[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return cudaStreamSynchronize(st);
  if (!ev[dev]) {
    cudaError_t e = cudaEventCreateWithFlags(&ev[dev], cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) { ev[dev] = nullptr; return cudaStreamSynchronize(st); }
  }
  cudaError_t e = cudaEventRecord(ev[dev], st);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ev[dev]);
}

int Staging::ensure() {
  for (int b = 0; b < 2; b++) {
    if (!pin[b] && cudaMallocHost(&pin[b], kChunk) != cudaSuccess) { pin[b] = nullptr; set_error("staging: cudaMallocHost failed"); return DVBT_B200_ENOMEM; }
    if (!ev[b]) DVBT_CUDA_TRY(cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming));
  }
  return 0;
}

void Staging::release() {
  for (int b = 0; b < 2; b++) {
    if (pin[b]) cudaFreeHost(pin[b]);
    if (ev[b]) cudaEventDestroy(ev[b]);
    pin[b] = nullptr;
    ev[b] = nullptr;
  }
}

static bool host_pointer_is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

static cudaError_t event_wait_yield(cudaEvent_t e) { return cudaEventSynchronize(e); }

int Staging::h2d(void *d_dst, const void *h_src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return 0;
  if (bytes < 65536 || host_pointer_is_pinned(h_src)) {
    DVBT_CUDA_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, st));
    return 0;
  }
  if (int rc = ensure()) return rc;
  size_t off = 0;
  for (int i = 0; off < bytes; i++) {
    const int b = i & 1;
    const size_t n = bytes - off < kChunk ? bytes - off : kChunk;
    if (i >= 2) DVBT_CUDA_TRY(event_wait_yield(ev[b]));       // the DMA that last read this pinned buffer is done
    memcpy(pin[b], (const char *)h_src + off, n);
    DVBT_CUDA_TRY(cudaMemcpyAsync((char *)d_dst + off, pin[b], n, cudaMemcpyHostToDevice, st));
    DVBT_CUDA_TRY(cudaEventRecord(ev[b], st));
    off += n;
  }
  // the pinned buffers are reused by the next call: its first two chunks must not overtake this call's DMAs
  DVBT_CUDA_TRY(event_wait_yield(ev[0]));
  DVBT_CUDA_TRY(event_wait_yield(ev[1]));
  return 0;
}

int Staging::d2h(void *h_dst, const void *d_src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return 0;
  if (bytes < 65536 || host_pointer_is_pinned(h_dst)) {
    DVBT_CUDA_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, st));
    DVBT_CUDA_TRY(stream_wait(st));
    return 0;
  }
  if (int rc = ensure()) return rc;
  const size_t nchunks = (bytes + kChunk - 1) / kChunk;
  auto issue = [&](size_t i) -> int {
    const size_t off = i * kChunk, n = bytes - off < kChunk ? bytes - off : kChunk;
    DVBT_CUDA_TRY(cudaMemcpyAsync(pin[i & 1], (const char *)d_src + off, n, cudaMemcpyDeviceToHost, st));
    DVBT_CUDA_TRY(cudaEventRecord(ev[i & 1], st));
    return 0;
  };
  if (int rc = issue(0)) return rc;
  for (size_t i = 0; i < nchunks; i++) {
    if (i + 1 < nchunks) { if (int rc = issue(i + 1)) return rc; }     // the other pinned buffer was drained in the previous iteration
    DVBT_CUDA_TRY(event_wait_yield(ev[i & 1]));
    const size_t off = i * kChunk, n = bytes - off < kChunk ? bytes - off : kChunk;
    memcpy((char *)h_dst + off, pin[i & 1], n);
  }
  return 0;
}

void energy_prbs_table(uint8_t tab[1504]) {
  unsigned reg = 0xa9;
  auto clock8 = [&]() {
    unsigned res = 0;
    for (int i = 0; i < 8; i++) {
      unsigned fb = ((reg >> 13) ^ (reg >> 14)) & 1u;
      reg = ((reg << 1) | fb) & 0x7fff;
      res = (res << 1) | fb;
    }
    return (uint8_t)res;
  };
  for (int pk = 0; pk < 8; pk++) {
    tab[pk * 188] = 0;
    for (int k = 1; k < 188; k++) tab[pk * 188 + k] = clock8();
    clock8();
  }
}

}  // namespace dvbt

extern "C" {

const char *dvbt_b200_last_error(void) { return dvbt::g_err; }

int dvbt_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int dvbt_b200_set_device(int device) {
  int rc = dvbt::ensure_device();
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaSetDevice(device));
  return 0;
}

unsigned long long dvbt_b200_kernel_launches(void) { return dvbt::g_launches.load(); }
int dvbt_b200_set_blocking_wait(int on) { dvbt::set_blocking_wait(on); return 0; }

}  // extern "C"
