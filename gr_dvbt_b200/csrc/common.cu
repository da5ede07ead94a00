// libdvbt_b200.so: error text, device selection, launch counter.
#include "common.cuh"
#include "chain_internal.cuh"

#include <atomic>

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

}  // extern "C"
