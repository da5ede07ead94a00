// Minimal host emulation of the CUDA execution model, for running the Viterbi device code of
// gr_dvbt_b200/csrc/viterbi.cu on a CPU (test infrastructure; see tests/emul/build_vit_emul.py).
//
// One CUDA thread = one host thread (threadIdx / blockIdx are thread_local), the threads of a block run
// concurrently and __syncthreads() is a barrier over them; blocks run one after another; dynamic shared memory is
// one static buffer per process.  Warp-level primitives are NOT emulated (the one-lane ACS kernels, the verify and
// repair kernels and the depuncture kernel do not use any); calling one aborts.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static   /* function-local shared variables; `extern __shared__` is rewritten to `extern` by the builder */
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)

struct emul_dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local emul_dim3 threadIdx, blockIdx;
static emul_dim3 blockDim, gridDim;

struct alignas(16) uint4 { uint32_t x, y, z, w; };

using std::max;
using std::min;

static std::barrier<> *emul_block_barrier = nullptr;
static inline void __syncthreads() { emul_block_barrier->arrive_and_wait(); }

[[noreturn]] static inline void emul_unsupported(const char *what) {
  fprintf(stderr, "cuda_host_emul: %s is not emulated\n", what);
  abort();
}
static inline void __syncwarp(unsigned = 0xffffffffu) { emul_unsupported("__syncwarp"); }
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t, int) { emul_unsupported("__shfl_xor_sync"); }
static inline uint32_t __shfl_sync(unsigned, uint32_t, int) { emul_unsupported("__shfl_sync"); }

static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
static int emul_block_or = 0;
static inline int __syncthreads_or(int pred) {
  if (pred) __atomic_fetch_or(&emul_block_or, 1, __ATOMIC_RELAXED);
  __syncthreads();
  const int r = __atomic_load_n(&emul_block_or, __ATOMIC_RELAXED);
  __syncthreads();
  if (threadIdx.x == 0) __atomic_store_n(&emul_block_or, 0, __ATOMIC_RELAXED);
  __syncthreads();
  return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t shift) {
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (shift & 31u));
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline unsigned __brev(unsigned v) {
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  return __builtin_bswap32(v);
}

// prmt.b32 (default mode): selector nibble = byte index into {a, b}; bit 3 replicates the sign bit of that byte
static inline uint32_t emul_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; i++) {
    const unsigned nib = (sel >> (4 * i)) & 0xfu;
    uint32_t byte = (uint32_t)(src >> (8 * (nib & 7u))) & 0xffu;
    if (nib & 8u) byte = (byte & 0x80u) ? 0xffu : 0x00u;
    r |= byte << (8 * i);
  }
  return r;
}
// add.u16x2 + max.u16x2 (what __viaddmax_u16x2 expands to on sm_90+): per halfword max((a + b) mod 2^16, c)
static inline uint32_t __viaddmax_u16x2(uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t lo = max((a + b) & 0xffffu, c & 0xffffu);
  const uint32_t hi = max(((a >> 16) + (b >> 16)) & 0xffffu, c >> 16);
  return lo | (hi << 16);
}
static inline uint32_t __vimax3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t lo = max(max(a & 0xffffu, b & 0xffffu), c & 0xffffu);
  const uint32_t hi = max(max(a >> 16, b >> 16), c >> 16);
  return lo | (hi << 16);
}

// kernel<<<grid, block, smem>>>(args...)
template <class Kernel, class... Args>
static void emul_launch(Kernel kern, unsigned grid, unsigned block, Args... args) {
  gridDim.x = grid;
  blockDim.x = block;
  for (unsigned b = 0; b < grid; b++) {
    std::barrier<> bar((std::ptrdiff_t)block);
    emul_block_barrier = &bar;
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; t++)
      th.emplace_back([=, &bar]() {
        threadIdx.x = t;
        blockIdx.x = b;
        kern(args...);
        bar.arrive_and_drop();   // a thread that has returned no longer takes part in later barriers
      });
    for (auto &x : th) x.join();
  }
  emul_block_barrier = nullptr;
}
