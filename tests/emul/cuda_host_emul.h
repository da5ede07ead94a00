// Minimal host emulation of the CUDA execution model, for running the Viterbi device code of
// gr_dvbt_b200/csrc/viterbi.cu on a CPU (test infrastructure; see tests/emul/build_vit_emul.py).
//
// One CUDA thread = one host thread (threadIdx / blockIdx are thread_local), the threads of a block run
// concurrently and __syncthreads() is a barrier over them; blocks run one after another; dynamic shared memory is
// one static buffer per process.  Warp shuffles / ballots / __syncwarp are lock-step exchanges over a per-warp barrier
// (converged, full-mask use only).
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

struct uint4 { uint32_t x, y, z, w; };   // not over-aligned: the host compiler then never assumes 16-byte alignment
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
struct uint2 { uint32_t x, y; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }

using std::max;
using std::min;

// dynamic shared memory of the running block for whole-file builds: `extern __shared__ T name[];` becomes
// `T *name = (T *)emul_dyn_smem;`
alignas(128) static unsigned char emul_dyn_smem[256 * 1024];

static unsigned emul_cur_block_y = 0;   // second grid dimension: set by the dim3 overload of emul_launch
static std::barrier<> *emul_block_barrier = nullptr;
static inline void __syncthreads() { emul_block_barrier->arrive_and_wait(); }

// Warp primitives: the threads of a warp (32 consecutive threads of the block) meet at a per-warp barrier and exchange
// through a per-warp buffer.  Every live thread of the warp must take part (the member masks are not interpreted: the
// kernels that are emulated call them converged, with full masks).
struct EmulWarp {
  std::barrier<> bar;
  uint32_t x[32];
  explicit EmulWarp(int n) : bar(n) {}
};
static thread_local EmulWarp *emul_warp = nullptr;
static inline void __syncwarp(unsigned = 0xffffffffu) { emul_warp->bar.arrive_and_wait(); }
template <class T, class F>
static inline T emul_exchange(T v, F pick) {
  static_assert(sizeof(T) == 4, "32-bit values only");
  const int lane = (int)(threadIdx.x & 31u);
  uint32_t u;
  memcpy(&u, &v, 4);
  emul_warp->x[lane] = u;
  emul_warp->bar.arrive_and_wait();
  const int src = pick(lane);
  uint32_t r = (src >= 0 && src < 32) ? emul_warp->x[src] : u;
  emul_warp->bar.arrive_and_wait();
  T out;
  memcpy(&out, &r, 4);
  return out;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int lanemask) { return emul_exchange(v, [=](int l) { return l ^ lanemask; }); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, int delta) { return emul_exchange(v, [=](int l) { return l - delta; }); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int delta) { return emul_exchange(v, [=](int l) { return l + delta < 32 ? l + delta : -1; }); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return emul_exchange(v, [=](int) { return src & 31; }); }
static inline unsigned __ballot_sync(unsigned, int pred) {
  const int lane = (int)(threadIdx.x & 31u);
  emul_warp->x[lane] = pred ? 1u : 0u;
  emul_warp->bar.arrive_and_wait();
  unsigned r = 0;
  const int n = (int)std::min(32u, blockDim.x - (threadIdx.x & ~31u));
  for (int i = 0; i < n; i++) r |= emul_warp->x[i] << i;
  emul_warp->bar.arrive_and_wait();
  return r;
}
static inline int __all_sync(unsigned m, int pred) {
  const unsigned n = std::min(32u, blockDim.x - (threadIdx.x & ~31u));
  return __ballot_sync(m, pred) == (n == 32 ? 0xffffffffu : ((1u << n) - 1u));
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <class T> static inline T __ldg(const T *p) { return *p; }
#define __constant__

static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
static inline int atomicMin(int *p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_RELAXED);
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
// funnel shift left: the upper 32 bits of (hi:lo) << (shift & 31)
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t shift) {
  return (uint32_t)(((((uint64_t)hi << 32) | lo) << (shift & 31u)) >> 32);
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
// __byte_perm(x, y, s): selector indices are the low three bits of each of the four low nibbles of s
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) { return emul_prmt(x, y, s & 0x7777u); }
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

// kernel<<<grid, block, smem>>>(args...): `block` host threads live for the whole launch and walk the blocks together
// (creating them per block cost more than the kernels themselves); between two blocks they meet twice at an outer
// barrier while thread 0 renews the per-block and per-warp barriers, which exited threads had dropped out of.
template <class Kernel, class... Args>
static void emul_launch(Kernel kern, unsigned grid, unsigned block, Args... args) {
  gridDim.x = grid;
  blockDim.x = block;
  if (grid == 0 || block == 0) return;
  std::barrier<> outer((std::ptrdiff_t)block);
  std::unique_ptr<std::barrier<>> inner;
  std::vector<std::unique_ptr<EmulWarp>> warps((block + 31) / 32);
  const unsigned y = emul_cur_block_y;
  auto renew = [&]() {
    inner.reset(new std::barrier<>((std::ptrdiff_t)block));
    emul_block_barrier = inner.get();
    for (unsigned w = 0; w < warps.size(); w++) warps[w].reset(new EmulWarp((int)std::min(32u, block - 32 * w)));
  };
  renew();
  std::vector<std::thread> th;
  th.reserve(block);
  for (unsigned t = 0; t < block; t++)
    th.emplace_back([&, t]() {
      threadIdx.x = t;
      blockIdx.y = y;
      for (unsigned b = 0; b < grid; b++) {
        blockIdx.x = b;
        emul_warp = warps[t >> 5].get();
        kern(args...);
        emul_warp->bar.arrive_and_drop();   // a thread that has returned no longer takes part in later barriers
        inner->arrive_and_drop();
        if (b + 1 < grid) {
          outer.arrive_and_wait();          // every thread has left block b
          if (t == 0) renew();
          outer.arrive_and_wait();          // the barriers of block b + 1 are in place
        }
      }
    });
  for (auto &x : th) x.join();
  emul_block_barrier = nullptr;
}
