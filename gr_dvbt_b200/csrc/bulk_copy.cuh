// Bulk asynchronous global -> shared copy (the 1-D form of TMA: cp.async.bulk, UBLKCP in SASS) with an mbarrier that
// counts the bytes in: one thread arms the barrier with the byte count and issues the copy, the copy engine moves the
// data while the block does other work, every thread then waits on the barrier's phase.  Source, destination and size
// must be multiples of 16 bytes.
#pragma once
#include <stdint.h>

namespace dvbt {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible to the async proxy before the copy is issued
}

// arm the barrier for `bytes` and start the copy (call from ONE thread, after a block barrier that follows mbar_init)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// wait until the phase with the given parity has completed (all bytes have landed)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

}  // namespace dvbt
