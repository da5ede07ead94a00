// Throughput of the 16x2 DPX instruction VIADDMNMX.U16x2 on sm_100a next to LOP3 / PRMT / IADD3 / IMAD, alone and
// mixed, to decide whether a Viterbi ACS on packed (metric:8 | path:8) halfwords pays.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dpx_rate dpx_rate.cu && ./dpx_rate
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

constexpr int ITERS = 4096, ACC = 8;

template <int MODE> __global__ void __launch_bounds__(256) rate_kernel(uint32_t *out, uint32_t a, uint32_t b, uint32_t one) {
  uint32_t v[ACC], w[ACC];
#pragma unroll
  for (int i = 0; i < ACC; i++) { v[i] = threadIdx.x * 2654435761u + i; w[i] = v[i] ^ a; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ACC; i++) {
      if (MODE == 0) v[i] = __viaddmax_u16x2(v[i], a, w[i]);                       // VIADDMNMX.U16x2
      if (MODE == 1) v[i] = (v[i] & a) ^ w[i];                                     // LOP3
      if (MODE == 2) asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(v[i]) : "r"(v[i]), "r"(w[i]));   // PRMT
      if (MODE == 3) v[i] = v[i] + w[i] + a;                                       // IADD3
      if (MODE == 4) asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v[i]) : "r"(v[i]), "r"(one), "r"(w[i]));  // IMAD
      if (MODE == 5) {                                                             // ACS pair: IMAD + VIADDMNMX
        uint32_t c; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(c) : "r"(w[i]), "r"(one), "r"(b));
        v[i] = __viaddmax_u16x2(v[i], a, c);
      }
      if (MODE == 6) v[i] = __vimax3_u16x2(v[i], w[i], w[i]);                         // VIMNMX.U16x2
      if (MODE == 7) v[i] = __vimax3_u16x2(v[i], w[i], a);                         // VIMNMX3.U16x2
      if (MODE == 8) {                                                             // 1 VIADDMNMX + 1 LOP3
        v[i] = __viaddmax_u16x2(v[i], a, w[i]);
        w[i] = (w[i] & b) ^ v[i];
      }
      if (MODE == 9) {                                                             // plain add (compiler's choice) + VIADDMNMX
        uint32_t c = w[i] + b;
        v[i] = __viaddmax_u16x2(v[i], a, c);
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < ACC; i++) r ^= v[i] ^ w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char *name, int ops_per_iter, uint32_t *d) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int tpb : {128, 256, 512}) {
    int blocks = sms * (1024 / tpb);
    rate_kernel<MODE><<<blocks, tpb>>>(d, 0x01000100u, 0x02010201u, 1u);
    cudaEventRecord(e0);
    rate_kernel<MODE><<<blocks, tpb>>>(d, 0x01000100u, 0x02010201u, 1u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * tpb * ITERS * ACC * ops_per_iter;
    printf("%-28s tpb %3d: %7.3f ms  %6.1f thread-ops/clk/SM (at %d MHz nominal)\n", name, tpb, ms, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  }
}

int main() {
  uint32_t *d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  run<0>("VIADDMNMX.U16x2", 1, d);
  run<1>("LOP3", 1, d);
  run<2>("PRMT", 1, d);
  run<3>("IADD3", 1, d);
  run<4>("IMAD", 1, d);
  run<5>("IMAD + VIADDMNMX", 2, d);
  run<6>("VIMNMX.U16x2", 1, d);
  run<7>("VIMNMX3.U16x2", 1, d);
  run<8>("VIADDMNMX + LOP3", 2, d);
  run<9>("add + VIADDMNMX", 2, d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
