// Stand-in for <cufft.h> (see cuda_runtime.h in this directory): batched 1-D complex transforms of power-of-two length
// as a plain radix-2 FFT in double precision (test infrastructure: the emulated library uses it where the real one
// calls cuFFT - the transmit chain's inverse FFT and the optional unfused acquisition FFT).
#pragma once
#include <math.h>
#include <vector>
typedef int cufftHandle;
typedef int cufftResult;
typedef float2 cufftComplex;
enum { CUFFT_SUCCESS = 0, CUFFT_EMULATED = 1, CUFFT_C2C = 0x29, CUFFT_FORWARD = -1, CUFFT_INVERSE = 1 };
struct emul_cufft_plan { int n, batch; };
static std::vector<emul_cufft_plan> &emul_cufft_plans() { static std::vector<emul_cufft_plan> p(1); return p; }
static inline cufftResult cufftPlanMany(cufftHandle *h, int rank, int *n, int *, int, int idist, int *, int, int odist, int type, int batch) {
  if (rank != 1 || type != CUFFT_C2C || (n[0] & (n[0] - 1)) || idist != n[0] || odist != n[0]) return CUFFT_EMULATED;
  emul_cufft_plans().push_back(emul_cufft_plan{n[0], batch});
  *h = (int)emul_cufft_plans().size() - 1;
  return CUFFT_SUCCESS;
}
static inline cufftResult cufftSetStream(cufftHandle, cudaStream_t) { return CUFFT_SUCCESS; }
static inline cufftResult cufftExecC2C(cufftHandle h, cufftComplex *in, cufftComplex *out, int dir) {
  if (h <= 0 || h >= (int)emul_cufft_plans().size()) return CUFFT_EMULATED;
  const int n = emul_cufft_plans()[h].n, batch = emul_cufft_plans()[h].batch;
  std::vector<double> re(n), im(n);
  for (int b = 0; b < batch; b++) {
    for (int i = 0; i < n; i++) {        // bit reversal
      int r = 0;
      for (int k = 1, v = i; k < n; k <<= 1, v >>= 1) r = (r << 1) | (v & 1);
      re[r] = in[(size_t)b * n + i].x;
      im[r] = in[(size_t)b * n + i].y;
    }
    for (int len = 2; len <= n; len <<= 1) {
      const double ang = (dir == CUFFT_FORWARD ? -2.0 : 2.0) * M_PI / len;
      for (int i = 0; i < n; i += len)
        for (int k = 0; k < len / 2; k++) {
          const double wr = cos(ang * k), wi = sin(ang * k);
          const int a = i + k, c = i + k + len / 2;
          const double xr = re[c] * wr - im[c] * wi, xi = re[c] * wi + im[c] * wr;
          re[c] = re[a] - xr; im[c] = im[a] - xi;
          re[a] += xr; im[a] += xi;
        }
    }
    for (int i = 0; i < n; i++) out[(size_t)b * n + i] = make_float2((float)re[i], (float)im[i]);
  }
  return CUFFT_SUCCESS;
}
static inline cufftResult cufftDestroy(cufftHandle) { return CUFFT_SUCCESS; }
