// Front end of the RX flowgraphs: rational_resampler_ccc(interp 64, decim 70) followed by
// multiply_const (apps/dvbt_rx_demo*.grc), i.e. the 10 Msps capture -> 64/7 Msps OFDM rate.
//
// These are stock GNU Radio blocks whose source is not under /root/reference:
// the kernel follows their documented GNU Radio 3.7 semantics and is "parity unpinned":
//   filter.rational_resampler_ccc(64, 70): gcd-reduced to 32/35; taps = firdes.low_pass(32, 32,
//     rate*0.5 - tw/2, tw = rate*(0.5-0.4), WIN_KAISER, beta 7.0) (python/rational_resampler.py
//     design_filter); polyphase arms h[phase + 32 j]; y[m] = sum_j h[(35 m mod 32) + 32 j] *
//     x[floor(35 m / 32) - j] with zero history (rational_resampler_base_ccc::general_work);
//   blocks.multiply_const_vcc((k,)): y * (k + 0j).
// A block stages the inputs of 512 outputs in shared memory (each input read once), taps in shared
// memory (36 per arm), float accumulation in tap order.  HBM bound: 8 B read (x 35/32) + 8 B written
// per output sample.
#include "chain_internal.cuh"

#include <math.h>
#include <vector>

namespace dvbt {

static double izero(double x) {  // gr::fft::window Izero, IzeroEPSILON = 1e-21
  double sum = 1, u = 1, halfx = x / 2.0, temp;
  int n = 1;
  do {
    temp = halfx / (double)n;
    n += 1;
    temp *= temp;
    u *= temp;
    sum += u;
  } while (u >= 1e-21 * sum);
  return sum;
}

// firdes::low_pass(gain, fs, fc, tw, WIN_KAISER, beta) of GNU Radio 3.7
static std::vector<float> firdes_low_pass_kaiser(double gain, double fs, double fc, double tw, double beta) {
  double a = beta / 0.1102 + 8.7;  // max_attenuation(WIN_KAISER)
  int ntaps = (int)(a * fs / (22.0 * tw));
  if ((ntaps & 1) == 0) ntaps++;
  std::vector<float> w(ntaps), taps(ntaps);
  double ibeta = 1.0 / izero(beta), inm1 = 1.0 / (double)(ntaps - 1);
  for (int i = 0; i < ntaps; i++) {
    double t = 2 * i * inm1 - 1;
    w[i] = (float)(izero(beta * sqrt(1.0 - t * t)) * ibeta);
  }
  int M = (ntaps - 1) / 2;
  double fwT0 = 2 * M_PI * fc / fs;
  for (int n = -M; n <= M; n++) {
    if (n == 0) taps[n + M] = (float)(fwT0 / M_PI * w[n + M]);
    else taps[n + M] = (float)(sin(n * fwT0) / (n * M_PI) * w[n + M]);
  }
  double fmax = taps[M];
  for (int n = 1; n <= M; n++) fmax += 2 * taps[n + M];
  gain /= fmax;
  for (int i = 0; i < ntaps; i++) taps[i] = (float)(taps[i] * gain);
  return taps;
}

constexpr int kInterp = 32, kDecim = 35;

int resampler_taps(std::vector<float> *out, int *per_arm) {
  double rate = (double)kInterp / kDecim;
  double tw = rate * (0.5 - 0.4);
  double mid = rate * 0.5 - tw / 2.0;
  std::vector<float> t = firdes_low_pass_kaiser(kInterp, kInterp, mid, tw, 7.0);
  while (t.size() % kInterp) t.push_back(0.f);  // install_taps pads to a multiple of the arm count
  *per_arm = (int)(t.size() / kInterp);
  *out = t;
  return 0;
}

// taps laid out arm-major: tap[phase * per_arm + j] = h[phase + 32 j].  A block produces 512 outputs
// (2 per thread); the 560 + per_arm input samples they need are staged in shared memory with coalesced
// loads, so every input sample is read from HBM/L2 once.
constexpr int kOutPerBlock = 512;
constexpr int kTileIn = (kOutPerBlock * kDecim) / kInterp + 2;  // inputs advanced by a block (+ slack)

__global__ void __launch_bounds__(256) resample_kernel(const float2 *__restrict__ x, long long nin, float2 *__restrict__ y, long long nout,
                                                       const float *__restrict__ taps_arm, int per_arm, float scale) {
  extern __shared__ float s_mem[];
  float *s_taps = s_mem;                                        // [32][per_arm]
  float2 *s_x = reinterpret_cast<float2 *>(s_mem + kInterp * (per_arm | 1));  // [kTileIn + per_arm], 8-byte aligned
  for (int i = threadIdx.x; i < kInterp * per_arm; i += blockDim.x) s_taps[(i / per_arm) * (per_arm | 1) + i % per_arm] = taps_arm[i];
  long long m0 = (long long)blockIdx.x * kOutPerBlock;
  long long a0 = (m0 * kDecim) / kInterp;                       // newest input of the block's first output
  long long lo = a0 - (per_arm - 1);                            // oldest input needed
  int span = kTileIn + per_arm;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    long long idx = lo + i;
    s_x[i] = (idx >= 0 && idx < nin) ? x[idx] : make_float2(0.f, 0.f);
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 2; r++) {
    long long m = m0 + threadIdx.x + 256 * r;
    if (m >= nout) return;
    long long t = m * kDecim;
    long long a = t / kInterp;
    int phase = (int)(t - a * kInterp);
    const float *h = s_taps + phase * (per_arm | 1);          // odd row stride: no bank conflicts between phases
    const float2 *xs = s_x + (int)(a - lo);                     // xs[-j] = x[a - j]
    float accr = 0.f, acci = 0.f;
    for (int j = 0; j < per_arm; j++) {
      float2 v = xs[-j];
      accr = fmaf(h[j], v.x, accr);
      acci = fmaf(h[j], v.y, acci);
    }
    y[m] = make_float2(accr * scale, acci * scale);
  }
}

// Fast path for the 36-taps-per-arm prototype.  The phase of output m is (35 m) mod 32, so a thread that
// steps m by a multiple of 32 always uses the same arms: it owns the output pair (2u, 2u+1), keeps both
// arms' 36 taps in registers, and loads the 38 input samples the two overlapping windows span once
// (the windows start 1 or 2 samples apart, fixed per thread): 19 8-byte loads and 72 FMAs per output.
constexpr int kRegArm = 36, kRegPairs = 4;

template <int DELTA>
__device__ __forceinline__ void resample_pair_loop(const float2 *__restrict__ x, long long nin, float2 *__restrict__ y, long long nout,
                                                   const float *__restrict__ taps_arm, float scale, long long u) {
  float h0[kRegArm], h1[kRegArm];
  {
    int p0 = (int)(((2 * u) * kDecim) % kInterp), p1 = (int)(((2 * u + 1) * kDecim) % kInterp);
#pragma unroll
    for (int j = 0; j < kRegArm; j++) { h0[j] = taps_arm[p0 * kRegArm + j]; h1[j] = taps_arm[p1 * kRegArm + j]; }
  }
#pragma unroll 1
  for (int r = 0; r < kRegPairs; r++, u += 256) {
    long long m0 = 2 * u, m1 = m0 + 1;
    if (m0 >= nout) return;
    long long a1 = (m1 * kDecim) / kInterp;   // newest input of the second output; the first one's is a1 - DELTA
    float2 xs[kRegArm + DELTA];
    if (a1 >= kRegArm + DELTA - 1 && a1 < nin) {
#pragma unroll
      for (int k = 0; k < kRegArm + DELTA; k++) xs[k] = __ldg(x + a1 - k);
    } else {
#pragma unroll
      for (int k = 0; k < kRegArm + DELTA; k++) {
        long long idx = a1 - k;
        xs[k] = (idx >= 0 && idx < nin) ? x[idx] : make_float2(0.f, 0.f);
      }
    }
    float r0 = 0.f, i0 = 0.f, r1 = 0.f, i1 = 0.f;
#pragma unroll
    for (int j = 0; j < kRegArm; j++) {
      r0 = fmaf(h0[j], xs[j + DELTA].x, r0);
      i0 = fmaf(h0[j], xs[j + DELTA].y, i0);
      r1 = fmaf(h1[j], xs[j].x, r1);
      i1 = fmaf(h1[j], xs[j].y, i1);
    }
    y[m0] = make_float2(r0 * scale, i0 * scale);
    if (m1 < nout) y[m1] = make_float2(r1 * scale, i1 * scale);
  }
}

__global__ void __launch_bounds__(256) resample_reg_kernel(const float2 *__restrict__ x, long long nin, float2 *__restrict__ y, long long nout,
                                                           const float *__restrict__ taps_arm, float scale) {
  long long u = (long long)blockIdx.x * (256 * kRegPairs) + threadIdx.x;
  long long a0 = ((2 * u) * kDecim) / kInterp, a1 = ((2 * u + 1) * kDecim) / kInterp;
  if (a1 - a0 == 1) resample_pair_loop<1>(x, nin, y, nout, taps_arm, scale, u);
  else resample_pair_loop<2>(x, nin, y, nout, taps_arm, scale, u);
}

struct Resampler {
  DevBuf d_taps;
  int per_arm = 0;
  int init() {
    std::vector<float> t;
    resampler_taps(&t, &per_arm);
    std::vector<float> arm((size_t)kInterp * per_arm);
    for (int p = 0; p < kInterp; p++)
      for (int j = 0; j < per_arm; j++) arm[(size_t)p * per_arm + j] = t[p + kInterp * j];
    int rc = d_taps.reserve(arm.size() * 4);
    if (rc) return rc;
    DVBT_CUDA_TRY(cudaMemcpy(d_taps.p, arm.data(), arm.size() * 4, cudaMemcpyHostToDevice));
    return 0;
  }
};

static Resampler *g_res[64] = {nullptr};

long long resample_out_count(long long nin) { return nin <= 0 ? 0 : ((nin - 1) * kInterp) / kDecim + 1; }

int resample_launch(const float2 *d_x, long long nin, float2 *d_y, long long nout, float scale, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64) dev = 63;
  if (!g_res[dev]) {
    g_res[dev] = new Resampler();
    int rc = g_res[dev]->init();
    if (rc) return rc;
  }
  Resampler *r = g_res[dev];
  if (nout <= 0) return 0;
  if (r->per_arm == kRegArm) {
    long long per_block = 2LL * 256 * kRegPairs;
    resample_reg_kernel<<<(unsigned)((nout + per_block - 1) / per_block), 256, 0, st>>>(d_x, nin, d_y, nout, r->d_taps.as<float>(), scale);
    count_launch();
    DVBT_CUDA_TRY(cudaGetLastError());
    return 0;
  }
  size_t smem = (size_t)(kInterp * (r->per_arm | 1) + 2) * 4 + (size_t)(kTileIn + r->per_arm) * 8;
  resample_kernel<<<(unsigned)((nout + kOutPerBlock - 1) / kOutPerBlock), 256, smem, st>>>(d_x, nin, d_y, nout, r->d_taps.as<float>(), r->per_arm, scale);
  count_launch();
  DVBT_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace dvbt

extern "C" {

// test/inspection hook: the 32/35 low-pass prototype (length is a multiple of 32)
int dvbt_b200_resampler_taps(float *taps, int capacity) {
  std::vector<float> t;
  int per_arm = 0;
  dvbt::resampler_taps(&t, &per_arm);
  if (!taps || capacity < (int)t.size()) return (int)t.size();
  for (size_t i = 0; i < t.size(); i++) taps[i] = t[i];
  return (int)t.size();
}

}  // extern "C"
