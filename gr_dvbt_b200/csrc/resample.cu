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
#include <stdlib.h>
#include <mutex>
#include <new>
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

// python/rational_resampler.py design_filter(interpolation, decimation, fractional_bw = 0.4) of GNU Radio 3.7
int resampler_taps_for(int interp, int decim, std::vector<float> *out, int *per_arm) {
  const double rate = (double)interp / decim, halfband = 0.5, fractional_bw = 0.4;
  double tw, mid;
  if (rate >= 1.0) { tw = halfband - fractional_bw; mid = halfband - tw / 2.0; }
  else { tw = rate * (halfband - fractional_bw); mid = rate * halfband - tw / 2.0; }
  std::vector<float> t = firdes_low_pass_kaiser(interp, interp, mid, tw, 7.0);
  while (t.size() % interp) t.push_back(0.f);  // install_taps pads to a multiple of the arm count
  *per_arm = (int)(t.size() / interp);
  *out = t;
  return 0;
}

int resampler_taps(std::vector<float> *out, int *per_arm) { return resampler_taps_for(kInterp, kDecim, out, per_arm); }

// taps laid out arm-major: tap[phase * per_arm + j] = h[phase + 32 j].  A block produces 512 outputs
// (2 per thread); the 560 + per_arm input samples they need are staged in shared memory with coalesced
// loads, so every input sample is read from HBM/L2 once.
constexpr int kOutPerBlock = 512;
constexpr int kTileIn = (kOutPerBlock * kDecim) / kInterp + 2;  // inputs advanced by a block (+ slack)

__global__ void __launch_bounds__(256) resample_kernel(const float2 *__restrict__ x, long long nin, float2 *__restrict__ y, long long nout,
                                                       const float *__restrict__ taps_arm, int per_arm, float scale, int nhist) {
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
    s_x[i] = (idx >= -(long long)nhist && idx < nin) ? x[idx] : make_float2(0.f, 0.f);
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

// Fast path for the 36-taps-per-arm prototype.
//   * A thread owns four consecutive outputs 4v..4v+3: their windows overlap, so 36+off3 (<= 40) input
//     samples held in registers feed 4 x 36 taps (10 shared-memory loads per output instead of 36).
//   * The polyphase arm of output m is (35 m) mod 32, so the four arms - and the distances between the four
//     windows - depend only on R = v mod 8.  Warp w of a block takes the quads with v mod 8 = w: arms and
//     window offsets are compile-time constants of the instantiation resample_quad<w>; the warp reads its
//     taps from shared memory with 16-byte broadcast loads (one address per warp: one wavefront each).
//     (Taps as constant-bank operands were tried: 36 dependent LDCU.128 per iteration left the kernel
//     latency bound at 0.59 ms.)
//   * Lane i of warp w owns quad v = 256 T + 8 i + w of tile T: lane windows start exactly 35 samples apart,
//     35 = 3 (mod 16), so the 8-byte shared-memory reads of a half-warp hit 16 different banks.
//   * A tile (1024 outputs, 1120 + 40 inputs) is staged with coalesced loads, the next tile's loads are in
//     flight during the FMAs, results leave through a padded shared-memory tile as coalesced 8-byte stores.
// Accumulation order is tap 0..35 per output, one float FMA chain each for re and im, as in the generic kernel.
constexpr int kRegArm = 36;
constexpr int kQuadOut = 1024;                       // outputs per tile
constexpr int kQuadIn = 1120;                        // inputs a tile advances by (1024 * 35 / 32)
constexpr int kQuadSpan = 1160;                      // staged inputs: 35 history + 1120 + window slack
constexpr int kQuadPerThread = (kQuadSpan + 255) / 256;

// o1, o2, o3: window starts of outputs 4v+1..4v+3 relative to output 4v.  Only three combinations occur
// over the eight residues ((1,2,3) six times, (1,2,4) for R = 2, (1,3,4) for R = 5), so the kernel holds
// three copies of the FMA block instead of eight (the eight-copy version was instruction-cache bound).
template <int o1, int o2, int o3>
__device__ __forceinline__ void resample_quad(const float2 *__restrict__ s_x, const float *__restrict__ s_taps, float4 *__restrict__ s_y, int lane,
                                              int R, float scale) {
  const int f = (12 * R) % 32;                       // (35 * 4v) mod 32 for v = R (mod 8)
  const int p0 = f, p1 = (f + 3) % 32, p2 = (f + 6) % 32, p3 = (f + 9) % 32;
  const int top = 35 * lane + (35 * R) / 8 + 35 + o3;  // tile-local index of the newest input of output 4v+3
  float2 xs[kRegArm + o3];
#pragma unroll
  for (int k = 0; k < kRegArm + o3; k++) xs[k] = s_x[top - k];
  float r0 = 0.f, i0 = 0.f, r1 = 0.f, i1 = 0.f, r2 = 0.f, i2 = 0.f, r3 = 0.f, i3 = 0.f;
  const float4 *t0 = reinterpret_cast<const float4 *>(s_taps + p0 * kRegArm), *t1 = reinterpret_cast<const float4 *>(s_taps + p1 * kRegArm),
               *t2 = reinterpret_cast<const float4 *>(s_taps + p2 * kRegArm), *t3 = reinterpret_cast<const float4 *>(s_taps + p3 * kRegArm);
#pragma unroll
  for (int jj = 0; jj < kRegArm / 4; jj++) {
    const float4 h0 = t0[jj], h1 = t1[jj], h2 = t2[jj], h3 = t3[jj];
    const float a0[4] = {h0.x, h0.y, h0.z, h0.w}, a1[4] = {h1.x, h1.y, h1.z, h1.w}, a2[4] = {h2.x, h2.y, h2.z, h2.w},
                a3[4] = {h3.x, h3.y, h3.z, h3.w};
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int j = 4 * jj + u;
      r0 = fmaf(a0[u], xs[j + o3].x, r0);
      i0 = fmaf(a0[u], xs[j + o3].y, i0);
      r1 = fmaf(a1[u], xs[j + o3 - o1].x, r1);
      i1 = fmaf(a1[u], xs[j + o3 - o1].y, i1);
      r2 = fmaf(a2[u], xs[j + o3 - o2].x, r2);
      i2 = fmaf(a2[u], xs[j + o3 - o2].y, i2);
      r3 = fmaf(a3[u], xs[j].x, r3);
      i3 = fmaf(a3[u], xs[j].y, i3);
    }
  }
  const int q = 8 * lane + R;                        // quad index inside the tile
  const int unit = 2 * q + q / 8;                    // 16-byte units, one pad unit per 8 quads: conflict-free stores
  s_y[unit] = make_float4(r0 * scale, i0 * scale, r1 * scale, i1 * scale);
  s_y[unit + 1] = make_float4(r2 * scale, i2 * scale, r3 * scale, i3 * scale);
}

__global__ void __launch_bounds__(256, 3) resample_quad_kernel(const float2 *__restrict__ x, long long nin, float2 *__restrict__ y, long long nout,
                                                               const float *__restrict__ taps_arm, long long ntiles, float scale, int nhist) {
  __shared__ float2 s_x[kQuadSpan];
  __shared__ __align__(16) float s_taps[kInterp * kRegArm];
  __shared__ float4 s_y[kQuadOut / 2 + kQuadOut / 32];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  float2 pre[kQuadPerThread];
  auto fetch = [&](long long tile) {
    long long lo = tile * kQuadIn - 35;
#pragma unroll
    for (int k = 0; k < kQuadPerThread; k++) {
      int i = t + 256 * k;
      long long idx = lo + i;
      pre[k] = (i < kQuadSpan && idx >= -(long long)nhist && idx < nin) ? __ldg(x + idx) : make_float2(0.f, 0.f);
    }
  };
  for (int i = t; i < kInterp * kRegArm; i += 256) s_taps[i] = taps_arm[i];
  long long tile = blockIdx.x;
  if (tile < ntiles) fetch(tile);
  for (; tile < ntiles; tile += gridDim.x) {
#pragma unroll
    for (int k = 0; k < kQuadPerThread; k++) {
      int i = t + 256 * k;
      if (i < kQuadSpan) s_x[i] = pre[k];
    }
    __syncthreads();
    if (tile + gridDim.x < ntiles) fetch(tile + gridDim.x);
    if (w == 2) resample_quad<1, 2, 4>(s_x, s_taps, s_y, lane, w, scale);
    else if (w == 5) resample_quad<1, 3, 4>(s_x, s_taps, s_y, lane, w, scale);
    else resample_quad<1, 2, 3>(s_x, s_taps, s_y, lane, w, scale);
    __syncthreads();
    const float2 *s_y2 = reinterpret_cast<const float2 *>(s_y);
    long long m0 = tile * kQuadOut;
#pragma unroll
    for (int k = 0; k < kQuadOut / 256; k++) {
      int o = t + 256 * k;
      int q = o >> 2;
      long long m = m0 + o;
      if (m < nout) y[m] = s_y2[2 * (2 * q + (q >> 3) + ((o >> 1) & 1)) + (o & 1)];
    }
  }
}

// Two (NQ) quads per thread.  The quad kernel above is bound by shared-memory bandwidth, and two thirds of it are
// tap loads: a 16-byte load costs the full four wavefronts even when every lane reads the same address, so the 36 tap
// loads of a quad cost 144 cycles per warp against 80 for its 40 input samples (ncu: l1tex 91 %, mio throttle).  A
// thread therefore takes NQ quads with the SAME residue - quads 256 apart, i.e. the same position in NQ consecutive
// 1024-output sub-tiles - so that one set of tap loads feeds NQ x 4 outputs (lanes stay 35 samples apart: the 8-byte
// input reads remain conflict free).  The input window of a quad slides through registers (4 + o3 samples live per
// quad instead of 36 + o3), and the tile is staged with cp.async into a double buffer instead of through registers.
template <int NQ> struct MultiCfg {
  static constexpr int kOut = kQuadOut * NQ;            // outputs per tile
  static constexpr int kIn = kQuadIn * NQ;              // inputs a tile advances by
  static constexpr int kSpan = kQuadIn * NQ + 40;       // staged inputs: 35 history + kIn + window slack
  static constexpr int kYUnits = kOut / 2 + kOut / 32;  // float4 units of the padded output tile
  static constexpr size_t kSmem = (size_t)2 * kSpan * 8 + (size_t)kYUnits * 16 + (size_t)kInterp * kRegArm * 4;
};

template <int o1, int o2, int o3, int NQ>
__device__ __forceinline__ void resample_multi(const float2 *__restrict__ s_x, const float *__restrict__ s_taps, float4 *__restrict__ s_y, int lane,
                                               int R, float scale) {
  const int f = (12 * R) % 32;                       // (35 * 4v) mod 32 for v = R (mod 8)
  const int p0 = f, p1 = (f + 3) % 32, p2 = (f + 6) % 32, p3 = (f + 9) % 32;
  const int top = 35 * lane + (35 * R) / 8 + 35 + o3;  // sub-tile-local index of the newest input of output 4v+3
  float2 xs[NQ][kRegArm + o3];                       // compile-time indexed: only a sliding part is ever live
  float acc[NQ][8];
#pragma unroll
  for (int q = 0; q < NQ; q++) {
#pragma unroll
    for (int k = 0; k < o3; k++) xs[q][k] = s_x[top + q * kQuadIn - k];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[q][k] = 0.f;
  }
  const float4 *t0 = reinterpret_cast<const float4 *>(s_taps + p0 * kRegArm), *t1 = reinterpret_cast<const float4 *>(s_taps + p1 * kRegArm),
               *t2 = reinterpret_cast<const float4 *>(s_taps + p2 * kRegArm), *t3 = reinterpret_cast<const float4 *>(s_taps + p3 * kRegArm);
#pragma unroll
  for (int jj = 0; jj < kRegArm / 4; jj++) {
    const float4 h0 = t0[jj], h1 = t1[jj], h2 = t2[jj], h3 = t3[jj];
    const float a0[4] = {h0.x, h0.y, h0.z, h0.w}, a1[4] = {h1.x, h1.y, h1.z, h1.w}, a2[4] = {h2.x, h2.y, h2.z, h2.w},
                a3[4] = {h3.x, h3.y, h3.z, h3.w};
#pragma unroll
    for (int q = 0; q < NQ; q++) {
#pragma unroll
      for (int u = 0; u < 4; u++) xs[q][4 * jj + o3 + u] = s_x[top + q * kQuadIn - (4 * jj + o3 + u)];
    }
#pragma unroll
    for (int q = 0; q < NQ; q++) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = 4 * jj + u;
        acc[q][0] = fmaf(a0[u], xs[q][j + o3].x, acc[q][0]);
        acc[q][1] = fmaf(a0[u], xs[q][j + o3].y, acc[q][1]);
        acc[q][2] = fmaf(a1[u], xs[q][j + o3 - o1].x, acc[q][2]);
        acc[q][3] = fmaf(a1[u], xs[q][j + o3 - o1].y, acc[q][3]);
        acc[q][4] = fmaf(a2[u], xs[q][j + o3 - o2].x, acc[q][4]);
        acc[q][5] = fmaf(a2[u], xs[q][j + o3 - o2].y, acc[q][5]);
        acc[q][6] = fmaf(a3[u], xs[q][j].x, acc[q][6]);
        acc[q][7] = fmaf(a3[u], xs[q][j].y, acc[q][7]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < NQ; q++) {
    const int qi = 8 * lane + R + 256 * q;           // quad index inside the tile
    const int unit = 2 * qi + qi / 8;                // 16-byte units, one pad unit per 8 quads: conflict-free stores
    s_y[unit] = make_float4(acc[q][0] * scale, acc[q][1] * scale, acc[q][2] * scale, acc[q][3] * scale);
    s_y[unit + 1] = make_float4(acc[q][4] * scale, acc[q][5] * scale, acc[q][6] * scale, acc[q][7] * scale);
  }
}

template <int NQ>
__global__ void __launch_bounds__(256, 3) resample_multi_kernel(const float2 *__restrict__ x, long long nin, float2 *__restrict__ y, long long nout,
                                                                const float *__restrict__ taps_arm, long long ntiles, float scale, int nhist) {
  using Cfg = MultiCfg<NQ>;
  extern __shared__ __align__(16) unsigned char s_raw[];
  float4 *s_y = reinterpret_cast<float4 *>(s_raw);
  float *s_taps = reinterpret_cast<float *>(s_raw + (size_t)Cfg::kYUnits * 16);
  float2 *s_xbuf = reinterpret_cast<float2 *>(s_raw + (size_t)Cfg::kYUnits * 16 + (size_t)kInterp * kRegArm * 4);
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  // stage the inputs of `tile` into buffer b: 8-byte cp.async, samples outside [0, nin) are zero filled (src-size 0)
  auto stage = [&](long long tile, int b) {
    const long long lo = tile * Cfg::kIn - 35;
    float2 *dst = s_xbuf + (size_t)b * Cfg::kSpan;
    for (int i = t; i < Cfg::kSpan; i += 256) {
      const long long idx = lo + i;
      const bool in = idx >= -(long long)nhist && idx < nin;
      const unsigned d = (unsigned)__cvta_generic_to_shared(dst + i);
      const float2 *src = x + (in ? idx : 0);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(in ? 8 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int i = t; i < kInterp * kRegArm; i += 256) s_taps[i] = taps_arm[i];
  long long tile = blockIdx.x;
  int b = 0;
  if (tile < ntiles) stage(tile, 0);
  for (; tile < ntiles; tile += gridDim.x, b ^= 1) {
    const bool more = tile + gridDim.x < ntiles;
    if (more) stage(tile + gridDim.x, b ^ 1);        // the other buffer was last read before the barrier that ended the previous tile
    if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const float2 *s_x = s_xbuf + (size_t)b * Cfg::kSpan;
    if (w == 2) resample_multi<1, 2, 4, NQ>(s_x, s_taps, s_y, lane, w, scale);
    else if (w == 5) resample_multi<1, 3, 4, NQ>(s_x, s_taps, s_y, lane, w, scale);
    else resample_multi<1, 2, 3, NQ>(s_x, s_taps, s_y, lane, w, scale);
    __syncthreads();
    const float2 *s_y2 = reinterpret_cast<const float2 *>(s_y);
    const long long m0 = tile * Cfg::kOut;
#pragma unroll
    for (int k = 0; k < Cfg::kOut / 256; k++) {
      int o = t + 256 * k;
      int q = o >> 2;
      long long m = m0 + o;
      if (m < nout) y[m] = s_y2[2 * (2 * q + (q >> 3) + ((o >> 1) & 1)) + (o & 1)];
    }
    // the next iteration's first barrier orders these reads of s_y before its writes
  }
}

struct Resampler {
  DevBuf d_taps;
  int per_arm = 0, sm_count = 148;
  int init() {
    std::vector<float> t;
    resampler_taps(&t, &per_arm);
    std::vector<float> arm((size_t)kInterp * per_arm);
    for (int p = 0; p < kInterp; p++)
      for (int j = 0; j < per_arm; j++) arm[(size_t)p * per_arm + j] = t[p + kInterp * j];
    int rc = d_taps.reserve(arm.size() * 4);
    if (rc) return rc;
    DVBT_CUDA_TRY(cudaMemcpy(d_taps.p, arm.data(), arm.size() * 4, cudaMemcpyHostToDevice));
    int dev = 0;
    cudaGetDevice(&dev);
    DVBT_CUDA_TRY(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    return 0;
  }
};

// one per device, built under a lock and published only when complete (several handles are driven from
// concurrent host threads); a failed init is not cached
static Resampler *g_res[64] = {nullptr};
static std::mutex g_res_mutex;
static int resampler_for_device(int dev, Resampler **out) {
  std::lock_guard<std::mutex> lock(g_res_mutex);
  if (!g_res[dev]) {
    Resampler *r = new (std::nothrow) Resampler();
    if (!r) { set_error("resampler: out of memory"); return DVBT_B200_ENOMEM; }
    int rc = r->init();
    if (rc) { r->d_taps.release(); delete r; return rc; }
    g_res[dev] = r;
  }
  *out = g_res[dev];
  return 0;
}

long long resample_out_count(long long nin) { return nin <= 0 ? 0 : ((nin - 1) * kInterp) / kDecim + 1; }

int resample_launch_variant(const float2 *d_x, long long nin, float2 *d_y, long long nout, float scale, cudaStream_t st, int variant, int nhist = 0);

int resample_launch(const float2 *d_x, long long nin, float2 *d_y, long long nout, float scale, cudaStream_t st, int nhist) {
  static const int nq = getenv("DVBT_B200_RESAMPLE_NQ") ? atoi(getenv("DVBT_B200_RESAMPLE_NQ")) : 2;
  return resample_launch_variant(d_x, nin, d_y, nout, scale, st, nq, nhist);
}

// variant: 0 generic kernel (any tap count), 1 quad kernel, 2 / 4 quads per thread (36 taps per arm)
int resample_launch_variant(const float2 *d_x, long long nin, float2 *d_y, long long nout, float scale, cudaStream_t st, int nq, int nhist) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 64) dev = 63;
  Resampler *r = nullptr;
  if (int rc = resampler_for_device(dev, &r)) return rc;
  if (nout <= 0) return 0;
  if (r->per_arm == kRegArm && (nq == 2 || nq == 4)) {
    const long long out_per_tile = (long long)kQuadOut * nq;
    long long ntiles = (nout + out_per_tile - 1) / out_per_tile;
    long long grid = ntiles < 3LL * r->sm_count ? ntiles : 3LL * r->sm_count;
    if (nq == 2) {
      DVBT_CUDA_TRY(cudaFuncSetAttribute(resample_multi_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MultiCfg<2>::kSmem));
      resample_multi_kernel<2><<<(unsigned)grid, 256, MultiCfg<2>::kSmem, st>>>(d_x, nin, d_y, nout, r->d_taps.as<float>(), ntiles, scale, nhist);
    } else {
      DVBT_CUDA_TRY(cudaFuncSetAttribute(resample_multi_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MultiCfg<4>::kSmem));
      resample_multi_kernel<4><<<(unsigned)grid, 256, MultiCfg<4>::kSmem, st>>>(d_x, nin, d_y, nout, r->d_taps.as<float>(), ntiles, scale, nhist);
    }
    count_launch();
    DVBT_CUDA_TRY(cudaGetLastError());
    return 0;
  }
  if (r->per_arm == kRegArm && nq == 1) {
    long long ntiles = (nout + kQuadOut - 1) / kQuadOut;
    long long grid = ntiles < 3LL * r->sm_count ? ntiles : 3LL * r->sm_count;  // persistent: three resident blocks per SM
    resample_quad_kernel<<<(unsigned)grid, 256, 0, st>>>(d_x, nin, d_y, nout, r->d_taps.as<float>(), ntiles, scale, nhist);
    count_launch();
    DVBT_CUDA_TRY(cudaGetLastError());
    return 0;
  }
  size_t smem = (size_t)(kInterp * (r->per_arm | 1) + 2) * 4 + (size_t)(kTileIn + r->per_arm) * 8;
  resample_kernel<<<(unsigned)((nout + kOutPerBlock - 1) / kOutPerBlock), 256, smem, st>>>(d_x, nin, d_y, nout, r->d_taps.as<float>(), r->per_arm, scale, nhist);
  count_launch();
  DVBT_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace dvbt

extern "C" {

// the front end alone on host buffers (parity tests of the resampler kernels): nin complex samples at 10 Msps ->
// resample_out_count(nin) samples at 64/7 Msps times gain.  variant selects the kernel (0 generic, 1 quad, 2 / 4
// quads per thread, -1 the default of the chain); all variants add the taps in the same order.
int dvbt_b200_resample_host(const void *in, size_t nin, float gain, void *out, size_t out_capacity, size_t *nout, int variant) {
  if ((nin && !in) || !out || !nout) { dvbt::set_error("resample_host: null argument"); return DVBT_B200_EINVAL; }
  *nout = 0;
  int rc = dvbt::ensure_device();
  if (rc) return rc;
  long long n = dvbt::resample_out_count((long long)nin);
  if ((size_t)n > out_capacity) { dvbt::set_error("resample_host: need %lld output samples, capacity %zu", n, out_capacity); return DVBT_B200_ENOSPC; }
  if (n <= 0) return 0;
  dvbt::DevBuf d_in, d_out;
  if ((rc = d_in.reserve(nin * 8)) || (rc = d_out.reserve((size_t)n * 8))) { d_in.release(); d_out.release(); return rc; }
  cudaError_t e = cudaMemcpy(d_in.p, in, nin * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    rc = variant < 0 ? dvbt::resample_launch(d_in.as<float2>(), (long long)nin, d_out.as<float2>(), n, gain, nullptr)
                     : dvbt::resample_launch_variant(d_in.as<float2>(), (long long)nin, d_out.as<float2>(), n, gain, nullptr, variant);
    if (!rc) e = cudaMemcpy(out, d_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost);
  }
  d_in.release();
  d_out.release();
  if (rc) return rc;
  if (e != cudaSuccess) { dvbt::set_error("resample_host: %s", cudaGetErrorString(e)); return DVBT_B200_ECUDA; }
  *nout = (size_t)n;
  return 0;
}

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
