// K2 — ofdm_sym_acquisition on sm_100a: van de Beek ML cyclic-prefix timing + fractional CFO,
// peak tracking, derotation and CP removal; K3 — batched cuFFT with the fft_vxx "shift" folded in.
//
// Replaces gr::dvbt::ofdm_sym_acquisition (lib/ofdm_sym_acquisition_impl.cc):
//   general_work        :488-568  initial acquisition / tracking (+-8) / restart state machine
//   ml_sync             :148-351  lambda[k] = |gamma| - rho/2 * phi over candidate symbol ends,
//                                 epsilon = atan2(gamma[peak]), phase accumulator, derot[]
//   peak_detect_process :72-146   rise/fall threshold detector with a running average
// and the stock fft_vxx(N, forward, rectangular, shift=True) that follows it in the flowgraphs.
//
// The reference handles one symbol per call and re-sums cp products for each of the 16
// candidates.  Here a whole capture (or one piece of a stream) is processed at once:
//   acq_init_lambda_kernel / acq_init_peak_kernel / acq_probe_kernel
//                       the one-off search over N candidates (same lambda values, same detector)
//                       and a 4-symbol look-ahead that sizes the first tracking batch;
//   acq_lambda_kernel   one thread per (symbol, candidate): gamma, phi, lambda with the
//                       reference's summation order (so lambda is bit-identical: it depends only
//                       on the absolute sample position);
//   acq_pass1_kernel / acq_pass2_kernel
//                       the detector's running average d_avg is the only state that crosses symbols
//                       besides cp_start; it is speculated (pass 1: from zero per window offset,
//                       pass 2: the detector for every (offset, previous offset) state from the pass-1
//                       value) and every speculated input is verified bit-for-bit;
//   acq_chunkmap_kernel / acq_compose_kernel / acq_walk_kernel
//                       tracking = composition of 187-state maps over 32-symbol chunks; where the
//                       speculation stops, the reference's sequential detector continues on the spot;
//   acq_post_kernel / acq_finish_kernel / acq_desc_kernel
//                       peak, epsilon and the phase-increment schedule (:285-312) as scans, giving
//                       per-symbol (first sample, start phase, increments);
//   acq_fftd_kernel     out = FFT((-1)^j expj(phase_j) in[cp_start-N+1+j]): derotation, CP removal,
//                       fft_vxx's shift and the forward FFT in one kernel (acq_derot_kernel + cuFFT
//                       for other sizes or time-domain output).
// A missed peak restarts acquisition half a symbol later (:545-558); every acquisition attempt sends
// sync_start at the current output position (:507) - acq_run records those offsets for the caller.
// Difference to the reference, by construction: the derotation phase is evaluated in closed
// form in double instead of N+cp sequential float additions per symbol (:285-309), so the
// output samples agree to ~1e-6 relative, not bit-for-bit; decisions (timing, peaks) are exact.
#include "chain_internal.cuh"

#include <cufft.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

namespace {

using dvbt::set_error;

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cmul_conjf(float2 a, float2 b) { return cmulf(a, make_float2(b.x, -b.y)); }
__device__ __forceinline__ float cnormf(float2 a) { return __fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)); }

// lambda/gamma for the candidate symbol end at absolute sample index p (ofdm_sym_acquisition_impl.cc:162-250).
// lo: lowest readable index of x (0, or -history).  The reference's window reaches up to 8 samples in front of its input
// pointer when the peak sits at the very start of the search range (earlier items of the scheduler's circular
// buffer; zeros on a fresh buffer), the candidate table here up to kD: what lies in front of `lo` reads as zero - the
// start of a stream - and callers that continue a stream keep kAcqHistory consumed samples in front of x.
__device__ __forceinline__ void ml_point(const float2 *__restrict__ x, long long p, int N, int cp, float rho2, float *lambda,
                                         float2 *gamma, long long lo = 0) {
  float2 g = make_float2(0.f, 0.f);
  float phi = 0.f;
  if (p - (cp - 1) - N < lo) {   // rare: only candidates within kD samples of the start of the buffer
    for (int j = 0; j < cp; j++) {
      const long long ia = p - j, ib = p - j - N;
      float2 a = ia >= lo ? x[ia] : make_float2(0.f, 0.f), b = ib >= lo ? x[ib] : make_float2(0.f, 0.f);
      float2 c = cmul_conjf(a, b);
      g = make_float2(__fadd_rn(g.x, c.x), __fadd_rn(g.y, c.y));
      phi = __fadd_rn(phi, __fadd_rn(cnormf(a), cnormf(b)));
    }
  } else
  for (int j = 0; j < cp; j++) {
    float2 a = x[p - j], b = x[p - j - N];
    float2 c = cmul_conjf(a, b);                       // d_corr[i-j-N] = in[i-j] * conj(in[i-j-N])
    g = make_float2(__fadd_rn(g.x, c.x), __fadd_rn(g.y, c.y));
    phi = __fadd_rn(phi, __fadd_rn(cnormf(a), cnormf(b)));
  }
  float mag = __fsqrt_rn(cnormf(g));                   // volk_32fc_magnitude_32f
  *lambda = __fsub_rn(mag, __fmul_rn(phi, rho2));      // s32f_multiply then x2_subtract
  *gamma = g;
}

struct PeakState {
  float avg;
};

// peak_detect_process (:72-146) over n values; returns number of peaks, *best = index of the peak of peaks
// The reference loop re-examines a value after a state change without advancing; written here as "one
// value, up to three state visits" so that the loads do not depend on the detector state and can be
// issued ahead of the (inherently serial) float recurrence on d_avg.
__device__ int peak_detect(const float *d, int n, float *avg_io, float rise, float fall, float alpha, int *best) {
  float avg = *avg_io;
  int state = 0, npeaks = 0, peak_index = 0, best_idx = 0;
  float peak_val = -INFINITY, peak_at = 0.f, best_val = 0.f;
  const float one_minus = 1.0f - alpha;  // (1 - d_avg_alpha) in float
  auto step = [&](float v, int i) {
    for (;;) {
      if (state == 0) {
        if (v > __fmul_rn(avg, rise)) { state = 1; continue; }
        break;
      }
      if (v > peak_val) { peak_val = v; peak_at = v; peak_index = i; break; }
      if (v > __fmul_rn(avg, fall)) break;
      // record the peak; the first strictly greatest recorded peak wins (:127-137)
      if (npeaks == 0 || peak_at > best_val) { best_val = peak_at; best_idx = peak_index; }
      npeaks++;
      state = 0;
      peak_val = -INFINITY;
    }
    avg = __fadd_rn(__fmul_rn(alpha, v), __fmul_rn(one_minus, avg));
  };
  int i = 0;
  for (; i + 4 <= n; i += 4) {
    float v0 = d[i], v1 = d[i + 1], v2 = d[i + 2], v3 = d[i + 3];
    step(v0, i);
    step(v1, i + 1);
    step(v2, i + 2);
    step(v3, i + 3);
  }
  for (; i < n; i++) step(d[i], i);
  *avg_io = avg;
  *best = best_idx;
  return npeaks;
}

struct AcqParams {
  int N, cp;
  float rho2;                 // (float)(d_rho / 2.0)
  float rise, fall, alpha;    // 0.8, 0.9, 0.9 (:448)
};

constexpr int kD = 16;        // candidate table half width around the speculated cp_start
constexpr int kCand = 2 * kD;

// sequential state of the block (device resident)
struct AcqState {
  int initial;        // d_initial_aquisition
  int cp_start;       // relative to the current read position
  float avg;          // peak detector running average
  float phase;        // d_phase
  double phaseinc, nextphaseinc;
  int nextpos;
  // results of the last batch
  long long consumed; // samples consumed
  int n_out;          // symbols produced
  int n_sync_tags;    // sync_start tags (first symbol of a (re)acquisition)
  int lost_at;        // symbol count at which tracking missed (restart), -1 if never
  int fallback;       // 1 if the sequential detector had to be used
  int n_run, n_single, n_seq;  // symbols handled in quiet runs / one at a time / by the sequential detector
  int probe_lost1;    // acq_probe_kernel: 1 + index of the first of the next kProbe symbols that misses its peak, 0 = none
};

struct SymOut {
  long long first;    // absolute index of the first of the N samples
  double phase0;      // phase before the first increment of this call
  double inc0, inc1;  // increment before / from nextpos
  int switch_at;      // sample index at which inc1 takes over (>= N+cp: never)
};

// ---- initial acquisition: lambda over N candidates of the window at `base`
__global__ void acq_init_lambda_kernel(AcqParams p, const float2 *__restrict__ x, long long base, float *__restrict__ lambda,
                                       float2 *__restrict__ gamma) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= p.N) return;
  ml_point(x, base + (p.N + p.cp - 1) + k, p.N, p.cp, p.rho2, &lambda[k], &gamma[k]);
}

// Initial acquisition: peak_detect over the N candidate positions of one symbol.  Every advance of the
// reference loop updates d_avg the same way whatever the detector state (:88-118), so the average seen by
// element i does not depend on the state machine: thread 0 runs only that float recurrence (one FMUL + one
// FADD of latency per element, alpha*v precomputed), all threads then evaluate the two threshold tests of
// every element in parallel into bit masks, and thread 0 replays the state machine on the masks, skipping
// the long below-threshold runs with bit scans.  Same decisions as peak_detect(), 5x less latency.
__global__ void __launch_bounds__(256) acq_init_peak_kernel(AcqParams p, const float *__restrict__ lambda_g, const float2 *__restrict__ gamma,
                                                            AcqState *st) {
  extern __shared__ float s_lambda[];          // [N] lambda, then [N] alpha*v -> average before element i
  float *s_av = s_lambda + p.N;
  __shared__ unsigned s_rise[256], s_fall[256];  // N <= 8192
  __shared__ float s_avg_end;
  const int N = p.N;
  const float one_minus = 1.0f - p.alpha;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float v = lambda_g[i];
    s_lambda[i] = v;
    s_av[i] = __fmul_rn(p.alpha, v);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float avg = st->avg;
    float4 *av4 = reinterpret_cast<float4 *>(s_av);
    for (int i = 0; i < N / 4; i++) {
      float4 a = av4[i], o;
      o.x = avg; avg = __fadd_rn(a.x, __fmul_rn(one_minus, avg));
      o.y = avg; avg = __fadd_rn(a.y, __fmul_rn(one_minus, avg));
      o.z = avg; avg = __fadd_rn(a.z, __fmul_rn(one_minus, avg));
      o.w = avg; avg = __fadd_rn(a.w, __fmul_rn(one_minus, avg));
      av4[i] = o;
    }
    s_avg_end = avg;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {   // N is a multiple of 32: whole warps
    float v = s_lambda[i], av = s_av[i];
    unsigned r = __ballot_sync(0xffffffffu, v > __fmul_rn(av, p.rise));
    unsigned f = __ballot_sync(0xffffffffu, v > __fmul_rn(av, p.fall));
    if ((threadIdx.x & 31) == 0) { s_rise[i >> 5] = r; s_fall[i >> 5] = f; }
  }
  __syncthreads();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float *lambda = s_lambda;
  int n = 0, best = 0;
  {
    int i = 0, state = 0, peak_index = 0;
    float peak_val = -INFINITY, best_val = 0.f;
    while (i < N) {
      if (state == 0) {
        // next element at or after i that exceeds avg*rise; the elements in between only advance
        int w = i >> 5;
        unsigned m = s_rise[w] & (0xffffffffu << (i & 31));
        while (m == 0u && ++w < N / 32) m = s_rise[w];
        if (m == 0u) break;
        i = 32 * w + __ffs(m) - 1;
        state = 1;
      } else {
        float v = lambda[i];
        if (v > peak_val) { peak_val = v; peak_index = i; i++; }
        else if ((s_fall[i >> 5] >> (i & 31)) & 1u) { i++; }
        else {
          if (n == 0 || lambda[peak_index] > best_val) { best_val = lambda[peak_index]; best = peak_index; }
          n++;
          state = 0;
          peak_val = -INFINITY;
        }
      }
    }
  }
  float avg = s_avg_end;
  st->avg = avg;
  st->initial = n;
  // phase loop of this ml_sync call (:285-309) happens whether or not a peak was found
  double ph = st->phase;
  double inc = st->phaseinc;
  int total = p.N + p.cp;
  int sw = st->nextpos;
  if (n > 0) {
    if (sw >= 0 && sw < total) { ph += sw * inc; inc = st->nextphaseinc; ph += (total - sw) * inc; }
    else ph += total * inc;
  } else {
    ph += total * inc;
  }
  ph = remainder(ph, 2.0 * M_PI);
  st->phase = (float)ph;
  st->phaseinc = inc;
  if (n > 0) {
    int peak = best + (p.N + p.cp - 1);
    st->cp_start = peak;
    float eps = atan2f(gamma[best].y, gamma[best].x);
    st->nextphaseinc = (-1.0 / (double)p.N) * (double)eps;
    st->nextpos = peak - (p.cp + p.N);
  }
}

// ---- look-ahead after an initial acquisition.  A peak found by the N-candidate search is sometimes lost again one or
// two symbols later (the detector's average still carries the search), and the reference then restarts acquisition
// half a symbol further on (:545-558).  The tracking tables of a whole capture would be computed for nothing, so this
// kernel follows the first kProbe symbols the way the reference does (same lambda values, same detector) and tells the
// host where the first miss is; the host then sizes the first tracking batch to end there.  Advisory only: the batch
// kernels decide what is output, whatever this kernel says.
constexpr int kProbe = 4;
__global__ void __launch_bounds__(kProbe * kCand) acq_probe_kernel(AcqParams p, const float2 *__restrict__ x, long long base, long long nsamples,
                                                                   AcqState *st, long long lo) {
  __shared__ float s_l[kProbe * kCand];
  const int t = threadIdx.x, n = t / kCand, c = t % kCand;
  if (!st->initial) { if (t == 0) st->probe_lost1 = 0; return; }
  const int c0 = st->cp_start, total = p.N + p.cp;
  long long q = base + (long long)n * total + c0 - kD + c;
  float lam = -INFINITY;
  float2 g;
  if (q < nsamples) ml_point(x, q, p.N, p.cp, p.rho2, &lam, &g, lo);
  s_l[t] = lam;
  __syncthreads();
  if (t != 0) return;
  float avg = st->avg;
  int off = kD - 8, lost1 = 0;
  for (int m = 0; m < kProbe; m++) {
    if (off < 0 || off > kCand - 16) break;
    if (base + (long long)m * total + c0 - kD + off + 15 >= nsamples) break;   // beyond the input: the batch stops there as well
    int best;
    int np = peak_detect(s_l + m * kCand + off, 16, &avg, p.rise, p.fall, p.alpha, &best);
    if (np <= 0) { lost1 = m + 1; break; }
    off += best - 8;
  }
  st->probe_lost1 = lost1;
}

// ---- tracking table: symbol n, candidate c <-> symbol end at base + n*(N+cp) + c0 - kD + c
__global__ void acq_lambda_kernel(AcqParams p, const float2 *__restrict__ x, long long base, int c0, int nsym,
                                  float *__restrict__ lambda, float2 *__restrict__ gamma, long long lo) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nsym * kCand) return;
  int n = (int)(t / kCand), c = (int)(t % kCand);
  ml_point(x, base + (long long)n * (p.N + p.cp) + c0 - kD + c, p.N, p.cp, p.rho2, &lambda[t], &gamma[t], lo);
}

// ---- tracking as a finite-state machine, evaluated in parallel -----------------------------------
// The reference state that crosses symbols is (cp_start, d_avg) plus the phase schedule.  cp_start
// moves by (best - 8) per symbol; the table covers window offsets c = 0..kNC-1 (window = candidates
// [c, c+16), cp_start = c0 - kD + 8 + c).  d_avg after the 16 values of a window is, to float
// precision, independent of the average it started from, so it is speculated and then verified:
//   pass 1: avg1[n][c] = average after window (n, c) starting from 0
//   pass 2: for state s = (c, d), d = previous offset - c in [-kHD, kHD]: run the detector on window (n, c)
//           starting from avg1[n-1][c+d]; best[n][s], avg2[n][s], and the successor state
//           next[n][s] = (c + best - 8, 8 - best) — or a stop code when the peak is missed, the next window
//           leaves the table, or the speculation cannot be continued (|move| > kHD, or avg2 != avg1[n][c]
//           bit-for-bit, i.e. the successor's assumed input would not be the true average).
// The true trajectory is then the composition next[n-1] o ... o next[0] applied to the start state:
// acq_compose_kernel composes 32-symbol chunks for all kNS start states in parallel, chains the chunk
// maps, re-walks every chunk from its now known start state, and acq_finish_kernel turns the per-symbol
// (offset, best) into output descriptors with a warp scan of the phase schedule (:285-312).  A stop code
// ends the batch early; the host loop continues from there with the true state (exactness never rests
// on the speculation).
constexpr int kNC = kCand - 16 + 1;  // 17 window offsets
constexpr int kHD = 5;               // |previous offset - offset| covered by the tables (peaks 3..13 of 0..15: noise-free
                                     // QAM64 captures already jitter by +-5 through the resampler, histograms in profiles/README.md)
constexpr int kND = 2 * kHD + 1;     // previous-offset deltas -5..5
constexpr int kNS = kNC * kND;       // 187 states (< kStop)
constexpr unsigned char kLost = 0xFE, kOff = 0xFD, kSplit = 0xFC, kStop = 0xF0;
constexpr int kChunk = 32;
static_assert(kNS < kStop, "state ids and stop codes share one byte");

// The running average does not depend on the detector's state (:88-118), so for a window and a start value it is a
// straight-line float recurrence, and the detector only sees it through the 16 + 16 threshold tests "v > avg * rise",
// "v > avg * fall".  Pass 1 needs the recurrence alone; pass 2 replays the state machine on the two bit masks (same
// visits per element as peak_detect() above, a third of its instructions: ncu had pass 2 issue bound at 700
// instructions per thread).
__global__ void acq_pass1_kernel(AcqParams p, int nsym, const float *__restrict__ lambda, float *__restrict__ avg1) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsym * kNC) return;
  int n = t / kNC, c = t - n * kNC;
  const float *d = lambda + (long long)n * kCand + c;
  const float one_minus = 1.0f - p.alpha;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; i++) v[i] = d[i];
  float avg = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) avg = __fadd_rn(__fmul_rn(p.alpha, v[i]), __fmul_rn(one_minus, avg));
  avg1[t] = avg;
}

__global__ void __launch_bounds__(128) acq_pass2_kernel(AcqParams p, int nsym, const float *__restrict__ lambda, const float *__restrict__ avg1,
                                                        float avg_first, signed char *__restrict__ best2, float *__restrict__ avg2,
                                                        unsigned char *__restrict__ next) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsym * kNS) return;
  int n = t / kNS, s = t - n * kNS, c = s / kND, d = s - c * kND - kHD;
  int cp = c + d;
  signed char res = -2;
  float avg = 0.f;
  unsigned char nx = kSplit;
  if (n == 0 || (cp >= 0 && cp < kNC)) {
    avg = n == 0 ? avg_first : avg1[(n - 1) * kNC + cp];
    const float *w = lambda + (long long)n * kCand + c;
    const float one_minus = 1.0f - p.alpha;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = w[i];
    unsigned rise = 0, fall = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      rise |= (v[i] > __fmul_rn(avg, p.rise) ? 1u : 0u) << i;
      fall |= (v[i] > __fmul_rn(avg, p.fall) ? 1u : 0u) << i;
      avg = __fadd_rn(__fmul_rn(p.alpha, v[i]), __fmul_rn(one_minus, avg));
    }
    int state = 0, npeaks = 0, peak_index = 0, best = 0;
    float peak_val = -INFINITY, peak_at = 0.f, best_val = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const float vi = v[i];
      for (;;) {
        if (state == 0) {
          if ((rise >> i) & 1u) { state = 1; continue; }
          break;
        }
        if (vi > peak_val) { peak_val = vi; peak_at = vi; peak_index = i; break; }
        if ((fall >> i) & 1u) break;
        // record the peak; the first strictly greatest recorded peak wins (:127-137)
        if (npeaks == 0 || peak_at > best_val) { best_val = peak_at; best = peak_index; }
        npeaks++;
        state = 0;
        peak_val = -INFINITY;
      }
    }
    if (npeaks <= 0) { res = -1; nx = kLost; }
    else {
      res = (signed char)best;
      int cn = c + best - 8, dn = 8 - best;
      if (cn < 0 || cn >= kNC) nx = kOff;
      else if (dn < -kHD || dn > kHD || __float_as_uint(avg) != __float_as_uint(avg1[n * kNC + c])) nx = kSplit;
      else nx = (unsigned char)(cn * kND + dn + kHD);
    }
  }
  best2[t] = res;
  avg2[t] = avg;
  next[t] = nx;
}

struct AcqWalk {
  int n_found;        // symbols that produced output
  int code;           // 0: all symbols processed, else kLost / kOff / kSplit
  float avg;          // true detector average after the last processed symbol (incl. a missed one)
  int n_override;     // symbols whose speculation did not hold and that were re-run sequentially
  int avg_from;       // symbol whose table entry holds the final average (acq_walk_kernel fetches it), or -1
  int n_seg, n_staged;            // table segments; chunks walked symbol by symbol (trace)
  long long cyc_maps, cyc_serial;
  long long cyc_fin[6];   // trace: phase boundaries of acq_finish_kernel
  long long cyc_cmp[4];   // trace: compose chain - staging, table walks, detector re-runs, everything else  // trace: clock64 deltas of the three phases (DVBT_B200_ACQ_TRACE)
};

// chunk maps: one warp per chunk of `per_thread` symbols, lane = start state (kNS states, 6 rounds).  The chunk's rows of
// `next` are staged in shared memory with one coalesced round trip when they fit (32-symbol chunks: 6 KB per warp); walking
// them in global memory is a chain of per_thread dependent L2 round trips per start state.
constexpr int kMapRows = kChunk;
__global__ void __launch_bounds__(128) acq_chunkmap_kernel(int nsym, int per_thread, int nchunks, const unsigned char *__restrict__ next,
                                                           unsigned char *__restrict__ maps) {
  __shared__ __align__(16) unsigned char s_rows[4][kMapRows * kNS + 32];
  int w = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (w >= nchunks) return;
  const int n0 = w * per_thread, n1 = min(nsym, n0 + per_thread);
  const bool staged = per_thread <= kMapRows;
  unsigned char *rows = s_rows[threadIdx.x >> 5];
  int shift = 0;
  if (staged) {
    // 16-byte units from the aligned address below the first row (the allocation carries slack at both ends)
    const long long off = (long long)n0 * kNS, a0 = off & ~15LL;
    shift = (int)(off - a0);
    const int n16 = (shift + (n1 - n0) * kNS + 15) >> 4;
    const uint4 *src = reinterpret_cast<const uint4 *>(next + a0);
    uint4 *dst = reinterpret_cast<uint4 *>(rows);
    for (int i = lane; i < n16; i += 32) dst[i] = src[i];
    __syncwarp();
  }
  for (int s0 = lane; s0 < kNS; s0 += 32) {
    unsigned char cur = (unsigned char)s0;
    if (staged) {
      for (int n = n0; n < n1; n++) {
        unsigned char nx = rows[shift + (n - n0) * kNS + cur];
        if (nx >= kStop) { cur = kStop; break; }
        cur = nx;
      }
    } else {
      for (int n = n0; n < n1; n++) {
        unsigned char nx = next[(long long)n * kNS + cur];
        if (nx >= kStop) { cur = kStop; break; }
        cur = nx;
      }
    }
    maps[(long long)w * kNS + s0] = cur;
  }
}

// Chains the chunk maps into a list of table segments (first symbol, end, start state) that
// acq_walk_kernel then re-walks in parallel, writing (offset c, best) per symbol.  Where the speculation
// stops (kSplit) the chain continues on the spot: the reference detector runs sequentially on the next
// symbol from the true average, and keeps doing so until a symbol's result re-validates the tables; those
// few symbols are written here.
//
// The chain is sequential, so it runs on warp 0 with warp-uniform control flow: every lane executes the
// same scalar logic, the lanes cooperate only to fetch table rows / lambda windows with one coalesced
// round trip (16-byte loads), and lane 0 alone writes results.
constexpr int kMaxSeg = 2048;
constexpr int kRowsCap = kChunk;

// 16-byte-granular copy of `bytes` bytes starting at byte offset `off` of `src` into shared memory;
// returns the shift to add to indices into `dst` (the copy starts at the aligned address below `off`).
// The DevBuf allocations are 256-byte aligned and carry >= 256 bytes of slack, so reading up to 15 bytes
// either side of the range stays inside the allocation.
__device__ __forceinline__ int stage_bytes16(const unsigned char *__restrict__ src, long long off, int bytes, unsigned char *dst, int lane,
                                             int nlanes) {
  long long a0 = off & ~15LL;
  int shift = (int)(off - a0);
  int n16 = (shift + bytes + 15) >> 4;
  const uint4 *s4 = reinterpret_cast<const uint4 *>(src + a0);
  uint4 *d4 = reinterpret_cast<uint4 *>(dst);
  // eight loads in flight per lane: a load-then-store loop would wait one full memory latency per 16 bytes
  for (int i0 = lane; i0 < n16; i0 += 8 * nlanes) {
    uint4 v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      int i = i0 + u * nlanes;
      v[u] = i < n16 ? s4[i] : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      int i = i0 + u * nlanes;
      if (i < n16) d4[i] = v[u];
    }
  }
  return shift;
}

__global__ void __launch_bounds__(256) acq_compose_kernel(AcqParams p, int nsym, int per_thread, int nchunks, int start_state, float avg_first,
                                                          const float *__restrict__ lambda, const float *__restrict__ avg1,
                                                          const signed char *__restrict__ best2, const float *__restrict__ avg2,
                                                          const unsigned char *__restrict__ next, const unsigned char *__restrict__ maps,
                                                          unsigned char *__restrict__ c_of, signed char *__restrict__ best_of, int4 *__restrict__ segs,
                                                          AcqWalk *walk) {
  extern __shared__ __align__(16) unsigned char s_maps[];  // [nchunks][kNS] (+ slack)
  __shared__ __align__(16) unsigned char s_rows[kRowsCap * kNS + 32];  // `next` rows of the chunk being walked symbol by symbol
  // everything else a speculation split can ask for, staged together with the rows when the chunk is short
  // (one round trip to L2 instead of four dependent ones): best2 / avg2 / lambda / avg1 rows of the chunk
  constexpr int kExt = kChunk;
  __shared__ __align__(16) unsigned char s_best[kExt * kNS + 32];
  __shared__ __align__(16) unsigned char s_avg2[kExt * kNS * 4 + 32];
  __shared__ __align__(16) unsigned char s_lam[kExt * kCand * 4 + 32];
  __shared__ __align__(16) unsigned char s_avg1[kExt * kNC * 4 + 32];
  __shared__ float s_win[16];
  __shared__ unsigned char s_cst[1024];   // start state of every chunk taken whole
  const int t = threadIdx.x;
  const long long cyc0 = clock64();
  stage_bytes16(maps, 0, nchunks * kNS, s_maps, t, blockDim.x);
  __syncthreads();
  const long long cyc1 = clock64();
  // Warps 1..7 serve warp 0: when the chain needs table rows they are fetched by all 256 threads (one round trip
  // to L2 for ~20 KB instead of one per 4 KB).  Protocol: warp 0 posts (first symbol, end) and everybody meets at
  // a block barrier, copies, meets again; end < 0 dismisses the helpers.
  __shared__ int s_cmd_n0, s_cmd_n1;
  // all five tables in ONE batch of loads: the 16-byte units of the five source ranges are numbered consecutively
  // and dealt to the threads, every thread issues its (<= 8) loads before the first store
  auto stage_chunk = [&](int n0, int n1) {
    const int r = n1 - n0;
    const unsigned char *srcs[5] = {next, reinterpret_cast<const unsigned char *>(best2), reinterpret_cast<const unsigned char *>(avg2),
                                    reinterpret_cast<const unsigned char *>(lambda), reinterpret_cast<const unsigned char *>(avg1)};
    unsigned char *dsts[5] = {s_rows, s_best, s_avg2, s_lam, s_avg1};
    const int rowb[5] = {kNS, kNS, kNS * 4, kCand * 4, kNC * 4};
    const int ntab = r <= kExt ? 5 : 1;
    long long a0[5];
    int first[6];
    first[0] = 0;
#pragma unroll
    for (int q = 0; q < 5; q++) {
      long long off = (long long)n0 * rowb[q];
      a0[q] = off & ~15LL;
      int n16 = q < ntab ? (int)((off - a0[q]) + (long long)r * rowb[q] + 15) >> 4 : 0;
      first[q + 1] = first[q] + n16;
    }
    const int total = first[5];
    for (int u0 = t; u0 < total; u0 += 8 * blockDim.x) {
      uint4 v[8];
      int tab[8], idx[8];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        int u = u0 + k * blockDim.x;
        int q = 0;
#pragma unroll
        for (int qq = 1; qq < 5; qq++) q += (u >= first[qq]) ? 1 : 0;
        tab[k] = q; idx[k] = u - first[q];
        v[k] = u < total ? reinterpret_cast<const uint4 *>(srcs[q] + a0[q])[idx[k]] : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 8; k++) {
        int u = u0 + k * blockDim.x;
        if (u < total) reinterpret_cast<uint4 *>(dsts[tab[k]])[idx[k]] = v[k];
      }
    }
  };
  if (t >= 32) {
    for (;;) {
      __syncthreads();
      const int n1 = s_cmd_n1;
      if (n1 < 0) return;
      stage_chunk(s_cmd_n0, n1);
      __syncthreads();
    }
  }
  const int lane = t;
  int n = 0, code = 0, n_found = 0, n_override = 0, nseg = 0, avg_from = -1, n_staged = 0;
  int st_n0 = 0, st_n1 = 0, st_shift = 0;   // symbols whose rows are in s_rows
  int ex_n0 = 0, ex_n1 = 0, sh_best = 0, sh_avg2 = 0, sh_lam = 0, sh_avg1 = 0;   // symbols covered by the extended staging
  auto ld_best2 = [&](int nn, int state) -> int {
    return (nn >= ex_n0 && nn < ex_n1) ? (int)(signed char)s_best[sh_best + (nn - ex_n0) * kNS + state] : (int)best2[(long long)nn * kNS + state];
  };
  auto ld_avg2 = [&](int nn, int state) -> float {
    return (nn >= ex_n0 && nn < ex_n1) ? *reinterpret_cast<const float *>(s_avg2 + sh_avg2 + ((nn - ex_n0) * kNS + state) * 4)
                                       : avg2[(long long)nn * kNS + state];
  };
  auto ld_lambda = [&](int nn, int cand) -> float {
    return (nn >= ex_n0 && nn < ex_n1) ? *reinterpret_cast<const float *>(s_lam + sh_lam + ((nn - ex_n0) * kCand + cand) * 4)
                                       : lambda[(long long)nn * kCand + cand];
  };
  auto ld_avg1 = [&](int nn, int c) -> float {
    return (nn >= ex_n0 && nn < ex_n1) ? *reinterpret_cast<const float *>(s_avg1 + sh_avg1 + ((nn - ex_n0) * kNC + c) * 4)
                                       : avg1[(long long)nn * kNC + c];
  };
  unsigned char st = (unsigned char)start_state;
  float avg = avg_first;
  int k = 0, kstart = 0;                    // chunk of symbol n and its first symbol
  long long c_stage = 0, c_walk = 0, c_over = 0;
  while (n < nsym && !code) {
    if (n == kstart) {
      // whole chunks by their maps, as far as they go: the loop carries only the state (one shared-memory byte
      // per chunk in, one out); the segments of the run are written afterwards, one per lane
      const int k0 = k, room = kMaxSeg - 2 - nseg;
      while (kstart < nsym && k - k0 < room) {
        unsigned char m = s_maps[k * kNS + st];
        if (m >= kStop) break;
        s_cst[k] = st;
        st = m;
        k++;
        kstart += per_thread;
      }
      if (k > k0) {
        __syncwarp();
        for (int kk = k0 + lane; kk < k; kk += 32)
          segs[nseg + kk - k0] = make_int4(kk * per_thread, min(nsym, (kk + 1) * per_thread), s_cst[kk], 0);
        __syncwarp();
        nseg += k - k0;
        n = min(nsym, kstart);
        n_found = n; avg_from = n - 1;
        continue;
      }
    }
    int nend = min(nsym, kstart + per_thread);
    if (nseg >= kMaxSeg - 2) { code = kSplit; break; }  // segment list full: end the batch here (the host loop continues)
    // table walk inside chunk k until its end or a stop code; the rows are staged in shared memory first
    if (nend - n > kRowsCap) nend = n + kRowsCap;
    if (n < st_n0 || nend > st_n1) {
      const long long cs0 = clock64();
      if (lane == 0) { s_cmd_n0 = n; s_cmd_n1 = nend; }
      __syncthreads();
      stage_chunk(n, nend);
      __syncthreads();
      c_stage += clock64() - cs0;
      auto shift_of = [](long long off) { return (int)(off & 15LL); };
      st_shift = shift_of((long long)n * kNS);
      st_n0 = n; st_n1 = nend;
      if (nend - n <= kExt) {
        sh_best = shift_of((long long)n * kNS);
        sh_avg2 = shift_of((long long)n * kNS * 4);
        sh_lam = shift_of((long long)n * kCand * 4);
        sh_avg1 = shift_of((long long)n * kNC * 4);
        ex_n0 = n; ex_n1 = nend;
      } else {
        ex_n0 = ex_n1 = 0;
      }
      n_staged++;
      __syncwarp();
    }
    const unsigned char *rows = s_rows + st_shift - st_n0 * kNS;   // rows[n * kNS + state]
    int seg0 = n;
    unsigned char seg_st = st, nx = 0, st_at = st;
    const long long cw0 = clock64();
    while (n < nend) {
      st_at = st;
      nx = rows[n * kNS + st];
      if (nx >= kStop) break;
      st = nx;
      n++;
    }
    c_walk += clock64() - cw0;
    if (n == nend) {  // reached the end of the staged rows without a stop
      if (lane == 0) segs[nseg] = make_int4(seg0, n, seg_st, 0);
      nseg++;
      n_found = n; avg_from = n - 1;
      if (n == kstart + per_thread) { k++; kstart += per_thread; }
      continue;
    }
    // stop code at symbol n, reached in state st_at
    if (nx == kLost) {
      if (n > seg0) { if (lane == 0) segs[nseg] = make_int4(seg0, n, seg_st, 0); nseg++; }
      avg = ld_avg2(n, st_at);  // the missed symbol still updated the average
      avg_from = -1;
      n_found = n;
      code = kLost;
      break;
    }
    // kOff / kSplit: symbol n itself is good (table entry valid)
    if (lane == 0) segs[nseg] = make_int4(seg0, n + 1, seg_st, 0);
    nseg++;
    int c = st_at / kND;
    int best = ld_best2(n, st_at);
    avg = ld_avg2(n, st_at);
    avg_from = -1;
    n++;
    n_found = n;
    if (nx == kOff) { code = kOff; break; }
    // kSplit: the tables cannot be trusted for the next symbol; run the detector from the true average
    int cn = c + best - 8;
    const long long co0 = clock64();
    while (n < nsym && !code) {
      if (cn < 0 || cn >= kNC) { code = kOff; break; }
      __syncwarp();
      if (lane < 16) s_win[lane] = ld_lambda(n, cn + lane);
      __syncwarp();
      int b2;
      float a2 = avg;
      int np = peak_detect(s_win, 16, &a2, p.rise, p.fall, p.alpha, &b2);
      n_override++;
      if (lane == 0) c_of[n] = (unsigned char)cn;
      avg = a2;
      if (np <= 0) { if (lane == 0) best_of[n] = -1; code = kLost; break; }
      if (lane == 0) best_of[n] = (signed char)b2;
      n++;
      n_found = n;
      int c2 = cn + b2 - 8, d2 = 8 - b2;
      if (c2 < 0 || c2 >= kNC) { code = kOff; break; }
      bool ok = d2 >= -kHD && d2 <= kHD && __float_as_uint(avg) == __float_as_uint(ld_avg1(n - 1, cn));
      if (ok) { st = (unsigned char)(c2 * kND + d2 + kHD); break; }   // tables valid again from symbol n
      cn = c2;
    }
    c_over += clock64() - co0;
    k = n / per_thread;
    kstart = k * per_thread;
  }
  if (lane == 0) {
    walk->code = code;
    walk->n_found = n_found;
    walk->n_override = n_override;
    walk->avg = avg;
    walk->avg_from = avg_from;
    walk->n_seg = nseg;
    walk->n_staged = n_staged;
    walk->cyc_maps = cyc1 - cyc0;
    walk->cyc_serial = clock64() - cyc1;
    walk->cyc_cmp[0] = c_stage; walk->cyc_cmp[1] = c_walk; walk->cyc_cmp[2] = c_over;
    walk->cyc_cmp[3] = walk->cyc_serial - c_stage - c_walk - c_over;
    s_cmd_n1 = -1;   // dismiss the helper warps
  }
  __syncthreads();
}

// One warp per table segment: the segment's `next` rows are staged with one coalesced round trip, lane 0
// walks the states through shared memory, then lane i writes (offset, best) of the segment's i-th symbol.
__global__ void __launch_bounds__(128) acq_walk_kernel(int rows_cap, const signed char *__restrict__ best2, const float *__restrict__ avg2,
                                                       const unsigned char *__restrict__ next, const int4 *__restrict__ segs,
                                                       unsigned char *__restrict__ c_of, signed char *__restrict__ best_of, AcqWalk *walk) {
  extern __shared__ __align__(16) unsigned char s_walk[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seg = blockIdx.x * (blockDim.x >> 5) + w;
  if (seg >= walk->n_seg) return;
  const int per_warp = rows_cap * kNS + 32 + rows_cap;          // rows (16-byte granular) + one state per symbol
  unsigned char *rows = s_walk + (size_t)w * ((per_warp + 15) & ~15);
  unsigned char *states = rows + rows_cap * kNS + 32;
  const int4 sg = segs[seg];
  const int n0 = sg.x, cnt = sg.y - sg.x;
  int shift = stage_bytes16(next, (long long)n0 * kNS, cnt * kNS, rows, lane, 32);
  __syncwarp();
  if (lane == 0) {
    unsigned char st = (unsigned char)sg.z;
    for (int i = 0; i < cnt; i++) {
      states[i] = st;
      st = rows[shift + i * kNS + st];
    }
  }
  __syncwarp();
  const int avg_from = walk->avg_from;
  for (int i = lane; i < cnt; i += 32) {
    int n = n0 + i;
    unsigned char st = states[i];
    c_of[n] = (unsigned char)(st / kND);
    best_of[n] = best2[(long long)n * kNS + st];
    if (n == avg_from) walk->avg = avg2[(long long)n * kNS + st];
  }
}

// per-symbol quantities that need no ordering: peak position and the carrier-offset estimate it leaves behind
__global__ void acq_post_kernel(AcqParams p, int c0, const float2 *__restrict__ gamma, const unsigned char *__restrict__ c_of,
                                const signed char *__restrict__ best_of, const AcqWalk *walk, int *__restrict__ peak_of,
                                float *__restrict__ eps_of) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= walk->n_found) return;
  int c = c_of[m], best = best_of[m];
  peak_of[m] = c0 - kD + c + best;                         // cp_start_before - 8 + best
  float2 g = gamma[(long long)m * kCand + c + best];
  eps_of[m] = atan2f(g.y, g.x);                            // d_nextphaseinc left by symbol m = -eps / N (:311)
}

// phase schedule (:285-312): one block.  (peak, eps) of all symbols are staged in shared memory with coalesced
// loads (kFinishMax symbols at most: the host splits longer batches).  Every thread owns a run of consecutive
// symbols; two small scans over the per-thread summaries give (a) the increment in force when a run starts
// ("value switched to by the last earlier symbol that switched", :287-288) and (b) the phase at the start of the
// run.  The output descriptors are then written one symbol per thread (coalesced), each thread re-adding its
// run's increments from the run start in the same order as the run owner did.
constexpr int kFinishMax = 24576;

__global__ void __launch_bounds__(1024) acq_finish_kernel(AcqParams p, long long base, const int *__restrict__ peak_of,
                                                          const float *__restrict__ eps_of, const AcqWalk *walk, AcqState *st,
                                                          double *__restrict__ run_start) {
  extern __shared__ __align__(16) unsigned char s_fin[];   // [nf] int peak, [nf] float eps
  __shared__ double s_val[1024], s_sum[1024];
  __shared__ unsigned char s_has[1024];
  const int t = threadIdx.x, nt = blockDim.x;
  const int total = p.N + p.cp;
  const double twopi = 2.0 * M_PI, minus_inv_n = -1.0 / (double)p.N;
  long long *trace = const_cast<AcqWalk *>(walk)->cyc_fin;
  const long long tr0 = clock64();
  const int nf = walk->n_found;
  int *s_pk = reinterpret_cast<int *>(s_fin);
  float *s_ep = reinterpret_cast<float *>(s_fin) + nf;
  for (int m0 = t; m0 < nf; m0 += 8 * nt) {
    int pk[8];
    float ep[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      int m = m0 + u * nt;
      pk[u] = m < nf ? peak_of[m] : 0;
      ep[u] = m < nf ? eps_of[m] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      int m = m0 + u * nt;
      if (m < nf) { s_pk[m] = pk[u]; s_ep[m] = ep[u]; }
    }
  }
  const int per = (nf + nt - 1) / nt;
  const int a = min(nf, t * per), b = min(nf, a + per);
  const double inc_init = st->phaseinc, pend_init = st->nextphaseinc;
  const int nextpos_init = st->nextpos;
  __syncthreads();
  // symbol m switches the increment at sample swm (if inside the symbol) to pend
  auto swm_of = [&](int m) { return m == 0 ? nextpos_init : s_pk[m - 1] - total; };
  auto pend_of = [&](int m) { return m == 0 ? pend_init : minus_inv_n * (double)s_ep[m - 1]; };
  // (a) does a symbol of my run switch the increment, and to what
  bool has = false;
  double val = 0.0;
  for (int m = a; m < b; m++) {
    int swm = swm_of(m);
    if (swm >= 0 && swm < total) { has = true; val = pend_of(m); }
  }
  s_has[t] = has; s_val[t] = val;
  __syncthreads();
  if (t == 0) trace[0] = clock64() - tr0;
  if (t == 0) {  // exclusive "last switched value" scan
    // serial on purpose (the order defines the result); unrolled in groups of 16 so that the shared-memory
    // loads are issued ahead of the dependent chain
    double cur = inc_init;
    for (int i0 = 0; i0 < nt; i0 += 16) {
      double v[16];
      unsigned char hs[16];
#pragma unroll
      for (int u = 0; u < 16; u++) { v[u] = s_val[i0 + u]; hs[u] = s_has[i0 + u]; }
#pragma unroll
      for (int u = 0; u < 16; u++) {
        s_val[i0 + u] = cur;
        if (hs[u]) cur = v[u];
      }
    }
    s_sum[0] = cur;  // increment in force after the last symbol (stashed, re-read below)
  }
  __syncthreads();
  const double inc_end = s_sum[0];
  __syncthreads();
  if (t == 0) trace[1] = clock64() - tr0;
  // (b) phase advance of my run
  double inc = s_val[t], sum = 0.0;
  for (int m = a; m < b; m++) {
    int swm = swm_of(m);
    double pendm = pend_of(m);
    bool ok = swm >= 0 && swm < total;
    sum += ok ? swm * inc + (total - swm) * pendm : total * inc;
    if (ok) inc = pendm;
  }
  s_sum[t] = sum;
  __syncthreads();
  if (t == 0) trace[2] = clock64() - tr0;
  if (t == 0) {
    double cur = st->phase;
    for (int i0 = 0; i0 < nt; i0 += 16) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; u++) v[u] = s_sum[i0 + u];
#pragma unroll
      for (int u = 0; u < 16; u++) { s_sum[i0 + u] = cur; cur += v[u]; }
    }
    // end state
    int code = walk->code;
    double ph = cur;
    if (code == kLost) ph += total * inc_end;  // the missed symbol still advances the phase (:335-343)
    ph -= twopi * rint(ph / twopi);
    st->avg = walk->avg;
    st->phase = (float)ph;
    st->phaseinc = inc_end;
    st->nextphaseinc = nf > 0 ? minus_inv_n * (double)s_ep[nf - 1] : pend_init;
    st->nextpos = nf > 0 ? s_pk[nf - 1] - total : nextpos_init;
    if (nf > 0) st->cp_start = s_pk[nf - 1];
    st->n_out = nf;
    st->lost_at = code == kLost ? nf : (code ? -2 - nf : -1);   // -2-nf: stopped after nf symbols without a miss
    st->fallback = code == kSplit ? 1 : 0;
    st->n_run += nf;
    st->n_single += code ? 1 : 0;
    st->n_seq += walk->n_override;
    st->consumed = (long long)nf * total;
    trace[3] = clock64() - tr0;
  }
  __syncthreads();
  // run starts for acq_desc_kernel: increment in force and phase when run t begins
  run_start[2 * t] = s_val[t];
  run_start[2 * t + 1] = s_sum[t];
  if (t == 0) trace[4] = clock64() - tr0;
}

// (c) output descriptors, one symbol per thread over the whole GPU (the double-precision work of this step on a
// single SM was the longest part of the phase schedule): each thread re-adds its run's increments from the run
// start in the same order as the run owner did in acq_finish_kernel.
__global__ void __launch_bounds__(128) acq_desc_kernel(AcqParams p, long long base, const int *__restrict__ peak_of,
                                                       const float *__restrict__ eps_of, const AcqWalk *walk, int nextpos_init,
                                                       double pend_init, const double *__restrict__ run_start, SymOut *__restrict__ out) {
  const int nf = walk->n_found;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nf) return;
  const int total = p.N + p.cp;
  const double twopi = 2.0 * M_PI, minus_inv_n = -1.0 / (double)p.N;
  const int per = (nf + 1023) / 1024;
  const int r = m / per;
  auto swm_of = [&](int mm) { return mm == 0 ? nextpos_init : peak_of[mm - 1] - total; };
  auto pend_of = [&](int mm) { return mm == 0 ? pend_init : minus_inv_n * (double)eps_of[mm - 1]; };
  double inc_m = run_start[2 * r], ph = run_start[2 * r + 1];
  for (int mm = r * per; mm < m; mm++) {
    int swm = swm_of(mm);
    double pendm = pend_of(mm);
    bool ok = swm >= 0 && swm < total;
    ph += ok ? swm * inc_m + (total - swm) * pendm : total * inc_m;
    if (ok) inc_m = pendm;
  }
  int swm = swm_of(m);
  bool ok = swm >= 0 && swm < total;
  SymOut so;
  so.first = base + (long long)m * total + peak_of[m] - p.N + 1;
  so.phase0 = ph - twopi * rint(ph / twopi);
  so.inc0 = inc_m; so.inc1 = pend_of(m); so.switch_at = ok ? swm : total;
  out[m] = so;
}

// out[n][j] = (-1)^j * expj(phase_j) * x[first + j].  A thread takes four samples 256 apart: four independent
// loads in flight per thread (one sample per thread left the kernel waiting on HBM latency at 2.8 TB/s).
constexpr int kDerotPer = 4;
__global__ void __launch_bounds__(256) acq_derot_kernel(int N, int nsym, const float2 *__restrict__ x, const SymOut *__restrict__ so,
                                                        float2 *__restrict__ out, int shift_sign) {
  int n = blockIdx.y;
  int j0 = blockIdx.x * (256 * kDerotPer) + threadIdx.x;
  if (n >= nsym) return;
  SymOut s = so[n];
  float2 v[kDerotPer];
#pragma unroll
  for (int u = 0; u < kDerotPer; u++) {
    int j = j0 + 256 * u;
    v[u] = j < N ? __ldg(x + s.first + j) : make_float2(0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < kDerotPer; u++) {
    int j = j0 + 256 * u;
    if (j >= N) break;
    int steps = j + 1;  // the phase is incremented before it is used (:291-307)
    double ph = s.phase0 + (steps <= s.switch_at ? steps * s.inc0 : s.switch_at * s.inc0 + (steps - s.switch_at) * s.inc1);
    ph -= (2.0 * M_PI) * rint(ph * (1.0 / (2.0 * M_PI)));
    float sn, cs;
    sincosf((float)ph, &sn, &cs);
    float2 r = cmulf(make_float2(cs, sn), v[u]);
    if (shift_sign && (j & 1)) r = make_float2(-r.x, -r.y);
    out[(long long)n * N + j] = r;
  }
}

// ---- derotation fused with the forward FFT (one block per OFDM symbol) ------------------------------------------
// Replaces acq_derot_kernel + cufftExecC2C when the block delivers frequency-domain symbols: the derotated
// samples never go to HBM (8N B read + 8N B written per symbol instead of 4 x 8N).  Stockham autosort passes
// through ONE shared-memory buffer, 16 points per thread and pass (N/16 threads): radices 16,16,8 for N = 2048 and
// 16,16,16,2 for N = 8192, i.e. two (three) shared-memory round trips.  Pass with radix R and p = product of the
// earlier radices, butterfly i (0 <= i < N/R): k = i mod p, inputs in[i + r N/R] * W_N^(r k N/(p R)), outputs to
// out[(i - k) R + k + q p].  The first pass reads global memory (derotation and (-1)^j applied on the fly; the 16
// inputs of a thread are N/16 apart = coalesced across the block), the last one writes global memory
// (q N/R + i: coalesced).  Index a of the buffer lives at a + a/16: with that padding every access pattern above
// is conflict free.  Twiddles: W_N^n from a table computed in double precision on the host; a butterfly loads the
// powers 1, 2, 4, 8 of its base twiddle and multiplies the others together (the first radix-8 version loaded all
// of them and kept L1 92 % busy, profiles/r01_side_kernels_v20_ncu_summary.txt).
__device__ __forceinline__ float2 cadd2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub2(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul2(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cmulmi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

__device__ __forceinline__ void dft2(float2 *u) {
  float2 a = u[0];
  u[0] = cadd2(a, u[1]); u[1] = csub2(a, u[1]);
}
__device__ __forceinline__ void dft4(float2 *u) {
  float2 a0 = cadd2(u[0], u[2]), a2 = csub2(u[0], u[2]), a1 = cadd2(u[1], u[3]), a3 = cmulmi(csub2(u[1], u[3]));
  u[0] = cadd2(a0, a1); u[1] = cadd2(a2, a3); u[2] = csub2(a0, a1); u[3] = csub2(a2, a3);
}
__device__ __forceinline__ void dft8(float2 *u) {
  const float h = 0.70710678118654752440f;
  float2 a0 = cadd2(u[0], u[4]), a4 = csub2(u[0], u[4]), a1 = cadd2(u[1], u[5]), a5 = csub2(u[1], u[5]);
  float2 a2 = cadd2(u[2], u[6]), a6 = csub2(u[2], u[6]), a3 = cadd2(u[3], u[7]), a7 = csub2(u[3], u[7]);
  a5 = make_float2(h * (a5.x + a5.y), h * (a5.y - a5.x));      // * W8^1 = (1 - i)/sqrt2
  a6 = cmulmi(a6);                                             // * W8^2 = -i
  a7 = make_float2(h * (a7.y - a7.x), -h * (a7.x + a7.y));     // * W8^3 = (-1 - i)/sqrt2
  float2 b0 = cadd2(a0, a2), b2 = csub2(a0, a2), b1 = cadd2(a1, a3), b3 = cmulmi(csub2(a1, a3));
  float2 b4 = cadd2(a4, a6), b6 = csub2(a4, a6), b5 = cadd2(a5, a7), b7 = cmulmi(csub2(a5, a7));
  u[0] = cadd2(b0, b1); u[4] = csub2(b0, b1); u[2] = cadd2(b2, b3); u[6] = csub2(b2, b3);
  u[1] = cadd2(b4, b5); u[5] = csub2(b4, b5); u[3] = cadd2(b6, b7); u[7] = csub2(b6, b7);
}
// 16 = 4 x 4: X[k1 + 4 k2] = sum_n2 W4^(n2 k2) W16^(n2 k1) sum_n1 W4^(n1 k1) x[4 n1 + n2]
__device__ __forceinline__ void dft16(float2 *u) {
  const float c1 = 0.92387953251128673848f, s1 = 0.38268343236508978178f, h = 0.70710678118654752440f;
  float2 y[4][4];   // [n2][k1]
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) {
    float2 a[4] = {u[n2], u[n2 + 4], u[n2 + 8], u[n2 + 12]};
    dft4(a);
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) y[n2][k1] = a[k1];
  }
  // W16^(n2 k1), W16^m = (cos, -sin)(2 pi m / 16)
  y[1][1] = cmul2(y[1][1], make_float2(c1, -s1));
  y[1][2] = make_float2(h * (y[1][2].x + y[1][2].y), h * (y[1][2].y - y[1][2].x));     // W16^2 = (1 - i)/sqrt2
  y[1][3] = cmul2(y[1][3], make_float2(s1, -c1));
  y[2][1] = make_float2(h * (y[2][1].x + y[2][1].y), h * (y[2][1].y - y[2][1].x));
  y[2][2] = cmulmi(y[2][2]);                                                            // W16^4 = -i
  y[2][3] = make_float2(h * (y[2][3].y - y[2][3].x), -h * (y[2][3].x + y[2][3].y));    // W16^6 = (-1 - i)/sqrt2
  y[3][1] = cmul2(y[3][1], make_float2(s1, -c1));
  y[3][2] = make_float2(h * (y[3][2].y - y[3][2].x), -h * (y[3][2].x + y[3][2].y));
  y[3][3] = cmul2(y[3][3], make_float2(-c1, s1));                                       // W16^9
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) {
    float2 a[4] = {y[0][k1], y[1][k1], y[2][k1], y[3][k1]};
    dft4(a);
#pragma unroll
    for (int k2 = 0; k2 < 4; k2++) u[k1 + 4 * k2] = a[k2];
  }
}
template <int R> __device__ __forceinline__ void dftR(float2 *u) {
  if (R == 16) dft16(u); else if (R == 8) dft8(u); else if (R == 4) dft4(u); else dft2(u);
}

__device__ __forceinline__ int fft_pad(int a) { return a + (a >> 4); }

// one Stockham pass over the shared buffer (or to global memory when LAST), radix R, p = product of the earlier radices
template <int N, int R, bool LAST>
__device__ __forceinline__ void fft16_pass(float2 *buf, float2 *__restrict__ gdst, int p, const float2 *__restrict__ W, int t) {
  constexpr int T = N / 16, NB = N / R, REPS = 16 / R;
  float2 u[REPS][R];
#pragma unroll
  for (int rep = 0; rep < REPS; rep++)
#pragma unroll
    for (int r = 0; r < R; r++) u[rep][r] = buf[fft_pad(t + rep * T + r * NB)];
  if (!LAST) __syncthreads();    // every input is in registers before the buffer is overwritten
#pragma unroll
  for (int rep = 0; rep < REPS; rep++) {
    const int i = t + rep * T;
    const int k = i & (p - 1);
    const int j = (i - k) * R + k;
    const int ws = k * (NB / p);             // W_N^(r ws), r = 1..R-1
    float2 w[R];
    w[1] = __ldg(W + (ws & (N - 1)));
    if (R >= 4) w[2] = __ldg(W + ((2 * ws) & (N - 1)));
    if (R >= 8) w[4] = __ldg(W + ((4 * ws) & (N - 1)));
    if (R >= 16) w[8] = __ldg(W + ((8 * ws) & (N - 1)));
    if (R >= 4) w[3] = cmul2(w[1], w[2]);
    if (R >= 8) { w[5] = cmul2(w[1], w[4]); w[6] = cmul2(w[2], w[4]); w[7] = cmul2(w[3], w[4]); }
    if (R >= 16) {
#pragma unroll
      for (int r = 1; r < 8; r++) w[8 + r] = cmul2(w[r], w[8]);
    }
#pragma unroll
    for (int r = 1; r < R; r++) u[rep][r] = cmul2(u[rep][r], w[r]);
    dftR<R>(u[rep]);
#pragma unroll
    for (int q = 0; q < R; q++) {
      if (LAST) gdst[j + q * p] = u[rep][q];
      else buf[fft_pad(j + q * p)] = u[rep][q];
    }
  }
  if (!LAST) __syncthreads();
}

template <int N>
__global__ void __launch_bounds__(N / 16) acq_fftd_kernel(int nsym, const float2 *__restrict__ x, const SymOut *__restrict__ so,
                                                          float2 *__restrict__ out, const float2 *__restrict__ W) {
  extern __shared__ __align__(16) float2 s_fft[];   // N + N/16
  constexpr int T = N / 16;
  const int n = blockIdx.x, t = threadIdx.x;
  if (n >= nsym) return;
  const SymOut s = so[n];
  // pass 0 (radix 16, p = 1) straight from global memory with the derotation
  {
    float2 u[16];
#pragma unroll
    for (int r = 0; r < 16; r++) u[r] = __ldg(x + s.first + t + r * T);
    // Derotation rotors.  The phase of sample j is linear in j on either side of switch_at (the phase is incremented
    // before it is used, :291-307): phase0 + (j+1) inc0, then phase0 + switch_at inc0 + (j+1-switch_at) inc1.  A thread's
    // 16 samples are T apart, so on each side its rotors are a geometric sequence: one sincos for the first sample and
    // one for the ratio exp(i T inc) per side (phases in double, reduced to (-pi, pi]), then 15 complex multiplies -
    // instead of 16 double-precision phases and 16 sincos, which were 35 % of this kernel's instructions
    // (profiles/r01_side_kernels_v36_ncu_summary.txt; the kernel is issue bound).  Rounding: <= 15 float multiplies
    // deep, ~1e-6 relative, inside the tolerance the closed-form phase already has against the reference's own
    // float accumulation (tests/test_acq_gpu.py).
    {
      auto rotor = [](double ph) {
        ph -= (2.0 * M_PI) * rint(ph * (1.0 / (2.0 * M_PI)));
        float sn, cs;
        sincosf((float)ph, &sn, &cs);
        return make_float2(cs, sn);
      };
      const int sw = s.switch_at;
      float2 ra = rotor(s.phase0 + (double)(t + 1) * s.inc0);
      float2 rb = rotor(s.phase0 + (double)sw * s.inc0 + (double)(t + 1 - sw) * s.inc1);
      const float2 sa = rotor((double)T * s.inc0), sb = rotor((double)T * s.inc1);
#pragma unroll
      for (int r = 0; r < 16; r++) {
        const int j = t + r * T;
        float2 v = cmul2(j + 1 <= sw ? ra : rb, u[r]);
        if (j & 1) v = make_float2(-v.x, -v.y);     // fft_vxx(shift = true)
        u[r] = v;
        ra = cmul2(ra, sa);
        rb = cmul2(rb, sb);
      }
    }
    dft16(u);
#pragma unroll
    for (int q = 0; q < 16; q++) s_fft[17 * t + q] = u[q];   // fft_pad(16 t + q)
    __syncthreads();
  }
  float2 *dst = out + (long long)n * N;
  if (N == 2048) {
    fft16_pass<N, 16, false>(s_fft, dst, 16, W, t);
    fft16_pass<N, 8, true>(s_fft, dst, 256, W, t);
  } else {
    fft16_pass<N, 16, false>(s_fft, dst, 16, W, t);
    fft16_pass<N, 16, false>(s_fft, dst, 256, W, t);
    fft16_pass<N, 2, true>(s_fft, dst, 4096, W, t);
  }
}

}  // namespace

struct dvbt_b200_acq {
  int device = dvbt::current_device();
  dvbt_b200_acq_params par;
  AcqParams kp;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cufftHandle plan = 0;
  int plan_batch = 0;
  dvbt::DevBuf d_hist;                      // acq_work: the last kAcqHistory samples consumed by the previous calls
  bool hist_valid = false;
  dvbt::DevBuf d_x, d_state, h_state, d_lambda, d_gamma, d_avg1, d_avg2, d_peak, d_sym, d_out, d_il, d_ig, d_eps, d_flag, d_maps, d_cof, d_bof, d_peakof, d_eof, d_seg, d_runs, d_tw;
  int tw_n = 0;
  static constexpr int kFftEv = 8;          // CUDA events around the derotation+FFT kernel of the first batches of a run
  cudaEvent_t ev_fft[2 * kFftEv] = {nullptr};
  int n_fft_ev = 0;
  bool pending_sync = false;                // a sync_start whose item the next acq_work call produces
  dvbt::Staging stg;                        // pinned staging of acq_work's pageable buffers
  bool state_is_zero = false;               // set by acq_reset(): acq_run need not read the state back
};

namespace dvbt {

// Runs acquisition + derotation (+ optional FFT) over device samples x[0..n).  Output symbols go to
// d_out (N complex each).  Returns counts through the host copy of the state.
// nhist: consumed samples of the stream that are still readable in front of x (x[-nhist .. -1]); what lies further back reads
// as zero (ml_point).
int acq_run(dvbt_b200_acq *h, const float2 *x, long long n, float2 *d_out, long long out_capacity_syms, int do_fft,
            AcqState *host_state_out, std::vector<long long> *sync_at = nullptr, int nhist = 0) {
  const AcqParams &p = h->kp;
  const int total = p.N + p.cp;
  cudaStream_t st = h->stream;
  AcqState *hs = h->h_state.as<AcqState>();
  long long pos = 0;       // read position (samples) within x
  long long produced = 0;
  int sync_tags = 0, lost_total = -1, fb = 0;
  int rc;
  if (h->state_is_zero) {
    memset(hs, 0, sizeof(AcqState));       // right after acq_reset(): the device state is the memset that is still in the stream - no round trip
    h->state_is_zero = false;
  } else {
    DVBT_CUDA_TRY(cudaMemcpyAsync(hs, h->d_state.p, sizeof(AcqState), cudaMemcpyDeviceToHost, st));
    DVBT_CUDA_TRY(dvbt::stream_wait(st));
  }
  // frequency-domain output: derotation + FFT in one kernel for the two DVB-T sizes (cuFFT otherwise)
  const bool fused_fft = do_fft && (p.N == 2048 || p.N == 8192) && !getenv("DVBT_B200_ACQ_CUFFT");
  if (fused_fft && h->tw_n != p.N) {
    std::vector<float2> tw(p.N);
    for (int i = 0; i < p.N; i++) {
      double a = -2.0 * M_PI * (double)i / (double)p.N;
      tw[i] = make_float2((float)cos(a), (float)sin(a));
    }
    if ((rc = h->d_tw.reserve((size_t)p.N * sizeof(float2)))) return rc;
    DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_tw.p, tw.data(), (size_t)p.N * sizeof(float2), cudaMemcpyHostToDevice, st));
    DVBT_CUDA_TRY(dvbt::stream_wait(st));
    h->tw_n = p.N;
  }
  h->n_fft_ev = 0;
  int guard = 0;
  while (guard++ < 1000000) {
    // ---- initial acquisition (needs 2N+cp+8 samples visible)
    int probe_lost1 = 0;
    if (!hs->initial) {
      if (n - pos < 2LL * p.N + p.cp + 8 || produced >= out_capacity_syms) break;
      if ((rc = h->d_il.reserve((size_t)p.N * 4)) || (rc = h->d_ig.reserve((size_t)p.N * 8))) return rc;
      acq_init_lambda_kernel<<<(p.N + 127) / 128, 128, 0, st>>>(p, x, pos, h->d_il.as<float>(), h->d_ig.as<float2>());
      DVBT_CUDA_TRY(cudaFuncSetAttribute(acq_init_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p.N * 8));
      acq_init_peak_kernel<<<1, 256, (size_t)p.N * 8, st>>>(p, h->d_il.as<float>(), h->d_ig.as<float2>(), h->d_state.as<AcqState>());
      acq_probe_kernel<<<1, kProbe * kCand, 0, st>>>(p, x, pos, n, h->d_state.as<AcqState>(), -(long long)nhist);
      count_launch(3);
      DVBT_CUDA_TRY(cudaGetLastError());
      DVBT_CUDA_TRY(cudaMemcpyAsync(hs, h->d_state.p, sizeof(AcqState), cudaMemcpyDeviceToHost, st));
      DVBT_CUDA_TRY(dvbt::stream_wait(st));
      probe_lost1 = hs->probe_lost1;
      sync_tags++;  // send_sync_start() on every attempt (:507), tagged at nitems_written = symbols produced so far
      if (sync_at && (sync_at->empty() || sync_at->back() != produced)) sync_at->push_back(produced);
      if (!hs->initial) {
        // nothing found: the reference consumes d_to_consume = N+cp (set by ml_sync's miss branch)
        pos += total;
        continue;
      }
    }
    // ---- tracking: symbols whose whole candidate table is inside the buffer
    int c0 = hs->cp_start;
    long long avail = n - pos - (c0 + kD + 1);
    long long nsym = avail < 0 ? 0 : avail / total + 1;
    if (nsym > out_capacity_syms - produced) nsym = out_capacity_syms - produced;
    if (nsym <= 0) break;
    bool capped = false;   // the finish kernel stages one batch in shared memory: longer inputs go in several batches
    if (nsym > kFinishMax) { nsym = kFinishMax; capped = true; }
    if (probe_lost1 > 0 && nsym > probe_lost1) {   // the look-ahead saw a miss: no tables beyond it (the miss ends the batch and
      nsym = probe_lost1;                          // acquisition restarts; if it does not, the next batch simply carries on)
      capped = true;
    }
    if ((rc = h->d_lambda.reserve((size_t)nsym * kCand * 4)) || (rc = h->d_gamma.reserve((size_t)nsym * kCand * 8)) ||
        (rc = h->d_avg1.reserve((size_t)nsym * 4)) || (rc = h->d_avg2.reserve((size_t)nsym * 4)) ||
        (rc = h->d_peak.reserve((size_t)nsym * 4)) || (rc = h->d_sym.reserve((size_t)nsym * sizeof(SymOut))))
      return rc;
    {
      long long threads = nsym * kCand;
      acq_lambda_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(p, x, pos, c0, (int)nsym, h->d_lambda.as<float>(),
                                                                          h->d_gamma.as<float2>(), -(long long)nhist);
      int per_thread = kChunk;
      while ((nsym + per_thread - 1) / per_thread > 1024) per_thread *= 2;
      int nthreads = (int)((nsym + per_thread - 1) / per_thread);
      if ((rc = h->d_avg1.reserve((size_t)nsym * kNC * 4)) || (rc = h->d_avg2.reserve((size_t)nsym * kNS * 4)) ||
          (rc = h->d_peak.reserve((size_t)nsym * kNS)) || (rc = h->d_flag.reserve((size_t)nsym * kNS)) ||
          (rc = h->d_eps.reserve(256)))
        return rc;
      long long t1 = nsym * kNC, t2 = nsym * kNS;
      acq_pass1_kernel<<<(unsigned)((t1 + 127) / 128), 128, 0, st>>>(p, (int)nsym, h->d_lambda.as<float>(), h->d_avg1.as<float>());
      acq_pass2_kernel<<<(unsigned)((t2 + 127) / 128), 128, 0, st>>>(p, (int)nsym, h->d_lambda.as<float>(), h->d_avg1.as<float>(), hs->avg,
                                                                    h->d_peak.as<signed char>(), h->d_avg2.as<float>(),
                                                                    h->d_flag.as<unsigned char>());
      if ((rc = h->d_maps.reserve((size_t)nthreads * kNS)) || (rc = h->d_cof.reserve((size_t)nsym)) || (rc = h->d_bof.reserve((size_t)nsym)))
        return rc;
      acq_chunkmap_kernel<<<(nthreads * 32 + 127) / 128, 128, 0, st>>>((int)nsym, per_thread, nthreads, h->d_flag.as<unsigned char>(),
                                                               h->d_maps.as<unsigned char>());
      AcqWalk *d_walk = h->d_eps.as<AcqWalk>();
      if ((rc = h->d_seg.reserve((size_t)kMaxSeg * sizeof(int4)))) return rc;
      DVBT_CUDA_TRY(cudaFuncSetAttribute(acq_compose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      acq_compose_kernel<<<1, 256, (size_t)nthreads * kNS + 64, st>>>(
          p, (int)nsym, per_thread, nthreads, (kD - 8) * kND + kHD, hs->avg, h->d_lambda.as<float>(), h->d_avg1.as<float>(),
          h->d_peak.as<signed char>(), h->d_avg2.as<float>(), h->d_flag.as<unsigned char>(), h->d_maps.as<unsigned char>(),
          h->d_cof.as<unsigned char>(), h->d_bof.as<signed char>(), h->d_seg.as<int4>(), d_walk);
      {
        int rows_cap = per_thread < kRowsCap ? per_thread : kRowsCap;
        size_t per_warp = (size_t)((rows_cap * kNS + 32 + rows_cap + 15) & ~15);
        DVBT_CUDA_TRY(cudaFuncSetAttribute(acq_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(4 * per_warp)));
        acq_walk_kernel<<<kMaxSeg / 4, 128, 4 * per_warp, st>>>(rows_cap, h->d_peak.as<signed char>(), h->d_avg2.as<float>(),
                                                               h->d_flag.as<unsigned char>(), h->d_seg.as<int4>(), h->d_cof.as<unsigned char>(),
                                                               h->d_bof.as<signed char>(), d_walk);
      }
      if ((rc = h->d_peakof.reserve((size_t)nsym * 4)) || (rc = h->d_eof.reserve((size_t)nsym * 4))) return rc;
      acq_post_kernel<<<(unsigned)((nsym + 127) / 128), 128, 0, st>>>(p, c0, h->d_gamma.as<float2>(), h->d_cof.as<unsigned char>(),
                                                                     h->d_bof.as<signed char>(), d_walk, h->d_peakof.as<int>(), h->d_eof.as<float>());
      if ((rc = h->d_runs.reserve(1024 * 2 * 8))) return rc;
      DVBT_CUDA_TRY(cudaFuncSetAttribute(acq_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFinishMax * 8));
      acq_finish_kernel<<<1, 1024, (size_t)nsym * 8, st>>>(p, pos, h->d_peakof.as<int>(), h->d_eof.as<float>(), d_walk, h->d_state.as<AcqState>(),
                                                         h->d_runs.as<double>());
      // the schedule's initial values are the host's copy of the state this batch started from
      acq_desc_kernel<<<(unsigned)((nsym + 127) / 128), 128, 0, st>>>(p, pos, h->d_peakof.as<int>(), h->d_eof.as<float>(), d_walk, hs->nextpos,
                                                                     hs->nextphaseinc, h->d_runs.as<double>(), h->d_sym.as<SymOut>());
      count_launch(9);
      DVBT_CUDA_TRY(cudaGetLastError());
    }
    DVBT_CUDA_TRY(cudaMemcpyAsync(hs, h->d_state.p, sizeof(AcqState), cudaMemcpyDeviceToHost, st));
    DVBT_CUDA_TRY(dvbt::stream_wait(st));
    if (getenv("DVBT_B200_ACQ_TRACE")) {
      AcqWalk wk;
      if (cudaMemcpy(&wk, h->d_eps.p, sizeof wk, cudaMemcpyDeviceToHost) == cudaSuccess)
        fprintf(stderr, "acq batch: nsym %lld found %d code %d override %d segments %d staged %d cycles maps %lld serial %lld (stage %lld walk %lld detect %lld rest %lld) | finish %lld %lld %lld %lld %lld\n", nsym,
                wk.n_found, wk.code, wk.n_override, wk.n_seg, wk.n_staged, wk.cyc_maps, wk.cyc_serial, wk.cyc_cmp[0], wk.cyc_cmp[1], wk.cyc_cmp[2],
                wk.cyc_cmp[3], wk.cyc_fin[0], wk.cyc_fin[1], wk.cyc_fin[2], wk.cyc_fin[3], wk.cyc_fin[4]);
    }
    if (getenv("DVBT_B200_ACQ_TRACE") && atoi(getenv("DVBT_B200_ACQ_TRACE")) >= 2) {
      // symbols whose peak is off centre, with their lambda windows (diagnostic)
      std::vector<unsigned char> hc((size_t)nsym);
      std::vector<signed char> hb((size_t)nsym);
      std::vector<float> hl((size_t)nsym * kCand), ha1((size_t)nsym * kNC);
      cudaMemcpy(hc.data(), h->d_cof.p, (size_t)nsym, cudaMemcpyDeviceToHost);
      cudaMemcpy(hb.data(), h->d_bof.p, (size_t)nsym, cudaMemcpyDeviceToHost);
      cudaMemcpy(hl.data(), h->d_lambda.p, (size_t)nsym * kCand * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(ha1.data(), h->d_avg1.p, (size_t)nsym * kNC * 4, cudaMemcpyDeviceToHost);
      int shown = 0;
      {
        long long hist[17] = {0}, run_far = 0, far = 0;
        for (long long m = 0; m < nsym; m++) {
          int b = hb[m] < 0 ? 16 : hb[m];
          hist[b]++;
          if (b < 6 || b > 10) { far++; if (m > 0 && (hb[m - 1] < 6 || hb[m - 1] > 10)) run_far++; }
        }
        fprintf(stderr, "acq best histogram:");
        for (int i = 0; i < 17; i++) fprintf(stderr, " %d:%lld", i, hist[i]);
        fprintf(stderr, " | far %lld, far after far %lld\n", far, run_far);
      }
      for (long long m = 1; m < nsym && shown < 0; m++) {
        if (hb[m] == 8 && hb[m - 1] == 8) continue;
        shown++;
        fprintf(stderr, "acq sym %lld c %d best %d avg1 %.9g | lambda:", m, hc[m], hb[m], ha1[m * kNC + hc[m]]);
        for (int i = 0; i < 16; i++) fprintf(stderr, " %.4g", hl[m * kCand + hc[m] + i]);
        fprintf(stderr, "\n");
      }
    }
    const bool time_fft = hs->n_out > 0 && fused_fft && h->n_fft_ev < dvbt_b200_acq::kFftEv;
    if (time_fft) {
      for (int e = 2 * h->n_fft_ev; e < 2 * h->n_fft_ev + 2; e++)
        if (!h->ev_fft[e]) DVBT_CUDA_TRY(cudaEventCreate(&h->ev_fft[e]));
      DVBT_CUDA_TRY(cudaEventRecord(h->ev_fft[2 * h->n_fft_ev], st));
    }
    if (hs->n_out > 0 && fused_fft) {
      if (p.N == 2048) {
        acq_fftd_kernel<2048><<<hs->n_out, 128, (2048 + 128) * sizeof(float2), st>>>(hs->n_out, x, h->d_sym.as<SymOut>(), d_out + produced * p.N,
                                                                             h->d_tw.as<float2>());
      } else {
        DVBT_CUDA_TRY(cudaFuncSetAttribute(acq_fftd_kernel<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (8192 + 512) * (int)sizeof(float2)));
        acq_fftd_kernel<8192><<<hs->n_out, 512, (8192 + 512) * sizeof(float2), st>>>(hs->n_out, x, h->d_sym.as<SymOut>(), d_out + produced * p.N,
                                                                             h->d_tw.as<float2>());
      }
      count_launch();
      DVBT_CUDA_TRY(cudaGetLastError());
      if (time_fft) { DVBT_CUDA_TRY(cudaEventRecord(h->ev_fft[2 * h->n_fft_ev + 1], st)); h->n_fft_ev++; }
    } else if (hs->n_out > 0) {
      dim3 grid((p.N + 256 * kDerotPer - 1) / (256 * kDerotPer), hs->n_out);
      acq_derot_kernel<<<grid, 256, 0, st>>>(p.N, hs->n_out, x, h->d_sym.as<SymOut>(), d_out + produced * p.N, do_fft ? 1 : 0);
      count_launch();
      DVBT_CUDA_TRY(cudaGetLastError());
    }
    produced += hs->n_out;
    fb |= hs->fallback;
    pos += hs->consumed;
    if (hs->lost_at >= 0) {
      // missed peak: the reference restarts acquisition and consumes half a symbol for that call (:545-558)
      if (lost_total < 0) lost_total = (int)produced;
      AcqState s2 = *hs;
      s2.initial = 0;
      pos += total / 2;
      DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_state.p, &s2, sizeof(AcqState), cudaMemcpyHostToDevice, st));
      DVBT_CUDA_TRY(dvbt::stream_wait(st));
      *hs = s2;
      continue;
    }
    if (hs->lost_at <= -2) continue;  // stopped early (table re-centre or speculation split): carry on from the true state
    if (capped) continue;             // batch limit: the rest of the input follows from the state just written
    break;
  }
  if (do_fft && !fused_fft && produced > 0) {
    if (h->plan == 0 || h->plan_batch != (int)produced) {
      if (h->plan) cufftDestroy(h->plan);
      h->plan = 0;
      int nn[1] = {p.N};
      if (cufftPlanMany(&h->plan, 1, nn, nullptr, 1, p.N, nullptr, 1, p.N, CUFFT_C2C, (int)produced) != CUFFT_SUCCESS) {
        set_error("acq: cufftPlanMany(%d x %lld) failed", p.N, produced);
        return DVBT_B200_ECUDA;
      }
      cufftSetStream(h->plan, st);
      h->plan_batch = (int)produced;
    }
    if (cufftExecC2C(h->plan, (cufftComplex *)d_out, (cufftComplex *)d_out, CUFFT_FORWARD) != CUFFT_SUCCESS) {
      set_error("acq: cufftExecC2C failed");
      return DVBT_B200_ECUDA;
    }
    count_launch();
  }
  hs->consumed = pos;
  hs->n_out = (int)produced;
  hs->n_sync_tags = sync_tags;
  hs->lost_at = lost_total;
  hs->fallback = fb;
  if (host_state_out) *host_state_out = *hs;
  return 0;
}

// device time of the derotation+FFT kernel(s) of the last run; call after the stream has been synchronised
float acq_last_fft_ms(dvbt_b200_acq *h) {
  float tot = 0.f, ms;
  for (int i = 0; i < h->n_fft_ev; i++)
    if (cudaEventElapsedTime(&ms, h->ev_fft[2 * i], h->ev_fft[2 * i + 1]) == cudaSuccess) tot += ms;
  return tot;
}

void acq_use_stream(dvbt_b200_acq *h, cudaStream_t st) {
  if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
  h->stream = st;
  h->own_stream = false;
  if (h->plan) { cufftDestroy(h->plan); h->plan = 0; h->plan_batch = 0; }
}

int acq_reset(dvbt_b200_acq *h) {
  DVBT_CUDA_TRY(cudaMemsetAsync(h->d_state.p, 0, sizeof(AcqState), h->stream));
  h->state_is_zero = true;
  h->pending_sync = false;
  h->hist_valid = false;
  return 0;
}

int acq_run_simple(dvbt_b200_acq *h, const float2 *x, long long n, float2 *d_out, long long out_capacity_syms, int do_fft,
                   AcqResult *res, std::vector<long long> *sync_at, int nhist) {
  AcqState hs;
  int rc = acq_run(h, x, n, d_out, out_capacity_syms, do_fft, &hs, sync_at, nhist);
  if (rc) return rc;
  if (res) { res->n_run = hs.n_run; res->n_single = hs.n_single; res->n_seq = hs.n_seq; res->consumed = hs.consumed; res->n_out = hs.n_out; res->lost_at = hs.lost_at; res->fallback = hs.fallback; res->cp_start = hs.cp_start; }
  return 0;
}

}  // namespace dvbt

extern "C" {

int dvbt_b200_acq_create(const dvbt_b200_acq_params *p, dvbt_b200_acq **out) {
  if (!p || !out) { set_error("acq_create: null argument"); return DVBT_B200_EINVAL; }
  *out = nullptr;
  if (p->blocks != 1 || (p->fft_length != 2048 && p->fft_length != 8192) || p->cp_length <= 0 || p->cp_length > p->fft_length / 4) {
    set_error("acq_create: blocks must be 1, fft_length 2048/8192, 0 < cp_length <= N/4");
    return DVBT_B200_EINVAL;
  }
  int rc = dvbt::ensure_device();
  if (rc) return rc;
  dvbt_b200_acq *h = new (std::nothrow) dvbt_b200_acq();
  if (!h) { set_error("acq_create: out of memory"); return DVBT_B200_ENOMEM; }
  h->par = *p;
  h->kp.N = p->fft_length;
  h->kp.cp = p->cp_length;
  // ofdm_sym_acquisition_impl.cc:390-391: d_snr = pow(10, snr/10) (float member), d_rho = d_snr/(d_snr+1.0)
  float snr = (float)pow(10, p->snr / 10.0);
  float rho = (float)(snr / (snr + 1.0));
  h->kp.rho2 = (float)(rho / 2.0);  // :236
  h->kp.rise = 0.8f; h->kp.fall = 0.9f; h->kp.alpha = 0.9f;  // :448
  h->h_state.host = true;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("acq_create: cannot create stream"); delete h; return DVBT_B200_ECUDA; }
  if ((rc = h->d_state.reserve(sizeof(AcqState))) || (rc = h->h_state.reserve(sizeof(AcqState)))) { dvbt_b200_acq_destroy(h); return rc; }
  cudaMemset(h->d_state.p, 0, sizeof(AcqState));
  *out = h;
  return 0;
}

void dvbt_b200_acq_destroy(dvbt_b200_acq *h) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) return;
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->plan) cufftDestroy(h->plan);
  dvbt::DevBuf *bufs[] = {&h->d_hist, &h->d_x, &h->d_state, &h->h_state, &h->d_lambda, &h->d_gamma, &h->d_avg1, &h->d_avg2, &h->d_peak, &h->d_sym, &h->d_out, &h->d_il, &h->d_ig, &h->d_eps, &h->d_flag, &h->d_maps, &h->d_cof, &h->d_bof, &h->d_peakof, &h->d_eof, &h->d_seg, &h->d_runs, &h->d_tw};
  for (auto *b : bufs) b->release();
  h->stg.release();
  for (auto &e : h->ev_fft) if (e) cudaEventDestroy(e);
  if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
}

int dvbt_b200_acq_work(dvbt_b200_acq *h, const void *in, size_t n_in_items, void *out, size_t out_capacity_items, size_t *consumed,
                       size_t *produced, dvbt_b200_tag *tags_out, size_t tags_out_capacity, size_t *n_tags_out, int apply_fft) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || !consumed || !produced) { set_error("acq_work: null argument"); return DVBT_B200_EINVAL; }
  *consumed = *produced = 0;
  if (n_tags_out) *n_tags_out = 0;
  if (!in || !out || out_capacity_items == 0) return 0;
  int rc;
  const int N = h->kp.N;
  // [kAcqHistory samples consumed by the previous calls (zeros after a reset) | this call's input]: the reference's window
  // reaches a few samples in front of its input pointer, into the scheduler's circular buffer (ml_point)
  constexpr int H = dvbt::kAcqHistory;
  if ((rc = h->d_x.reserve((H + n_in_items) * 8)) || (rc = h->d_hist.reserve((size_t)H * 8)) || (rc = h->d_out.reserve(out_capacity_items * N * 8))) return rc;
  if (!h->hist_valid) { DVBT_CUDA_TRY(cudaMemsetAsync(h->d_hist.p, 0, (size_t)H * 8, h->stream)); h->hist_valid = true; }
  DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_x.p, h->d_hist.p, (size_t)H * 8, cudaMemcpyDeviceToDevice, h->stream));
  if ((rc = h->stg.h2d(h->d_x.as<float2>() + H, in, n_in_items * 8, h->stream))) return rc;
  AcqState hs;
  std::vector<long long> sync_at;
  rc = dvbt::acq_run(h, h->d_x.as<float2>() + H, (long long)n_in_items, h->d_out.as<float2>(), (long long)out_capacity_items, apply_fft, &hs, &sync_at, H);
  if (rc) return rc;
  // the H samples in front of the new read position (d_x still holds the old history in front of the input)
  DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_hist.p, h->d_x.as<float2>() + hs.consumed, (size_t)H * 8, cudaMemcpyDeviceToDevice, h->stream));
  if (hs.n_out > 0 && (rc = h->stg.d2h(out, h->d_out.p, (size_t)hs.n_out * N * 8, h->stream))) return rc;
  DVBT_CUDA_TRY(dvbt::stream_wait(h->stream));
  *consumed = (size_t)hs.consumed;
  *produced = (size_t)hs.n_out;
  // send_sync_start() (:353-360) on every acquisition attempt (:507), at the output position it happened at: the first
  // item of the call for a fresh or still unlocked receiver, later items when the lock was lost and regained inside the
  // call.  A tag beyond the last produced item belongs to the next call's first item (the attempt is repeated there).
  if (h->pending_sync && (sync_at.empty() || sync_at.front() != 0)) sync_at.insert(sync_at.begin(), 0);
  h->pending_sync = !sync_at.empty() && sync_at.back() >= (long long)hs.n_out;
  if (tags_out && n_tags_out) {
    size_t nt = 0;
    for (long long off : sync_at)
      if (off < (long long)hs.n_out && nt < tags_out_capacity) tags_out[nt++] = dvbt_b200_tag{(uint64_t)off, DVBT_TAG_SYNC_START, 1};
    *n_tags_out = nt;
  }
  return 0;
}

}  // extern "C"
