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
// candidates.  Here a whole capture is processed at once:
//   acq_lambda_kernel   one thread per (symbol, candidate): gamma, phi, lambda with the
//                       reference's summation order (so lambda is bit-identical: it depends only
//                       on the absolute sample position);
//   acq_initial_kernel  the one-off search over N candidates (sequential peak detector);
//   acq_track_kernel    one thread per symbol runs the 16-step peak detector.  The detector's
//                       running average d_avg is the only state that crosses symbols; it is
//                       speculated (pass 1: from zero, pass 2: from the previous symbol's pass-1
//                       value) and acq_chain_kernel verifies the chain bit-for-bit, falling back to
//                       the sequential detector from the first symbol where speculation, the peak
//                       position or a missed peak breaks the assumption;
//   acq_chain_kernel    also carries cp_start and the phase-increment schedule (:285-312) and
//                       produces per-symbol (first sample, start phase, increments);
//   acq_derot_kernel    out[j] = expj(phase_j) * in[cp_start-N+1+j], times (-1)^j so that the
//                       unshifted cuFFT output equals fft_vxx's shifted output.
// Difference to the reference, by construction: the derotation phase is evaluated in closed
// form in double instead of N+cp sequential float additions per symbol (:285-309), so the
// output samples agree to ~1e-6 relative, not bit-for-bit; decisions (timing, peaks) are exact.
#include "chain_internal.cuh"

#include <cufft.h>
#include <math.h>
#include <string.h>
#include <new>

namespace {

using dvbt::set_error;

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
  return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cmul_conjf(float2 a, float2 b) { return cmulf(a, make_float2(b.x, -b.y)); }
__device__ __forceinline__ float cnormf(float2 a) { return __fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)); }

// lambda/gamma for the candidate symbol end at absolute sample index p (ofdm_sym_acquisition_impl.cc:162-250)
__device__ __forceinline__ void ml_point(const float2 *__restrict__ x, long long p, int N, int cp, float rho2, float *lambda,
                                         float2 *gamma) {
  float2 g = make_float2(0.f, 0.f);
  float phi = 0.f;
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
__device__ int peak_detect(const float *d, int n, float *avg_io, float rise, float fall, float alpha, int *best) {
  float avg = *avg_io;
  int state = 0, npeaks = 0, peak_index = 0, best_idx = 0;
  float peak_val = -INFINITY, best_val = 0.f;
  float one_minus = 1.0f - alpha;  // (1 - d_avg_alpha) in float
  int i = 0;
  while (i < n) {
    float v = d[i];
    if (state == 0) {
      if (v > __fmul_rn(avg, rise)) {
        state = 1;
      } else {
        avg = __fadd_rn(__fmul_rn(alpha, v), __fmul_rn(one_minus, avg));
        i++;
      }
    } else {
      if (v > peak_val) {
        peak_val = v;
        peak_index = i;
        avg = __fadd_rn(__fmul_rn(alpha, v), __fmul_rn(one_minus, avg));
        i++;
      } else if (v > __fmul_rn(avg, fall)) {
        avg = __fadd_rn(__fmul_rn(alpha, v), __fmul_rn(one_minus, avg));
        i++;
      } else {
        // record the peak; the first strictly greatest recorded peak wins (:127-137)
        if (npeaks == 0 || d[peak_index] > best_val) { best_val = d[peak_index]; best_idx = peak_index; }
        npeaks++;
        state = 0;
        peak_val = -INFINITY;
      }
    }
  }
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

__global__ void acq_init_peak_kernel(AcqParams p, const float *__restrict__ lambda, const float2 *__restrict__ gamma, AcqState *st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float avg = st->avg;
  int best = 0;
  int n = peak_detect(lambda, p.N, &avg, p.rise, p.fall, p.alpha, &best);
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

// ---- tracking table: symbol n, candidate c <-> symbol end at base + n*(N+cp) + c0 - kD + c
__global__ void acq_lambda_kernel(AcqParams p, const float2 *__restrict__ x, long long base, int c0, int nsym,
                                  float *__restrict__ lambda, float2 *__restrict__ gamma) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)nsym * kCand) return;
  int n = (int)(t / kCand), c = (int)(t % kCand);
  ml_point(x, base + (long long)n * (p.N + p.cp) + c0 - kD + c, p.N, p.cp, p.rho2, &lambda[t], &gamma[t]);
}

// speculative per-symbol detector: window = candidates [kD-8, kD+8) (cp_start == c0)
__global__ void acq_track_kernel(AcqParams p, int nsym, const float *__restrict__ lambda, const float *__restrict__ avg_in,
                                 float avg_first, float *__restrict__ avg_out, int *__restrict__ peak_out,
                                 const float2 *__restrict__ gamma, float *__restrict__ eps_out) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nsym) return;
  float avg = avg_in ? (n == 0 ? avg_first : avg_in[n - 1]) : 0.f;
  int best = 0;
  int np = peak_detect(lambda + (long long)n * kCand + (kD - 8), 16, &avg, p.rise, p.fall, p.alpha, &best);
  avg_out[n] = avg;
  if (peak_out) peak_out[n] = np > 0 ? best : -1;
  if (eps_out) {
    float2 g = gamma[(long long)n * kCand + (kD - 8) + best];
    eps_out[n] = np > 0 ? atan2f(g.y, g.x) : 0.f;  // fast_atan2f(d_gamma[peak]) (:277)
  }
}

// speculation holds for the whole batch iff every symbol found its peak at the centre of the window
// (cp_start unchanged) and every average fed forward in pass 2 equals the one pass 2 produced
__global__ void acq_verify_kernel(int nsym, const float *__restrict__ avg1, const float *__restrict__ avg2,
                                  const int *__restrict__ peak2, int *first_bad) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nsym) return;
  bool bad = peak2[n] != 8;
  if (n + 1 < nsym && __float_as_uint(avg1[n]) != __float_as_uint(avg2[n])) bad = true;
  if (bad) atomicMin(first_bad, n);
}

// whole batch verified: per-symbol output descriptors and the end state, one warp, no dependent loads
__global__ void __launch_bounds__(32) acq_fast_kernel(AcqParams p, int nsym, long long base, int c0, const float *__restrict__ eps,
                                                      const float *__restrict__ avg2, const int *first_bad, AcqState *st,
                                                      SymOut *__restrict__ out) {
  if (*first_bad < nsym) return;
  const int lane = threadIdx.x;
  const int total = p.N + p.cp;
  const double invN = -1.0 / (double)p.N;
  const int sw = c0 - total;                       // d_nextpos left by every symbol (:312)
  const bool sw_ok = sw >= 0 && sw < total;
  const int sw0 = st->nextpos;
  const bool sw0_ok = sw0 >= 0 && sw0 < total;
  const double inc_init = st->phaseinc, pend_init = st->nextphaseinc;
  const double inc_after0 = sw0_ok ? pend_init : inc_init;  // d_phaseinc after symbol 0
  double carry = st->phase;                         // phase before symbol `base_n`
  double last_inc = inc_init;
  for (int bn = 0; bn < nsym; bn += 32) {
    int n = bn + lane;
    double i0 = 0, i1 = 0, adv = 0;
    int swn = total;
    if (n < nsym) {
      if (n == 0) {
        i0 = inc_init; i1 = pend_init; swn = sw0_ok ? sw0 : total;
      } else {
        double e1 = invN * (double)eps[n - 1];
        // increment in force when symbol n starts: what symbol n-1 switched to (or kept)
        double start = (n == 1) ? inc_after0 : (sw_ok ? invN * (double)eps[n - 2] : inc_after0);
        i0 = start; i1 = e1; swn = sw_ok ? sw : total;
      }
      adv = swn < total ? swn * i0 + (total - swn) * i1 : total * i0;
    }
    // inclusive warp scan of the advances
    double incl = adv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (n < nsym) {
      SymOut so;
      so.first = base + (long long)n * total + c0 - p.N + 1;
      so.phase0 = remainder(carry + (incl - adv), 2.0 * M_PI);
      so.inc0 = i0; so.inc1 = i1; so.switch_at = swn;
      out[n] = so;
    }
    int lastl = min(31, nsym - 1 - bn);
    carry = remainder(carry + __shfl_sync(0xffffffffu, incl, lastl), 2.0 * M_PI);
    double li = (swn < total) ? i1 : i0;
    last_inc = __shfl_sync(0xffffffffu, li, lastl);
  }
  if (lane == 0) {
    st->avg = avg2[nsym - 1];
    st->phase = (float)carry;
    st->phaseinc = last_inc;
    st->nextphaseinc = invN * (double)eps[nsym - 1];
    st->nextpos = sw;
    st->cp_start = c0;
    st->n_out = nsym;
    st->lost_at = -1;
    st->fallback = 0;
    st->consumed = (long long)nsym * total;
  }
}

// verification + the light sequential bookkeeping; falls back to the sequential detector from the
// first symbol whose speculation does not hold
__global__ void acq_chain_kernel(AcqParams p, int nsym, long long base, int c0, const float *__restrict__ lambda,
                                 const float2 *__restrict__ gamma, const float *__restrict__ avg1, const float *__restrict__ avg2,
                                 const int *__restrict__ peak2, AcqState *st, SymOut *__restrict__ out, const int *first_bad) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (*first_bad >= nsym) return;  // acq_fast_kernel handled the batch
  const int total = p.N + p.cp;
  int cp_start = c0;
  float avg = st->avg;
  double ph = st->phase, inc = st->phaseinc, nextinc = st->nextphaseinc;
  int nextpos = st->nextpos;
  int n_out = 0, lost_at = -1, fallback = 0;
  bool spec_ok = true;
  for (int n = 0; n < nsym; n++) {
    int found, best;
    float2 g;
    // speculation holds for symbol n if (a) we are still on the speculated timing, (b) the average
    // it started from is the true one.  avg2[n] was computed from avg1[n-1]; the true incoming average
    // is `avg` (exact by induction).
    float spec_in = n == 0 ? st->avg : avg1[n - 1];
    if (spec_ok && cp_start == c0 && __float_as_uint(spec_in) == __float_as_uint(avg)) {
      best = peak2[n];
      found = best >= 0;
      avg = avg2[n];
    } else {
      // sequential detector on the true state
      int lo = cp_start - 8 - (c0 - kD);
      if (lo < 0 || lo + 16 > kCand) { lost_at = n; break; }  // drifted out of the table: stop here
      fallback = 1;
      found = peak_detect(lambda + (long long)n * kCand + lo, 16, &avg, p.rise, p.fall, p.alpha, &best) > 0;
      if (!found) best = -1;
      spec_ok = false;
    }
    int lo = cp_start - 8 - (c0 - kD);
    SymOut so;
    so.phase0 = ph;
    so.inc0 = inc;
    so.inc1 = inc;
    so.switch_at = total;
    if (found) {
      if (nextpos >= 0 && nextpos < total) {  // :287-288
        so.inc1 = nextinc;
        so.switch_at = nextpos;
        ph += nextpos * inc;
        inc = nextinc;
        ph += (total - nextpos) * inc;
      } else {
        ph += total * inc;
      }
      ph = remainder(ph, 2.0 * M_PI);
      int peak = best + cp_start - 8;
      g = gamma[(long long)n * kCand + lo + best];
      float eps = atan2f(g.y, g.x);
      nextinc = (-1.0 / (double)p.N) * (double)eps;  // :311
      nextpos = peak - total;                        // :312
      cp_start = peak;
      so.first = base + (long long)n * total + cp_start - p.N + 1;
      out[n_out++] = so;
    } else {
      // missed peak: phase still advances (:335-343); timeout is 0 so acquisition restarts (:545-558)
      ph = remainder(ph + total * inc, 2.0 * M_PI);
      lost_at = n;
      break;
    }
  }
  st->avg = avg;
  st->phase = (float)ph;
  st->phaseinc = inc;
  st->nextphaseinc = nextinc;
  st->nextpos = nextpos;
  st->cp_start = cp_start;
  st->n_out = n_out;
  st->lost_at = lost_at;
  st->fallback = 1;
  (void)fallback;
  if (lost_at >= 0) {
    // symbols 0..lost_at-1 consumed N+cp each; the miss consumes N+cp (table overrun) or half of it (restart)
    st->consumed = (long long)lost_at * total;
  } else {
    st->consumed = (long long)nsym * total;
  }
}

// out[n][j] = (-1)^j * expj(phase_j) * x[first + j]
__global__ void __launch_bounds__(256) acq_derot_kernel(int N, int nsym, const float2 *__restrict__ x, const SymOut *__restrict__ so,
                                                        float2 *__restrict__ out, int shift_sign) {
  int n = blockIdx.y;
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nsym || j >= N) return;
  SymOut s = so[n];
  int steps = j + 1;  // the phase is incremented before it is used (:291-307)
  double ph = s.phase0 + (steps <= s.switch_at ? steps * s.inc0 : s.switch_at * s.inc0 + (steps - s.switch_at) * s.inc1);
  ph = remainder(ph, 2.0 * M_PI);
  float sn, cs;
  sincosf((float)ph, &sn, &cs);
  float2 v = cmulf(make_float2(cs, sn), x[s.first + j]);
  if (shift_sign && (j & 1)) v = make_float2(-v.x, -v.y);
  out[(long long)n * N + j] = v;
}

}  // namespace

struct dvbt_b200_acq {
  dvbt_b200_acq_params par;
  AcqParams kp;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  cufftHandle plan = 0;
  int plan_batch = 0;
  dvbt::DevBuf d_x, d_state, h_state, d_lambda, d_gamma, d_avg1, d_avg2, d_peak, d_sym, d_out, d_il, d_ig, d_eps, d_flag;
};

namespace dvbt {

// Runs acquisition + derotation (+ optional FFT) over device samples x[0..n).  Output symbols go to
// d_out (N complex each).  Returns counts through the host copy of the state.
int acq_run(dvbt_b200_acq *h, const float2 *x, long long n, float2 *d_out, long long out_capacity_syms, int do_fft,
            AcqState *host_state_out) {
  const AcqParams &p = h->kp;
  const int total = p.N + p.cp;
  cudaStream_t st = h->stream;
  AcqState *hs = h->h_state.as<AcqState>();
  long long pos = 0;       // read position (samples) within x
  long long produced = 0;
  int sync_tags = 0, lost_total = -1, fb = 0;
  int rc;
  DVBT_CUDA_TRY(cudaMemcpyAsync(hs, h->d_state.p, sizeof(AcqState), cudaMemcpyDeviceToHost, st));
  DVBT_CUDA_TRY(cudaStreamSynchronize(st));
  int guard = 0;
  while (guard++ < 64) {
    // ---- initial acquisition (needs 2N+cp+8 samples visible)
    if (!hs->initial) {
      if (n - pos < 2LL * p.N + p.cp + 8 || produced >= out_capacity_syms) break;
      if ((rc = h->d_il.reserve((size_t)p.N * 4)) || (rc = h->d_ig.reserve((size_t)p.N * 8))) return rc;
      acq_init_lambda_kernel<<<(p.N + 127) / 128, 128, 0, st>>>(p, x, pos, h->d_il.as<float>(), h->d_ig.as<float2>());
      acq_init_peak_kernel<<<1, 32, 0, st>>>(p, h->d_il.as<float>(), h->d_ig.as<float2>(), h->d_state.as<AcqState>());
      count_launch(2);
      DVBT_CUDA_TRY(cudaGetLastError());
      DVBT_CUDA_TRY(cudaMemcpyAsync(hs, h->d_state.p, sizeof(AcqState), cudaMemcpyDeviceToHost, st));
      DVBT_CUDA_TRY(cudaStreamSynchronize(st));
      sync_tags++;  // send_sync_start() on every attempt (:507)
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
    if ((rc = h->d_lambda.reserve((size_t)nsym * kCand * 4)) || (rc = h->d_gamma.reserve((size_t)nsym * kCand * 8)) ||
        (rc = h->d_avg1.reserve((size_t)nsym * 4)) || (rc = h->d_avg2.reserve((size_t)nsym * 4)) ||
        (rc = h->d_peak.reserve((size_t)nsym * 4)) || (rc = h->d_sym.reserve((size_t)nsym * sizeof(SymOut))))
      return rc;
    {
      long long threads = nsym * kCand;
      acq_lambda_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(p, x, pos, c0, (int)nsym, h->d_lambda.as<float>(),
                                                                          h->d_gamma.as<float2>());
      unsigned g = (unsigned)((nsym + 127) / 128);
      if ((rc = h->d_eps.reserve((size_t)nsym * 4)) || (rc = h->d_flag.reserve(16))) return rc;
      int big = 0x7fffffff;
      DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_flag.p, &big, 4, cudaMemcpyHostToDevice, st));
      acq_track_kernel<<<g, 128, 0, st>>>(p, (int)nsym, h->d_lambda.as<float>(), nullptr, 0.f, h->d_avg1.as<float>(), nullptr, nullptr, nullptr);
      acq_track_kernel<<<g, 128, 0, st>>>(p, (int)nsym, h->d_lambda.as<float>(), h->d_avg1.as<float>(), hs->avg, h->d_avg2.as<float>(),
                                          h->d_peak.as<int>(), h->d_gamma.as<float2>(), h->d_eps.as<float>());
      acq_verify_kernel<<<g, 128, 0, st>>>((int)nsym, h->d_avg1.as<float>(), h->d_avg2.as<float>(), h->d_peak.as<int>(), h->d_flag.as<int>());
      acq_fast_kernel<<<1, 32, 0, st>>>(p, (int)nsym, pos, c0, h->d_eps.as<float>(), h->d_avg2.as<float>(), h->d_flag.as<int>(),
                                        h->d_state.as<AcqState>(), h->d_sym.as<SymOut>());
      acq_chain_kernel<<<1, 32, 0, st>>>(p, (int)nsym, pos, c0, h->d_lambda.as<float>(), h->d_gamma.as<float2>(), h->d_avg1.as<float>(),
                                         h->d_avg2.as<float>(), h->d_peak.as<int>(), h->d_state.as<AcqState>(), h->d_sym.as<SymOut>(),
                                         h->d_flag.as<int>());
      count_launch(6);
      DVBT_CUDA_TRY(cudaGetLastError());
    }
    DVBT_CUDA_TRY(cudaMemcpyAsync(hs, h->d_state.p, sizeof(AcqState), cudaMemcpyDeviceToHost, st));
    DVBT_CUDA_TRY(cudaStreamSynchronize(st));
    if (hs->n_out > 0) {
      dim3 grid((p.N + 255) / 256, hs->n_out);
      acq_derot_kernel<<<grid, 256, 0, st>>>(p.N, hs->n_out, x, h->d_sym.as<SymOut>(), d_out + produced * p.N, do_fft ? 1 : 0);
      count_launch();
      DVBT_CUDA_TRY(cudaGetLastError());
    }
    produced += hs->n_out;
    fb |= hs->fallback;
    pos += hs->consumed;
    if (hs->lost_at >= 0) {
      // either a missed peak (restart: consume half a symbol, :557) or the timing left the table
      // (re-centre the table; nothing is consumed for the symbol that could not be evaluated)
      int lo = hs->cp_start - 8 - (c0 - kD);
      bool off_table = (lo < 0 || lo + 16 > kCand);
      if (!off_table) {
        if (lost_total < 0) lost_total = (int)produced;
        AcqState s2 = *hs;
        s2.initial = 0;
        pos += total / 2;
        DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_state.p, &s2, sizeof(AcqState), cudaMemcpyHostToDevice, st));
        DVBT_CUDA_TRY(cudaStreamSynchronize(st));
        *hs = s2;
      }
      continue;
    }
    break;
  }
  if (do_fft && produced > 0) {
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

void acq_use_stream(dvbt_b200_acq *h, cudaStream_t st) {
  if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
  h->stream = st;
  h->own_stream = false;
  if (h->plan) { cufftDestroy(h->plan); h->plan = 0; h->plan_batch = 0; }
}

int acq_reset(dvbt_b200_acq *h) {
  DVBT_CUDA_TRY(cudaMemsetAsync(h->d_state.p, 0, sizeof(AcqState), h->stream));
  return 0;
}

int acq_run_simple(dvbt_b200_acq *h, const float2 *x, long long n, float2 *d_out, long long out_capacity_syms, int do_fft,
                   AcqResult *res) {
  AcqState hs;
  int rc = acq_run(h, x, n, d_out, out_capacity_syms, do_fft, &hs);
  if (rc) return rc;
  if (res) { res->consumed = hs.consumed; res->n_out = hs.n_out; res->lost_at = hs.lost_at; res->fallback = hs.fallback; res->cp_start = hs.cp_start; }
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
  if (!h) return;
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->plan) cufftDestroy(h->plan);
  dvbt::DevBuf *bufs[] = {&h->d_x, &h->d_state, &h->h_state, &h->d_lambda, &h->d_gamma, &h->d_avg1, &h->d_avg2, &h->d_peak, &h->d_sym, &h->d_out, &h->d_il, &h->d_ig, &h->d_eps, &h->d_flag};
  for (auto *b : bufs) b->release();
  if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
}

int dvbt_b200_acq_work(dvbt_b200_acq *h, const void *in, size_t n_in_items, void *out, size_t out_capacity_items, size_t *consumed,
                       size_t *produced, dvbt_b200_tag *tags_out, size_t tags_out_capacity, size_t *n_tags_out, int apply_fft) {
  if (!h || !consumed || !produced) { set_error("acq_work: null argument"); return DVBT_B200_EINVAL; }
  *consumed = *produced = 0;
  if (n_tags_out) *n_tags_out = 0;
  if (!in || !out || out_capacity_items == 0) return 0;
  int rc;
  const int N = h->kp.N;
  if ((rc = h->d_x.reserve(n_in_items * 8)) || (rc = h->d_out.reserve(out_capacity_items * N * 8))) return rc;
  DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_x.p, in, n_in_items * 8, cudaMemcpyHostToDevice, h->stream));
  AcqState hs;
  bool was_initial = false;
  {
    DVBT_CUDA_TRY(cudaMemcpyAsync(&hs, h->d_state.p, sizeof hs, cudaMemcpyDeviceToHost, h->stream));
    DVBT_CUDA_TRY(cudaStreamSynchronize(h->stream));
    was_initial = hs.initial != 0;
  }
  rc = dvbt::acq_run(h, h->d_x.as<float2>(), (long long)n_in_items, h->d_out.as<float2>(), (long long)out_capacity_items, apply_fft, &hs);
  if (rc) return rc;
  if (hs.n_out > 0) DVBT_CUDA_TRY(cudaMemcpyAsync(out, h->d_out.p, (size_t)hs.n_out * N * 8, cudaMemcpyDeviceToHost, h->stream));
  DVBT_CUDA_TRY(cudaStreamSynchronize(h->stream));
  *consumed = (size_t)hs.consumed;
  *produced = (size_t)hs.n_out;
  if (tags_out && n_tags_out && tags_out_capacity > 0 && !was_initial && hs.n_sync_tags > 0) {
    tags_out[0] = dvbt_b200_tag{0, DVBT_TAG_SYNC_START, 1};  // :353-360
    *n_tags_out = 1;
  }
  return 0;
}

}  // extern "C"
