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
  int n_run, n_single, n_seq;  // symbols handled in quiet runs / one at a time / by the sequential detector
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

// ---- speculative tracking tables --------------------------------------------------------------
// The reference state that crosses symbols is (cp_start, d_avg) plus the phase schedule.  cp_start
// moves by (best - 8) per symbol; the table covers window offsets c = 0..kNC-1 (window = candidates
// [c, c+16), cp_start = c0 - kD + 8 + c).  d_avg after the 16 values of a window is, to float
// precision, independent of the average it started from, so it is speculated:
//   pass 1: avg1[n][c]      = average after window (n, c) starting from 0
//   pass 2: best2/avg2[n][c][d] = detector result of window (n, c) starting from avg1[n-1][c+d], d in {-1,0,1}
//           (the previous symbol sat at offset c+d); ok0[n][c] = "no timing move and avg2 == avg1" for d = 0
// acq_walk_kernel then follows the true path, checking bit-for-bit that every speculated input equals
// the true average, and runs the plain sequential detector for any symbol where it does not.
constexpr int kNC = kCand - 16 + 1;  // 17 window offsets

__global__ void acq_pass1_kernel(AcqParams p, int nsym, const float *__restrict__ lambda, float *__restrict__ avg1) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsym * kNC) return;
  int n = t / kNC, c = t - n * kNC;
  float avg = 0.f;
  int best;
  peak_detect(lambda + (long long)n * kCand + c, 16, &avg, p.rise, p.fall, p.alpha, &best);
  avg1[t] = avg;
}

__global__ void acq_pass2_kernel(AcqParams p, int nsym, const float *__restrict__ lambda, const float *__restrict__ avg1, float avg_first,
                                 signed char *__restrict__ best2, float *__restrict__ avg2, unsigned char *__restrict__ ok0) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsym * kNC * 3) return;
  int n = t / (kNC * 3), r = t - n * (kNC * 3), c = r / 3, d = r - c * 3 - 1;
  int cp = c + d;
  signed char res = -2;  // -2: no speculation available
  float avg = 0.f;
  if (n == 0 || (cp >= 0 && cp < kNC)) {
    avg = n == 0 ? avg_first : avg1[(n - 1) * kNC + cp];
    int best;
    int np = peak_detect(lambda + (long long)n * kCand + c, 16, &avg, p.rise, p.fall, p.alpha, &best);
    res = np > 0 ? (signed char)best : (signed char)-1;
  }
  best2[t] = res;
  avg2[t] = avg;
  if (d == 0) ok0[n * kNC + c] = (res == 8 && __float_as_uint(avg) == __float_as_uint(avg1[n * kNC + c])) ? 1 : 0;
}

// The sequential walk (one warp).  Runs of "nothing happens" symbols (ok0) are handled 32 at a time;
// everything else one symbol at a time.  Produces the per-symbol output descriptors and the end state.
__global__ void __launch_bounds__(32) acq_walk_kernel(AcqParams p, int nsym, long long base, int c0, const float *__restrict__ lambda,
                                                      const float2 *__restrict__ gamma, const float *__restrict__ avg1,
                                                      const signed char *__restrict__ best2, const float *__restrict__ avg2,
                                                      const unsigned char *__restrict__ ok0, AcqState *st, SymOut *__restrict__ out) {
  const int lane = threadIdx.x;
  const int total = p.N + p.cp;
  const double invN = -1.0 / (double)p.N;
  int c = kD - 8;            // window offset of the current symbol (cp_start == c0)
  int d = 0;                 // offset of the previous symbol minus c
  bool spec_valid = true;    // the speculated input of entry (n, c, d) equals the true average (n == 0: avg_first is the truth)
  float avg = st->avg;
  double ph = st->phase, inc = st->phaseinc, pend = st->nextphaseinc;
  int nextpos = st->nextpos;
  int n = 0, n_out = 0, lost_at = -1, fallback = 0;
  int c_run = 0, c_single = 0, c_seq = 0;
  while (n < nsym) {
    // ---------- run of quiet symbols: entry (m, c, 0) valid, best == 8, avg2 == avg1
    if (spec_valid && d == 0) {
      int m = n + lane;
      bool q = m < nsym && ok0[m * kNC + c];
      unsigned bal = __ballot_sync(0xffffffffu, q);
      int run = bal == 0xffffffffu ? 32 : __ffs(~bal) - 1;
      if (run > 0) {
        // all symbols of the run: peak = cp_start (unchanged), nextpos = cp_start - total, eps from gamma at the peak
        int cp_start = c0 - kD + 8 + c;
        int sw = cp_start - total;
        bool sw_ok = sw >= 0 && sw < total;
        bool sw0_ok = nextpos >= 0 && nextpos < total;
        float2 g = (lane < run) ? gamma[(long long)m * kCand + c + 8] : make_float2(1.f, 0.f);
        double e = invN * (double)atan2f(g.y, g.x);           // nextphaseinc produced by symbol m
        double e1 = __shfl_up_sync(0xffffffffu, e, 1), e2 = __shfl_up_sync(0xffffffffu, e, 2);
        double inc_after0 = sw0_ok ? pend : inc;              // increment in force after the first symbol of the run
        double i0, i1;
        int swn;
        if (lane == 0) { i0 = inc; i1 = pend; swn = sw0_ok ? nextpos : total; }
        else {
          i0 = lane == 1 ? inc_after0 : (sw_ok ? e2 : inc_after0);
          i1 = e1;
          swn = sw_ok ? sw : total;
        }
        double adv = lane < run ? (swn < total ? swn * i0 + (total - swn) * i1 : total * i0) : 0.0;
        double incl = adv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          double t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        if (lane < run) {
          SymOut so;
          so.first = base + (long long)m * total + cp_start - p.N + 1;
          so.phase0 = remainder(ph + (incl - adv), 2.0 * M_PI);
          so.inc0 = i0; so.inc1 = i1; so.switch_at = swn;
          out[n_out + lane] = so;
        }
        int lastl = run - 1;
        ph = remainder(ph + __shfl_sync(0xffffffffu, incl, lastl), 2.0 * M_PI);
        double li = (swn < total) ? i1 : i0;
        inc = __shfl_sync(0xffffffffu, li, lastl);
        pend = __shfl_sync(0xffffffffu, e, lastl);
        nextpos = sw;
        avg = avg1[(n + lastl) * kNC + c];  // == avg2 of the last symbol of the run
        n_out += run;
        n += run;
        c_run += run;
        continue;
      }
    }
    // ---------- one symbol
    c_single++;
    int best;
    bool found;
    signed char sp = spec_valid ? best2[(n * kNC + c) * 3 + (d + 1)] : (signed char)-2;
    if (sp != -2) {
      best = sp;
      found = best >= 0;
      avg = avg2[(n * kNC + c) * 3 + (d + 1)];
    } else {
      fallback = 1;
      c_seq++;
      found = peak_detect(lambda + (long long)n * kCand + c, 16, &avg, p.rise, p.fall, p.alpha, &best) > 0;
    }
    int cp_start = c0 - kD + 8 + c;
    SymOut so;
    so.phase0 = ph; so.inc0 = inc; so.inc1 = inc; so.switch_at = total;
    if (!found) {
      ph = remainder(ph + total * inc, 2.0 * M_PI);  // :335-343
      lost_at = n;
      break;
    }
    if (nextpos >= 0 && nextpos < total) {           // :287-288
      so.inc1 = pend; so.switch_at = nextpos;
      ph += nextpos * inc;
      inc = pend;
      ph += (total - nextpos) * inc;
    } else {
      ph += total * inc;
    }
    ph = remainder(ph, 2.0 * M_PI);
    int peak = best + cp_start - 8;
    float2 g = gamma[(long long)n * kCand + c + best];
    pend = invN * (double)atan2f(g.y, g.x);          // :311
    nextpos = peak - total;                          // :312
    so.first = base + (long long)n * total + peak - p.N + 1;
    if (lane == 0) out[n_out] = so;
    n_out++;
    // next symbol's table entry: offset c' = c + best - 8, came from offset c (d' = c - c')
    int cn = c + best - 8;
    int dn = c - cn;
    // its speculated input is avg1[n][c]; valid iff that is the true average now
    bool next_valid = (dn >= -1 && dn <= 1) && __float_as_uint(avg1[n * kNC + c]) == __float_as_uint(avg);
    n++;
    if (cn < 0 || cn >= kNC) { c = cn; lost_at = n; break; }  // left the table: the host re-centres it
    c = cn; d = dn; spec_valid = next_valid;
  }
  if (lane == 0) {
    st->avg = avg;
    st->phase = (float)ph;
    st->phaseinc = inc;
    st->nextphaseinc = pend;
    st->nextpos = nextpos;
    st->cp_start = c0 - kD + 8 + c;
    st->n_out = n_out;
    st->lost_at = lost_at;
    st->fallback = fallback;
    st->n_run += c_run; st->n_single += c_single; st->n_seq += c_seq;
    st->consumed = (long long)(lost_at >= 0 ? lost_at : nsym) * total;
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
      if ((rc = h->d_avg1.reserve((size_t)nsym * kNC * 4)) || (rc = h->d_avg2.reserve((size_t)nsym * kNC * 3 * 4)) ||
          (rc = h->d_peak.reserve((size_t)nsym * kNC * 3)) || (rc = h->d_flag.reserve((size_t)nsym * kNC)))
        return rc;
      long long t1 = nsym * kNC, t2 = t1 * 3;
      acq_pass1_kernel<<<(unsigned)((t1 + 127) / 128), 128, 0, st>>>(p, (int)nsym, h->d_lambda.as<float>(), h->d_avg1.as<float>());
      acq_pass2_kernel<<<(unsigned)((t2 + 127) / 128), 128, 0, st>>>(p, (int)nsym, h->d_lambda.as<float>(), h->d_avg1.as<float>(), hs->avg,
                                                                    h->d_peak.as<signed char>(), h->d_avg2.as<float>(),
                                                                    h->d_flag.as<unsigned char>());
      acq_walk_kernel<<<1, 32, 0, st>>>(p, (int)nsym, pos, c0, h->d_lambda.as<float>(), h->d_gamma.as<float2>(), h->d_avg1.as<float>(),
                                        h->d_peak.as<signed char>(), h->d_avg2.as<float>(), h->d_flag.as<unsigned char>(),
                                        h->d_state.as<AcqState>(), h->d_sym.as<SymOut>());
      count_launch(4);
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
