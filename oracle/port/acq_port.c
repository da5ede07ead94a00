/* TEST INFRASTRUCTURE ONLY — CPU restatement of gr::dvbt::ofdm_sym_acquisition.
 *   lib/ofdm_sym_acquisition_impl.cc: peak_detect_process :72-146, ml_sync :148-351,
 *   constructor :379-449 (rho, detector parameters), general_work :488-568.
 * VOLK kernels are restated from their generic definitions, gr_expj as (cosf, sinf),
 * gr::fast_atan2f as atan2f (un-vendored third-party code: parity unpinned at those calls).
 * Out-of-bounds accesses of the reference are avoided by indexing the caller's
 * buffer with absolute positions; the values computed are the same.
 */
#define _GNU_SOURCE
#include "dvbt_oracle.h"
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef float complex cf;

struct dvbt_oracle_acq {
  int N, cp;
  float rho;
  float rise, fall, alpha, avg; /* peak detector, :448 */
  float phase;
  double phaseinc, nextphaseinc;
  int nextpos;
  int initial, cp_start, freq_correction_count;
  cf *derot, *gamma;
  float *lambda, *phi;
  int *peak_pos;
};

dvbt_oracle_acq *dvbt_oracle_acq_create(int fft_length, int cp_length, float snr_db) {
  dvbt_oracle_acq *a = (dvbt_oracle_acq *)calloc(1, sizeof *a);
  a->N = fft_length;
  a->cp = cp_length;
  float snr = pow(10, snr_db / 10.0); /* :390 */
  a->rho = snr / (snr + 1.0);         /* :391 */
  a->rise = 0.8f; a->fall = 0.9f; a->alpha = 0.9f; a->avg = 0;
  a->derot = (cf *)calloc(fft_length + cp_length, sizeof(cf));
  a->gamma = (cf *)calloc(fft_length, sizeof(cf));
  a->lambda = (float *)calloc(fft_length, sizeof(float));
  a->phi = (float *)calloc(fft_length, sizeof(float));
  a->peak_pos = (int *)calloc(fft_length, sizeof(int));
  return a;
}

void dvbt_oracle_acq_destroy(dvbt_oracle_acq *a) {
  if (!a) return;
  free(a->derot); free(a->gamma); free(a->lambda); free(a->phi); free(a->peak_pos); free(a);
}

/* :72-146 */
static int peak_detect(dvbt_oracle_acq *a, const float *d, int n, int *peak_pos, int *peak_max) {
  int state = 0, peak_index = 0, npeaks = 0, i = 0;
  float peak_val = -(float)INFINITY;
  while (i < n) {
    if (state == 0) {
      if (d[i] > a->avg * a->rise) state = 1;
      else { a->avg = a->alpha * d[i] + (1 - a->alpha) * a->avg; i++; }
    } else {
      if (d[i] > peak_val) { peak_val = d[i]; peak_index = i; a->avg = a->alpha * d[i] + (1 - a->alpha) * a->avg; i++; }
      else if (d[i] > a->avg * a->fall) { a->avg = a->alpha * d[i] + (1 - a->alpha) * a->avg; i++; }
      else { peak_pos[npeaks++] = peak_index; state = 0; peak_val = -(float)INFINITY; }
    }
  }
  if (npeaks) {
    float max = d[peak_pos[0]];
    int maxi = 0;
    for (int k = 1; k < npeaks; k++)
      if (d[peak_pos[k]] > max) { max = d[peak_pos[k]]; maxi = k; }
    *peak_max = maxi;
  }
  return npeaks;
}

/* ml_sync :148-351 on the window starting at in[0]; returns number of peaks */
static int ml_sync(dvbt_oracle_acq *a, const cf *in, int lookup_start, int lookup_stop, int *cp_pos) {
  const int N = a->N, cp = a->cp;
  int size = lookup_start - lookup_stop;
  for (int i = lookup_start - 1; i >= lookup_stop; i--) { /* :196-211 */
    int k = i - lookup_stop;
    a->phi[k] = 0.0;
    a->gamma[k] = 0.0;
    for (int j = 0; j < cp; j++) {
      cf x1 = in[i - j], x0 = in[i - j - N];
      a->gamma[k] += x1 * conjf(x0);                                         /* d_corr, :184 */
      float n1 = crealf(x1) * crealf(x1) + cimagf(x1) * cimagf(x1);          /* d_norm, :168 */
      float n0 = crealf(x0) * crealf(x0) + cimagf(x0) * cimagf(x0);
      a->phi[k] += n1 + n0;
    }
  }
  float rho2 = a->rho / 2.0; /* :236: scalar argument of volk_32f_s32f_multiply_32f is a float */
  for (int k = 0; k < size; k++) {
    float mag = sqrtf(crealf(a->gamma[k]) * crealf(a->gamma[k]) + cimagf(a->gamma[k]) * cimagf(a->gamma[k])); /* :219 */
    float p2 = a->phi[k] * rho2;
    a->lambda[k] = mag - p2; /* :237 */
  }
  int peak_max = 0;
  int npeaks = peak_detect(a, a->lambda, size, a->peak_pos, &peak_max);
  if (npeaks) {
    int peak = a->peak_pos[peak_max] + lookup_stop;
    *cp_pos = peak;
    float eps = atan2f(cimagf(a->gamma[a->peak_pos[peak_max]]), crealf(a->gamma[a->peak_pos[peak_max]])); /* :277 */
    double sensitivity = (double)(-1) / (double)N;
    for (int i = 0; i < cp + N; i++) { /* :285-309 */
      if (i == a->nextpos) a->phaseinc = a->nextphaseinc;
      a->phase += a->phaseinc;
      while (a->phase > (float)M_PI) a->phase -= (float)(2.0 * M_PI);
      while (a->phase < (float)(-M_PI)) a->phase += (float)(2.0 * M_PI);
      float sn, cs;
      sincosf(a->phase, &sn, &cs);
      a->derot[i] = cs + sn * I;
    }
    a->nextphaseinc = sensitivity * eps;   /* :311 */
    a->nextpos = peak - (cp + N);          /* :312 */
  } else {
    for (int i = 0; i < cp + N; i++) { /* :335-343 */
      a->phase += a->phaseinc;
      while (a->phase > (float)M_PI) a->phase -= (float)(2.0 * M_PI);
      while (a->phase < (float)(-M_PI)) a->phase += (float)(2.0 * M_PI);
    }
  }
  return npeaks;
}

/* general_work (:488-568) called repeatedly over x[0..n): returns symbols produced; *consumed = samples.
 * first_sync_tag: set to 1 if a sync_start tag was emitted on the first produced item. */
long dvbt_oracle_acq_run(dvbt_oracle_acq *a, const float *x_re_im, long n, float *out_re_im, long out_capacity, long *consumed,
                         int *first_sync_tag) {
  const cf *x = (const cf *)x_re_im;
  cf *out = (cf *)out_re_im;
  const int N = a->N, cp = a->cp;
  long pos = 0, nout = 0;
  int tagged = 0, sent = 0;
  while (pos + 2 * N + cp + 16 <= n && nout < out_capacity) {
    const cf *in = x + pos;
    int to_consume = cp + N, to_out = 0;
    if (!a->initial) {
      a->initial = ml_sync(a, in, 2 * N + cp - 1, N + cp - 1, &a->cp_start); /* :501-503 */
      if (nout == 0 && !sent) tagged = 1;                                   /* send_sync_start(), :507 */
      sent = 1;
    }
    if (a->initial) {
      int found = ml_sync(a, in, a->cp_start + 8, a->cp_start - 8, &a->cp_start); /* :514-515 */
      if (found) {
        a->freq_correction_count = 0;
        for (int j = 0; j < N; j++) out[nout * N + j] = a->derot[j] * in[a->cp_start - N + 1 + j]; /* :526-535 */
        to_out = 1;
      } else if (++a->freq_correction_count > 0) { /* d_freq_correction_timeout = 0, :548 */
        a->initial = 0;
        a->freq_correction_count = 0;
        to_consume = to_consume / 2; /* :557 */
      }
    }
    pos += to_consume;
    nout += to_out;
  }
  *consumed = pos;
  if (first_sync_tag) *first_sync_tag = tagged;
  return nout;
}
