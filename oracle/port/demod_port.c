/* TEST INFRASTRUCTURE ONLY — CPU restatement of demod_reference_signals / pilot_gen (RX half).
 *   lib/reference_signals_impl.cc: tables :54-126, generate_prbs :333-345, verify_bch_code :384-425,
 *     process_spilot_data :535-689, process_cpilot_data :714-744, compute_oneshot_csft :746-790,
 *     frequency_correction :792-819, process_tps_data :918-1032, process_payload_data :1064-1124,
 *     parse_input :1188-1248
 *   lib/demod_reference_signals_impl.cc:96-150 (block level gating and tags)
 * Complex arithmetic uses C99 `float complex`, which gcc lowers exactly like the reference's
 * std::complex<float> (same libgcc __mulsc3/__divsc3), so operand order and rounding are the reference's.
 */
#define _GNU_SOURCE
#include "dvbt_oracle.h"
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef float complex cf;

static const int CP2K[45] = {0,   48,  54,  87,  141, 156, 192, 201, 255,  279,  282,  333,  432,  450,  483,
                             525, 531, 618, 636, 714, 759, 765, 780, 804,  873,  888,  918,  939,  942,  969,
                             984, 1050, 1101, 1107, 1110, 1137, 1140, 1146, 1206, 1269, 1323, 1377, 1491, 1683, 1704};
static const int TPS2K[17] = {34, 50, 209, 346, 413, 569, 595, 688, 790, 901, 1073, 1219, 1262, 1286, 1469, 1594, 1687};
static const int SYNC_EVEN[16] = {0, 0, 1, 1, 0, 1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 0};

struct dvbt_oracle_demod {
  int N, P, K, zl, cp, ncp, ntps, spsize, fi_start;
  int *cpilot, *tpsc;
  char *wk;
  float *known;
  cf *gain, *derot, *prev_tps;
  int *chanestim, *payload;
  int freq_offset;
  float carrier_corr;
  int symbol_index, symbol_index_known, frame_index, mod_symbol_index, prev_mod_symbol_index;
  unsigned char fifo[68];
  int d_init;
};

static cf pilot_value(const dvbt_oracle_demod *d, int k) { return (float)(4 * 2 * (0.5 - d->wk[k]) / 3) + 0.0f * I; } /* :474-479 */

dvbt_oracle_demod *dvbt_oracle_demod_create(int constellation, int tm) {
  dvbt_oracle_demod *d = (dvbt_oracle_demod *)calloc(1, sizeof *d);
  d->N = tm == 0 ? 2048 : 8192;
  d->P = tm == 0 ? 1512 : 6048;
  d->K = tm == 0 ? 1705 : 6817;
  d->zl = (int)ceil((d->N - d->K) / 2.0);
  d->cp = d->N / 32;
  d->spsize = tm == 0 ? 142 : 568;
  d->fi_start = (constellation == 2 && tm == 1) ? 2 : 3; /* demod_reference_signals_impl.cc:73-77 */
  int reps = tm == 0 ? 1 : 4;
  d->cpilot = (int *)malloc(sizeof(int) * 45 * reps);
  d->ncp = 0;
  for (int r = 0; r < reps; r++)
    for (int i = 0; i < 45; i++) {
      int v = CP2K[i] + 1704 * r;
      if (d->ncp == 0 || d->cpilot[d->ncp - 1] != v) d->cpilot[d->ncp++] = v;
    }
  d->ntps = 17 * reps;
  d->tpsc = (int *)malloc(sizeof(int) * d->ntps);
  for (int r = 0; r < reps; r++)
    for (int i = 0; i < 17; i++) d->tpsc[r * 17 + i] = TPS2K[i] + 1704 * r;
  d->wk = (char *)malloc(d->K);
  unsigned reg = (1u << 11) - 1; /* :333-345 */
  for (int k = 0; k < d->K; k++) {
    d->wk[k] = (char)(reg & 1);
    int nb = ((reg >> 2) ^ reg) & 1;
    reg = (reg >> 1) | (nb << 10);
  }
  d->known = (float *)malloc(sizeof(float) * (d->ncp - 1));
  for (int i = 0; i < d->ncp - 1; i++) {
    cf df = pilot_value(d, d->cpilot[i + 1]) - pilot_value(d, d->cpilot[i]);
    d->known[i] = crealf(df) * crealf(df) + cimagf(df) * cimagf(df); /* norm(), :224-228 */
  }
  d->gain = (cf *)calloc(d->K, sizeof(cf));
  d->derot = (cf *)calloc(d->N, sizeof(cf));
  d->prev_tps = (cf *)calloc(d->ntps, sizeof(cf));
  d->chanestim = (int *)malloc(sizeof(int) * d->K);
  d->payload = (int *)malloc(sizeof(int) * d->K);
  return d;
}

void dvbt_oracle_demod_destroy(dvbt_oracle_demod *d) {
  if (!d) return;
  free(d->cpilot); free(d->tpsc); free(d->wk); free(d->known); free(d->gain); free(d->derot); free(d->prev_tps);
  free(d->chanestim); free(d->payload); free(d);
}

static float cnorm(cf z) { return crealf(z) * crealf(z) + cimagf(z) * cimagf(z); }

static int bch_ok(const unsigned char *f) { /* :384-425 */
  unsigned reg = 0;
  for (int i = 0; i < 113; i++) {
    unsigned bit = i < 60 ? 0u : f[1 + (i - 60)];
    unsigned fb = 1u & (bit ^ reg);
    reg >>= 1;
    reg |= fb << 13;
    reg ^= (fb << 12) ^ (fb << 11) ^ (fb << 9) ^ (fb << 8) ^ (fb << 7) ^ (fb << 5) ^ (fb << 4);
  }
  for (int i = 0; i < 14; i++)
    if (f[i + 54] != (1u & (reg >> i))) return 0;
  return 1;
}

/* scattered pilot k of phase r: r == 0 has one more (Kmax) (:445-468) */
static int is_spilot(const dvbt_oracle_demod *d, int k, int r) { return k >= 3 * r && (k - 3 * r) % 12 == 0; }

/* parse_input (:1188-1248): in = this symbol, in + N = the next one.  Returns 1. */
static void parse_input(dvbt_oracle_demod *d, const cf *in, cf *out, int *symbol_index, int *frame_index) {
  const int zl = d->zl, N = d->N;
  /* process_cpilot_data :714-744 */
  float max = 0;
  int start = 0;
  for (int i = zl - 8; i < zl + 8; i++) {
    float sum = 0;
    for (int j = 0; j < d->ncp - 1; j++) {
      float phase = cnorm(in[i + d->cpilot[j + 1]] - in[i + d->cpilot[j]]);
      sum += d->known[j] * phase;
    }
    if (sum > max) { max = sum; start = i; }
  }
  d->freq_offset = max > 0 ? start - zl : 0; /* all-zero input: the reference reads out of bounds */
  /* compute_oneshot_csft :746-790 */
  {
    cf left = 0.0f, right = 0.0f;
    int half = (d->ncp - 1) / 2;
    float carrier_coeff = 1.0 / (2 * M_PI * (1 + (float)d->cp / (float)N) * 2);
    for (int j = 0; j < half; j++) {
      int idx = d->freq_offset + zl + d->cpilot[j];
      left += in[idx] * conjf(in[idx + N]);
    }
    for (int j = half + 1; j < d->ncp; j++) {
      int idx = d->freq_offset + zl + d->cpilot[j];
      right += in[idx] * conjf(in[idx + N]);
    }
    float la = cargf(left), ra = cargf(right);
    d->carrier_corr = (ra + la) * carrier_coeff;
  }
  /* frequency_correction :792-819 */
  {
    float correction = (float)d->freq_offset + d->carrier_corr;
    float ang = (float)(-2 * M_PI * correction * (N + d->cp) / N * 1);
    float sn, cs;
    sincosf(ang, &sn, &cs);
    cf c = cs + sn * I;
    for (int k = 0; k < N; k++) {
      int src = k + d->freq_offset;
      d->derot[k] = (src >= 0 && src < 2 * N) ? c * in[src] : 0;
    }
  }
  const cf *x = d->derot;
  /* process_spilot_data :535-689 */
  {
    float smax = 0;
    for (int sc = 0; sc < 4; sc++) {
      cf c = 0.0f;
      for (int j = 0; j < 10; j++) {
        int k = 3 * sc + 12 * j;
        c += pilot_value(d, k) * conjf(x[zl + k]);
      }
      float sum = cnorm(c);
      if (sum > smax) { smax = sum; d->mod_symbol_index = sc; }
    }
    int r = d->mod_symbol_index, n = 0;
    for (int k = 0; k < d->K; k++) { /* :594-614: scattered first, then continual (duplicates kept) */
      if (is_spilot(d, k, r)) d->chanestim[n++] = k;
      int isc = 0;
      for (int j = 0; j < d->ncp; j++)
        if (d->cpilot[j] == k) isc = 1;
      if (isc) d->chanestim[n++] = k;
    }
    int startk = d->chanestim[0];
    for (int i = 0; i < n; i++) { /* :617-642 */
      int k = d->chanestim[i];
      d->gain[k] = pilot_value(d, k) / x[k + zl];
      cf tg = (d->gain[k] - d->gain[startk]) / (11.0f + 0.0f * I);
      for (int j = 1; j < k - startk; j++) d->gain[startk + j] = d->gain[startk] + tg * ((float)j + 0.0f * I);
      startk = k;
    }
  }
  int diff = (d->mod_symbol_index - d->prev_mod_symbol_index + 4) % 4;
  d->prev_mod_symbol_index = d->mod_symbol_index;
  d->symbol_index = (d->symbol_index + diff) % 68; /* :1228 */
  *symbol_index = d->symbol_index;
  *frame_index = d->frame_index;
  /* process_tps_data :918-1032 */
  int end_frame = 0;
  {
    int vote = 0;
    for (int k = 0; k < d->ntps; k++) {
      cf val = x[zl + d->tpsc[k]] * d->gain[d->tpsc[k]];
      if (!d->symbol_index_known || d->symbol_index != 0) {
        cf ph = val * conjf(d->prev_tps[k]);
        if (crealf(ph) >= 0.0) vote++; else vote--;
      }
      d->prev_tps[k] = val;
    }
    for (int i = 0; i < diff; i++) {
      memmove(d->fifo, d->fifo + 1, 67);
      if (!d->symbol_index_known || d->symbol_index != 0) d->fifo[67] = vote >= 0 ? 0 : 1;
      else d->fifo[67] = 0;
    }
    int even = 1, odd = 1;
    for (int i = 0; i < 15; i++) { /* std::equal over 15 elements, :975/:1002 */
      if (d->fifo[1 + i] != SYNC_EVEN[i]) even = 0;
      if (d->fifo[1 + i] != 1 - SYNC_EVEN[i]) odd = 0;
    }
    if (even || odd) {
      if (bch_ok(d->fifo)) {
        d->frame_index = (d->fifo[23] << 1) | d->fifo[24];
        d->symbol_index_known = 1;
        end_frame = 1;
      } else {
        d->symbol_index_known = 0;
      }
      memset(d->fifo, 0, 68);
    }
  }
  if (end_frame) d->symbol_index = 67; /* :1240-1241 */
  /* process_payload_data :1064-1124 */
  {
    int r = d->mod_symbol_index, n = 0;
    for (int k = 0; k < d->K; k++) {
      int pay = !is_spilot(d, k, r);
      for (int j = 0; j < d->ncp && pay; j++)
        if (d->cpilot[j] == k) pay = 0;
      for (int j = 0; j < d->ntps && pay; j++)
        if (d->tpsc[j] == k) pay = 0;
      if (pay) d->payload[n++] = k;
    }
    for (int i = 0; i < n && i < d->P; i++) out[i] = x[zl + d->payload[i]] * d->gain[d->payload[i]];
  }
}

/* The block, one item per call as the reference is driven (demod_reference_signals_impl.cc:96-150):
 * in holds nsym symbols of N cells, nsym-1 are parsed.  sync_start_at0: a sync_start tag sits on item 0.
 * out receives P cells per produced item; symbol_index_out[i] is the symbol_index tag of produced item i;
 * *superframe_tag_at = produced-item index that carries the superframe_start tag (or -1).
 * Returns the number of produced items. */
long dvbt_oracle_demod_run(dvbt_oracle_demod *d, const float *in_re_im, long nsym, int sync_start_at0, float *out_re_im,
                           int *symbol_index_out, long *superframe_tag_at) {
  const cf *in = (const cf *)in_re_im;
  cf *out = (cf *)out_re_im;
  long nout = 0;
  *superframe_tag_at = -1;
  cf *tmp = (cf *)malloc(sizeof(cf) * d->P);
  for (long s = 0; s + 1 < nsym; s++) {
    int si, fi;
    parse_input(d, in + s * d->N, tmp, &si, &fi);
    if (s == 0 && sync_start_at0) d->d_init = 0; /* :115-116 */
    if (d->d_init == 0) {
      if ((si % 68) == 0 && (fi % 4) == d->fi_start) { d->d_init = 1; *superframe_tag_at = nout; }
      else continue;
    }
    memcpy(out + nout * d->P, tmp, sizeof(cf) * d->P);
    symbol_index_out[nout] = si;
    nout++;
  }
  free(tmp);
  return nout;
}
