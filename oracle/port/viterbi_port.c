/* TEST INFRASTRUCTURE ONLY — CPU restatement of the reference Viterbi decoder.
 *
 * Follows, step for step and with the same 8-bit modular arithmetic:
 *   lib/viterbi_decoder_impl.cc:77-170  (constructor: k, n, m, ntraceback, block sizes)
 *   lib/viterbi_decoder_impl.cc:231-293 (depuncture/unpack + step/output cadence)
 *   lib/d_viterbi.c:261-285             (d_viterbi_chunks_init_sse2: zero state, branch table)
 *   lib/d_viterbi.c:461-576             (d_viterbi_butterfly2_sse2: 64-state ACS, two steps)
 *   lib/d_viterbi.c:680-735             (d_viterbi_get_output_sse2: argmax, ring traceback, renormalise)
 * in scalar C (one state at a time instead of 16 per SSE2 register).
 */
#include "dvbt_oracle.h"
#include <stdlib.h>
#include <string.h>

#define POLYA 0x4f /* d_viterbi.c:38 */
#define POLYB 0x6d /* d_viterbi.c:39 */
#define TRACEBACK_MAX 24 /* d_viterbi.c:71 */

static const int RATE_K[5] = {1, 2, 3, 5, 7};
static const int RATE_N[5] = {2, 3, 4, 6, 8};
static const int RATE_NTB[5] = {5, 9, 10, 15, 24}; /* viterbi_decoder_impl.cc:95-124 */
/* viterbi_decoder_impl.cc:61-65: 1 = transmitted, order X1 Y1 X2 Y2 ... */
static const unsigned char PUNCT[5][14] = {
    {1, 1},
    {1, 1, 0, 1},
    {1, 1, 0, 1, 1, 0},
    {1, 1, 0, 1, 1, 0, 0, 1, 1, 0},
    {1, 1, 0, 1, 0, 1, 0, 1, 1, 0, 0, 1, 1, 0}};

static int parity8(unsigned v) { /* lib/d_tab.c:24-57 (d_Partab) */
  v ^= v >> 4; v ^= v >> 2; v ^= v >> 1;
  return v & 1;
}

struct dvbt_oracle_viterbi {
  int k, n, m, ntb, bsize;
  int nsymbols; /* input bytes per block: bsize*n/m          (viterbi_decoder_impl.cc:149) */
  int nbits;    /* depunctured symbols per block: 2*k*bsize  (:151) */
  int nout;     /* output bytes per block: nbits/2/8         (:153) */
  const unsigned char *punct;
  unsigned char *inbits;
  int init; /* d_init */
  uint8_t branch[2][32];            /* Branchtab27_sse2 */
  uint8_t metric0[64], metric1[64]; /* metric0/metric1 of viterbi_decoder_impl.cc:49-50 */
  uint8_t path0[64], path1[64];     /* path0/path1 :51-52 */
  uint8_t mmresult[64];
  uint8_t ppresult[TRACEBACK_MAX][64];
  int store_pos;
};

/* d_viterbi.c:261-285 */
static void chunks_init(dvbt_oracle_viterbi *v) {
  memset(v->metric0, 0, 64);
  memset(v->path0, 0, 64);
  for (int i = 0; i < 32; i++) {
    v->branch[0][i] = (uint8_t)parity8((2 * i) & POLYA);
    v->branch[1][i] = (uint8_t)parity8((2 * i) & POLYB);
  }
  memset(v->mmresult, 0, 64);
  memset(v->ppresult, 0, sizeof v->ppresult);
  /* store_pos is deliberately NOT reset (d_viterbi.c:77 is a static, :261-285 leaves it) */
}

dvbt_oracle_viterbi *dvbt_oracle_viterbi_create(int m, int rate, int bsize) {
  if (rate < 0 || rate > 4 || (m != 2 && m != 4 && m != 6)) return NULL;
  dvbt_oracle_viterbi *v = (dvbt_oracle_viterbi *)calloc(1, sizeof *v);
  v->k = RATE_K[rate]; v->n = RATE_N[rate]; v->m = m; v->ntb = RATE_NTB[rate];
  v->punct = PUNCT[rate];
  v->bsize = bsize;
  v->nsymbols = bsize * v->n / m;
  v->nbits = 2 * v->k * bsize;
  v->nout = v->nbits / 2 / 8;
  v->inbits = (unsigned char *)malloc((size_t)v->nbits + 32);
  v->init = 0;
  v->store_pos = 0;
  chunks_init(v);
  return v;
}

void dvbt_oracle_viterbi_destroy(dvbt_oracle_viterbi *v) {
  if (!v) return;
  free(v->inbits);
  free(v);
}

void dvbt_oracle_viterbi_reset(dvbt_oracle_viterbi *v) { /* viterbi_decoder_impl.cc:217-221 */
  v->init = 0;
  chunks_init(v);
}

int dvbt_oracle_viterbi_in_bytes_per_block(const dvbt_oracle_viterbi *v) { return v->nsymbols; }
int dvbt_oracle_viterbi_out_bytes_per_block(const dvbt_oracle_viterbi *v) { return v->nout; }
int dvbt_oracle_viterbi_ntraceback(const dvbt_oracle_viterbi *v) { return v->ntb; }
void dvbt_oracle_viterbi_metrics(const dvbt_oracle_viterbi *v, uint8_t metrics[64]) { memcpy(metrics, v->metric0, 64); }

/* One trellis step, d_viterbi.c:477-524 (first half) == :534-575 (second half). */
static void acs_step(const dvbt_oracle_viterbi *v, unsigned char sym0, unsigned char sym1,
                     const uint8_t *M, const uint8_t *P, uint8_t *Mn, uint8_t *Pn) {
  for (int i = 0; i < 32; i++) {
    uint8_t metsv, metsvm;
    if (sym0 == 2) { /* :487-491 */
      metsvm = (uint8_t)(v->branch[1][i] ^ sym1);
      metsv = (uint8_t)(1 - metsvm);
    } else if (sym1 == 2) { /* :492-496 */
      metsvm = (uint8_t)(v->branch[0][i] ^ sym0);
      metsv = (uint8_t)(1 - metsvm);
    } else { /* :497-501 */
      metsvm = (uint8_t)((v->branch[0][i] ^ sym0) + (v->branch[1][i] ^ sym1));
      metsv = (uint8_t)(2 - metsvm);
    }
    uint8_t m0 = (uint8_t)(M[i] + metsv);       /* :503 */
    uint8_t m1 = (uint8_t)(M[i + 32] + metsvm); /* :504 */
    uint8_t m2 = (uint8_t)(M[i] + metsvm);      /* :505 */
    uint8_t m3 = (uint8_t)(M[i + 32] + metsv);  /* :506 */
    int d0 = (int8_t)(uint8_t)(m0 - m1) > 0;    /* :508 signed compare of the 8-bit difference */
    int d1 = (int8_t)(uint8_t)(m2 - m3) > 0;    /* :509 */
    uint8_t shift0 = (uint8_t)(P[i] << 1);            /* :513 */
    uint8_t shift1 = (uint8_t)((P[i + 32] << 1) + 1); /* :514-515 */
    Mn[2 * i] = d0 ? m0 : m1;                   /* :510,517,520 (unpacklo/hi = states 2i, 2i+1) */
    Mn[2 * i + 1] = d1 ? m2 : m3;               /* :511 */
    Pn[2 * i] = d0 ? shift0 : shift1;           /* :518,523 */
    Pn[2 * i + 1] = d1 ? shift0 : shift1;       /* :521,524 */
  }
}

/* d_viterbi.c:461-576: two steps, ping-pong metric0 -> metric1 -> metric0 */
static void butterfly2(dvbt_oracle_viterbi *v, const unsigned char *s) {
  acs_step(v, s[0], s[1], v->metric0, v->path0, v->metric1, v->path1);
  acs_step(v, s[2], s[3], v->metric1, v->path1, v->metric0, v->path0);
}

/* d_viterbi.c:680-735 */
static unsigned char get_output(dvbt_oracle_viterbi *v) {
  int ntb = v->ntb;
  v->store_pos = (v->store_pos + 1) % ntb; /* :689 */
  memcpy(v->mmresult, v->metric0, 64);       /* :692-696 */
  memcpy(v->ppresult[v->store_pos], v->path0, 64);
  int beststate = 0;
  int bestmetric = v->mmresult[0], minmetric = v->mmresult[0]; /* :699-700 */
  for (int i = 1; i < 64; i++) { /* :702-711 unsigned, first strict maximum */
    if (v->mmresult[i] > bestmetric) { bestmetric = v->mmresult[i]; beststate = i; }
    if (v->mmresult[i] < minmetric) minmetric = v->mmresult[i];
  }
  int pos = v->store_pos;
  for (int i = 0; i < ntb - 1; i++) { /* :714-721 */
    beststate = v->ppresult[pos][beststate] >> 2;
    pos = (pos - 1 + ntb) % ntb;
  }
  unsigned char out = v->ppresult[pos][beststate]; /* :724 */
  for (int i = 0; i < 64; i++) {                   /* :728-732 */
    v->path0[i] = 0;
    v->metric0[i] = (uint8_t)(v->metric0[i] - (uint8_t)minmetric);
  }
  return out;
}

long dvbt_oracle_viterbi_work(dvbt_oracle_viterbi *v, const uint8_t *in, long nblocks, uint8_t *out) {
  long out_count = 0;
  int period = 2 * v->k;
  for (long nb = 0; nb < nblocks; nb++) {
    /* viterbi_decoder_impl.cc:241-256 */
    int count = 0;
    for (int i = 0; i < v->nsymbols; i++) {
      for (int j = v->m - 1; j >= 0; j--) {
        while (v->punct[count % period] == 0) v->inbits[count++] = 2;
        v->inbits[count++] = (in[nb * v->nsymbols + i] >> j) & 1;
        while (v->punct[count % period] == 0) v->inbits[count++] = 2;
      }
    }
    /* :261-292 */
    for (int in_count = 0; in_count < v->nbits; in_count++) {
      if ((in_count % 4) == 0) {
        butterfly2(v, &v->inbits[in_count & ~3]);
        if (in_count > 0 && (in_count % 16) == 8) {
          unsigned char c = get_output(v);
          if (v->init == 0) {
            if (out_count >= v->ntb) out[out_count - v->ntb] = c; /* :277-281 */
          } else {
            out[out_count] = c; /* :285 */
          }
          out_count++;
        }
      }
    }
  }
  long to_out = nblocks * v->nout; /* noutput_items */
  if (v->init == 0) {            /* :298-312 */
    to_out -= v->ntb;
    v->init = 1;
  }
  return to_out;
}

long dvbt_oracle_conv_encode(const uint8_t *data, long nbytes, int m, int rate, uint8_t *out) {
  if (rate < 0 || rate > 4) return -1;
  int k = RATE_K[rate];
  const unsigned char *p = PUNCT[rate];
  unsigned char encstate = 0; /* d_viterbi.c:106-124 */
  long nout = 0;
  unsigned acc = 0;
  int nacc = 0, phase = 0;
  for (long b = 0; b < nbytes; b++) {
    for (int i = 7; i >= 0; i--) {
      encstate = (unsigned char)((encstate << 1) | ((data[b] >> i) & 1));
      int sym[2] = {parity8(encstate & POLYA), parity8(encstate & POLYB)};
      for (int s = 0; s < 2; s++) {
        if (p[2 * phase + s]) {
          acc = (acc << 1) | (unsigned)sym[s];
          if (++nacc == m) { out[nout++] = (uint8_t)acc; acc = 0; nacc = 0; }
        }
      }
      phase = (phase + 1) % k;
    }
  }
  return nout;
}

/* ---- soft-decision generalisation (NOT in the reference: lib/d_metrics.c:57-74 is a stub, TODO.txt:25) ----------------
 * The same decoder - trellis and ACS tie rule of d_viterbi.c:477-524, output cadence of viterbi_decoder_impl.cc:261-292,
 * argmax / ring traceback / renormalisation of d_viterbi.c:680-735 - with the agreement count of a branch replaced by
 *     w(c0, v0) + w(c1, v1),   w(1, v) = max(v, 0),  w(0, v) = max(-v, 0),
 * v the signed soft value of a received code bit (clamped to +-6; 0 for a punctured position).  v = +-1 gives exactly
 * metsv / metsvm of :487-501, so dvbt_oracle_viterbi_soft on +-1 values equals dvbt_oracle_viterbi_work on the hard bits
 * (checked in tests/test_soft_decision_cpu.py): that is the pin of this function; beyond it, it is the definition of the
 * mode ("parity unpinned": the reference has no soft path to compare with). */
static void acs_step_soft(const dvbt_oracle_viterbi *v, int v0, int v1, const uint8_t *M, const uint8_t *P, uint8_t *Mn, uint8_t *Pn) {
  for (int i = 0; i < 32; i++) {
    int c0 = v->branch[0][i], c1 = v->branch[1][i];
    int w_same = (c0 ? (v0 > 0 ? v0 : 0) : (v0 < 0 ? -v0 : 0)) + (c1 ? (v1 > 0 ? v1 : 0) : (v1 < 0 ? -v1 : 0));
    int w_inv = (c0 ? (v0 < 0 ? -v0 : 0) : (v0 > 0 ? v0 : 0)) + (c1 ? (v1 < 0 ? -v1 : 0) : (v1 > 0 ? v1 : 0));
    uint8_t metsv = (uint8_t)w_same, metsvm = (uint8_t)w_inv;
    uint8_t m0 = (uint8_t)(M[i] + metsv), m1 = (uint8_t)(M[i + 32] + metsvm);
    uint8_t m2 = (uint8_t)(M[i] + metsvm), m3 = (uint8_t)(M[i + 32] + metsv);
    int d0 = (int8_t)(uint8_t)(m0 - m1) > 0, d1 = (int8_t)(uint8_t)(m2 - m3) > 0;
    uint8_t shift0 = (uint8_t)(P[i] << 1), shift1 = (uint8_t)((P[i + 32] << 1) + 1);
    Mn[2 * i] = d0 ? m0 : m1;
    Mn[2 * i + 1] = d1 ? m2 : m3;
    Pn[2 * i] = d0 ? shift0 : shift1;
    Pn[2 * i + 1] = d1 ? shift0 : shift1;
  }
}

/* One stream from a reset: `in` = one int8 per transmitted code bit (order X1 Y1 X2 ..., punctured positions absent);
 * n_in * k must be a multiple of 8 n.  Writes n_in*k/(8n) - ntraceback bytes; returns that count. */
long dvbt_oracle_viterbi_soft(const int8_t *in, long n_in, int rate, uint8_t *out) {
  if (rate < 0 || rate > 4) return -1;
  dvbt_oracle_viterbi *v = dvbt_oracle_viterbi_create(2, rate, 768);
  if (!v) return -1;
  const int k = v->k, n = v->n, period = 2 * k;
  if (((long long)n_in * k) % (8LL * n) != 0) { dvbt_oracle_viterbi_destroy(v); return -1; }
  const long nsteps = (long)((long long)n_in * k / n);
  long idx = 0, nout = 0, nbt = 0;
  int ph = 0;
  for (long t = 0; t < nsteps; t += 2) {
    int sv[4];
    for (int q = 0; q < 4; q++) {
      int val = 0;
      if (v->punct[ph]) { val = in[idx++]; if (val > 6) val = 6; if (val < -6) val = -6; }
      sv[q] = val;
      ph = (ph + 1) % period;
    }
    acs_step_soft(v, sv[0], sv[1], v->metric0, v->path0, v->metric1, v->path1);
    acs_step_soft(v, sv[2], sv[3], v->metric1, v->path1, v->metric0, v->path0);
    if ((t % 8) == 4) { /* after the 6th step of a byte time: in_count % 16 == 8 of viterbi_decoder_impl.cc:270 */
      unsigned char c = get_output(v);
      if (nbt >= v->ntb) out[nout++] = c;
      nbt++;
    }
  }
  dvbt_oracle_viterbi_destroy(v);
  return nout;
}
