/* TEST INFRASTRUCTURE ONLY — CPU restatement of the reference Reed-Solomon decoder.
 *
 * Follows lib/reed_solomon.cc:
 *   gf_init   :48-89   (GF(2^8), poly 0x11d, exp/log tables, exp[255] = 0, log[0] = 255)
 *   gf_mul/div/pow/exp :98-148
 *   rs_decode :246-489 (syndromes by Horner over n = 255 symbols, errors-only Berlekamp-Massey
 *                       as the block calls it with no_eras = 0, Chien search, Forney)
 * and lib/reed_solomon_dec_impl.cc:77-116 (51 zero symbols prepended, 204 -> 188 bytes,
 * return value of rs_decode ignored).
 *
 * `as_built` selects how the reference's out-of-bounds store `omega[2*d_t] = 0`
 * (reed_solomon.cc:434, array declared with 2*d_t elements at :255) is modelled:
 *   0: as the source intends (omega has room; nothing else is touched) — equals the
 *      reference built with omega[2*d_t+1] (oracle/_ref/libdvbt_ref_rsfix.so);
 *   1: as gcc 13.3 -O3 lays the VLAs out in this container: the store lands on loc[0], so the
 *      first (lowest-position) error is "corrected" at data[0] — a symbol of the zero prefix
 *      that the block discards — and stays wrong in the payload.
 */
#include "dvbt_oracle.h"
#include <string.h>

#define RS_N 255
#define RS_K 239
#define RS_T 8
#define RS_S 51

static unsigned char gf_exp_t[256], gf_log_t[256];
static int gf_ready = 0;

static void gf_init(void) { /* reed_solomon.cc:48-89 with p=2, m=8, gfpoly=0x11d */
  int reg = 1;
  gf_exp_t[255] = 0;
  gf_log_t[0] = 255;
  for (int i = 0; i < 255; i++) {
    gf_exp_t[i] = (unsigned char)reg;
    gf_log_t[reg] = (unsigned char)i;
    reg <<= 1;
    if (reg & 0x100) reg ^= 0x11d;
    reg &= 0xff;
  }
  gf_ready = 1;
}
static int gf_exp(int a) { return gf_exp_t[a % RS_N]; }                 /* :98-102 */
static int gf_mul(int a, int b) { return (a == 0 || b == 0) ? 0 : gf_exp(gf_log_t[a] + gf_log_t[b]); } /* :118-125 */
static int gf_div(int a, int b) { return (a == 0 || b == 0) ? 0 : gf_exp(RS_N + gf_log_t[a] - gf_log_t[b]); } /* :127-134 */
static int gf_pow(int a, int p) { return a == 0 ? 0 : gf_exp(RS_N + gf_log_t[a] + p); } /* :136-143 */

/* reed_solomon.cc:246-489 with eras = NULL, no_eras = 0.  data has RS_N symbols. */
static int rs_decode(unsigned char *data, int as_built) {
  unsigned char sigma[2 * RS_T + 1], b[2 * RS_T + 1], T[2 * RS_T + 1], reg[2 * RS_T + 1];
  unsigned char root[2 * RS_T + 1], loc[2 * RS_T + 1], omega[2 * RS_T + 1], syn[2 * RS_T + 1];
  memset(sigma, 0, sizeof sigma);
  sigma[0] = 1;
  for (int j = 0; j < 2 * RS_T; j++) syn[j] = data[0];               /* :281-282 */
  for (int j = 1; j < RS_N; j++)                                      /* :284-288 */
    for (int i = 0; i < 2 * RS_T; i++) syn[i] = (unsigned char)(data[j] ^ gf_pow(syn[i], i));
  int syn_error = 0;
  for (int i = 0; i < 2 * RS_T; i++) syn_error |= syn[i];
  if (!syn_error) return 0;                                           /* :299-305 */
  int r = 0, el = 0;                                                  /* :310-311 */
  memcpy(b, sigma, sizeof b);
  while (++r <= 2 * RS_T) {                                           /* :315-354 */
    int discr = 0;
    for (int i = 0; i < r; i++) discr ^= gf_mul(sigma[i], syn[r - i - 1]);
    if (discr == 0) {
      memmove(&b[1], b, 2 * RS_T);
      b[0] = 0;
    } else {
      T[0] = sigma[0];
      for (int i = 0; i < 2 * RS_T; i++) T[i + 1] = (unsigned char)(sigma[i + 1] ^ gf_mul(discr, b[i]));
      if (2 * el <= r - 1) {
        el = r - el;
        for (int i = 0; i <= 2 * RS_T; i++) b[i] = (unsigned char)gf_div(sigma[i], discr);
      } else {
        memmove(&b[1], b, 2 * RS_T);
        b[0] = 0;
      }
      memcpy(sigma, T, sizeof sigma);
    }
  }
  int deg_sigma = 0;                                                  /* :357-364 */
  for (int i = 0; i < 2 * RS_T + 1; i++)
    if (sigma[i] != 0) deg_sigma = i;
  int no_roots = 0;                                                   /* :376-403 */
  memcpy(&reg[1], &sigma[1], 2 * RS_T);
  for (int i = 1; i <= RS_N; i++) {
    int q = 1;
    for (int j = deg_sigma; j > 0; j--) {
      reg[j] = (unsigned char)gf_pow(reg[j], j);
      q ^= reg[j];
    }
    if (q != 0) continue;
    root[no_roots] = (unsigned char)i;
    loc[no_roots] = (unsigned char)(i - 1);
    if (++no_roots == deg_sigma) break;
  }
  if (no_roots != deg_sigma) return -1;                               /* :405-415 */
  int deg_omega = 0;                                                  /* :419-433 */
  for (int i = 0; i < 2 * RS_T; i++) {
    int tmp = 0;
    int j = (deg_sigma < i) ? deg_sigma : i;
    for (; j >= 0; j--) tmp ^= gf_mul(syn[i - j], sigma[j]);
    if (tmp != 0) deg_omega = i;
    omega[i] = (unsigned char)tmp;
  }
  omega[2 * RS_T] = 0;                                                /* :434 (in bounds here) */
  if (as_built) loc[0] = 0;                                           /* where :434 lands with gcc 13.3 */
  for (int j = no_roots - 1; j >= 0; j--) {                           /* :445-486 */
    int num1 = 0;
    for (int i = deg_omega; i >= 0; i--) num1 ^= gf_pow(omega[i], i * root[j]);
    int num2 = gf_exp(root[j] * (-1) + RS_N);
    int den = 0;
    int deg_max = deg_sigma < 2 * RS_T - 1 ? deg_sigma : 2 * RS_T - 1;
    for (int i = 1; i <= deg_max; i += 2)
      if (sigma[i] != 0) den ^= gf_exp(gf_log_t[sigma[i]] + (i - 1) * root[j]);
    if (den == 0) return -1;                                          /* :470-479 */
    int err = gf_div(gf_mul(num1, num2), den);
    data[loc[j]] ^= (unsigned char)err;
  }
  return no_roots;
}

/* reed_solomon_dec_impl.cc:77-116 over npackets packets of 204 bytes.
 * status (optional) receives rs_decode's return value per packet. */
void dvbt_oracle_rs_decode(const uint8_t *in, long npackets, uint8_t *out, int as_built, int *status) {
  if (!gf_ready) gf_init();
  unsigned char d[RS_N];
  for (long p = 0; p < npackets; p++) {
    memset(d, 0, RS_S);
    memcpy(d + RS_S, in + p * (RS_N - RS_S), RS_N - RS_S);
    int r = rs_decode(d, as_built);
    if (status) status[p] = r;
    memcpy(out + p * (RS_K - RS_S), d + RS_S, RS_K - RS_S);
  }
}

/* systematic RS(204,188) encoder for test-vector generation: parity = remainder of
 * data(x) * x^16 by g(x) = prod_{i=0..15} (x - a^i) (reed_solomon.cc:176-192, 205-243) */
void dvbt_oracle_rs_encode(const uint8_t *in, long npackets, uint8_t *out) {
  if (!gf_ready) gf_init();
  unsigned char g[2 * RS_T + 1];
  memset(g, 0, sizeof g);
  g[0] = 1;
  for (int i = 1; i <= 2 * RS_T; i++) {
    int li = gf_exp_t[(i - 1) % 255];
    for (int j = i; j > 0; j--) g[j] = (unsigned char)(g[j - 1] ^ gf_mul(g[j], li));
    g[0] = (unsigned char)gf_mul(g[0], li);
  }
  for (long p = 0; p < npackets; p++) {
    unsigned char par[2 * RS_T];
    memset(par, 0, sizeof par);
    const uint8_t *d = in + p * 188;
    for (int i = 0; i < 188; i++) {
      int fb = d[i] ^ par[0];
      for (int j = 0; j < 2 * RS_T - 1; j++) par[j] = (unsigned char)(par[j + 1] ^ gf_mul(fb, g[2 * RS_T - 1 - j]));
      par[2 * RS_T - 1] = (unsigned char)gf_mul(fb, g[0]);
    }
    memcpy(out + p * 204, d, 188);
    memcpy(out + p * 204 + 188, par, 16);
  }
}
