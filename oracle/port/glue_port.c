/* TEST INFRASTRUCTURE ONLY — CPU restatements of the glue blocks between the hot blocks.
 *   symbol_inner_interleaver (deinterleave): lib/symbol_inner_interleaver_impl.cc:35-96 (H(q)), :161-219
 *   bit_inner_deinterleaver (non-hierarchical): lib/bit_inner_deinterleaver_impl.cc:34-58, :91-99, :120-184
 *   convolutional_deinterleaver(136,12,17): lib/convolutional_deinterleaver_impl.cc:55-68, :93-150
 *   energy_descramble: lib/energy_descramble_impl.cc:46-67 (PRBS), :108-174
 */
#include "dvbt_oracle.h"
#include <stdlib.h>
#include <string.h>

/* H(q) for 2k (tm = 0) / 8k (tm = 1); h must hold P = 1512 / 6048 ints */
void dvbt_oracle_symbol_H(int tm, int *h) {
  static const char perm2k[] = {4, 3, 9, 6, 2, 8, 1, 5, 7, 0};
  static const char perm8k[] = {7, 1, 4, 2, 9, 6, 8, 10, 0, 3, 11, 5};
  const int Mmax = tm == 0 ? 2048 : 8192, Nmax = tm == 0 ? 1512 : 6048, Nr = tm == 0 ? 11 : 13;
  const char *perm = tm == 0 ? perm2k : perm8k;
  int q = 0;
  for (int i = 0; i < Mmax; i++) {
    /* calculate_R(i), :55-95 (recomputed from scratch for every i, as the reference does) */
    int reg = 0;
    if (i >= 2) {
      reg = 1;
      for (int k = 3; k <= i; k++) {
        int nb = tm == 0 ? ((reg ^ (reg >> 3)) & 1) : ((reg ^ (reg >> 1) ^ (reg >> 4) ^ (reg >> 6)) & 1);
        reg = ((reg >> 1) | (nb << (Nr - 2))) & ((1 << Nr) - 1);
      }
    }
    int newreg = 0;
    for (int k = 0; k < Nr - 1; k++) newreg |= ((reg >> k) & 1) << perm[k];
    int v = ((i % 2) << (Nr - 1)) + newreg; /* :43 */
    if (v < Nmax) h[q++] = v;
  }
}

/* nsym items of P cells; symbol_index[k] is the tag value of item k (:199) */
void dvbt_oracle_symbol_deinterleave(const uint8_t *in, long nsym, int tm, const int *symbol_index, uint8_t *out) {
  const int P = tm == 0 ? 1512 : 6048;
  int *h = (int *)malloc(sizeof(int) * P);
  dvbt_oracle_symbol_H(tm, h);
  for (long k = 0; k < nsym; k++) {
    const uint8_t *i = in + k * P;
    uint8_t *o = out + k * P;
    for (int q = 0; q < P; q++) {
      if (symbol_index[k] % 2) o[h[q]] = i[q]; /* :202-205 */
      else o[q] = i[h[q]];                     /* :206-207 */
    }
  }
  free(h);
}

/* ncells must be a multiple of 126; v = bits per cell (2, 4, 6), non-hierarchical */
void dvbt_oracle_bit_deinterleave(const uint8_t *in, long ncells, int v, uint8_t *out) {
  static const int off[6] = {0, 63, 105, 42, 21, 84}; /* H(e, w) = (w + off[e]) % 126, :34-58 */
  unsigned char b[6][126];
  for (long blk = 0; blk < ncells / 126; blk++) {
    for (int w = 0; w < 126; w++) {
      int c = in[blk * 126 + w];
      for (int e = 0; e < v; e++) b[e][(w + off[e]) % 126] = (c >> (v - e - 1)) & 1; /* :141-147 */
    }
    for (int i = 0; i < 126; i++) {
      int c = 0;
      for (int k = 0; k < v; k++) {
        int idx = v * i + k;
        int perm = ((idx % v) / (v / 2)) + 2 * (idx % (v / 2)); /* d_perm, :94 */
        c = (c << 1) | b[perm][i];                                /* :155-156 */
      }
      out[blk * 126 + i] = (uint8_t)c;
    }
  }
}

/* 12 FIFOs of 17*(11-b) cells, zero initialised, byte t goes through FIFO t % 12 (:109-137).
 * in/out: n bytes (n a multiple of 12), stream aligned at t = 0 */
void dvbt_oracle_conv_deinterleave(const uint8_t *in, long n, uint8_t *out) {
  unsigned char *fifo[12];
  int len[12], pos[12];
  for (int b = 0; b < 12; b++) {
    len[b] = 17 * (11 - b);
    fifo[b] = (unsigned char *)calloc(len[b] + 1, 1);
    pos[b] = 0;
  }
  for (long t = 0; t < n; t++) {
    int b = (int)(t % 12);
    if (len[b] == 0) { out[t] = in[t]; continue; }
    out[t] = fifo[b][pos[b]];     /* front() after the push == oldest entry */
    fifo[b][pos[b]] = in[t];
    pos[b] = (pos[b] + 1) % len[b];
  }
  for (int b = 0; b < 12; b++) free(fifo[b]);
}

/* The block as the scheduler drives it with everything available at once: search NSYNC (0xB8) over the
 * first two 8-packet groups (:121-134); if absent skip two groups and search again; then descramble
 * complete groups from the found packet on (:140-165).  Returns bytes written (multiple of 1504);
 * *first_packet = index of the first packet output, or -1. */
long dvbt_oracle_descramble(const uint8_t *in, long npackets, uint8_t *out, long *first_packet) {
  long p0 = -1;
  for (long w = 0; w + 16 <= npackets && p0 < 0; w += 16)
    for (int i = 0; i < 16; i++)
      if (in[(w + i) * 188] == 0xB8) { p0 = w + i; break; }
  if (first_packet) *first_packet = p0;
  if (p0 < 0) return 0;
  long ngroups = (npackets - p0) / 8, count = 0;
  for (long g = 0; g < ngroups; g++) {
    unsigned reg = 0xa9; /* init_prbs, :46-49 */
    const uint8_t *src = in + (p0 + g * 8) * 188;
    for (int pk = 0; pk < 8; pk++) {
      out[count++] = 0x47;
      for (int k = 1; k < 188; k++) {
        unsigned res = 0;
        for (int i = 0; i < 8; i++) { /* clock_prbs(8), :52-67 */
          unsigned fb = ((reg >> 13) ^ (reg >> 14)) & 1u;
          reg = ((reg << 1) | fb) & 0x7fff;
          res = (res << 1) | fb;
        }
        out[count] = (uint8_t)(src[pk * 188 + k] ^ res);
        count++;
      }
      for (int i = 0; i < 8; i++) { /* clocked on the next sync byte, output unused (:162-164) */
        unsigned fb = ((reg >> 13) ^ (reg >> 14)) & 1u;
        reg = ((reg << 1) | fb) & 0x7fff;
      }
    }
  }
  return count;
}

/* One group of 8 packets (:140-165): PRBS restarted, SYNC bytes restored */
static long descramble_group(const uint8_t *src, uint8_t *out) {
  unsigned reg = 0xa9; /* init_prbs, :46-49 */
  long count = 0;
  for (int pk = 0; pk < 8; pk++) {
    out[count++] = 0x47;
    for (int k = 1; k < 188; k++) {
      unsigned res = 0;
      for (int i = 0; i < 8; i++) { /* clock_prbs(8), :52-67 */
        unsigned fb = ((reg >> 13) ^ (reg >> 14)) & 1u;
        reg = ((reg << 1) | fb) & 0x7fff;
        res = (res << 1) | fb;
      }
      out[count] = (uint8_t)(src[pk * 188 + k] ^ res);
      count++;
    }
    for (int i = 0; i < 8; i++) { /* clocked on the next sync byte, output unused (:162-164) */
      unsigned fb = ((reg >> 13) ^ (reg >> 14)) & 1u;
      reg = ((reg << 1) | fb) & 0x7fff;
    }
  }
  return count;
}

/* energy_descramble_impl::general_work (:108-174) the way the scheduler calls it with the smallest output it may ask for
 * (noutput_items = 4 x 1504, set_output_multiple :84-87, i.e. to_consume = 2 items, to_out = 2 groups) whenever 4 items
 * (forecast :102-106 asks for more; the data the call touches is 4 items) are visible.  The only state is d_index (here
 * *pk_io = d_index / 188): the search starts where NSYNC was seen last (:121-123) and walks one packet at a time up to
 * d_search = 2 groups; not found -> d_index = 0, consume 2 items, no output (:128-134); found -> the two groups at d_index
 * are descrambled blindly (:140-165) and 2 items consumed.  Processes `npackets` pending packets (8 per item); returns bytes
 * written, *items_used = items consumed, *first_packet = index of the first packet output (-1: none).
 * flush != 0 (not part of the reference: what a run that knows its input has ended can still deliver): with d_index in
 * place, keep descrambling complete pairs, then one last complete group. */
long dvbt_oracle_descramble_calls(const uint8_t *in, long npackets, int flush, int *pk_io, uint8_t *out, long *items_used,
                                  long *first_packet) {
  long i = 0, count = 0, first = -1;
  int pk = *pk_io;
  while (8 * i + 32 <= npackets) {
    while (in[(8 * i + pk) * 188] != 0xB8 && pk < 16) pk++; /* :121-123 (d_index < d_search tested second, as there) */
    if (pk >= 16) {
      pk = 0;
    } else {
      if (first < 0) first = 8 * i + pk;
      count += descramble_group(in + (8 * i + pk) * 188, out + count);
      count += descramble_group(in + (8 * i + pk + 8) * 188, out + count);
    }
    i += 2;
  }
  if (flush && pk < 16) {
    while (8 * i + pk + 16 <= npackets && in[(8 * i + pk) * 188] == 0xB8) {
      if (first < 0) first = 8 * i + pk;
      count += descramble_group(in + (8 * i + pk) * 188, out + count);
      count += descramble_group(in + (8 * i + pk + 8) * 188, out + count);
      i += 2;
    }
    if (8 * i + pk + 8 <= npackets && in[(8 * i + pk) * 188] == 0xB8 && !(8 * i + pk + 16 <= npackets)) {
      if (first < 0) first = 8 * i + pk;
      count += descramble_group(in + (8 * i + pk) * 188, out + count);
    }
  }
  *pk_io = pk;
  if (items_used) *items_used = i;
  if (first_packet) *first_packet = first;
  return count;
}
