/* TEST INFRASTRUCTURE.  <gnuradio/fxpt.h> is the *last* header that
 * ofdm_sym_acquisition_impl.cc includes (:32) and the only file that includes it,
 * so it is where the oracle build pads that block's work arrays: the reference
 * writes d_corr[-2] on initial acquisition (ofdm_sym_acquisition_impl.cc:180-186
 * with low = N-2).  Padding keeps the verbatim code from corrupting
 * the heap of the test process; results are unchanged. */
#pragma once
#include <stdlib.h>
#include <string.h>
static inline int dvbt_oracle_padded_memalign(void **p, size_t, size_t size) {
  void *base = 0;
  if (posix_memalign(&base, 64, size + 256)) return 1;
  memset(base, 0, size + 256);
  *p = (char *)base + 128;
  return 0;
}
static inline void dvbt_oracle_padded_free(void *p) { free((char *)p - 128); }
#define posix_memalign dvbt_oracle_padded_memalign
#define free dvbt_oracle_padded_free
