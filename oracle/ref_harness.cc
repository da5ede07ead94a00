/* TEST INFRASTRUCTURE ONLY — oracle/_ref/libdvbt_ref.so
 *
 * A C API around the reference's own block implementations, which are compiled
 * verbatim from /root/reference/lib (see oracle/Makefile) against oracle/fake_gr.
 * It constructs a *_impl object, lets the caller attach input tags, runs one
 * general_work()/work() call and returns produced/consumed counts and output
 * tags, keeping the nitems_read/nitems_written counters the way the GNU Radio
 * scheduler would.  Python (oracle/refchain.py, tests/) scripts whole chains
 * with it.  Only tests/, __graft_entry__.smoke() and bench.py's CPU baseline may
 * load this library; it never ships in the product path.
 *
 * Known reference defects that the caller must respect:
 *  - the Viterbi decoder state is process-global (viterbi_decoder_impl.cc:49-52,
 *    d_viterbi.c:68-77): one live viterbi_decoder per process;
 *  - rs_decode has a stack overflow on corrupted packets (reed_solomon.cc:255,434).
 */
#include "bit_inner_deinterleaver_impl.h"
#include "bit_inner_interleaver_impl.h"
#include "convolutional_deinterleaver_impl.h"
#include "convolutional_interleaver_impl.h"
#include "demod_reference_signals_impl.h"
#include "dvbt_demap_impl.h"
#include "dvbt_map_impl.h"
#include "energy_descramble_impl.h"
#include "energy_dispersal_impl.h"
#include "inner_coder_impl.h"
#include "ofdm_sym_acquisition_impl.h"
#include "reed_solomon_dec_impl.h"
#include "reed_solomon_enc_impl.h"
#include "reference_signals_impl.h"
#include "symbol_inner_interleaver_impl.h"
#include "viterbi_decoder_impl.h"
/* d_viterbi.h has no include guard and already arrives through viterbi_decoder_impl.h */

#include <cstdio>
#include <cstring>
#include <string>

using namespace gr;
using namespace gr::dvbt;

namespace {
struct handle {
  block *b;
  bool is_sync;  /* derives from sync_interpolator: call work() */
  handle() : b(0), is_sync(false) {}
};
inline int I(const double *a, int i) { return (int)a[i]; }
}  // namespace

extern "C" {

/* args follow the reference make() signatures (include/dvbt/<block>.h) in order. */
void *dvbt_ref_create(const char *name, const double *a, int nargs) {
  std::string n(name);
  handle *h = new handle;
  (void)nargs;
  if (n == "viterbi_decoder")
    h->b = new viterbi_decoder_impl((dvbt_constellation_t)I(a, 0), (dvbt_hierarchy_t)I(a, 1), (dvbt_code_rate_t)I(a, 2), I(a, 3), I(a, 4), I(a, 5));
  else if (n == "ofdm_sym_acquisition")
    h->b = new ofdm_sym_acquisition_impl(I(a, 0), I(a, 1), I(a, 2), I(a, 3), (float)a[4]);
  else if (n == "demod_reference_signals")
    h->b = new demod_reference_signals_impl(I(a, 0), I(a, 1), I(a, 2), (dvbt_constellation_t)I(a, 3), (dvbt_hierarchy_t)I(a, 4), (dvbt_code_rate_t)I(a, 5),
                                            (dvbt_code_rate_t)I(a, 6), (dvbt_guard_interval_t)I(a, 7), (dvbt_transmission_mode_t)I(a, 8), I(a, 9), I(a, 10));
  else if (n == "reference_signals")
    h->b = new reference_signals_impl(I(a, 0), I(a, 1), I(a, 2), (dvbt_constellation_t)I(a, 3), (dvbt_hierarchy_t)I(a, 4), (dvbt_code_rate_t)I(a, 5),
                                      (dvbt_code_rate_t)I(a, 6), (dvbt_guard_interval_t)I(a, 7), (dvbt_transmission_mode_t)I(a, 8), I(a, 9), I(a, 10));
  else if (n == "dvbt_demap")
    h->b = new dvbt_demap_impl(I(a, 0), (dvbt_constellation_t)I(a, 1), (dvbt_hierarchy_t)I(a, 2), (dvbt_transmission_mode_t)I(a, 3), (float)a[4]);
  else if (n == "dvbt_map")
    h->b = new dvbt_map_impl(I(a, 0), (dvbt_constellation_t)I(a, 1), (dvbt_hierarchy_t)I(a, 2), (dvbt_transmission_mode_t)I(a, 3), (float)a[4]);
  else if (n == "reed_solomon_dec")
    h->b = new reed_solomon_dec_impl(I(a, 0), I(a, 1), I(a, 2), I(a, 3), I(a, 4), I(a, 5), I(a, 6), I(a, 7));
  else if (n == "reed_solomon_enc")
    h->b = new reed_solomon_enc_impl(I(a, 0), I(a, 1), I(a, 2), I(a, 3), I(a, 4), I(a, 5), I(a, 6), I(a, 7));
  else if (n == "symbol_inner_interleaver")
    h->b = new symbol_inner_interleaver_impl(I(a, 0), (dvbt_transmission_mode_t)I(a, 1), I(a, 2));
  else if (n == "bit_inner_interleaver")
    h->b = new bit_inner_interleaver_impl(I(a, 0), (dvbt_constellation_t)I(a, 1), (dvbt_hierarchy_t)I(a, 2), (dvbt_transmission_mode_t)I(a, 3));
  else if (n == "bit_inner_deinterleaver")
    h->b = new bit_inner_deinterleaver_impl(I(a, 0), (dvbt_constellation_t)I(a, 1), (dvbt_hierarchy_t)I(a, 2), (dvbt_transmission_mode_t)I(a, 3));
  else if (n == "inner_coder")
    h->b = new inner_coder_impl(I(a, 0), I(a, 1), (dvbt_constellation_t)I(a, 2), (dvbt_hierarchy_t)I(a, 3), (dvbt_code_rate_t)I(a, 4));
  else if (n == "convolutional_interleaver") {
    h->b = new convolutional_interleaver_impl(I(a, 0), I(a, 1), I(a, 2));
    h->is_sync = true;
  } else if (n == "convolutional_deinterleaver")
    h->b = new convolutional_deinterleaver_impl(I(a, 0), I(a, 1), I(a, 2));
  else if (n == "energy_dispersal")
    h->b = new energy_dispersal_impl(I(a, 0));
  else if (n == "energy_descramble")
    h->b = new energy_descramble_impl(I(a, 0));
  else {
    delete h;
    return 0;
  }
  return h;
}

void dvbt_ref_destroy(void *hv) {
  handle *h = (handle *)hv;
  if (!h) return;
  delete h->b;
  delete h;
}

void dvbt_ref_add_in_tag(void *hv, unsigned long long offset, const char *key, long value) {
  handle *h = (handle *)hv;
  tag_t t;
  t.offset = offset;
  t.key = pmt::string_to_symbol(key);
  t.value = pmt::from_long(value);
  h->b->h_in_tags.push_back(t);
}

void dvbt_ref_clear_tags(void *hv, int in_tags, int out_tags) {
  handle *h = (handle *)hv;
  if (in_tags) h->b->h_in_tags.clear();
  if (out_tags) h->b->h_out_tags.clear();
}

int dvbt_ref_num_out_tags(void *hv) { return (int)((handle *)hv)->b->h_out_tags.size(); }

/* key is copied into key_buf (at most key_len-1 chars) */
int dvbt_ref_get_out_tag(void *hv, int i, unsigned long long *offset, char *key_buf, int key_len, long *value) {
  handle *h = (handle *)hv;
  if (i < 0 || i >= (int)h->b->h_out_tags.size()) return -1;
  const tag_t &t = h->b->h_out_tags[i];
  *offset = t.offset;
  snprintf(key_buf, key_len, "%s", t.key.text.c_str());
  *value = t.value.number;
  return 0;
}

unsigned long long dvbt_ref_nitems_read(void *hv) { return ((handle *)hv)->b->h_nread; }
unsigned long long dvbt_ref_nitems_written(void *hv) { return ((handle *)hv)->b->h_nwritten; }

int dvbt_ref_forecast(void *hv, int noutput) {
  handle *h = (handle *)hv;
  gr_vector_int need(1, 0);
  h->b->forecast(noutput, need);
  return need[0];
}

/* One scheduler call.  nports = number of pointers handed to the block on each
 * side (1, or 2 for the bit (de)interleavers which index port 1 unconditionally,
 * bit_inner_deinterleaver_impl.cc:128).  Returns general_work's return value;
 * *consumed receives what the block passed to consume_each(). */
int dvbt_ref_general_work(void *hv, int noutput, int ninput_items, int nports, const void *in0, const void *in1, void *out0, void *out1, int *consumed) {
  handle *h = (handle *)hv;
  gr_vector_int ni(nports, ninput_items);
  gr_vector_const_void_star iv(nports);
  gr_vector_void_star ov(nports);
  iv[0] = in0;
  ov[0] = out0;
  if (nports > 1) {
    iv[1] = in1 ? in1 : in0;
    ov[1] = out1 ? out1 : out0;
  }
  int r;
  h->b->h_consumed = 0;
  if (h->is_sync) {
    sync_interpolator *s = static_cast<sync_interpolator *>(h->b);
    r = s->work(noutput, iv, ov);
    h->b->h_consumed = (s->h_interp ? noutput / (int)s->h_interp : noutput);
  } else {
    r = h->b->general_work(noutput, ni, iv, ov);
  }
  h->b->h_nread += h->b->h_consumed;
  if (r > 0) h->b->h_nwritten += r;
  if (consumed) *consumed = h->b->h_consumed;
  return r;
}

/* Direct access to the reference's L0 Viterbi kernels and encoder
 * (lib/d_viterbi.c) for microbenchmarks and vector generation. */
unsigned char dvbt_ref_d_encode(unsigned char *symbols, unsigned char *data, unsigned int nbytes, unsigned char encstate) {
  return d_encode(symbols, data, nbytes, encstate);
}

}  /* extern "C" */
