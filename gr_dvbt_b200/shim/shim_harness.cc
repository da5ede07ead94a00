// TEST DRIVER for the shims: the same C entry points as oracle/ref_harness.cc (dvbt_ref_*), but the
// five hot blocks are created through their public make() — which this library defines with the B200
// shims — so oracle/refchain.py can run a chain in which those blocks are ours and every other block
// is still the reference's.  Compiled against oracle/fake_gr (no GNU Radio in this image).
#include <dvbt/demod_reference_signals.h>
#include <dvbt/dvbt_demap.h>
#include <dvbt/ofdm_sym_acquisition.h>
#include <dvbt/reed_solomon_dec.h>
#include <dvbt/viterbi_decoder.h>

#include <cstdio>
#include <string>

using namespace gr;
using namespace gr::dvbt;

namespace {
struct handle {
  boost::shared_ptr<gr::block> b;
};
inline int I(const double *a, int i) { return (int)a[i]; }
}  // namespace

extern "C" {

void *dvbt_ref_create(const char *name, const double *a, int) {
  std::string n(name);
  handle *h = new handle;
  try {
    if (n == "viterbi_decoder")
      h->b = viterbi_decoder::make((dvbt_constellation_t)I(a, 0), (dvbt_hierarchy_t)I(a, 1), (dvbt_code_rate_t)I(a, 2), I(a, 3), I(a, 4), I(a, 5));
    else if (n == "ofdm_sym_acquisition")
      h->b = ofdm_sym_acquisition::make(I(a, 0), I(a, 1), I(a, 2), I(a, 3), (float)a[4]);
    else if (n == "demod_reference_signals")
      h->b = demod_reference_signals::make(I(a, 0), I(a, 1), I(a, 2), (dvbt_constellation_t)I(a, 3), (dvbt_hierarchy_t)I(a, 4),
                                           (dvbt_code_rate_t)I(a, 5), (dvbt_code_rate_t)I(a, 6), (dvbt_guard_interval_t)I(a, 7),
                                           (dvbt_transmission_mode_t)I(a, 8), I(a, 9), I(a, 10));
    else if (n == "dvbt_demap")
      h->b = dvbt_demap::make(I(a, 0), (dvbt_constellation_t)I(a, 1), (dvbt_hierarchy_t)I(a, 2), (dvbt_transmission_mode_t)I(a, 3), (float)a[4]);
    else if (n == "reed_solomon_dec")
      h->b = reed_solomon_dec::make(I(a, 0), I(a, 1), I(a, 2), I(a, 3), I(a, 4), I(a, 5), I(a, 6), I(a, 7));
  } catch (const std::exception &e) {
    fprintf(stderr, "shim_harness: %s\n", e.what());
  }
  if (!h->b) { delete h; return 0; }
  return h;
}
void dvbt_ref_destroy(void *hv) { delete (handle *)hv; }
void dvbt_ref_add_in_tag(void *hv, unsigned long long offset, const char *key, long value) {
  tag_t t;
  t.offset = offset;
  t.key = pmt::string_to_symbol(key);
  t.value = pmt::from_long(value);
  ((handle *)hv)->b->h_in_tags.push_back(t);
}
void dvbt_ref_clear_tags(void *hv, int in_tags, int out_tags) {
  handle *h = (handle *)hv;
  if (in_tags) h->b->h_in_tags.clear();
  if (out_tags) h->b->h_out_tags.clear();
}
int dvbt_ref_num_out_tags(void *hv) { return (int)((handle *)hv)->b->h_out_tags.size(); }
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
  gr_vector_int need(1, 0);
  ((handle *)hv)->b->forecast(noutput, need);
  return need[0];
}
int dvbt_ref_general_work(void *hv, int noutput, int ninput_items, int nports, const void *in0, const void *, void *out0, void *, int *consumed) {
  handle *h = (handle *)hv;
  gr_vector_int ni(nports, ninput_items);
  gr_vector_const_void_star iv(nports, in0);
  gr_vector_void_star ov(nports, out0);
  h->b->h_consumed = 0;
  int r;
  try {
    r = h->b->general_work(noutput, ni, iv, ov);
  } catch (const std::exception &e) {
    fprintf(stderr, "shim_harness: %s\n", e.what());
    return -2;
  }
  h->b->h_nread += h->b->h_consumed;
  if (r > 0) h->b->h_nwritten += r;
  if (consumed) *consumed = h->b->h_consumed;
  return r;
}
}
