// Drop-in for lib/reed_solomon_dec_impl.cc: gr::dvbt::reed_solomon_dec on the B200.
#include <dvbt/reed_solomon_dec.h>
#include "shim_common.h"

namespace gr {
namespace dvbt {

class reed_solomon_dec_b200 : public reed_solomon_dec {
  dvbt_b200_rsdec *d_h;

 public:
  reed_solomon_dec_b200(int p, int m, int gfpoly, int n, int k, int t, int s, int blocks)
      : block("reed_solomon_dec", io_signature::make(1, 1, sizeof(unsigned char) * blocks * (n - s)),
              io_signature::make(1, 1, sizeof(unsigned char) * blocks * (k - s))),
        d_h(0) {
    dvbt_b200_rsdec_params par = {p, m, gfpoly, n, k, t, s, blocks};
    b200::check(dvbt_b200_rsdec_create(&par, &d_h), "reed_solomon_dec");
    // bit-for-bit the binary the reference builds with gcc when asked for
    if (getenv("DVBT_B200_RS_AS_BUILT")) dvbt_b200_rsdec_set_compat(d_h, 1);
    set_min_noutput_items(256);
    set_min_output_buffer(0, 2 * 256);
  }
  ~reed_solomon_dec_b200() { dvbt_b200_rsdec_destroy(d_h); }

  void forecast(int noutput_items, gr_vector_int &ninput_items_required) { ninput_items_required[0] = noutput_items; }

  int general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) {
    size_t consumed = 0, produced = 0;
    b200::check(dvbt_b200_rsdec_work(d_h, (const uint8_t *)input_items[0], (size_t)ninput_items[0], (uint8_t *)output_items[0],
                                     (size_t)noutput_items, &consumed, &produced),
                "reed_solomon_dec");
    consume_each((int)consumed);
    return (int)produced;
  }
};

reed_solomon_dec::sptr reed_solomon_dec::make(int p, int m, int gfpoly, int n, int k, int t, int s, int blocks) {
  return gnuradio::get_initial_sptr(new reed_solomon_dec_b200(p, m, gfpoly, n, k, t, s, blocks));
}

}  // namespace dvbt
}  // namespace gr
