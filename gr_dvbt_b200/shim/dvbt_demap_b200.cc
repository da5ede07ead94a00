// Drop-in for lib/dvbt_demap_impl.cc: gr::dvbt::dvbt_demap on the B200.
#include <dvbt/dvbt_demap.h>
#include "shim_common.h"

namespace gr {
namespace dvbt {

class dvbt_demap_b200 : public dvbt_demap {
  dvbt_b200_demap *d_h;

 public:
  dvbt_demap_b200(int nsize, dvbt_constellation_t constellation, dvbt_hierarchy_t hierarchy, dvbt_transmission_mode_t transmission, float gain)
      : block("dvbt_demap", io_signature::make(1, 1, sizeof(gr_complex) * nsize), io_signature::make(1, 1, sizeof(unsigned char) * nsize)), d_h(0) {
    dvbt_b200_demap_params p = {nsize, (int)constellation, (int)hierarchy, (int)transmission, gain};
    b200::check(dvbt_b200_demap_create(&p, &d_h), "dvbt_demap");
    set_min_noutput_items(512);
    set_min_output_buffer(0, 2 * 512);
  }
  ~dvbt_demap_b200() { dvbt_b200_demap_destroy(d_h); }

  void forecast(int noutput_items, gr_vector_int &ninput_items_required) { ninput_items_required[0] = noutput_items; }

  int general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) {
    size_t consumed = 0, produced = 0;
    b200::check(dvbt_b200_demap_work(d_h, input_items[0], (size_t)ninput_items[0], (uint8_t *)output_items[0], (size_t)noutput_items, &consumed,
                                     &produced),
                "dvbt_demap");
    consume_each((int)consumed);
    return (int)produced;
  }
};

dvbt_demap::sptr dvbt_demap::make(int nsize, dvbt_constellation_t constellation, dvbt_hierarchy_t hierarchy,
                                  dvbt_transmission_mode_t transmission, float gain) {
  return gnuradio::get_initial_sptr(new dvbt_demap_b200(nsize, constellation, hierarchy, transmission, gain));
}

}  // namespace dvbt
}  // namespace gr
