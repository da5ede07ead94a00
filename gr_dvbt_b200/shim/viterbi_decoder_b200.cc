// Drop-in for lib/viterbi_decoder_impl.cc: gr::dvbt::viterbi_decoder on the B200.
#include <dvbt/viterbi_decoder.h>
#include "shim_common.h"

namespace gr {
namespace dvbt {

class viterbi_decoder_b200 : public viterbi_decoder {
  dvbt_b200_viterbi *d_h;
  int d_ntb;

 public:
  viterbi_decoder_b200(dvbt_constellation_t constellation, dvbt_hierarchy_t hierarchy, dvbt_code_rate_t coderate, int bsize, int S0, int SK)
      : block("viterbi_decoder", io_signature::make(1, 1, sizeof(unsigned char)), io_signature::make(1, 1, sizeof(unsigned char))), d_h(0) {
    dvbt_b200_viterbi_params p = {(int)constellation, (int)hierarchy, (int)coderate, bsize, S0, SK};
    b200::check(dvbt_b200_viterbi_create(&p, &d_h), "viterbi_decoder");
    d_ntb = dvbt_b200_viterbi_ntraceback(d_h);
    int om = dvbt_b200_viterbi_output_multiple(d_h);
    // the reference passes (k*m)/(8*n) in integer arithmetic, i.e. 0 (viterbi_decoder_impl.cc:138)
    set_relative_rate((double)om / (double)dvbt_b200_viterbi_forecast(d_h, om));
    set_output_multiple(om);                      // :141
    set_min_noutput_items(512 * om);              // one decode per 512 blocks at least (64 blocks are launch-latency bound: 0.54 ms per call)
    set_min_output_buffer(0, 2 * 512 * om);
  }
  ~viterbi_decoder_b200() { dvbt_b200_viterbi_destroy(d_h); }

  void forecast(int noutput_items, gr_vector_int &ninput_items_required) {
    for (size_t i = 0; i < ninput_items_required.size(); i++) ninput_items_required[i] = dvbt_b200_viterbi_forecast(d_h, noutput_items);
  }

  int general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) {
    const unsigned char *in = (const unsigned char *)input_items[0];
    unsigned char *out = (unsigned char *)output_items[0];
    std::vector<dvbt_b200_tag> tin;
    b200::collect_tags(this, "superframe_start", DVBT_TAG_SUPERFRAME_START, nitems_read(0), (uint64_t)ninput_items[0], tin);
    dvbt_b200_tag tout[4];
    size_t consumed = 0, produced = 0, ntout = 0;
    b200::check(dvbt_b200_viterbi_work(d_h, in, (size_t)ninput_items[0], out, (size_t)noutput_items, &consumed, &produced,
                                       tin.empty() ? 0 : &tin[0], tin.size(), tout, 4, &ntout),
                "viterbi_decoder");
    b200::emit_tags(this, nitems_written(0), tout, ntout);
    consume_each((int)consumed);
    return (int)produced;
  }
};

viterbi_decoder::sptr viterbi_decoder::make(dvbt_constellation_t constellation, dvbt_hierarchy_t hierarchy, dvbt_code_rate_t coderate,
                                            int bsize, int S0, int SK) {
  return gnuradio::get_initial_sptr(new viterbi_decoder_b200(constellation, hierarchy, coderate, bsize, S0, SK));
}

}  // namespace dvbt
}  // namespace gr
