// Drop-in for lib/demod_reference_signals_impl.cc (+ the RX half of pilot_gen in
// lib/reference_signals_impl.cc): gr::dvbt::demod_reference_signals on the B200.
#include <dvbt/demod_reference_signals.h>
#include "shim_common.h"

namespace gr {
namespace dvbt {

class demod_reference_signals_b200 : public demod_reference_signals {
  dvbt_b200_demod *d_h;

 public:
  demod_reference_signals_b200(int itemsize, int ninput, int noutput, dvbt_constellation_t constellation, dvbt_hierarchy_t hierarchy,
                               dvbt_code_rate_t code_rate_HP, dvbt_code_rate_t code_rate_LP, dvbt_guard_interval_t guard_interval,
                               dvbt_transmission_mode_t transmission_mode, int include_cell_id, int cell_id)
      : block("demod_reference_signals", io_signature::make(1, 1, itemsize * ninput), io_signature::make(1, 1, itemsize * noutput)), d_h(0) {
    dvbt_b200_demod_params p = {itemsize, ninput, noutput, (int)constellation, (int)hierarchy, (int)code_rate_HP, (int)code_rate_LP,
                                (int)guard_interval, (int)transmission_mode, include_cell_id, cell_id};
    b200::check(dvbt_b200_demod_create(&p, &d_h), "demod_reference_signals");
    set_min_noutput_items(8 * 68);               // eight TPS frames per call (bench.py drop_in_blocks: 64 items -> 18 x, 512 -> 74 x real time)
    set_min_output_buffer(0, 2 * 8 * 68);
  }
  ~demod_reference_signals_b200() { dvbt_b200_demod_destroy(d_h); }

  void forecast(int noutput_items, gr_vector_int &ninput_items_required) {
    for (size_t i = 0; i < ninput_items_required.size(); i++) ninput_items_required[i] = 2 * noutput_items;  // :87-94
  }

  int general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) {
    std::vector<dvbt_b200_tag> tin;
    b200::collect_tags(this, "sync_start", DVBT_TAG_SYNC_START, nitems_read(0), (uint64_t)ninput_items[0], tin);
    std::vector<dvbt_b200_tag> tout((size_t)noutput_items + 4);
    size_t consumed = 0, produced = 0, ntout = 0;
    b200::check(dvbt_b200_demod_work(d_h, input_items[0], (size_t)ninput_items[0], output_items[0], (size_t)noutput_items, &consumed, &produced,
                                     tin.empty() ? 0 : &tin[0], tin.size(), &tout[0], tout.size(), &ntout),
                "demod_reference_signals");
    // one symbol_index tag per produced item (the reference emits one per call because it parses
    // one item per call, demod_reference_signals_impl.cc:138-143)
    b200::emit_tags(this, nitems_written(0), &tout[0], ntout);
    consume_each((int)consumed);
    return (int)produced;
  }
};

demod_reference_signals::sptr demod_reference_signals::make(int itemsize, int ninput, int noutput, dvbt_constellation_t constellation,
                                                            dvbt_hierarchy_t hierarchy, dvbt_code_rate_t code_rate_HP,
                                                            dvbt_code_rate_t code_rate_LP, dvbt_guard_interval_t guard_interval,
                                                            dvbt_transmission_mode_t transmission_mode, int include_cell_id, int cell_id) {
  return gnuradio::get_initial_sptr(new demod_reference_signals_b200(itemsize, ninput, noutput, constellation, hierarchy, code_rate_HP,
                                                                     code_rate_LP, guard_interval, transmission_mode, include_cell_id, cell_id));
}

}  // namespace dvbt
}  // namespace gr
