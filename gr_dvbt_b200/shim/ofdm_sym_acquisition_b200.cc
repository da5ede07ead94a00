// Drop-in for lib/ofdm_sym_acquisition_impl.cc: gr::dvbt::ofdm_sym_acquisition on the B200.
#include <dvbt/ofdm_sym_acquisition.h>
#include "shim_common.h"

namespace gr {
namespace dvbt {

class ofdm_sym_acquisition_b200 : public ofdm_sym_acquisition {
  dvbt_b200_acq *d_h;
  int d_fft_length, d_cp_length;

 public:
  ofdm_sym_acquisition_b200(int blocks, int fft_length, int occupied_tones, int cp_length, float snr)
      : block("ofdm_sym_acquisition", io_signature::make(1, 1, sizeof(gr_complex) * blocks),
              io_signature::make(1, 1, sizeof(gr_complex) * blocks * fft_length)),
        d_h(0), d_fft_length(fft_length), d_cp_length(cp_length) {
    dvbt_b200_acq_params p = {blocks, fft_length, occupied_tones, cp_length, snr};
    b200::check(dvbt_b200_acq_create(&p, &d_h), "ofdm_sym_acquisition");
    set_relative_rate(1.0 / (double)(cp_length + fft_length));  // :388
    set_min_noutput_items(512);                   // large work items: a call costs ~0.3 ms + 3 us per symbol (bench.py drop_in_blocks)
    set_min_output_buffer(0, 2 * 512);
  }
  ~ofdm_sym_acquisition_b200() { dvbt_b200_acq_destroy(d_h); }

  void forecast(int noutput_items, gr_vector_int &ninput_items_required) {
    // the reference asks for (2N+cp) per output item (:473-481); a batch needs one window plus N+cp per further symbol
    for (size_t i = 0; i < ninput_items_required.size(); i++)
      ninput_items_required[i] = 2 * d_fft_length + d_cp_length + 32 + (noutput_items - 1) * (d_fft_length + d_cp_length);
  }

  int general_work(int noutput_items, gr_vector_int &ninput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items) {
    dvbt_b200_tag tout[64];   // one sync_start per (re)acquisition inside the call
    size_t consumed = 0, produced = 0, ntout = 0;
    b200::check(dvbt_b200_acq_work(d_h, input_items[0], (size_t)ninput_items[0], output_items[0], (size_t)noutput_items, &consumed, &produced, tout,
                                   64, &ntout, 0),
                "ofdm_sym_acquisition");
    b200::emit_tags(this, nitems_written(0), tout, ntout);
    consume_each((int)consumed);
    return (int)produced;
  }
};

ofdm_sym_acquisition::sptr ofdm_sym_acquisition::make(int blocks, int fft_length, int occupied_tones, int cp_length, float snr) {
  return gnuradio::get_initial_sptr(new ofdm_sym_acquisition_b200(blocks, fft_length, occupied_tones, cp_length, snr));
}

}  // namespace dvbt
}  // namespace gr
