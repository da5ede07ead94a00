// Internal C++ entry points of the stand-alone block implementations that the fused receive
// chain (rx_chain.cu) reuses.
#pragma once
#include "demod.cuh"

#include <vector>

struct dvbt_b200_viterbi;

namespace dvbt {
cudaStream_t vit_stream(dvbt_b200_viterbi *h);
bool vit_is_soft(const dvbt_b200_viterbi *h);
int vit_params(const dvbt_b200_viterbi *h, int *k, int *n, int *m, int *ntb, int *nsymbols, int *nout);
uint32_t *vit_reserve_codes(dvbt_b200_viterbi *h, size_t nbt);
int vit_decode_prepared(dvbt_b200_viterbi *h, int nbt, uint8_t *d_out);
int vit_collect_stats(dvbt_b200_viterbi *h);
// Streaming decode (state carried from call to call, as dvbt_b200_viterbi_work does): vit_stream_codes returns where
// the new_bt step codes of the next run go (behind the retained history), vit_stream_decode decodes them to d_out
// (*nprod bytes: none for the first ntraceback byte times after a reset) and tells whether this was the first
// producing run after a reset (= the run whose first byte carries the superframe_start tag, viterbi_decoder_impl.cc:298-312).
// retain = false: the stream ends here (no history kept).  Stats accumulate over the runs of one call when enabled.
void vit_stream_reset(dvbt_b200_viterbi *h);
void vit_stream_accumulate_stats(dvbt_b200_viterbi *h, bool on);
int vit_stream_codes(dvbt_b200_viterbi *h, int new_bt, uint32_t **codes);
int vit_stream_decode(dvbt_b200_viterbi *h, int new_bt, uint8_t *d_out, size_t *nprod, bool *first_after_reset, bool retain);
// gather_stream_bytes >= 0: d_in is the Viterbi output stream and the outer (Forney) deinterleaver
// is applied while loading; -1: d_in holds packed 204-byte packets.  history_bytes (multiple of 4): bytes of the
// same stream that sit in front of d_in (a continuing stream; 2244 cover the deepest delay line), 0 = the delay
// lines start zeroed at d_in
int rs_launch(const uint8_t *d_in, uint8_t *d_out, int *d_status, long long npackets, int as_built, int sm_count,
              cudaStream_t st, long long gather_stream_bytes, long long history_bytes = 0);

struct AcqResult { long long consumed; int n_out, lost_at, fallback, cp_start, n_run, n_single, n_seq; };
}  // namespace dvbt
struct dvbt_b200_acq;
namespace dvbt {
void acq_use_stream(dvbt_b200_acq *h, cudaStream_t st);
float acq_last_fft_ms(dvbt_b200_acq *h);
int acq_reset(dvbt_b200_acq *h);
// samples x[0..n) on the device -> up to out_capacity_syms symbols of N complex at d_out
// (FFT applied, DC at bin N/2, when do_fft)
// sync_at (optional) receives, in increasing order, the output-symbol offsets (relative to this call) at which an
// acquisition attempt sent sync_start (ofdm_sym_acquisition_impl.cc:507); an offset equal to n_out refers to the
// symbol the NEXT call will produce first
// nhist: consumed samples of the stream still readable in front of x (the acquisition window reaches up to 16 samples back)
constexpr int kAcqHistory = 64;
int acq_run_simple(dvbt_b200_acq *h, const float2 *x, long long n, float2 *d_out, long long out_capacity_syms, int do_fft,
                   AcqResult *res, std::vector<long long> *sync_at = nullptr, int nhist = 0);
}  // namespace dvbt

namespace dvbt {
// rational_resampler_ccc(64,70) + multiply_const(scale): nin samples -> resample_out_count(nin) samples
long long resample_out_count(long long nin);
// nhist: valid samples in front of d_x (a continuing stream whose d_x[0] is input sample 35 q: the polyphase pattern
// restarts there, so output m of this call is output 32 q + m of the stream); 0 = the stream starts at d_x (zero history)
int resample_launch(const float2 *d_x, long long nin, float2 *d_y, long long nout, float scale, cudaStream_t st, int nhist = 0);
}  // namespace dvbt

namespace dvbt {
// transmit side (tx_chain.cu)
int tx_outer_launch(const uint8_t *d_ts, long long npk, const uint8_t *d_prbs, uint8_t *d_ed, uint8_t *d_rs, uint8_t *d_ci, cudaStream_t st);
// the 8-packet PRBS of energy_dispersal / energy_descramble (energy_descramble_impl.cc:46-67): 1 + x^14 + x^15, init 0xa9, 8
// clocks per byte, clocked but unused on the sync bytes
void energy_prbs_table(uint8_t tab[1504]);
// polyphase prototype of rational_resampler_ccc(interp, decim) after gcd reduction (GNU Radio 3.7 design_filter, beta 7, fractional_bw 0.4)
int resampler_taps_for(int interp, int decim, std::vector<float> *taps, int *per_arm);
}  // namespace dvbt
