// Internal C++ entry points of the stand-alone block implementations that the fused receive
// chain (rx_chain.cu) reuses.
#pragma once
#include "demod.cuh"

struct dvbt_b200_viterbi;

namespace dvbt {
cudaStream_t vit_stream(dvbt_b200_viterbi *h);
int vit_params(const dvbt_b200_viterbi *h, int *k, int *n, int *m, int *ntb, int *nsymbols, int *nout);
uint32_t *vit_reserve_codes(dvbt_b200_viterbi *h, size_t nbt);
int vit_decode_prepared(dvbt_b200_viterbi *h, int nbt, uint8_t *d_out);
int vit_collect_stats(dvbt_b200_viterbi *h);
// gather_stream_bytes >= 0: d_in is the Viterbi output stream and the outer (Forney) deinterleaver
// is applied while loading; -1: d_in holds packed 204-byte packets
int rs_launch(const uint8_t *d_in, uint8_t *d_out, int *d_status, long long npackets, int as_built, int sm_count,
              cudaStream_t st, long long gather_stream_bytes);

struct AcqResult { long long consumed; int n_out, lost_at, fallback, cp_start, n_run, n_single, n_seq; };
}  // namespace dvbt
struct dvbt_b200_acq;
namespace dvbt {
void acq_use_stream(dvbt_b200_acq *h, cudaStream_t st);
float acq_last_fft_ms(dvbt_b200_acq *h);
int acq_reset(dvbt_b200_acq *h);
// samples x[0..n) on the device -> up to out_capacity_syms symbols of N complex at d_out
// (FFT applied, DC at bin N/2, when do_fft)
int acq_run_simple(dvbt_b200_acq *h, const float2 *x, long long n, float2 *d_out, long long out_capacity_syms, int do_fft,
                   AcqResult *res);
}  // namespace dvbt

namespace dvbt {
// rational_resampler_ccc(64,70) + multiply_const(scale): nin samples -> resample_out_count(nin) samples
long long resample_out_count(long long nin);
int resample_launch(const float2 *d_x, long long nin, float2 *d_y, long long nout, float scale, cudaStream_t st);
}  // namespace dvbt
