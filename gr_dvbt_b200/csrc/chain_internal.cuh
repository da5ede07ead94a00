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
}  // namespace dvbt
