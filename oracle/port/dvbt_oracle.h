/* TEST INFRASTRUCTURE ONLY — oracle/libdvbt_oracle.so
 *
 * Plain-C, single-thread restatement of the gr-dvbt receive hot path
 * (BogdanDIA/gr-dvbt; every function cites the reference file:line it follows).
 * It is the checker for the CUDA path: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (gr_dvbt_b200/) never links or calls it and has no CPU fallback.
 *
 * Pinning: the reference ships no golden vectors (all qa_* tests are empty
 * templates).  Each restatement is pinned against the reference's
 * own sources compiled verbatim (oracle/_ref, built by oracle/Makefile) in
 * tests/test_oracle_vs_ref.py, and against fixtures generated from that build and
 * committed under tests/golden/ (script: tests/golden/make_golden.py).
 */
#ifndef DVBT_ORACLE_H
#define DVBT_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Viterbi (viterbi_decoder_impl.cc + d_viterbi.c) ------------------------------ */
typedef struct dvbt_oracle_viterbi dvbt_oracle_viterbi;
/* m = bits per constellation symbol (2,4,6); rate = 0..4 for 1/2,2/3,3/4,5/6,7/8 */
dvbt_oracle_viterbi *dvbt_oracle_viterbi_create(int m, int rate, int bsize);
void dvbt_oracle_viterbi_destroy(dvbt_oracle_viterbi *);
/* what a superframe_start tag does (viterbi_decoder_impl.cc:217-221) */
void dvbt_oracle_viterbi_reset(dvbt_oracle_viterbi *);
/* one general_work() body without the tag search: decodes nblocks blocks of
 * bsize*n/m input bytes, writes the output bytes and returns how many are valid
 * (nblocks*bsize*k/8, minus ntraceback on the first call after a reset). */
long dvbt_oracle_viterbi_work(dvbt_oracle_viterbi *, const uint8_t *in, long nblocks, uint8_t *out);
int dvbt_oracle_viterbi_in_bytes_per_block(const dvbt_oracle_viterbi *);
int dvbt_oracle_viterbi_out_bytes_per_block(const dvbt_oracle_viterbi *);
int dvbt_oracle_viterbi_ntraceback(const dvbt_oracle_viterbi *);
/* the 64 path metrics as the reference holds them right now (after the last
 * get_output's min subtraction), for boundary-state tests */
void dvbt_oracle_viterbi_metrics(const dvbt_oracle_viterbi *, uint8_t metrics[64]);
/* soft-decision generalisation of the same decoder (not in the reference; see viterbi_port.c): one stream from a reset,
 * one int8 per transmitted code bit, returns the number of decoded bytes written (n_in*k/(8n) - ntraceback) */
long dvbt_oracle_viterbi_soft(const int8_t *in, long n_in, int rate, uint8_t *out);
/* K=7 encoder + DVB-T puncturing + packing of m bits per byte, MSB first:
 * d_viterbi.c:106-124 (d_encode) with the puncture order of
 * viterbi_decoder_impl.cc:61-65.  nbytes*8 must be a multiple of k, and
 * nbytes*8*n/k a multiple of m.  Returns the number of output bytes. */
long dvbt_oracle_conv_encode(const uint8_t *data, long nbytes, int m, int rate, uint8_t *out);

/* ---- Reed-Solomon (reed_solomon.cc, reed_solomon_dec_impl.cc) ----------------------- */
/* npackets packets of 204 bytes -> 188 bytes each; as_built: see rs_port.c; status optional */
void dvbt_oracle_rs_decode(const uint8_t *in, long npackets, uint8_t *out, int as_built, int *status);
void dvbt_oracle_rs_encode(const uint8_t *in, long npackets, uint8_t *out);

/* ---- dvbt_demap (dvbt_demap_impl.cc) ------------------------------------------------ */
int dvbt_oracle_constellation(int constellation, int alpha, float gain, float *points);
void dvbt_oracle_demap(const float *in, long n, int constellation, int alpha, float gain, uint8_t *out);

/* ---- glue blocks (glue_port.c) ------------------------------------------------------ */
void dvbt_oracle_symbol_H(int tm, int *h);
void dvbt_oracle_symbol_deinterleave(const uint8_t *in, long nsym, int tm, const int *symbol_index, uint8_t *out);
void dvbt_oracle_bit_deinterleave(const uint8_t *in, long ncells, int v, uint8_t *out);
void dvbt_oracle_conv_deinterleave(const uint8_t *in, long n, uint8_t *out);
long dvbt_oracle_descramble(const uint8_t *in, long npackets, uint8_t *out, long *first_packet);
long dvbt_oracle_descramble_calls(const uint8_t *in, long npackets, int flush, int *pk_io, uint8_t *out, long *items_used,
                                  long *first_packet);

/* ---- demod_reference_signals (demod_port.c) ------------------------------------------ */
typedef struct dvbt_oracle_demod dvbt_oracle_demod;
dvbt_oracle_demod *dvbt_oracle_demod_create(int constellation, int tm);
void dvbt_oracle_demod_destroy(dvbt_oracle_demod *);
long dvbt_oracle_demod_run(dvbt_oracle_demod *, const float *in_re_im, long nsym, int sync_start_at0, float *out_re_im,
                           int *symbol_index_out, long *superframe_tag_at);

/* ---- ofdm_sym_acquisition (acq_port.c) ----------------------------------------------- */
typedef struct dvbt_oracle_acq dvbt_oracle_acq;
dvbt_oracle_acq *dvbt_oracle_acq_create(int fft_length, int cp_length, float snr_db);
void dvbt_oracle_acq_destroy(dvbt_oracle_acq *);
long dvbt_oracle_acq_run(dvbt_oracle_acq *, const float *x_re_im, long n, float *out_re_im, long out_capacity, long *consumed,
                         int *first_sync_tag);

#ifdef __cplusplus
}
#endif
#endif
