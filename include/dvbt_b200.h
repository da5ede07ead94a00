/* dvbt_b200.h — C ABI of libdvbt_b200.so, the B200 (sm_100a) receive hot path that sits
 * behind gr-dvbt's GNU Radio block API.
 *
 * Every entry point below replaces one reference interface (cited as file:line relative to
 * the BogdanDIA/gr-dvbt tree).  The gr::block shims in gr_dvbt_b200/shim/ and any other
 * host (ctypes, a GR-free harness) call only these functions.  Conventions:
 *   - plain pointers and sizes, no C++ or torch types; never throws across the ABI;
 *   - return 0 on success, a negative DVBT_B200_E* code on failure;
 *     dvbt_b200_last_error() returns a thread-local description of the last failure;
 *   - "host" buffers are borrowed for the duration of the call: they are copied to and from
 *     device buffers the handle owns with cudaMemcpyAsync on the handle's stream (pageable
 *     memory is staged by the CUDA driver, pinned memory - cudaHostAlloc / cudaHostRegister by
 *     the caller - is copied at PCIe speed), and the call returns after the stream has been
 *     synchronised; "dev" entry points take CUDA device pointers on the current device and
 *     enqueue on the handle's stream, then synchronise before returning unless stated otherwise;
 *   - a handle is not thread-safe; distinct handles are independent (the reference's
 *     process-global Viterbi state, viterbi_decoder_impl.cc:49-52, is not reproduced);
 *   - there is NO CPU fallback: without a CUDA device every create() fails with
 *     DVBT_B200_ENODEV.
 */
#ifndef DVBT_B200_H
#define DVBT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DVBT_B200_OK 0
#define DVBT_B200_EINVAL (-22)  /* bad argument */
#define DVBT_B200_ENOMEM (-12)  /* host or device allocation failed */
#define DVBT_B200_ENODEV (-19)  /* no usable CUDA device */
#define DVBT_B200_ECUDA (-5)    /* a CUDA call or kernel failed */
#define DVBT_B200_ENOSPC (-28)  /* output capacity too small */

/* enums of include/dvbt/dvbt_config.h:34-76 (values are the TPS codes) */
enum { DVBT_QPSK = 0, DVBT_QAM16 = 1, DVBT_QAM64 = 2 };
enum { DVBT_NH = 0, DVBT_ALPHA1 = 1, DVBT_ALPHA2 = 2, DVBT_ALPHA4 = 3 };
enum { DVBT_C1_2 = 0, DVBT_C2_3 = 1, DVBT_C3_4 = 2, DVBT_C5_6 = 3, DVBT_C7_8 = 4 };
enum { DVBT_T2K = 0, DVBT_T8K = 1 };
enum { DVBT_G1_32 = 0, DVBT_G1_16 = 1, DVBT_G1_8 = 2, DVBT_G1_4 = 3 };

/* Stream tags (part of the block ABI).  offset is relative to the first item
 * of the buffer handed to the call (the shim subtracts nitems_read / adds nitems_written). */
enum { DVBT_TAG_SYNC_START = 1, DVBT_TAG_SUPERFRAME_START = 2, DVBT_TAG_SYMBOL_INDEX = 3 };
typedef struct dvbt_b200_tag {
  uint64_t offset;
  int32_t key;   /* DVBT_TAG_* */
  int64_t value; /* pmt::from_long payload */
} dvbt_b200_tag;

const char *dvbt_b200_last_error(void);
/* number of CUDA devices visible (0 if none / no driver); selects nothing */
int dvbt_b200_device_count(void);
/* bind the calling thread to a device.  A handle belongs to the device that was current in the thread
 * that created it; every later call on the handle switches to that device for its duration, so a handle
 * may be driven from any host thread (one call at a time per handle). */
int dvbt_b200_set_device(int device);
/* how many kernels of this library have been launched by this process (bench evidence) */
unsigned long long dvbt_b200_kernel_launches(void);
/* How the library waits for its CUDA streams inside a call: 0 (default) = cudaStreamSynchronize, the driver spins - lowest
 * latency, right while every calling thread has a host core to itself; 1 = a blocking event, the thread sleeps - for
 * processes that drive more handles than they have cores.  Process-wide; DVBT_B200_BLOCKING_WAIT=1 sets the initial value. */
int dvbt_b200_set_blocking_wait(int on);

/* ------------------------------------------------------------------------------------
 * viterbi_decoder  — replaces gr::dvbt::viterbi_decoder
 *   make():         include/dvbt/viterbi_decoder.h:51-52
 *   forecast():     lib/viterbi_decoder_impl.cc:180-189
 *   general_work(): lib/viterbi_decoder_impl.cc:191-324 (+ lib/d_viterbi.c:461-576,680-735)
 * Input: one constellation symbol per byte, m hard bits in the low bits, MSB first.
 * Output: decoded bytes, MSB first; out[i] is information byte i after a reset.
 * ------------------------------------------------------------------------------------ */
typedef struct dvbt_b200_viterbi dvbt_b200_viterbi;

typedef struct dvbt_b200_viterbi_params { /* the make() arguments, in order */
  int constellation; /* DVBT_QPSK / QAM16 / QAM64 */
  int hierarchy;     /* DVBT_NH (hierarchical modes are parameterised but untested upstream) */
  int code_rate;     /* DVBT_C1_2 .. DVBT_C7_8 */
  int bsize;         /* 768 in every shipped flowgraph */
  int S0, SK;        /* unused by the reference decoder (kept for signature parity) */
} dvbt_b200_viterbi_params;

/* Tunables of the chunk-parallel decoder (0 = library default).  They change speed only:
 * every chunk boundary is verified against the sequential decoder state and repaired when
 * it differs, so the output is bit-identical for any setting. */
typedef struct dvbt_b200_viterbi_tuning {
  int chunk_bytes; /* output bytes decoded by one GPU thread */
  int warmup_bytes; /* byte times of warm-up before a chunk's first output */
  int threads_per_block;
  int ring_depth; /* newest byte times whose survivor rows stay in shared memory (older ones: global ring) */
} dvbt_b200_viterbi_tuning;

int dvbt_b200_viterbi_create(const dvbt_b200_viterbi_params *p, dvbt_b200_viterbi **out);
void dvbt_b200_viterbi_destroy(dvbt_b200_viterbi *h);
int dvbt_b200_viterbi_set_tuning(dvbt_b200_viterbi *h, const dvbt_b200_viterbi_tuning *t);
/* what a superframe_start tag does (viterbi_decoder_impl.cc:217-221) */
int dvbt_b200_viterbi_reset(dvbt_b200_viterbi *h);
/* forecast(): input items needed for noutput_items */
int dvbt_b200_viterbi_forecast(const dvbt_b200_viterbi *h, int noutput_items);
/* set_output_multiple() value, bsize*k/8, and ntraceback */
int dvbt_b200_viterbi_output_multiple(const dvbt_b200_viterbi *h);
int dvbt_b200_viterbi_ntraceback(const dvbt_b200_viterbi *h);

/* One general_work() call on HOST buffers.  noutput_items must be a multiple of
 * output_multiple; `in` must hold forecast(noutput_items) items.  Reproduces the tag
 * behaviour: a superframe_start tag inside the window resets the decoder, and if it is not
 * at the first item the call consumes up to it and produces nothing; the first producing
 * call after a reset emits a superframe_start tag (value 1) at output offset 0 and
 * produces noutput_items - ntraceback.  The decoder state is carried from call to call. */
int dvbt_b200_viterbi_work(dvbt_b200_viterbi *h, const uint8_t *in, size_t n_in_items,
                           uint8_t *out, size_t noutput_items, size_t *consumed,
                           size_t *produced, const dvbt_b200_tag *tags_in, size_t n_tags_in,
                           dvbt_b200_tag *tags_out, size_t tags_out_capacity, size_t *n_tags_out);

/* Batch entry points: nstreams independent streams, each decoded from a reset.
 * Stream s reads n_in bytes at in + s*in_stride and writes n_in*k*m/(8n) - ntraceback bytes
 * at out + s*out_stride (*n_out receives that count).  n_in*m*k must be a multiple of 8n.
 * _host: pageable or pinned host pointers (copies are inside the call).
 * _dev:  device pointers; returns after the stream has been synchronised.  Every *_dev entry
 *        point of this library first makes its own (non-blocking) stream wait for the work
 *        already queued on the legacy default stream, so buffers that the caller filled or
 *        zeroed there (cudaMemset, cudaMemcpy, PyTorch) are ordered before the kernels. */
int dvbt_b200_viterbi_decode_host(dvbt_b200_viterbi *h, const uint8_t *in, size_t in_stride,
                                  size_t n_in, int nstreams, uint8_t *out, size_t out_stride,
                                  size_t *n_out);
int dvbt_b200_viterbi_decode_dev(dvbt_b200_viterbi *h, const uint8_t *d_in, size_t in_stride,
                                 size_t n_in, int nstreams, uint8_t *d_out, size_t out_stride,
                                 size_t *n_out);
/* Soft-decision mode - BEYOND the reference: gr-dvbt decodes hard decisions only (its soft metric table,
 * lib/d_metrics.c:57-74, is a stub and TODO.txt:25 lists soft decoding as future work), so there is no
 * reference output to match; the mode is off by default and nothing else changes when it is off.  The decoder
 * is the same trellis, traceback cadence and tie rules (viterbi_decoder_impl.cc:261-292, d_viterbi.c:461-576,
 * 680-735) with the branch metric generalised: a code bit arrives as a signed value v in [-6, 6] (clamped),
 * v > 0 meaning "1", and a branch that expects bit c earns max(v, 0) if c = 1, max(-v, 0) if c = 0.  With
 * v = +1 / -1 for hard bits 1 / 0 this IS the reference's metric (number of agreeing bits), which is how the
 * tests pin it: soft decode of +-1 values == hard decode, bit for bit; other values against oracle/port's
 * scalar restatement of the same rule.
 *   set_soft(h, 1): switch the handle (resets the stream state); work()/decode_*() then refuse.
 *   decode_soft_*:  one stream from a reset; `in` holds one int8 per TRANSMITTED code bit in the order of the
 *                   reference's input stream (X1 Y1 X2 ..., punctured positions absent); n_in*k must be a
 *                   multiple of 8n.  Writes n_in*k/(8n) - ntraceback decoded bytes. */
int dvbt_b200_viterbi_set_soft(dvbt_b200_viterbi *h, int on);
int dvbt_b200_viterbi_decode_soft_host(dvbt_b200_viterbi *h, const int8_t *in, size_t n_in, uint8_t *out, size_t *n_out);
int dvbt_b200_viterbi_decode_soft_dev(dvbt_b200_viterbi *h, const int8_t *d_in, size_t n_in, uint8_t *d_out, size_t *n_out);
/* statistics of the last decode: chunks launched, chunks whose warm-up state differed from
 * the sequential state and were re-decoded, and device time of the ACS kernel in ms */
int dvbt_b200_viterbi_last_stats(const dvbt_b200_viterbi *h, long long *chunks,
                                 long long *repaired, float *acs_kernel_ms);

/* ------------------------------------------------------------------------------------
 * reed_solomon_dec — replaces gr::dvbt::reed_solomon_dec
 *   make():         include/dvbt/reed_solomon_dec.h:49
 *   general_work(): lib/reed_solomon_dec_impl.cc:77-116 -> reed_solomon::rs_decode
 *                   (lib/reed_solomon.cc:246-489)
 * Items: `blocks` packets of n-s = 204 bytes in, `blocks` packets of k-s = 188 bytes out.
 * Packets rs_decode reports as uncorrectable pass through unmodified (its return value is
 * dropped by the block, reed_solomon_dec_impl.cc:100).
 * ------------------------------------------------------------------------------------ */
typedef struct dvbt_b200_rsdec dvbt_b200_rsdec;
typedef struct dvbt_b200_rsdec_params { /* the make() arguments, in order */
  int p, m, gfpoly, n, k, t, s, blocks; /* 2, 8, 0x11d, 255, 239, 8, 51, 8 in every flowgraph */
} dvbt_b200_rsdec_params;

int dvbt_b200_rsdec_create(const dvbt_b200_rsdec_params *p, dvbt_b200_rsdec **out);
void dvbt_b200_rsdec_destroy(dvbt_b200_rsdec *h);
/* as_built = 0 (default): the decoder the reference source describes (corrects up to t = 8
 * byte errors).  as_built = 1: reproduce the reference *binary* as gcc 13 builds it, where the
 * out-of-bounds store at reed_solomon.cc:434 (array declared :255) zeroes loc[0], so the
 * lowest-position error of every corrupted packet is left uncorrected. */
int dvbt_b200_rsdec_set_compat(dvbt_b200_rsdec *h, int as_built);
/* one general_work() call on HOST buffers; forecast is 1:1 */
int dvbt_b200_rsdec_work(dvbt_b200_rsdec *h, const uint8_t *in, size_t n_in_items, uint8_t *out,
                         size_t noutput_items, size_t *consumed, size_t *produced);
/* npackets packets of 204 bytes at d_in -> 188 bytes each at d_out (device pointers);
 * d_status (nullable, int[npackets]) receives rs_decode's return value per packet:
 * 0 clean, >0 corrected symbols, -1 uncorrectable */
int dvbt_b200_rsdec_decode_dev(dvbt_b200_rsdec *h, const uint8_t *d_in, size_t npackets,
                               uint8_t *d_out, int *d_status);

/* ------------------------------------------------------------------------------------
 * dvbt_demap — replaces gr::dvbt::dvbt_demap
 *   make():         include/dvbt/dvbt_demap.h:50
 *   general_work(): lib/dvbt_demap_impl.cc:217-240 (find_constellation_value :167-203,
 *                   make_constellation_points :117-165)
 * Items: nsize gr_complex (float re, im) in, nsize bytes out (constellation index < 2^m).
 * ------------------------------------------------------------------------------------ */
typedef struct dvbt_b200_demap dvbt_b200_demap;
typedef struct dvbt_b200_demap_params { /* the make() arguments, in order */
  int nsize;         /* cells per item: 1512 (2k) or 6048 (8k) */
  int constellation; /* DVBT_QPSK / QAM16 / QAM64 */
  int hierarchy;     /* DVBT_NH, ALPHA1, ALPHA2, ALPHA4 */
  int transmission;  /* DVBT_T2K / DVBT_T8K (unused by the decision, kept for signature parity) */
  float gain;
} dvbt_b200_demap_params;

int dvbt_b200_demap_create(const dvbt_b200_demap_params *p, dvbt_b200_demap **out);
void dvbt_b200_demap_destroy(dvbt_b200_demap *h);
/* the constellation table (re, im pairs indexed by output value); returns its size */
int dvbt_b200_demap_points(const dvbt_b200_demap *h, float *re_im, int capacity_points);
/* one general_work() call on HOST buffers; forecast is 1:1 */
int dvbt_b200_demap_work(dvbt_b200_demap *h, const void *in, size_t n_in_items, uint8_t *out,
                         size_t noutput_items, size_t *consumed, size_t *produced);
/* ncells complex cells at d_in -> ncells bytes at d_out (device pointers, 16-byte aligned) */
int dvbt_b200_demap_run_dev(dvbt_b200_demap *h, const void *d_in, size_t ncells, uint8_t *d_out);

/* ------------------------------------------------------------------------------------
 * demod_reference_signals — replaces gr::dvbt::demod_reference_signals
 *   make():         include/dvbt/demod_reference_signals.h:50-54
 *   forecast():     lib/demod_reference_signals_impl.cc:87-94 (2 input items per output item)
 *   general_work(): lib/demod_reference_signals_impl.cc:96-150 -> pilot_gen::parse_input
 *                   (lib/reference_signals_impl.cc:1188-1248 and callees)
 * Items: ninput = N gr_complex (one FFT output, DC at bin N/2) in; noutput = P gr_complex
 * (equalised payload cells) out.  Symbol i needs symbol i+1 to be visible.
 * Tags in:  sync_start (re-arms the wait for a superframe start).
 * Tags out: superframe_start (value 0xaa) on the first item produced after sync,
 *           symbol_index (0..67) on every produced item.
 * ------------------------------------------------------------------------------------ */
typedef struct dvbt_b200_demod dvbt_b200_demod;
typedef struct dvbt_b200_demod_params { /* the make() arguments, in order */
  int itemsize;          /* sizeof(gr_complex) = 8 */
  int ninput, noutput;   /* 2048/1512 or 8192/6048 */
  int constellation, hierarchy, code_rate_HP, code_rate_LP, guard_interval, transmission_mode;
  int include_cell_id, cell_id;
} dvbt_b200_demod_params;

int dvbt_b200_demod_create(const dvbt_b200_demod_params *p, dvbt_b200_demod **out);
void dvbt_b200_demod_destroy(dvbt_b200_demod *h);
/* One scheduler call on HOST buffers.  Parses up to min(n_in_items - 1, out_capacity_items)
 * symbols (the reference parses one per call; the stream behaviour is the same): consumed =
 * symbols parsed, produced = symbols emitted (none until the superframe start is found). */
int dvbt_b200_demod_work(dvbt_b200_demod *h, const void *in, size_t n_in_items, void *out,
                         size_t out_capacity_items, size_t *consumed, size_t *produced,
                         const dvbt_b200_tag *tags_in, size_t n_tags_in, dvbt_b200_tag *tags_out,
                         size_t tags_out_capacity, size_t *n_tags_out);

/* ------------------------------------------------------------------------------------
 * ofdm_sym_acquisition — replaces gr::dvbt::ofdm_sym_acquisition (and, with apply_fft, the
 * fft_vxx(N, forward, rectangular, shift=True) block that follows it in every RX flowgraph)
 *   make():         include/dvbt/ofdm_sym_acquisition.h:49
 *   forecast():     lib/ofdm_sym_acquisition_impl.cc:473-481 ((2N+cp) samples per output item)
 *   general_work(): lib/ofdm_sym_acquisition_impl.cc:488-568 (ml_sync :148-351,
 *                   peak_detect_process :72-146)
 * Items: gr_complex samples in (64/7 Msps), vectors of N gr_complex out (CP removed, derotated).
 * Tags out: sync_start (value 1) on the first item after a (re)acquisition.
 * ------------------------------------------------------------------------------------ */
typedef struct dvbt_b200_acq dvbt_b200_acq;
typedef struct dvbt_b200_acq_params { /* the make() arguments, in order */
  int blocks;         /* 1 in every flowgraph */
  int fft_length;     /* 2048 / 8192 */
  int occupied_tones; /* 1705 / 6817 (unused by the reference algorithm) */
  int cp_length;      /* N/32 ... N/4 */
  float snr;          /* dB; 30 in every flowgraph */
} dvbt_b200_acq_params;

int dvbt_b200_acq_create(const dvbt_b200_acq_params *p, dvbt_b200_acq **out);
void dvbt_b200_acq_destroy(dvbt_b200_acq *h);
/* One scheduler call on HOST buffers, as many symbols as the input allows (the reference does
 * one per call; the stream behaviour is the same).  consumed = samples consumed (N+cp per
 * symbol; half of that once after a lost peak), produced = symbols written.  apply_fft != 0
 * additionally applies the forward FFT with DC moved to bin N/2 (= fft_vxx shift=True). */
int dvbt_b200_acq_work(dvbt_b200_acq *h, const void *in, size_t n_in_items, void *out, size_t out_capacity_items,
                       size_t *consumed, size_t *produced, dvbt_b200_tag *tags_out, size_t tags_out_capacity,
                       size_t *n_tags_out, int apply_fft);

/* ------------------------------------------------------------------------------------
 * Fused receive chain (device resident).  One call = what the RX flowgraph
 * apps/dvbt_rx_demo*.grc does to a capture, from the FFT output onwards:
 *   demod_reference_signals -> dvbt_demap -> symbol_inner_interleaver(deinterleave) ->
 *   bit_inner_deinterleaver -> viterbi_decoder -> convolutional_deinterleaver(136,12,17) ->
 *   reed_solomon_dec -> energy_descramble
 * with the tags (sync_start at the first symbol and at every re-acquisition, symbol_index,
 * superframe_start) carried as batch metadata.  The TS written is what energy_descramble delivers when it
 * is called with 4 items visible and its smallest output (pairs of 8-packet groups from the NSYNC packet
 * on, energy_descramble_impl.cc:121-141); with larger calls the reference flowgraph writes a
 * scheduler-dependent prefix of the same bytes (:140-141 holds back two groups per call).
 * ------------------------------------------------------------------------------------ */
typedef struct dvbt_b200_rx dvbt_b200_rx;
typedef struct dvbt_b200_rx_params {
  int constellation, hierarchy, code_rate, guard_interval, transmission_mode;
} dvbt_b200_rx_params;
typedef struct dvbt_b200_rx_info { /* counts since the stream was reset (= of the run, for the one-shot entry points) */
  long long symbols_parsed; /* OFDM symbols run through parse_input */
  long long first_symbol;   /* stream index of the first symbol output by demod (superframe start), -1 */
  long long symbols_out;
  long long viterbi_bytes;  /* decoded bytes */
  long long viterbi_repaired; /* chunks whose boundary state had to be re-decoded */
  long long rs_packets;
  long long first_packet;   /* RS packet index where the descrambler locked (NSYNC), -1 */
  long long ts_bytes;
  long long acq_symbols;    /* symbols produced by acquisition (baseband entry) */
  long long acq_cp_start;   /* d_cp_start after the run */
  long long acq_lost_at;    /* symbol count at which tracking lost the peak, -1 never */
  long long acq_run_symbols, acq_single_symbols, acq_sequential_symbols; /* how the tracker handled the symbols */
  float ms_resample, ms_acq_fft;
  float ms_demod, ms_inner, ms_viterbi, ms_viterbi_acs, ms_rs, ms_descramble; /* device time per stage */
  float ms_fft, ms_equalise; /* single kernels inside the stages above: derotation+FFT (in ms_acq_fft), equalise+demap (in ms_demod) */
  long long n_sync_start;        /* sync_start tags acquisition sent that demod has seen (1 + re-acquisitions) */
  long long n_superframe_start;  /* superframe_start tags the Viterbi block honoured (decoder resets) */
  long long n_viterbi_runs;      /* chunk-parallel decodes launched (one per call and stream segment) */
  long long ts_total;            /* TS bytes written since the stream was reset (ts_bytes: by the last call) */
} dvbt_b200_rx_info;
enum { DVBT_RX_STAGE_CELLS = 0, DVBT_RX_STAGE_DEMAP = 1, DVBT_RX_STAGE_BITDEINT = 2, DVBT_RX_STAGE_VITERBI = 3,
       DVBT_RX_STAGE_RS = 4, DVBT_RX_STAGE_RS_STATUS = 5, DVBT_RX_STAGE_SYMBOL_INDEX = 6,
       DVBT_RX_STAGE_SOFT_CELLS = 7,  /* soft mode: uint32 per cell of the output symbols, value + 8 of bit e (0 = first) in nibble e */
       DVBT_RX_STAGE_SOFT_VALUES = 8  /* soft mode: int8 per code bit, in the order of the Viterbi block's input stream */ };

int dvbt_b200_rx_create(const dvbt_b200_rx_params *p, dvbt_b200_rx **out);
void dvbt_b200_rx_destroy(dvbt_b200_rx *h);
int dvbt_b200_rx_set_rs_compat(dvbt_b200_rx *h, int as_built); /* see dvbt_b200_rsdec_set_compat */
/* Soft-decision mode of the chain - BEYOND the reference (hard decisions only; lib/d_metrics.c:57-74 is a stub,
 * TODO.txt:25), off by default; with it off nothing changes.  on = 1: the demapper emits, per bit of a cell, the max-log
 * metric (nearest level with the bit 0 vs. nearest with the bit 1, along the axis the bit rides on) scaled so that a cell
 * on its constellation point next to a decision boundary gets +-scale (0 = default 4), weighted by the cell's channel
 * state |H|^2 / mean |H|^2 (from the pilot-based channel estimate the equaliser already holds), rounded and clamped to
 * +-6; the inner deinterleavers move these values instead of bits and the Viterbi decoder runs in soft mode
 * (dvbt_b200_viterbi_set_soft).  Everything after the Viterbi decoder is unchanged.  The sign of a non-zero value is the
 * hard decision, so a noise-free capture gives the same TS in both modes.  Switching resets the stream state. */
int dvbt_b200_rx_set_soft_decision(dvbt_b200_rx *h, int on, float scale);
/* nsym post-FFT symbols (N gr_complex each, DC at bin N/2).  _host: X and ts are host buffers;
 * _dev: device pointers.  *ts_bytes receives the TS bytes written (multiple of 1504). */
int dvbt_b200_rx_run_freq_host(dvbt_b200_rx *h, const void *X, size_t nsym, uint8_t *ts, size_t ts_capacity, size_t *ts_bytes);
int dvbt_b200_rx_run_freq_dev(dvbt_b200_rx *h, const void *dX, size_t nsym, uint8_t *d_ts, size_t ts_capacity, size_t *ts_bytes);
/* The same from time-domain baseband at the OFDM sample rate (64/7 Msps for 8 MHz channels), i.e.
 * including ofdm_sym_acquisition and the FFT: nsamples gr_complex. */
int dvbt_b200_rx_run_baseband_host(dvbt_b200_rx *h, const void *samples, size_t nsamples, uint8_t *ts, size_t ts_capacity, size_t *ts_bytes);
int dvbt_b200_rx_run_baseband_dev(dvbt_b200_rx *h, const void *d_samples, size_t nsamples, uint8_t *d_ts, size_t ts_capacity, size_t *ts_bytes);
/* The whole RX flowgraph from the capture file: complex samples at 10 Msps ->
 * rational_resampler_ccc(64,70) -> multiply_const(gain) -> the chain above.  gain is the flowgraph's
 * multiply_const value (0.0022097087 for 2k, 0.00055242272 for 8k).  The resampler and FFT are stock
 * GNU Radio blocks (not part of gr-dvbt); they follow GNU Radio 3.7's documented behaviour. */
int dvbt_b200_rx_run_file_host(dvbt_b200_rx *h, const void *samples, size_t nsamples, float gain, uint8_t *ts, size_t ts_capacity, size_t *ts_bytes);
int dvbt_b200_rx_run_file_dev(dvbt_b200_rx *h, const void *d_samples, size_t nsamples, float gain, uint8_t *d_ts, size_t ts_capacity, size_t *ts_bytes);
/* Streaming: the same chain fed in pieces of any size, the way the GNU Radio scheduler feeds the flowgraph.  Every
 * block's state is carried from call to call - the resampler's FIR history, the samples ofdm_sym_acquisition has not
 * consumed and its tracking state, the symbol demod_reference_signals is still waiting to see the successor of
 * (forecast: 2 items, demod_reference_signals_impl.cc:87-94) and its TPS/frame state, the cells short of a whole
 * 768-block and the Viterbi decoder state, the outer deinterleaver's delay lines (never cleared,
 * convolutional_deinterleaver_impl.cc:109-120), the packets energy_descramble has not consumed and its NSYNC index.
 * level: 0 = capture file samples at 10 Msps (gain = the flowgraph's multiply_const), 1 = baseband at the OFDM rate,
 * 2 = post-FFT symbols (count = symbols); a stream keeps its level until it ends or is reset.  Each call writes the TS
 * bytes that became available (*ts_bytes) at the start of ts.  end_of_stream != 0: no more input follows (a tail that
 * would need more input is dropped, exactly as at the end of a one-shot run); the next push starts a new stream.
 * The concatenated output of the pieces is the output of the one-shot run on the concatenated input.
 * ts_capacity must cover what a call can deliver (188/204 of the Viterbi bytes the piece completes, plus up to 48 packets
 * held back earlier): a buffer that might not fit is refused with DVBT_B200_ENOSPC.  After an error the stream state is
 * undefined and the next push starts a new stream.
 * Re-synchronisation inside a stream: a missed peak restarts acquisition (ofdm_sym_acquisition_impl.cc:545-558), every
 * attempt sends sync_start (:507), demod re-arms on it and waits for the next superframe start
 * (demod_reference_signals_impl.cc:112-116), whose tag resets the Viterbi decoder and re-aligns the outer
 * deinterleaver.  What the reference drops at such a tag depends on the size of the scheduler's calls (input in front
 * of a tag inside a call's window is consumed undecoded, viterbi_decoder_impl.cc:213-229,
 * convolutional_deinterleaver_impl.cc:109-120; energy_descramble re-checks NSYNC once per call); this library behaves
 * as the reference does with the smallest calls the scheduler can make (one 768-block, 2 deinterleaver items,
 * 4 x 1504 descrambler output bytes), which lose the least. */
int dvbt_b200_rx_stream_reset(dvbt_b200_rx *h);
int dvbt_b200_rx_stream_push_host(dvbt_b200_rx *h, int level, const void *data, size_t count, float gain, int end_of_stream, uint8_t *ts,
                                  size_t ts_capacity, size_t *ts_bytes);
int dvbt_b200_rx_stream_push_dev(dvbt_b200_rx *h, int level, const void *d_data, size_t count, float gain, int end_of_stream, uint8_t *d_ts,
                                 size_t ts_capacity, size_t *ts_bytes);
enum { DVBT_RX_LEVEL_FILE = 0, DVBT_RX_LEVEL_BASEBAND = 1, DVBT_RX_LEVEL_FREQ = 2 };
/* the 32/35 low-pass prototype used by the resampler; returns its length */
int dvbt_b200_resampler_taps(float *taps, int capacity);
/* The front end alone, host buffers: rational_resampler_ccc(64,70) + multiply_const(gain) (stock GNU Radio blocks of
 * apps/dvbt_rx_demo*.grc, see above).  nin gr_complex at 10 Msps -> *nout gr_complex at 64/7 Msps,
 * y[m] = gain * sum_j h[(35 m mod 32) + 32 j] * x[floor(35 m / 32) - j].  variant: -1 = the kernel the chain uses,
 * 0 generic, 1 one quad per thread, 2 / 4 quads per thread (same summation order in all of them; parity tests). */
int dvbt_b200_resample_host(const void *in, size_t nin, float gain, void *out, size_t out_capacity, size_t *nout, int variant);
int dvbt_b200_rx_last_info(const dvbt_b200_rx *h, dvbt_b200_rx_info *info);
/* copies an intermediate of the last run to the host (parity tests): DVBT_RX_STAGE_* */
int dvbt_b200_rx_read_stage(dvbt_b200_rx *h, int stage, void *host_out, size_t capacity_bytes, size_t *nbytes);

/* ------------------------------------------------------------------------------------
 * Transmit chain on the device: apps/dvbt_tx_demo*.grc as a synthetic-input generator
 * for the receive path - energy_dispersal -> reed_solomon_enc -> convolutional_interleaver(136,12,17) ->
 * inner_coder (lib/inner_coder_impl.cc:34-121, :226-262) -> bit_inner_interleaver -> symbol_inner_interleaver ->
 * dvbt_map (lib/dvbt_map_impl.cc:100-170) -> reference_signals (lib/reference_signals_impl.cc:1126-1186) ->
 * fft_vxx(reverse, shift=True) -> ofdm_cyclic_prefixer -> multiply_const(gain) -> rational_resampler_ccc(70, 64).
 * The gr-dvbt stages are bit-exact against the reference's blocks (read_stage taps); IFFT / prefix / resampler
 * are stock GNU Radio blocks restated from their documented behaviour (parity unpinned, like the RX front end).
 * npackets TS packets of 188 bytes (sync byte 0x47 first; whole groups of 8 are used) give
 * nsym = floor(npackets*204 / (P*m*k/(8n)) / 4) * 4 OFDM symbols.  level: DVBT_RX_LEVEL_FREQ = nsym x N
 * frequency-domain symbols (DC at bin N/2, what demod_reference_signals receives after a perfect channel),
 * DVBT_RX_LEVEL_BASEBAND = nsym x (N + cp) samples at the OFDM rate, DVBT_RX_LEVEL_FILE = the 10 Msps capture.
 * *count = complex values written.
 * ------------------------------------------------------------------------------------ */
typedef struct dvbt_b200_tx dvbt_b200_tx;
enum { DVBT_TX_STAGE_ENERGY = 0, DVBT_TX_STAGE_RS = 1, DVBT_TX_STAGE_OUTER = 2, DVBT_TX_STAGE_INNER_CODER = 3,
       DVBT_TX_STAGE_BIT_INTERLEAVER = 4, DVBT_TX_STAGE_SYMBOL_INTERLEAVER = 5 };
int dvbt_b200_tx_create(const dvbt_b200_rx_params *p, dvbt_b200_tx **out);
void dvbt_b200_tx_destroy(dvbt_b200_tx *h);
int dvbt_b200_tx_run_host(dvbt_b200_tx *h, const uint8_t *ts, size_t npackets, int level, float gain, void *out, size_t capacity,
                          size_t *count, size_t *nsym);
int dvbt_b200_tx_run_dev(dvbt_b200_tx *h, const uint8_t *d_ts, size_t npackets, int level, float gain, void *d_out, size_t capacity,
                         size_t *count, size_t *nsym);
/* intermediates of the last run, to the host: DVBT_TX_STAGE_* (bytes) */
int dvbt_b200_tx_read_stage(dvbt_b200_tx *h, int stage, void *host_out, size_t capacity_bytes, size_t *nbytes);

#ifdef __cplusplus
}
#endif
#endif /* DVBT_B200_H */
