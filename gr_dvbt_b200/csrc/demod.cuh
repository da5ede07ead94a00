// Internal interface of the demod_reference_signals kernels (demod.cu) used by the fused
// receive chain (rx_chain.cu).
#pragma once
#include "common.cuh"

namespace dvbt {

struct DemapTable {
  float2 pts[64];
  int size;
};
int make_demap_table(int constellation, int hierarchy, float gain, DemapTable *t);
int demap_launch(const DemapTable &t, const float2 *d_in, uint8_t *d_out, long long ncells, cudaStream_t st);

// Per transmission mode constants and carrier tables (device pointers are owned by ModeTables).
struct ModeDev {
  int N, P, K, zl, cp, ncp, ntps;
  float carrier_coeff;     // reference_signals_impl.cc:753
  const short *cpilot;     // [ncp] continual pilot carriers (:54-117)
  const short *tps;        // [ntps] TPS carriers
  const float *known;      // [ncp-1] |c_{j+1} - c_j|^2 of the transmitted continual pilots (:224-228)
  const float *pval;       // [K] boosted pilot value 4/3*(1-2w_k) as float (:474-479, :700-705)
  const unsigned char *kind;  // [4][K] bit0: channel-estimation carrier (scattered(r) or continual), bit1: TPS
  const short *prevp;      // [4][K] nearest channel-estimation carrier <= k for scattered phase r
  const short *nextp;      // [4][K] nearest channel-estimation carrier  > k
  const short *payload;    // [4][P] payload carriers in increasing order
  const short *H;          // [P] symbol interleaver permutation H(q) (symbol_inner_interleaver_impl.cc:35-96)
  const short *Hinv;       // [P]
};

struct ModeTables {
  ModeDev dev;
  DevBuf blob;
  int init(int transmission_mode, int guard_interval);
  void release() { blob.release(); }
};

// Sequential receiver state of pilot_gen + demod_reference_signals_impl (device resident).
struct DemodState {
  int symbol_index, known, frame_index, prev_mod, mod, d_init;
  unsigned char fifo[68];
  float2 prev_tps[68];
  // results of the last scan
  int first_out;   // index (within the batch) of the first symbol that was output, -1 if none
  int n_out;       // symbols output by the last scan
  int sf_tag_at;   // output index carrying the superframe_start tag in the last scan, -1 if none
};

struct DemodBuffers {
  // per symbol of the batch
  int *fo;          // integer frequency offset
  float2 *rot;      // frequency-correction rotor
  int *modidx;      // scattered-pilot phase (symbol index mod 4), -1 = keep previous
  float2 *tpsval;   // [nsym][ntps] equalised TPS carriers
  int *vote;        // tps_majority_zero
  int *out_symidx;  // [n_out] symbol_index tag value of each output symbol
  int *out_src;     // [n_out] batch index of each output symbol
};

// Runs process_cpilot_data / compute_oneshot_csft / frequency_correction / process_spilot_data
// (phase detection) for symbols [0, nparse) of X (nparse+1 symbols must be readable), then the
// channel estimate + equalisation (+ optional demap) of every parsed symbol, the TPS votes and the
// sequential scan.  Y (optional) receives P equalised cells per *parsed* symbol (not compacted);
// dm (optional) the demapped bytes per parsed symbol.
int demod_run(const ModeDev &md, const DemapTable *demap, const float2 *X, int nparse, DemodBuffers b,
              DemodState *d_state, int fi_start, int sync_start_at0, float2 *Y, uint8_t *dm, cudaStream_t st);

}  // namespace dvbt
