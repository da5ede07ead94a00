// Transmit chain on the device: what apps/dvbt_tx_demo*.grc does to a transport stream, as a
// synthetic-input generator for the receive path - no host code between the TS and the 10 Msps capture.
//
//   energy_dispersal -> reed_solomon_enc -> convolutional_interleaver          rs.cu: tx_outer_kernel, tx_outer_interleave_kernel
//   inner_coder (inner_coder_impl.cc:34-121, :226-262) -> bit_inner_interleaver -> symbol_inner_interleaver
//                                                                               tx_cells_kernel: ONE index map per output cell
//   dvbt_map (dvbt_map_impl.cc:100-170) -> reference_signals (reference_signals_impl.cc:1126-1186, TPS :832-915)
//                                                                               tx_symbols_kernel
//   fft_vxx(reverse, shift = True) -> ofdm_cyclic_prefixer -> multiply_const    cuFFT C2C inverse + tx_prefix_kernel
//   rational_resampler_ccc(70, 64)                                              tx_resample_kernel (35/32 polyphase)
//
// The first three rows are the reference's own integer / table arithmetic and are bit-exact against its blocks
// (oracle/_ref), stage by stage; the last two are stock GNU Radio blocks (parity unpinned, like their receive-side
// counterparts): unnormalised inverse DFT of the half-swapped vector, cyclic prefix, gain; polyphase FIR with the default
// Kaiser design of rational_resampler.
//
// The convolutional encoder has no state beyond the six previous input bits, so every coded bit is a parity of a 7-bit
// window of the Forney-interleaved byte stream, and the puncturing, the packing into m-bit cells, the six bit
// interleavers and the symbol interleaver are index arithmetic on top of it: the cell at position q of symbol s is
// computed directly from the byte stream, nothing in between is materialised.
#include "chain_internal.cuh"

#include <cufft.h>
#include <math.h>
#include <string.h>
#include <new>
#include <vector>

namespace {

using dvbt::set_error;

__host__ __device__ constexpr int tx_rate_k(int r) { return r == 0 ? 1 : r == 1 ? 2 : r == 2 ? 3 : r == 3 ? 5 : 7; }
// puncture masks, bit ph = 1 if X (resp. Y) of step phase ph is transmitted (viterbi_decoder_impl.cc:61-65 / the
// X1 Y1 Y2 X3 ... lists of inner_coder_impl.cc:58-121), transmitted in the order X then Y of each step
__host__ __device__ constexpr unsigned tx_rate_px(int r) { return r == 0 ? 0x1u : r == 1 ? 0x1u : r == 2 ? 0x5u : r == 3 ? 0x15u : 0x51u; }
__host__ __device__ constexpr unsigned tx_rate_py(int r) { return r == 0 ? 0x1u : r == 1 ? 0x3u : r == 2 ? 0x3u : r == 3 ? 0x0bu : 0x2fu; }

// coded bit number `tbit` of the punctured stream: which trellis step, X or Y
template <int RATE>
__device__ __forceinline__ uint32_t tx_coded_bit(const uint8_t *__restrict__ ci, long long tbit) {
  constexpr int K = tx_rate_k(RATE), N = K + 1;
  constexpr unsigned PX = tx_rate_px(RATE), PY = tx_rate_py(RATE);
  const long long sp = tbit / N;
  int r = (int)(tbit - sp * N);
  int ph = 0, is_y = 0;
#pragma unroll
  for (int p = 0; p < K; p++) {          // walk the period: position r of the transmitted list
    if ((PX >> p) & 1u) { if (r == 0) { ph = p; is_y = 0; } r--; }
    if ((PY >> p) & 1u) { if (r == 0) { ph = p; is_y = 1; } r--; }
  }
  const long long t = sp * K + ph;       // trellis step = input bit number
  const long long b = t >> 3;
  const uint32_t w16 = (b > 0 ? (uint32_t)ci[b - 1] << 8 : 0u) | ci[b];
  const uint32_t v = w16 >> (7 - (int)(t & 7));    // bit d = input bit t - d (generate_codeword: newest at the top of d_reg)
  return __popc(v & (is_y ? 0x6Du : 0x4Fu)) & 1u;    // G1 = 171, G2 = 133 octal
}

// stage 0: inner_coder output, 1: + bit_inner_interleaver, 2: + symbol_inner_interleaver (what dvbt_map receives)
template <int RATE, int M>
__global__ void tx_cells_kernel(const uint8_t *__restrict__ ci, const short *__restrict__ H, const short *__restrict__ Hinv, int P, long long ncells,
                                int stage, uint8_t *__restrict__ cells) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ncells) return;
  const long long sym = g / P;
  const int q = (int)(g - sym * P);
  constexpr int HALF = M / 2;
  int x = q;
  if (stage >= 2) x = (sym & 1) ? H[q] : Hinv[q];      // odd symbols: out[H(q')] = in[q'] (symbol_inner_interleaver_impl.cc:202-208), TX direction
  uint32_t v = 0;
  if (stage == 0) {
#pragma unroll
    for (int kb = 0; kb < M; kb++) v |= tx_coded_bit<RATE>(ci, (sym * P + x) * M + kb) << (M - 1 - kb);
  } else {
    const int blk = (x / 126) * 126, w = x - blk;
#pragma unroll
    for (int e = 0; e < M; e++) {
      constexpr int kOff[6] = {0, 63, 105, 42, 21, 84};                 // bit interleaver e: H(e, w) = (w + off) % 126 (bit_inner_interleaver_impl.cc)
      int ii = w + kOff[e];
      if (ii >= 126) ii -= 126;
      const int kbit = (e & 1) * HALF + (e >> 1);                       // demultiplexer: stream e takes input bit kbit of every cell
      v |= tx_coded_bit<RATE>(ci, (sym * P + blk + ii) * M + kbit) << (M - 1 - e);
    }
  }
  cells[g] = (uint8_t)v;
}

// dvbt_map + reference_signals: one block per OFDM symbol.  tps_flip[frame][symbol]: parity of the TPS bits 1..symbol of that
// frame (DBPSK against the previous symbol, re-initialised at symbol 0 of every frame, reference_signals_impl.cc:832-845)
__global__ void __launch_bounds__(256) tx_symbols_kernel(dvbt::ModeDev md, const __grid_constant__ dvbt::DemapTable pts, const uint8_t *__restrict__ cells,
                                                         const uint8_t *__restrict__ tps_flip, long long first_symbol, float2 *__restrict__ X) {
  const long long s = blockIdx.x;
  const long long sa = first_symbol + s;
  const int r = (int)(sa & 3), sidx = (int)(sa % 68), frame = (int)((sa / 68) & 3);
  float2 *out = X + s * md.N;
  const int right = md.zl + md.K;
  for (int i = threadIdx.x; i < md.zl; i += blockDim.x) out[i] = make_float2(0.f, 0.f);
  for (int i = right + threadIdx.x; i < md.N; i += blockDim.x) out[i] = make_float2(0.f, 0.f);
  const unsigned char *kind = md.kind + r * md.K;
  const bool flip = tps_flip[frame * 68 + sidx] != 0;
  for (int k = threadIdx.x; k < md.K; k += blockDim.x) {
    const unsigned char kd = kind[k];
    if (kd & 1) out[md.zl + k] = make_float2(md.pval[k], 0.f);                       // scattered / continual pilot: 4/3 (1 - 2 w_k)
    else if (kd & 2) {
      const float base = md.pval[k] > 0.f ? 1.0f : -1.0f;                            // 2 (0.5 - w_k)
      out[md.zl + k] = make_float2(flip ? -base : base, 0.f);
    }
  }
  const short *pay = md.payload + r * md.P;
  const uint8_t *c = cells + s * md.P;
  for (int i = threadIdx.x; i < md.P; i += blockDim.x) out[md.zl + pay[i]] = pts.pts[c[i]];
}

// out[s][j] = gain (-1)^n ifft[s][n], n = (j - cp) mod N: the half swap of fft_vxx(shift = True) as a sign, the cyclic prefix
__global__ void tx_prefix_kernel(const float2 *__restrict__ t, int N, int cp, long long nsym, float gain, float2 *__restrict__ out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int total = N + cp;
  if (g >= nsym * total) return;
  const long long s = g / total;
  const int j = (int)(g - s * total);
  const int n = j < cp ? N - cp + j : j - cp;
  float2 v = t[s * N + n];
  const float sg = (n & 1) ? -gain : gain;
  out[g] = make_float2(v.x * sg, v.y * sg);
}

// rational_resampler_ccc(70, 64) = 35/32: y[m] = sum_j h[(32 m mod 35) + 35 j] x[floor(32 m / 35) - j], zero history
__global__ void tx_resample_kernel(const float2 *__restrict__ x, long long nin, float2 *__restrict__ y, long long nout, const float *__restrict__ taps_arm,
                                   int per_arm) {
  extern __shared__ float s_taps[];   // [35][per_arm]
  for (int i = threadIdx.x; i < 35 * per_arm; i += blockDim.x) s_taps[i] = taps_arm[i];
  __syncthreads();
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nout) return;
  const long long t = m * 32;
  const long long a = t / 35;
  const int phase = (int)(t - a * 35);
  const float *h = s_taps + phase * per_arm;
  float accr = 0.f, acci = 0.f;
  for (int j = 0; j < per_arm; j++) {
    const long long idx = a - j;
    if (idx < 0) break;
    if (idx >= nin) continue;
    const float2 v = __ldg(x + idx);
    accr = fmaf(h[j], v.x, accr);
    acci = fmaf(h[j], v.y, acci);
  }
  y[m] = make_float2(accr, acci);
}

}  // namespace

struct dvbt_b200_tx {
  int device = dvbt::current_device();
  dvbt_b200_rx_params par;
  dvbt::ModeTables tables;
  dvbt::DemapTable map;
  int k = 1, n = 2, m = 4, per_arm = 0;
  cudaStream_t stream = nullptr;
  cufftHandle plan = 0;
  int plan_batch = 0;
  dvbt::DevBuf d_ts, d_prbs, d_ed, d_rs, d_ci, d_cells, d_tap, d_X, d_t, d_bb, d_cap, d_taps, d_flip;
  long long npk = 0, nsym = 0;
};

namespace {

// the 68 TPS bits of frame f (reference_signals_impl.cc:883-915, BCH :351-382) and the running parity the DBPSK needs
void tps_flip_table(const dvbt_b200_rx_params &p, const float *pval_host0, uint8_t flip[4 * 68]) {
  for (int f = 0; f < 4; f++) {
    unsigned char d[68];
    memset(d, 0, sizeof d);
    auto set_bits = [&](int start, int stop, unsigned data) { for (int i = start; i >= stop; i--) { d[i] = data & 1u; data >>= 1; } };
    set_bits(0, 0, pval_host0[0] < 0.f ? 1u : 0u);     // d_wk[0]
    set_bits(16, 1, (f % 2) ? 0xca11u : 0x35eeu);
    set_bits(22, 17, 0x17u);                            // no cell id
    set_bits(24, 23, (unsigned)f);
    set_bits(26, 25, (unsigned)p.constellation);
    set_bits(29, 27, (unsigned)p.hierarchy);
    set_bits(32, 30, (unsigned)p.code_rate);
    set_bits(35, 33, (unsigned)p.code_rate);
    set_bits(37, 36, (unsigned)p.guard_interval);
    set_bits(39, 38, (unsigned)p.transmission_mode);
    set_bits(47, 40, 0u);
    set_bits(53, 48, 0u);
    unsigned reg = 0;
    for (int i = 0; i < 113; i++) {
      unsigned b = i < 60 ? 0u : d[1 + i - 60];
      unsigned fb = 1u & (b ^ reg);
      reg >>= 1;
      reg |= fb << 13;
      reg ^= (fb << 12) ^ (fb << 11) ^ (fb << 9) ^ (fb << 8) ^ (fb << 7) ^ (fb << 5) ^ (fb << 4);
    }
    for (int i = 0; i < 14; i++) d[i + 54] = 1u & (reg >> i);
    unsigned par = 0;
    flip[f * 68] = 0;
    for (int s = 1; s < 68; s++) { par ^= d[s]; flip[f * 68 + s] = (uint8_t)par; }
  }
}

template <int RATE>
int launch_cells(int m, const uint8_t *ci, const short *H, const short *Hinv, int P, long long ncells, int stage, uint8_t *cells, cudaStream_t st) {
  const unsigned grid = (unsigned)((ncells + 255) / 256);
  if (m == 2) tx_cells_kernel<RATE, 2><<<grid, 256, 0, st>>>(ci, H, Hinv, P, ncells, stage, cells);
  else if (m == 4) tx_cells_kernel<RATE, 4><<<grid, 256, 0, st>>>(ci, H, Hinv, P, ncells, stage, cells);
  else tx_cells_kernel<RATE, 6><<<grid, 256, 0, st>>>(ci, H, Hinv, P, ncells, stage, cells);
  dvbt::count_launch();
  DVBT_CUDA_TRY(cudaGetLastError());
  return 0;
}

int tx_cells(dvbt_b200_tx *h, int stage, uint8_t *cells) {
  const dvbt::ModeDev &md = h->tables.dev;
  const long long ncells = h->nsym * md.P;
  const uint8_t *ci = h->d_ci.as<uint8_t>();
  switch (h->par.code_rate) {
    case 0: return launch_cells<0>(h->m, ci, md.H, md.Hinv, md.P, ncells, stage, cells, h->stream);
    case 1: return launch_cells<1>(h->m, ci, md.H, md.Hinv, md.P, ncells, stage, cells, h->stream);
    case 2: return launch_cells<2>(h->m, ci, md.H, md.Hinv, md.P, ncells, stage, cells, h->stream);
    case 3: return launch_cells<3>(h->m, ci, md.H, md.Hinv, md.P, ncells, stage, cells, h->stream);
    default: return launch_cells<4>(h->m, ci, md.H, md.Hinv, md.P, ncells, stage, cells, h->stream);
  }
}

// TS on the device -> the requested level in one of the handle's buffers; *count = complex values produced
int tx_run(dvbt_b200_tx *h, const uint8_t *d_ts, size_t npackets, int level, float gain, const float2 **result, size_t *count) {
  const dvbt::ModeDev &md = h->tables.dev;
  cudaStream_t st = h->stream;
  int rc;
  *count = 0;
  *result = nullptr;
  const long long npk = (long long)(npackets / 8 * 8);                     // whole 8-packet groups (energy_dispersal_impl.cc:113)
  const long long per_sym = (long long)md.P * h->m * h->k / (8 * h->n);    // bytes of the inner coder's input per OFDM symbol
  const long long nsym = npk * 204 / per_sym / 4 * 4;                      // inner_coder: set_output_multiple(4)
  h->npk = npk;
  h->nsym = nsym;
  if (npk <= 0 || nsym <= 0) return 0;
  if ((rc = h->d_ed.reserve((size_t)npk * 188)) || (rc = h->d_rs.reserve((size_t)npk * 204)) || (rc = h->d_ci.reserve((size_t)npk * 204 + 16)) ||
      (rc = h->d_cells.reserve((size_t)nsym * md.P)) || (rc = h->d_X.reserve((size_t)nsym * md.N * 8)))
    return rc;
  if ((rc = dvbt::tx_outer_launch(d_ts, npk, h->d_prbs.as<uint8_t>(), h->d_ed.as<uint8_t>(), h->d_rs.as<uint8_t>(), h->d_ci.as<uint8_t>(), st))) return rc;
  if ((rc = tx_cells(h, 2, h->d_cells.as<uint8_t>()))) return rc;
  tx_symbols_kernel<<<(unsigned)nsym, 256, 0, st>>>(md, h->map, h->d_cells.as<uint8_t>(), h->d_flip.as<uint8_t>(), 0, h->d_X.as<float2>());
  dvbt::count_launch();
  DVBT_CUDA_TRY(cudaGetLastError());
  if (level == DVBT_RX_LEVEL_FREQ) { *result = h->d_X.as<float2>(); *count = (size_t)nsym * md.N; return 0; }
  // inverse FFT in batches (one cuFFT plan), then sign / cyclic prefix / gain
  const long long total = md.N + md.cp;
  if ((rc = h->d_t.reserve((size_t)nsym * md.N * 8)) || (rc = h->d_bb.reserve((size_t)nsym * total * 8))) return rc;
  const int batch = (int)(nsym < 4096 ? nsym : 4096);
  for (long long s0 = 0; s0 < nsym; s0 += batch) {
    const int nb = (int)(nsym - s0 < batch ? nsym - s0 : batch);
    if (h->plan == 0 || h->plan_batch != nb) {
      if (h->plan) cufftDestroy(h->plan);
      h->plan = 0;
      int nn[1] = {md.N};
      if (cufftPlanMany(&h->plan, 1, nn, nullptr, 1, md.N, nullptr, 1, md.N, CUFFT_C2C, nb) != CUFFT_SUCCESS) { set_error("tx: cufftPlanMany(%d x %d) failed", md.N, nb); return DVBT_B200_ECUDA; }
      cufftSetStream(h->plan, st);
      h->plan_batch = nb;
    }
    if (cufftExecC2C(h->plan, (cufftComplex *)(h->d_X.as<float2>() + s0 * md.N), (cufftComplex *)(h->d_t.as<float2>() + s0 * md.N), CUFFT_INVERSE) != CUFFT_SUCCESS) {
      set_error("tx: cufftExecC2C failed");
      return DVBT_B200_ECUDA;
    }
    dvbt::count_launch();
  }
  {
    const long long nn = nsym * total;
    tx_prefix_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, st>>>(h->d_t.as<float2>(), md.N, md.cp, nsym, gain, h->d_bb.as<float2>());
    dvbt::count_launch();
    DVBT_CUDA_TRY(cudaGetLastError());
  }
  if (level == DVBT_RX_LEVEL_BASEBAND) { *result = h->d_bb.as<float2>(); *count = (size_t)(nsym * total); return 0; }
  // 64/7 Msps -> 10 Msps
  const long long nin = nsym * total;
  const long long nout = nin <= 0 ? 0 : ((nin - 1) * 35) / 32 + 1;
  if ((rc = h->d_cap.reserve((size_t)nout * 8))) return rc;
  const size_t smem = (size_t)35 * h->per_arm * 4;
  tx_resample_kernel<<<(unsigned)((nout + 255) / 256), 256, smem, st>>>(h->d_bb.as<float2>(), nin, h->d_cap.as<float2>(), nout, h->d_taps.as<float>(), h->per_arm);
  dvbt::count_launch();
  DVBT_CUDA_TRY(cudaGetLastError());
  *result = h->d_cap.as<float2>();
  *count = (size_t)nout;
  return 0;
}

}  // namespace

extern "C" {

int dvbt_b200_tx_create(const dvbt_b200_rx_params *p, dvbt_b200_tx **out) {
  if (!p || !out) { set_error("tx_create: null argument"); return DVBT_B200_EINVAL; }
  *out = nullptr;
  if (p->hierarchy != DVBT_NH) { set_error("tx_create: non-hierarchical transmission only"); return DVBT_B200_EINVAL; }
  if (p->code_rate < DVBT_C1_2 || p->code_rate > DVBT_C7_8 || p->constellation < DVBT_QPSK || p->constellation > DVBT_QAM64) {
    set_error("tx_create: bad constellation / code rate (%d, %d)", p->constellation, p->code_rate);
    return DVBT_B200_EINVAL;
  }
  int rc = dvbt::ensure_device();
  if (rc) return rc;
  dvbt_b200_tx *h = new (std::nothrow) dvbt_b200_tx();
  if (!h) { set_error("tx_create: out of memory"); return DVBT_B200_ENOMEM; }
  h->par = *p;
  h->k = tx_rate_k(p->code_rate);
  h->n = h->k + 1;
  h->m = 2 * (p->constellation + 1);
  rc = h->tables.init(p->transmission_mode, p->guard_interval);
  if (!rc && dvbt::make_demap_table(p->constellation, p->hierarchy, 1.0f, &h->map)) { set_error("tx_create: bad constellation"); rc = DVBT_B200_EINVAL; }
  if (!rc && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { set_error("tx_create: cannot create stream"); rc = DVBT_B200_ECUDA; }
  if (!rc) {
    uint8_t tab[1504];
    dvbt::energy_prbs_table(tab);
    if (!(rc = h->d_prbs.reserve(1504)) && cudaMemcpy(h->d_prbs.p, tab, 1504, cudaMemcpyHostToDevice) != cudaSuccess) rc = DVBT_B200_ECUDA;
  }
  if (!rc) {
    // w_0 of the pilot PRBS (all ones register: first output bit 1) decides TPS bit 0; read it back from the device table
    float pv0 = 0.f;
    if (cudaMemcpy(&pv0, h->tables.dev.pval, 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = DVBT_B200_ECUDA;
    uint8_t flip[4 * 68];
    tps_flip_table(*p, &pv0, flip);
    if (!rc && !(rc = h->d_flip.reserve(sizeof flip)) && cudaMemcpy(h->d_flip.p, flip, sizeof flip, cudaMemcpyHostToDevice) != cudaSuccess) rc = DVBT_B200_ECUDA;
  }
  if (!rc) {
    std::vector<float> t;
    dvbt::resampler_taps_for(35, 32, &t, &h->per_arm);
    std::vector<float> arm((size_t)35 * h->per_arm);
    for (int ph = 0; ph < 35; ph++)
      for (int j = 0; j < h->per_arm; j++) arm[(size_t)ph * h->per_arm + j] = t[ph + 35 * j];
    if (!(rc = h->d_taps.reserve(arm.size() * 4)) && cudaMemcpy(h->d_taps.p, arm.data(), arm.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) rc = DVBT_B200_ECUDA;
  }
  if (rc) {
    if (rc == DVBT_B200_ECUDA) set_error("tx_create: table upload failed");
    dvbt_b200_tx_destroy(h);
    return rc;
  }
  *out = h;
  return 0;
}

void dvbt_b200_tx_destroy(dvbt_b200_tx *h) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) return;
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->plan) cufftDestroy(h->plan);
  dvbt::DevBuf *bufs[] = {&h->d_ts, &h->d_prbs, &h->d_ed, &h->d_rs, &h->d_ci, &h->d_cells, &h->d_tap, &h->d_X, &h->d_t, &h->d_bb, &h->d_cap, &h->d_taps, &h->d_flip};
  for (auto *b : bufs) b->release();
  h->tables.release();
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int dvbt_b200_tx_run_host(dvbt_b200_tx *h, const uint8_t *ts, size_t npackets, int level, float gain, void *out, size_t capacity, size_t *count,
                          size_t *nsym) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (npackets && !ts) || !out || !count || level < DVBT_RX_LEVEL_FILE || level > DVBT_RX_LEVEL_FREQ) { set_error("tx_run_host: bad argument"); return DVBT_B200_EINVAL; }
  *count = 0;
  int rc = h->d_ts.reserve(npackets * 188 + 16);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_ts.p, ts, npackets * 188, cudaMemcpyHostToDevice, h->stream));
  const float2 *res = nullptr;
  size_t n = 0;
  if ((rc = tx_run(h, h->d_ts.as<uint8_t>(), npackets, level, gain, &res, &n))) return rc;
  if (nsym) *nsym = (size_t)h->nsym;
  if (n > capacity) { set_error("tx_run_host: %zu complex values, capacity %zu", n, capacity); return DVBT_B200_ENOSPC; }
  if (n) DVBT_CUDA_TRY(cudaMemcpyAsync(out, res, n * 8, cudaMemcpyDeviceToHost, h->stream));
  DVBT_CUDA_TRY(cudaStreamSynchronize(h->stream));
  *count = n;
  return 0;
}

int dvbt_b200_tx_run_dev(dvbt_b200_tx *h, const uint8_t *d_ts, size_t npackets, int level, float gain, void *d_out, size_t capacity, size_t *count,
                         size_t *nsym) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (npackets && !d_ts) || !d_out || !count || level < DVBT_RX_LEVEL_FILE || level > DVBT_RX_LEVEL_FREQ) { set_error("tx_run_dev: bad argument"); return DVBT_B200_EINVAL; }
  *count = 0;
  if (int rc = dvbt::join_default_stream(h->stream)) return rc;
  const float2 *res = nullptr;
  size_t n = 0;
  int rc = tx_run(h, d_ts, npackets, level, gain, &res, &n);
  if (rc) return rc;
  if (nsym) *nsym = (size_t)h->nsym;
  if (n > capacity) { set_error("tx_run_dev: %zu complex values, capacity %zu", n, capacity); return DVBT_B200_ENOSPC; }
  if (n) DVBT_CUDA_TRY(cudaMemcpyAsync(d_out, res, n * 8, cudaMemcpyDeviceToDevice, h->stream));
  DVBT_CUDA_TRY(cudaStreamSynchronize(h->stream));
  *count = n;
  return 0;
}

// intermediates of the last run (parity tests against the reference's TX blocks): DVBT_TX_STAGE_*
int dvbt_b200_tx_read_stage(dvbt_b200_tx *h, int stage, void *host_out, size_t capacity_bytes, size_t *nbytes) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || !host_out || !nbytes) { set_error("tx_read_stage: null argument"); return DVBT_B200_EINVAL; }
  *nbytes = 0;
  const dvbt::ModeDev &md = h->tables.dev;
  const void *src = nullptr;
  size_t n = 0;
  switch (stage) {
    case DVBT_TX_STAGE_ENERGY: src = h->d_ed.p; n = (size_t)h->npk * 188; break;
    case DVBT_TX_STAGE_RS: src = h->d_rs.p; n = (size_t)h->npk * 204; break;
    case DVBT_TX_STAGE_OUTER: src = h->d_ci.p; n = (size_t)h->npk * 204; break;
    case DVBT_TX_STAGE_INNER_CODER:
    case DVBT_TX_STAGE_BIT_INTERLEAVER: {
      n = (size_t)h->nsym * md.P;
      int rc = h->d_tap.reserve(n);
      if (rc) return rc;
      if (n && (rc = tx_cells(h, stage == DVBT_TX_STAGE_INNER_CODER ? 0 : 1, h->d_tap.as<uint8_t>()))) return rc;
      src = h->d_tap.p;
      break;
    }
    case DVBT_TX_STAGE_SYMBOL_INTERLEAVER: src = h->d_cells.p; n = (size_t)h->nsym * md.P; break;
    default: set_error("tx_read_stage: unknown stage %d", stage); return DVBT_B200_EINVAL;
  }
  if (n > capacity_bytes) { set_error("tx_read_stage: need %zu bytes, capacity %zu", n, capacity_bytes); return DVBT_B200_ENOSPC; }
  if (n && src) {
    DVBT_CUDA_TRY(cudaMemcpyAsync(host_out, src, n, cudaMemcpyDeviceToHost, h->stream));
    DVBT_CUDA_TRY(cudaStreamSynchronize(h->stream));
    *nbytes = n;
  }
  return 0;
}

}  // extern "C"
