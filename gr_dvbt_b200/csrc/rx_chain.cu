// Device-resident receive chain: post-FFT OFDM symbols -> transport stream, one H2D and one
// D2H per batch.  Stages and the reference blocks they stand for:
//
//   demod_run (demod.cu)        demod_reference_signals + dvbt_demap (fused epilogue)
//   rx_inner_codes_kernel       symbol_inner_interleaver (deinterleave,
//                               symbol_inner_interleaver_impl.cc:161-219), bit_inner_deinterleaver
//                               (bit_inner_deinterleaver_impl.cc:120-184), vector_to_stream and the
//                               Viterbi block's unpack/depuncture (viterbi_decoder_impl.cc:241-256),
//                               all as ONE index map from demapped cells to Viterbi step codes
//   vit_* kernels (viterbi.cu)  viterbi_decoder
//   rs_decode_kernel<GATHER>    convolutional_deinterleaver (index map while loading) +
//                               reed_solomon_dec
//   rx_descramble_kernel        energy_descramble (energy_descramble_impl.cc:108-174): NSYNC search
//                               over 2 groups, PRBS 1 + x^14 + x^15 restarted every 8 packets
//
// Tags become batch metadata: symbol_index per output symbol, superframe_start = first output
// symbol (the Viterbi reset and the outer deinterleaver alignment).
#include "chain_internal.cuh"

#include <string.h>
#include <new>
#include <vector>

namespace {

using dvbt::set_error;

struct InnerMap {
  const uint8_t *dm;        // demapped cells, P per parsed symbol
  const int *out_src;       // batch symbol index of output symbol o
  const int *out_symidx;    // symbol_index tag of output symbol o
  const short *H, *Hinv;
  int P, m, n_out;
};

// hard bit number `tbit` of the Viterbi block's input stream (m bits per cell, MSB first)
__device__ __forceinline__ uint32_t inner_bit(const InnerMap &im, long long tbit) {
  long long b = tbit / im.m;
  int kbit = (int)(tbit - b * im.m);
  int sym = (int)(b / im.P);
  int i = (int)(b - (long long)sym * im.P);
  int blk = i / 126, ii = i - blk * 126;
  int half = im.m >> 1;
  int e = kbit / half + 2 * (kbit % half);  // demultiplexer permutation (bit_inner_deinterleaver_impl.cc:91-99)
  // bit interleaver e delays by H(e,w) = (w + off) % 126 (:34-58)
  int off = e == 0 ? 0 : e == 1 ? 63 : e == 2 ? 105 : e == 3 ? 42 : e == 4 ? 21 : 84;
  int w = ii - off;
  if (w < 0) w += 126;
  int x = blk * 126 + w;
  // symbol deinterleaver: odd symbols out[H(q)] = in[q], even symbols out[q] = in[H(q)] (:202-208)
  int q = (im.out_symidx[sym] & 1) ? im.Hinv[x] : im.H[x];
  uint32_t cell = im.dm[(long long)im.out_src[sym] * im.P + q];
  return (cell >> (im.m - 1 - e)) & 1u;
}

__host__ __device__ constexpr int rate_k(int r) { return r == 0 ? 1 : r == 1 ? 2 : r == 2 ? 3 : r == 3 ? 5 : 7; }
__host__ __device__ constexpr unsigned rate_px(int r) { return r == 0 ? 0x1u : r == 1 ? 0x1u : r == 2 ? 0x5u : r == 3 ? 0x15u : 0x51u; }
__host__ __device__ constexpr unsigned rate_py(int r) { return r == 0 ? 0x1u : r == 1 ? 0x3u : r == 2 ? 0x3u : r == 3 ? 0x0bu : 0x2fu; }

// bits of the Viterbi input stream consumed by the first `steps` trellis steps (puncturing matrix of RATE)
template <int RATE>
__device__ __forceinline__ long long inner_bits_before(long long steps) {
  constexpr int K = rate_k(RATE), N = K + 1;
  constexpr unsigned PX = rate_px(RATE), PY = rate_py(RATE);
  long long sp = steps / K;
  int ph = (int)(steps - sp * K);
  return sp * N + __popc(PX & ((1u << ph) - 1u)) + __popc(PY & ((1u << ph) - 1u));
}

// One block per tile of 6048 cells (4 symbols in 2k mode, 1 in 8k mode).
//   phase A  one thread per cell of the symbol-deinterleaved tile: ONE gather through H / H^-1 from the
//            demapped cells, then the cell's m bits are scattered to the places the bit deinterleaver and
//            the demultiplexer give them in the Viterbi block's input stream - one byte per bit in
//            shared memory (m, the interleaver offsets and the demultiplexer map are compile-time).
//            (A gather per stream bit, as a straight index map would do, costs m times the loads.)
//   phase A' the bit bytes are packed 32 to a word: stream bit i = bit (i & 31) of word i >> 5.
//   phase B  one thread per Viterbi byte time whose first bit lies in the tile: a 32-bit window of the
//            stream at the byte time's first bit, then 8 trellis steps whose bit positions in the window
//            are compile-time for each puncturing phase -> one 32-bit step code.
// Step codes that straddle the end of the tile need <= 16 bits of the next one: those few are fetched
// with the bit-wise index map.
constexpr int kInnerTileCells = 6048, kInnerTail = 64;

// step code of 8 trellis steps starting in puncturing phase PH0, from a window whose bit k is stream bit idx + k
template <int RATE, int PH0>
__device__ __forceinline__ uint32_t inner_code_from_window(uint32_t win) {
  constexpr int K = rate_k(RATE);
  constexpr unsigned PX = rate_px(RATE), PY = rate_py(RATE);
  uint32_t wv = 0;
  int pos = 0, ph = PH0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if ((PX >> ph) & 1u) { wv |= (((win >> pos) & 1u) | 2u) << (4 * i); pos++; }
    if ((PY >> ph) & 1u) { wv |= ((((win >> pos) & 1u) << 2) | 8u) << (4 * i); pos++; }
    ph = (ph + 1 == K) ? 0 : ph + 1;
  }
  return wv;
}

template <int RATE, int M>
__global__ void __launch_bounds__(256) rx_inner_codes_kernel(InnerMap im, uint32_t *__restrict__ codes, int nbt) {
  extern __shared__ __align__(16) uint8_t s_bit[];  // [tile cells * M + kInnerTail] bit bytes, then the packed words
  constexpr int HALF = M / 2;
  const int G = kInnerTileCells / im.P;
  const int sym0 = blockIdx.x * G;
  const int nsym = min(G, im.n_out - sym0);
  const int ncell = nsym * im.P, nbits = ncell * M;
  const long long lo = (long long)sym0 * im.P * M, hi = lo + nbits;
  uint32_t *s_words = reinterpret_cast<uint32_t *>(s_bit + ((kInnerTileCells * M + kInnerTail + 15) & ~15));
  for (int ls = 0; ls < nsym; ls++) {
    const int sym = sym0 + ls;
    const short *perm = (im.out_symidx[sym] & 1) ? im.Hinv : im.H;   // symbol_inner_interleaver_impl.cc:202-208
    const uint8_t *row = im.dm + (long long)im.out_src[sym] * im.P;
    uint8_t *base = s_bit + ls * im.P * M;
    for (int x = threadIdx.x; x < im.P; x += blockDim.x) {
      uint32_t cell = row[perm[x]];
      int blk126 = (x / 126) * 126, w = x - blk126;
      uint8_t *dst = base + blk126 * M;
#pragma unroll
      for (int e = 0; e < M; e++) {
        constexpr int kOff[6] = {0, 63, 105, 42, 21, 84};              // bit interleaver e: H(e,w) = (w + off) % 126 (:34-58)
        int ii = w + kOff[e];
        if (ii >= 126) ii -= 126;
        const int kbit = (e & 1) * HALF + (e >> 1);                    // inverse of the demultiplexer permutation (:91-99)
        dst[ii * M + kbit] = (uint8_t)((cell >> (M - 1 - e)) & 1u);
      }
    }
  }
  const long long total_bits = (long long)im.n_out * im.P * M;
  for (int b = threadIdx.x; b < kInnerTail; b += blockDim.x) s_bit[nbits + b] = (hi + b < total_bits) ? (uint8_t)inner_bit(im, hi + b) : 0;
  __syncthreads();
  // ---- A': pack (4 bit bytes -> 4 bits with one multiply: bytes are 0/1, so no carries meet)
  {
    const int nw = (nbits + kInnerTail + 31) / 32;
    const uint32_t *b4 = reinterpret_cast<const uint32_t *>(s_bit);
    for (int wi = threadIdx.x; wi < nw; wi += blockDim.x) {
      uint32_t word = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        uint32_t v = b4[wi * 8 + q];                       // bit bytes 4q..4q+3 of this word
        uint32_t nib = ((v * 0x01020408u) >> 24) & 0xFu;   // byte k (bit 8k) -> bit 24 + k; no two partial products share a bit
        word |= nib << (4 * q);
      }
      s_words[wi] = word;
    }
  }
  __syncthreads();
  constexpr int K = rate_k(RATE), N = K + 1;
  // first byte time whose first bit is at or after `lo`, and the same for `hi`
  long long jlo = (lo * K / N) / 8 - 2, jhi = (hi * K / N) / 8 - 2;
  if (jlo < 0) jlo = 0;
  if (jhi < 0) jhi = 0;
  while (inner_bits_before<RATE>(8 * jlo) < lo) jlo++;
  while (inner_bits_before<RATE>(8 * jhi) < hi) jhi++;
  if (jhi > nbt) jhi = nbt;
  // The puncturing phase of byte time j is (8 j) mod K.  Byte times are taken class by class (j = jlo + c + K i):
  // within a class the phase is the same for every thread (no divergence in the switch below) and the first
  // bit advances by exactly 8 N per i (no division).  The step codes go through shared memory (the bit
  // bytes are dead after packing) so that the global stores are contiguous.
  uint32_t *s_codes = reinterpret_cast<uint32_t *>(s_bit);
  const int nj = jhi > jlo ? (int)(jhi - jlo) : 0;
  for (int c = 0; c < K; c++) {
    const long long t0 = 8 * (jlo + c);
    const int ph = (int)(t0 % K);
    const int local0 = (int)(inner_bits_before<RATE>(t0) - lo);
    const int cnt = (nj - c + K - 1) / K;
#define INNER_CLASS(PH)                                                              \
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {                            \
      int local = local0 + 8 * N * i;                                                \
      uint32_t win = __funnelshift_r(s_words[local >> 5], s_words[(local >> 5) + 1], local & 31); \
      s_codes[c + K * i] = inner_code_from_window<RATE, (PH) % K>(win);              \
    }
    switch (ph) {
      case 0: INNER_CLASS(0) break;
      case 1: INNER_CLASS(1) break;
      case 2: INNER_CLASS(2) break;
      case 3: INNER_CLASS(3) break;
      case 4: INNER_CLASS(4) break;
      case 5: INNER_CLASS(5) break;
      default: INNER_CLASS(6) break;
    }
#undef INNER_CLASS
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nj; i += blockDim.x) codes[jlo + i] = s_codes[i];
}

// test tap: the bit_inner_deinterleaver output bytes (what the reference feeds its Viterbi block)
__global__ void rx_inner_bytes_kernel(InnerMap im, uint8_t *__restrict__ out, long long nbytes) {
  long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbytes) return;
  uint32_t v = 0;
  for (int k = 0; k < im.m; k++) v = (v << 1) | inner_bit(im, b * im.m + k);
  out[b] = (uint8_t)v;
}

struct DescrInfo {
  int p0;             // index of the first packet that is output (NSYNC found there), -1: none
  long long ngroups;  // 8-packet groups written
};

// Every block repeats the (tiny) NSYNC search of energy_descramble_impl.cc:121-134 - windows of 2 groups,
// first packet whose first byte is 0xB8 - with one packet per thread, then descrambles 8-packet groups a
// 32-bit word per thread (packets are 47 words; the PRBS table sits in shared memory).
__global__ void __launch_bounds__(256) rx_descramble_kernel(const uint8_t *__restrict__ rs, long long npk, const uint32_t *__restrict__ prbs,
                                                            uint8_t *__restrict__ ts, long long ts_capacity, DescrInfo *info) {
  __shared__ long long s_p0;
  __shared__ uint32_t s_prbs[376];
  for (int i = threadIdx.x; i < 376; i += blockDim.x) s_prbs[i] = prbs[i];
  if (threadIdx.x == 0) s_p0 = -1;
  __syncthreads();
  const long long limit = (npk / 16) * 16;   // only whole 16-packet windows are searched
  for (long long base = 0; base < limit; base += blockDim.x) {
    long long pk = base + threadIdx.x;
    bool hit = pk < limit && rs[pk * 188] == 0xB8;
    unsigned any = __syncthreads_or(hit ? 1 : 0);
    if (any) {
      if (hit) atomicMin((unsigned long long *)&s_p0, (unsigned long long)pk);  // -1 reads as the largest value
      __syncthreads();
      break;
    }
  }
  long long p0 = s_p0;
  long long ngroups = p0 < 0 ? 0 : (npk - p0) / 8;
  if (ngroups * 1504 > ts_capacity) ngroups = ts_capacity / 1504;
  if (blockIdx.x == 0 && threadIdx.x == 0) { info->p0 = (int)p0; info->ngroups = ngroups; }
  if (ngroups <= 0) return;
  const uint8_t *src0 = rs + p0 * 188;
  if (((((uintptr_t)src0) | ((uintptr_t)ts)) & 3u) == 0) {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(src0);
    uint32_t *dst = reinterpret_cast<uint32_t *>(ts);
    const long long nwords = ngroups * 376;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (long long)gridDim.x * blockDim.x) {
      int k = (int)(i % 376);
      uint32_t v = src[i] ^ s_prbs[k];                       // :146-165
      if (k % 47 == 0) v = (v & 0xffffff00u) | 0x47u;          // sync byte of every packet
      dst[i] = v;
    }
  } else {
    const uint8_t *pb = reinterpret_cast<const uint8_t *>(s_prbs);
    const long long nbytes = ngroups * 1504;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (long long)gridDim.x * blockDim.x) {
      int k = (int)(i % 1504);
      ts[i] = (k % 188 == 0) ? (uint8_t)0x47 : (uint8_t)(src0[i] ^ pb[k]);
    }
  }
}

}  // namespace

struct dvbt_b200_rx {
  int device = dvbt::current_device();
  dvbt_b200_rx_params par;
  dvbt::ModeTables tables;
  dvbt::DemapTable demap;
  dvbt_b200_viterbi *vit = nullptr;
  dvbt_b200_acq *acq = nullptr;
  dvbt::DevBuf d_samples, d_sym, d_file;
  cudaStream_t stream = nullptr;
  int fi_start = 3, rs_as_built = 0, sm_count = 148;
  int k = 1, n = 2, m = 4, ntb = 5, vit_in_block = 0, vit_out_block = 0;
  dvbt::DevBuf d_X, d_state, d_fo, d_rot, d_mod, d_tps, d_vote, d_osym, d_osrc, d_dm, d_Y, d_vit, d_rs, d_rsst, d_ts, d_info, d_prbs, h_state, h_info;
  dvbt_b200_rx_info info;
  long long last_nparse = 0;
  cudaEvent_t ev[10];
};

extern "C" {

int dvbt_b200_rx_create(const dvbt_b200_rx_params *p, dvbt_b200_rx **out) {
  if (!p || !out) { set_error("rx_create: null argument"); return DVBT_B200_EINVAL; }
  *out = nullptr;
  // the fused inner index map implements the non-hierarchical demultiplexer only (bit_inner_deinterleaver_impl.cc:91-99,
  // NH branch); a hierarchical stream needs the HP/LP split of the reference flowgraph, which the block-level entry points keep
  if (p->hierarchy != DVBT_NH) { set_error("rx_create: hierarchy %d: the fused chain is non-hierarchical only", p->hierarchy); return DVBT_B200_EINVAL; }
  if (p->guard_interval < DVBT_G1_32 || p->guard_interval > DVBT_G1_4) { set_error("rx_create: bad guard interval %d", p->guard_interval); return DVBT_B200_EINVAL; }
  if (p->code_rate < DVBT_C1_2 || p->code_rate > DVBT_C7_8) { set_error("rx_create: bad code rate %d", p->code_rate); return DVBT_B200_EINVAL; }
  int rc = dvbt::ensure_device();
  if (rc) return rc;
  dvbt_b200_rx *h = new (std::nothrow) dvbt_b200_rx();
  if (!h) { set_error("rx_create: out of memory"); return DVBT_B200_ENOMEM; }
  for (auto &e : h->ev) e = nullptr;
  h->par = *p;
  rc = h->tables.init(p->transmission_mode, p->guard_interval);
  if (!rc && dvbt::make_demap_table(p->constellation, p->hierarchy, 1.0f, &h->demap)) {
    set_error("rx_create: bad constellation %d", p->constellation);
    rc = DVBT_B200_EINVAL;
  }
  if (!rc) {
    dvbt_b200_viterbi_params vp{p->constellation, p->hierarchy, p->code_rate, 768, 0, -1};
    rc = dvbt_b200_viterbi_create(&vp, &h->vit);
  }
  if (rc) { dvbt_b200_rx_destroy(h); return rc; }
  h->stream = dvbt::vit_stream(h->vit);
  {
    dvbt_b200_acq_params ap{1, h->tables.dev.N, h->tables.dev.K, h->tables.dev.cp, 30.0f};
    rc = dvbt_b200_acq_create(&ap, &h->acq);
    if (rc) { dvbt_b200_rx_destroy(h); return rc; }
    dvbt::acq_use_stream(h->acq, h->stream);
  }
  dvbt::vit_params(h->vit, &h->k, &h->n, &h->m, &h->ntb, &h->vit_in_block, &h->vit_out_block);
  h->fi_start = (p->constellation == DVBT_QAM64 && p->transmission_mode == DVBT_T8K) ? 2 : 3;
  h->h_state.host = h->h_info.host = true;
  if ((rc = h->d_state.reserve(sizeof(dvbt::DemodState))) || (rc = h->h_state.reserve(sizeof(dvbt::DemodState))) ||
      (rc = h->d_info.reserve(sizeof(DescrInfo))) || (rc = h->h_info.reserve(sizeof(DescrInfo)))) {
    dvbt_b200_rx_destroy(h);
    return rc;
  }
  for (auto &e : h->ev) cudaEventCreate(&e);
  // energy_descramble PRBS (energy_descramble_impl.cc:46-67): 1 + x^14 + x^15, init 0xa9, 8 clocks per
  // byte, clocked but unused on every sync byte except the first of the group
  {
    uint8_t tab[1504];
    unsigned reg = 0xa9;
    auto clock8 = [&]() {
      unsigned res = 0;
      for (int i = 0; i < 8; i++) {
        unsigned fb = ((reg >> 13) ^ (reg >> 14)) & 1u;
        reg = ((reg << 1) | fb) & 0x7fff;
        res = (res << 1) | fb;
      }
      return (uint8_t)res;
    };
    for (int pk = 0; pk < 8; pk++) {
      tab[pk * 188] = 0;
      for (int k = 1; k < 188; k++) tab[pk * 188 + k] = clock8();
      clock8();
    }
    if ((rc = h->d_prbs.reserve(1504))) { dvbt_b200_rx_destroy(h); return rc; }
    if (cudaMemcpy(h->d_prbs.p, tab, 1504, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("rx_create: PRBS table upload failed"); dvbt_b200_rx_destroy(h); return DVBT_B200_ECUDA; }
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
  memset(&h->info, 0, sizeof h->info);
  *out = h;
  return 0;
}

void dvbt_b200_rx_destroy(dvbt_b200_rx *h) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) return;
  if (h->stream) cudaStreamSynchronize(h->stream);
  dvbt::DevBuf *bufs[] = {&h->d_X, &h->d_state, &h->d_fo, &h->d_rot, &h->d_mod, &h->d_tps, &h->d_vote, &h->d_osym, &h->d_osrc, &h->d_dm,
                          &h->d_Y, &h->d_vit, &h->d_rs, &h->d_rsst, &h->d_ts, &h->d_info, &h->d_prbs, &h->h_state, &h->h_info};
  for (auto *b : bufs) b->release();
  for (auto &e : h->ev) if (e) cudaEventDestroy(e);
  h->tables.release();
  h->d_samples.release();
  h->d_file.release();
  h->d_sym.release();
  if (h->acq) dvbt_b200_acq_destroy(h->acq);
  if (h->vit) dvbt_b200_viterbi_destroy(h->vit);  // owns the stream
  delete h;
}

int dvbt_b200_rx_set_rs_compat(dvbt_b200_rx *h, int as_built) {
  if (!h) { set_error("rx_set_rs_compat: null handle"); return DVBT_B200_EINVAL; }
  h->rs_as_built = as_built ? 1 : 0;
  return 0;
}

// X: nsym post-FFT symbols on the device.  TS goes to d_ts (device) and optionally to a host buffer.
static int rx_run_freq(dvbt_b200_rx *h, const float2 *dX, size_t nsym, uint8_t *ts_host, uint8_t *ts_dev, size_t ts_capacity,
                       size_t *ts_bytes, int keep_cells) {
  const dvbt::ModeDev &md = h->tables.dev;
  memset(&h->info, 0, sizeof h->info);
  h->info.first_symbol = -1;
  h->info.first_packet = -1;
  if (ts_bytes) *ts_bytes = 0;
  if (nsym < 2) return 0;
  size_t nparse = nsym - 1;
  int rc;
  if ((rc = h->d_fo.reserve(nparse * 4)) || (rc = h->d_rot.reserve(nparse * 8)) || (rc = h->d_mod.reserve(nparse * 4)) ||
      (rc = h->d_tps.reserve(nparse * md.ntps * 8)) || (rc = h->d_vote.reserve(nparse * 4)) || (rc = h->d_osym.reserve(nparse * 4)) ||
      (rc = h->d_osrc.reserve(nparse * 4)) || (rc = h->d_dm.reserve(nparse * md.P)))
    return rc;
  if (keep_cells && (rc = h->d_Y.reserve(nparse * md.P * 8))) return rc;
  cudaStream_t st = h->stream;
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[0], st));
  DVBT_CUDA_TRY(cudaMemsetAsync(h->d_state.p, 0, sizeof(dvbt::DemodState), st));
  dvbt::DemodBuffers b{h->d_fo.as<int>(), h->d_rot.as<float2>(), h->d_mod.as<int>(), h->d_tps.as<float2>(), h->d_vote.as<int>(),
                       h->d_osym.as<int>(), h->d_osrc.as<int>(), h->ev[8], h->ev[9]};
  rc = dvbt::demod_run(md, &h->demap, dX, (int)nparse, b, h->d_state.as<dvbt::DemodState>(), h->fi_start, 1,
                       keep_cells ? h->d_Y.as<float2>() : nullptr, h->d_dm.as<uint8_t>(), st);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[1], st));
  DVBT_CUDA_TRY(cudaMemcpyAsync(h->h_state.p, h->d_state.p, sizeof(dvbt::DemodState), cudaMemcpyDeviceToHost, st));
  DVBT_CUDA_TRY(cudaStreamSynchronize(st));
  const dvbt::DemodState *S = h->h_state.as<dvbt::DemodState>();
  h->last_nparse = (long long)nparse;
  h->info.symbols_parsed = (long long)nparse;
  h->info.first_symbol = S->first_out;
  h->info.symbols_out = S->n_out;
  if (S->n_out <= 0) return 0;
  // Viterbi: whole 768-blocks only (viterbi_decoder_impl.cc:198)
  long long vin_bytes = (long long)S->n_out * md.P;
  long long nblocks = vin_bytes / h->vit_in_block;
  long long nbt = nblocks * h->vit_out_block;
  if (nbt <= h->ntb) return 0;
  if (nbt >= (1LL << 30)) { set_error("rx_run: batch too large (%lld byte times)", nbt); return DVBT_B200_EINVAL; }
  uint32_t *codes = dvbt::vit_reserve_codes(h->vit, (size_t)nbt);
  if (!codes) return DVBT_B200_ENOMEM;
  InnerMap im{h->d_dm.as<uint8_t>(), h->d_osrc.as<int>(), h->d_osym.as<int>(), md.H, md.Hinv, md.P, h->m, S->n_out};
  {
    const int G = kInnerTileCells / md.P;
    unsigned grid = (unsigned)((S->n_out + G - 1) / G);
    size_t smem = (size_t)((kInnerTileCells * h->m + kInnerTail + 15) & ~15) + (size_t)(kInnerTileCells * h->m + kInnerTail) / 8 + 16;
#define RX_INNER_LAUNCH(R, M) rx_inner_codes_kernel<R, M><<<grid, 256, smem, st>>>(im, codes, (int)nbt)
#define RX_INNER_RATE(R) (h->m == 2 ? RX_INNER_LAUNCH(R, 2) : h->m == 4 ? RX_INNER_LAUNCH(R, 4) : RX_INNER_LAUNCH(R, 6))
    switch (h->par.code_rate) {
      case 0: RX_INNER_RATE(0); break;
      case 1: RX_INNER_RATE(1); break;
      case 2: RX_INNER_RATE(2); break;
      case 3: RX_INNER_RATE(3); break;
      default: RX_INNER_RATE(4); break;
    }
#undef RX_INNER_RATE
#undef RX_INNER_LAUNCH
    dvbt::count_launch();
    DVBT_CUDA_TRY(cudaGetLastError());
  }
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[2], st));
  long long vout = nbt - h->ntb;
  if ((rc = h->d_vit.reserve((size_t)vout + 16))) return rc;
  if ((rc = dvbt::vit_decode_prepared(h->vit, (int)nbt, h->d_vit.as<uint8_t>()))) return rc;
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[3], st));
  h->info.viterbi_bytes = vout;
  long long npk = vout / 204;
  h->info.rs_packets = npk;
  if (npk <= 0) return dvbt::vit_collect_stats(h->vit);
  if ((rc = h->d_rs.reserve((size_t)npk * 188)) || (rc = h->d_rsst.reserve((size_t)npk * 4))) return rc;
  rc = dvbt::rs_launch(h->d_vit.as<uint8_t>(), h->d_rs.as<uint8_t>(), h->d_rsst.as<int>(), npk, h->rs_as_built, h->sm_count, st, vout);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[4], st));
  size_t cap = ts_capacity;
  uint8_t *ts_out = ts_dev;
  if (!ts_out) {
    if ((rc = h->d_ts.reserve((size_t)npk * 188))) return rc;
    ts_out = h->d_ts.as<uint8_t>();
    if (cap > (size_t)npk * 188 || ts_host == nullptr) cap = (size_t)npk * 188;
  }
  {
    long long blocks = (npk / 8 + 1) * 376 / 256 / 4 + 1;   // ~4 words per thread
    unsigned grid = (unsigned)(blocks < 8LL * h->sm_count ? blocks : 8LL * h->sm_count);
    rx_descramble_kernel<<<grid, 256, 0, st>>>(h->d_rs.as<uint8_t>(), npk, h->d_prbs.as<uint32_t>(), ts_out, (long long)cap,
                                               h->d_info.as<DescrInfo>());
    dvbt::count_launch();
    DVBT_CUDA_TRY(cudaGetLastError());
  }
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[5], st));
  DVBT_CUDA_TRY(cudaMemcpyAsync(h->h_info.p, h->d_info.p, sizeof(DescrInfo), cudaMemcpyDeviceToHost, st));
  if ((rc = dvbt::vit_collect_stats(h->vit))) return rc;  // synchronises the stream
  const DescrInfo *di = h->h_info.as<DescrInfo>();
  h->info.first_packet = di->p0;
  h->info.ts_bytes = di->ngroups * 1504;
  if (ts_host && h->info.ts_bytes > 0) {
    DVBT_CUDA_TRY(cudaMemcpyAsync(ts_host, ts_out, (size_t)h->info.ts_bytes, cudaMemcpyDeviceToHost, st));
    DVBT_CUDA_TRY(cudaStreamSynchronize(st));
  }
  if (ts_bytes) *ts_bytes = (size_t)h->info.ts_bytes;
  float ms;
  const int pairs[5][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 4}, {4, 5}};
  float *dst[5] = {&h->info.ms_demod, &h->info.ms_inner, &h->info.ms_viterbi, &h->info.ms_rs, &h->info.ms_descramble};
  for (int i = 0; i < 5; i++)
    if (cudaEventElapsedTime(&ms, h->ev[pairs[i][0]], h->ev[pairs[i][1]]) == cudaSuccess) *dst[i] = ms;
  long long chunks = 0, rep = 0;
  float acs = 0;
  dvbt_b200_viterbi_last_stats(h->vit, &chunks, &rep, &acs);
  h->info.ms_viterbi_acs = acs;
  if (cudaEventElapsedTime(&ms, h->ev[8], h->ev[9]) == cudaSuccess) h->info.ms_equalise = ms;
  h->info.viterbi_repaired = rep;
  return 0;
}

int dvbt_b200_rx_run_freq_host(dvbt_b200_rx *h, const void *X, size_t nsym, uint8_t *ts, size_t ts_capacity, size_t *ts_bytes) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (nsym && !X) || !ts) { set_error("rx_run_freq_host: bad argument"); return DVBT_B200_EINVAL; }
  const dvbt::ModeDev &md = h->tables.dev;
  int rc = h->d_X.reserve(nsym * md.N * 8);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_X.p, X, nsym * md.N * 8, cudaMemcpyHostToDevice, h->stream));
  return rx_run_freq(h, h->d_X.as<float2>(), nsym, ts, nullptr, ts_capacity, ts_bytes, 1);
}

int dvbt_b200_rx_run_freq_dev(dvbt_b200_rx *h, const void *dX, size_t nsym, uint8_t *d_ts, size_t ts_capacity, size_t *ts_bytes) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (nsym && !dX) || !d_ts) { set_error("rx_run_freq_dev: bad argument"); return DVBT_B200_EINVAL; }
  if (int rc = dvbt::join_default_stream(h->stream)) return rc;
  return rx_run_freq(h, (const float2 *)dX, nsym, nullptr, d_ts, ts_capacity, ts_bytes, 0);
}

static int rx_run_baseband(dvbt_b200_rx *h, const float2 *d_x, size_t nsamples, uint8_t *ts_host, uint8_t *ts_dev, size_t ts_capacity,
                           size_t *ts_bytes, int keep_cells) {
  const dvbt::ModeDev &md = h->tables.dev;
  if (ts_bytes) *ts_bytes = 0;
  long long cap_syms = (long long)(nsamples / (size_t)(md.N + md.cp)) + 2;
  int rc = h->d_sym.reserve((size_t)cap_syms * md.N * 8);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[6], h->stream));
  if ((rc = dvbt::acq_reset(h->acq))) return rc;
  dvbt::AcqResult ar;
  rc = dvbt::acq_run_simple(h->acq, d_x, (long long)nsamples, h->d_sym.as<float2>(), cap_syms, 1, &ar);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[7], h->stream));
  rc = rx_run_freq(h, h->d_sym.as<float2>(), (size_t)ar.n_out, ts_host, ts_dev, ts_capacity, ts_bytes, keep_cells);
  h->info.acq_symbols = ar.n_out;
  h->info.acq_cp_start = ar.cp_start;
  h->info.acq_lost_at = ar.lost_at;
  h->info.acq_run_symbols = ar.n_run;
  h->info.acq_single_symbols = ar.n_single;
  h->info.acq_sequential_symbols = ar.n_seq;
  float ms;
  if (cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]) == cudaSuccess) h->info.ms_acq_fft = ms;
  h->info.ms_fft = dvbt::acq_last_fft_ms(h->acq);
  return rc;
}

int dvbt_b200_rx_run_baseband_host(dvbt_b200_rx *h, const void *samples, size_t nsamples, uint8_t *ts, size_t ts_capacity, size_t *ts_bytes) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (nsamples && !samples) || !ts) { set_error("rx_run_baseband_host: bad argument"); return DVBT_B200_EINVAL; }
  int rc = h->d_samples.reserve(nsamples * 8);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_samples.p, samples, nsamples * 8, cudaMemcpyHostToDevice, h->stream));
  return rx_run_baseband(h, h->d_samples.as<float2>(), nsamples, ts, nullptr, ts_capacity, ts_bytes, 1);
}

int dvbt_b200_rx_run_baseband_dev(dvbt_b200_rx *h, const void *d_samples, size_t nsamples, uint8_t *d_ts, size_t ts_capacity, size_t *ts_bytes) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (nsamples && !d_samples) || !d_ts) { set_error("rx_run_baseband_dev: bad argument"); return DVBT_B200_EINVAL; }
  if (int rc = dvbt::join_default_stream(h->stream)) return rc;
  return rx_run_baseband(h, (const float2 *)d_samples, nsamples, nullptr, d_ts, ts_capacity, ts_bytes, 0);
}

static int rx_run_file(dvbt_b200_rx *h, const float2 *d_file, size_t nfile, float gain, uint8_t *ts_host, uint8_t *ts_dev,
                       size_t ts_capacity, size_t *ts_bytes, int keep_cells) {
  long long nout = dvbt::resample_out_count((long long)nfile);
  int rc = h->d_samples.reserve((size_t)nout * 8);
  if (rc) return rc;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0, h->stream);
  rc = dvbt::resample_launch(d_file, (long long)nfile, h->d_samples.as<float2>(), nout, gain, h->stream);
  cudaEventRecord(e1, h->stream);
  if (!rc) rc = rx_run_baseband(h, h->d_samples.as<float2>(), (size_t)nout, ts_host, ts_dev, ts_capacity, ts_bytes, keep_cells);
  float ms = 0;
  if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) h->info.ms_resample = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

int dvbt_b200_rx_run_file_host(dvbt_b200_rx *h, const void *samples, size_t nsamples, float gain, uint8_t *ts, size_t ts_capacity,
                               size_t *ts_bytes) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (nsamples && !samples) || !ts) { set_error("rx_run_file_host: bad argument"); return DVBT_B200_EINVAL; }
  int rc = h->d_file.reserve(nsamples * 8);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_file.p, samples, nsamples * 8, cudaMemcpyHostToDevice, h->stream));
  return rx_run_file(h, h->d_file.as<float2>(), nsamples, gain, ts, nullptr, ts_capacity, ts_bytes, 0);
}

int dvbt_b200_rx_run_file_dev(dvbt_b200_rx *h, const void *d_samples, size_t nsamples, float gain, uint8_t *d_ts, size_t ts_capacity,
                              size_t *ts_bytes) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (nsamples && !d_samples) || !d_ts) { set_error("rx_run_file_dev: bad argument"); return DVBT_B200_EINVAL; }
  if (int rc = dvbt::join_default_stream(h->stream)) return rc;
  return rx_run_file(h, (const float2 *)d_samples, nsamples, gain, nullptr, d_ts, ts_capacity, ts_bytes, 0);
}

int dvbt_b200_rx_last_info(const dvbt_b200_rx *h, dvbt_b200_rx_info *info) {
  if (!h || !info) { set_error("rx_last_info: null argument"); return DVBT_B200_EINVAL; }
  *info = h->info;
  return 0;
}

// stage taps of the last run, for stage-by-stage parity tests
int dvbt_b200_rx_read_stage(dvbt_b200_rx *h, int stage, void *host_out, size_t capacity_bytes, size_t *nbytes) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || !host_out || !nbytes) { set_error("rx_read_stage: null argument"); return DVBT_B200_EINVAL; }
  const dvbt::ModeDev &md = h->tables.dev;
  *nbytes = 0;
  const void *src = nullptr;
  size_t n = 0;
  dvbt::DevBuf tmp;
  long long nout = h->info.symbols_out, first = h->info.first_symbol;
  switch (stage) {
    case DVBT_RX_STAGE_CELLS:  // equalised cells of the output symbols (contiguous when no resync happened)
      if (nout > 0 && h->d_Y.p) { src = h->d_Y.as<float2>() + first * md.P; n = (size_t)nout * md.P * 8; }
      break;
    case DVBT_RX_STAGE_DEMAP:
      if (nout > 0) { src = h->d_dm.as<uint8_t>() + first * md.P; n = (size_t)nout * md.P; }
      break;
    case DVBT_RX_STAGE_BITDEINT: {
      if (nout <= 0) break;
      n = (size_t)nout * md.P;
      int rc = tmp.reserve(n);
      if (rc) return rc;
      InnerMap im{h->d_dm.as<uint8_t>(), h->d_osrc.as<int>(), h->d_osym.as<int>(), md.H, md.Hinv, md.P, h->m, (int)nout};
      rx_inner_bytes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(im, tmp.as<uint8_t>(), (long long)n);
      dvbt::count_launch();
      DVBT_CUDA_TRY(cudaGetLastError());
      src = tmp.p;
      break;
    }
    case DVBT_RX_STAGE_VITERBI: src = h->d_vit.p; n = (size_t)h->info.viterbi_bytes; break;
    case DVBT_RX_STAGE_RS: src = h->d_rs.p; n = (size_t)h->info.rs_packets * 188; break;
    case DVBT_RX_STAGE_RS_STATUS: src = h->d_rsst.p; n = (size_t)h->info.rs_packets * 4; break;
    case DVBT_RX_STAGE_SYMBOL_INDEX: src = h->d_osym.p; n = (size_t)(nout > 0 ? nout : 0) * 4; break;
    default: set_error("rx_read_stage: unknown stage %d", stage); return DVBT_B200_EINVAL;
  }
  if (n > capacity_bytes) { tmp.release(); set_error("rx_read_stage: need %zu bytes, capacity %zu", n, capacity_bytes); return DVBT_B200_ENOSPC; }
  if (n && src) {
    cudaError_t e = cudaMemcpyAsync(host_out, src, n, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) { tmp.release(); set_error("rx_read_stage: %s", cudaGetErrorString(e)); return DVBT_B200_ECUDA; }
    *nbytes = n;
  }
  tmp.release();
  return 0;
}

}  // extern "C"
