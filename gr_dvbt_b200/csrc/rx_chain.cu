// Device-resident receive chain: 10 Msps capture / baseband / post-FFT OFDM symbols -> transport stream, one H2D and one
// D2H per call, as a STREAM: every block's call-to-call state is carried (dvbt_b200_rx_stream_push_*),
// the one-shot entry points are "reset, one piece, end of stream".  Stages and the reference blocks they stand for:
//
//   resample_launch (resample.cu)  rational_resampler_ccc(64, 70) + multiply_const (stock GNU Radio, parity unpinned)
//   acq_run (acq.cu)               ofdm_sym_acquisition + fft_vxx(shift = True); a lost peak restarts acquisition and
//                                  sends sync_start, as the reference does
//   demod_run (demod.cu)        demod_reference_signals + dvbt_demap (fused epilogue; soft decisions as a separate mode)
//   rx_inner_codes_kernel       symbol_inner_interleaver (deinterleave,
//                               symbol_inner_interleaver_impl.cc:161-219), bit_inner_deinterleaver
//                               (bit_inner_deinterleaver_impl.cc:120-184), vector_to_stream and the
//                               Viterbi block's unpack/depuncture (viterbi_decoder_impl.cc:241-256),
//                               all as ONE index map from demapped cells to Viterbi step codes
//                               (rx_inner_soft_kernel: the same map on 4-bit soft values)
//   vit_* kernels (viterbi.cu)  viterbi_decoder, one chunk-parallel run per stretch between superframe_start tags
//   rs_decode_kernel<GATHER>    convolutional_deinterleaver (index map while loading, 2244 bytes of stream history =
//                               the delay lines that are never cleared) + reed_solomon_dec
//   rx_descr_syncbytes_kernel / rx_descr_plan_kernel / rx_descramble_kernel
//                               energy_descramble (energy_descramble_impl.cc:108-174): the block's per-call NSYNC
//                               state machine replayed over all pending packets, then PRBS 1 + x^14 + x^15 over the
//                               planned pairs of groups
//
// Tags are metadata of the stream: sync_start (acquisition -> demod re-arm), symbol_index per output symbol,
// superframe_start (demod -> Viterbi reset -> outer deinterleaver re-alignment), at the positions and with the
// consequences they have in the reference chain driven with the scheduler's smallest calls.
#include "chain_internal.cuh"

#include <string.h>
#include <new>
#include <utility>
#include <vector>

namespace {

using dvbt::set_error;

struct InnerMap {
  const uint8_t *dm;        // demapped cells, P per parsed symbol
  const int *out_src;       // batch symbol index of output symbol o
  const int *out_symidx;    // symbol_index tag of output symbol o
  const short *H, *Hinv;
  int P, m, n_out;
  int shift_bits;           // bit of the first row that is bit 0 of the Viterbi input stream of this run (a run starts on a
                            // 768-block boundary, not on a symbol boundary, when the stream continues from an earlier call)
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
  // first byte time whose first bit is at or after `lo`, and the same for `hi` (tile positions = stream positions + shift)
  const long long shift = im.shift_bits;
  long long jlo = lo > shift ? ((lo - shift) * K / N) / 8 - 2 : 0, jhi = hi > shift ? ((hi - shift) * K / N) / 8 - 2 : 0;
  if (jlo < 0) jlo = 0;
  if (jhi < 0) jhi = 0;
  while (inner_bits_before<RATE>(8 * jlo) + shift < lo) jlo++;
  while (inner_bits_before<RATE>(8 * jhi) + shift < hi) jhi++;
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
    const int local0 = (int)(inner_bits_before<RATE>(t0) + shift - lo);
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

// ---- soft decisions (beyond the reference): the same index map on 4-bit values --------------------------------------
// dm holds one word per cell (demap_cell_soft: value + 8 of bit e in nibble e); the step codes are two words per byte time
// (8 bits per step: X value + 8 | Y value + 8 << 4, punctured positions = 8, see viterbi.cu).
__device__ __forceinline__ uint32_t inner_soft(const InnerMap &im, long long tbit) {
  long long b = tbit / im.m;
  int kbit = (int)(tbit - b * im.m);
  int sym = (int)(b / im.P);
  int i = (int)(b - (long long)sym * im.P);
  int blk = i / 126, ii = i - blk * 126;
  int half = im.m >> 1;
  int e = kbit / half + 2 * (kbit % half);
  int off = e == 0 ? 0 : e == 1 ? 63 : e == 2 ? 105 : e == 3 ? 42 : e == 4 ? 21 : 84;
  int w = ii - off;
  if (w < 0) w += 126;
  int x = blk * 126 + w;
  int q = (im.out_symidx[sym] & 1) ? im.Hinv[x] : im.H[x];
  uint32_t cell = reinterpret_cast<const uint32_t *>(im.dm)[(long long)im.out_src[sym] * im.P + q];
  return (cell >> (4 * e)) & 15u;
}

// ---- phase B of the soft kernel: two code words of a byte time from its window of <= 16 value bytes ----------------------
// Selector (PRMT) / keep mask / fill constant of the four steps of one half of a byte time: the half starts in puncturing
// phase PH, its values are the next bytes of the window; a punctured position becomes the value 0 (byte 8).
template <int RATE, int PH, int XY>   // XY: 0 = the X bits of the four steps, 1 = the Y bits
struct InnerSoftHalf {
  static constexpr unsigned compute(int what) {
    constexpr int K = rate_k(RATE);
    constexpr unsigned PX = rate_px(RATE), PY = rate_py(RATE);
    unsigned sel = 0, keep = 0, fill = 0;
    int pos = 0, ph = PH;
    for (int i = 0; i < 4; i++) {
      const bool hx = (PX >> ph) & 1u, hy = (PY >> ph) & 1u;
      const int px = pos, py = pos + (hx ? 1 : 0);
      const bool have = XY ? hy : hx;
      const int at = XY ? py : px;
      if (have) { sel |= (unsigned)at << (4 * i); keep |= 0xffu << (8 * i); }
      else fill |= 0x08u << (8 * i);
      pos += (hx ? 1 : 0) + (hy ? 1 : 0);
      ph = (ph + 1 == K) ? 0 : ph + 1;
    }
    return what == 0 ? sel : what == 1 ? keep : fill;
  }
  static constexpr unsigned sel = compute(0), keep = compute(1), fill = compute(2);
};

// bits consumed by 4 trellis steps that start in puncturing phase PH0
template <int RATE, int PH0>
__host__ __device__ constexpr int inner_half_values() {
  constexpr int K = rate_k(RATE);
  constexpr unsigned PX = rate_px(RATE), PY = rate_py(RATE);
  int n = 0, ph = PH0;
  for (int i = 0; i < 4; i++) {
    n += (int)((PX >> ph) & 1u) + (int)((PY >> ph) & 1u);
    ph = (ph + 1 == K) ? 0 : ph + 1;
  }
  return n;
}

// A[0..3] = the 16 window bytes (byte k of the window = byte k & 3 of A[k >> 2]); 4 steps -> one code word
template <int RATE, int PH, int OFF>
__device__ __forceinline__ uint32_t inner_soft_half(const uint32_t (&A)[4]) {
  constexpr int q = OFF >> 2, r = OFF & 3;
  // window bytes [OFF, OFF + 8) as two registers (OFF <= 8, so q + 2 <= 4; the register behind A[3] is never selected)
  const uint32_t a0 = A[q], a1 = q + 1 < 4 ? A[q + 1] : 0u, a2 = q + 2 < 4 ? A[q + 2] : 0u;
  const uint32_t b0 = r ? __byte_perm(a0, a1, 0x3210u + 0x1111u * r) : a0;
  const uint32_t b1 = r ? __byte_perm(a1, a2, 0x3210u + 0x1111u * r) : a1;
  using HX = InnerSoftHalf<RATE, PH, 0>;
  using HY = InnerSoftHalf<RATE, PH, 1>;
  const uint32_t xw = (__byte_perm(b0, b1, HX::sel) & HX::keep) | HX::fill;
  const uint32_t yw = (__byte_perm(b0, b1, HY::sel) & HY::keep) | HY::fill;
  return xw | (yw << 4);
}

template <int RATE, int PH0>
__device__ __forceinline__ uint2 inner_soft_code(const uint32_t (&A)[4]) {
  constexpr int K = rate_k(RATE);
  constexpr int H1 = inner_half_values<RATE, PH0>();
  return make_uint2(inner_soft_half<RATE, PH0, 0>(A), inner_soft_half<RATE, (PH0 + 4) % K, H1>(A));
}

// Same tiling and phases as rx_inner_codes_kernel, without the bit packing: the value bytes are read where they lie.
// Shared-memory layout of the value bytes: one pad word behind every 16 (byte i lives at i + 4 (i >> 6)).  The byte times of
// one puncturing class are 8 N = 16 .. 64 bytes apart, so without the pad a warp's window loads hit two to eight banks
// (measured: the whole kernel was those bank conflicts, 0.55 ms against 0.16 ms for the hard-decision kernel); a window is
// fetched as five aligned words and aligned with funnel shifts, the steps' values are picked with PRMT.
__device__ __forceinline__ int inner_soft_phys(int i) { return i + ((i >> 6) << 2); }

template <int RATE, int M>
__global__ void __launch_bounds__(256) rx_inner_soft_kernel(InnerMap im, uint2 *__restrict__ codes, int nbt) {
  extern __shared__ __align__(16) uint8_t s_val[];  // [tile cells * M + kInnerTail] values + 8 (padded), then the step codes of the tile
  constexpr int HALF = M / 2;
  const int G = kInnerTileCells / im.P;
  const int sym0 = blockIdx.x * G;
  const int nsym = min(G, im.n_out - sym0);
  const int ncell = nsym * im.P, nbits = ncell * M;
  const long long lo = (long long)sym0 * im.P * M, hi = lo + nbits;
  constexpr int kValBytes = (kInnerTileCells * M + kInnerTail + 63) / 64 * 68 + 64;   // padded size, + the words a window may read past the tail
  uint2 *s_codes = reinterpret_cast<uint2 *>(s_val + ((kValBytes + 15) & ~15));
  const uint32_t *dm32 = reinterpret_cast<const uint32_t *>(im.dm);
  for (int ls = 0; ls < nsym; ls++) {
    const int sym = sym0 + ls;
    const short *perm = (im.out_symidx[sym] & 1) ? im.Hinv : im.H;
    const uint32_t *row = dm32 + (long long)im.out_src[sym] * im.P;
    const int base = ls * im.P * M;
    for (int x = threadIdx.x; x < im.P; x += blockDim.x) {
      uint32_t cell = row[perm[x]];
      int blk126 = (x / 126) * 126, w = x - blk126;
      const int dst = base + blk126 * M;
#pragma unroll
      for (int e = 0; e < M; e++) {
        constexpr int kOff[6] = {0, 63, 105, 42, 21, 84};
        int ii = w + kOff[e];
        if (ii >= 126) ii -= 126;
        const int kbit = (e & 1) * HALF + (e >> 1);
        s_val[inner_soft_phys(dst + ii * M + kbit)] = (uint8_t)((cell >> (4 * e)) & 15u);
      }
    }
  }
  const long long total_bits = (long long)im.n_out * im.P * M;
  for (int b = threadIdx.x; b < kInnerTail + 32; b += blockDim.x)
    s_val[inner_soft_phys(nbits + b)] = (b < kInnerTail && hi + b < total_bits) ? (uint8_t)inner_soft(im, hi + b) : 8;
  __syncthreads();
  constexpr int K = rate_k(RATE), N = K + 1;
  const long long shift = im.shift_bits;
  long long jlo = lo > shift ? ((lo - shift) * K / N) / 8 - 2 : 0, jhi = hi > shift ? ((hi - shift) * K / N) / 8 - 2 : 0;
  if (jlo < 0) jlo = 0;
  if (jhi < 0) jhi = 0;
  while (inner_bits_before<RATE>(8 * jlo) + shift < lo) jlo++;
  while (inner_bits_before<RATE>(8 * jhi) + shift < hi) jhi++;
  if (jhi > nbt) jhi = nbt;
  const int nj = jhi > jlo ? (int)(jhi - jlo) : 0;
  const uint32_t *s_w = reinterpret_cast<const uint32_t *>(s_val);
  for (int c = 0; c < K; c++) {
    const long long t0 = 8 * (jlo + c);
    const int ph = (int)(t0 % K);
    const int local0 = (int)(inner_bits_before<RATE>(t0) + shift - lo);
    const int cnt = (nj - c + K - 1) / K;
#define INNER_CLASS(PH)                                                              \
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {                            \
      const int local = local0 + 8 * N * i, w0 = local >> 2, sh = 8 * (local & 3);   \
      uint32_t W[5], A[4];                                                           \
      _Pragma("unroll") for (int k = 0; k < 5; k++) W[k] = s_w[(w0 + k) + ((w0 + k) >> 4)]; \
      _Pragma("unroll") for (int k = 0; k < 4; k++) A[k] = __funnelshift_r(W[k], W[k + 1], sh); \
      s_codes[c + K * i] = inner_soft_code<RATE, (PH) % K>(A);                       \
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

// test tap: the soft values in the order of the Viterbi block's input stream, one signed byte per code bit
__global__ void rx_inner_soft_values_kernel(InnerMap im, int8_t *__restrict__ out, long long nbits) {
  long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbits) return;
  out[b] = (int8_t)((int)inner_soft(im, b) - 8);
}

// test tap: the bit_inner_deinterleaver output bytes (what the reference feeds its Viterbi block)
__global__ void rx_inner_bytes_kernel(InnerMap im, uint8_t *__restrict__ out, long long nbytes) {
  long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbytes) return;
  uint32_t v = 0;
  for (int k = 0; k < im.m; k++) v = (v << 1) | inner_bit(im, b * im.m + k);
  out[b] = (uint8_t)v;
}

// ---- energy_descramble (energy_descramble_impl.cc:108-174) ----------------------------------------------------------
// The block keeps ONE piece of state, d_index: where, relative to its read pointer, it last saw NSYNC (0xB8).  A call
// (the scheduler's smallest: noutput = 4 x 1504 bytes, >= 4 items of 8 packets visible) looks for NSYNC from d_index on
// in steps of one packet, up to 2 items ahead; if none is there it forgets the index and consumes 2 items, else it
// descrambles the 2 groups that start at the NSYNC packet (PRBS restarted per group, :146-165) and consumes 2 items.
// Nothing is re-checked inside a call, and the index survives from call to call - which is how the block re-locks
// after the outer deinterleaver was re-aligned by a mid-stream superframe_start.
//
// rx_descr_plan_kernel replays that state machine over all pending packets (pk = d_index in packets, 0..16): in the
// locked state the NSYNC tests of up to 4096 consecutive calls are made at once (one per thread and round); the
// first call whose test fails is replayed by thread 0 exactly as the reference loop does.  Output: plan[j] = first
// packet of the j-th emitted pair of groups.  rx_descramble_kernel then descrambles the pairs in parallel.
struct DescrState {
  int pk;                  // d_index / 188 (persists across calls of a stream)
  int n_pairs;             // pairs of groups emitted by the last plan
  int n_tail;              // 1: one more single group follows the pairs (end of stream only)
  long long items_used;    // 8-packet items consumed by the last plan
  long long first_packet1; // 1 + stream packet number of the first packet ever emitted, 0: none yet (info; all-zero = a new stream)
  long long packets_seen;  // packets consumed before the pending ones (stream position of pending packet 0)
};

// At the end of a stream (`end`) the locked state keeps going while whole pairs are left (the scheduler would not call the
// block without 4 items visible, :102-106, so the reference stops earlier: its file is a prefix of this), and one last
// complete group is delivered on its own.
// sb[p] = first byte of pending packet p (rx_descr_syncbytes_kernel): the plan only ever looks at those, and while the
// block is out of lock one thread replays the reference's search call by call - on a compact array its loads are cache hits
// instead of one 188-byte stride each (a capture that lost lock nine times spent 6.6 ms here before, profiles/README.md).
__global__ void rx_descr_syncbytes_kernel(const uint8_t *__restrict__ rs, long long npk, uint8_t *__restrict__ sb) {
  long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < npk) sb[p] = rs[p * 188];
}

__global__ void __launch_bounds__(1024) rx_descr_plan_kernel(const uint8_t *__restrict__ sb, long long npk, DescrState *st,
                                                             int *__restrict__ plan, long long plan_capacity, int end) {
  __shared__ int s_fail;
  const int t = threadIdx.x;
  long long i = 0;          // item index of the next call
  int pk = st->pk, np = 0;
  long long first_packet = st->first_packet1 - 1;
  const long long seen = st->packets_seen;
  __syncthreads();          // every thread has read the state before thread 0 may rewrite it (a short input leaves the loop below
                            // without meeting a barrier; the values would be the same ones, but a race is a race - found by the
                            // ThreadSanitizer build of tests/emul)
  const long long calls_strict = npk >= 32 ? (npk - 32) / 16 + 1 : 0;   // a call at item i needs items i .. i+3 visible
  long long calls_total = calls_strict;
  long long c = 0;          // calls made so far (i == 2 c)
  int tail = 0;
  while (np < plan_capacity) {
    // end of stream: calls whose pair is complete, as long as the index stays where it is (no search beyond the data)
    calls_total = calls_strict;
    if (end && pk < 16 && npk >= pk + 16 && (npk - pk - 16) / 16 + 1 > calls_total) calls_total = (npk - pk - 16) / 16 + 1;
    if (c >= calls_total) break;
    // ---- locked fast path: thread t tests the call c + t of this round
    if (t == 0) s_fail = 1 << 30;
    __syncthreads();
    if (pk < 16) {
      for (int r = 0; r < 4; r++) {
        long long cc = c + t + 1024LL * r;
        if (cc < calls_total && sb[16 * cc + pk] != 0xB8) atomicMin(&s_fail, t + 1024 * r);
      }
    }
    __syncthreads();
    long long good = 0;
    if (pk < 16) {
      good = s_fail;
      if (good > calls_total - c) good = calls_total - c;
      if (good > 4096) good = 4096;
      if (good > plan_capacity - np) good = plan_capacity - np;
    }
    for (long long j = t; j < good; j += 1024) plan[np + j] = (int)(16 * (c + j) + pk);
    if (good > 0 && first_packet < 0) first_packet = seen + 16 * c + pk;
    np += (int)good;
    c += good;
    __syncthreads();
    if (c >= calls_total || np >= plan_capacity) break;
    if (pk < 16 && good == 4096) continue;   // a full round without a failed test: next round
    if (pk < 16 && good > 0 && sb[16 * c + pk] == 0xB8) continue;   // the round ended at a limit, not at a failed test
    if (c >= calls_strict) break;            // a failed test beyond the strict range: the search would run past the data
    // ---- calls of the reference loop (:121-134), every thread the same scalar code: call after call while the block is out
    // of lock (no barrier in here), back to the parallel test above as soon as a call has found its inverted sync byte
    while (c < calls_strict && np < plan_capacity) {
      while (pk < 16 && sb[16 * c + pk] != 0xB8) pk++;
      bool found = pk < 16;
      if (!found) {
        pk = 0;             // d_index = 0, consume 2, no output
      } else {
        if (t == 0) plan[np] = (int)(16 * c + pk);
        if (first_packet < 0) first_packet = seen + 16 * c + pk;
        np++;
      }
      c++;
      if (found) break;
    }
  }
  i = 2 * c;
  if (end && pk < 16 && np < plan_capacity && 16 * c + pk + 8 <= npk && sb[16 * c + pk] == 0xB8) {
    if (t == 0) plan[np] = (int)(16 * c + pk);    // the last complete group, on its own
    if (first_packet < 0) first_packet = seen + 16 * c + pk;
    tail = 1;
  }
  if (t == 0) {
    st->pk = pk;
    st->n_pairs = np;
    st->n_tail = tail;
    st->items_used = i;
    st->first_packet1 = first_packet + 1;
    st->packets_seen = seen + 8 * i;
  }
}

// descrambles the planned pairs: 2 groups = 16 packets = 752 words per pair, one word per thread and step
__global__ void __launch_bounds__(256) rx_descramble_kernel(const uint8_t *__restrict__ rs, const DescrState *__restrict__ st,
                                                            const int *__restrict__ plan, const uint32_t *__restrict__ prbs,
                                                            uint8_t *__restrict__ ts, long long ts_capacity) {
  __shared__ uint32_t s_prbs[376];
  for (int i = threadIdx.x; i < 376; i += blockDim.x) s_prbs[i] = prbs[i];
  __syncthreads();
  long long npairs = st->n_pairs;
  int tail = st->n_tail;
  if (npairs * 3008 + tail * 1504 > ts_capacity) { npairs = ts_capacity / 3008; tail = 0; }
  const long long nwords = npairs * 752 + tail * 376;   // the single last group is the first half of pair `npairs`
  const bool aligned = ((((uintptr_t)rs) | ((uintptr_t)ts)) & 3u) == 0;
  for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (long long)gridDim.x * blockDim.x) {
    const long long pair = w / 752;
    const int k = (int)(w - pair * 752);
    const int kk = k >= 376 ? k - 376 : k;                       // word inside the group
    const uint8_t *src = rs + ((long long)plan[pair] * 188 + 4LL * k);
    uint32_t v = aligned ? *reinterpret_cast<const uint32_t *>(src)
                         : (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
    v ^= s_prbs[kk];                                             // :146-165
    if (kk % 47 == 0) v = (v & 0xffffff00u) | 0x47u;             // sync byte of every packet
    uint8_t *dst = ts + 4 * w;
    if (aligned) *reinterpret_cast<uint32_t *>(dst) = v;
    else { dst[0] = (uint8_t)v; dst[1] = (uint8_t)(v >> 8); dst[2] = (uint8_t)(v >> 16); dst[3] = (uint8_t)(v >> 24); }
  }
}

// Output symbols whose cells the Viterbi block has not consumed yet (less than one 768-block's worth) are carried to
// the next call: their demapped cells and symbol_index descriptors move to the front of the other buffer set.
__global__ void __launch_bounds__(256) rx_carry_kernel(int P, int first, int count, const uint8_t *__restrict__ dm, const int *__restrict__ src,
                                                       const int *__restrict__ symidx, uint8_t *__restrict__ dm_to, int *__restrict__ src_to,
                                                       int *__restrict__ symidx_to) {
  const int r = blockIdx.x;
  if (r >= count) return;
  const uint8_t *row = dm + (long long)src[first + r] * P;
  for (int i = threadIdx.x; i < P; i += blockDim.x) dm_to[(long long)r * P + i] = row[i];
  if (threadIdx.x == 0) { src_to[r] = r; symidx_to[r] = symidx[first + r]; }
}

}  // namespace



struct dvbt_b200_rx {
  int device = dvbt::current_device();
  dvbt_b200_rx_params par;
  dvbt::ModeTables tables;
  dvbt::DemapTable demap;
  dvbt_b200_viterbi *vit = nullptr;
  dvbt_b200_acq *acq = nullptr;
  cudaStream_t stream = nullptr;
  int fi_start = 3, rs_as_built = 0, sm_count = 148;
  bool soft = false;                 // soft-decision mode (dvbt_b200_rx_set_soft_decision): d_dm holds one word per cell
  int esz() const { return soft ? 4 : 1; }
  int k = 1, n = 2, m = 4, ntb = 5, vit_in_block = 0, vit_out_block = 0;
  dvbt::DevBuf d_state, d_fo, d_rot, d_mod, d_tps, d_vote, d_Y, d_rsst, d_ts, d_prbs, h_state, h_info, d_sync, h_sync, d_plan, d_dstate, d_syncb;
  dvbt_b200_rx_info info;
  cudaEvent_t ev[10];
  cudaStream_t side = nullptr;       // the demod scan runs here, beside the equalise kernel (demod.cuh: DemodBuffers::side)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_mid = nullptr;

  // ---- stream state: what the blocks of the flowgraph carry from one scheduler call to the next --------------------
  // The one-shot entry points are "reset, one piece, end of stream"; dvbt_b200_rx_stream_* feed a capture in pieces.
  int level = -1;                    // entry level of the stream: 0 capture file (10 Msps), 1 baseband, 2 post-FFT symbols
  bool fresh = true;                 // nothing pushed since the reset
  // front end (rational_resampler + multiply_const): input history + capture samples not yet used up
  dvbt::DevBuf d_file, d_file2;      // [kResHist history samples | pending capture samples]
  long long res_pend = 0;            // capture samples behind the history
  int res_hist = 0;                  // valid history samples in front of them
  // ofdm_sym_acquisition: baseband samples not yet consumed (the block's input buffer)
  dvbt::DevBuf d_samples, d_samples2;   // [kAcqHistory consumed samples (zeros at the start of a stream) | pending samples]
  long long bb_pend = 0;
  long long acq_total = 0;           // symbols acquisition has produced since the reset
  std::vector<long long> sync_abs;   // sync_start tags (absolute symbol numbers) whose symbol demod has not parsed yet
  // demod_reference_signals: the last symbol is parsed when its successor is visible (forecast 2 items)
  dvbt::DevBuf d_sym, d_sym2;        // [carried unparsed symbol | new symbols]
  int sym_carry = 0;
  long long parsed_total = 0;        // symbols parsed since the reset = absolute number of row 0 of d_sym
  // inner deinterleavers + viterbi_decoder: output symbols whose cells are not consumed yet (< one 768-block)
  dvbt::DevBuf d_dm[2], d_osym[2], d_osrc[2];   // demapped cells per row, symbol_index and row of each pending output symbol
  int cur = 0;                       // buffer set in use
  int out_carry = 0;                 // pending output symbols carried at the front of set `cur`
  long long cell_head = 0;           // cells of the pending output symbols already consumed by the Viterbi block
  std::vector<long long> sf_cells;   // superframe_start tags of demod, as cell offsets into the pending output symbols
  // convolutional_deinterleaver: the delay lines are never cleared (:109-120) = 2244 bytes of consumed stream history
  dvbt::DevBuf d_D, d_D2;            // [kOuterHist history | pending Viterbi output]
  long long d_pend = 0;              // bytes behind the history
  std::vector<long long> d_tags;     // superframe_start tags of the Viterbi block, offsets into the pending bytes
  // reed_solomon_dec output not yet consumed by energy_descramble (it wants 4 items visible)
  dvbt::DevBuf d_rs, d_rs2;
  long long rs_pend = 0;             // packets
  bool descr_pending_shift = false;  // the last plan's consumption has not been applied to d_rs yet
  long long ts_total = 0;
  // test taps of the last call
  long long tap_vit_off = 0, tap_vit_bytes = 0, tap_rs_first = 0, tap_rs_packets = 0, tap_first_row = 0, tap_rows_out = 0;
  int tap_set = 0, tap_carry = 0;
  bool tap_vit_copied = false;
  dvbt::DevBuf d_vit_tap;
};

namespace {

constexpr int kResHist = 64;          // >= 35 input samples of FIR history, kept 16-byte friendly
constexpr int kBbHist = dvbt::kAcqHistory;   // consumed baseband samples kept in front of the pending ones (acq.cu: ml_point)
constexpr int kOuterHist = 204 * 11;  // deepest delay line of the outer deinterleaver, in stream bytes
enum { kLevelFile = 0, kLevelBaseband = 1, kLevelFreq = 2 };

// event time that never leaves an error behind (an event that was not recorded in this call is not an error here)
bool elapsed_ms(float *ms, cudaEvent_t a, cudaEvent_t b) {
  if (cudaEventElapsedTime(ms, a, b) == cudaSuccess) return true;
  cudaGetLastError();
  return false;
}

// grows b to `bytes` keeping its first `keep` bytes
int reserve_keep(dvbt::DevBuf &b, size_t bytes, size_t keep, cudaStream_t st) {
  if (bytes <= b.cap) return 0;
  if (keep == 0 || !b.p) return b.reserve(bytes);
  dvbt::DevBuf bigger;
  int rc = bigger.reserve(bytes + bytes / 4);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaMemcpyAsync(bigger.p, b.p, keep, cudaMemcpyDeviceToDevice, st));
  DVBT_CUDA_TRY(dvbt::stream_wait(st));
  b.release();
  b = bigger;
  return 0;
}

// moves bytes [from, from + n) of a to the front of b (+ dst_off) and swaps the two, i.e. "drop the front of the buffer"
int shift_front(dvbt::DevBuf &a, dvbt::DevBuf &b, size_t from, size_t n, size_t dst_off, size_t min_cap, cudaStream_t st) {
  int rc = b.reserve(min_cap > dst_off + n ? min_cap : dst_off + n);
  if (rc) return rc;
  if (n) DVBT_CUDA_TRY(cudaMemcpyAsync((char *)b.p + dst_off, (const char *)a.p + from, n, cudaMemcpyDeviceToDevice, st));
  std::swap(a, b);
  return 0;
}

void rx_stream_reset(dvbt_b200_rx *h) {
  h->level = -1;
  h->fresh = true;
  h->res_pend = 0; h->res_hist = 0;
  h->bb_pend = 0; h->acq_total = 0; h->sync_abs.clear();
  h->sym_carry = 0; h->parsed_total = 0;
  h->out_carry = 0; h->cell_head = 0; h->sf_cells.clear();
  h->d_pend = 0; h->d_tags.clear();
  h->rs_pend = 0; h->descr_pending_shift = false;
  h->ts_total = 0;
}

void rx_info_reset(dvbt_b200_rx *h) {   // at the start of a stream: the info of the previous one stays readable until then
  memset(&h->info, 0, sizeof h->info);
  h->info.first_symbol = -1;
  h->info.first_packet = -1;
  h->info.acq_lost_at = -1;
}

int launch_inner(dvbt_b200_rx *h, const InnerMap &im, int rows, uint32_t *codes, int nbt) {
  const dvbt::ModeDev &md = h->tables.dev;
  const int G = kInnerTileCells / md.P;
  unsigned grid = (unsigned)((rows + G - 1) / G);
  size_t smem = (size_t)((kInnerTileCells * h->m + kInnerTail + 15) & ~15) + (size_t)(kInnerTileCells * h->m + kInnerTail) / 8 + 16;
  cudaStream_t st = h->stream;
  if (h->soft) {
    // value bytes + the tile's step codes (at most one byte time per 8 N / K >= 64 / 7 stream bits), 8 bytes each
    smem = (size_t)((((kInnerTileCells * h->m + kInnerTail + 63) / 64 * 68 + 64) + 15) & ~15) + ((size_t)(kInnerTileCells * h->m) * 7 / 64 + 8) * 8;
    uint2 *codes2 = reinterpret_cast<uint2 *>(codes);
#define RX_SOFT_LAUNCH(R, M) do { \
      DVBT_CUDA_TRY(cudaFuncSetAttribute(rx_inner_soft_kernel<R, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      rx_inner_soft_kernel<R, M><<<grid, 256, smem, st>>>(im, codes2, nbt); } while (0)
#define RX_SOFT_RATE(R) do { if (h->m == 2) RX_SOFT_LAUNCH(R, 2); else if (h->m == 4) RX_SOFT_LAUNCH(R, 4); else RX_SOFT_LAUNCH(R, 6); } while (0)
    switch (h->par.code_rate) {
      case 0: RX_SOFT_RATE(0); break;
      case 1: RX_SOFT_RATE(1); break;
      case 2: RX_SOFT_RATE(2); break;
      case 3: RX_SOFT_RATE(3); break;
      default: RX_SOFT_RATE(4); break;
    }
#undef RX_SOFT_RATE
#undef RX_SOFT_LAUNCH
    dvbt::count_launch();
    DVBT_CUDA_TRY(cudaGetLastError());
    return 0;
  }
#define RX_INNER_LAUNCH(R, M) rx_inner_codes_kernel<R, M><<<grid, 256, smem, st>>>(im, codes, nbt)
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
  return 0;
}

// ---- stages 3..6: pending output symbols -> Viterbi -> outer deinterleaver + RS -> descrambler ----------------------
// `nrows` output symbols are pending in buffer set h->cur (rows/descriptors 0..nrows-1), tags in h->sf_cells.
int rx_back_end(dvbt_b200_rx *h, int nrows, bool end, uint8_t *ts_host, uint8_t *ts_dev, size_t ts_capacity, size_t *ts_bytes) {
  const dvbt::ModeDev &md = h->tables.dev;
  cudaStream_t st = h->stream;
  int rc;
  const long long P = md.P, nsymb = h->vit_in_block;
  uint8_t *dm = h->d_dm[h->cur].as<uint8_t>();
  int *osym = h->d_osym[h->cur].as<int>(), *osrc = h->d_osrc[h->cur].as<int>();

  // ---- viterbi_decoder, one 768-block per scheduler call (viterbi_decoder_impl.cc:198-229): a superframe_start tag
  // inside the block's window resets the decoder; if it is not on the window's first item everything in front of it is
  // consumed undecoded.  Consecutive blocks without a tag are decoded as one chunk-parallel run.
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[2], st));
  const long long cells_end = (long long)nrows * P;
  long long pos = h->cell_head;
  // room for everything this call can produce
  {
    long long max_new = (cells_end - pos) / nsymb * h->vit_out_block + 64;
    if ((rc = reserve_keep(h->d_D, (size_t)(kOuterHist + h->d_pend + max_new), (size_t)(kOuterHist + h->d_pend), st))) return rc;
  }
  h->tap_vit_off = kOuterHist + h->d_pend;
  h->tap_vit_bytes = 0;
  h->tap_vit_copied = false;
  dvbt::vit_stream_accumulate_stats(h->vit, true);
  float ms_inner = 0.f;
  int nruns = 0;
  while (pos + nsymb <= cells_end) {
    // lowest tag at or after pos (tags behind the read position were passed over inside a decoded block: never seen again)
    while (!h->sf_cells.empty() && h->sf_cells.front() < pos) h->sf_cells.erase(h->sf_cells.begin());
    if (!h->sf_cells.empty() && h->sf_cells.front() < pos + nsymb) {
      dvbt::vit_stream_reset(h->vit);                               // :217-221
      const long long t = h->sf_cells.front();
      h->sf_cells.erase(h->sf_cells.begin());
      h->info.n_superframe_start++;
      if (t > pos) { pos = t; continue; }                          // :223-228 consume up to the tag, produce nothing
    }
    // blocks from pos on until a window that holds a tag, or the end of the input
    long long nb = 1;
    {
      long long next_tag = -1;
      for (long long t : h->sf_cells) if (t >= pos + nsymb) { next_tag = t; break; }
      long long room = (cells_end - pos) / nsymb;
      if (next_tag >= 0) { long long upto = (next_tag - pos) / nsymb; if (upto < room) room = upto; }   // window [pos + j nsymb, +nsymb) is tag free for j < upto
      if (room > 1) nb = room;
    }
    const long long new_bt = nb * h->vit_out_block;
    if (new_bt >= (1LL << 30)) { set_error("rx: batch too large (%lld byte times)", new_bt); return DVBT_B200_EINVAL; }
    uint32_t *codes = nullptr;
    if ((rc = dvbt::vit_stream_codes(h->vit, (int)new_bt, &codes))) return rc;
    const int row0 = (int)(pos / P);
    const int rows = (int)((pos + nb * nsymb + P - 1) / P) - row0;
    InnerMap im{dm, osrc + row0, osym + row0, md.H, md.Hinv, md.P, h->m, nrows - row0, (int)((pos - (long long)row0 * P) * h->m)};
    if ((rc = launch_inner(h, im, rows, codes, (int)new_bt))) return rc;
    if (nruns == 0) DVBT_CUDA_TRY(cudaEventRecord(h->ev[3], st));
    size_t nprod = 0;
    bool first = false;
    if ((rc = dvbt::vit_stream_decode(h->vit, (int)new_bt, h->d_D.as<uint8_t>() + kOuterHist + h->d_pend, &nprod, &first, !end))) return rc;
    if (first) h->d_tags.push_back(h->d_pend);                      // :298-312 superframe_start on the first byte produced after a reset
    h->d_pend += (long long)nprod;
    h->tap_vit_bytes += (long long)nprod;
    h->info.viterbi_bytes += (long long)nprod;
    pos += nb * nsymb;
    nruns++;
  }
  if (nruns == 0) DVBT_CUDA_TRY(cudaEventRecord(h->ev[3], st));
  h->info.n_viterbi_runs += nruns;
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[4], st));
  (void)ms_inner;
  // carry the output symbols that still hold unconsumed cells
  {
    const int drop = (int)(pos / P) < nrows ? (int)(pos / P) : nrows;
    const int keep = end ? 0 : nrows - drop;
    if (keep > 0) {
      const int o = h->cur ^ 1;
      if ((rc = h->d_dm[o].reserve((size_t)keep * P * h->esz())) || (rc = h->d_osym[o].reserve((size_t)keep * 4)) || (rc = h->d_osrc[o].reserve((size_t)keep * 4))) return rc;
      rx_carry_kernel<<<keep, 256, 0, st>>>(md.P * h->esz(), drop, keep, dm, osrc, osym, h->d_dm[o].as<uint8_t>(), h->d_osrc[o].as<int>(), h->d_osym[o].as<int>());
      dvbt::count_launch();
      DVBT_CUDA_TRY(cudaGetLastError());
      h->cur = o;
    }
    h->out_carry = keep;
    h->cell_head = pos - (long long)drop * P;
    for (auto &t : h->sf_cells) t -= (long long)drop * P;
  }

  // ---- convolutional_deinterleaver (2 items of 1632 bytes per call, :93-150) + reed_solomon_dec: the same tag rule,
  // delay lines that are never cleared = the gather reaches 2244 bytes back into the consumed stream
  h->tap_rs_packets = 0;
  if (h->descr_pending_shift) {
    // apply the previous call's descrambler consumption to the pending RS packets
    const DescrState *ds = h->h_info.as<DescrState>();
    long long used = ds->items_used * 8;
    if (used > h->rs_pend) used = h->rs_pend;
    if (used > 0) {
      if ((rc = shift_front(h->d_rs, h->d_rs2, (size_t)used * 188, (size_t)(h->rs_pend - used) * 188, 0, 0, st))) return rc;
      h->rs_pend -= used;
    }
    h->descr_pending_shift = false;
  }
  h->tap_rs_first = h->rs_pend;
  {
    long long dpos = 0;
    uint8_t *D = h->d_D.as<uint8_t>() + kOuterHist;
    while (dpos + 3264 <= h->d_pend) {
      while (!h->d_tags.empty() && h->d_tags.front() < dpos) h->d_tags.erase(h->d_tags.begin());
      if (!h->d_tags.empty() && h->d_tags.front() < dpos + 3264 && h->d_tags.front() > dpos) {
        // consume up to the tag, produce nothing (:109-120): those bytes never enter the delay lines
        const long long t = h->d_tags.front(), gap = t - dpos, tail = h->d_pend - t;
        if (!h->tap_vit_copied && h->tap_vit_off >= 0 && h->tap_vit_bytes > 0) {
          // test tap: the Viterbi output of this call as it was produced (the bytes in front of the tag are about to go)
          if ((rc = h->d_vit_tap.reserve((size_t)h->tap_vit_bytes))) return rc;
          DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_vit_tap.p, h->d_D.as<uint8_t>() + h->tap_vit_off, (size_t)h->tap_vit_bytes, cudaMemcpyDeviceToDevice, st));
          h->tap_vit_copied = true;
        }
        if ((rc = h->d_D2.reserve((size_t)tail + 16))) return rc;
        DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_D2.p, D + t, (size_t)tail, cudaMemcpyDeviceToDevice, st));
        DVBT_CUDA_TRY(cudaMemcpyAsync(D + dpos, h->d_D2.p, (size_t)tail, cudaMemcpyDeviceToDevice, st));
        h->d_pend -= gap;
        for (auto &x : h->d_tags) x -= gap;
        continue;
      }
      if (!h->d_tags.empty() && h->d_tags.front() == dpos) h->d_tags.erase(h->d_tags.begin());
      long long nw = (h->d_pend - dpos) / 3264;
      for (long long t : h->d_tags) if (t >= dpos + 3264) { long long upto = (t - dpos) / 3264; if (upto < nw) nw = upto; break; }
      if (nw < 1) nw = 1;
      const long long npk = 16 * nw;
      if ((rc = reserve_keep(h->d_rs, (size_t)(h->rs_pend + npk) * 188 + 64, (size_t)h->rs_pend * 188, st)) ||
          (rc = reserve_keep(h->d_rsst, (size_t)(h->rs_pend + npk) * 4, (size_t)h->rs_pend * 4, st)))
        return rc;
      // history_bytes: 2244 bytes of consumed stream sit in front of the pending ones (zeros at the start of a stream)
      rc = dvbt::rs_launch(D + dpos, h->d_rs.as<uint8_t>() + h->rs_pend * 188, h->d_rsst.as<int>() + h->rs_pend, npk, h->rs_as_built, h->sm_count, st,
                           h->d_pend - dpos, kOuterHist + dpos);
      if (rc) return rc;
      h->rs_pend += npk;
      h->tap_rs_packets += npk;
      h->info.rs_packets += npk;
      dpos += 3264 * nw;
    }
    // keep the last 2244 consumed bytes + what is still pending at the front
    if (!end && dpos > 0) {
      if ((rc = shift_front(h->d_D, h->d_D2, (size_t)dpos, (size_t)(kOuterHist + h->d_pend - dpos), 0, 0, st))) return rc;
      h->tap_vit_off = -1;   // the tap region moved
      h->d_pend -= dpos;
      for (auto &x : h->d_tags) x -= dpos;
    }
  }
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[5], st));

  // ---- energy_descramble
  uint8_t *ts_out = ts_dev;
  size_t cap = ts_capacity;
  const long long max_pairs = h->rs_pend / 16 + 2;
  if (!ts_out) {
    if ((rc = h->d_ts.reserve((size_t)max_pairs * 3008))) return rc;
    ts_out = h->d_ts.as<uint8_t>();
    if (cap > (size_t)max_pairs * 3008 || ts_host == nullptr) cap = (size_t)max_pairs * 3008;
  }
  const size_t ts_need = (size_t)(h->rs_pend / 16) * 3008 + ((h->rs_pend % 16) >= 8 ? 1504 : 0);
  if (ts_capacity < ts_need) {
    // the descrambler's consumption is decided on the device: refuse up front what might not fit rather than drop groups
    set_error("rx: ts_capacity %zu is below what this call can deliver (%lld pending packets: up to %lld bytes)", ts_capacity, h->rs_pend,
              (long long)ts_need);
    return DVBT_B200_ENOSPC;
  }
  if ((rc = h->d_plan.reserve((size_t)max_pairs * 4))) return rc;
  if ((rc = h->d_syncb.reserve((size_t)h->rs_pend + 16))) return rc;
  if (h->rs_pend > 0) rx_descr_syncbytes_kernel<<<(unsigned)((h->rs_pend + 255) / 256), 256, 0, st>>>(h->d_rs.as<uint8_t>(), h->rs_pend, h->d_syncb.as<uint8_t>());
  rx_descr_plan_kernel<<<1, 1024, 0, st>>>(h->d_syncb.as<uint8_t>(), h->rs_pend, h->d_dstate.as<DescrState>(), h->d_plan.as<int>(), max_pairs, end ? 1 : 0);
  dvbt::count_launch(h->rs_pend > 0 ? 1 : 0);
  {
    long long blocks = max_pairs * 752 / 256 / 4 + 1;   // ~4 words per thread
    unsigned grid = (unsigned)(blocks < 8LL * h->sm_count ? blocks : 8LL * h->sm_count);
    rx_descramble_kernel<<<grid, 256, 0, st>>>(h->d_rs.as<uint8_t>(), h->d_dstate.as<DescrState>(), h->d_plan.as<int>(), h->d_prbs.as<uint32_t>(), ts_out,
                                               (long long)cap);
  }
  dvbt::count_launch(2);
  DVBT_CUDA_TRY(cudaGetLastError());
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[6], st));
  DVBT_CUDA_TRY(cudaMemcpyAsync(h->h_info.p, h->d_dstate.p, sizeof(DescrState), cudaMemcpyDeviceToHost, st));
  if ((rc = dvbt::vit_collect_stats(h->vit))) return rc;  // synchronises the stream
  dvbt::vit_stream_accumulate_stats(h->vit, false);
  const DescrState *ds = h->h_info.as<DescrState>();
  h->descr_pending_shift = true;
  long long nbytes = (long long)ds->n_pairs * 3008 + (long long)ds->n_tail * 1504;
  if (nbytes > (long long)cap) nbytes = (long long)(cap / 3008) * 3008;
  h->info.first_packet = ds->first_packet1 - 1;
  h->info.ts_bytes = nbytes;
  h->ts_total += nbytes;
  h->info.ts_total = h->ts_total;
  if (ts_host && nbytes > 0) {
    DVBT_CUDA_TRY(cudaMemcpyAsync(ts_host, ts_out, (size_t)nbytes, cudaMemcpyDeviceToHost, st));
    DVBT_CUDA_TRY(dvbt::stream_wait(st));
  }
  if (ts_bytes) *ts_bytes = (size_t)nbytes;
  float ms;
  if (elapsed_ms(&ms, h->ev[2], h->ev[3])) h->info.ms_inner = ms;
  if (elapsed_ms(&ms, h->ev[3], h->ev[4])) h->info.ms_viterbi = ms;
  if (elapsed_ms(&ms, h->ev[4], h->ev[5])) h->info.ms_rs = ms;
  if (elapsed_ms(&ms, h->ev[5], h->ev[6])) h->info.ms_descramble = ms;
  long long chunks = 0, rep = 0;
  float acs = 0;
  dvbt_b200_viterbi_last_stats(h->vit, &chunks, &rep, &acs);
  h->info.ms_viterbi_acs = acs;
  h->info.viterbi_repaired += rep;
  return 0;
}

// ---- stage 2: demod_reference_signals (+ fused demap) over the pending symbols, then the back end -------------------
// X: rows [0, nsym) = [carried unparsed symbol | new symbols] on the device; absolute number of row 0 = h->parsed_total
int rx_from_symbols(dvbt_b200_rx *h, const float2 *X, size_t nsym, bool x_is_internal, bool end, uint8_t *ts_host, uint8_t *ts_dev,
                    size_t ts_capacity, size_t *ts_bytes, int keep_cells) {
  const dvbt::ModeDev &md = h->tables.dev;
  cudaStream_t st = h->stream;
  int rc;
  const size_t nparse = nsym >= 2 ? nsym - 1 : 0;
  const int c = h->out_carry;
  int n_out = 0;
  h->tap_first_row = 0; h->tap_rows_out = 0; h->tap_set = h->cur; h->tap_carry = c;
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[0], st));
  if (nparse > 0) {
    if ((rc = h->d_fo.reserve(nparse * 4)) || (rc = h->d_rot.reserve(nparse * 8)) || (rc = h->d_mod.reserve(nparse * 4)) ||
        (rc = h->d_tps.reserve(nparse * md.ntps * 8)) || (rc = h->d_vote.reserve(nparse * 4)))
      return rc;
    if ((rc = reserve_keep(h->d_dm[h->cur], (size_t)(c + nparse) * md.P * h->esz(), (size_t)c * md.P * h->esz(), st)) ||
        (rc = reserve_keep(h->d_osym[h->cur], (size_t)(c + nparse) * 4, (size_t)c * 4, st)) ||
        (rc = reserve_keep(h->d_osrc[h->cur], (size_t)(c + nparse) * 4, (size_t)c * 4, st)))
      return rc;
    if (keep_cells && (rc = h->d_Y.reserve(nparse * md.P * 8))) return rc;
    // sync_start tags on symbols parsed in this call (absolute -> row)
    std::vector<int> sync_rows;
    {
      size_t kept = 0;
      for (long long a : h->sync_abs) {
        long long row = a - h->parsed_total;
        if (row < 0) row = 0;                                   // cannot happen; a tag is never behind the read position
        if (row < (long long)nparse) sync_rows.push_back((int)row);
        else h->sync_abs[kept++] = a;                           // its symbol is parsed by a later call
      }
      h->sync_abs.resize(kept);
    }
    h->info.n_sync_start += (long long)sync_rows.size();
    if (!sync_rows.empty()) {
      // through a pinned buffer of the handle: it is not written again before the call's final stream synchronisation
      h->h_sync.host = true;
      if ((rc = h->d_sync.reserve(sync_rows.size() * 4)) || (rc = h->h_sync.reserve(sync_rows.size() * 4))) return rc;
      memcpy(h->h_sync.p, sync_rows.data(), sync_rows.size() * 4);
      DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_sync.p, h->h_sync.p, sync_rows.size() * 4, cudaMemcpyHostToDevice, st));
    }
    dvbt::DemodBuffers b{h->d_fo.as<int>(), h->d_rot.as<float2>(), h->d_mod.as<int>(), h->d_tps.as<float2>(), h->d_vote.as<int>(),
                         h->d_osym[h->cur].as<int>() + c, h->d_osrc[h->cur].as<int>() + c, h->ev[8], h->ev[9], h->side, h->ev_fork, h->ev_join, h->ev_mid};
    rc = dvbt::demod_run(md, &h->demap, X, (int)nparse, b, h->d_state.as<dvbt::DemodState>(), h->fi_start, 0,
                         keep_cells ? h->d_Y.as<float2>() : nullptr, h->soft ? nullptr : h->d_dm[h->cur].as<uint8_t>() + (size_t)c * md.P, st,
                         sync_rows.empty() ? nullptr : h->d_sync.as<int>(), (int)sync_rows.size(), c,
                         h->soft ? h->d_dm[h->cur].as<uint32_t>() + (size_t)c * md.P : nullptr);
    if (rc) return rc;
    DVBT_CUDA_TRY(cudaEventRecord(h->ev[1], st));
    DVBT_CUDA_TRY(cudaMemcpyAsync(h->h_state.p, h->d_state.p, sizeof(dvbt::DemodState), cudaMemcpyDeviceToHost, st));
    // carry the unparsed last symbol while the scan result travels
    if (!end) {
      if (x_is_internal) {
        if ((rc = shift_front(h->d_sym, h->d_sym2, nparse * md.N * 8, (size_t)md.N * 8, 0, 0, st))) return rc;
      } else {
        if ((rc = h->d_sym.reserve((size_t)md.N * 8))) return rc;
        DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_sym.p, X + nparse * md.N, (size_t)md.N * 8, cudaMemcpyDeviceToDevice, st));
      }
      h->sym_carry = 1;
    }
    DVBT_CUDA_TRY(dvbt::stream_wait(st));
    const dvbt::DemodState *S = h->h_state.as<dvbt::DemodState>();
    if (S->n_sf > dvbt::kMaxSfTags) {
      set_error("rx: %d re-synchronisations in one batch (at most %d): feed the capture in smaller pieces", S->n_sf, dvbt::kMaxSfTags);
      return DVBT_B200_EINVAL;
    }
    n_out = S->n_out;
    for (int i = 0; i < S->n_sf; i++) h->sf_cells.push_back((long long)(c + S->sf_at[i]) * md.P);
    h->info.symbols_parsed += (long long)nparse;
    if (h->info.first_symbol < 0 && S->first_out >= 0) h->info.first_symbol = h->parsed_total + S->first_out;
    h->info.symbols_out += n_out;
    h->tap_first_row = S->first_out < 0 ? 0 : S->first_out;
    h->tap_rows_out = n_out;
    h->parsed_total += (long long)nparse;
  } else {
    DVBT_CUDA_TRY(cudaEventRecord(h->ev[1], st));
    if (!end && nsym == 1 && !x_is_internal) {
      if ((rc = h->d_sym.reserve((size_t)md.N * 8))) return rc;
      DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_sym.p, X, (size_t)md.N * 8, cudaMemcpyDeviceToDevice, st));
    }
    if (!end) h->sym_carry = (int)nsym;
  }
  rc = rx_back_end(h, c + n_out, end, ts_host, ts_dev, ts_capacity, ts_bytes);
  float ms;
  if (elapsed_ms(&ms, h->ev[0], h->ev[1])) h->info.ms_demod = ms;
  if (nparse > 0 && elapsed_ms(&ms, h->ev[8], h->ev[9])) h->info.ms_equalise = ms;
  return rc;
}

// ---- stage 1: ofdm_sym_acquisition + FFT over the pending baseband samples ------------------------------------------
// x: [0, n) = [samples acquisition has not consumed yet | new samples] on the device
int rx_from_baseband(dvbt_b200_rx *h, const float2 *x, size_t n, bool x_is_internal, bool end, uint8_t *ts_host, uint8_t *ts_dev,
                     size_t ts_capacity, size_t *ts_bytes, int keep_cells) {
  const dvbt::ModeDev &md = h->tables.dev;
  cudaStream_t st = h->stream;
  const long long total = md.N + md.cp;
  long long cap_syms = (long long)(n / (size_t)total) + 2;
  const int carry = h->sym_carry;
  int rc = reserve_keep(h->d_sym, (size_t)(carry + cap_syms) * md.N * 8, (size_t)carry * md.N * 8, st);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaEventRecord(h->ev[7], st));
  dvbt::AcqResult ar;
  std::vector<long long> sync_at;
  rc = dvbt::acq_run_simple(h->acq, x, (long long)n, h->d_sym.as<float2>() + (size_t)carry * md.N, cap_syms, 1, &ar, &sync_at, x_is_internal ? kBbHist : 0);
  if (rc) return rc;
  cudaEvent_t ev_acq_end = h->ev[0];   // recorded next by rx_from_symbols: end of the acquisition stage
  for (long long o : sync_at) {
    long long a = h->acq_total + o;
    if (h->sync_abs.empty() || h->sync_abs.back() != a) h->sync_abs.push_back(a);
  }
  if (ar.lost_at >= 0 && h->info.acq_lost_at < 0) h->info.acq_lost_at = h->acq_total + ar.lost_at;
  h->acq_total += ar.n_out;
  h->info.acq_symbols += ar.n_out;
  h->info.acq_cp_start = ar.cp_start;
  h->info.acq_run_symbols += ar.n_run;
  h->info.acq_single_symbols += ar.n_single;
  h->info.acq_sequential_symbols += ar.n_seq;
  // the samples acquisition has not consumed stay in its input buffer
  if (!end) {
    const long long left = (long long)n - ar.consumed;
    if (left < 0) { set_error("rx: acquisition consumed more than it was given"); return DVBT_B200_ECUDA; }
    if (x_is_internal) {
      // x = d_samples + kBbHist: the kBbHist samples in front of the new read position travel along
      if ((rc = shift_front(h->d_samples, h->d_samples2, (size_t)ar.consumed * 8, (size_t)(kBbHist + left) * 8, 0, 0, st))) return rc;
    } else {
      // the caller's buffer (first piece of a stream read in place): history = its last consumed samples, zeros in front
      const long long have = ar.consumed < kBbHist ? ar.consumed : kBbHist;
      if ((rc = h->d_samples.reserve((size_t)(kBbHist + left) * 8 + 16))) return rc;
      if (have < kBbHist) DVBT_CUDA_TRY(cudaMemsetAsync(h->d_samples.p, 0, (size_t)(kBbHist - have) * 8, st));
      if (have + left) DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_samples.as<float2>() + (kBbHist - have), x + ar.consumed - have, (size_t)(have + left) * 8, cudaMemcpyDeviceToDevice, st));
    }
    h->bb_pend = left;
  }
  rc = rx_from_symbols(h, h->d_sym.as<float2>(), (size_t)(carry + ar.n_out), true, end, ts_host, ts_dev, ts_capacity, ts_bytes, keep_cells);
  float ms;
  if (elapsed_ms(&ms, h->ev[7], ev_acq_end)) h->info.ms_acq_fft = ms;
  h->info.ms_fft = dvbt::acq_last_fft_ms(h->acq);
  return rc;
}

// ---- stage 0: rational_resampler_ccc(64,70) + multiply_const over the pending capture samples ------------------------
// x: capture samples of this piece on the device.  Streaming pieces are appended behind the FIR history in d_file by
// the caller (x == nullptr); a fresh one-piece run reads the caller's buffer in place.
int rx_from_file(dvbt_b200_rx *h, const float2 *x_direct, size_t n_direct, float gain, bool end, uint8_t *ts_host, uint8_t *ts_dev,
                 size_t ts_capacity, size_t *ts_bytes) {
  cudaStream_t st = h->stream;
  int rc;
  const float2 *x;
  long long nin;
  int nhist;
  if (x_direct) { x = x_direct; nin = (long long)n_direct; nhist = 0; }
  else { x = h->d_file.as<float2>() + kResHist; nin = h->res_pend; nhist = h->res_hist; }
  // outputs whose newest input sample is available; a continuing stream stops at a multiple of 32 outputs so that the
  // next piece starts on input sample 35 q, where the polyphase pattern restarts
  long long nout = dvbt::resample_out_count(nin);
  if (!end) nout = nout / 32 * 32;
  if ((rc = reserve_keep(h->d_samples, (size_t)(kBbHist + h->bb_pend + nout) * 8 + 16, (size_t)(kBbHist + h->bb_pend) * 8, st))) return rc;
  cudaEvent_t e0, e1;
  DVBT_CUDA_TRY(cudaEventCreate(&e0));
  DVBT_CUDA_TRY(cudaEventCreate(&e1));
  cudaEventRecord(e0, st);
  rc = dvbt::resample_launch(x, nin, h->d_samples.as<float2>() + kBbHist + h->bb_pend, nout, gain, st, nhist);
  cudaEventRecord(e1, st);
  if (!rc && !end) {
    // keep the FIR history + the inputs of the outputs not produced yet
    const long long used = nout / 32 * 35;            // input samples the produced outputs have moved past
    const long long from = used - kResHist;           // may reach into the old history (negative: before x)
    const long long keep = nin - from;                // history + pending
    if ((rc = h->d_file2.reserve((size_t)(keep > kResHist ? keep : kResHist) * 8 + 16))) return rc;
    // d_file holds kResHist samples in front of x: index `from` >= -kResHist is always inside it (a piece read in
    // place from the caller's buffer is always the whole stream, so it never gets here)
    DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_file2.p, x + from, (size_t)keep * 8, cudaMemcpyDeviceToDevice, st));
    std::swap(h->d_file, h->d_file2);
    h->res_hist = kResHist;
    h->res_pend = nin - used;
  }
  if (!rc) {
    const size_t nbb = (size_t)(h->bb_pend + nout);
    rc = rx_from_baseband(h, h->d_samples.as<float2>() + kBbHist, nbb, true, end, ts_host, ts_dev, ts_capacity, ts_bytes, 0);
  }
  float ms = 0;
  if (elapsed_ms(&ms, e0, e1)) h->info.ms_resample = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

// One piece of a stream at entry level `level`.  data: host pointer (host = true) or device pointer.
int rx_push(dvbt_b200_rx *h, int level, const void *data, size_t count, float gain, bool host, bool end, uint8_t *ts_host, uint8_t *ts_dev,
            size_t ts_capacity, size_t *ts_bytes) {
  const dvbt::ModeDev &md = h->tables.dev;
  cudaStream_t st = h->stream;
  if (ts_bytes) *ts_bytes = 0;
  const bool first = h->fresh;
  if (!first && h->level != level) { set_error("rx: a stream keeps its entry level (%d) until it ends or is reset; got %d", h->level, level); return DVBT_B200_EINVAL; }
  if (first) {
    // decoder state of a new stream
    int rc;
    rx_stream_reset(h);
    rx_info_reset(h);
    DVBT_CUDA_TRY(cudaMemsetAsync(h->d_state.p, 0, sizeof(dvbt::DemodState), st));
    DVBT_CUDA_TRY(cudaMemsetAsync(h->d_dstate.p, 0, sizeof(DescrState), st));   // NSYNC index 0, nothing emitted yet
    if ((rc = h->d_D.reserve(kOuterHist + 4096))) return rc;
    DVBT_CUDA_TRY(cudaMemsetAsync(h->d_D.p, 0, kOuterHist, st));   // the delay lines start zeroed (convolutional_deinterleaver_impl.cc:62-64)
    dvbt::vit_stream_reset(h->vit);
    if (level <= kLevelBaseband && (rc = dvbt::acq_reset(h->acq))) return rc;
    h->sync_abs.clear();
    if (level == kLevelFreq) h->sync_abs.push_back(0);   // what acquisition sends with its first symbol (ofdm_sym_acquisition_impl.cc:507)
    h->level = level;
    h->fresh = false;
  }
  int rc;
  const bool in_place = first && end && !host;   // a whole capture already on the device: no staging copy
  if (level == kLevelFreq) {
    const int carry = h->sym_carry;
    if (in_place) return rx_from_symbols(h, (const float2 *)data, count, false, end, ts_host, ts_dev, ts_capacity, ts_bytes, 0);
    if ((rc = reserve_keep(h->d_sym, (size_t)(carry + count) * md.N * 8 + 16, (size_t)carry * md.N * 8, st))) return rc;
    if (count) DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_sym.as<float2>() + (size_t)carry * md.N, data, count * md.N * 8, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    return rx_from_symbols(h, h->d_sym.as<float2>(), (size_t)carry + count, true, end, ts_host, ts_dev, ts_capacity, ts_bytes, host ? 1 : 0);
  }
  if (level == kLevelBaseband) {
    if (in_place) return rx_from_baseband(h, (const float2 *)data, count, false, end, ts_host, ts_dev, ts_capacity, ts_bytes, 0);
    if (first) {
      if ((rc = h->d_samples.reserve((size_t)(kBbHist + count) * 8 + 16))) return rc;
      DVBT_CUDA_TRY(cudaMemsetAsync(h->d_samples.p, 0, (size_t)kBbHist * 8, st));   // nothing in front of the stream's first sample
    }
    if ((rc = reserve_keep(h->d_samples, (size_t)(kBbHist + h->bb_pend + count) * 8 + 16, (size_t)(kBbHist + h->bb_pend) * 8, st))) return rc;
    if (count) DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_samples.as<float2>() + kBbHist + h->bb_pend, data, count * 8, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
    return rx_from_baseband(h, h->d_samples.as<float2>() + kBbHist, (size_t)h->bb_pend + count, true, end, ts_host, ts_dev, ts_capacity, ts_bytes, host ? 1 : 0);
  }
  // capture file
  if (first) {   // the resampler's output buffer starts with kBbHist zeros: nothing in front of the stream's first baseband sample
    if ((rc = h->d_samples.reserve((size_t)kBbHist * 8 + 16))) return rc;
    DVBT_CUDA_TRY(cudaMemsetAsync(h->d_samples.p, 0, (size_t)kBbHist * 8, st));
  }
  if (in_place) return rx_from_file(h, (const float2 *)data, count, gain, end, ts_host, ts_dev, ts_capacity, ts_bytes);
  if (first) {
    if ((rc = h->d_file.reserve((size_t)(kResHist + count) * 8 + 16))) return rc;
    DVBT_CUDA_TRY(cudaMemsetAsync(h->d_file.p, 0, (size_t)kResHist * 8, st));   // zero history (rational_resampler_base: d_history zeros)
    h->res_hist = kResHist;
    h->res_pend = 0;
  }
  if ((rc = reserve_keep(h->d_file, (size_t)(kResHist + h->res_pend + count) * 8 + 16, (size_t)(kResHist + h->res_pend) * 8, st))) return rc;
  if (count) DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_file.as<float2>() + kResHist + h->res_pend, data, count * 8, host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
  h->res_pend += (long long)count;
  return rx_from_file(h, nullptr, 0, gain, end, ts_host, ts_dev, ts_capacity, ts_bytes);
}

}  // namespace

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
      (rc = h->d_dstate.reserve(sizeof(DescrState))) || (rc = h->h_info.reserve(sizeof(DescrState)))) {
    dvbt_b200_rx_destroy(h);
    return rc;
  }
  for (auto &e : h->ev) cudaEventCreate(&e);
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  if (cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&h->ev_mid, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    set_error("rx_create: side stream: %s", cudaGetErrorString(cudaGetLastError()));
    dvbt_b200_rx_destroy(h);
    return DVBT_B200_ECUDA;
  }
  // energy_descramble PRBS (energy_descramble_impl.cc:46-67): 1 + x^14 + x^15, init 0xa9, 8 clocks per
  // byte, clocked but unused on every sync byte except the first of the group
  {
    uint8_t tab[1504];
    dvbt::energy_prbs_table(tab);
    if ((rc = h->d_prbs.reserve(1504))) { dvbt_b200_rx_destroy(h); return rc; }
    if (cudaMemcpy(h->d_prbs.p, tab, 1504, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("rx_create: PRBS table upload failed"); dvbt_b200_rx_destroy(h); return DVBT_B200_ECUDA; }
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
  rx_stream_reset(h);
  rx_info_reset(h);
  *out = h;
  return 0;
}

void dvbt_b200_rx_destroy(dvbt_b200_rx *h) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) return;
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->side) { cudaStreamSynchronize(h->side); cudaStreamDestroy(h->side); }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_mid) cudaEventDestroy(h->ev_mid);
  dvbt::DevBuf *bufs[] = {&h->d_state, &h->d_fo, &h->d_rot, &h->d_mod, &h->d_tps, &h->d_vote, &h->d_Y, &h->d_rsst, &h->d_ts, &h->d_prbs, &h->h_state,
                          &h->h_info, &h->d_sync, &h->h_sync, &h->d_plan, &h->d_dstate, &h->d_file, &h->d_file2, &h->d_samples, &h->d_samples2, &h->d_sym,
                          &h->d_sym2, &h->d_dm[0], &h->d_dm[1], &h->d_osym[0], &h->d_osym[1], &h->d_osrc[0], &h->d_osrc[1], &h->d_D, &h->d_D2,
                          &h->d_rs, &h->d_rs2, &h->d_vit_tap, &h->d_syncb};
  for (auto *b : bufs) b->release();
  for (auto &e : h->ev) if (e) cudaEventDestroy(e);
  h->tables.release();
  if (h->acq) dvbt_b200_acq_destroy(h->acq);
  if (h->vit) dvbt_b200_viterbi_destroy(h->vit);  // owns the stream
  delete h;
}

int dvbt_b200_rx_set_rs_compat(dvbt_b200_rx *h, int as_built) {
  if (!h) { set_error("rx_set_rs_compat: null handle"); return DVBT_B200_EINVAL; }
  h->rs_as_built = as_built ? 1 : 0;
  return 0;
}

// Soft-decision mode of the fused chain (beyond the reference; include/dvbt_b200.h): the demapper hands 4-bit values to the
// inner deinterleavers and the Viterbi decoder instead of hard bits.  Switching resets the stream state.
int dvbt_b200_rx_set_soft_decision(dvbt_b200_rx *h, int on, float scale) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) { set_error("rx_set_soft_decision: null handle"); return DVBT_B200_EINVAL; }
  if (on && !(scale >= 0.f && scale <= 64.f)) { set_error("rx_set_soft_decision: scale must be in [0, 64] (0 = default 4)"); return DVBT_B200_EINVAL; }
  if (int rc = dvbt_b200_viterbi_set_soft(h->vit, on ? 1 : 0)) return rc;
  DVBT_CUDA_TRY(dvbt::stream_wait(h->stream));
  h->soft = on != 0;
  h->demap.soft_scale = (on && scale > 0.f) ? scale : 4.0f;
  rx_stream_reset(h);
  h->fresh = true;
  return 0;
}

// ---- one-shot entry points: a whole capture = reset, one piece, end of stream ----
#define RX_ONE_SHOT(NAME, LEVEL, GAIN, HOST)                                                                              \
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);                                                                      \
  if (!h || (count && !data) || !ts) { set_error(NAME ": bad argument"); return DVBT_B200_EINVAL; }                       \
  if (!(HOST)) { if (int rc = dvbt::join_default_stream(h->stream)) return rc; }                                          \
  h->fresh = true;                                                                                                        \
  int rc__ = rx_push(h, LEVEL, data, count, GAIN, HOST, true, (HOST) ? ts : nullptr, (HOST) ? nullptr : ts, ts_capacity, ts_bytes); \
  h->fresh = true;   /* a one-shot run leaves no stream behind */                                                       \
  h->level = -1;                                                                                                          \
  return rc__

int dvbt_b200_rx_run_freq_host(dvbt_b200_rx *h, const void *data, size_t count, uint8_t *ts, size_t ts_capacity, size_t *ts_bytes) {
  RX_ONE_SHOT("rx_run_freq_host", kLevelFreq, 1.0f, true);
}
int dvbt_b200_rx_run_freq_dev(dvbt_b200_rx *h, const void *data, size_t count, uint8_t *ts, size_t ts_capacity, size_t *ts_bytes) {
  RX_ONE_SHOT("rx_run_freq_dev", kLevelFreq, 1.0f, false);
}
int dvbt_b200_rx_run_baseband_host(dvbt_b200_rx *h, const void *data, size_t count, uint8_t *ts, size_t ts_capacity, size_t *ts_bytes) {
  RX_ONE_SHOT("rx_run_baseband_host", kLevelBaseband, 1.0f, true);
}
int dvbt_b200_rx_run_baseband_dev(dvbt_b200_rx *h, const void *data, size_t count, uint8_t *ts, size_t ts_capacity, size_t *ts_bytes) {
  RX_ONE_SHOT("rx_run_baseband_dev", kLevelBaseband, 1.0f, false);
}
int dvbt_b200_rx_run_file_host(dvbt_b200_rx *h, const void *data, size_t count, float gain, uint8_t *ts, size_t ts_capacity, size_t *ts_bytes) {
  RX_ONE_SHOT("rx_run_file_host", kLevelFile, gain, true);
}
int dvbt_b200_rx_run_file_dev(dvbt_b200_rx *h, const void *data, size_t count, float gain, uint8_t *ts, size_t ts_capacity, size_t *ts_bytes) {
  RX_ONE_SHOT("rx_run_file_dev", kLevelFile, gain, false);
}
#undef RX_ONE_SHOT

// ---- streaming entry points ----
int dvbt_b200_rx_stream_reset(dvbt_b200_rx *h) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) { set_error("rx_stream_reset: null handle"); return DVBT_B200_EINVAL; }
  DVBT_CUDA_TRY(dvbt::stream_wait(h->stream));
  rx_stream_reset(h);
  rx_info_reset(h);
  return 0;
}

int dvbt_b200_rx_stream_push_host(dvbt_b200_rx *h, int level, const void *data, size_t count, float gain, int end_of_stream, uint8_t *ts,
                                  size_t ts_capacity, size_t *ts_bytes) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (count && !data) || !ts || level < kLevelFile || level > kLevelFreq) { set_error("rx_stream_push_host: bad argument"); return DVBT_B200_EINVAL; }
  int rc = rx_push(h, level, data, count, gain, true, end_of_stream != 0, ts, nullptr, ts_capacity, ts_bytes);
  if (end_of_stream || rc) { h->fresh = true; h->level = -1; }   // after an error the stream state is undefined: the next push starts a new stream
  return rc;
}

int dvbt_b200_rx_stream_push_dev(dvbt_b200_rx *h, int level, const void *d_data, size_t count, float gain, int end_of_stream, uint8_t *d_ts,
                                 size_t ts_capacity, size_t *ts_bytes) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (count && !d_data) || !d_ts || level < kLevelFile || level > kLevelFreq) { set_error("rx_stream_push_dev: bad argument"); return DVBT_B200_EINVAL; }
  if (int rc = dvbt::join_default_stream(h->stream)) return rc;
  int rc = rx_push(h, level, d_data, count, gain, false, end_of_stream != 0, nullptr, d_ts, ts_capacity, ts_bytes);
  if (end_of_stream || rc) { h->fresh = true; h->level = -1; }
  return rc;
}

int dvbt_b200_rx_last_info(const dvbt_b200_rx *h, dvbt_b200_rx_info *info) {
  if (!h || !info) { set_error("rx_last_info: null argument"); return DVBT_B200_EINVAL; }
  *info = h->info;
  return 0;
}

// stage taps of the last call, for stage-by-stage parity tests (a one-piece run; after a streaming call the taps
// cover what that call added)
int dvbt_b200_rx_read_stage(dvbt_b200_rx *h, int stage, void *host_out, size_t capacity_bytes, size_t *nbytes) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || !host_out || !nbytes) { set_error("rx_read_stage: null argument"); return DVBT_B200_EINVAL; }
  const dvbt::ModeDev &md = h->tables.dev;
  *nbytes = 0;
  const void *src = nullptr;
  size_t n = 0;
  dvbt::DevBuf tmp;
  const long long nout = h->tap_rows_out, first = h->tap_first_row;
  switch (stage) {
    case DVBT_RX_STAGE_CELLS:  // equalised cells of the output symbols (contiguous when no resync happened)
      if (nout > 0 && h->d_Y.p) { src = h->d_Y.as<float2>() + first * md.P; n = (size_t)nout * md.P * 8; }
      break;
    case DVBT_RX_STAGE_SOFT_CELLS:   // soft mode: one word per cell of the output symbols (value + 8 of bit e in nibble e)
      if (!h->soft) { set_error("rx_read_stage: the handle is not in soft-decision mode"); return DVBT_B200_EINVAL; }
      if (nout > 0) { src = h->d_dm[h->tap_set].as<uint32_t>() + (h->tap_carry + first) * md.P; n = (size_t)nout * md.P * 4; }
      break;
    case DVBT_RX_STAGE_SOFT_VALUES: {   // soft mode: one int8 per code bit in the order of the Viterbi block's input stream
      if (!h->soft) { set_error("rx_read_stage: the handle is not in soft-decision mode"); return DVBT_B200_EINVAL; }
      if (nout <= 0) break;
      n = (size_t)nout * md.P * h->m;
      int rc = tmp.reserve(n);
      if (rc) return rc;
      InnerMap im{h->d_dm[h->tap_set].as<uint8_t>(), h->d_osrc[h->tap_set].as<int>() + h->tap_carry, h->d_osym[h->tap_set].as<int>() + h->tap_carry,
                  md.H, md.Hinv, md.P, h->m, (int)nout, 0};
      rx_inner_soft_values_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(im, tmp.as<int8_t>(), (long long)n);
      dvbt::count_launch();
      DVBT_CUDA_TRY(cudaGetLastError());
      src = tmp.p;
      break;
    }
    case DVBT_RX_STAGE_DEMAP:
      if (h->soft) { set_error("rx_read_stage: soft-decision mode keeps no hard demapper output (read DVBT_RX_STAGE_SOFT_CELLS)"); return DVBT_B200_EINVAL; }
      if (nout > 0) { src = h->d_dm[h->tap_set].as<uint8_t>() + (h->tap_carry + first) * md.P; n = (size_t)nout * md.P; }
      break;
    case DVBT_RX_STAGE_BITDEINT: {
      if (h->soft) { set_error("rx_read_stage: soft-decision mode keeps no hard bits (read DVBT_RX_STAGE_SOFT_VALUES)"); return DVBT_B200_EINVAL; }
      if (nout <= 0) break;
      n = (size_t)nout * md.P;
      int rc = tmp.reserve(n);
      if (rc) return rc;
      InnerMap im{h->d_dm[h->tap_set].as<uint8_t>(), h->d_osrc[h->tap_set].as<int>() + h->tap_carry, h->d_osym[h->tap_set].as<int>() + h->tap_carry,
                  md.H, md.Hinv, md.P, h->m, (int)nout, 0};
      rx_inner_bytes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(im, tmp.as<uint8_t>(), (long long)n);
      dvbt::count_launch();
      DVBT_CUDA_TRY(cudaGetLastError());
      src = tmp.p;
      break;
    }
    case DVBT_RX_STAGE_VITERBI:
      if (h->tap_vit_copied) { src = h->d_vit_tap.p; n = (size_t)h->tap_vit_bytes; }
      else if (h->tap_vit_off >= 0) { src = h->d_D.as<uint8_t>() + h->tap_vit_off; n = (size_t)h->tap_vit_bytes; }
      break;
    case DVBT_RX_STAGE_RS: src = h->d_rs.as<uint8_t>() + h->tap_rs_first * 188; n = (size_t)h->tap_rs_packets * 188; break;
    case DVBT_RX_STAGE_RS_STATUS: src = h->d_rsst.as<int>() + h->tap_rs_first; n = (size_t)h->tap_rs_packets * 4; break;
    case DVBT_RX_STAGE_SYMBOL_INDEX: src = h->d_osym[h->tap_set].as<int>() + h->tap_carry; n = (size_t)(nout > 0 ? nout : 0) * 4; break;
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
