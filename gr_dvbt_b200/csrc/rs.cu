// K5 — RS(204,188) decoder, one warp per transport packet, GF(2^8) poly 0x11d.
//
// Replaces gr::dvbt::reed_solomon_dec (lib/reed_solomon_dec_impl.cc:77-116) and
// reed_solomon::rs_decode (lib/reed_solomon.cc:246-489).  The algebra is the reference's:
// syndromes S_i = r(a^i), i = 0..15 (:281-288; the 51 zero symbols of the shortened code
// contribute nothing, so only the 204 received bytes are read), the reference's
// Berlekamp-Massey iteration (:315-354), Chien search in increasing position order
// (:376-403), error evaluator and Forney values (:419-486) including the "uncorrectable:
// leave the packet alone" and "null denominator: keep what was already corrected" exits.
//
// Mapping: a block takes 128 packets.  The clean-packet test is a division by the generator polynomial,
// one packet per thread (rs_decode_kernel); only packets with a non-zero remainder go through the
// syndrome / locator / Chien / Forney path, one warp per packet (rs_warp_decode): the 32 lanes split the 204
// bytes (7 per lane), each lane accumulates its 16 partial syndromes from log/exp tables in shared memory, a
// butterfly of warp shuffles XOR-reduces them (16 syndromes packed in 4 registers), the locator iteration
// runs on lane 0, the Chien search on all lanes (8 positions per lane, ballots keep the order) and Forney on
// lane 0.
#include "common.cuh"

#include <new>

namespace {

constexpr int kN = 255, kT = 8, kS = 51, kPktIn = 204, kPktOut = 188;

__constant__ uint8_t c_exp2[512];
__constant__ uint8_t c_log[256];
bool g_tables_ready[64] = {false};

struct GfTables {
  const uint8_t *exp2;  // exp2[i] = a^(i mod 255), 0 <= i < 512
  const uint8_t *log;   // log[0] = 255
  __device__ __forceinline__ int mul(int a, int b) const { return (a && b) ? exp2[log[a] + log[b]] : 0; }
  __device__ __forceinline__ int div(int a, int b) const { return (a && b) ? exp2[255 + log[a] - log[b]] : 0; }
  // a * alpha^p for any p >= 0 (gf_pow, reed_solomon.cc:136-143)
  __device__ __forceinline__ int mulpow(int a, int p) const { return a ? exp2[(log[a] + p) % 255] : 0; }
};

// Berlekamp-Massey exactly as reed_solomon.cc:307-364 (no erasures), one polynomial coefficient per lane:
// lane i holds sigma[i] and b[i] (i <= 16).  Every step of the reference's loop is element-wise in i except the
// discrepancy (a warp XOR-reduce) and the shift b[i] = b[i-1] (a shuffle), so the 16 iterations are the same
// arithmetic in the same order.  Writes sigma[0..16] to shared memory; returns deg(sigma) in every lane.
__device__ int rs_locator_warp(const GfTables &gf, const uint8_t *syn, uint8_t *sigma_out, int lane) {
  int sg = lane == 0 ? 1 : 0, bb = sg;
  int el = 0;
  for (int r = 1; r <= 2 * kT; r++) {
    int term = (lane < r && lane <= 2 * kT) ? gf.mul(sg, syn[r - lane - 1]) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) term ^= __shfl_xor_sync(0xffffffffu, term, o);
    const int discr = term;
    int b_up = __shfl_up_sync(0xffffffffu, bb, 1);   // b[i-1]
    if (lane == 0) b_up = 0;
    if (discr == 0) {
      bb = b_up;
    } else {
      int T = lane == 0 ? sg : (sg ^ gf.mul(discr, b_up));   // T[0] = sigma[0]; T[i+1] = sigma[i+1] ^ discr * b[i]
      if (2 * el <= r - 1) {
        el = r - el;
        bb = gf.div(sg, discr);
      } else {
        bb = b_up;
      }
      sg = T;
    }
    if (lane > 2 * kT) { sg = 0; bb = 0; }
  }
  if (lane <= 2 * kT) sigma_out[lane] = (uint8_t)sg;
  unsigned nz = __ballot_sync(0xffffffffu, lane <= 2 * kT && sg != 0);
  return nz ? 31 - __clz(nz) : 0;
}

// Forney, reed_solomon.cc:417-486, one root per lane.  root[]/loc[] hold no_roots == deg_sigma entries.
// pkt = the 204 received bytes (positions 51..254 of the code word).  The reference walks the roots from the
// last to the first and stops at the first null denominator, keeping the corrections made so far (:470-479):
// here every lane evaluates its root, the highest root index with a null denominator is found by ballot, and
// only the roots above it are applied.
__device__ int rs_forney_warp(const GfTables &gf, const uint8_t *syn, const uint8_t *sigma, int deg_sigma, const uint8_t *root,
                              const uint8_t *loc, int no_roots, uint8_t *pkt, uint8_t *omega, int lane) {
  // omega[i] = sum_j syn[i-j] sigma[j], i < 2t: lane i
  int om = 0;
  if (lane < 2 * kT) {
    int j = deg_sigma < lane ? deg_sigma : lane;
    for (; j >= 0; j--) om ^= gf.mul(syn[lane - j], sigma[j]);
    omega[lane] = (uint8_t)om;
  }
  unsigned onz = __ballot_sync(0xffffffffu, lane < 2 * kT && om != 0);
  const int deg_omega = onz ? 31 - __clz(onz) : 0;
  __syncwarp();
  int den = 1, err = 0, pos = 0;
  if (lane < no_roots) {
    int rt = root[lane];
    int num1 = 0;
    for (int i = deg_omega; i >= 0; i--) num1 ^= gf.mulpow(omega[i], i * rt);
    int num2 = gf.exp2[(kN - rt) % kN];
    den = 0;
    int deg_max = deg_sigma < 2 * kT - 1 ? deg_sigma : 2 * kT - 1;
    for (int i = 1; i <= deg_max; i += 2)
      if (sigma[i]) den ^= gf.exp2[(gf.log[sigma[i]] + (i - 1) * rt) % kN];
    err = den ? gf.div(gf.mul(num1, num2), den) : 0;
    pos = loc[lane];
  }
  unsigned zero_den = __ballot_sync(0xffffffffu, lane < no_roots && den == 0);
  const int first_ok = zero_den ? 32 - __clz(zero_den) : 0;   // roots first_ok .. no_roots-1 are applied
  if (lane < no_roots && lane >= first_ok && pos >= kS) pkt[pos - kS] ^= (uint8_t)err;  // positions < 51 are the discarded zero prefix
  __syncwarp();
  return zero_den ? -1 : no_roots;
}

// Syndromes, locator, Chien and Forney for one packet held in shared memory (204 bytes at `pkt`), by one
// warp; returns the reference's per-packet result (0 clean, > 0 corrected symbols, -1 uncorrectable) in
// every lane.  `work`: 96 bytes of per-warp scratch.
__device__ int rs_warp_decode(const GfTables &gf, uint8_t *pkt, uint8_t *work, int lane, int as_built) {
  uint8_t *syn = work, *sigma = syn + 16, *root = sigma + 17, *loc = root + 17, *misc = loc + 17;
  // ---- syndromes: byte j carries x^(203-j); lane takes j = lane, lane+32, ...
  uint32_t S0 = 0, S1 = 0, S2 = 0, S3 = 0;
#pragma unroll
  for (int t = 0; t < 7; t++) {
    int j = lane + 32 * t;
    if (j < kPktIn) {
      int v = pkt[j];
      if (v) {
        int ex = 203 - j;
        int e = gf.log[v];
        uint32_t acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 16; i++) {
          acc[i >> 2] |= (uint32_t)gf.exp2[e] << (8 * (i & 3));
          e += ex;
          if (e >= 255) e -= 255;
        }
        S0 ^= acc[0]; S1 ^= acc[1]; S2 ^= acc[2]; S3 ^= acc[3];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    S0 ^= __shfl_xor_sync(0xffffffffu, S0, o);
    S1 ^= __shfl_xor_sync(0xffffffffu, S1, o);
    S2 ^= __shfl_xor_sync(0xffffffffu, S2, o);
    S3 ^= __shfl_xor_sync(0xffffffffu, S3, o);
  }
  int st = 0;
  if ((S0 | S1 | S2 | S3) != 0u) {  // warp-uniform
    if (lane == 0) {
      uint32_t S[4] = {S0, S1, S2, S3};
      for (int i = 0; i < 16; i++) syn[i] = (uint8_t)(S[i >> 2] >> (8 * (i & 3)));
    }
    __syncwarp();
    const int deg = rs_locator_warp(gf, syn, sigma, lane);
    __syncwarp();
    // ---- Chien: q(i) = 1 + sum_j sigma[j] a^(j*i), i = 1..255 in increasing order
    int nroots = 0;
    for (int t = 0; t < 8; t++) {
      int i = 32 * t + lane + 1;
      int q = 1;
      if (i <= kN) {
        for (int j = deg; j > 0; j--) q ^= gf.mulpow(sigma[j], j * i);
      }
      unsigned hit = __ballot_sync(0xffffffffu, i <= kN && q == 0);
      if (q == 0 && i <= kN) {
        int idx = nroots + __popc(hit & ((1u << lane) - 1u));
        if (idx < 17) { root[idx] = (uint8_t)i; loc[idx] = (uint8_t)(i - 1); }
      }
      nroots += __popc(hit);
    }
    __syncwarp();
    if (nroots != deg) {
      st = -1;  // uncorrectable: data untouched (:405-415)
    } else {
      // the reference's out-of-bounds omega[2t] = 0 lands on loc[0] with gcc 13.3
      if (as_built && lane == 0) loc[0] = 0;
      __syncwarp();
      st = rs_forney_warp(gf, syn, sigma, deg, root, loc, nroots, pkt, misc, lane);
    }
  }
  return st;
}

// Row f of the division table: the 16 bytes f * g_k, k = 0..15, of the generator polynomial
// g(x) = prod_{i=0..15} (x + a^i) (reed_solomon.cc:95-133 builds the same g), as 4 little-endian words.
__device__ uint32_t d_lfsr[256 * 4];

constexpr int kTilePk = 256;            // packets per block
constexpr int kHalo = 204 * 11;         // bytes of stream history the deepest deinterleaver branch reaches back

// One block per 128 packets.
//   1. The contiguous input range of the tile is staged in shared memory with coalesced loads.  GATHER: `in` is
//      the Viterbi output stream (in_bytes long, index 0 = the byte carrying the superframe_start tag) and the
//      12-branch Forney deinterleaver of convolutional_deinterleaver_impl.cc:93-150 (branch b = t % 12 is delayed
//      by 17*(11-b) cells, i.e. 204*(11-b) stream positions; the FIFOs start zeroed) is the index map
//      packet q, byte j  ->  staged[q*204 + j + 204*(j % 12)]  applied when a byte is read.
//   2. One THREAD per packet divides the received word by g(x) (LFSR, one 16-byte table row per byte): the
//      remainder is zero exactly when all 16 syndromes r(a^i) are zero.  This is the whole cost of a clean
//      packet - 204 table rows instead of 204 x 16 log/exp lookups.
//   3. All packets are copied out (bytes 0..187) with coalesced stores.
//   4. Packets with a non-zero remainder are decoded by a warp each exactly as before (rs_warp_decode) and
//      written over their copy.
template <bool GATHER>
__global__ void __launch_bounds__(kTilePk) rs_decode_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out,
                                                            int *__restrict__ status, long long npackets, int as_built,
                                                            long long in_bytes, long long in_lo) {
  extern __shared__ __align__(16) uint8_t s_dyn[];   // division table (8 copies), then the staged input [kTilePk*204 (+ kHalo)]
  // Copy c of row f sits at 16-byte slot 8 f + c: lane l reads copy l & 7, so the eight lanes of a quarter warp -
  // what a 128-bit shared-memory access serves per pass - always hit eight different bank groups, whatever their
  // rows (one shared copy cost ~3 passes per quarter warp and left the kernel LSU bound).
  uint4 *s_lfsr = reinterpret_cast<uint4 *>(s_dyn);
  uint8_t *s_raw = s_dyn + 256 * 8 * 16;
  __shared__ __align__(16) uint8_t s_exp2[512];
  __shared__ __align__(16) uint8_t s_log[256];
  __shared__ __align__(16) uint32_t s_pkt[kTilePk / 32][52];
  __shared__ uint8_t s_work[kTilePk / 32][96];
  __shared__ int s_dirty[kTilePk];
  __shared__ int s_ndirty;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < 512; i += blockDim.x) s_exp2[i] = c_exp2[i];
  for (int i = t; i < 256; i += blockDim.x) s_log[i] = c_log[i];
  for (int i = t; i < 256 * 8; i += blockDim.x) s_lfsr[i] = reinterpret_cast<const uint4 *>(d_lfsr)[i >> 3];
  if (t == 0) s_ndirty = 0;
  const long long p0 = (long long)blockIdx.x * kTilePk;
  const int np = (int)((npackets - p0) < kTilePk ? (npackets - p0) : kTilePk);
  const int halo = GATHER ? kHalo : 0;
  const long long total = GATHER ? in_bytes : npackets * (long long)kPktIn;
  {
    const long long src0 = p0 * kPktIn - halo;             // multiple of 4
    const int nwords = (np * kPktIn + halo) / 4;
    uint32_t *raw32 = reinterpret_cast<uint32_t *>(s_raw);
    // eight loads in flight per thread (a load-then-store loop waits one full memory latency per word)
    for (int i0 = t; i0 < nwords; i0 += 8 * kTilePk) {
      uint32_t v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int i = i0 + u * kTilePk;
        long long pos = src0 + 4LL * i;
        v[u] = 0;
        if (i < nwords) {
          // in_lo <= 0: stream positions [in_lo, 0) before `in` hold real history (a continuing stream: the delay
          // lines keep their contents, convolutional_deinterleaver_impl.cc:109-120); positions below read as the
          // zeros the FIFOs start with
          if (pos >= in_lo && pos + 4 <= total) {
            v[u] = __ldg(reinterpret_cast<const uint32_t *>(in + pos));
          } else {
            for (int b = 0; b < 4; b++)
              if (pos + b >= in_lo && pos + b < total) v[u] |= (uint32_t)in[pos + b] << (8 * b);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int i = i0 + u * kTilePk;
        if (i < nwords) raw32[i] = v[u];
      }
    }
  }
  __syncthreads();
  // ---- division by g(x), one packet per thread
  if (t < np) {
    const uint8_t *src = s_raw + t * kPktIn;
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
    int jm = 0;  // j % 12
    for (int j = 0; j < kPktIn; j++) {
      uint32_t d = src[j + (GATHER ? kPktIn * jm : 0)];
      if (++jm == 12) jm = 0;
      uint32_t fb = d ^ (r3 >> 24);
      r3 = __funnelshift_l(r2, r3, 8);
      r2 = __funnelshift_l(r1, r2, 8);
      r1 = __funnelshift_l(r0, r1, 8);
      r0 <<= 8;
      uint4 row = s_lfsr[fb * 8 + (lane & 7)];
      r0 ^= row.x; r1 ^= row.y; r2 ^= row.z; r3 ^= row.w;
    }
    if ((r0 | r1 | r2 | r3) != 0u) s_dirty[atomicAdd(&s_ndirty, 1)] = t;
    if (status) status[p0 + t] = 0;
  }
  __syncthreads();
  // ---- copy out the 188 data bytes of every packet (47 words each)
  {
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + p0 * kPktOut);
    const int nwords = np * 47;
    for (int k = t; k < nwords; k += blockDim.x) {
      int q = k / 47, j = 4 * (k - q * 47);
      const uint8_t *src = s_raw + q * kPktIn;
      uint32_t v = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        int jj = j + b;
        v |= (uint32_t)src[jj + (GATHER ? kPktIn * (jj % 12) : 0)] << (8 * b);
      }
      dst[k] = v;
    }
  }
  __syncthreads();
  // ---- packets with errors: one warp each
  GfTables gf{s_exp2, s_log};
  const int ndirty = s_ndirty;
  uint8_t *pkt = reinterpret_cast<uint8_t *>(s_pkt[warp]);
  for (int d = warp; d < ndirty; d += kTilePk / 32) {
    const int q = s_dirty[d];
    const uint8_t *src = s_raw + q * kPktIn;
#pragma unroll
    for (int u = 0; u < 7; u++) {
      int j = lane + 32 * u;
      if (j < kPktIn) pkt[j] = src[j + (GATHER ? kPktIn * (j % 12) : 0)];
    }
    __syncwarp();
    int st = rs_warp_decode(gf, pkt, s_work[warp], lane, as_built);
    __syncwarp();
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + (p0 + q) * kPktOut);
    dst[lane] = s_pkt[warp][lane];
    if (lane + 32 < 47) dst[lane + 32] = s_pkt[warp][lane + 32];
    if (status && lane == 0) status[p0 + q] = st;
    __syncwarp();  // pkt is rewritten by the next iteration
  }
}


// ---- transmit side (a synthetic-input generator on the device) ------------------------------------
// energy_dispersal (energy_dispersal_impl.cc:93-140: PRBS 1 + x^14 + x^15 restarted every 8 packets, first sync byte
// inverted) + reed_solomon_enc (reed_solomon.cc:216-244: systematic, parity = remainder of data x^16 modulo g(x) - the
// same division the decoder's clean-packet test makes, with the same table): one thread per packet.
__global__ void __launch_bounds__(128) tx_outer_kernel(const uint8_t *__restrict__ ts, long long npk, const uint8_t *__restrict__ prbs,
                                                       uint8_t *__restrict__ ed, uint8_t *__restrict__ rs) {
  __shared__ uint4 s_lfsr[256];
  __shared__ uint8_t s_prbs[1504];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lfsr[i] = reinterpret_cast<const uint4 *>(d_lfsr)[i];
  for (int i = threadIdx.x; i < 1504; i += blockDim.x) s_prbs[i] = prbs[i];
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npk) return;
  const int j = (int)(p & 7);
  const uint8_t *in = ts + p * 188;
  uint8_t *e = ed + p * 188, *o = rs + p * 204;
  uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
  for (int k = 0; k < 188; k++) {
    const uint8_t d = k == 0 ? (j == 0 ? 0xB8 : 0x47) : (uint8_t)(in[k] ^ s_prbs[j * 188 + k]);
    e[k] = d;
    o[k] = d;
    const uint32_t fb = d ^ (r3 >> 24);
    r3 = __funnelshift_l(r2, r3, 8);
    r2 = __funnelshift_l(r1, r2, 8);
    r1 = __funnelshift_l(r0, r1, 8);
    r0 <<= 8;
    const uint4 row = s_lfsr[fb];
    r0 ^= row.x; r1 ^= row.y; r2 ^= row.z; r3 ^= row.w;
  }
  const uint32_t r[4] = {r3, r2, r1, r0};          // parity[0] = coefficient of x^15 ... parity[15] = x^0
#pragma unroll
  for (int q = 0; q < 16; q++) o[188 + q] = (uint8_t)(r[q >> 2] >> (24 - 8 * (q & 3)));
}

// convolutional_interleaver(136, 12, 17) (convolutional_interleaver_impl.cc:66-88): branch j = t % 12 delays by 17 j cells
// of its own = 204 j stream positions, the FIFOs start zeroed
__global__ void tx_outer_interleave_kernel(const uint8_t *__restrict__ rs, long long nbytes, uint8_t *__restrict__ ci) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nbytes) return;
  const long long src = t - 204LL * (t % 12);
  ci[t] = src >= 0 ? rs[src] : (uint8_t)0;
}

}  // namespace

struct dvbt_b200_rsdec {
  int device = dvbt::current_device();
  dvbt_b200_rsdec_params par;
  int as_built = 0;
  cudaStream_t stream = nullptr;
  dvbt::DevBuf d_in, d_out;
  dvbt::Staging stg;
  int sm_count = 148;
};

namespace dvbt {
int rs_upload_tables() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && g_tables_ready[dev]) return 0;
  uint8_t e2[512], lg[256];
  int reg = 1;  // reed_solomon.cc:48-89 with p = 2, m = 8, gfpoly = 0x11d
  lg[0] = 255;
  for (int i = 0; i < 255; i++) {
    e2[i] = (uint8_t)reg;
    lg[reg] = (uint8_t)i;
    reg <<= 1;
    if (reg & 0x100) reg ^= 0x11d;
    reg &= 0xff;
  }
  for (int i = 255; i < 512; i++) e2[i] = e2[i - 255];
  DVBT_CUDA_TRY(cudaMemcpyToSymbol(c_exp2, e2, 512));
  DVBT_CUDA_TRY(cudaMemcpyToSymbol(c_log, lg, 256));
  {
    auto mul = [&](int a, int b) { return (a && b) ? (int)e2[lg[a] + lg[b]] : 0; };
    uint8_t g[17] = {1};          // g(x) = prod (x + a^i), i = 0..15; g[k] = coefficient of x^k
    int deg = 0;
    for (int i = 0; i < 16; i++) {
      int root = e2[i];
      for (int k = deg + 1; k > 0; k--) g[k] = (uint8_t)(g[k - 1] ^ mul(g[k], root));
      g[0] = (uint8_t)mul(g[0], root);
      deg++;
    }
    static uint8_t rows[256 * 16];
    for (int f = 0; f < 256; f++)
      for (int k = 0; k < 16; k++) rows[f * 16 + k] = (uint8_t)mul(f, g[k]);
    DVBT_CUDA_TRY(cudaMemcpyToSymbol(d_lfsr, rows, sizeof rows));
  }
  if (dev < 64) g_tables_ready[dev] = true;
  return 0;
}

int rs_launch(const uint8_t *d_in, uint8_t *d_out, int *d_status, long long npackets, int as_built, int sm_count,
              cudaStream_t st, long long gather_stream_bytes, long long history_bytes) {
  if (npackets <= 0) return 0;
  int rc = rs_upload_tables();
  if (rc) return rc;
  unsigned grid = (unsigned)((npackets + kTilePk - 1) / kTilePk);
  (void)sm_count;
  const size_t table = 256 * 8 * 16;
  if (gather_stream_bytes >= 0) {
    const size_t smem = (size_t)kTilePk * kPktIn + kHalo + table;
    DVBT_CUDA_TRY(cudaFuncSetAttribute(rs_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rs_decode_kernel<true><<<grid, kTilePk, smem, st>>>(d_in, d_out, d_status, npackets, as_built, gather_stream_bytes, -history_bytes);
  } else {
    const size_t smem = (size_t)kTilePk * kPktIn + table;
    DVBT_CUDA_TRY(cudaFuncSetAttribute(rs_decode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rs_decode_kernel<false><<<grid, kTilePk, smem, st>>>(d_in, d_out, d_status, npackets, as_built, 0, 0);
  }
  count_launch();
  DVBT_CUDA_TRY(cudaGetLastError());
  return 0;
}

// TS packets -> energy dispersal -> RS(204,188) -> Forney interleaver (all on the device); ed / rs / ci are npk x 188 / 204 / 204 bytes
int tx_outer_launch(const uint8_t *d_ts, long long npk, const uint8_t *d_prbs, uint8_t *d_ed, uint8_t *d_rs, uint8_t *d_ci, cudaStream_t st) {
  if (npk <= 0) return 0;
  int rc = rs_upload_tables();
  if (rc) return rc;
  tx_outer_kernel<<<(unsigned)((npk + 127) / 128), 128, 0, st>>>(d_ts, npk, d_prbs, d_ed, d_rs);
  const long long nb = npk * 204;
  tx_outer_interleave_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(d_rs, nb, d_ci);
  count_launch(2);
  DVBT_CUDA_TRY(cudaGetLastError());
  return 0;
}
}  // namespace dvbt

extern "C" {

int dvbt_b200_rsdec_create(const dvbt_b200_rsdec_params *p, dvbt_b200_rsdec **out) {
  if (!p || !out) { dvbt::set_error("rsdec_create: null argument"); return DVBT_B200_EINVAL; }
  *out = nullptr;
  if (p->p != 2 || p->m != 8 || p->gfpoly != 0x11d || p->n != 255 || p->k != 239 || p->t != 8 || p->s != 51 || p->blocks <= 0) {
    dvbt::set_error("rsdec_create: only the DVB-T outer code is built: (p,m,gfpoly,n,k,t,s) = (2,8,0x11d,255,239,8,51)");
    return DVBT_B200_EINVAL;
  }
  int rc = dvbt::ensure_device();
  if (rc) return rc;
  dvbt_b200_rsdec *h = new (std::nothrow) dvbt_b200_rsdec();
  if (!h) { dvbt::set_error("rsdec_create: out of memory"); return DVBT_B200_ENOMEM; }
  h->par = *p;
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    dvbt::set_error("rsdec_create: cannot create stream");
    delete h;
    return DVBT_B200_ECUDA;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
  *out = h;
  return 0;
}

void dvbt_b200_rsdec_destroy(dvbt_b200_rsdec *h) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) return;
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  h->d_in.release();
  h->d_out.release();
  h->stg.release();
  delete h;
}

int dvbt_b200_rsdec_set_compat(dvbt_b200_rsdec *h, int as_built) {
  if (!h) { dvbt::set_error("rsdec_set_compat: null handle"); return DVBT_B200_EINVAL; }
  h->as_built = as_built ? 1 : 0;
  return 0;
}

int dvbt_b200_rsdec_decode_dev(dvbt_b200_rsdec *h, const uint8_t *d_in, size_t npackets, uint8_t *d_out, int *d_status) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || (npackets && (!d_in || !d_out))) { dvbt::set_error("rsdec_decode_dev: bad argument"); return DVBT_B200_EINVAL; }
  int rc = dvbt::join_default_stream(h->stream);
  if (rc) return rc;
  rc = dvbt::rs_launch(d_in, d_out, d_status, (long long)npackets, h->as_built, h->sm_count, h->stream, -1, 0);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaStreamSynchronize(h->stream));
  return 0;
}

int dvbt_b200_rsdec_work(dvbt_b200_rsdec *h, const uint8_t *in, size_t n_in_items, uint8_t *out, size_t noutput_items,
                         size_t *consumed, size_t *produced) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || !consumed || !produced) { dvbt::set_error("rsdec_work: null argument"); return DVBT_B200_EINVAL; }
  *consumed = *produced = 0;
  if (n_in_items < noutput_items) {  // forecast: 1:1 (reed_solomon_dec_impl.cc:71-75)
    dvbt::set_error("rsdec_work: %zu input items for %zu output items", n_in_items, noutput_items);
    return DVBT_B200_EINVAL;
  }
  if (noutput_items == 0) return 0;
  if (!in || !out) { dvbt::set_error("rsdec_work: null buffer"); return DVBT_B200_EINVAL; }
  size_t npk = noutput_items * (size_t)h->par.blocks;
  int rc;
  if ((rc = h->d_in.reserve(npk * kPktIn))) return rc;
  if ((rc = h->d_out.reserve(npk * kPktOut))) return rc;
  if ((rc = h->stg.h2d(h->d_in.p, in, npk * kPktIn, h->stream))) return rc;
  rc = dvbt::rs_launch(h->d_in.as<uint8_t>(), h->d_out.as<uint8_t>(), nullptr, (long long)npk, h->as_built, h->sm_count, h->stream, -1, 0);
  if (rc) return rc;
  if ((rc = h->stg.d2h(out, h->d_out.p, npk * kPktOut, h->stream))) return rc;
  *consumed = noutput_items;
  *produced = noutput_items;
  return 0;
}

}  // extern "C"
