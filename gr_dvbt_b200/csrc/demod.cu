// K4 — demod_reference_signals on sm_100a: post-FFT frequency correction, scattered-pilot
// channel estimate, equalisation (+ fused demap), TPS decoding and frame synchronisation.
//
// Replaces gr::dvbt::demod_reference_signals (lib/demod_reference_signals_impl.cc:96-150) and
// pilot_gen::parse_input with its callees (lib/reference_signals_impl.cc:1188-1248):
//   process_cpilot_data   :714-744   integer CFO in [-8,8) from continual-pilot differences
//   compute_oneshot_csft  :746-790   fractional CFO from this and the next symbol
//   frequency_correction  :792-819   one rotor per symbol + bin shift
//   process_spilot_data   :535-689   scattered-pilot phase (symbol index mod 4), pilot gains,
//                                    linear interpolation with the reference's constant /11 step
//   process_tps_data      :918-1032  DBPSK majority vote, 68-bit FIFO, sync word + BCH(67,53)
//   process_payload_data  :1064-1124 payload carriers x gain
//
// The reference is one sequential loop per symbol.  Here every float operation that feeds a
// decision or an output keeps the reference's operand order and rounding (explicit
// __f*_rn / __d*_rn, no FMA contraction; complex division as libgcc's __divsc3 does it for
// float operands: straight formula in double), but the work is split:
//   demod_stage1_kernel : one warp per symbol   - CFO sums (16 lanes), one-shot sums (2 lanes), rotor,
//                                                scattered-pilot phase (4 lanes); in the fused chain also the symbol's
//                                                17 / 68 TPS carriers, equalised exactly as the symbol kernel would
//   demod_symbol_kernel : one block per symbol  - the symbol staged once by a bulk async copy (TMA engine); pilot
//                                                gains and slopes; (TPS carriers,) payload cells (+ demap), four cells
//                                                per thread.  <true>: stage 1 fused in as well (opt-in, slower);
//                                                <.., SOFT>: soft decisions instead of the hard demap (demod.cuh)
//   demod_vote_kernel   : one thread per symbol - TPS majority vote against the previous symbol
//   demod_scan_kernel   : one block             - the genuinely sequential part (symbol/frame
//                                                index, TPS FIFO, BCH, superframe gating, sync_start re-arming)
// In the fused chain vote + scan run on a high-priority side stream BESIDE the symbol kernel (demod_run).
// HBM traffic per symbol: 8 (K + 17) B read once (+ 128 B around each continual pilot of the next symbol, L2 hits) and
// P (demapped) written, + 8P when the caller wants the equalised cells.
#include "demod.cuh"
#include "bulk_copy.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include <vector>

namespace dvbt {

// ---------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------
static const short kCpilot2k[45] = {0,   48,  54,  87,  141, 156, 192, 201, 255,  279,  282,  333,  432,  450,  483,
                                    525, 531, 618, 636, 714, 759, 765, 780, 804,  873,  888,  918,  939,  942,  969,
                                    984, 1050, 1101, 1107, 1110, 1137, 1140, 1146, 1206, 1269, 1323, 1377, 1491, 1683, 1704};
static const short kTps2k[17] = {34, 50, 209, 346, 413, 569, 595, 688, 790, 901, 1073, 1219, 1262, 1286, 1469, 1594, 1687};

int ModeTables::init(int tm, int gi) {
  if (tm != DVBT_T2K && tm != DVBT_T8K) { set_error("mode tables: bad transmission mode %d", tm); return DVBT_B200_EINVAL; }
  if (gi < DVBT_G1_32 || gi > DVBT_G1_4) { set_error("mode tables: bad guard interval %d (dvbt_config.cc:194-208 knows 1/32 .. 1/4)", gi); return DVBT_B200_EINVAL; }
  ModeDev &d = dev;
  d.N = tm == DVBT_T2K ? 2048 : 8192;
  d.P = tm == DVBT_T2K ? 1512 : 6048;
  int Kmax = tm == DVBT_T2K ? 1704 : 6816;
  d.K = Kmax + 1;
  d.zl = (int)ceil((d.N - d.K) / 2.0);  // dvbt_config.cc:124
  d.cp = gi == DVBT_G1_16 ? d.N / 16 : gi == DVBT_G1_8 ? d.N / 8 : gi == DVBT_G1_4 ? d.N / 4 : d.N / 32;
  d.ncp = tm == DVBT_T2K ? 45 : 177;
  d.ntps = tm == DVBT_T2K ? 17 : 68;
  // reference_signals_impl.cc:753 — float(cp)/float(N) in float, the rest in double, stored in a float
  d.carrier_coeff = (float)(1.0 / (2 * M_PI * (1 + float(d.cp) / float(d.N)) * 2));
  const int K = d.K, P = d.P;
  // the 8k tables are the 2k tables repeated with +1704*r (reference_signals_impl.cc:77-117)
  std::vector<short> cpl, tpl;
  int reps = tm == DVBT_T2K ? 1 : 4;
  for (int r = 0; r < reps; r++)
    for (int i = 0; i < 45; i++) {
      short v = (short)(kCpilot2k[i] + 1704 * r);
      if (cpl.empty() || cpl.back() != v) cpl.push_back(v);
    }
  for (int r = 0; r < reps; r++)
    for (int i = 0; i < 17; i++) tpl.push_back((short)(kTps2k[i] + 1704 * r));
  if ((int)cpl.size() != d.ncp || (int)tpl.size() != d.ntps) { set_error("mode tables: pilot table size mismatch"); return DVBT_B200_EINVAL; }
  // PRBS x^11 + x^2 + 1 (:333-345)
  std::vector<float> pval(K);
  {
    unsigned reg = (1u << 11) - 1;
    for (int k = 0; k < K; k++) {
      int w = reg & 1;
      int nb = ((reg >> 2) ^ reg) & 1;
      reg = (reg >> 1) | (nb << 10);
      pval[k] = (float)(4 * 2 * (0.5 - w) / 3);  // :474-479 / :700-705
    }
  }
  std::vector<float> known(d.ncp - 1);
  for (int i = 0; i < d.ncp - 1; i++) {
    float dr = pval[cpl[i + 1]] - pval[cpl[i]];
    known[i] = dr * dr + 0.0f * 0.0f;  // norm() of a real difference (:224-228)
  }
  std::vector<unsigned char> kind(4 * K, 0);
  std::vector<short> prevp(4 * K), nextp(4 * K), payload(4 * P);
  for (int r = 0; r < 4; r++) {
    unsigned char *kd = &kind[r * K];
    for (int k = 3 * r; k < K; k += 12) kd[k] |= 1;
    for (short c : cpl) kd[c] |= 1;
    for (short t : tpl) kd[t] |= 2;
    int last = 0;
    for (int k = 0; k < K; k++) {
      if (kd[k] & 1) last = k;
      prevp[r * K + k] = (short)last;
    }
    int nxt = K - 1;
    for (int k = K - 1; k >= 0; k--) {
      nextp[r * K + k] = (short)nxt;
      if (kd[k] & 1) nxt = k;
    }
    int n = 0;
    for (int k = 0; k < K; k++)
      if (kd[k] == 0) {
        if (n < P) payload[r * P + n] = (short)k;
        n++;
      }
    if (n != P) { set_error("mode tables: %d payload carriers for scattered phase %d, expected %d", n, r, P); return DVBT_B200_EINVAL; }
  }
  std::vector<int> pay32(4 * P);
  for (int r = 0; r < 4; r++)
    for (int i = 0; i < P; i++) {
      int k = payload[r * P + i];
      pay32[r * P + i] = k | ((k - prevp[r * K + k]) << 16);
    }
  // dense list of the channel-estimation carriers of every scattered phase (the kernels loop over these
  // instead of testing kind[] on all K carriers: the gain computation is a double-precision complex division,
  // and with one pilot in twelve it would otherwise run with three lanes of a warp active)
  d.pil_stride = (K + 11) / 12 + d.ncp + 8;
  std::vector<short> pil((size_t)4 * d.pil_stride, 0);
  for (int r = 0; r < 4; r++) {
    int n = 0;
    for (int k = 0; k < K; k++)
      if (kind[r * K + k] & 1) pil[(size_t)r * d.pil_stride + n++] = (short)k;
    if (n > d.pil_stride) { set_error("mode tables: %d pilots for phase %d", n, r); return DVBT_B200_EINVAL; }
    d.npil[r] = n;
  }
  // The fused symbol kernel keeps pilot gains and interval slopes by pilot ORDINAL (their number, not the K carriers, sets
  // the shared-memory footprint): payload entry = k | (k - k0) << 13 | ordinal(k0) << 17 with k0 the channel-estimation
  // carrier at or below k; the same for the TPS carriers; and for every pilot the ordinal of the next one up.
  std::vector<int> pay3(4 * P), tps3(4 * d.ntps);
  std::vector<short> ordk(4 * K, 0);
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < d.npil[r]; i++) ordk[r * K + pil[(size_t)r * d.pil_stride + i]] = (short)i;
    auto pack = [&](int k) { int k0 = prevp[r * K + k]; return k | ((k - k0) << 13) | ((int)ordk[r * K + k0] << 17); };
    for (int i = 0; i < P; i++) pay3[r * P + i] = pack(payload[r * P + i]);
    for (int i = 0; i < d.ntps; i++) tps3[r * d.ntps + i] = pack(tpl[i]);
  }
  // symbol interleaver H(q) (symbol_inner_interleaver_impl.cc:35-96)
  std::vector<short> H(P), Hinv(P);
  {
    const int Nr = tm == DVBT_T2K ? 11 : 13;
    static const char perm2k[] = {4, 3, 9, 6, 2, 8, 1, 5, 7, 0};
    static const char perm8k[] = {7, 1, 4, 2, 9, 6, 8, 10, 0, 3, 11, 5};
    const char *perm = tm == DVBT_T2K ? perm2k : perm8k;
    int q = 0;
    unsigned reg = 0;
    for (int i = 0; i < d.N; i++) {
      if (i < 2) reg = 0;
      else if (i == 2) reg = 1;
      else {
        unsigned nb = tm == DVBT_T2K ? ((reg ^ (reg >> 3)) & 1u) : ((reg ^ (reg >> 1) ^ (reg >> 4) ^ (reg >> 6)) & 1u);
        reg = ((reg >> 1) | (nb << (Nr - 2))) & ((1u << Nr) - 1u);
      }
      unsigned newreg = 0;
      for (int k = 0; k < Nr - 1; k++) newreg |= ((reg >> k) & 1u) << perm[k];
      int h = ((i % 2) << (Nr - 1)) + (int)newreg;
      if (h < P) {
        if (q < P) H[q] = (short)h;
        q++;
      }
    }
    if (q != P) { set_error("mode tables: symbol interleaver produced %d entries", q); return DVBT_B200_EINVAL; }
    for (int i = 0; i < P; i++) Hinv[H[i]] = (short)i;
  }
  // pack into one device blob
  size_t off = 0;
  auto place = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
  size_t o_cp = place(cpl.size() * 2), o_tps = place(tpl.size() * 2), o_known = place(known.size() * 4), o_pval = place(pval.size() * 4),
         o_kind = place(kind.size()), o_prev = place(prevp.size() * 2), o_next = place(nextp.size() * 2), o_pay = place(payload.size() * 2),
         o_H = place(H.size() * 2), o_Hi = place(Hinv.size() * 2), o_pil = place(pil.size() * 2), o_pay32 = place(pay32.size() * 4),
         o_pay3 = place(pay3.size() * 4), o_tps3 = place(tps3.size() * 4);
  std::vector<unsigned char> host(off);
  memcpy(&host[o_cp], cpl.data(), cpl.size() * 2);
  memcpy(&host[o_tps], tpl.data(), tpl.size() * 2);
  memcpy(&host[o_known], known.data(), known.size() * 4);
  memcpy(&host[o_pval], pval.data(), pval.size() * 4);
  memcpy(&host[o_kind], kind.data(), kind.size());
  memcpy(&host[o_prev], prevp.data(), prevp.size() * 2);
  memcpy(&host[o_next], nextp.data(), nextp.size() * 2);
  memcpy(&host[o_pay], payload.data(), payload.size() * 2);
  memcpy(&host[o_H], H.data(), H.size() * 2);
  memcpy(&host[o_Hi], Hinv.data(), Hinv.size() * 2);
  memcpy(&host[o_pil], pil.data(), pil.size() * 2);
  memcpy(&host[o_pay32], pay32.data(), pay32.size() * 4);
  memcpy(&host[o_pay3], pay3.data(), pay3.size() * 4);
  memcpy(&host[o_tps3], tps3.data(), tps3.size() * 4);
  int rc = blob.reserve(off);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaMemcpy(blob.p, host.data(), off, cudaMemcpyHostToDevice));
  unsigned char *base = blob.as<unsigned char>();
  d.cpilot = (const short *)(base + o_cp);
  d.tps = (const short *)(base + o_tps);
  d.known = (const float *)(base + o_known);
  d.pval = (const float *)(base + o_pval);
  d.kind = base + o_kind;
  d.prevp = (const short *)(base + o_prev);
  d.nextp = (const short *)(base + o_next);
  d.payload = (const short *)(base + o_pay);
  d.H = (const short *)(base + o_H);
  d.Hinv = (const short *)(base + o_Hi);
  d.pilots = (const short *)(base + o_pil);
  d.pay32 = (const int *)(base + o_pay32);
  d.pay3 = (const int *)(base + o_pay3);
  d.tps3 = (const int *)(base + o_tps3);
  return 0;
}

// ---------------------------------------------------------------------------------------
// exact float helpers (operand order of libstdc++ / libgcc, no contraction)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {  // (a.x*b.x - a.y*b.y, a.x*b.y + a.y*b.x)
  return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)), __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) { return cmul(a, make_float2(b.x, -b.y)); }
__device__ __forceinline__ float cnorm(float2 a) { return __fadd_rn(__fmul_rn(a.x, a.x), __fmul_rn(a.y, a.y)); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y)); }
// libgcc __divsc3 for float operands (GCC >= 12): the textbook formula evaluated in double
__device__ __forceinline__ float2 cdiv(float2 n, float2 d) {
  double a = n.x, b = n.y, c = d.x, e = d.y;
  double denom = __dadd_rn(__dmul_rn(c, c), __dmul_rn(e, e));
  double x = __ddiv_rn(__dadd_rn(__dmul_rn(a, c), __dmul_rn(b, e)), denom);
  double y = __ddiv_rn(__dsub_rn(__dmul_rn(b, c), __dmul_rn(a, e)), denom);
  return make_float2((float)x, (float)y);
}

// ---------------------------------------------------------------------------------------
// stage 1: one warp per symbol
// ---------------------------------------------------------------------------------------
// tpsval (optional): the symbol's TPS carriers, equalised exactly as demod_symbol_kernel does it (gain of the pilot below,
// slope of its interval, :617-642, :929-945) - 17 / 68 carriers per symbol, so that the TPS vote and the sequential scan can
// start before the payload is equalised.
__global__ void __launch_bounds__(128) demod_stage1_kernel(ModeDev md, const float2 *__restrict__ X, int nparse,
                                                           int *__restrict__ fo_out, float2 *__restrict__ rot_out,
                                                           int *__restrict__ mod_out, float2 *__restrict__ tpsval) {
  int s = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (s >= nparse) return;
  const float2 *x0 = X + (long long)s * md.N;
  const float2 *x1 = x0 + md.N;
  // ---- process_cpilot_data (:714-744): lanes 0..15 hold candidate i = zl - 8 + lane
  float sum = 0.f;
  if (lane < 16) {
    int i = md.zl - 8 + lane;
    float2 prev = x0[i + md.cpilot[0]];
    for (int j = 0; j < md.ncp - 1; j++) {
      float2 cur = x0[i + md.cpilot[j + 1]];
      float phase = cnorm(csub(cur, prev));
      sum = __fadd_rn(sum, __fmul_rn(md.known[j], phase));
      prev = cur;
    }
  }
  float best = 0.f;
  int start = 0;
  for (int l = 0; l < 16; l++) {  // sequential first-strict-maximum scan
    float v = __shfl_sync(0xffffffffu, sum, l);
    if (v > best) { best = v; start = md.zl - 8 + l; }
  }
  // all-zero input leaves start = 0 in the reference (offset -zl, out-of-bounds reads there);
  // a zero offset is used instead
  int fo = (best > 0.f) ? start - md.zl : 0;
  // ---- compute_oneshot_csft (:746-790): lane 0 = left half, lane 1 = right half
  float angle = 0.f;
  if (lane < 2) {
    int half = (md.ncp - 1) / 2;
    int j0 = lane == 0 ? 0 : half + 1, j1 = lane == 0 ? half : md.ncp;
    float2 acc = make_float2(0.f, 0.f);
    for (int j = j0; j < j1; j++) {
      int idx = fo + md.zl + md.cpilot[j];
      acc = cadd(acc, cmul_conj(x0[idx], x1[idx]));
    }
    angle = atan2f(acc.y, acc.x);
  }
  float left = __shfl_sync(0xffffffffu, angle, 0), right = __shfl_sync(0xffffffffu, angle, 1);
  float corr = __fmul_rn(__fadd_rn(right, left), md.carrier_coeff);
  // ---- frequency_correction (:792-819): one rotor for the whole symbol
  float correction = __fadd_rn((float)fo, corr);
  double ang = __ddiv_rn(__dmul_rn(__dmul_rn(-2.0 * M_PI, (double)correction), (double)(md.N + md.cp)), (double)md.N);
  float sn, cs;
  sincosf((float)ang, &sn, &cs);
  float2 rot = make_float2(cs, sn);
  // ---- process_spilot_data, phase detection (:547-582): lanes 0..3 = candidate phase
  float ssum = 0.f;
  if (lane < 4) {
    float2 c = make_float2(0.f, 0.f);
    for (int j = 0; j < 10; j++) {
      int k = 3 * lane + 12 * j;
      float2 dv = cmul(rot, x0[md.zl + k + fo]);
      c = cadd(c, cmul_conj(make_float2(md.pval[k], 0.f), dv));
    }
    ssum = cnorm(c);
  }
  float smax = 0.f;
  int mod = -1;  // -1: no candidate exceeded 0, the reference keeps the previous value
  for (int l = 0; l < 4; l++) {
    float v = __shfl_sync(0xffffffffu, ssum, l);
    if (v > smax) { smax = v; mod = l; }
  }
  if (lane == 0) {
    fo_out[s] = fo;
    rot_out[s] = rot;
    mod_out[s] = mod;
  }
  if (tpsval) {
    const int r = mod < 0 ? 0 : mod;
    const float2 *x = x0 + md.zl + fo;                             // x[k] = carrier k, integer offset applied
    const short *pil = md.pilots + r * md.pil_stride;
    const int npil = md.npil[r];
    auto pilot_gain = [&](int i) {                                 // :484-489 set_channel_gain
      const int k = pil[i];
      return cdiv(make_float2(md.pval[k], 0.f), cmul(rot, x[k]));
    };
    for (int t = lane; t < md.ntps; t += 32) {
      const int e = md.tps3[r * md.ntps + t];
      const int k = e & 0x1fff, dk = (e >> 13) & 15, o = e >> 17;
      const int nx = o + 1 < npil ? o + 1 : o;
      const float2 g0 = pilot_gain(o);
      const float2 slope = cdiv(csub(pilot_gain(nx), g0), make_float2(11.0f, 0.0f));
      const float2 g = cadd(g0, cmul(slope, make_float2((float)dk, 0.0f)));
      tpsval[(long long)s * md.ntps + t] = cmul(cmul(rot, x[k]), g);
    }
  }
}


// ---------------------------------------------------------------------------------------
// one block per symbol: everything parse_input does to a symbol except the sequential bookkeeping
// ---------------------------------------------------------------------------------------
// The symbol's active carriers (K + 16 around them for the 16 integer-CFO candidates) are staged in shared memory ONCE by
// a bulk asynchronous copy (cp.async.bulk / TMA engine, bulk_copy.cuh) - 8 (K + 17) bytes per symbol instead of the three
// passes over global memory the separate stage-1 and equalise kernels made - while the threads fetch the only other
// input, the next symbol's carriers around the continual pilots (16 x ncp values, one-shot estimate).  Then:
//   all threads   products of process_cpilot_data (:714-744), 16 candidates x (ncp - 1) pilot pairs
//   warp 0        the sequential sums the reference's loops define (same operand order: bit-identical floats) - 16 lanes =
//                 the 16 candidates; first-strict-maximum scan -> integer offset; compute_oneshot_csft (:746-790): terms by all
//                 lanes, the two half sums by lanes 0 / 1; rotor (:792-819); scattered-pilot phase (:547-582): 40 terms, 4 sums
//   all threads   pilot gains tx / rx (:484-489; complex division as libgcc's __divsc3), interval slopes with the reference's
//                 constant /11 (:617-642) - kept by pilot ordinal, not by carrier
//   all threads   TPS carriers (:929-945) and payload cells (:1104-1113), four consecutive cells per thread: one 16-byte
//                 table load, gains interpolated where they are used, fused demap, ONE 4-byte store of the four demapped
//                 cells (and two 16-byte stores of the equalised cells when the caller wants them)
// FRONT = false (default): the integer offset, rotor and scattered-pilot phase come from demod_stage1_kernel (one warp per
// symbol, good occupancy for its latency-bound sums) and this kernel does the wide part only; FRONT = true: everything in
// one kernel (DVBT_B200_DEMOD_FUSED=1; measured slower on B200: the other eleven warps of the block wait for warp 0's sums).
// SOFT (only with FRONT = false): the demap writes one word of soft decisions per cell to `soft` instead of the hard byte to
// `dm` (demap_cell_soft); the channel-state weight of a cell is |H|^2 = 1 / |gain|^2 over its mean at the symbol's pilots.
template <bool FRONT, int kSymThreads, bool SOFT = false>
__global__ void __launch_bounds__(kSymThreads, FRONT ? 1 : SOFT ? 1 : 1536 / kSymThreads) demod_symbol_kernel(ModeDev md, const __grid_constant__ DemapTable dt, int do_demap, int use_bulk,
                                                                   const float2 *__restrict__ X, int *fo_out, float2 *rot_out, int *mod_out,
                                                                   float2 *__restrict__ tpsval, float2 *__restrict__ Y,
                                                                   uint8_t *__restrict__ dm, uint32_t *__restrict__ soft) {
  extern __shared__ __align__(16) unsigned char s_sym[];
  __shared__ uint64_t s_bar;
  __shared__ int s_fo, s_mod;
  __shared__ float2 s_rot;
  const int s = blockIdx.x, t = threadIdx.x, lane = t & 31;
  const int K = md.K, ncp = md.ncp, nxs = K + 17;
  float2 *xs = reinterpret_cast<float2 *>(s_sym);                  // xs[i] = X[s][zl - 8 + i]
  float2 *x1seg = xs + ((nxs + 1) & ~1);                           // [ncp][16]: next symbol at zl - 8 + cpilot[j] + 0..15
  float *prod = reinterpret_cast<float *>(x1seg + ncp * 16);       // [ncp - 1][16]
  float2 *t1 = reinterpret_cast<float2 *>(prod + 16 * ncp);        // [ncp] one-shot terms
  float2 *sp = t1 + ncp;                                           // [40] scattered-pilot terms
  float2 *gain = FRONT ? sp + 40 : xs + ((nxs + 1) & ~1);          // [npil] by pilot ordinal
  float2 *slope = gain + md.pil_stride;
  const float2 *x0g = X + (long long)s * md.N + md.zl - 8;
  const float2 *x1g = x0g + md.N;
  if (t == 0) mbar_init(&s_bar, 1);
  __syncthreads();
  if (use_bulk && t == 0) bulk_g2s(xs, x0g, (unsigned)nxs * 8u, &s_bar);
  if (FRONT)
    for (int i = t; i < ncp * 16; i += kSymThreads) x1seg[i] = x1g[md.cpilot[i >> 4] + (i & 15)];
  if (!FRONT && t == 0) { s_fo = fo_out[s]; s_rot = rot_out[s]; s_mod = mod_out[s]; }
  if (use_bulk) {
    mbar_wait(&s_bar, 0);
  } else {   // the caller's buffer is not 16-byte aligned: plain loads
    for (int i = t; i < nxs; i += kSymThreads) xs[i] = x0g[i];
  }
  __syncthreads();
  if (FRONT) {
  // ---- process_cpilot_data (:714-744): candidate l <-> offset l - 8
  for (int i = t; i < 16 * (ncp - 1); i += kSymThreads) {
    const int j = i >> 4, l = i & 15;
    const float2 prev = xs[l + md.cpilot[j]], cur = xs[l + md.cpilot[j + 1]];
    prod[i] = __fmul_rn(md.known[j], cnorm(csub(cur, prev)));
  }
  __syncthreads();
  if (t < 32) {
    float sum = 0.f;
    if (lane < 16)
      for (int j = 0; j < ncp - 1; j++) sum = __fadd_rn(sum, prod[16 * j + lane]);
    float best = 0.f;
    int start = 0;
    for (int l = 0; l < 16; l++) {  // sequential first-strict-maximum scan
      float v = __shfl_sync(0xffffffffu, sum, l);
      if (v > best) { best = v; start = md.zl - 8 + l; }
    }
    // all-zero input leaves start = 0 in the reference (offset -zl, out-of-bounds reads there); a zero offset is used instead
    const int fo = (best > 0.f) ? start - md.zl : 0;
    const float2 *x = xs + 8 + fo;                                 // x[k] = carrier k of this symbol, integer offset applied
    // ---- compute_oneshot_csft (:746-790): terms by all lanes, left half by lane 0, right half by lane 1
    for (int j = lane; j < ncp; j += 32) t1[j] = cmul_conj(x[md.cpilot[j]], x1seg[16 * j + 8 + fo]);
    __syncwarp();
    float angle = 0.f;
    if (lane < 2) {
      const int half = (ncp - 1) / 2;
      const int j0 = lane == 0 ? 0 : half + 1, j1 = lane == 0 ? half : ncp;
      float2 acc = make_float2(0.f, 0.f);
      for (int j = j0; j < j1; j++) acc = cadd(acc, t1[j]);
      angle = atan2f(acc.y, acc.x);
    }
    const float left = __shfl_sync(0xffffffffu, angle, 0), right = __shfl_sync(0xffffffffu, angle, 1);
    const float corr = __fmul_rn(__fadd_rn(right, left), md.carrier_coeff);
    // ---- frequency_correction (:792-819): one rotor for the whole symbol
    const float correction = __fadd_rn((float)fo, corr);
    const double ang = __ddiv_rn(__dmul_rn(__dmul_rn(-2.0 * M_PI, (double)correction), (double)(md.N + md.cp)), (double)md.N);
    float sn, cs;
    sincosf((float)ang, &sn, &cs);
    const float2 rot = make_float2(cs, sn);
    // ---- process_spilot_data, phase detection (:547-582): candidate phase p, pilot j -> term 10 p + j
    for (int i = lane; i < 40; i += 32) {
      const int k = 3 * (i / 10) + 12 * (i % 10);
      sp[i] = cmul_conj(make_float2(md.pval[k], 0.f), cmul(rot, x[k]));
    }
    __syncwarp();
    float ssum = 0.f;
    if (lane < 4) {
      float2 c = make_float2(0.f, 0.f);
      for (int j = 0; j < 10; j++) c = cadd(c, sp[10 * lane + j]);
      ssum = cnorm(c);
    }
    float smax = 0.f;
    int mod = -1;  // -1: no candidate exceeded 0, the reference keeps the previous value
    for (int l = 0; l < 4; l++) {
      float v = __shfl_sync(0xffffffffu, ssum, l);
      if (v > smax) { smax = v; mod = l; }
    }
    if (lane == 0) {
      s_fo = fo; s_mod = mod; s_rot = rot;
      fo_out[s] = fo;
      rot_out[s] = rot;
      mod_out[s] = mod;
    }
  }
  __syncthreads();
  }   // FRONT
  const int fo = s_fo;
  const float2 rot = s_rot;
  int r = s_mod;
  if (r < 0) r = 0;  // degenerate all-zero symbol; the scan resolves the index bookkeeping
  const float2 *x = xs + 8 + fo;
  // ---- pilot gains: gain = tx / rx (:484-489 set_channel_gain), then the slope of every pilot interval (:617-642)
  const short *pil = md.pilots + r * md.pil_stride;
  const int npil = md.npil[r];
  for (int i = t; i < npil; i += kSymThreads) {
    const int k = pil[i];
    gain[i] = cdiv(make_float2(md.pval[k], 0.f), cmul(rot, x[k]));
  }
  __syncthreads();
  for (int i = t; i < npil; i += kSymThreads) {
    const int nx = i + 1 < npil ? i + 1 : i;                       // the last pilot (carrier Kmax) closes its own interval
    slope[i] = cdiv(csub(gain[nx], gain[i]), make_float2(11.0f, 0.0f));
  }
  __syncthreads();
  float soft_f = 0.f;   // soft_scale / (4 g^2 mean |H|^2)
  if constexpr (SOFT) {
    __shared__ float s_part[kSymThreads / 32];
    __shared__ float s_f;
    float acc = 0.f;
    for (int i = t; i < npil; i += kSymThreads) acc += 1.0f / (gain[i].x * gain[i].x + gain[i].y * gain[i].y);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_part[t >> 5] = acc;
    __syncthreads();
    if (t == 0) {
      float sum = 0.f;
      for (int i = 0; i < kSymThreads / 32; i++) sum += s_part[i];
      const float mean = sum / (float)npil;
      s_f = (mean > 0.f && mean < 3.0e38f) ? dt.soft_scale / (4.0f * dt.g * dt.g * mean) : 0.f;   // a dead symbol: everything erased
    }
    __syncthreads();
    soft_f = s_f;
  }
  // gain of a data carrier: g[k0] + (k - k0) * slope[k0], k0 the channel-estimation carrier below (same operations in the
  // same order as the reference's loop over the interval, so the same floats)
  auto cell = [&](int e) {
    const int k = e & 0x1fff, dk = (e >> 13) & 15, o = e >> 17;
    const float2 g = cadd(gain[o], cmul(slope[o], make_float2((float)dk, 0.0f)));
    return cmul(cmul(rot, x[k]), g);
  };
  if (tpsval && t < md.ntps) tpsval[(long long)s * md.ntps + t] = cell(md.tps3[r * md.ntps + t]);   // :929-945 (null: stage 1 did it)
  const int4 *pay = reinterpret_cast<const int4 *>(md.pay3 + r * md.P);
  for (int q = t; q < md.P / 4; q += kSymThreads) {                // :1104-1113
    const int4 e = pay[q];
    const float2 c0 = cell(e.x), c1 = cell(e.y), c2 = cell(e.z), c3 = cell(e.w);
    const long long o = (long long)s * md.P + 4 * q;
    if (Y) {
      float4 *yo = reinterpret_cast<float4 *>(Y + o);
      yo[0] = make_float4(c0.x, c0.y, c1.x, c1.y);
      yo[1] = make_float4(c2.x, c2.y, c3.x, c3.y);
    }
    if constexpr (SOFT) {
      auto weight = [&](int e) {
        const int dk = (e >> 13) & 15, o = e >> 17;
        const float2 g = cadd(gain[o], cmul(slope[o], make_float2((float)dk, 0.0f)));
        return soft_f / (g.x * g.x + g.y * g.y);
      };
      *reinterpret_cast<uint4 *>(soft + o) = make_uint4(demap_cell_soft_any(dt, c0, weight(e.x)), demap_cell_soft_any(dt, c1, weight(e.y)),
                                                        demap_cell_soft_any(dt, c2, weight(e.z)), demap_cell_soft_any(dt, c3, weight(e.w)));
    } else if (do_demap) {
      const uint32_t d = (uint32_t)demap_cell_any(dt, c0) | ((uint32_t)demap_cell_any(dt, c1) << 8) | ((uint32_t)demap_cell_any(dt, c2) << 16) |
                         ((uint32_t)demap_cell_any(dt, c3) << 24);
      *reinterpret_cast<uint32_t *>(dm + o) = d;
    }
  }
}

size_t demod_symbol_smem(const ModeDev &md, bool front) {
  const size_t nxs = (size_t)md.K + 17;
  if (!front) return ((nxs + 1) & ~(size_t)1) * 8 + (size_t)md.pil_stride * 16 + 64;
  return ((nxs + 1) & ~(size_t)1) * 8 + (size_t)md.ncp * 16 * 8 + (size_t)md.ncp * 16 * 4 + (size_t)md.ncp * 8 + 40 * 8 + (size_t)md.pil_stride * 16 + 64;
}

// TPS DBPSK majority vote against the previous symbol (:929-948).  The low half of vote[s] is the vote (|v| <= 68);
// bit 16 marks a symbol that carries a sync_start tag (demod_reference_signals_impl.cc:44-52), so that the scan
// below finds the (rare) tags in the data it stages anyway.
__global__ void demod_vote_kernel(int ntps, int nparse, const float2 *__restrict__ tpsval,
                                  const DemodState *__restrict__ state, int *__restrict__ vote,
                                  int sync_start_at0, const int *__restrict__ sync_at, int nsync) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nparse) return;
  int flag = (sync_start_at0 && s == 0) ? 1 : 0;
  for (int q = 0; q < nsync; q++) flag |= (sync_at[q] == s) ? 1 : 0;
  const float2 *cur = tpsval + (long long)s * ntps;
  int v = 0;
  for (int k = 0; k < ntps; k++) {
    float2 prev = s > 0 ? tpsval[(long long)(s - 1) * ntps + k] : state->prev_tps[k];
    float2 ph = cmul_conj(cur[k], prev);
    v += (ph.x >= 0.0f) ? 1 : -1;
  }
  vote[s] = (v & 0xffff) | (flag << 16);
}

// BCH(127,113) shortened to (67,53): verify_bch_code (:384-425), on the bit-packed FIFO
// TPS helpers of the scan kernel below.  The 68-entry TPS FIFO (d_rcv_tps_data) is a bit set: entry i = bit i
// of (lo, hi).  The BCH remainder is linear in the 53 data bits (FIFO entries 1..53): lane j keeps the 14-bit
// remainder R[j] of the unit vector at data bit j (and j+32), computed once per launch with the
// reference's LFSR; a check is then one masked XOR per lane and a 5-step warp reduction.
__device__ unsigned tps_bch_unit_remainder(int j) {
  unsigned reg = 0;
  for (int i = 0; i < 113; i++) {
    unsigned b = (i == 60 + j) ? 1u : 0u;
    unsigned fb = 1u & (b ^ reg);
    reg >>= 1;
    reg |= fb << 13;
    reg ^= (fb << 12) ^ (fb << 11) ^ (fb << 9) ^ (fb << 8) ^ (fb << 7) ^ (fb << 5) ^ (fb << 4);
  }
  return reg;
}
// all 32 lanes must call this together
__device__ bool tps_bch_ok_bits(unsigned long long lo, unsigned hi, unsigned r_lo, unsigned r_hi, int lane) {
  unsigned long long data = lo >> 1;  // data bit j = FIFO entry 1 + j
  unsigned acc = ((data >> lane) & 1ull) ? r_lo : 0u;
  if (lane + 32 < 53 && ((data >> (lane + 32)) & 1ull)) acc ^= r_hi;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, o);
  unsigned parity = (unsigned)((lo >> 54) & 0x3FFull) | ((hi & 0xFu) << 10);  // entries 54..67
  return parity == (acc & 0x3FFFu);
}

constexpr int kScanChunk = 4096;     // symbols staged in shared memory for the symbol-by-symbol path
constexpr int kScanFrames = 2048;    // frames validated per parallel round
constexpr int kScanWarps = 32;

// The sequential bookkeeping of parse_input (:1188-1248) and of the block (:108-149) as one block of 32 warps.
//   * Warp 0 runs the reference's per-symbol state machine (all lanes the same scalar code, lane 0 writes)
//     until the receiver is in lock at a frame boundary: known, symbol_index == 67, FIFO just cleared.
//   * From there on a whole 68-symbol frame is equivalent to 68 single steps provided every symbol advances
//     the index by one, no sync word shows up early in the partly filled FIFO, and the frame ends with a
//     valid sync word + BCH.  A good frame leaves the scattered-pilot phase where it was (68 = 0 mod 4), so
//     these conditions do not depend on the frames before: ALL following frames are checked in parallel
//     (one warp per frame: ballots build the 68 TPS bits, the BCH remainder is a warp XOR-reduce).
//   * Warp 0 then walks the per-frame verdicts (superframe gating is the only state), all threads write the
//     output descriptors of the accepted frames, and the first rejected frame is handed back to the
//     per-symbol machine, exactly as the one-frame-at-a-time loop did.
//   * A sync_start tag (bit 16 of vote[], see demod_vote_kernel) re-arms the wait for a superframe start on the symbol
//     that carries it (demod_reference_signals_impl.cc:112-116); a frame that contains one is never taken whole.
__global__ void __launch_bounds__(32 * kScanWarps) demod_scan_kernel(int ntps, int nparse, int fi_start, int src_base,
                                                                     const int *__restrict__ mod_in, const int *__restrict__ vote,
                                                                     const float2 *__restrict__ tpsval, DemodState *st,
                                                                     int *__restrict__ out_symidx, int *__restrict__ out_src) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // sync words of :121-126 as FIFO entries 1..15 (std::equal compares 15 elements, :975/:1002)
  const unsigned long long kEven = (1ull << 3) | (1ull << 4) | (1ull << 6) | (1ull << 8) | (1ull << 9) | (1ull << 10) | (1ull << 11) |
                                   (1ull << 13) | (1ull << 14) | (1ull << 15);
  const unsigned long long kMask = 0xFFFEull;
  __shared__ signed char s_mod[kScanChunk];   // phase in -1..3
  __shared__ short s_vote[kScanChunk];        // |vote| <= 68 in the low byte (sign extended), bit 8: sync_start on this symbol
  __shared__ unsigned char s_ok[kScanFrames], s_fi[kScanFrames];
  __shared__ int s_lock, s_prev_mod, s_nframes, s_emit_from, s_nf, s_out_base;
  const unsigned r_lo = tps_bch_unit_remainder(lane), r_hi = lane + 32 < 53 ? tps_bch_unit_remainder(lane + 32) : 0u;

  // ---- state of the sequential machine (meaningful in warp 0 only)
  int symbol_index = 0, known = 0, frame_index = 0, prev_mod = 0, cur_mod = 0, d_init = 0;
  unsigned long long lo = 0;
  unsigned hi = 0;
  int first_out = -1, n_out = 0, sf_tag_at = -1, n_sf = 0;
  auto record_sf = [&](int at) {   // warp 0, all lanes the same value
    if (sf_tag_at < 0) sf_tag_at = at;
    if (n_sf < kMaxSfTags && lane == 0) st->sf_at[n_sf] = at;
    n_sf++;
  };
  int cbase = -(1 << 30);
  int s = 0;
  bool skip_fast = false;
  if (warp == 0) {
    symbol_index = st->symbol_index; known = st->known; frame_index = st->frame_index; prev_mod = st->prev_mod;
    cur_mod = st->mod; d_init = st->d_init;
    unsigned long long b0 = lane < 32 ? (unsigned long long)(st->fifo[lane] & 1) : 0ull, b1 = (unsigned long long)(st->fifo[32 + lane] & 1);
    unsigned b2 = lane < 4 ? (unsigned)(st->fifo[64 + lane] & 1) : 0u;
    unsigned w0 = __ballot_sync(0xffffffffu, b0 != 0), w1 = __ballot_sync(0xffffffffu, b1 != 0), w2 = __ballot_sync(0xffffffffu, b2 != 0);
    lo = ((unsigned long long)w1 << 32) | w0;
    hi = w2 & 0xFu;
  }
  auto load_chunk = [&](int s0) {   // warp 0 only
    __syncwarp();
    for (int i0 = 0; i0 < kScanChunk && s0 + i0 < nparse; i0 += 32 * 8) {
      int m[8], v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        int idx = s0 + i0 + lane + 32 * u;
        m[u] = idx < nparse ? mod_in[idx] : 0;
        v[u] = idx < nparse ? vote[idx] : 0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        s_mod[i0 + lane + 32 * u] = (signed char)m[u];
        s_vote[i0 + lane + 32 * u] = (short)(((int)(signed char)(v[u] & 0xff) & 0xff) | (((v[u] >> 16) & 1) << 8));
      }
    }
    cbase = s0;
    __syncwarp();
  };

  for (;;) {
    // ================= warp 0: symbol by symbol until in lock at a frame boundary (or out of input)
    if (warp == 0) {
      while (s < nparse) {
        if (!skip_fast && known && symbol_index == 67 && lo == 0ull && hi == 0u && s + 68 <= nparse) break;
        skip_fast = false;
        // ---- one symbol (parse_input :1188-1248 bookkeeping)
        if (s < cbase || s >= cbase + kScanChunk) load_chunk(s);
        int m_in = s_mod[s - cbase];
        int v_pk = s_vote[s - cbase];
        int v_in = (int)(signed char)(v_pk & 0xff);
        if (v_pk & 0x100) d_init = 0;           // sync_start on this item (demod_reference_signals_impl.cc:115-116)
        int mod = m_in >= 0 ? m_in : cur_mod;
        cur_mod = mod;
        int diff = (mod - prev_mod + 4) & 3;  // :684-688
        prev_mod = mod;
        symbol_index += diff;                 // :1228
        if (symbol_index >= 68) symbol_index -= 68;
        int sym_out = symbol_index, frame_out = frame_index;
        bool cond = !known || symbol_index != 0;
        unsigned long long bitv = cond ? (v_in >= 0 ? 0ull : 1ull) : 0ull;
        for (int d = 0; d < diff; d++) {      // :957-972: pop front, push back
          lo = (lo >> 1) | ((unsigned long long)(hi & 1u) << 63);
          hi = (hi >> 1) | ((unsigned)bitv << 3);
        }
        bool even = (lo & kMask) == kEven, odd = (lo & kMask) == (kEven ^ kMask);
        int end_frame = 0;
        if (even || odd) {
          if (tps_bch_ok_bits(lo, hi, r_lo, r_hi, lane)) {
            frame_index = (int)((((lo >> 23) & 1ull) << 1) | ((lo >> 24) & 1ull));
            known = 1;
            end_frame = 1;
          } else {
            known = 0;
          }
          lo = 0;
          hi = 0;
        }
        if (end_frame) symbol_index = 67;     // :1240-1241
        // block level (demod_reference_signals_impl.cc:118-143)
        bool emit = true;
        if (d_init == 0) {
          if (sym_out == 0 && (frame_out & 3) == fi_start) {
            d_init = 1;
            record_sf(n_out);
          } else {
            emit = false;
          }
        }
        if (emit) {
          if (first_out < 0) first_out = s;
          if (lane == 0) { out_symidx[n_out] = sym_out; out_src[n_out] = src_base + s; }
          n_out++;
        }
        s++;
      }
      if (lane == 0) {
        s_lock = s < nparse ? s : -1;
        s_prev_mod = prev_mod;
        int nfr = s < nparse ? (nparse - s) / 68 : 0;
        s_nframes = nfr < kScanFrames ? nfr : kScanFrames;
      }
    }
    __syncthreads();
    const int lock = s_lock;
    if (lock < 0) break;
    const int nframes = s_nframes, pm = s_prev_mod;
    // ================= all warps: verdict of every frame that follows
    for (int f = warp; f < nframes; f += kScanWarps) {
      const int fs = lock + 68 * f;
      bool good = true;
      unsigned w[3];
#pragma unroll
      for (int q = 0; q < 3; q++) {
        int i = lane + 32 * q;
        bool valid = i < 68;
        int cm = valid ? mod_in[fs + i] : 0;
        int cv = valid ? vote[fs + i] : 0;
        if (valid && cm != ((pm + 1 + i) & 3)) good = false;
        if (valid && (cv & (1 << 16))) good = false;        // a sync_start inside the frame: symbol by symbol
        w[q] = __ballot_sync(0xffffffffu, valid && (cv & 0x8000));
      }
      good = __all_sync(0xffffffffu, good);
      unsigned long long flo = (((unsigned long long)w[1] << 32) | w[0]) & ~1ull;  // entry 0: index 0 pushes 0 (:964-971)
      unsigned fhi = w[2] & 0xFu;
      bool early = false;
#pragma unroll
      for (int k = 1; k <= 3; k++) {
        unsigned long long e = (flo << k) & kMask;
        if (e == kEven || e == (kEven ^ kMask)) early = true;
      }
      bool match = ((flo & kMask) == kEven) || ((flo & kMask) == (kEven ^ kMask));
      bool ok = good && !early && match && tps_bch_ok_bits(flo, fhi, r_lo, r_hi, lane);
      if (lane == 0) {
        s_ok[f] = ok ? 1 : 0;
        s_fi[f] = (unsigned char)((((flo >> 23) & 1ull) << 1) | ((flo >> 24) & 1ull));
      }
    }
    __syncthreads();
    // ================= warp 0: walk the verdicts (superframe gating), first rejected frame ends the run
    if (warp == 0) {
      int f = 0, emit_from = -1;
      const int out_base = n_out;
      while (f < nframes && s_ok[f]) {
        bool emit = true;
        if (d_init == 0) {
          if ((frame_index & 3) == fi_start) { d_init = 1; record_sf(n_out); }
          else emit = false;
        }
        if (emit) {
          if (first_out < 0) first_out = lock + 68 * f;
          if (emit_from < 0) emit_from = f;
          n_out += 68;
        }
        frame_index = s_fi[f];
        f++;
      }
      // every accepted frame: cur_mod = prev_mod = (prev_mod + 68) & 3, FIFO cleared, symbol_index 67 - all as they were
      s = lock + 68 * f;
      if (f > 0) cur_mod = prev_mod;       // (:684-688 over 68 symbols; 68 = 0 mod 4)
      if (f < nframes) skip_fast = true;   // the frame at s is not a clean one: the per-symbol machine takes it
      if (lane == 0) { s_nf = f; s_emit_from = emit_from; s_out_base = out_base; }
    }
    __syncthreads();
    // ================= all threads: descriptors of the accepted, emitted frames
    {
      const int nf = s_nf, ef = s_emit_from, ob = s_out_base;
      if (ef >= 0) {
        const int count = (nf - ef) * 68;
        for (int i = threadIdx.x; i < count; i += blockDim.x) {
          out_symidx[ob + i] = i % 68;
          out_src[ob + i] = src_base + lock + 68 * ef + i;
        }
      }
    }
    __syncthreads();
  }
  if (warp == 0) {
    if (lane == 0) {
      st->symbol_index = symbol_index; st->known = known; st->frame_index = frame_index; st->prev_mod = prev_mod;
      st->mod = cur_mod; st->d_init = d_init;
      for (int i = 0; i < 64; i++) st->fifo[i] = (unsigned char)((lo >> i) & 1ull);
      for (int i = 64; i < 68; i++) st->fifo[i] = (unsigned char)((hi >> (i - 64)) & 1u);
      st->first_out = first_out; st->n_out = n_out; st->sf_tag_at = sf_tag_at; st->n_sf = n_sf;
    }
    if (nparse > 0)
      for (int k = lane; k < ntps; k += 32) st->prev_tps[k] = tpsval[(long long)(nparse - 1) * ntps + k];
  }
}

int demod_run(const ModeDev &md, const DemapTable *demap, const float2 *X, int nparse, DemodBuffers b, DemodState *d_state,
              int fi_start, int sync_start_at0, float2 *Y, uint8_t *dm, cudaStream_t st, const int *sync_at, int nsync, int src_base,
              uint32_t *soft) {
  if (nparse <= 0) return 0;
  if (soft && (!demap || ((uintptr_t)soft & 15))) { set_error("demod: soft decisions need a demap table and a 16-byte aligned buffer"); return DVBT_B200_EINVAL; }
  {
    static const bool fused_env = getenv("DVBT_B200_DEMOD_FUSED") && atoi(getenv("DVBT_B200_DEMOD_FUSED")) != 0;
    const bool fused_front = fused_env && !soft;
    static const bool side_env = !(getenv("DVBT_B200_DEMOD_SIDE_SCAN") && atoi(getenv("DVBT_B200_DEMOD_SIDE_SCAN")) == 0);   // =0: scan behind the equalise kernel (A/B)
    const bool tps_early = !fused_front && side_env && b.side && b.ev_fork && b.ev_join;
    if (!fused_front) {
      const int threads = 128;
      const long long total = (long long)nparse * 32;
      demod_stage1_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(md, X, nparse, b.fo, b.rot, b.modidx,
                                                                                            tps_early ? b.tpsval : nullptr);
      DVBT_CUDA_TRY(cudaGetLastError());
      count_launch();
    }
    if (tps_early) {
      // fork: the TPS vote and the scan on the side stream, beside the equalise + demap kernel below.  The scan is ONE block
      // that must be resident before the wide kernel takes every register of every SM (streams do not preempt): the main
      // stream therefore waits for the (short) vote, so that scan and equalise become ready together, and the side stream
      // has the higher priority - its block is placed first, the wide kernel fills the rest of the GPU.
      DVBT_CUDA_TRY(cudaEventRecord(b.ev_fork, st));
      DVBT_CUDA_TRY(cudaStreamWaitEvent(b.side, b.ev_fork, 0));
      demod_vote_kernel<<<(nparse + 127) / 128, 128, 0, b.side>>>(md.ntps, nparse, b.tpsval, d_state, b.vote, sync_start_at0, sync_at, sync_at ? nsync : 0);
      DVBT_CUDA_TRY(cudaGetLastError());
      if (b.ev_mid) {
        DVBT_CUDA_TRY(cudaEventRecord(b.ev_mid, b.side));
        DVBT_CUDA_TRY(cudaStreamWaitEvent(st, b.ev_mid, 0));
      }
      demod_scan_kernel<<<1, 32 * kScanWarps, 0, b.side>>>(md.ntps, nparse, fi_start, src_base, b.modidx, b.vote, b.tpsval, d_state, b.out_symidx, b.out_src);
      DVBT_CUDA_TRY(cudaGetLastError());
      DVBT_CUDA_TRY(cudaEventRecord(b.ev_join, b.side));
      count_launch(2);
    }
    const size_t smem = demod_symbol_smem(md, fused_front);
    // threads per block: the payload loop takes P / 4 = 378 (1512) quads of cells per symbol - 192 threads make it two (eight)
    // full passes, and eight blocks (1536 threads, <= 42 registers) fit an SM
    static const int nt_env = getenv("DVBT_B200_DEMOD_THREADS") ? atoi(getenv("DVBT_B200_DEMOD_THREADS")) : 192;
    const int kSymThreads = fused_front ? 384 : soft ? 192 : (nt_env == 128 || nt_env == 256 || nt_env == 384) ? nt_env : 192;
    void (*symk)(ModeDev, const DemapTable, int, int, const float2 *, int *, float2 *, int *, float2 *, float2 *, uint8_t *, uint32_t *) =
        soft ? demod_symbol_kernel<false, 192, true> :
        fused_front ? demod_symbol_kernel<true, 384>
        : kSymThreads == 128 ? demod_symbol_kernel<false, 128>
        : kSymThreads == 256 ? demod_symbol_kernel<false, 256>
        : kSymThreads == 384 ? demod_symbol_kernel<false, 384> : demod_symbol_kernel<false, 192>;
    // per launch, like every other kernel of the library: the attribute is per device and handles live on any device
    DVBT_CUDA_TRY(cudaFuncSetAttribute(symk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DemapTable dummy;
    dummy.size = 0;
    // bulk copies and the vector stores want 16-byte aligned buffers (every buffer of this library is; a caller's may not be)
    const int aligned = (((uintptr_t)X & 15) == 0) ? 1 : 0;
    if ((Y && ((uintptr_t)Y & 15)) || (dm && ((uintptr_t)dm & 3))) { set_error("demod: output buffers must be 16-byte (cells) / 4-byte (demapped) aligned"); return DVBT_B200_EINVAL; }
    if (b.ev_eq0) DVBT_CUDA_TRY(cudaEventRecord(b.ev_eq0, st));
    symk<<<nparse, kSymThreads, smem, st>>>(md, demap ? *demap : dummy, (demap && dm) ? 1 : 0, aligned, X, b.fo, b.rot, b.modidx,
                                                          tps_early ? nullptr : b.tpsval, Y, dm, soft);
    DVBT_CUDA_TRY(cudaGetLastError());
    if (b.ev_eq1) DVBT_CUDA_TRY(cudaEventRecord(b.ev_eq1, st));
    if (tps_early) {
      DVBT_CUDA_TRY(cudaStreamWaitEvent(st, b.ev_join, 0));       // join: the scan's results are ordered before everything behind
      count_launch(1);
      return 0;
    }
  }
  demod_vote_kernel<<<(nparse + 127) / 128, 128, 0, st>>>(md.ntps, nparse, b.tpsval, d_state, b.vote, sync_start_at0, sync_at, sync_at ? nsync : 0);
  DVBT_CUDA_TRY(cudaGetLastError());
  demod_scan_kernel<<<1, 32 * kScanWarps, 0, st>>>(md.ntps, nparse, fi_start, src_base, b.modidx, b.vote, b.tpsval, d_state, b.out_symidx, b.out_src);
  DVBT_CUDA_TRY(cudaGetLastError());
  count_launch(3);
  return 0;
}

}  // namespace dvbt

// ---------------------------------------------------------------------------------------
// block ABI
// ---------------------------------------------------------------------------------------
struct dvbt_b200_demod {
  int device = dvbt::current_device();
  dvbt_b200_demod_params par;
  dvbt::ModeTables tables;
  cudaStream_t stream = nullptr;
  dvbt::DevBuf d_X, d_Y, d_Yc, d_state, d_fo, d_rot, d_mod, d_tps, d_vote, d_osym, d_osrc, h_state;
  int fi_start = 3;
  bool pending_sync = false;
  dvbt::Staging stg;
};

namespace {
__global__ void demod_compact_kernel(int P, const int *__restrict__ src, const float2 *__restrict__ Y, float2 *__restrict__ out) {
  int o = blockIdx.x;
  const float2 *y = Y + (long long)src[o] * P;
  for (int i = threadIdx.x; i < P; i += blockDim.x) out[(long long)o * P + i] = y[i];
}
}  // namespace

extern "C" {

int dvbt_b200_demod_create(const dvbt_b200_demod_params *p, dvbt_b200_demod **out) {
  if (!p || !out) { dvbt::set_error("demod_create: null argument"); return DVBT_B200_EINVAL; }
  *out = nullptr;
  int rc = dvbt::ensure_device();
  if (rc) return rc;
  dvbt_b200_demod *h = new (std::nothrow) dvbt_b200_demod();
  if (!h) { dvbt::set_error("demod_create: out of memory"); return DVBT_B200_ENOMEM; }
  h->par = *p;
  rc = h->tables.init(p->transmission_mode, p->guard_interval);
  if (!rc && (p->ninput != h->tables.dev.N || p->noutput != h->tables.dev.P || p->itemsize != 8)) {
    dvbt::set_error("demod_create: itemsize/ninput/noutput (%d,%d,%d) do not match the mode (8,%d,%d)", p->itemsize, p->ninput,
                    p->noutput, h->tables.dev.N, h->tables.dev.P);
    rc = DVBT_B200_EINVAL;
  }
  if (!rc && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    dvbt::set_error("demod_create: cannot create stream");
    rc = DVBT_B200_ECUDA;
  }
  if (!rc) rc = h->d_state.reserve(sizeof(dvbt::DemodState));
  h->h_state.host = true;
  if (!rc) rc = h->h_state.reserve(sizeof(dvbt::DemodState));
  if (rc) { dvbt_b200_demod_destroy(h); return rc; }
  cudaMemset(h->d_state.p, 0, sizeof(dvbt::DemodState));
  // demod_reference_signals_impl.cc:73-77 ("TODO investigate" upstream)
  h->fi_start = (p->constellation == DVBT_QAM64 && p->transmission_mode == DVBT_T8K) ? 2 : 3;
  *out = h;
  return 0;
}

void dvbt_b200_demod_destroy(dvbt_b200_demod *h) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) return;
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  dvbt::DevBuf *bufs[] = {&h->d_X, &h->d_Y, &h->d_Yc, &h->d_state, &h->d_fo, &h->d_rot, &h->d_mod, &h->d_tps, &h->d_vote, &h->d_osym, &h->d_osrc, &h->h_state};
  for (auto *b : bufs) b->release();
  h->stg.release();
  h->tables.release();
  delete h;
}

int dvbt_b200_demod_work(dvbt_b200_demod *h, const void *in, size_t n_in_items, void *out, size_t out_capacity_items,
                         size_t *consumed, size_t *produced, const dvbt_b200_tag *tags_in, size_t n_tags_in,
                         dvbt_b200_tag *tags_out, size_t tags_out_capacity, size_t *n_tags_out) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || !consumed || !produced) { dvbt::set_error("demod_work: null argument"); return DVBT_B200_EINVAL; }
  *consumed = *produced = 0;
  if (n_tags_out) *n_tags_out = 0;
  if (n_in_items < 2 || out_capacity_items == 0) return 0;  // forecast: 2 input items per output item
  if (!in || !out) { dvbt::set_error("demod_work: null buffer"); return DVBT_B200_EINVAL; }
  const dvbt::ModeDev &md = h->tables.dev;
  size_t nparse = n_in_items - 1;
  if (nparse > out_capacity_items) nparse = out_capacity_items;
  // a sync_start tag anywhere in the window re-arms the gating (is_sync_start, :44-52); the reference
  // parses one item per call, so honour only a tag on the first item and stop the window before a later one
  bool sync0 = false;
  for (size_t i = 0; i < n_tags_in; i++) {
    if (tags_in[i].key != DVBT_TAG_SYNC_START) continue;
    if (tags_in[i].offset == 0) sync0 = true;
    else if (tags_in[i].offset < nparse) nparse = (size_t)tags_in[i].offset;
  }
  int rc;
  size_t nsym = nparse + 1;
  if ((rc = h->d_X.reserve(nsym * md.N * 8))) return rc;
  if ((rc = h->d_Y.reserve(nparse * md.P * 8))) return rc;
  if ((rc = h->d_Yc.reserve(nparse * md.P * 8))) return rc;
  if ((rc = h->d_fo.reserve(nparse * 4))) return rc;
  if ((rc = h->d_rot.reserve(nparse * 8))) return rc;
  if ((rc = h->d_mod.reserve(nparse * 4))) return rc;
  if ((rc = h->d_tps.reserve(nparse * md.ntps * 8))) return rc;
  if ((rc = h->d_vote.reserve(nparse * 4))) return rc;
  if ((rc = h->d_osym.reserve(nparse * 4))) return rc;
  if ((rc = h->d_osrc.reserve(nparse * 4))) return rc;
  if ((rc = h->stg.h2d(h->d_X.p, in, nsym * md.N * 8, h->stream))) return rc;
  dvbt::DemodBuffers b{h->d_fo.as<int>(), h->d_rot.as<float2>(), h->d_mod.as<int>(), h->d_tps.as<float2>(), h->d_vote.as<int>(),
                       h->d_osym.as<int>(), h->d_osrc.as<int>()};
  rc = dvbt::demod_run(md, nullptr, h->d_X.as<float2>(), (int)nparse, b, h->d_state.as<dvbt::DemodState>(), h->fi_start, sync0 ? 1 : 0,
                       h->d_Y.as<float2>(), nullptr, h->stream);
  if (rc) return rc;
  DVBT_CUDA_TRY(cudaMemcpyAsync(h->h_state.p, h->d_state.p, sizeof(dvbt::DemodState), cudaMemcpyDeviceToHost, h->stream));
  DVBT_CUDA_TRY(cudaStreamSynchronize(h->stream));
  const dvbt::DemodState *S = h->h_state.as<dvbt::DemodState>();
  size_t nout = (size_t)S->n_out;
  std::vector<int> symidx(nout);
  if (nout) {
    demod_compact_kernel<<<(unsigned)nout, 256, 0, h->stream>>>(md.P, h->d_osrc.as<int>(), h->d_Y.as<float2>(), h->d_Yc.as<float2>());
    dvbt::count_launch();
    DVBT_CUDA_TRY(cudaGetLastError());
    DVBT_CUDA_TRY(cudaMemcpyAsync(symidx.data(), h->d_osym.p, nout * 4, cudaMemcpyDeviceToHost, h->stream));
    if ((rc = h->stg.d2h(out, h->d_Yc.p, nout * md.P * 8, h->stream))) return rc;
    DVBT_CUDA_TRY(cudaStreamSynchronize(h->stream));
  }
  // tags: superframe_start (0xaa) on the first output after sync, symbol_index on every output (:126-143)
  size_t nt = 0;
  if (tags_out && n_tags_out) {
    if (S->sf_tag_at >= 0 && nt < tags_out_capacity) tags_out[nt++] = dvbt_b200_tag{(uint64_t)S->sf_tag_at, DVBT_TAG_SUPERFRAME_START, 0xaa};
    for (size_t i = 0; i < nout && nt < tags_out_capacity; i++) tags_out[nt++] = dvbt_b200_tag{(uint64_t)i, DVBT_TAG_SYMBOL_INDEX, symidx[i]};
    *n_tags_out = nt;
  }
  *consumed = nparse;
  *produced = nout;
  return 0;
}

}  // extern "C"
