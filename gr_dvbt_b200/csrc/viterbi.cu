// K1 — chunk-parallel K=7 punctured Viterbi decoder for sm_100a (hard decisions as the reference; soft decisions as a
// separate mode beyond it).
//
// Replaces gr::dvbt::viterbi_decoder (lib/viterbi_decoder_impl.cc:191-324) and its SSE2
// kernels (lib/d_viterbi.c:461-576 d_viterbi_butterfly2_sse2, :680-735
// d_viterbi_get_output_sse2).  Bit-exact with the reference by construction:
//
//  * vit_depuncture_kernel turns the block's input (one constellation symbol per byte,
//    m bits, MSB first) into one 32-bit "step code" word per byte time (8 trellis steps,
//    4 bits per step: sym0, valid0, sym1, valid1), following the depuncture loop of
//    viterbi_decoder_impl.cc:241-256 and the puncture tables :61-65.
//  * vit_acs_kernel: ONE THREAD PER CHUNK of the stream, all 64 states in registers.  Default schedule (H16 = 1,
//    viterbi_acs_h16_gen.cuh, generated and simulated by gen_viterbi_acs_h16.py): a state is one halfword
//    metric << 8 | path byte, two states per register, and one VIADDMNMX.U16x2 per two states and step selects the
//    survivor's metric and path at once (the reference's tie rule falls out of the packing).  H16 = 2 (h16b): the same
//    steps with a cheaper event; the byte-SWAR and two-lane schedules of round 1 are a legacy build option
//    (DVBT_B200_LEGACY_ACS).  Every 8 steps (after the 6th step of a byte time, the reference's cadence :263-272) the path
//    bytes go to a per-thread survivor ring - the newest D rows in shared memory ([slot][word][thread], conflict free),
//    for ntraceback > 13 all rows also written through to a global ring - the best state is found with the reference's
//    tie rule (first strict maximum, d_viterbi.c:699-711) and the ring is traced back ntraceback-1 hops (:714-721).  The
//    trace stops early when it meets the previous byte's trace (identical result, fewer hops).  Metrics are
//    renormalised by a lower bound of the minimum (decisions depend on differences only; spread <= 12).
//  * Chunks start `warm` byte times early from the all-zero state.  Exactness does not rest
//    on that heuristic: every chunk records its metrics at its first byte time (G) and at
//    the next chunk's first byte time (F); vit_verify_kernel compares G[c] with F[c-1]
//    (min-normalised); two parallel rounds of vit_repair_round_kernel re-decode, from the true state, every flagged chunk
//    whose predecessor is settled, and vit_repair_kernel walks what is left sequentially (cascading while states differ).
//  * SOFT = true (dvbt_b200_viterbi_set_soft; the reference has no soft path, include/dvbt_b200.h): the same schedule with
//    8 bits per step in the step codes (two values in [-6, 6]) and a 4 x 256-entry addend table; spread bound 72.
#include "common.cuh"
// FMAADD (a template parameter of the functions that expand the generated macros) selects which additions
// of the schedule are forced onto the FMA pipe as IMAD with a run-time multiplier (vit_one / vit_neg1 /
// vit_two are kernel parameters so that ptxas cannot turn them back into ALU-pipe IADD3):
//   bit 0, bit 1: two / the other two of the four branch-metric additions per butterfly pair
//                 (ptxas already issues most plain additions as IMAD.IADD, so these change little);
//   bit 2:        the compare word and the path update - the only 3-input additions, which otherwise
//                 are IADD3 on the ALU pipe next to the LOP3/PRMT that cannot move.
#define VIT_ADDF(a, b) ((FMAADD & 1) ? vit_mad((a), vit_one, (b)) : (a) + (b))
#define VIT_ADDG(a, b) ((FMAADD & 2) ? vit_mad((a), vit_one, (b)) : (a) + (b))
#define VIT_CMP(m1, mj, bh, m0) ((FMAADD & 4) ? vit_mad((m0), vit_neg1, (mj) + (bh)) : (m1) + 0x80808080u - (m0))
#define VIT_PATH2(p) ((FMAADD & 4) ? vit_mad((p), vit_two, 0x01010101u) : (p) + (p) + 0x01010101u)
#include "viterbi_acs_gen.cuh"
#include "viterbi_acs2_gen.cuh"
#define VITH_ADDC(a, c) vit_mad((a), vit_one, (c))   // FMA pipe (ptxas picks VIADD otherwise)
#define VITH_HIMASK vit_himask   // 0xff00ff00 as a run-time register value (vit_neg1 ^ 0x00ff00ff)
#include "viterbi_acs_h16_gen.cuh"
#include "viterbi_acs_h16b_gen.cuh"

#include <stdlib.h>
#include <string.h>
#include <new>

namespace {

constexpr int kRowWords = 17;  // 16 path words + 1 trace-state word per ring slot
constexpr int kMaxStatRuns = 32;  // host slots for the verify/repair counters of the runs of one call
constexpr int kMaxNtb = 24;

__host__ __device__ constexpr int rate_k(int r) { return r == 0 ? 1 : r == 1 ? 2 : r == 2 ? 3 : r == 3 ? 5 : 7; }
__host__ __device__ constexpr int rate_n(int r) { return rate_k(r) + 1; }
// puncture masks, bit ph = 1 if X (resp. Y) of step phase ph is transmitted
// (viterbi_decoder_impl.cc:61-65, order X1 Y1 X2 Y2 ...)
__host__ __device__ constexpr unsigned rate_px(int r) { return r == 0 ? 0x1u : r == 1 ? 0x1u : r == 2 ? 0x5u : r == 3 ? 0x15u : 0x51u; }
__host__ __device__ constexpr unsigned rate_py(int r) { return r == 0 ? 0x1u : r == 1 ? 0x3u : r == 2 ? 0x3u : r == 3 ? 0x0bu : 0x2fu; }
static const int kNtb[5] = {5, 9, 10, 15, 24};  // viterbi_decoder_impl.cc:95-124

// ---------------------------------------------------------------------------------------
// depuncture: reference-format bytes -> step codes
// ---------------------------------------------------------------------------------------
template <int RATE, int MBITS>
__global__ void vit_depuncture_kernel(const uint8_t *__restrict__ in, long long in_stride,
                                      uint32_t *__restrict__ codes, long long codes_stride,
                                      long long nbt, int nstreams, long long code_offset) {
  constexpr int K = rate_k(RATE), N = rate_n(RATE);
  constexpr unsigned PX = rate_px(RATE), PY = rate_py(RATE);
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nbt * nstreams) return;
  int s = (int)(g / nbt);
  long long j = g - (long long)s * nbt;
  const uint8_t *src = in + (long long)s * in_stride;
  long long t = 8 * j;
  long long sp = t / K;
  int ph = (int)(t - sp * K);
  long long idx = sp * N + __popc(PX & ((1u << ph) - 1u)) + __popc(PY & ((1u << ph) - 1u));
  uint32_t w = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t nib = 0;
    if ((PX >> ph) & 1u) {
      long long b = idx / MBITS;
      int sh = MBITS - 1 - (int)(idx - b * MBITS);
      nib |= ((src[b] >> sh) & 1u) | 2u;
      idx++;
    }
    if ((PY >> ph) & 1u) {
      long long b = idx / MBITS;
      int sh = MBITS - 1 - (int)(idx - b * MBITS);
      nib |= (((src[b] >> sh) & 1u) << 2) | 8u;
      idx++;
    }
    w |= nib << (4 * i);
    ph = (ph + 1 == K) ? 0 : ph + 1;
  }
  codes[(long long)s * codes_stride + code_offset + j] = w;
}

// soft decisions: one signed value per transmitted code bit (order X1 Y1 X2 ..., viterbi_decoder_impl.cc:61-65), positive =
// "the bit is 1", clamped to +-6 -> two words of step codes per byte time (8 bits per step, punctured positions = 0)
template <int RATE>
__global__ void vit_depuncture_soft_kernel(const int8_t *__restrict__ in, uint2 *__restrict__ codes, long long nbt) {
  constexpr int K = rate_k(RATE), N = rate_n(RATE);
  constexpr unsigned PX = rate_px(RATE), PY = rate_py(RATE);
  long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nbt) return;
  long long t = 8 * j;
  long long sp = t / K;
  int ph = (int)(t - sp * K);
  long long idx = sp * N + __popc(PX & ((1u << ph) - 1u)) + __popc(PY & ((1u << ph) - 1u));
  uint32_t w[2] = {0u, 0u};
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int vx = 0, vy = 0;
    if ((PX >> ph) & 1u) vx = min(max((int)in[idx++], -6), 6);
    if ((PY >> ph) & 1u) vy = min(max((int)in[idx++], -6), 6);
    w[i >> 2] |= ((uint32_t)(vx + 8) | ((uint32_t)(vy + 8) << 4)) << (8 * (i & 3));
    ph = (ph + 1 == K) ? 0 : ph + 1;
  }
  codes[j] = make_uint2(w[0], w[1]);
}

// ---------------------------------------------------------------------------------------
// chunk geometry shared by the kernels
// ---------------------------------------------------------------------------------------
struct VitGeom {
  const uint32_t *codes;   // step codes; local byte time j of stream s at codes[s*codes_stride + j]
  long long codes_stride;
  uint8_t *out;            // out index i (= byte time - ntb) of stream s at out[s*out_stride + i - O0]
  long long out_stride;
  uint32_t *G, *F;         // [nstreams*nchunks][gf_words] metrics at a chunk's first / end byte time
  const uint32_t *prevF;   // [nstreams][gf_words] true metrics at byte time O0 from the previous call, or null
  uint8_t *bad;            // [nstreams*nchunks]
  unsigned int *counters;  // [0] = chunks flagged by verify, [1] = chunks re-decoded
  int nstreams, nchunks, L, W, ntb;
  int nbt;                 // byte times present in codes (local)
  int O0, O1;              // out indices this launch is responsible for: [O0, O1), O1 = nbt - ntb
  int reset_at_0;          // local byte time 0 is a decoder reset (true all-zero state)
  uint32_t neg1, two, one; // 0xffffffff, 2 and 1 as run-time values (see vit_mad in viterbi_acs_gen.cuh)
  int ring_depth;          // path rows kept in shared memory (== ntb: all of them, no global ring)
  uint32_t *gring;         // [blocks][ntb][16][block threads] path rows of all ntb slots when ring_depth < ntb
  int gf_words;            // 16: event-layout metric words (4 per word); 32: zipped halfword registers (h16b schedule)
};

struct ChunkSpan {
  int A, B, jstart, jend, first_real;
  bool true_start;
};

__device__ __forceinline__ ChunkSpan chunk_span(const VitGeom &g, int c) {
  ChunkSpan s;
  s.A = g.O0 + c * g.L;
  s.B = min(s.A + g.L, g.O1);
  s.jstart = max(s.A - g.W, 0);
  s.true_start = (s.jstart == 0) && g.reset_at_0;
  s.jend = s.B + g.ntb;
  s.first_real = s.A + g.ntb;
  return s;
}

constexpr int kLutWords = 32 + 256;
constexpr int kLutWordsSoft = 32 + 4 * 256 * 4;   // soft decisions: Q[f][two 4-bit values] (16 KB)
constexpr int kSoftMax = 6;                        // |soft value| <= 6: spread 72 + 8 steps x 12 stays below 256 (see below)

// Step codes of one byte time.  Hard decisions: one word, 4 bits per step (sym0, valid0, sym1, valid1).  Soft decisions: two
// words, 8 bits per step: value of the X bit + 8 in the low nibble, of the Y bit + 8 in the high one; a value v in
// [-kSoftMax, kSoftMax] says "the bit is 1" with weight v (0 = punctured or erased; +-1 only = the hard decoder's metrics).
struct StepCodes { uint32_t x, y; };
template <bool SOFT>
__device__ __forceinline__ StepCodes load_codes(const uint32_t *__restrict__ codes, int j) {
  StepCodes c;
  if constexpr (SOFT) { const uint2 v = reinterpret_cast<const uint2 *>(codes)[j]; c.x = v.x; c.y = v.y; }
  else { c.x = codes[j]; c.y = 0u; }
  return c;
}

// addend words of step i (0..7) of the byte time with step codes `code` (halfword schedule)
template <bool SOFT>
__device__ __forceinline__ uint4 vith_addends(const uint32_t *lut, int i, StepCodes code) {
  if constexpr (SOFT)
    return reinterpret_cast<const uint4 *>(lut + 32)[VITH_STEP_F(i) * 256 + (((i < 4 ? code.x : code.y) >> (8 * (i & 3))) & 255u)];
  else
    return reinterpret_cast<const uint4 *>(lut + 32)[VITH_STEP_F(i) * 16 + ((code.x >> (4 * i)) & 15u)];
}

__device__ __forceinline__ uint32_t vit_signmask(uint32_t t) { return vit_prmt(t, 0u, 0xba98u); }

// Decode byte times [jstart, jend) of one stream in this thread.
//   Survivor storage, per thread (column `stride` apart in shared memory):
//     trace[slot * stride]                 slot = byte time mod ntb: state of the last traceback that passed
//                                          through the slot | path byte it read there << 8
//     pring[(slotd*16 + word) * stride]    path rows (64 path bytes = 16 words) of the newest D byte times,
//                                          slotd = byte time mod D
//     gring[(slot*16 + word) * gstride]    GRING only (D < ntb): path rows of all ntb slots in global
//                                          memory (written through every byte time, L2 resident); read only
//                                          by a traceback that has not merged within D byte times
//   The reference traces ntb-1 hops back from the best state every byte time (d_viterbi.c:714-721).  A
//   traceback that meets the previous one (same state in the same slot) follows it from there on, and the
//   path byte the reference would output was already read by the earlier traceback and sits in `trace`:
//   in the steady state a byte time costs one hop.  Shared memory holds 4*ntb + 64*D bytes per thread
//   instead of 68*ntb, which is what bounds the number of resident warps at rate 7/8 (ntb = 24).
//   RESUME: start from `init` = event-layout metrics valid after the event of byte time
//           jstart-1 (used by the repair kernel); otherwise from the all-zero state.
//   H16: 1 = the halfword schedule (viterbi_acs_h16_gen.cuh: metric << 8 | path per state, one VIADDMNMX.U16x2 per two
//        states and step) instead of the byte-SWAR one (0); its path bytes are bit reversed and its event layout differs,
//        everything else (ring, merge shortcut, G/F, renormalisation) is shared.
//        2 = the same steps with the cheaper event of viterbi_acs_h16b_gen.cuh: no metric words (the boundary vectors
//        G / F / init are the 32 zipped registers, not renormalised), renormalisation folded into step 7.
//   BD:  threads per block (= column stride of the ring arrays) as a compile-time constant, or 0 for the
//        run-time `stride`/`gstride`: with a constant the 32 row stores of a byte time need no address arithmetic.
//   SOFT: soft-decision step codes (H16 == 1 only): the same schedule with addends A[L] = w(c0, v0) + w(c1, v1),
//        w(1, v) = max(v, 0), w(0, v) = max(-v, 0) instead of the agreement counts; the renormalisation bound becomes
//        12 * kSoftMax (a state is 6 steps from any other, each worth at most 2 * kSoftMax).
template <bool RESUME, bool GRING, int FMAADD, int H16, int BD, bool SOFT = false>
__device__ __forceinline__ void vit_decode_range(const uint32_t *__restrict__ codes, int jstart,
                                                 int jend, int first_real, int ntb,
                                                 uint8_t *__restrict__ out, int O0, int saveA,
                                                 uint32_t *__restrict__ Gdst, int saveB,
                                                 uint32_t *__restrict__ Fdst,
                                                 const uint32_t *__restrict__ init,
                                                 uint32_t *trace, uint32_t *pring, int stride_rt, int D,
                                                 uint32_t *gring, int gstride_rt,
                                                 const uint32_t *__restrict__ lut, const uint32_t vit_neg1,
                                                 const uint32_t vit_two, const uint32_t vit_one) {
  const int stride = BD ? BD : stride_rt, gstride = BD ? BD : gstride_rt;
  const uint32_t vit_himask = vit_neg1 ^ 0x00ff00ffu;
  constexpr int GFW = H16 == 2 ? 32 : 16;   // words of a boundary metric vector
  constexpr uint32_t kSpread = SOFT ? 12u * kSoftMax : 12u;
  static_assert(!SOFT || H16 == 1, "soft decisions use the h16 schedule");
  uint32_t vm[16], vp[16];   // byte-SWAR state: metrics, paths (4 states per word)
  uint32_t vh[32];           // H16 state: metric << 8 | path (2 states per word)
#pragma unroll
  for (int i = 0; i < 16; i++) { vm[i] = 0u; vp[i] = 0u; }
#pragma unroll
  for (int i = 0; i < 32; i++) vh[i] = 0u;
  int j = jstart;
  if (RESUME) {
    // metrics are given just after the event of byte time jstart-1: finish that byte time
    const StepCodes code = load_codes<SOFT>(codes, jstart - 1);
    uint32_t vme[GFW];
#pragma unroll
    for (int i = 0; i < GFW; i++) vme[i] = init[i];
    if constexpr (H16 == 2) {
      uint4 q6 = vith_addends<false>(lut, 6, code), q7 = vith_addends<false>(lut, 7, code);
      const uint32_t sub = max((vme[0] >> 8) & 0xffu, 12u) - 12u;     // as below, from the metric of state 0
      const uint32_t neg2 = ((0u - sub) & 0xffu) * 0x01000100u;       // -sub in the metric byte of both halfwords
      const uint32_t bitsub = 0x00010001u - sub * 0x01000100u;        // decision bit of step 7 minus sub, 32-bit
      VITB_ACS_PART2(vme, vh, q6, q7, neg2, bitsub);
    } else if constexpr (H16 == 1) {
      uint4 q6 = vith_addends<SOFT>(lut, 6, code), q7 = vith_addends<SOFT>(lut, 7, code);
      VITH_ACS_PART2(vme, vh, q6, q7);
    } else {
      uint32_t a6 = lut[(code.x >> 24) & 15u], a7 = lut[(code.x >> 28) & 15u];
      VIT_ACS_PART2(vme, vp, vm, vp, a6, a7);
    }
  }
  int slot = j % ntb;
  int slotd = GRING ? j % D : slot;
  bool have_trace = false;
  StepCodes code = (j < jend) ? load_codes<SOFT>(codes, j) : StepCodes{0u, 0u};
  // The byte times are walked in three segments that end at saveA, saveB and jend: the boundary metrics G / F are
  // stored between the segments, not by 32 predicated stores inside the unrolled byte-time body.
  uint32_t vme[GFW];  // metrics after the event (renormalised; H16 == 2: the zipped registers, not renormalised);
                      // live across iterations for the stores below
#pragma unroll 1
  for (int seg = 0; seg < 3; ++seg) {
  const int jstop = seg == 0 ? min(saveA + 1, jend) : seg == 1 ? min(saveB + 1, jend) : jend;
  const bool ran = j < jstop;
#pragma unroll 1
  for (; j < jstop; ++j) {
    const StepCodes next_code = (j + 1 < jend) ? load_codes<SOFT>(codes, j + 1) : StepCodes{0u, 0u};
    uint32_t vpe[16], vhe[32];
    if constexpr (H16 == 2) {
      uint4 q0 = vith_addends<false>(lut, 0, code), q1 = vith_addends<false>(lut, 1, code), q2 = vith_addends<false>(lut, 2, code),
            q3 = vith_addends<false>(lut, 3, code), q4 = vith_addends<false>(lut, 4, code), q5 = vith_addends<false>(lut, 5, code);
      VITB_ACS_PART1(vh, vme, vpe, q0, q1, q2, q3, q4, q5);
    } else if constexpr (H16 == 1) {
      uint4 q0 = vith_addends<SOFT>(lut, 0, code), q1 = vith_addends<SOFT>(lut, 1, code), q2 = vith_addends<SOFT>(lut, 2, code),
            q3 = vith_addends<SOFT>(lut, 3, code), q4 = vith_addends<SOFT>(lut, 4, code), q5 = vith_addends<SOFT>(lut, 5, code);
      VITH_ACS_PART1(vh, vme, vpe, vhe, q0, q1, q2, q3, q4, q5);
    } else {
      uint32_t a0 = lut[code.x & 15u], a1 = lut[(code.x >> 4) & 15u], a2 = lut[(code.x >> 8) & 15u],
               a3 = lut[(code.x >> 12) & 15u], a4 = lut[(code.x >> 16) & 15u], a5 = lut[(code.x >> 20) & 15u];
      VIT_ACS_PART1(vm, vp, vme, vpe, a0, a1, a2, a3, a4, a5);
    }

    // ---------------- event of byte time j (d_viterbi_get_output_sse2, d_viterbi.c:680-735)
    {
      uint32_t *row = pring + slotd * 16 * stride;
#pragma unroll
      for (int w = 0; w < 16; w++) row[w * stride] = vpe[w];  // ppresult[store_pos] (:692-696)
      if (GRING) {
        uint32_t *grow = gring + (long long)slot * 16 * gstride;
#pragma unroll
        for (int w = 0; w < 16; w++) grow[(long long)w * gstride] = vpe[w];
      }
    }

    if (j >= first_real) {
      // best state: lowest index with strictly greatest metric (:699-711).  Lane-wise
      // tournament over the 16 words (for a fixed lane the state index grows with the word
      // index), ties keep the lower word; then the 4 lanes are ranked on (metric, -state).
      uint32_t s;
      if constexpr (H16 == 2) {
        uint32_t bw;
        VITB_ARGMAX(vme, bw);
        s = 63u - (max(bw & 0xffffu, bw >> 16) & 63u);
      } else if constexpr (H16 == 1) {
        // 16-bit maximum of (metric << 8 | 63 - state) over all 64 halfwords (VIMNMX3.U16x2)
        uint32_t bw;
        VITH_ARGMAX(vhe, bw);
        s = 63u - (max(bw & 0xffffu, bw >> 16) & 63u);
      } else {
        uint32_t bv[16], bi[16];
#pragma unroll
        for (int w = 0; w < 16; w++) { bv[w] = vme[w]; bi[w] = 0x01010101u * (uint32_t)w; }
#pragma unroll
        for (int st = 1; st < 16; st *= 2) {
#pragma unroll
          for (int w = 0; w < 16; w += 2 * st) {
            uint32_t mk = vit_signmask(bv[w] + 0x80808080u - bv[w + st]);  // 0xff: left >= right
            bv[w] = vit_sel(bv[w + st], bv[w], mk);
            bi[w] = vit_sel(bi[w + st], bi[w], mk);
          }
        }
        uint32_t best = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) {
          uint32_t mb = (bv[0] >> (8 * b)) & 0xffu, wb = (bi[0] >> (8 * b)) & 0xffu;
          uint32_t state = ((wb & 8u) << 2) | ((uint32_t)b << 3) | (wb & 7u);
          best = max(best, (mb << 6) | (63u - state));
        }
        s = 63u - (best & 63u);
      }

      // traceback ntb-1 hops (:714-721) with early exit on meeting the previous trace
      bool merged = false;
      int q = slot, qd = slotd;
      for (int h = 0; h < ntb - 1; h++) {
        uint32_t *tq = trace + q * stride;
        if (h > 0 && have_trace && (*tq & 0xffu) == s) { merged = true; break; }
        uint32_t bidx = H16 ? vith_event_byte_index(s) : vit_event_byte_index(s);
        uint32_t pb;
        if (!GRING || h < D) pb = ((const uint8_t *)(pring + (qd * 16 + (int)(bidx >> 2)) * stride))[bidx & 3u];
        else pb = (gring[((long long)q * 16 + (bidx >> 2)) * gstride] >> (8u * (bidx & 3u))) & 0xffu;
        if (H16) pb = __brev(pb) >> 24;   // the halfword schedule fills the path byte from bit 0 upwards
        *tq = s | (pb << 8);
        s = pb >> 2;
        q = (q == 0) ? ntb - 1 : q - 1;
        qd = (qd == 0) ? D - 1 : qd - 1;
      }
      int qf = (slot + 1 == ntb) ? 0 : slot + 1;  // slot of byte time j-(ntb-1)
      uint32_t *tf = trace + qf * stride;
      uint32_t ob;
      if (merged) {
        ob = *tf >> 8;
      } else {
        uint32_t bidx = H16 ? vith_event_byte_index(s) : vit_event_byte_index(s);
        if (GRING) ob = (gring[((long long)qf * 16 + (bidx >> 2)) * gstride] >> (8u * (bidx & 3u))) & 0xffu;
        else ob = ((const uint8_t *)(pring + (qf * 16 + (int)(bidx >> 2)) * stride))[bidx & 3u];
        if (H16) ob = __brev(ob) >> 24;
        *tf = s | (ob << 8);
      }
      have_trace = true;
      out[j - ntb - O0] = (uint8_t)ob;  // :724
    }

    // renormalise (:728-732 subtracts the exact minimum; any lower bound of it gives the same
    // decisions).  All metrics >= metric(state 0) - 12 (soft decisions: - 12 kSoftMax).
    if constexpr (H16 != 2) {
      uint32_t x = vme[0] & 0xffu;
      uint32_t sub = (max(x, kSpread) - kSpread) * 0x01010101u;
#pragma unroll
      for (int w = 0; w < 16; w++) vme[w] -= sub;
    }
    if constexpr (H16 == 2) {
      // the subtraction rides on the addends of step 7 (8 words instead of 32 registers); state 0 is Z[0].lo
      uint4 q6 = vith_addends<false>(lut, 6, code), q7 = vith_addends<false>(lut, 7, code);
      const uint32_t sub = max((vme[0] >> 8) & 0xffu, 12u) - 12u;     // as below, from the metric of state 0
      const uint32_t neg2 = ((0u - sub) & 0xffu) * 0x01000100u;       // -sub in the metric byte of both halfwords
      const uint32_t bitsub = 0x00010001u - sub * 0x01000100u;        // decision bit of step 7 minus sub, 32-bit
      VITB_ACS_PART2(vme, vh, q6, q7, neg2, bitsub);
    } else if constexpr (H16 == 1) {
      uint4 q6 = vith_addends<SOFT>(lut, 6, code), q7 = vith_addends<SOFT>(lut, 7, code);
      VITH_ACS_PART2(vme, vh, q6, q7);           // paths := 0 (:730) is part of the zip
    } else {
      uint32_t a6 = lut[(code.x >> 24) & 15u], a7 = lut[(code.x >> 28) & 15u];
#pragma unroll
      for (int w = 0; w < 16; w++) vpe[w] = 0u;  // paths := 0 (:730)
      VIT_ACS_PART2(vme, vpe, vm, vp, a6, a7);
    }
    slot = (slot + 1 == ntb) ? 0 : slot + 1;
    if (GRING) slotd = (slotd + 1 == D) ? 0 : slotd + 1; else slotd = slot;
    code = next_code;
  }
  uint32_t *sv = seg == 0 ? Gdst : seg == 1 ? Fdst : nullptr;   // the last byte time walked was saveA / saveB
  if (ran && sv) {
#pragma unroll
    for (int w = 0; w < GFW; w++) sv[w] = vme[w];
  }
  }
}

// Shared-memory tables of a block, kLutWords words: [0,16) byte-SWAR agreement bytes per step code nibble;
// [32, 32+256) halfword schedule: uint4 Q[f][nib] = addend words A[L] << 8 | A[L ^ f] << 24, L = 0..3.
template <bool SOFT = false>
__device__ __forceinline__ void vit_fill_lut(uint32_t *lut) {
  if constexpr (SOFT) {
    // uint4 Q[f][vy + 8 << 4 | vx + 8]: A[L] = w(c0, vx) + w(c1, vy), L = 2 c0 + c1 (X is the first code bit of a step)
    for (uint32_t t = threadIdx.x; t < 1024u; t += blockDim.x) {
      const uint32_t f = t >> 8;
      const int vx = min(max((int)(t & 15u) - 8, -kSoftMax), kSoftMax), vy = min(max((int)((t >> 4) & 15u) - 8, -kSoftMax), kSoftMax);
      uint32_t A[4];
      for (int L = 0; L < 4; L++) A[L] = (uint32_t)(((L & 2) ? max(vx, 0) : max(-vx, 0)) + ((L & 1) ? max(vy, 0) : max(-vy, 0)));
      for (uint32_t L = 0; L < 4u; L++) lut[32 + 4 * t + L] = (A[L] << 8) | (A[L ^ f] << 24);
    }
    return;
  }
  for (uint32_t t = threadIdx.x; t < 64u; t += blockDim.x) {
    uint32_t nib = t & 15u, f = t >> 4;
    uint32_t s0 = nib & 1u, v0 = (nib >> 1) & 1u, s1 = (nib >> 2) & 1u, v1 = (nib >> 3) & 1u;
    // byte L = 2*c0 + c1 holds A[L] = v0*[c0==s0] + v1*[c1==s1]
    uint32_t apk = v0 * (s0 ? 0x01010000u : 0x00000101u) + v1 * (s1 ? 0x01000100u : 0x00010001u);
    if (f == 0) lut[nib] = apk;
    for (uint32_t L = 0; L < 4u; L++)
      lut[32 + 4 * t + L] = (((apk >> (8 * L)) & 0xffu) << 8) | (((apk >> (8 * (L ^ f))) & 0xffu) << 24);
  }
}

template <bool GRING, int FMAADD, int H16, int BD, bool SOFT = false>
__global__ void __launch_bounds__(512, 1) vit_acs_kernel(VitGeom g) {
  extern __shared__ uint32_t smem[];
  uint32_t *lut = smem;
  uint32_t *trace = smem + (SOFT ? kLutWordsSoft : kLutWords);   // [ntb][threads]
  uint32_t *pring = trace + g.ntb * blockDim.x;         // [ring_depth][16][threads]
  vit_fill_lut<SOFT>(lut);
  __syncthreads();
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)g.nstreams * g.nchunks) return;
  int s = (int)(gid / g.nchunks);
  int c = (int)(gid - (long long)s * g.nchunks);
  ChunkSpan sp = chunk_span(g, c);
  if (sp.A >= g.O1) return;
  uint32_t *gr = GRING ? g.gring + (long long)blockIdx.x * g.ntb * 16 * blockDim.x + threadIdx.x : nullptr;
  vit_decode_range<false, GRING, FMAADD, H16, BD, SOFT>(g.codes + s * g.codes_stride * (SOFT ? 2 : 1), sp.jstart, sp.jend, sp.first_real, g.ntb,
                                 g.out + s * g.out_stride, g.O0, sp.true_start ? -1 : sp.A,
                                 g.G + gid * (H16 == 2 ? 32 : 16), sp.B, g.F + gid * (H16 == 2 ? 32 : 16), nullptr, trace + threadIdx.x, pring + threadIdx.x,
                                 (int)blockDim.x, g.ring_depth, gr, (int)blockDim.x, lut, g.neg1, g.two, g.one);
}

// ---------------------------------------------------------------------------------------
// two lanes per chunk (viterbi_acs2_gen.cuh): lane t of a pair holds the states with s5 = t at the event.
// Same ring row format, same G/F format, same results as the one-lane kernel; twice the threads per SM.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void vit_decode_range2(const uint32_t *__restrict__ codes, int jstart, int jend, int first_real, int ntb,
                                                  uint8_t *__restrict__ out, int O0, int saveA, uint32_t *__restrict__ Gdst, int saveB,
                                                  uint32_t *__restrict__ Fdst, uint32_t *ring, int stride,
                                                  const uint32_t *__restrict__ lut) {
  const int lane = threadIdx.x & 31;
  const bool vit_t = (lane & 1) != 0;
  const unsigned vit_pairmask = 3u << (lane & ~1);
  const uint32_t vit_tsel1 = vit_t ? 0x2301u : 0x3210u, vit_tsel2 = vit_t ? 0x1032u : 0x3210u, vit_tsel3 = vit_t ? 0x0123u : 0x3210u;
  const int tw = vit_t ? 8 : 0;  // my first word in a ring row / G / F vector
  uint32_t vm[8], vp[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { vm[i] = 0u; vp[i] = 0u; }
  int j = jstart;
  int slot = j % ntb;
  bool have_trace = false;
  uint32_t code = (j < jend) ? codes[j] : 0u;
  for (; j < jend; ++j) {
    uint32_t next_code = (j + 1 < jend) ? codes[j + 1] : 0u;
    uint32_t a0 = lut[code & 15u], a1 = lut[(code >> 4) & 15u], a2 = lut[(code >> 8) & 15u],
             a3 = lut[(code >> 12) & 15u], a4 = lut[(code >> 16) & 15u], a5 = lut[(code >> 20) & 15u],
             a6 = lut[(code >> 24) & 15u], a7 = lut[(code >> 28) & 15u];
    uint32_t vme[8], vpe[8];
    VIT2_ACS_PART1(vm, vp, vme, vpe, a0, a1, a2, a3, a4, a5);

    uint32_t *row = ring + slot * kRowWords * stride;
#pragma unroll
    for (int w = 0; w < 8; w++) row[(tw + w) * stride] = vpe[w];
    __syncwarp(vit_pairmask);

    if (j >= first_real) {
      uint32_t bv[8], bi[8];
#pragma unroll
      for (int w = 0; w < 8; w++) { bv[w] = vme[w]; bi[w] = 0x01010101u * (uint32_t)w; }
#pragma unroll
      for (int st = 1; st < 8; st *= 2) {
#pragma unroll
        for (int w = 0; w < 8; w += 2 * st) {
          uint32_t mk = vit_signmask(bv[w] + 0x80808080u - bv[w + st]);
          bv[w] = vit_sel(bv[w + st], bv[w], mk);
          bi[w] = vit_sel(bi[w + st], bi[w], mk);
        }
      }
      uint32_t best = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) {
        uint32_t mb = (bv[0] >> (8 * b)) & 0xffu, wb = (bi[0] >> (8 * b)) & 0xffu;
        uint32_t state = (vit_t ? 32u : 0u) | ((uint32_t)b << 3) | wb;
        best = max(best, (mb << 6) | (63u - state));
      }
      best = max(best, __shfl_xor_sync(vit_pairmask, best, 1));
      uint32_t s = 63u - (best & 63u);

      bool merged = false;
      int q = slot;
      for (int h = 0; h < ntb - 1; h++) {
        uint32_t *qrow = ring + q * kRowWords * stride;
        if (h > 0 && have_trace && qrow[16 * stride] == s) { merged = true; break; }
        __syncwarp(vit_pairmask);  // both lanes have read the old trace entry before either overwrites it
        qrow[16 * stride] = s;
        uint32_t bidx = vit_event_byte_index(s);
        s = (uint32_t)((const uint8_t *)(qrow + (bidx >> 2) * stride))[bidx & 3u] >> 2;
        q = (q == 0) ? ntb - 1 : q - 1;
      }
      int qf = (slot + 1 == ntb) ? 0 : slot + 1;
      uint32_t *frow = ring + qf * kRowWords * stride;
      if (merged) s = frow[16 * stride];
      __syncwarp(vit_pairmask);
      if (!merged) frow[16 * stride] = s;
      have_trace = true;
      uint32_t bidx = vit_event_byte_index(s);
      if (!vit_t) out[j - ntb - O0] = ((const uint8_t *)(frow + (bidx >> 2) * stride))[bidx & 3u];
    }
    {
      uint32_t x = __shfl_sync(vit_pairmask, vme[0] & 0xffu, lane & ~1);  // metric of state 0 lives in lane 0 of the pair
      uint32_t sub = (max(x, 12u) - 12u) * 0x01010101u;
#pragma unroll
      for (int w = 0; w < 8; w++) vme[w] -= sub;
    }
    if (j == saveA && Gdst) {
#pragma unroll
      for (int w = 0; w < 8; w++) Gdst[tw + w] = vme[w];
    }
    if (j == saveB && Fdst) {
#pragma unroll
      for (int w = 0; w < 8; w++) Fdst[tw + w] = vme[w];
    }
#pragma unroll
    for (int w = 0; w < 8; w++) vpe[w] = 0u;
    __syncwarp(vit_pairmask);  // the row written above is read by both lanes during the traceback
    VIT2_ACS_PART2(vme, vpe, vm, vp, a6, a7);
    slot = (slot + 1 == ntb) ? 0 : slot + 1;
    code = next_code;
  }
}

__global__ void __launch_bounds__(512, 1) vit_acs2_kernel(VitGeom g) {
  extern __shared__ uint32_t smem[];
  uint32_t *lut = smem;
  uint32_t *ring = smem + kLutWords;
  vit_fill_lut(lut);
  __syncthreads();
  const int npairs = blockDim.x >> 1;
  long long gid = (long long)blockIdx.x * npairs + (threadIdx.x >> 1);
  if (gid >= (long long)g.nstreams * g.nchunks) return;
  int s = (int)(gid / g.nchunks);
  int c = (int)(gid - (long long)s * g.nchunks);
  ChunkSpan sp = chunk_span(g, c);
  if (sp.A >= g.O1) return;
  vit_decode_range2(g.codes + s * g.codes_stride, sp.jstart, sp.jend, sp.first_real, g.ntb, g.out + s * g.out_stride, g.O0,
                    sp.true_start ? -1 : sp.A, g.G + gid * 16, sp.B, g.F + gid * 16, ring + (threadIdx.x >> 1), npairs, lut);
}

// min-normalised comparison of two 64-metric vectors stored as 16 SWAR words
__device__ __forceinline__ bool vit_same_state(const uint32_t *a, const uint32_t *b) {
  uint32_t mina = 255u, minb = 255u;
  for (int w = 0; w < 16; w++)
    for (int l = 0; l < 4; l++) {
      mina = min(mina, (a[w] >> (8 * l)) & 0xffu);
      minb = min(minb, (b[w] >> (8 * l)) & 0xffu);
    }
  for (int w = 0; w < 16; w++)
    if (a[w] - mina * 0x01010101u != b[w] - minb * 0x01010101u) return false;
  return true;
}

// the same for two vectors of 32 zipped halfword registers (h16b schedule: metric << 8 per halfword, path bytes empty)
__device__ __forceinline__ bool vit_same_state_z(const uint32_t *a, const uint32_t *b) {
  uint32_t mina = 255u, minb = 255u;
  for (int w = 0; w < 32; w++) {
    mina = min(mina, min((a[w] >> 8) & 0xffu, a[w] >> 24));
    minb = min(minb, min((b[w] >> 8) & 0xffu, b[w] >> 24));
  }
  for (int w = 0; w < 32; w++)
    if (a[w] - mina * 0x01000100u != b[w] - minb * 0x01000100u) return false;
  return true;
}

template <int GFW>
__device__ __forceinline__ bool vit_same_state_any(const uint32_t *a, const uint32_t *b) {
  if constexpr (GFW == 32) return vit_same_state_z(a, b);
  else return vit_same_state(a, b);
}

template <int GFW>
__global__ void vit_verify_kernel(VitGeom g) {
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)g.nstreams * g.nchunks) return;
  int s = (int)(gid / g.nchunks);
  int c = (int)(gid - (long long)s * g.nchunks);
  ChunkSpan sp = chunk_span(g, c);
  uint8_t bad = 0;
  if (sp.A < g.O1 && !sp.true_start) {
    const uint32_t *ref = (c > 0) ? g.F + (gid - 1) * GFW : (g.prevF ? g.prevF + s * GFW : nullptr);
    if (ref && !vit_same_state_any<GFW>(g.G + gid * GFW, ref)) bad = 1;
  }
  g.bad[gid] = bad;
  if (bad) atomicAdd(&g.counters[0], 1u);
}

// Repair of the chunks whose warm-up had not converged (verify found G[c] != F[c-1]).
//
// vit_repair_round_kernel - one THREAD PER CHUNK, all bad chunks of a round re-decoded in parallel.  Chunk c can be
// re-decoded exactly when the true state at its start is known, i.e. when its predecessor needs nothing in this round:
//     need(c)  = bad_in[c]  ||  chg_in[c-1]          (chg: the predecessor's end state F changed in the previous round)
//     ready(c) = need(c) && !need(c-1)
// A ready chunk whose only reason is chg_in[c-1] first compares G[c] with the new F[c-1] (often equal again: nothing to
// do).  Every thread writes its own bad_out[c] / chg_out[c], so the rounds need no atomics and no clearing; a chunk that
// was not ready carries its flag to the next round.  Isolated bad chunks - the usual case under noise - are all done in
// the first round (one chunk's decode time instead of one per bad chunk); runs of consecutive bad chunks take one round
// per chunk of the run, and whatever is left after the rounds goes to the sequential kernel below.  Exactness never
// rests on the number of rounds.
template <int H16, bool SOFT = false>
__global__ void __launch_bounds__(32, 1) vit_repair_round_kernel(VitGeom g, const uint8_t *__restrict__ bad_in, const uint8_t *__restrict__ chg_in,
                                                                 uint8_t *__restrict__ bad_out, uint8_t *__restrict__ chg_out, int left_counter) {
  constexpr int GFW = H16 == 2 ? 32 : 16;
  extern __shared__ uint32_t smem[];
  if (g.counters[0] == 0u) return;                       // verify found nothing (the normal case): no round touches anything
  uint32_t *lut = smem;
  uint32_t *trace = smem + (SOFT ? kLutWordsSoft : kLutWords);
  uint32_t *pring = trace + g.ntb * blockDim.x;
  vit_fill_lut<SOFT>(lut);
  __syncthreads();
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)g.nstreams * g.nchunks) return;
  const int s = (int)(gid / g.nchunks);
  const int c = (int)(gid - (long long)s * g.nchunks);
  const ChunkSpan sp = chunk_span(g, c);
  auto need = [&](int cc) -> bool {
    if (cc < 0) return false;
    const long long q = (long long)s * g.nchunks + cc;
    return bad_in[q] != 0 || (cc > 0 && chg_in && chg_in[q - 1] != 0);
  };
  uint8_t nb = 0, nc = 0;
  const uint32_t *ref = (c > 0) ? g.F + (gid - 1) * GFW : (g.prevF ? g.prevF + s * GFW : nullptr);
  if (sp.A < g.O1 && !sp.true_start && ref && need(c)) {
    if (need(c - 1)) {
      nb = 1;                                            // the predecessor's end state is not final yet: next round
    } else {
      bool redo = bad_in[gid] != 0 || !vit_same_state_any<GFW>(g.G + gid * GFW, ref);
      if (redo) {
        uint32_t oldF[GFW], init[GFW];
        for (int w = 0; w < GFW; w++) { oldF[w] = g.F[gid * GFW + w]; init[w] = ref[w]; }
        vit_decode_range<true, false, 0, H16, 0, SOFT>(g.codes + s * g.codes_stride * (SOFT ? 2 : 1), sp.A + 1, sp.jend, sp.first_real, g.ntb,
                                      g.out + s * g.out_stride, g.O0, -1, nullptr, sp.B, g.F + gid * GFW, init,
                                      trace + threadIdx.x, pring + threadIdx.x, (int)blockDim.x, g.ntb, nullptr, 0, lut, g.neg1, g.two, g.one);
        atomicAdd(&g.counters[1], 1u);
        nc = vit_same_state_any<GFW>(oldF, g.F + gid * GFW) ? 0 : 1;
      }
    }
  }
  bad_out[gid] = nb;
  chg_out[gid] = nc;
  if (nb | nc) atomicAdd(&g.counters[left_counter], 1u);   // something is left for the next round / the sequential kernel
}

// One thread per stream walks its chunks in order and re-decodes, from the true state of the previous chunk, every chunk
// that the parallel rounds left flagged (bad: still to do; chg of the predecessor: to be re-checked), cascading while end
// states change.  Does nothing when verify found no mismatch (the normal case).
template <int H16, bool SOFT = false>
__global__ void __launch_bounds__(32, 1) vit_repair_kernel(VitGeom g, const uint8_t *__restrict__ bad_in, const uint8_t *__restrict__ chg_in, int left_counter) {
  constexpr int GFW = H16 == 2 ? 32 : 16;
  extern __shared__ uint32_t smem[];
  uint32_t *lut = smem;
  uint32_t *trace = smem + (SOFT ? kLutWordsSoft : kLutWords);   // the repair kernel keeps the whole ring in shared memory
  uint32_t *pring = trace + g.ntb * blockDim.x;
  vit_fill_lut<SOFT>(lut);
  __syncthreads();
  if (g.counters[0] == 0u) return;
  if (left_counter >= 0 && g.counters[left_counter] == 0u) return;   // the parallel rounds settled everything
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= g.nstreams) return;
  bool recheck = false;
  for (int c = 0; c < g.nchunks; c++) {
    long long gid = (long long)s * g.nchunks + c;
    // skip clean stretches 16 chunks at a time (one load each instead of a dependent global round trip per chunk: with
    // tens of thousands of chunks that walk alone took 10 ms)
    if (!recheck && c + 16 <= g.nchunks && (gid & 15) == 0) {
      const uint4 b = *reinterpret_cast<const uint4 *>(bad_in + gid);
      uint4 q = make_uint4(0u, 0u, 0u, 0u);
      uint32_t prev = 0u;
      if (chg_in) { q = *reinterpret_cast<const uint4 *>(chg_in + gid); prev = c > 0 ? chg_in[gid - 1] : 0u; }
      if ((b.x | b.y | b.z | b.w | q.x | q.y | q.z | q.w | prev) == 0u) { c += 15; continue; }
    }
    ChunkSpan sp = chunk_span(g, c);
    if (sp.A >= g.O1) break;
    if (sp.true_start) { recheck = false; continue; }
    const uint32_t *ref = (c > 0) ? g.F + (gid - 1) * GFW : (g.prevF ? g.prevF + s * GFW : nullptr);
    if (!ref) { recheck = false; continue; }
    bool redo = bad_in[gid] != 0;
    if (!redo && (recheck || (c > 0 && chg_in && chg_in[gid - 1] != 0))) redo = !vit_same_state_any<GFW>(g.G + gid * GFW, ref);
    if (!redo) { recheck = false; continue; }
    uint32_t oldF[GFW], init[GFW];
    for (int w = 0; w < GFW; w++) { oldF[w] = g.F[gid * GFW + w]; init[w] = ref[w]; }
    // resume right after the event of byte time A (the state `ref` describes)
    vit_decode_range<true, false, 0, H16, 0, SOFT>(g.codes + s * g.codes_stride * (SOFT ? 2 : 1), sp.A + 1, sp.jend, sp.first_real, g.ntb,
                                  g.out + s * g.out_stride, g.O0, -1, nullptr, sp.B, g.F + gid * GFW, init,
                                  trace + threadIdx.x, pring + threadIdx.x, (int)blockDim.x, g.ntb, nullptr, 0, lut, g.neg1, g.two, g.one);
    atomicAdd(&g.counters[1], 1u);
    recheck = !vit_same_state_any<GFW>(oldF, g.F + gid * GFW);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
struct dvbt_b200_viterbi {
  int device = dvbt::current_device();
  dvbt_b200_viterbi_params par;
  int k, n, m, ntb;
  int nsymbols, nout;  // input bytes / output bytes per bsize block
  dvbt_b200_viterbi_tuning tune;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  dvbt::DevBuf d_in, d_codes, d_out, d_G, d_F, d_bad, d_counters, d_prevF, d_tmp, d_gring, h_stage_in, h_stage_out, h_counters;
  // streaming state (work()): history of step codes kept at the front of d_codes
  bool init_done = false;      // d_init of the reference
  long long hist_bt = 0;       // byte times of history present at the front of d_codes
  bool hist_from_reset = true; // history starts at the decoder reset
  bool have_prevF = false;
  // stats
  long long st_chunks = 0, st_repaired = 0;
  float st_ms = 0.f;
  dvbt::Staging stg;           // pinned staging of work()'s pageable buffers
  int st_runs = 0;             // run_decode() launches since the last finish_stats(): one 16-byte slot of h_counters each
  bool st_accumulate = false;  // the fused chain decodes several runs per call (one per stream segment)
  int sm_count = 148;
  int lanes = 1;  // lanes per chunk in the ACS kernel; DVBT_B200_VIT_LANES=2 selects the two-lane kernel (measured slower)
  bool h16 = true;  // halfword (VIADDMNMX.U16x2) ACS schedule; DVBT_B200_VIT_ACS=swar selects the byte-SWAR schedule
  bool soft = false; // soft-decision step codes (two words per byte time, set_soft()); h16 schedule only
  int cw() const { return soft ? 2 : 1; }                          // words per byte time in d_codes
  int lutw() const { return soft ? kLutWordsSoft : kLutWords; }    // table words at the front of a block's shared memory
  bool h16b = false; // DVBT_B200_VIT_ACS=h16b: halfword schedule with the cheaper event (viterbi_acs_h16b_gen.cuh);
                     // written without GPU access, opt-in until it has been measured and its parity tests have run
};

namespace {

using dvbt::set_error;

template <int RATE>
int launch_depuncture_m(int m, const uint8_t *in, long long in_stride, uint32_t *codes,
                        long long codes_stride, long long nbt, int nstreams, long long code_offset,
                        cudaStream_t st) {
  long long total = nbt * nstreams;
  if (total == 0) return 0;
  int bd = 256;
  unsigned grid = (unsigned)((total + bd - 1) / bd);
  if (m == 2) vit_depuncture_kernel<RATE, 2><<<grid, bd, 0, st>>>(in, in_stride, codes, codes_stride, nbt, nstreams, code_offset);
  else if (m == 4) vit_depuncture_kernel<RATE, 4><<<grid, bd, 0, st>>>(in, in_stride, codes, codes_stride, nbt, nstreams, code_offset);
  else vit_depuncture_kernel<RATE, 6><<<grid, bd, 0, st>>>(in, in_stride, codes, codes_stride, nbt, nstreams, code_offset);
  dvbt::count_launch();
  DVBT_CUDA_TRY(cudaGetLastError());
  return 0;
}

int launch_depuncture(int rate, int m, const uint8_t *in, long long in_stride, uint32_t *codes,
                      long long codes_stride, long long nbt, int nstreams, long long code_offset,
                      cudaStream_t st) {
  switch (rate) {
    case 0: return launch_depuncture_m<0>(m, in, in_stride, codes, codes_stride, nbt, nstreams, code_offset, st);
    case 1: return launch_depuncture_m<1>(m, in, in_stride, codes, codes_stride, nbt, nstreams, code_offset, st);
    case 2: return launch_depuncture_m<2>(m, in, in_stride, codes, codes_stride, nbt, nstreams, code_offset, st);
    case 3: return launch_depuncture_m<3>(m, in, in_stride, codes, codes_stride, nbt, nstreams, code_offset, st);
    default: return launch_depuncture_m<4>(m, in, in_stride, codes, codes_stride, nbt, nstreams, code_offset, st);
  }
}

size_t acs_smem_bytes(int ntb, int bd, int lutw = kLutWords) { return (size_t)(lutw + ntb * kRowWords * bd) * 4; }

// returns the number of chunks (ring columns) per block
int pick_block_dim(const dvbt_b200_viterbi *h) {
  if (h->tune.threads_per_block > 0) return h->lanes == 2 ? h->tune.threads_per_block / 2 : h->tune.threads_per_block;
  const size_t budget = 227 * 1024;
  int bd = h->lanes == 2 ? 256 : 512;
  while (bd > 32 && acs_smem_bytes(h->ntb, bd) > budget) bd -= 32;
  // two blocks per SM when two fit (more warps, same shared memory in total)
  if (h->lanes == 2 && acs_smem_bytes(h->ntb, bd) * 2 > budget && bd > 32) {
    int half = (bd / 2) / 16 * 16;
    if (half >= 32 && bd >= 128) { /* keep one big block: ring columns are the scarce resource */ }
  }
  return bd;
}

// Launch geometry of the one-lane ACS kernel: threads per block, path rows kept in shared memory, blocks
// that fit one SM.  The whole ring stays in shared memory when that leaves >= 256 threads per SM
// (ntb <= 13: rates 1/2 .. 3/4).  Otherwise (5/6, 7/8: ntb 15, 24) the older rows live in the global ring
// and shared memory is spent on resident warps instead: the kernel is ALU/issue bound and one warp per
// scheduler cannot hide its own dependent-issue latency (profiles/: 65 % ALU pipe at 4 warps/SM).
struct AcsPlan {
  int bd, depth, blocks_per_sm;
  size_t smem;
  bool gring;
};

size_t acs_ring_smem(int ntb, int depth, int bd, int lutw = kLutWords) { return (size_t)(lutw + (ntb + 16 * depth) * bd) * 4; }

AcsPlan plan_acs(const dvbt_b200_viterbi *h) {
  const size_t budget = 227 * 1024, reserved = 1024;  // per-block reservation of the runtime
  AcsPlan pl;
  const int ntb = h->ntb;
  if (h->lanes == 2) {
    pl.bd = pick_block_dim(h);
    pl.depth = ntb; pl.gring = false;
    pl.smem = acs_smem_bytes(ntb, pl.bd);
    pl.blocks_per_sm = pl.smem * 2 <= budget ? 2 : 1;
    return pl;
  }
  int depth = h->tune.ring_depth;
  if (const char *e = getenv("DVBT_B200_VIT_DEPTH")) depth = atoi(e);
  int tpsm = 384;                                     // resident threads per SM when the ring is split (tools/sweep_acs.py:
                                                      // 128/256/320/384/448/512 -> 1.28/1.16/1.27/1.13/1.28/1.28 ms)
  if (const char *e = getenv("DVBT_B200_VIT_TPSM")) tpsm = atoi(e);
  if (h->tune.threads_per_block > 0) {
    pl.bd = h->tune.threads_per_block;
    pl.depth = depth > 0 && depth < ntb ? depth : ntb;
    while (pl.depth > 1 && acs_ring_smem(ntb, pl.depth, pl.bd, h->lutw()) > budget) pl.depth--;
  } else {
    int full = 512;
    while (full > 32 && acs_ring_smem(ntb, ntb, full, h->lutw()) > budget) full -= 32;
    if (depth <= 0 && full >= 256) {
      pl.bd = full; pl.depth = ntb;
    } else {
      pl.bd = tpsm <= 512 ? tpsm : 128;               // one block per SM measured slightly better than 3 x 128
      if (const char *e = getenv("DVBT_B200_VIT_BD")) pl.bd = atoi(e);
      int bps = tpsm / pl.bd < 1 ? 1 : tpsm / pl.bd;
      long long words = (long long)(budget / bps - reserved) / 4 - h->lutw();   // per block
      int d = (int)((words / pl.bd - ntb) / 16);
      if (depth > 0) d = depth;
      pl.depth = d < 1 ? 1 : d >= ntb ? ntb : d;
    }
  }
  pl.gring = pl.depth < ntb;
  pl.smem = acs_ring_smem(ntb, pl.depth, pl.bd, h->lutw());
  int bps = (int)(budget / (pl.smem + reserved));
  if (bps < 1) bps = 1;
  if (bps * pl.bd > 1024) bps = 1024 / pl.bd;           // 128 registers per thread at most (launch bounds 512,1)
  pl.blocks_per_sm = bps < 1 ? 1 : bps;
  return pl;
}

// Decode local byte times [0, nbt) of nstreams streams whose step codes are in h->d_codes;
// responsible for out indices [O0, nbt-ntb).  Output goes to d_out (device).
int run_decode(dvbt_b200_viterbi *h, long long codes_stride, int nbt, int nstreams, int O0,
               bool reset_at_0, bool use_prevF, uint8_t *d_out, long long out_stride) {
  int O1 = nbt - h->ntb;
  if (!h->st_accumulate || h->st_runs == 0) {
    h->st_chunks = 0;
    h->st_repaired = 0;
    h->st_ms = 0.f;
    h->st_runs = 0;
  }
  if (O1 <= O0) return 0;
  const AcsPlan pl = plan_acs(h);
  const int bd = pl.bd;
  int W = h->tune.warmup_bytes > 0 ? h->tune.warmup_bytes : 40;
  int L = h->tune.chunk_bytes;
  if (L <= 0) {
    // enough chunks to fill every SM with the blocks it can hold, but not shorter than 4x the overhead
    long long want = (long long)h->sm_count * bd * pl.blocks_per_sm;
    // DVBT_B200_VIT_SM_DIV=d: a launch asks for 1/d of the machine with d times longer chunks.  With several captures in
    // flight (one stream each) the launches still fill the GPU together, and the fixed warm-up + traceback overlap of a
    // chunk (W + ntb byte times, 17 % at the default length) is spread over a longer chunk.  Experiment knob, default off.
    if (const char *e = getenv("DVBT_B200_VIT_SM_DIV")) {
      int d = atoi(e);
      if (d > 1) want = (want + d - 1) / d;
    }
    long long per_stream = (want + nstreams - 1) / nstreams;
    long long l = ((long long)(O1 - O0) + per_stream - 1) / per_stream;
    long long lmin = 4LL * (W + h->ntb);
    if (l < lmin) l = lmin;
    L = (int)l;
  }
  // Survivor slot = byte time mod ntb.  With the chunk length a multiple of ntb every thread of a warp is in
  // the same slot at the same instruction, so the write-through to the global ring is one 128-byte line per
  // warp and word; otherwise the slots differ from thread to thread and every 4-byte store is its own sector
  // (measured: 9x the write traffic and 1.8 ms instead of 1.1 ms at L = 449).
  if (pl.gring) L = (L + h->ntb - 1) / h->ntb * h->ntb;
  int nchunks = (O1 - O0 + L - 1) / L;
  long long total = (long long)nchunks * nstreams;
  int rc;
  const int gfw = h->h16b ? 32 : 16;
  if ((rc = h->d_G.reserve((size_t)total * gfw * 4))) return rc;
  if ((rc = h->d_F.reserve((size_t)total * gfw * 4))) return rc;
  const long long total16 = (total + 15) / 16 * 16;             // flag arrays are scanned 16 bytes at a time
  if ((rc = h->d_bad.reserve((size_t)total16 * 5))) return rc;   // verify's flags + two (bad, changed) pairs for the repair rounds
  if ((rc = h->d_counters.reserve(16))) return rc;
  if ((rc = h->h_counters.reserve(16 * kMaxStatRuns))) return rc;
  VitGeom g;
  g.codes = h->d_codes.as<uint32_t>();
  g.codes_stride = codes_stride;
  g.out = d_out;
  g.out_stride = out_stride;
  g.G = h->d_G.as<uint32_t>();
  g.F = h->d_F.as<uint32_t>();
  g.prevF = use_prevF ? h->d_prevF.as<uint32_t>() : nullptr;
  g.bad = h->d_bad.as<uint8_t>();
  g.counters = h->d_counters.as<unsigned int>();
  g.nstreams = nstreams; g.nchunks = nchunks; g.L = L; g.W = W; g.ntb = h->ntb;
  g.nbt = nbt; g.O0 = O0; g.O1 = O1; g.reset_at_0 = reset_at_0 ? 1 : 0;
  g.neg1 = 0xffffffffu; g.two = 2u; g.one = 1u;
  g.ring_depth = pl.depth;
  g.gring = nullptr;
  g.gf_words = gfw;

  DVBT_CUDA_TRY(cudaMemsetAsync(g.counters, 0, 16, h->stream));
  size_t smem = pl.smem;
  unsigned grid = (unsigned)((total + bd - 1) / bd);
  if (pl.gring) {
    if ((rc = h->d_gring.reserve((size_t)grid * h->ntb * 16 * bd * 4))) return rc;
    g.gring = h->d_gring.as<uint32_t>();
  }
  DVBT_CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
#ifdef DVBT_B200_LEGACY_ACS
  if (h->lanes == 2) {
    DVBT_CUDA_TRY(cudaFuncSetAttribute(vit_acs2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vit_acs2_kernel<<<grid, 2 * bd, smem, h->stream>>>(g);
  } else
#endif
  {
#ifdef DVBT_B200_LEGACY_ACS
    int fma = 0;
    if (const char *e = getenv("DVBT_B200_VIT_FMAADD")) fma = atoi(e);
#endif
    void (*kern)(VitGeom) =
        h->soft ? (pl.gring ? vit_acs_kernel<true, 0, 1, 0, true> : vit_acs_kernel<false, 0, 1, 0, true>) :
        h->h16b ? (pl.gring ? (bd == 384 ? vit_acs_kernel<true, 0, 2, 384> : bd == 512 ? vit_acs_kernel<true, 0, 2, 512>
                                                                                         : vit_acs_kernel<true, 0, 2, 0>)
                            : (bd == 512 ? vit_acs_kernel<false, 0, 2, 512> : bd == 256 ? vit_acs_kernel<false, 0, 2, 256>
                                                                                           : vit_acs_kernel<false, 0, 2, 0>))
        : h->h16 ? (pl.gring ? (bd == 384 ? vit_acs_kernel<true, 0, true, 384> : bd == 512 ? vit_acs_kernel<true, 0, true, 512>
                                                                                         : vit_acs_kernel<true, 0, true, 0>)
                           : (bd == 512 ? vit_acs_kernel<false, 0, true, 512> : bd == 256 ? vit_acs_kernel<false, 0, true, 256>
                                                                                           : vit_acs_kernel<false, 0, true, 0>))
#ifdef DVBT_B200_LEGACY_ACS
               : pl.gring ? (fma == 7 ? vit_acs_kernel<true, 7, false, 0> : fma == 4 ? vit_acs_kernel<true, 4, false, 0> : vit_acs_kernel<true, 0, false, 0>)
                          : (fma == 7 ? vit_acs_kernel<false, 7, false, 0> : fma == 4 ? vit_acs_kernel<false, 4, false, 0> : vit_acs_kernel<false, 0, false, 0>);
#else
               : nullptr;   // the byte-SWAR schedule is a legacy build option (create() refuses it otherwise)
    if (!kern) { set_error("viterbi: built without the legacy ACS schedules"); return DVBT_B200_EINVAL; }
#endif
    DVBT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, bd, smem, h->stream>>>(g);
  }
  DVBT_CUDA_TRY(cudaGetLastError());
  DVBT_CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
  if (gfw == 32) vit_verify_kernel<32><<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(g);
  else vit_verify_kernel<16><<<(unsigned)((total + 255) / 256), 256, 0, h->stream>>>(g);
  DVBT_CUDA_TRY(cudaGetLastError());
  size_t rsmem = acs_smem_bytes(h->ntb, 32, h->lutw());
  {
    // two parallel rounds (every isolated bad chunk, and pairs of consecutive ones), then the sequential kernel for the rest
    void (*round)(VitGeom, const uint8_t *, const uint8_t *, uint8_t *, uint8_t *, int) =
        h->soft ? vit_repair_round_kernel<1, true> :
#ifdef DVBT_B200_LEGACY_ACS
        h->h16b ? vit_repair_round_kernel<2> : (h->h16 && h->lanes == 1) ? vit_repair_round_kernel<1> : vit_repair_round_kernel<0>;
#else
        h->h16b ? vit_repair_round_kernel<2> : vit_repair_round_kernel<1>;
#endif
    void (*repair)(VitGeom, const uint8_t *, const uint8_t *, int) =
        h->soft ? vit_repair_kernel<1, true> :
#ifdef DVBT_B200_LEGACY_ACS
        h->h16b ? vit_repair_kernel<2> : (h->h16 && h->lanes == 1) ? vit_repair_kernel<1> : vit_repair_kernel<0>;
#else
        h->h16b ? vit_repair_kernel<2> : vit_repair_kernel<1>;
#endif
    DVBT_CUDA_TRY(cudaFuncSetAttribute(round, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
    DVBT_CUDA_TRY(cudaFuncSetAttribute(repair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
    uint8_t *f0 = g.bad, *b1 = g.bad + total16, *c1 = g.bad + 2 * total16, *b2 = g.bad + 3 * total16, *c2 = g.bad + 4 * total16;
    const unsigned rgrid = (unsigned)((total + 31) / 32);
    int rounds = 2;
    if (const char *e = getenv("DVBT_B200_VIT_REPAIR_ROUNDS")) rounds = atoi(e);   // 0: sequential repair only (A/B, tests)
    const uint8_t *bad_last = f0, *chg_last = nullptr;
    int left = -1;
    if (rounds >= 1) { round<<<rgrid, 32, rsmem, h->stream>>>(g, f0, nullptr, b1, c1, 2); bad_last = b1; chg_last = c1; left = 2; }
    if (rounds >= 2) { round<<<rgrid, 32, rsmem, h->stream>>>(g, b1, c1, b2, c2, 3); bad_last = b2; chg_last = c2; left = 3; }
    repair<<<(unsigned)((nstreams + 31) / 32), 32, rsmem, h->stream>>>(g, bad_last, chg_last, left);
    DVBT_CUDA_TRY(cudaGetLastError());
    dvbt::count_launch(3 + (rounds >= 2 ? 2 : rounds >= 1 ? 1 : 0));
  }
  {
    const int slot = h->st_runs < kMaxStatRuns ? h->st_runs : kMaxStatRuns - 1;
    DVBT_CUDA_TRY(cudaMemcpyAsync(h->h_counters.as<unsigned char>() + 16 * slot, g.counters, 16, cudaMemcpyDeviceToHost, h->stream));
    h->st_runs++;
  }
  // keep the true end state for a continuing stream (work()): F of each stream's last chunk
  if (nstreams == 1) {
    if ((rc = h->d_prevF.reserve((size_t)gfw * 4))) return rc;
    DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_prevF.p, g.F + (long long)(nchunks - 1) * gfw, (size_t)gfw * 4, cudaMemcpyDeviceToDevice, h->stream));
  }
  h->st_chunks += total;
  return 0;
}

int finish_stats(dvbt_b200_viterbi *h) {
  DVBT_CUDA_TRY(dvbt::stream_wait(h->stream));
  if (h->st_runs > 0) {
    h->st_repaired = 0;
    const int nslots = h->st_runs < kMaxStatRuns ? h->st_runs : kMaxStatRuns;
    for (int r = 0; r < nslots; r++) h->st_repaired += h->h_counters.as<unsigned int>()[4 * r + 1];
    DVBT_CUDA_TRY(cudaEventElapsedTime(&h->st_ms, h->ev0, h->ev1));   // the ACS kernel of the last run
  }
  h->st_runs = 0;
  return 0;
}

// ---- streaming decode shared by work() and the fused chain ---------------------------------------------------
// The step codes of the stream since the last reset are kept at the front of d_codes while they are few, else the
// last W + ntb + 8 byte times (warm-up + traceback of the first chunk of the next call).  stream_codes() returns
// where the caller writes the new_bt step codes of this call; stream_decode() decodes them and appends the bytes
// the reference would have produced (viterbi_decoder_impl.cc:275-312: nothing for the first ntb byte times after
// a reset, every byte time afterwards).
int stream_codes(dvbt_b200_viterbi *h, int new_bt, uint32_t **codes) {
  const int hist = (int)h->hist_bt;
  const size_t cwb = 4 * (size_t)h->cw();   // bytes per byte time
  const size_t want = (size_t)(hist + new_bt) * cwb;
  if (want > h->d_codes.cap) {
    dvbt::DevBuf bigger;
    int rc = bigger.reserve(want + want / 4);
    if (rc) return rc;
    if (hist > 0) DVBT_CUDA_TRY(cudaMemcpyAsync(bigger.p, h->d_codes.p, (size_t)hist * cwb, cudaMemcpyDeviceToDevice, h->stream));
    DVBT_CUDA_TRY(dvbt::stream_wait(h->stream));
    h->d_codes.release();
    h->d_codes = bigger;
  }
  *codes = h->d_codes.as<uint32_t>() + (size_t)hist * h->cw();
  return 0;
}

int stream_decode(dvbt_b200_viterbi *h, int new_bt, uint8_t *d_out, size_t *nprod_out, bool *first_after_reset, bool retain) {
  const int W = h->tune.warmup_bytes > 0 ? h->tune.warmup_bytes : 40;
  const int hist = (int)h->hist_bt;
  const int nbt = hist + new_bt;
  int O0 = h->init_done ? hist - h->ntb : 0;  // first out index owed (local)
  if (O0 < 0) O0 = 0;
  int rc = run_decode(h, nbt, nbt, 1, O0, h->hist_from_reset, h->have_prevF, d_out, 0);
  if (rc) return rc;
  const int O1 = nbt - h->ntb;
  const size_t nprod = O1 > O0 ? (size_t)(O1 - O0) : 0;
  // retain history for the next call
  const int keep = W + h->ntb + 8;
  if (!retain) {
    h->hist_bt = 0;   // end of the stream: the next decode starts from a reset anyway
  } else if (nbt > keep && !(h->hist_from_reset && nbt <= 2 * keep)) {
    const size_t cwb = 4 * (size_t)h->cw();
    if ((rc = h->d_tmp.reserve((size_t)keep * cwb))) return rc;
    DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_tmp.p, h->d_codes.as<uint8_t>() + (size_t)(nbt - keep) * cwb, (size_t)keep * cwb, cudaMemcpyDeviceToDevice, h->stream));
    DVBT_CUDA_TRY(cudaMemcpyAsync(h->d_codes.p, h->d_tmp.p, (size_t)keep * cwb, cudaMemcpyDeviceToDevice, h->stream));
    h->hist_bt = keep;
    h->hist_from_reset = false;
  } else {
    h->hist_bt = nbt;
  }
  h->have_prevF = nprod > 0 || h->have_prevF;
  if (first_after_reset) *first_after_reset = !h->init_done;
  h->init_done = true;
  if (nprod_out) *nprod_out = nprod;
  return 0;
}

}  // namespace

namespace dvbt {
// internal interface used by the fused receive chain (rx_chain.cu)
cudaStream_t vit_stream(dvbt_b200_viterbi *h) { return h->stream; }
bool vit_is_soft(const dvbt_b200_viterbi *h) { return h->soft; }
int vit_params(const dvbt_b200_viterbi *h, int *k, int *n, int *m, int *ntb, int *nsymbols, int *nout) {
  *k = h->k; *n = h->n; *m = h->m; *ntb = h->ntb; *nsymbols = h->nsymbols; *nout = h->nout;
  return 0;
}
uint32_t *vit_reserve_codes(dvbt_b200_viterbi *h, size_t nbt) {
  if (h->d_codes.reserve(nbt * 4 * h->cw())) return nullptr;
  return h->d_codes.as<uint32_t>();
}
// decode byte times [0, nbt) already present in the codes buffer, from a reset; no synchronisation
int vit_decode_prepared(dvbt_b200_viterbi *h, int nbt, uint8_t *d_out) {
  return run_decode(h, nbt, nbt, 1, 0, true, false, d_out, 0);
}
int vit_collect_stats(dvbt_b200_viterbi *h) { return finish_stats(h); }
// streaming flavour for the fused chain: see stream_codes() / stream_decode() above
void vit_stream_reset(dvbt_b200_viterbi *h) { dvbt_b200_viterbi_reset(h); }
void vit_stream_accumulate_stats(dvbt_b200_viterbi *h, bool on) { h->st_accumulate = on; if (!on) h->st_runs = 0; }
int vit_stream_codes(dvbt_b200_viterbi *h, int new_bt, uint32_t **codes) { return stream_codes(h, new_bt, codes); }
int vit_stream_decode(dvbt_b200_viterbi *h, int new_bt, uint8_t *d_out, size_t *nprod, bool *first_after_reset, bool retain) {
  return stream_decode(h, new_bt, d_out, nprod, first_after_reset, retain);
}
}  // namespace dvbt

extern "C" {

int dvbt_b200_viterbi_create(const dvbt_b200_viterbi_params *p, dvbt_b200_viterbi **out) {
  if (!p || !out) { set_error("viterbi_create: null argument"); return DVBT_B200_EINVAL; }
  *out = nullptr;
  if (p->constellation < DVBT_QPSK || p->constellation > DVBT_QAM64 || p->code_rate < DVBT_C1_2 ||
      p->code_rate > DVBT_C7_8 || p->bsize <= 0) {
    set_error("viterbi_create: bad constellation/code_rate/bsize (%d,%d,%d)", p->constellation, p->code_rate, p->bsize);
    return DVBT_B200_EINVAL;
  }
  int rc = dvbt::ensure_device();
  if (rc) return rc;
  dvbt_b200_viterbi *h = new (std::nothrow) dvbt_b200_viterbi();
  if (!h) { set_error("viterbi_create: out of memory"); return DVBT_B200_ENOMEM; }
  h->par = *p;
  h->k = rate_k(p->code_rate);
  h->n = rate_n(p->code_rate);
  h->m = 2 * (p->constellation + 1);  // dvbt_config.cc: QPSK 2, QAM16 4, QAM64 6
  h->ntb = kNtb[p->code_rate];
  // viterbi_decoder_impl.cc:137-153; the reference asserts these under !NDEBUG
  if ((p->bsize * h->n) % h->m != 0 || (p->bsize * h->k) % 8 != 0) {
    set_error("viterbi_create: bsize %d does not give whole symbols/bytes", p->bsize);
    delete h;
    return DVBT_B200_EINVAL;
  }
  h->nsymbols = p->bsize * h->n / h->m;
  h->nout = p->bsize * h->k / 8;
  h->tune = dvbt_b200_viterbi_tuning{0, 0, 0, 0};
  {
    const char *e = getenv("DVBT_B200_VIT_LANES");
    // one lane per chunk is the default: on B200 the two-lane kernel (twice the warps, 24 % more instructions)
    // reaches 72 Gbit/s at rate 7/8 against 104 Gbit/s — the kernel is ALU-pipe bound, not latency bound
    h->lanes = (e && e[0] == '2') ? 2 : 1;
#ifndef DVBT_B200_LEGACY_ACS
    // the byte-SWAR and two-lane schedules (measured dead ends) are compiled only with -DDVBT_B200_LEGACY_ACS
    if (h->lanes == 2 || (getenv("DVBT_B200_VIT_ACS") && getenv("DVBT_B200_VIT_ACS")[0] == 's')) {
      set_error("viterbi_create: DVBT_B200_VIT_LANES=2 / DVBT_B200_VIT_ACS=swar need a library built with -DDVBT_B200_LEGACY_ACS");
      delete h;
      return DVBT_B200_EINVAL;
    }
#endif
    const char *a = getenv("DVBT_B200_VIT_ACS");
    h->h16 = !(a && a[0] == 's') && h->lanes == 1;
    h->h16b = h->h16 && a && !strcmp(a, "h16b");
  }
  h->h_stage_in.host = h->h_stage_out.host = h->h_counters.host = true;
  cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
  if (e != cudaSuccess) {
    set_error("viterbi_create: %s", cudaGetErrorString(e));
    dvbt_b200_viterbi_destroy(h);
    return DVBT_B200_ECUDA;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
  *out = h;
  return 0;
}

void dvbt_b200_viterbi_destroy(dvbt_b200_viterbi *h) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) return;
  if (h->stream) cudaStreamSynchronize(h->stream);
  dvbt::DevBuf *bufs[] = {&h->d_in, &h->d_codes, &h->d_out, &h->d_G, &h->d_F, &h->d_bad, &h->d_counters,
                          &h->d_prevF, &h->d_tmp, &h->d_gring, &h->h_stage_in, &h->h_stage_out, &h->h_counters};
  for (auto *b : bufs) b->release();
  h->stg.release();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

int dvbt_b200_viterbi_set_tuning(dvbt_b200_viterbi *h, const dvbt_b200_viterbi_tuning *t) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || !t) { set_error("viterbi_set_tuning: null argument"); return DVBT_B200_EINVAL; }
  if (t->threads_per_block < 0 || t->threads_per_block > 512 || (t->threads_per_block % 32) != 0 ||
      t->chunk_bytes < 0 || t->warmup_bytes < 0 || t->ring_depth < 0) {
    set_error("viterbi_set_tuning: threads_per_block must be a multiple of 32 <= 512, sizes >= 0");
    return DVBT_B200_EINVAL;
  }
  if (t->threads_per_block > 0 && (h->lanes == 2 ? acs_smem_bytes(h->ntb, t->threads_per_block / 2)
                                                 : acs_ring_smem(h->ntb, t->ring_depth > 0 && t->ring_depth < h->ntb ? t->ring_depth : h->ntb,
                                                                 t->threads_per_block)) > 227 * 1024) {
    set_error("viterbi_set_tuning: %d threads need %zu B of shared memory (> 227 KB)", t->threads_per_block,
              acs_smem_bytes(h->ntb, t->threads_per_block));
    return DVBT_B200_EINVAL;
  }
  h->tune = *t;
  return 0;
}

int dvbt_b200_viterbi_reset(dvbt_b200_viterbi *h) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) { set_error("viterbi_reset: null handle"); return DVBT_B200_EINVAL; }
  h->init_done = false;
  h->hist_bt = 0;
  h->hist_from_reset = true;
  h->have_prevF = false;
  return 0;
}

int dvbt_b200_viterbi_forecast(const dvbt_b200_viterbi *h, int noutput_items) {
  if (!h) return DVBT_B200_EINVAL;
  return noutput_items * 8 * h->n / (h->k * h->m);  // viterbi_decoder_impl.cc:183
}
int dvbt_b200_viterbi_output_multiple(const dvbt_b200_viterbi *h) { return h ? h->nout : DVBT_B200_EINVAL; }
int dvbt_b200_viterbi_ntraceback(const dvbt_b200_viterbi *h) { return h ? h->ntb : DVBT_B200_EINVAL; }

int dvbt_b200_viterbi_last_stats(const dvbt_b200_viterbi *h, long long *chunks, long long *repaired, float *ms) {
  if (!h) return DVBT_B200_EINVAL;
  if (chunks) *chunks = h->st_chunks;
  if (repaired) *repaired = h->st_repaired;
  if (ms) *ms = h->st_ms;
  return 0;
}

static int decode_batch(dvbt_b200_viterbi *h, const uint8_t *in, size_t in_stride, size_t n_in, int nstreams,
                        uint8_t *out, size_t out_stride, size_t *n_out, bool host) {
  if (!h || !in || !out || nstreams <= 0) { set_error("viterbi_decode: bad argument"); return DVBT_B200_EINVAL; }
  if (h->soft) { set_error("viterbi_decode: the handle is in soft-decision mode (use dvbt_b200_viterbi_decode_soft_*)"); return DVBT_B200_EINVAL; }
  unsigned long long bits = (unsigned long long)n_in * h->m * h->k;
  if (bits % (8ull * h->n) != 0) {
    set_error("viterbi_decode: %zu input bytes are not a whole number of decoded bytes", n_in);
    return DVBT_B200_EINVAL;
  }
  unsigned long long nbt64 = bits / (8ull * h->n);
  if (nbt64 >= (1ull << 30)) { set_error("viterbi_decode: stream too long (%llu byte times)", nbt64); return DVBT_B200_EINVAL; }
  int nbt = (int)nbt64;
  size_t produced = nbt > h->ntb ? (size_t)(nbt - h->ntb) : 0;
  if (n_out) *n_out = produced;
  if (produced == 0) return 0;
  if (in_stride < n_in || out_stride < produced) { set_error("viterbi_decode: stride smaller than the stream"); return DVBT_B200_EINVAL; }
  int rc;
  const uint8_t *d_in = in;
  uint8_t *d_out = out;
  if (host) {
    if ((rc = h->d_in.reserve(n_in * nstreams))) return rc;
    if ((rc = h->d_out.reserve(produced * nstreams))) return rc;
    DVBT_CUDA_TRY(cudaMemcpy2DAsync(h->d_in.p, n_in, in, in_stride, n_in, nstreams, cudaMemcpyHostToDevice, h->stream));
    d_in = h->d_in.as<uint8_t>();
    d_out = h->d_out.as<uint8_t>();
  }
  if ((rc = h->d_codes.reserve((size_t)nbt * nstreams * 4))) return rc;
  rc = launch_depuncture(h->par.code_rate, h->m, d_in, host ? (long long)n_in : (long long)in_stride,
                         h->d_codes.as<uint32_t>(), nbt, nbt, nstreams, 0, h->stream);
  if (rc) return rc;
  rc = run_decode(h, nbt, nbt, nstreams, 0, true, false, d_out, host ? (long long)produced : (long long)out_stride);
  if (rc) return rc;
  if (host)
    DVBT_CUDA_TRY(cudaMemcpy2DAsync(out, out_stride, h->d_out.p, produced, produced, nstreams, cudaMemcpyDeviceToHost, h->stream));
  return finish_stats(h);
}

// Soft-decision mode (beyond the reference, whose d_metrics.c is a stub - TODO.txt:25): see include/dvbt_b200.h
int dvbt_b200_viterbi_set_soft(dvbt_b200_viterbi *h, int on) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h) { set_error("viterbi_set_soft: null handle"); return DVBT_B200_EINVAL; }
  if (on && (!h->h16 || h->h16b || h->lanes != 1)) {
    set_error("viterbi_set_soft: soft decisions use the default (h16) ACS schedule");
    return DVBT_B200_EINVAL;
  }
  if ((on != 0) != h->soft) {
    if (h->stream) DVBT_CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->soft = on != 0;
    dvbt_b200_viterbi_reset(h);   // the step-code history has the other format
  }
  return 0;
}

static int decode_soft(dvbt_b200_viterbi *h, const int8_t *in, size_t n_in, uint8_t *out, size_t *n_out, bool host) {
  if (!h || !in || !out) { set_error("viterbi_decode_soft: bad argument"); return DVBT_B200_EINVAL; }
  if (!h->soft) { set_error("viterbi_decode_soft: call dvbt_b200_viterbi_set_soft(h, 1) first"); return DVBT_B200_EINVAL; }
  unsigned long long bits = (unsigned long long)n_in * h->k;
  if (bits % (8ull * h->n) != 0) {
    set_error("viterbi_decode_soft: %zu code bits are not a whole number of decoded bytes", n_in);
    return DVBT_B200_EINVAL;
  }
  unsigned long long nbt64 = bits / (8ull * h->n);
  if (nbt64 >= (1ull << 30)) { set_error("viterbi_decode_soft: stream too long (%llu byte times)", nbt64); return DVBT_B200_EINVAL; }
  const int nbt = (int)nbt64;
  const size_t produced = nbt > h->ntb ? (size_t)(nbt - h->ntb) : 0;
  if (n_out) *n_out = produced;
  if (produced == 0) return 0;
  int rc;
  const int8_t *d_in = in;
  uint8_t *d_out = out;
  if (host) {
    if ((rc = h->d_in.reserve(n_in))) return rc;
    if ((rc = h->d_out.reserve(produced))) return rc;
    if ((rc = h->stg.h2d(h->d_in.p, in, n_in, h->stream))) return rc;
    d_in = h->d_in.as<int8_t>();
    d_out = h->d_out.as<uint8_t>();
  }
  if ((rc = h->d_codes.reserve((size_t)nbt * 8))) return rc;
  {
    const unsigned grid = (unsigned)((nbt + 255) / 256);
    uint2 *c2 = h->d_codes.as<uint2>();
    switch (h->par.code_rate) {
      case 0: vit_depuncture_soft_kernel<0><<<grid, 256, 0, h->stream>>>(d_in, c2, nbt); break;
      case 1: vit_depuncture_soft_kernel<1><<<grid, 256, 0, h->stream>>>(d_in, c2, nbt); break;
      case 2: vit_depuncture_soft_kernel<2><<<grid, 256, 0, h->stream>>>(d_in, c2, nbt); break;
      case 3: vit_depuncture_soft_kernel<3><<<grid, 256, 0, h->stream>>>(d_in, c2, nbt); break;
      default: vit_depuncture_soft_kernel<4><<<grid, 256, 0, h->stream>>>(d_in, c2, nbt); break;
    }
    dvbt::count_launch();
    DVBT_CUDA_TRY(cudaGetLastError());
  }
  if ((rc = run_decode(h, nbt, nbt, 1, 0, true, false, d_out, 0))) return rc;
  if (host && (rc = h->stg.d2h(out, h->d_out.p, produced, h->stream))) return rc;
  return finish_stats(h);
}

int dvbt_b200_viterbi_decode_soft_host(dvbt_b200_viterbi *h, const int8_t *in, size_t n_in, uint8_t *out, size_t *n_out) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  return decode_soft(h, in, n_in, out, n_out, true);
}
int dvbt_b200_viterbi_decode_soft_dev(dvbt_b200_viterbi *h, const int8_t *in, size_t n_in, uint8_t *out, size_t *n_out) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (h) {
    if (int rc = dvbt::join_default_stream(h->stream)) return rc;
  }
  return decode_soft(h, in, n_in, out, n_out, false);
}

int dvbt_b200_viterbi_decode_host(dvbt_b200_viterbi *h, const uint8_t *in, size_t in_stride, size_t n_in,
                                  int nstreams, uint8_t *out, size_t out_stride, size_t *n_out) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  return decode_batch(h, in, in_stride, n_in, nstreams, out, out_stride, n_out, true);
}
int dvbt_b200_viterbi_decode_dev(dvbt_b200_viterbi *h, const uint8_t *in, size_t in_stride, size_t n_in,
                                 int nstreams, uint8_t *out, size_t out_stride, size_t *n_out) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (h) {
    if (int rc = dvbt::join_default_stream(h->stream)) return rc;
  }
  return decode_batch(h, in, in_stride, n_in, nstreams, out, out_stride, n_out, false);
}

int dvbt_b200_viterbi_work(dvbt_b200_viterbi *h, const uint8_t *in, size_t n_in_items, uint8_t *out,
                           size_t noutput_items, size_t *consumed, size_t *produced,
                           const dvbt_b200_tag *tags_in, size_t n_tags_in, dvbt_b200_tag *tags_out,
                           size_t tags_out_capacity, size_t *n_tags_out) {
  dvbt::DeviceScope dev_scope__(h ? h->device : -1);
  if (!h || !consumed || !produced) { set_error("viterbi_work: null argument"); return DVBT_B200_EINVAL; }
  if (h->soft) { set_error("viterbi_work: the handle is in soft-decision mode; the block interface takes hard decisions"); return DVBT_B200_EINVAL; }
  *consumed = 0;
  *produced = 0;
  if (n_tags_out) *n_tags_out = 0;
  if (noutput_items % (size_t)h->nout != 0) {
    set_error("viterbi_work: noutput_items %zu is not a multiple of %d", noutput_items, h->nout);
    return DVBT_B200_EINVAL;
  }
  size_t nblocks = 8 * noutput_items / ((size_t)h->par.bsize * h->k);  // viterbi_decoder_impl.cc:198
  size_t need = nblocks * (size_t)h->nsymbols;
  if (n_in_items < need) {
    set_error("viterbi_work: %zu input items given, forecast() asks for %zu", n_in_items, need);
    return DVBT_B200_EINVAL;
  }
  if (nblocks == 0) return 0;
  if (!in || !out) { set_error("viterbi_work: null buffer"); return DVBT_B200_EINVAL; }
  // superframe_start inside the window (viterbi_decoder_impl.cc:213-229); only the first matters
  bool tagged = false;
  uint64_t tag_off = 0;
  for (size_t i = 0; i < n_tags_in; i++) {
    if (tags_in[i].key == DVBT_TAG_SUPERFRAME_START && tags_in[i].offset < need) {
      if (!tagged || tags_in[i].offset < tag_off) tag_off = tags_in[i].offset;
      tagged = true;
    }
  }
  if (tagged) {
    dvbt_b200_viterbi_reset(h);
    if (tag_off != 0) {
      *consumed = (size_t)tag_off;
      return 0;
    }
  }
  int new_bt = (int)(nblocks * (size_t)h->nout);
  int rc;
  if ((rc = h->d_in.reserve(need))) return rc;
  if ((rc = h->d_out.reserve((size_t)new_bt))) return rc;
  uint32_t *codes = nullptr;
  if ((rc = stream_codes(h, new_bt, &codes))) return rc;
  if ((rc = h->stg.h2d(h->d_in.p, in, need, h->stream))) return rc;
  // each call starts on a block boundary, where the puncture phase and the bit position in the
  // symbol are both zero (2*k*bsize symbols per block), so the new codes are position independent
  rc = launch_depuncture(h->par.code_rate, h->m, h->d_in.as<uint8_t>(), (long long)need, codes, new_bt, new_bt, 1, 0, h->stream);
  if (rc) return rc;
  size_t nprod = 0;
  bool first = false;
  if ((rc = stream_decode(h, new_bt, h->d_out.as<uint8_t>(), &nprod, &first, true))) return rc;
  if (nprod && (rc = h->stg.d2h(out, h->d_out.p, nprod, h->stream))) return rc;
  if ((rc = finish_stats(h))) return rc;
  *consumed = need;  // :320
  if (first) {
    // first producing call after a reset (:298-312)
    if (tags_out && tags_out_capacity > 0 && n_tags_out) {
      tags_out[0].offset = 0;
      tags_out[0].key = DVBT_TAG_SUPERFRAME_START;
      tags_out[0].value = 1;
      *n_tags_out = 1;
    }
  }
  *produced = nprod;
  return 0;
}

}  // extern "C"
