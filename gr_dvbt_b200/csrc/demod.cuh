// Internal interface of the demod_reference_signals kernels (demod.cu) used by the fused
// receive chain (rx_chain.cu).
#pragma once
#include "common.cuh"

namespace dvbt {

struct DemapTable {
  float2 pts[64];
  int size;
  // per-axis view for demap_cell_near: the 2^(m/2) levels of an axis are g * n, n = ..., -(alpha+2), -alpha,
  // alpha, alpha+2, ... (dvbt_demap_impl.cc:98-159); level k (ascending) contributes the index bits
  // (bx >> 8k) & 0xff on the I axis, (by >> 8k) & 0xff on the Q axis.  near_ok = 0 disables the shortcut.
  float g, inv_step;
  int alpha, near_ok;
  unsigned long long bx, by;
  // soft-decision view (demap_cell_soft; beyond the reference): the levels of the I / Q axis in ascending order, and for
  // bit e of a cell (e = 0 the first = most significant bit; even e ride on I, odd e on Q) the set of levels where it is 1
  float soft_lv[2][8];
  unsigned char soft_ones[8];
  float soft_scale;   // soft value of a cell sitting exactly on its constellation point next to a decision boundary (default 4)
};
int make_demap_table(int constellation, int hierarchy, float gain, DemapTable *t);

#ifdef __CUDACC__
// find_constellation_value (dvbt_demap_impl.cc:167-203): first index with the strictly smallest
// squared distance, every float operation rounded on its own.  The index bits alternate between
// the axes (b0 b2 b4 -> I, b1 b3 b5 -> Q, :141-156), so re - point.re takes only 2^(M/2) distinct
// values: the per-axis squares ax[], ay[] are computed once and every distance the reference
// compares is fl(ax[xa] + ay[ya]) of the same two floats.
//
// Shortcut (exact): rounding is monotone, so the smallest distance is V = fl(min ax + min ay), and a
// pair (xa, ya) can only reach V if fl(ax[xa] + min ay) == V and fl(min ax + ay[ya]) == V.  When
// exactly one xa and one ya pass that test the argmin is unique and no scan is needed (2^(M/2+1)
// additions instead of 2^M).  Otherwise - a tie between constellation points, or Inf/NaN input, for
// which every comparison is false - the reference's scan in index order decides, as before.
template <int M>
__device__ __forceinline__ uint8_t demap_cell_exact(const DemapTable &t, float2 v) {
  constexpr int H = M / 2, L = 1 << H, SIZE = 1 << M;
  float ax[L], ay[L];
#pragma unroll
  for (int a = 0; a < L; a++) {
    int ix = 0, iy = 0;
#pragma unroll
    for (int j = 0; j < H; j++) {
      int bit = (a >> (H - 1 - j)) & 1;
      ix |= bit << (M - 1 - 2 * j);
      iy |= bit << (M - 2 - 2 * j);
    }
    float dr = __fsub_rn(v.x, t.pts[ix].x), di = __fsub_rn(v.y, t.pts[iy].y);
    ax[a] = __fmul_rn(dr, dr);
    ay[a] = __fmul_rn(di, di);
  }
  float mx = ax[0], my = ay[0];
#pragma unroll
  for (int a = 1; a < L; a++) { mx = fminf(mx, ax[a]); my = fminf(my, ay[a]); }
  const float V = __fadd_rn(mx, my);
  int cx = 0, cy = 0, bx = 0, by = 0;
#pragma unroll
  for (int a = 0; a < L; a++) {
    int ix = 0, iy = 0;
#pragma unroll
    for (int j = 0; j < H; j++) {
      int bit = (a >> (H - 1 - j)) & 1;
      ix |= bit << (M - 1 - 2 * j);
      iy |= bit << (M - 2 - 2 * j);
    }
    bool ex = __fadd_rn(ax[a], my) == V, ey = __fadd_rn(mx, ay[a]) == V;
    cx += ex ? 1 : 0;
    cy += ey ? 1 : 0;
    bx = ex ? ix : bx;
    by = ey ? iy : by;
  }
  if (cx == 1 && cy == 1) return (uint8_t)(bx | by);
  float min_dist = __fadd_rn(ax[0], ay[0]);
  int min_index = 0;
#pragma unroll
  for (int i = 1; i < SIZE; i++) {
    int xa = 0, ya = 0;
#pragma unroll
    for (int j = 0; j < H; j++) {
      xa |= ((i >> (M - 1 - 2 * j)) & 1) << (H - 1 - j);
      ya |= ((i >> (M - 2 - 2 * j)) & 1) << (H - 1 - j);
    }
    float d = __fadd_rn(ax[xa], ay[ya]);
    if (d < min_dist) { min_dist = d; min_index = i; }
  }
  return (uint8_t)min_index;
}

// Nearest-level shortcut in front of demap_cell_exact (same result, a third of the operations).  Along one
// axis the squared distance fl(fl(v - level)^2) is unimodal in the (sorted) level index, because rounding is
// monotone.  Guess the nearest level of each axis by division, compute its squared distance c and those of
// its two neighbours l, r, and V = fl(cx + cy).  If fl(lx + cy), fl(rx + cy), fl(cx + ly), fl(cx + ry) are
// all > V, then cx and cy are the strict per-axis minima and every other point of the constellation is at a
// distance >= one of those four sums > V: the reference's scan (first strictly smallest distance) returns
// exactly the guessed point.  Anything else - a tie, a guess off by one, Inf/NaN - goes to demap_cell_exact.
template <int M>
__device__ __forceinline__ uint8_t demap_cell_near(const DemapTable &t, float2 v) {
  constexpr float TOP = (float)((1 << (M / 2)) - 1);   // the levels of an axis are g * n, n = -TOP, ..., -1, 1, ..., TOP (alpha = 1)
  const float inf = __int_as_float(0x7f800000);
  // nearest level as a float odd integer (no int<->float conversions on the hot path), its two neighbours
  float nx = fminf(fmaxf(fmaf(2.0f, floorf(v.x * t.inv_step), 1.0f), -TOP), TOP);
  float ny = fminf(fmaxf(fmaf(2.0f, floorf(v.y * t.inv_step), 1.0f), -TOP), TOP);
  auto sq = [&](float a, float n) {
    float d = __fsub_rn(a, __fmul_rn(t.g, n));   // the level exactly as make_demap_table rounds it
    return __fmul_rn(d, d);
  };
  // Clear cases first: the cell is within 0.49 level spacings of the guessed level on both axes (on the open side of
  // an edge level: anywhere up to 16 spacings out).  With s = 2 g the spacing, every other level of an axis is then
  // at least 0.5099 s away (the levels are fl(g n): spacing error < 1e-6 s), so its squared distance exceeds the
  // guessed one by more than 0.0197 s^2, while fl(cx + cy) < 512 s^2: the gap is > 300 ulp of any sum the scan
  // compares, every other point is strictly farther after rounding, and the reference's scan (first strictly
  // smallest distance) returns the guessed point.  Inf/NaN fail the comparisons and take the paths below.
  {
    const float lim = 0.98f * t.g, far = 32.0f * t.g;
    const float dx = __fsub_rn(v.x, __fmul_rn(t.g, nx)), dy = __fsub_rn(v.y, __fmul_rn(t.g, ny));
    const bool okx = (dx > -lim || nx == -TOP) && (dx < lim || nx == TOP) && fabsf(dx) < far;
    const bool oky = (dy > -lim || ny == -TOP) && (dy < lim || ny == TOP) && fabsf(dy) < far;
    if (okx && oky) {
      int kx = __float2int_rn((nx + TOP) * 0.5f), ky = __float2int_rn((ny + TOP) * 0.5f);
      unsigned bx = __byte_perm((unsigned)t.bx, (unsigned)(t.bx >> 32), kx), by = __byte_perm((unsigned)t.by, (unsigned)(t.by >> 32), ky);
      return (uint8_t)((bx | by) & 0xffu);
    }
  }
  float cx = sq(v.x, nx), cy = sq(v.y, ny);
  float lx = nx > -TOP ? sq(v.x, nx - 2.0f) : inf, rx = nx < TOP ? sq(v.x, nx + 2.0f) : inf;
  float ly = ny > -TOP ? sq(v.y, ny - 2.0f) : inf, ry = ny < TOP ? sq(v.y, ny + 2.0f) : inf;
  float V = __fadd_rn(cx, cy);
  bool ok = (__fadd_rn(lx, cy) > V) & (__fadd_rn(rx, cy) > V) & (__fadd_rn(cx, ly) > V) & (__fadd_rn(cx, ry) > V);
  if (ok) {
    int kx = __float2int_rn((nx + TOP) * 0.5f), ky = __float2int_rn((ny + TOP) * 0.5f);   // level number 0 .. 2^(M/2)-1, ascending
    unsigned bx = __byte_perm((unsigned)t.bx, (unsigned)(t.bx >> 32), kx), by = __byte_perm((unsigned)t.by, (unsigned)(t.by >> 32), ky);
    return (uint8_t)((bx | by) & 0xffu);
  }
  return demap_cell_exact<M>(t, v);
}

// Soft decisions (beyond the reference, whose demapper is hard only): per bit the max-log metric
//     (min over the levels where the bit is 0 of (x - level)^2  -  min over the levels where it is 1 of the same) * w,
// x the I or Q coordinate the bit rides on, rounded to the nearest integer and clamped to +-6; > 0 = "1".  w carries the
// scale (soft_scale / (4 g^2): a cell on its constellation point next to a boundary gets +-soft_scale) and the cell's
// channel-state weight |H|^2 / mean |H|^2.  Result: value + 8 of bit e in nibble e (Inf / NaN cells give 0 = erased).
// The sign of a non-zero value is the hard decision of demap_cell_exact (nearest level along each axis).
template <int M>
__device__ __forceinline__ uint32_t demap_cell_soft(const DemapTable &t, float2 v, float w) {
  constexpr int H = M / 2, L = 1 << H;
  const float inf = __int_as_float(0x7f800000);
  uint32_t word = 0;
#pragma unroll
  for (int a = 0; a < 2; a++) {
    const float x = a ? v.y : v.x;
    float d[L];
#pragma unroll
    for (int k = 0; k < L; k++) { const float df = x - t.soft_lv[a][k]; d[k] = df * df; }
#pragma unroll
    for (int j = 0; j < H; j++) {
      const int e = 2 * j + a;
      const unsigned ones = t.soft_ones[e];
      float d0 = inf, d1 = inf;
#pragma unroll
      for (int k = 0; k < L; k++) {
        const bool one = (ones >> k) & 1u;
        d0 = one ? d0 : fminf(d0, d[k]);
        d1 = one ? fminf(d1, d[k]) : d1;
      }
      int q = __float2int_rn((d0 - d1) * w);   // NaN -> 0
      q = max(-6, min(6, q));
      word |= (uint32_t)(q + 8) << (4 * e);
    }
  }
  return word;
}
__device__ __forceinline__ uint32_t demap_cell_soft_any(const DemapTable &t, float2 v, float w) {
  if (t.size == 64) return demap_cell_soft<6>(t, v, w);
  if (t.size == 16) return demap_cell_soft<4>(t, v, w);
  return demap_cell_soft<2>(t, v, w);
}

__device__ __forceinline__ uint8_t demap_cell_any(const DemapTable &t, float2 v) {
  if (t.near_ok) {
    if (t.size == 64) return demap_cell_near<6>(t, v);
    if (t.size == 16) return demap_cell_near<4>(t, v);
    return demap_cell_near<2>(t, v);
  }
  if (t.size == 64) return demap_cell_exact<6>(t, v);
  if (t.size == 16) return demap_cell_exact<4>(t, v);
  return demap_cell_exact<2>(t, v);
}
#endif
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
  const int *pay32;        // [4][P] the same with the distance to the channel-estimation carrier below: k | (k - prevp[k]) << 16
  const int *pay3;         // [4][P] payload carrier k | (k - k0) << 13 | ordinal of k0 among the phase's pilots << 17
  const int *tps3;         // [4][ntps] the same for the TPS carriers
  const short *H;          // [P] symbol interleaver permutation H(q) (symbol_inner_interleaver_impl.cc:35-96)
  const short *Hinv;       // [P]
  const short *pilots;     // [4][pil_stride] channel-estimation carriers of scattered phase r, increasing
  int npil[4], pil_stride;
};

struct ModeTables {
  ModeDev dev;
  DevBuf blob;
  int init(int transmission_mode, int guard_interval);
  void release() { blob.release(); }
};

constexpr int kMaxSfTags = 64;   // superframe_start tags one scan can report (one per re-synchronisation inside the batch)

// Sequential receiver state of pilot_gen + demod_reference_signals_impl (device resident).
struct DemodState {
  int symbol_index, known, frame_index, prev_mod, mod, d_init;
  unsigned char fifo[68];
  float2 prev_tps[68];
  // results of the last scan
  int first_out;   // index (within the batch) of the first symbol that was output, -1 if none
  int n_out;       // symbols output by the last scan
  int sf_tag_at;   // output index carrying the (first) superframe_start tag in the last scan, -1 if none
  int n_sf;        // superframe_start tags sent in the last scan (a sync_start inside the batch re-arms the gating,
  int sf_at[kMaxSfTags];   // demod_reference_signals_impl.cc:112-116) and the output indices of the first kMaxSfTags
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
  cudaEvent_t ev_eq0 = nullptr, ev_eq1 = nullptr;   // optional: recorded around demod_equalise_kernel (bench timing)
  // optional: a second stream + two events.  The TPS vote and the sequential scan (one block, ~0.1 ms per 20 000 symbols)
  // then run beside the wide equalise + demap kernel instead of behind it (the TPS carriers are equalised by stage 1)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_mid = nullptr;   // ev_mid: recorded behind the vote (see demod_run)
};

// Runs process_cpilot_data / compute_oneshot_csft / frequency_correction / process_spilot_data
// (phase detection) for symbols [0, nparse) of X (nparse+1 symbols must be readable), then the
// channel estimate + equalisation (+ optional demap) of every parsed symbol, the TPS votes and the
// sequential scan.  Y (optional) receives P equalised cells per *parsed* symbol (not compacted);
// dm (optional) the demapped bytes per parsed symbol; soft (optional, instead of dm): one word of soft decisions per cell
// (demap_cell_soft), P per parsed symbol.
// sync_at / nsync: batch indices of the symbols that carry a sync_start tag (device array, may be null);
// sync_start_at0 = 1 is the same as listing symbol 0.  src_base is added to the values written to out_src.
int demod_run(const ModeDev &md, const DemapTable *demap, const float2 *X, int nparse, DemodBuffers b,
              DemodState *d_state, int fi_start, int sync_start_at0, float2 *Y, uint8_t *dm, cudaStream_t st,
              const int *sync_at = nullptr, int nsync = 0, int src_base = 0, uint32_t *soft = nullptr);

}  // namespace dvbt
