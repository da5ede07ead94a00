"""Builds tests/emul/_build/libvit_emul.so: the DEVICE code of gr_dvbt_b200/csrc/viterbi.cu (depuncture, the one-lane ACS
kernels of every schedule, verify, repair - text taken from the file as it is) compiled for the host on top of
tests/emul/cuda_host_emul.h, plus a launcher that mirrors run_decode()'s geometry.  Test infrastructure: it lets a
CPU-only box run the kernels' own source against the oracle (tests/test_viterbi_host_emul_cpu.py).  What it cannot show:
anything that depends on the hardware (PRMT / VIADDMNMX semantics are restated in the shim from the PTX ISA and from
CUDA's own host fallback), launch limits, shared-memory sizes, timing."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "gr_dvbt_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD, "libvit_emul.so")

LAUNCHER = r'''
// ---------------------------------------------------------------------------------------------------------------
// launcher (mirrors run_decode() of viterbi.cu: one stream, from a reset)
// ---------------------------------------------------------------------------------------------------------------
namespace {
template <int RATE>
void emul_depuncture(int m, const uint8_t *in, uint32_t *codes, long long nbt) {
  unsigned grid = (unsigned)((nbt + 63) / 64);
  if (m == 2) emul_launch(vit_depuncture_kernel<RATE, 2>, grid, 64u, in, 0LL, codes, 0LL, nbt, 1, 0LL);
  else if (m == 4) emul_launch(vit_depuncture_kernel<RATE, 4>, grid, 64u, in, 0LL, codes, 0LL, nbt, 1, 0LL);
  else emul_launch(vit_depuncture_kernel<RATE, 6>, grid, 64u, in, 0LL, codes, 0LL, nbt, 1, 0LL);
}
template <int V>
void emul_acs(bool gring, unsigned grid, unsigned bd, VitGeom g) {
  if (gring) emul_launch(vit_acs_kernel<true, 0, V, 0>, grid, bd, g);
  else emul_launch(vit_acs_kernel<false, 0, V, 0>, grid, bd, g);
}
}  // namespace

extern "C" int emul_viterbi(const uint8_t *in, long long n_in, int rate, int m, int variant, int L, int W, int bd, int depth,
                            uint8_t *out, long long *n_out, unsigned *counters_out) {
  const int k = rate_k(rate), n = rate_n(rate), ntb = kNtb[rate];
  const long long nbt = n_in * m * k / (8 * n);
  const int O1 = (int)nbt - ntb;
  *n_out = O1 > 0 ? O1 : 0;
  if (O1 <= 0) return 0;
  std::vector<uint32_t> codes((size_t)nbt + 1);
  switch (rate) {
    case 0: emul_depuncture<0>(m, in, codes.data(), nbt); break;
    case 1: emul_depuncture<1>(m, in, codes.data(), nbt); break;
    case 2: emul_depuncture<2>(m, in, codes.data(), nbt); break;
    case 3: emul_depuncture<3>(m, in, codes.data(), nbt); break;
    default: emul_depuncture<4>(m, in, codes.data(), nbt); break;
  }
  if (depth <= 0 || depth > ntb) depth = ntb;
  const bool gring = depth < ntb;
  if (gring) L = (L + ntb - 1) / ntb * ntb;      // run_decode: chunk length a multiple of ntb with the split ring
  const int nchunks = (O1 + L - 1) / L;
  const int gfw = variant == 2 ? 32 : 16;
  const unsigned grid = (unsigned)((nchunks + bd - 1) / bd);
  std::vector<uint32_t> G((size_t)nchunks * gfw), F((size_t)nchunks * gfw), gr(gring ? (size_t)grid * ntb * 16 * bd : 1);
  const int nch16 = (nchunks + 15) / 16 * 16;
  std::vector<uint8_t> bad_store((size_t)nch16 * 5 + 16);
  uint8_t *bad_base = (uint8_t *)(((uintptr_t)bad_store.data() + 15) & ~(uintptr_t)15);
  unsigned counters[4] = {0, 0, 0, 0};
  VitGeom g;
  g.codes = codes.data(); g.codes_stride = nbt; g.out = out; g.out_stride = 0;
  g.G = G.data(); g.F = F.data(); g.prevF = nullptr; g.bad = bad_base; g.counters = counters;
  g.nstreams = 1; g.nchunks = nchunks; g.L = L; g.W = W; g.ntb = ntb; g.nbt = (int)nbt; g.O0 = 0; g.O1 = O1; g.reset_at_0 = 1;
  g.neg1 = 0xffffffffu; g.two = 2u; g.one = 1u; g.ring_depth = depth; g.gring = gring ? gr.data() : nullptr; g.gf_words = gfw;
  if ((size_t)(kLutWords + (ntb + 16 * depth) * bd) > sizeof(smem) / 4 || (size_t)(kLutWords + ntb * kRowWords * 32) > sizeof(smem) / 4) return -1;
  if (variant == 2) emul_acs<2>(gring, grid, (unsigned)bd, g);
  else if (variant == 1) emul_acs<1>(gring, grid, (unsigned)bd, g);
  else emul_acs<0>(gring, grid, (unsigned)bd, g);
  if (gfw == 32) emul_launch(vit_verify_kernel<32>, (unsigned)((nchunks + 63) / 64), 64u, g);
  else emul_launch(vit_verify_kernel<16>, (unsigned)((nchunks + 63) / 64), 64u, g);
  // the repair as run_decode() launches it: `rounds` parallel rounds (DVBT_EMUL_REPAIR_ROUNDS, default 2), then the sequential kernel
  {
    const char *e = getenv("DVBT_EMUL_REPAIR_ROUNDS");
    const int rounds = e ? atoi(e) : 2;
    uint8_t *f0 = bad_base, *b1 = f0 + nch16, *c1 = f0 + 2 * nch16, *b2 = f0 + 3 * nch16, *c2 = f0 + 4 * nch16;
    int left = -1;
    const uint8_t *bad_last = f0, *chg_last = nullptr;
    const unsigned rgrid = (unsigned)((nchunks + 31) / 32);
#define EMUL_ROUND(V) \
    if (rounds >= 1) { emul_launch(vit_repair_round_kernel<V>, rgrid, 32u, g, (const uint8_t *)f0, (const uint8_t *)nullptr, b1, c1, 2); bad_last = b1; chg_last = c1; left = 2; } \
    if (rounds >= 2) { emul_launch(vit_repair_round_kernel<V>, rgrid, 32u, g, (const uint8_t *)b1, (const uint8_t *)c1, b2, c2, 3); bad_last = b2; chg_last = c2; left = 3; } \
    emul_launch(vit_repair_kernel<V>, 1u, 32u, g, bad_last, chg_last, left);
    if (variant == 2) { EMUL_ROUND(2) } else if (variant == 1) { EMUL_ROUND(1) } else { EMUL_ROUND(0) }
#undef EMUL_ROUND
  }
  for (int i = 0; i < 4; i++) counters_out[i] = counters[i];
  return 0;
}
'''


def device_text():
    src = open(os.path.join(CSRC, "viterbi.cu")).read()
    a = src.index('#include "common.cuh"') + len('#include "common.cuh"')
    b = src.index("\nstruct dvbt_b200_viterbi {")
    text = src[a:b]
    # the banner of the host part trails the device part: cut after the namespace that holds the kernels
    text = text[: text.rindex("}  // namespace") + len("}  // namespace")] + "\n"
    text = text.replace("extern __shared__", "extern")
    for needle in ("vit_acs_kernel(VitGeom g)", "vit_repair_kernel(VitGeom g,", "vit_repair_round_kernel(VitGeom g,", "vit_verify_kernel(VitGeom g)", "vit_decode_range("):
        assert needle in text, "viterbi.cu changed shape: %r not in the extracted device part" % needle
    assert "cudaMalloc" not in text and "cudaStream" not in text
    return text


def build(force=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.startswith("viterbi")] + [os.path.join(HERE, "cuda_host_emul.h"), __file__]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    for h in ("viterbi_acs_gen.cuh", "viterbi_acs2_gen.cuh", "viterbi_acs_h16_gen.cuh", "viterbi_acs_h16b_gen.cuh"):
        t = open(os.path.join(CSRC, h)).read()
        if h == "viterbi_acs_gen.cuh":   # the two inline-PTX helpers, restated for the host
            t, n1 = re.subn(r'asm\("prmt\.b32 [^;]*;"[^;]*;', "r = emul_prmt(a, b, sel);", t)
            t, n2 = re.subn(r'asm\("mad\.lo\.u32 [^;]*;"[^;]*;', "r = a * b + c;", t)
            assert n1 == 1 and n2 == 1, (n1, n2)
        assert "asm(" not in t, h
        open(os.path.join(BUILD, h), "w").write(t)
    tu = ('// GENERATED by tests/emul/build_vit_emul.py from gr_dvbt_b200/csrc/viterbi.cu -- test infrastructure\n'
          '#include "../cuda_host_emul.h"\n'
          'namespace { alignas(16) uint32_t smem[1 << 17]; }   // the dynamic shared memory of the running block\n'
          + device_text() + LAUNCHER)
    path = os.path.join(BUILD, "vit_emul.cpp")
    open(path, "w").write(tu)
    cmd = ["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-Wno-unused", "-o", LIB, path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of the Viterbi device code failed:\n" + (r.stdout + r.stderr)[-6000:])
    return LIB


RX_LIB = os.path.join(BUILD, "librx_emul.so")

RX_LAUNCHER = r'''
// shift_bits: bit of row 0 that is bit 0 of the Viterbi input stream (a run that continues a stream starts inside a symbol)
extern "C" int emul_inner_codes_at(const uint8_t *dm, const int *out_src, const int *out_symidx, const short *H, const short *Hinv, int P, int m,
                                   int n_out, int rate, uint32_t *codes, int nbt, int shift_bits) {
  InnerMap im{dm, out_src, out_symidx, H, Hinv, P, m, n_out, shift_bits};
  const int G = kInnerTileCells / P;
  const unsigned grid = (unsigned)((n_out + G - 1) / G);
#define EMUL_INNER(R, M) emul_launch(rx_inner_codes_kernel<R, M>, grid, 256u, im, codes, nbt)
#define EMUL_RATE(R) (m == 2 ? EMUL_INNER(R, 2) : m == 4 ? EMUL_INNER(R, 4) : EMUL_INNER(R, 6))
  switch (rate) {
    case 0: EMUL_RATE(0); break;
    case 1: EMUL_RATE(1); break;
    case 2: EMUL_RATE(2); break;
    case 3: EMUL_RATE(3); break;
    default: EMUL_RATE(4); break;
  }
  return 0;
}

extern "C" int emul_inner_codes(const uint8_t *dm, const int *out_src, const int *out_symidx, const short *H, const short *Hinv, int P, int m,
                                int n_out, int rate, uint32_t *codes, int nbt) {
  return emul_inner_codes_at(dm, out_src, out_symidx, H, Hinv, P, m, n_out, rate, codes, nbt, 0);
}

// one call of the descrambler stage: plan (the block's NSYNC state machine over the pending packets) + descramble.
// pk_io: d_index in packets, carried from call to call; end: end of stream (the tail is flushed)
extern "C" int emul_descramble_stream(const uint8_t *rs, long long npk, const uint32_t *prbs, uint8_t *ts, long long ts_capacity, int grid,
                                      int end, int *pk_io, long long *first_packet, long long *ngroups, long long *items_used) {
  DescrState st;
  memset(&st, 0, sizeof st);
  st.pk = *pk_io;
  st.first_packet1 = 0;
  std::vector<int> plan((size_t)(npk / 16 + 2));
  std::vector<uint8_t> sb((size_t)npk + 16);
  if (npk > 0) emul_launch(rx_descr_syncbytes_kernel, (unsigned)((npk + 255) / 256), 256u, rs, npk, sb.data());
  emul_launch(rx_descr_plan_kernel, 1u, 1024u, (const uint8_t *)sb.data(), npk, &st, plan.data(), (long long)plan.size(), end);
  emul_launch(rx_descramble_kernel, (unsigned)grid, 256u, rs, (const DescrState *)&st, (const int *)plan.data(), prbs, ts, ts_capacity);
  *pk_io = st.pk;
  *first_packet = st.first_packet1 - 1;
  *ngroups = 2LL * st.n_pairs + st.n_tail;
  *items_used = st.items_used;
  return 0;
}

extern "C" int emul_descramble(const uint8_t *rs, long long npk, const uint32_t *prbs, uint8_t *ts, long long ts_capacity, int grid,
                               int *p0, long long *ngroups) {
  int pk = 0;
  long long first = -1, used = 0;
  emul_descramble_stream(rs, npk, prbs, ts, ts_capacity, grid, 1, &pk, &first, ngroups, &used);
  *p0 = (int)first;
  return 0;
}
'''


def rx_device_text():
    src = open(os.path.join(CSRC, "rx_chain.cu")).read()
    a = src.index("struct InnerMap {")
    b = src.index("\nstruct dvbt_b200_rx {")
    text = src[a:b]
    text = text[: text.rindex("}  // namespace")]
    for needle in ("rx_inner_codes_kernel(InnerMap im", "rx_descramble_kernel(", "rx_descr_plan_kernel(", "struct DescrState"):
        assert needle in text, "rx_chain.cu changed shape: %r not in the extracted device part" % needle
    assert "cudaMalloc" not in text and "cudaStream" not in text
    return text.replace("extern __shared__", "extern")


def build_rx(force=False):
    """tests/emul/_build/librx_emul.so: rx_inner_codes_kernel and rx_descramble_kernel of rx_chain.cu for the host"""
    deps = [os.path.join(CSRC, "rx_chain.cu"), os.path.join(HERE, "cuda_host_emul.h"), __file__]
    if not force and os.path.exists(RX_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(RX_LIB) for d in deps):
        return RX_LIB
    os.makedirs(BUILD, exist_ok=True)
    tu = ('// GENERATED by tests/emul/build_vit_emul.py from gr_dvbt_b200/csrc/rx_chain.cu -- test infrastructure\n'
          '#include <string.h>\n#include <vector>\n#include "../cuda_host_emul.h"\n'
          'namespace {\nalignas(16) uint8_t s_bit[1 << 17];   // the dynamic shared memory of the running block\n'
          + rx_device_text() + RX_LAUNCHER + "}  // namespace\n")
    # the launchers are extern "C": they must sit outside the unnamed namespace
    tu = tu.replace(RX_LAUNCHER + "}  // namespace\n", "}  // namespace\n" + RX_LAUNCHER)
    path = os.path.join(BUILD, "rx_emul.cpp")
    open(path, "w").write(tu)
    cmd = ["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-Wno-unused", "-o", RX_LIB, path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of the rx_chain device code failed:\n" + (r.stdout + r.stderr)[-6000:])
    return RX_LIB


RS_LIB = os.path.join(BUILD, "librs_emul.so")

RS_LAUNCHER = r'''
extern "C" int emul_rs(const uint8_t *in, uint8_t *out, int *status, long long npackets, int as_built, long long gather_stream_bytes) {
  if (npackets <= 0) return 0;
  if (dvbt::rs_upload_tables()) return -1;
  const unsigned grid = (unsigned)((npackets + kTilePk - 1) / kTilePk);
  if (gather_stream_bytes >= 0) emul_launch(rs_decode_kernel<true>, grid, (unsigned)kTilePk, in, out, status, npackets, as_built, gather_stream_bytes, 0LL);
  else emul_launch(rs_decode_kernel<false>, grid, (unsigned)kTilePk, in, out, status, npackets, as_built, 0LL, 0LL);
  return 0;
}
'''


def build_rs(force=False):
    """tests/emul/_build/librs_emul.so: rs_decode_kernel (both load paths) and the table construction of rs.cu for the host"""
    deps = [os.path.join(CSRC, "rs.cu"), os.path.join(HERE, "cuda_host_emul.h"), __file__]
    if not force and os.path.exists(RS_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(RS_LIB) for d in deps):
        return RS_LIB
    os.makedirs(BUILD, exist_ok=True)
    src = open(os.path.join(CSRC, "rs.cu")).read()
    a = src.index("namespace {\n\nconstexpr int kN = 255")
    b = src.index("\nstruct dvbt_b200_rsdec {")
    kernels = src[a:b]
    kernels = kernels[: kernels.rindex("}  // namespace") + len("}  // namespace")] + "\n"
    c = src.index("int rs_upload_tables() {")
    tables = src[c: src.index("\n}\n", c) + 3]
    assert "rs_decode_kernel(" in kernels and "rs_warp_decode(" in kernels and "cudaMemcpyToSymbol(d_lfsr" in tables
    tu = ('// GENERATED by tests/emul/build_vit_emul.py from gr_dvbt_b200/csrc/rs.cu -- test infrastructure\n'
          '#include "../cuda_host_emul.h"\n'
          '#define DVBT_CUDA_TRY(x) (void)(x)\n'
          '#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void *)&(sym), (src), (n)), 0)\n'
          'static inline int cudaGetDevice(int *d) { *d = 0; return 0; }\n'
          'namespace { alignas(16) uint8_t s_dyn[1 << 18]; }   // the dynamic shared memory of the running block\n'
          + kernels.replace("extern __shared__", "extern") + "namespace dvbt {\n" + tables + "}  // namespace dvbt\n" + RS_LAUNCHER)
    path = os.path.join(BUILD, "rs_emul.cpp")
    open(path, "w").write(tu)
    cmd = ["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-Wno-unused", "-o", RS_LIB, path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of the RS device code failed:\n" + (r.stdout + r.stderr)[-6000:])
    return RS_LIB


DEMAP_LIB = os.path.join(BUILD, "libdemap_emul.so")

DEMAP_SHIM = r'''
#include <math.h>
#include "../../../include/dvbt_b200.h"
struct float2 { float x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
// round-to-nearest single operations: plain operators (the TU is built with -ffp-contract=off, SSE arithmetic)
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float2int_rn(float f) { return (int)lrintf(f); }   // default rounding mode: to nearest even
#define __grid_constant__
'''

DEMAP_LAUNCHER = r'''
extern "C" int emul_demap(const float *in, long long ncells, int constellation, int hierarchy, float gain, uint8_t *out, float *pts_out) {
  dvbt::DemapTable t;
  if (dvbt::make_demap_table(constellation, hierarchy, gain, &t)) return -1;
  for (int i = 0; i < t.size; i++) { pts_out[2 * i] = t.pts[i].x; pts_out[2 * i + 1] = t.pts[i].y; }
  const long long threads = (ncells + 3) / 4;
  emul_launch(dvbt::demap_kernel, (unsigned)((threads + 63) / 64), 64u, (const float2 *)in, out, ncells, t);
  return t.near_ok;
}
'''


def build_demap(force=False):
    """tests/emul/_build/libdemap_emul.so: make_demap_table + demap_kernel (demap.cu) with the cell decision of demod.cuh"""
    deps = [os.path.join(CSRC, "demap.cu"), os.path.join(CSRC, "demod.cuh"), os.path.join(HERE, "cuda_host_emul.h"), __file__]
    if not force and os.path.exists(DEMAP_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(DEMAP_LIB) for d in deps):
        return DEMAP_LIB
    os.makedirs(BUILD, exist_ok=True)
    hdr = open(os.path.join(CSRC, "demod.cuh")).read()
    a = hdr.index("namespace dvbt {")
    b = hdr.index("#endif", hdr.index("#ifdef __CUDACC__"))
    cells = hdr[a:b].replace("#ifdef __CUDACC__", "") + "}  // namespace dvbt\n"
    src = open(os.path.join(CSRC, "demap.cu")).read()
    c = src.index("namespace dvbt {")
    d = src.index("int demap_launch(")
    kern = src[c:d] + "}  // namespace dvbt\n"
    assert "demap_cell_exact" in cells and "demap_cell_near" in cells and "demap_kernel(" in kern and "make_demap_table(" in kern
    tu = ('// GENERATED by tests/emul/build_vit_emul.py from gr_dvbt_b200/csrc/demap.cu and demod.cuh -- test infrastructure\n'
          '#include "../cuda_host_emul.h"\n' + DEMAP_SHIM + cells + kern + DEMAP_LAUNCHER)
    path = os.path.join(BUILD, "demap_emul.cpp")
    open(path, "w").write(tu)
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-Wno-unused", "-o", DEMAP_LIB, path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of the demap device code failed:\n" + (r.stdout + r.stderr)[-6000:])
    return DEMAP_LIB


DEMOD_LIB = os.path.join(BUILD, "libdemod_emul.so")

DEMOD_SHIM = r'''
#include <new>
#include <vector>
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
enum { cudaMemcpyHostToDevice = 1 };
static inline int cudaMemcpy(void *d, const void *s, size_t n, int) { memcpy(d, s, n); return 0; }
#define DVBT_CUDA_TRY(x) (void)(x)
namespace dvbt {
static inline void set_error(const char *, ...) {}
struct DevBuf {            // "device" memory is host memory here
  void *p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) { if (bytes > cap) { free(p); p = calloc(bytes + 64, 1); cap = bytes; } return p ? 0 : -1; }
  void release() { free(p); p = nullptr; cap = 0; }
  template <class T> T *as() const { return (T *)p; }
};
}  // namespace dvbt
'''

DEMOD_LAUNCHER = r'''
// mirrors demod_run() of demod.cu (the four launches in order) on host buffers, from a fresh receiver state
extern "C" int emul_demod(const float *Xf, int nsym, int constellation, int tm, int fi_start, int sync_start_at0, float *Y_out, uint8_t *dm_out,
                          int *symidx_out, int *src_out, int *n_out, int *first_out, int *sf_tag_at) {
  using namespace dvbt;
  ModeTables tabs;
  if (tabs.init(tm, DVBT_G1_32)) return -1;
  const ModeDev &md = tabs.dev;
  DemapTable dt;
  if (make_demap_table(constellation, DVBT_NH, 1.0f, &dt)) return -2;
  const int nparse = nsym - 1;
  if (nparse <= 0) return -3;
  const float2 *X = (const float2 *)Xf;
  std::vector<int> fo(nparse), mod(nparse), vote(nparse), osym(nparse), osrc(nparse);
  std::vector<float2> rot(nparse), tps((size_t)nparse * md.ntps);
  DemodState st;
  memset(&st, 0, sizeof st);
  const char *fused = getenv("DVBT_EMUL_DEMOD_FUSED");
  if (fused && atoi(fused)) {
    emul_launch(demod_symbol_kernel<true, 384>, (unsigned)nparse, 384u, md, dt, 1, (int)((((uintptr_t)X) & 15) == 0), X, fo.data(), rot.data(), mod.data(),
                tps.data(), (float2 *)Y_out, dm_out, (uint32_t *)nullptr);
  } else {
    // as demod_run() launches them with a side stream: stage 1 equalises the TPS carriers, the symbol kernel the payload only
    emul_launch(demod_stage1_kernel, (unsigned)(((long long)nparse * 32 + 127) / 128), 128u, md, X, nparse, fo.data(), rot.data(), mod.data(), tps.data());
    emul_launch(demod_symbol_kernel<false, 192>, (unsigned)nparse, 192u, md, dt, 1, (int)((((uintptr_t)X) & 15) == 0), X, fo.data(), rot.data(), mod.data(),
                (float2 *)nullptr, (float2 *)Y_out, dm_out, (uint32_t *)nullptr);
  }
  emul_launch(demod_vote_kernel, (unsigned)((nparse + 127) / 128), 128u, md.ntps, nparse, (const float2 *)tps.data(), &st, vote.data(),
              sync_start_at0, (const int *)nullptr, 0);
  emul_launch(demod_scan_kernel, 1u, (unsigned)(32 * kScanWarps), md.ntps, nparse, fi_start, 0, (const int *)mod.data(), (const int *)vote.data(),
              (const float2 *)tps.data(), &st, osym.data(), osrc.data());
  *n_out = st.n_out; *first_out = st.first_out; *sf_tag_at = st.sf_tag_at;
  for (int i = 0; i < st.n_out; i++) { symidx_out[i] = osym[i]; src_out[i] = osrc[i]; }
  tabs.release();
  return 0;
}
'''


# bulk_copy.cuh for the host: the bulk asynchronous copy is a memcpy that has landed when the call returns, the mbarrier
# is a no-op (the block barrier that follows in every user orders the threads)
BULK_COPY_HOST = r'''
#include <string.h>
namespace dvbt {
static inline void mbar_init(uint64_t *bar, unsigned) { *bar = 0; }
static inline void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, uint64_t *) { memcpy(smem_dst, gmem_src, bytes); }
static inline void mbar_wait(uint64_t *, unsigned) {}
}  // namespace dvbt
'''


def build_demod(force=False):
    """tests/emul/_build/libdemod_emul.so: ModeTables::init and the four demod_reference_signals kernels of demod.cu"""
    deps = [os.path.join(CSRC, "demod.cu"), os.path.join(CSRC, "demap.cu"), os.path.join(CSRC, "demod.cuh"), os.path.join(HERE, "cuda_host_emul.h"), __file__]
    if not force and os.path.exists(DEMOD_LIB) and all(os.path.getmtime(d) <= os.path.getmtime(DEMOD_LIB) for d in deps):
        return DEMOD_LIB
    os.makedirs(BUILD, exist_ok=True)
    hdr = open(os.path.join(CSRC, "demod.cuh")).read()
    hdr = hdr[hdr.index("namespace dvbt {"):].replace("#ifdef __CUDACC__", "").replace("#endif", "")
    dm = open(os.path.join(CSRC, "demap.cu")).read()
    dm = dm[dm.index("namespace dvbt {"): dm.index("__device__ __forceinline__ uint8_t demap_cell(")] + "}  // namespace dvbt\n"
    src = open(os.path.join(CSRC, "demod.cu")).read()
    body = src[src.index("namespace dvbt {"): src.index("int demod_run(")] + "}  // namespace dvbt\n"
    for needle in ("demod_stage1_kernel(", "demod_symbol_kernel(", "demod_vote_kernel(", "demod_scan_kernel(", "ModeTables::init("):
        assert needle in body, "demod.cu changed shape: %r" % needle
    tu = ('// GENERATED by tests/emul/build_vit_emul.py from gr_dvbt_b200/csrc/demod.cu, demod.cuh, demap.cu -- test infrastructure\n'
          '#include "../cuda_host_emul.h"\n' + BULK_COPY_HOST + DEMAP_SHIM + DEMOD_SHIM + hdr + dm
          + 'namespace dvbt { alignas(16) unsigned char s_sym[1 << 18]; }   // the dynamic shared memory of the running block\n'
          + body.replace("extern __shared__", "extern") + DEMOD_LAUNCHER)
    path = os.path.join(BUILD, "demod_emul.cpp")
    open(path, "w").write(tu)
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-D_GNU_SOURCE", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-Wno-unused",
           "-o", DEMOD_LIB, path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("host build of the demod device code failed:\n" + (r.stdout + r.stderr)[-8000:])
    return DEMOD_LIB


# ---------------------------------------------------------------------------------------------------------------
# whole-file builds: a .cu file with its host orchestration, on the stand-in runtime of tests/emul/fake_cuda/
# ---------------------------------------------------------------------------------------------------------------
def rewrite_cuda(text):
    """`kernel<<<grid, block[, smem[, stream]]>>>(args)` -> `emul_launch(kernel, grid, block, args)`;
    `extern __shared__ T name[];` -> `T *name = (T *)emul_dyn_smem;`"""
    out, i = [], 0
    while True:
        j = text.find("<<<", i)
        if j < 0:
            out.append(text[i:])
            break
        k = j                                  # kernel name (with template arguments) ends at j
        depth = 0
        while k > 0:
            c = text[k - 1]
            if c == ">":
                depth += 1
            elif c == "<":
                depth -= 1
            elif depth == 0 and not (c.isalnum() or c in "_:"):
                break
            k -= 1
        name = text[k:j]
        e = text.index(">>>", j)
        cfg, parts, depth, cur = text[j + 3: e], [], 0, ""
        for c in cfg:
            if c in "([":
                depth += 1
            elif c in ")]":
                depth -= 1
            if c == "," and depth == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += c
        parts.append(cur)
        assert text[e + 3] == "(", text[j - 40: e + 10]
        close = text[e + 4:].lstrip().startswith(")")
        out.append(text[i:k])
        out.append("emul_launch(%s, %s, (unsigned)(%s)%s" % (name, parts[0].strip(), parts[1].strip(), "" if close else ", "))
        i = e + 4
    text = "".join(out)
    text = re.sub(r"extern __shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];", r"\1 *\2 = (\1 *)emul_dyn_smem;", text)
    assert "<<<" not in text and "extern __shared__" not in text
    return text


def build_whole(libname, files, force=False, extra_flags=()):
    lib = os.path.join(BUILD, libname)
    srcs = [os.path.join(CSRC, f) for f in files]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [
        os.path.join(HERE, "cuda_host_emul.h"), os.path.join(HERE, "fake_cuda", "cuda_runtime.h"), os.path.join(HERE, "fake_cuda", "cufft.h"), __file__]
    if not force and os.path.exists(lib) and all(os.path.getmtime(d) <= os.path.getmtime(lib) for d in deps):
        return lib
    wdir = os.path.join(BUILD, "whole")
    os.makedirs(wdir, exist_ok=True)
    # headers: copies next to the generated sources (a quoted #include looks there first), patched where they hold inline PTX
    # or hide device functions from a host compiler
    for h in os.listdir(CSRC):
        if not h.endswith(".cuh"):
            continue
        t = open(os.path.join(CSRC, h)).read()
        if h == "bulk_copy.cuh":   # inline PTX only: the host version (a synchronous copy is one of its legal executions)
            open(os.path.join(wdir, h), "w").write("#pragma once\n#include <stdint.h>\n" + BULK_COPY_HOST)
            continue
        if h == "viterbi_acs_gen.cuh":
            t, n1 = re.subn(r'asm\("prmt\.b32 [^;]*;"[^;]*;', "r = emul_prmt(a, b, sel);", t)
            t, n2 = re.subn(r'asm\("mad\.lo\.u32 [^;]*;"[^;]*;', "r = a * b + c;", t)
            assert n1 == 1 and n2 == 1
        if h == "demod.cuh":
            t = t.replace("#ifdef __CUDACC__", "#if 1")
        assert "asm(" not in t and "asm volatile" not in t, h
        open(os.path.join(wdir, h), "w").write(rewrite_cuda(t))
    cpps = []
    for f in srcs:
        out = os.path.join(wdir, os.path.basename(f).replace(".cu", "_emul.cpp"))
        t = open(f).read()
        if f.endswith("resample.cu"):   # cp.async staging: a synchronous copy is one of its legal executions
            t, n1 = re.subn(r'asm volatile\("cp\.async\.ca\.shared\.global [^;]*;"[^;]*;', "dst[i] = in ? *src : make_float2(0.f, 0.f);", t)
            t, n2 = re.subn(r'asm volatile\("cp\.async\.(?:commit_group|wait_group \d)+;" ::: "memory"\);', "(void)0;", t)
            assert n1 == 1 and n2 == 3, (n1, n2)
        assert "asm" not in re.sub(r"//.*", "", t), f
        open(out, "w").write("// GENERATED by tests/emul/build_vit_emul.py from %s -- test infrastructure\n" % os.path.relpath(f, ROOT) + rewrite_cuda(t))
        cpps.append(out)
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-D_GNU_SOURCE", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-Wno-unused",
           "-Wno-attributes", "-I", os.path.join(HERE, "fake_cuda"), "-I", CSRC, "-o", lib] + list(extra_flags) + cpps
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("whole-file host build failed:\n" + (r.stdout + r.stderr)[-8000:])
    return lib


def build_acq(force=False):
    """tests/emul/_build/libacq_emul.so: acq.cu (ofdm_sym_acquisition: kernels AND host orchestration) + common.cu, exporting
    the same C ABI (dvbt_b200_acq_*) on the stand-in runtime"""
    return build_whole("libacq_emul.so", ["acq.cu", "common.cu"], force)


def build_all(force=False):
    """tests/emul/_build/libdvbt_b200_emul.so: EVERY source file of the library (kernels and host code) on the stand-in
    runtime - the whole C ABI of include/dvbt_b200.h, running on the CPU.  Loaded by tests only, by explicit path."""
    return build_whole("libdvbt_b200_emul.so", sorted(f for f in os.listdir(CSRC) if f.endswith(".cu")), force,
                       extra_flags=("-DDVBT_B200_LEGACY_ACS",))   # the round-1 byte-SWAR schedule stays covered here


def build_all_asan(force=False):
    """the same with AddressSanitizer and exact-size "device" allocations: a memcheck of kernels and host code.
    Run with  LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 DVBT_EMUL_ASAN=1 python -m pytest
    tests -m gpu --emulated-library  (see tests/conftest.py)"""
    return build_whole("libdvbt_b200_emul_asan.so", sorted(f for f in os.listdir(CSRC) if f.endswith(".cu")), force,
                       ("-fsanitize=address", "-fno-omit-frame-pointer", "-g", "-DDVBT_B200_EXACT_ALLOC"))


def build_all_tsan(force=False):
    """the same with ThreadSanitizer: CUDA threads are host threads and __syncthreads / warp primitives are barriers, so a
    missing barrier between a shared-memory write and a read by another thread is a data race TSan can see (a racecheck).
    Run with LD_PRELOAD=$(gcc -print-file-name=libtsan.so) DVBT_EMUL_TSAN=1 ... --emulated-library"""
    return build_whole("libdvbt_b200_emul_tsan.so", sorted(f for f in os.listdir(CSRC) if f.endswith(".cu")), force,
                       ("-fsanitize=thread", "-fno-omit-frame-pointer", "-g"))


if __name__ == "__main__":
    print(build_all(force=True))
    print(build_acq(force=True))
    print(build_demod(force=True))
    print(build(force=True))
    print(build_rx(force=True))
    print(build_rs(force=True))
    print(build_demap(force=True))
