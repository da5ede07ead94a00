"""TEST INFRASTRUCTURE ONLY — ctypes driver for oracle/_ref/libdvbt_ref.so.

The library is the reference's own block sources compiled verbatim (oracle/Makefile);
this module plays the part of the GNU Radio scheduler for them: one block object, one
general_work() call at a time, item counters and tags carried by hand, exactly the
recipe of SURVEY.md Appendix B.  It is used to
  * generate the golden fixtures under tests/golden/ (tests/golden/make_golden.py),
  * pin the C restatement in oracle/port against the real reference, and
  * time the reference CPU path (bench.py --impl reference / cpu_baseline).
Nothing in gr_dvbt_b200/ imports it.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libdvbt_ref.so")
REF_RSFIX_SO = os.path.join(HERE, "_ref", "libdvbt_ref_rsfix.so")
# the B200 gr::block shims behind the same harness entry points (gr_dvbt_b200/shim/shim_harness.cc):
# names listed in SHIM_BLOCKS are created from it instead of from the reference build, so a chain can
# be run with the hot blocks swapped and every other block still the reference's
SHIM_SO = os.path.join(os.path.dirname(HERE), "gr_dvbt_b200", "shim", "libdvbt_b200_shim_test.so")
SHIM_BLOCKS = set()
HOT_BLOCKS = ("ofdm_sym_acquisition", "demod_reference_signals", "dvbt_demap", "viterbi_decoder", "reed_solomon_dec")

# enums of include/dvbt/dvbt_config.h:34-76 (values are the TPS codes)
QPSK, QAM16, QAM64 = 0, 1, 2
NH = 0
C1_2, C2_3, C3_4, C5_6, C7_8 = 0, 1, 2, 3, 4
T2k, T8k = 0, 1
G1_32, G1_16, G1_8, G1_4 = 0, 1, 2, 3

RATE_KN = {C1_2: (1, 2), C2_3: (2, 3), C3_4: (3, 4), C5_6: (5, 6), C7_8: (7, 8)}
BITS_PER_CELL = {QPSK: 2, QAM16: 4, QAM64: 6}
NTRACEBACK = {C1_2: 5, C2_3: 9, C3_4: 10, C5_6: 15, C7_8: 24}  # viterbi_decoder_impl.cc:95-124


def available(fixed_rs=False):
    return os.path.exists(REF_RSFIX_SO if fixed_rs else REF_SO)


_libs = {}


def shim_available():
    return os.path.exists(SHIM_SO)


def _lib(fixed_rs=False, shim=False):
    path = SHIM_SO if shim else (REF_RSFIX_SO if fixed_rs else REF_SO)
    if path not in _libs:
        # RTLD_LOCAL: the two variants define the same symbols
        lib = C.CDLL(path, mode=os.RTLD_LOCAL | os.RTLD_NOW)
        lib.dvbt_ref_create.restype = C.c_void_p
        lib.dvbt_ref_create.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_int]
        lib.dvbt_ref_destroy.argtypes = [C.c_void_p]
        lib.dvbt_ref_add_in_tag.argtypes = [C.c_void_p, C.c_ulonglong, C.c_char_p, C.c_long]
        lib.dvbt_ref_clear_tags.argtypes = [C.c_void_p, C.c_int, C.c_int]
        lib.dvbt_ref_num_out_tags.argtypes = [C.c_void_p]
        lib.dvbt_ref_get_out_tag.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_ulonglong), C.c_char_p, C.c_int, C.POINTER(C.c_long)]
        lib.dvbt_ref_nitems_read.restype = C.c_ulonglong
        lib.dvbt_ref_nitems_read.argtypes = [C.c_void_p]
        lib.dvbt_ref_nitems_written.restype = C.c_ulonglong
        lib.dvbt_ref_nitems_written.argtypes = [C.c_void_p]
        lib.dvbt_ref_forecast.argtypes = [C.c_void_p, C.c_int]
        lib.dvbt_ref_general_work.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        _libs[path] = lib
    return _libs[path]


def _quiet(fn, *a):
    """The reference constructors printf their parameters; keep test logs readable."""
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        return fn(*a)
    finally:
        C.CDLL(None).fflush(None)
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)


class RefBlock:
    """One reference block instance; args follow its make() signature."""

    def __init__(self, name, *args, fixed_rs=False, quiet=True):
        self.lib = _lib(fixed_rs, shim=name in SHIM_BLOCKS)
        self.name = name
        self.quiet = quiet
        arr = (C.c_double * max(1, len(args)))(*[float(a) for a in args])
        mk = lambda: self.lib.dvbt_ref_create(name.encode(), arr, len(args))
        self.h = _quiet(mk) if quiet else mk()
        if not self.h:
            raise ValueError("unknown reference block " + name)

    def close(self):
        if self.h:
            self.lib.dvbt_ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_tag(self, offset, key, value=0):
        self.lib.dvbt_ref_add_in_tag(self.h, int(offset), key.encode(), int(value))

    def clear_tags(self, in_tags=True, out_tags=True):
        self.lib.dvbt_ref_clear_tags(self.h, int(in_tags), int(out_tags))

    def out_tags(self):
        res = []
        n = self.lib.dvbt_ref_num_out_tags(self.h)
        off = C.c_ulonglong()
        val = C.c_long()
        key = C.create_string_buffer(64)
        for i in range(n):
            self.lib.dvbt_ref_get_out_tag(self.h, i, C.byref(off), key, 64, C.byref(val))
            res.append((int(off.value), key.value.decode(), int(val.value)))
        return res

    @property
    def nread(self):
        return int(self.lib.dvbt_ref_nitems_read(self.h))

    @property
    def nwritten(self):
        return int(self.lib.dvbt_ref_nitems_written(self.h))

    def forecast(self, noutput):
        return int(self.lib.dvbt_ref_forecast(self.h, int(noutput)))

    def work(self, noutput, ninput_items, inp, out, nports=1, inp1=None, out1=None):
        """inp/out: numpy arrays (contiguous) or raw addresses. Returns (produced, consumed)."""
        def addr(x):
            if x is None:
                return None
            if isinstance(x, np.ndarray):
                return x.ctypes.data
            return int(x)
        cons = C.c_int(0)
        fn = lambda: self.lib.dvbt_ref_general_work(self.h, int(noutput), int(ninput_items), int(nports), addr(inp), addr(inp1), addr(out), addr(out1), C.byref(cons))
        r = _quiet(fn) if self.quiet else fn()
        return int(r), int(cons.value)


# ----------------------------------------------------------------------------------------
# Stage helpers (whole arrays in, whole arrays out), SURVEY Appendix B recipe
# ----------------------------------------------------------------------------------------

def mode_dims(tm, gi=G1_32):
    """(N, P, K, cp): cp = N/32, N/16, N/8, N/4 for the guard intervals G1_32 .. G1_4 (lib/dvbt_config.cc:194-208)"""
    N, P, K = (2048, 1512, 1705) if tm == T2k else (8192, 6048, 6817)
    return N, P, K, N // (32 >> gi)


def tx_outer(ts_bytes):
    """TS packets -> energy dispersal -> RS(204,188) -> Forney interleaver.
    energy_dispersal_impl.cc:93-140, reed_solomon_enc_impl.cc, convolutional_interleaver_impl.cc."""
    ts = np.ascontiguousarray(ts_bytes, dtype=np.uint8)
    ng = len(ts) // (8 * 188)
    ed = np.zeros(ng * 1504 + 16, np.uint8)
    b = RefBlock("energy_dispersal", 1)
    b.work(ng, 0, ts, ed)
    rs = np.zeros(ng * 1632, np.uint8)
    b = RefBlock("reed_solomon_enc", 2, 8, 0x11D, 255, 239, 8, 51, 8)
    b.work(ng, ng, ed, rs)
    ci = np.zeros(ng * 1632, np.uint8)
    b = RefBlock("convolutional_interleaver", 136, 12, 17)
    b.work(ng * 1632, ng * 1632 // 12, rs, ci)
    return ed[: ng * 1504], rs, ci


def tx_inner(ci, con, cr, tm, nsym=None, gi=G1_32):
    """Forney-interleaved bytes -> inner coder -> bit/symbol interleave -> map -> pilots.
    Returns dict with every intermediate; X is (nsym, N) complex64 frequency-domain symbols."""
    N, P, _, _ = mode_dims(tm)
    k, n = RATE_KN[cr]
    m = BITS_PER_CELL[con]
    per_item = P * k * m // (8 * n)
    if nsym is None:
        nsym = (len(ci) // per_item) // 4 * 4
    nsym = nsym // 4 * 4
    ic = np.zeros(nsym * P, np.uint8)
    RefBlock("inner_coder", 1, P, con, NH, cr).work(nsym, 0, np.ascontiguousarray(ci), ic)
    bi = np.zeros(nsym * P, np.uint8)
    RefBlock("bit_inner_interleaver", P, con, NH, tm).work(nsym, nsym, ic, bi, nports=2, inp1=ic, out1=bi)
    si = np.zeros(nsym * P, np.uint8)
    RefBlock("symbol_inner_interleaver", P, tm, 1).work(nsym, nsym, bi, si)
    ma = np.zeros(nsym * P, np.complex64)
    RefBlock("dvbt_map", P, con, NH, tm, 1.0).work(nsym, nsym, si, ma)
    X = np.zeros((nsym + 1) * N + 64, np.complex64)
    RefBlock("reference_signals", 8, P, N, con, NH, cr, cr, gi, tm, 0, 0).work(nsym, nsym, ma, X[32:])
    X = X[32 : 32 + nsym * N].reshape(nsym, N).copy()
    return dict(ic=ic, bi=bi, si=si, ma=ma, X=X, nsym=nsym, per_item=per_item)


def rx_demod(Xf, con, cr, tm, sync_offsets=(0,), gi=G1_32):
    """demod_reference_signals one item per call, two items visible, sync_start tags at the item
    offsets `sync_offsets` (what ofdm_sym_acquisition sends on every acquisition attempt; a tag on the
    item being parsed re-arms the wait for a superframe start, demod_reference_signals_impl.cc:96-150).
    Xf: (nsym, N) complex64 post-FFT symbols.  Returns (Y (nout,P) complex64, tags)."""
    N, P, _, _ = mode_dims(tm)
    nsym = Xf.shape[0]
    buf = np.zeros((nsym + 1) * N + 64, np.complex64)
    buf[32 : 32 + nsym * N] = Xf.reshape(-1)
    Y = np.zeros(nsym * P, np.complex64)
    b = RefBlock("demod_reference_signals", 8, N, P, con, NH, cr, cr, gi, tm, 0, 0)
    for off in sorted(set(int(o) for o in sync_offsets)):
        b.add_tag(off, "sync_start", 1)
    nout = 0
    base = buf.ctypes.data + 32 * 8
    for i in range(nsym - 1):
        r, _ = b.work(1, 2, base + i * N * 8, Y.ctypes.data + nout * P * 8)
        if r > 0:
            nout += r
    return Y[: nout * P].reshape(nout, P).copy(), b.out_tags()


def rx_demap(Y, con, tm):
    nout, P = Y.shape
    dm = np.zeros(nout * P, np.uint8)
    RefBlock("dvbt_demap", P, con, NH, tm, 1.0).work(nout, nout, np.ascontiguousarray(Y), dm)
    return dm.reshape(nout, P)


def rx_deinterleave(dm, tags, con, tm):
    nout, P = dm.shape
    sd = np.zeros(nout * P, np.uint8)
    b = RefBlock("symbol_inner_interleaver", P, tm, 0)
    for off, key, val in tags:
        if key == "symbol_index":
            b.add_tag(off, key, val)
    b.work(nout, nout, np.ascontiguousarray(dm), sd)
    bd = np.zeros(nout * P, np.uint8)
    bd2 = np.zeros(nout * P, np.uint8)
    RefBlock("bit_inner_deinterleaver", P, con, NH, tm).work(nout, nout, sd, bd, nports=2, inp1=sd, out1=bd2)
    return sd.reshape(nout, P), bd.reshape(nout, P)


def rx_viterbi(vin, con, cr, sf_tag_offset=None, blocks_per_call=16, bsize=768):
    """viterbi_decoder in calls of j*768*k/8 output bytes (viterbi_decoder_impl.cc:191-324).
    NOTE: process-global decoder state in the reference: one live instance at a time."""
    k, n = RATE_KN[cr]
    m = BITS_PER_CELL[con]
    vin = np.ascontiguousarray(vin, np.uint8).reshape(-1)
    om = bsize * k // 8
    nsymb = bsize * n // m
    vo = np.zeros(len(vin) * k * m // (8 * n) + 4096, np.uint8)
    b = RefBlock("viterbi_decoder", con, NH, cr, bsize, 0, -1)
    if sf_tag_offset is not None:
        # one offset or several (a receiver that re-synchronised mid-stream): with more than one tag the result
        # depends on the call size - everything between the start of a call's window and a tag inside it is
        # dropped (:213-229) - so multi-tag streams are driven one 768-block per call (blocks_per_call=1), the
        # smallest call the scheduler can make (set_output_multiple, :141)
        for off in ([sf_tag_offset] if np.isscalar(sf_tag_offset) else sf_tag_offset):
            b.add_tag(int(off), "superframe_start", 0xAA)
    vout = 0
    while True:
        avail = len(vin) - b.nread
        nb = min(blocks_per_call, avail // nsymb)
        if nb < 1:
            break
        r, cons = b.work(nb * om, avail, vin.ctypes.data + b.nread, vo.ctypes.data + vout)
        if r > 0:
            vout += r
        if r == 0 and cons == 0:
            break
    tags = b.out_tags()
    b.close()
    return vo[:vout].copy(), tags


def rx_outer(vo, vtags, fixed_rs=False, min_calls=False):
    """convolutional_deinterleaver -> reed_solomon_dec -> energy_descramble.
    min_calls: the smallest calls the scheduler can make - 2 items for the deinterleaver (set_output_multiple(2),
    convolutional_deinterleaver_impl.cc:61), 4 x 1504 output bytes for the descrambler (energy_descramble_impl.cc:84-87).
    With a single superframe_start at offset 0 and an undisturbed stream the call sizes do not matter; after a
    mid-stream re-synchronisation they do (bytes in front of a tag inside a call's window are dropped, and the
    descrambler re-checks NSYNC once per call), so those cases are pinned at the minimal call size."""
    vo = np.ascontiguousarray(vo, np.uint8)
    cd = np.zeros(len(vo) + 4096, np.uint8)
    b = RefBlock("convolutional_deinterleaver", 136, 12, 17)
    for off, key, val in vtags:
        b.add_tag(off, key, val)
    cdo = 0
    while True:
        avail = len(vo) - b.nread
        ni = min(2 if min_calls else 64, avail // 1632) // 2 * 2
        if ni < 2:
            break
        r, cons = b.work(ni, avail, vo.ctypes.data + b.nread, cd.ctypes.data + cdo * 1632)
        if r > 0:
            cdo += r
        if r == 0 and cons == 0:
            break
    cd = cd[: cdo * 1632]
    rd = np.zeros(cdo * 1504 + 16, np.uint8)
    RefBlock("reed_solomon_dec", 2, 8, 0x11D, 255, 239, 8, 51, 8, fixed_rs=fixed_rs).work(cdo, cdo, cd, rd)
    rd = rd[: cdo * 1504]
    out = np.zeros(cdo * 1504 + 16, np.uint8)
    b = RefBlock("energy_descramble", 8)
    oo = 0
    while True:
        avail = cdo - b.nread
        if avail < 4:
            break
        nitems = 4 if min_calls else avail
        r, cons = b.work(nitems * 1504, avail, rd.ctypes.data + b.nread * 1504, out.ctypes.data + oo)
        if r > 0:
            oo += r
        if cons == 0:
            break
    return cd, rd, out[:oo].copy()


def rx_acquisition(x, tm, max_symbols=None, gi=G1_32):
    """ofdm_sym_acquisition one symbol per call with >= 2N+cp+16 samples visible
    (ofdm_sym_acquisition_impl.cc:488-568).  x: complex64 samples at the OFDM rate.
    Returns (symbols (nout, N) complex64, consumed samples, tags)."""
    N, P, K, cp = mode_dims(tm, gi)
    x = np.ascontiguousarray(x, np.complex64)
    pad = np.zeros(64, np.complex64)
    buf = np.concatenate([pad, x, pad])  # the reference reads in[-2] on initial acquisition
    b = RefBlock("ofdm_sym_acquisition", 1, N, K, cp, 30.0)
    out = np.zeros((len(x) // (N + cp) + 2, N), np.complex64)
    pos, nout = 0, 0
    need = 2 * N + cp + 16
    while pos + need <= len(x) and (max_symbols is None or nout < max_symbols):
        r, cons = b.work(1, len(x) - pos, buf.ctypes.data + (64 + pos) * 8, out.ctypes.data + nout * N * 8)
        if r > 0:
            nout += r
        pos += cons
        if cons == 0:
            break
    return out[:nout].copy(), pos, b.out_tags()
