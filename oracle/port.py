"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/libdvbt_oracle.so (oracle/port/*.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libdvbt_oracle.so")

RATE_KN = {0: (1, 2), 1: (2, 3), 2: (3, 4), 3: (5, 6), 4: (7, 8)}
NTRACEBACK = {0: 5, 1: 9, 2: 10, 3: 15, 4: 24}

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            build()
        L = C.CDLL(SO)
        L.dvbt_oracle_viterbi_create.restype = C.c_void_p
        L.dvbt_oracle_viterbi_create.argtypes = [C.c_int, C.c_int, C.c_int]
        L.dvbt_oracle_viterbi_destroy.argtypes = [C.c_void_p]
        L.dvbt_oracle_viterbi_reset.argtypes = [C.c_void_p]
        L.dvbt_oracle_viterbi_work.restype = C.c_long
        L.dvbt_oracle_viterbi_work.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        L.dvbt_oracle_viterbi_in_bytes_per_block.argtypes = [C.c_void_p]
        L.dvbt_oracle_viterbi_out_bytes_per_block.argtypes = [C.c_void_p]
        L.dvbt_oracle_viterbi_ntraceback.argtypes = [C.c_void_p]
        L.dvbt_oracle_viterbi_metrics.argtypes = [C.c_void_p, C.c_void_p]
        L.dvbt_oracle_viterbi_soft.restype = C.c_long
        L.dvbt_oracle_viterbi_soft.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p]
        L.dvbt_oracle_conv_encode.restype = C.c_long
        L.dvbt_oracle_conv_encode.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_void_p]
        L.dvbt_oracle_rs_decode.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_int, C.c_void_p]
        L.dvbt_oracle_rs_encode.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.dvbt_oracle_constellation.argtypes = [C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.dvbt_oracle_demap.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.dvbt_oracle_symbol_deinterleave.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p]
        L.dvbt_oracle_bit_deinterleave.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p]
        L.dvbt_oracle_conv_deinterleave.argtypes = [C.c_void_p, C.c_long, C.c_void_p]
        L.dvbt_oracle_descramble.restype = C.c_long
        L.dvbt_oracle_descramble.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
        L.dvbt_oracle_descramble_calls.restype = C.c_long
        L.dvbt_oracle_descramble_calls.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.dvbt_oracle_demod_create.restype = C.c_void_p
        L.dvbt_oracle_demod_create.argtypes = [C.c_int, C.c_int]
        L.dvbt_oracle_demod_destroy.argtypes = [C.c_void_p]
        L.dvbt_oracle_demod_run.restype = C.c_long
        L.dvbt_oracle_demod_run.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.dvbt_oracle_acq_create.restype = C.c_void_p
        L.dvbt_oracle_acq_create.argtypes = [C.c_int, C.c_int, C.c_float]
        L.dvbt_oracle_acq_destroy.argtypes = [C.c_void_p]
        L.dvbt_oracle_acq_run.restype = C.c_long
        L.dvbt_oracle_acq_run.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class Viterbi:
    """Restated viterbi_decoder block (one instance = one decoder; no global state)."""

    def __init__(self, m, rate, bsize=768):
        self.L = lib()
        self.h = self.L.dvbt_oracle_viterbi_create(m, rate, bsize)
        if not self.h:
            raise ValueError("bad viterbi parameters")
        self.in_per_block = self.L.dvbt_oracle_viterbi_in_bytes_per_block(self.h)
        self.out_per_block = self.L.dvbt_oracle_viterbi_out_bytes_per_block(self.h)
        self.ntb = self.L.dvbt_oracle_viterbi_ntraceback(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.dvbt_oracle_viterbi_destroy(self.h)
            self.h = None

    def reset(self):
        self.L.dvbt_oracle_viterbi_reset(self.h)

    def work(self, inp, nblocks=None):
        inp = np.ascontiguousarray(inp, np.uint8).reshape(-1)
        if nblocks is None:
            nblocks = len(inp) // self.in_per_block
        assert nblocks * self.in_per_block <= len(inp)
        out = np.zeros(nblocks * self.out_per_block, np.uint8)
        n = self.L.dvbt_oracle_viterbi_work(self.h, inp.ctypes.data, nblocks, out.ctypes.data)
        return out[: max(n, 0)]

    def metrics(self):
        m = np.zeros(64, np.uint8)
        self.L.dvbt_oracle_viterbi_metrics(self.h, m.ctypes.data)
        return m


def viterbi_soft(values, rate):
    """Soft-decision generalisation of the restated decoder (not in the reference; viterbi_port.c): one stream from a
    reset, one int8 per transmitted code bit (> 0 = "1", clamped to +-6).  +-1 values decode like the hard decoder."""
    v = np.ascontiguousarray(values, np.int8).reshape(-1)
    k, n = RATE_KN[rate]
    assert (len(v) * k) % (8 * n) == 0
    out = np.zeros(len(v) * k // (8 * n), np.uint8)
    r = lib().dvbt_oracle_viterbi_soft(v.ctypes.data, len(v), rate, out.ctypes.data)
    assert r >= 0
    return out[:r]


def conv_encode(data, m, rate):
    """K=7 encode + puncture + pack m bits per byte (the Viterbi block's input format)."""
    data = np.ascontiguousarray(data, np.uint8).reshape(-1)
    k, n = RATE_KN[rate]
    nbits = len(data) * 8
    assert nbits % k == 0 and (nbits * n // k) % m == 0
    out = np.zeros(nbits * n // k // m, np.uint8)
    r = lib().dvbt_oracle_conv_encode(data.ctypes.data, len(data), m, rate, out.ctypes.data)
    assert r == len(out), (r, len(out))
    return out


def flip_bits(coded, m, ber, seed):
    """Hard-decision channel: flip each of the m used bits of every byte with probability ber."""
    rng = np.random.default_rng(seed)
    coded = coded.copy()
    for j in range(m):
        coded ^= (rng.random(len(coded)) < ber).astype(np.uint8) << j
    return coded


def rs_encode(packets188):
    """(npk,188) -> (npk,204) systematic RS(204,188) code words (generator roots a^0..a^15)."""
    d = np.ascontiguousarray(packets188, np.uint8).reshape(-1, 188)
    out = np.zeros((d.shape[0], 204), np.uint8)
    lib().dvbt_oracle_rs_encode(d.ctypes.data, d.shape[0], out.ctypes.data)
    return out


def rs_decode(packets204, as_built=False):
    """(npk,204) -> ((npk,188), status[npk]) following reed_solomon_dec_impl.cc / reed_solomon.cc."""
    d = np.ascontiguousarray(packets204, np.uint8).reshape(-1, 204)
    out = np.zeros((d.shape[0], 188), np.uint8)
    st = np.zeros(d.shape[0], np.int32)
    lib().dvbt_oracle_rs_decode(d.ctypes.data, d.shape[0], out.ctypes.data, int(as_built), st.ctypes.data)
    return out, st


def constellation_points(constellation, alpha=1, gain=1.0):
    pts = np.zeros(128, np.float32)
    n = lib().dvbt_oracle_constellation(constellation, alpha, gain, pts.ctypes.data)
    return pts[: 2 * n].view(np.complex64).copy()


def demap(cells, constellation, alpha=1, gain=1.0):
    """complex64 cells -> hard-decision bytes (dvbt_demap_impl.cc:167-203)."""
    c = np.ascontiguousarray(cells, np.complex64).reshape(-1)
    out = np.zeros(len(c), np.uint8)
    lib().dvbt_oracle_demap(c.ctypes.data, len(c), constellation, alpha, gain, out.ctypes.data)
    return out


def symbol_deinterleave(cells, tm, symbol_index):
    P = 1512 if tm == 0 else 6048
    c = np.ascontiguousarray(cells, np.uint8).reshape(-1, P)
    si = np.ascontiguousarray(symbol_index, np.int32)
    out = np.zeros_like(c)
    lib().dvbt_oracle_symbol_deinterleave(c.ctypes.data, c.shape[0], tm, si.ctypes.data, out.ctypes.data)
    return out


def bit_deinterleave(cells, m):
    c = np.ascontiguousarray(cells, np.uint8).reshape(-1)
    out = np.zeros_like(c)
    lib().dvbt_oracle_bit_deinterleave(c.ctypes.data, len(c), m, out.ctypes.data)
    return out


def conv_deinterleave(stream):
    s = np.ascontiguousarray(stream, np.uint8).reshape(-1)
    s = s[: len(s) // 12 * 12]
    out = np.zeros_like(s)
    lib().dvbt_oracle_conv_deinterleave(s.ctypes.data, len(s), out.ctypes.data)
    return out


def descramble(packets188):
    p = np.ascontiguousarray(packets188, np.uint8).reshape(-1, 188)
    out = np.zeros(p.size, np.uint8)
    first = C.c_long(-1)
    n = lib().dvbt_oracle_descramble(p.ctypes.data, p.shape[0], out.ctypes.data, C.byref(first))
    return out[:n], int(first.value)


def descramble_calls(packets188, pk=0, flush=False):
    """energy_descramble at the scheduler's smallest call size over the pending packets (energy_descramble_impl.cc:108-174):
    returns (ts bytes, items consumed, d_index in packets after the calls, index of the first packet output or -1)"""
    p = np.ascontiguousarray(packets188, np.uint8).reshape(-1, 188)
    out = np.zeros(p.size + 16, np.uint8)
    first, used, pkc = C.c_long(-1), C.c_long(0), C.c_int(int(pk))
    n = lib().dvbt_oracle_descramble_calls(p.ctypes.data, p.shape[0], int(flush), C.byref(pkc), out.ctypes.data, C.byref(used), C.byref(first))
    return out[:n], int(used.value), int(pkc.value), int(first.value)


def demod(X, constellation, tm, sync_start_at0=True):
    """post-FFT symbols (nsym, N) -> (cells (nout, P) complex64, symbol_index tags, superframe tag index)"""
    N, P = (2048, 1512) if tm == 0 else (8192, 6048)
    X = np.ascontiguousarray(X, np.complex64).reshape(-1, N)
    pad = np.zeros((X.shape[0] + 1) * N + 64, np.complex64)   # the reference reads up to 8 bins around an item
    pad[32: 32 + X.size] = X.reshape(-1)
    out = np.zeros((X.shape[0], P), np.complex64)
    si = np.zeros(X.shape[0], np.int32)
    tag = C.c_long(-1)
    h = lib().dvbt_oracle_demod_create(constellation, tm)
    n = lib().dvbt_oracle_demod_run(h, pad.ctypes.data + 32 * 8, X.shape[0], int(sync_start_at0), out.ctypes.data, si.ctypes.data, C.byref(tag))
    lib().dvbt_oracle_demod_destroy(h)
    return out[:n].copy(), si[:n].copy(), int(tag.value)


def acquisition(x, N, cp, snr_db=30.0):
    """samples -> (symbols (nout, N) complex64, consumed samples, sync_start on first item)"""
    x = np.ascontiguousarray(x, np.complex64).reshape(-1)
    cap = len(x) // (N + cp) + 2
    out = np.zeros((cap, N), np.complex64)
    cons, tag = C.c_long(0), C.c_int(0)
    h = lib().dvbt_oracle_acq_create(N, cp, snr_db)
    n = lib().dvbt_oracle_acq_run(h, x.ctypes.data, len(x), out.ctypes.data, cap, C.byref(cons), C.byref(tag))
    lib().dvbt_oracle_acq_destroy(h)
    return out[:n].copy(), int(cons.value), bool(tag.value)
