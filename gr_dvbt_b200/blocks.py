"""Host-side mirror of the reference block constructors on top of the C ABI.

Names, argument order and meaning follow the reference's make() functions
(include/dvbt/<block>.h) so parity tests read like tests of the reference blocks.
Buffers are numpy arrays (host) or raw device pointers (ints) for the *_dev calls.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import check, lib

__all__ = ["viterbi_decoder", "reed_solomon_dec", "dvbt_demap", "demod_reference_signals", "rx_chain", "tx_chain", "ofdm_sym_acquisition", "QPSK", "QAM16", "QAM64", "NH", "C1_2", "C2_3", "C3_4", "C5_6", "C7_8", "T2k", "T8k",
           "G1_32", "G1_16", "G1_8", "G1_4"]

QPSK, QAM16, QAM64 = 0, 1, 2
NH = 0
C1_2, C2_3, C3_4, C5_6, C7_8 = 0, 1, 2, 3, 4
T2k, T8k = 0, 1
G1_32, G1_16, G1_8, G1_4 = 0, 1, 2, 3


def _addr(x):
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    return int(x)


class viterbi_decoder:
    """dvbt.viterbi_decoder(constellation, hierarchy, coderate, bsize, S0, SK)
    (include/dvbt/viterbi_decoder.h:51-52)."""

    def __init__(self, constellation, hierarchy, coderate, bsize=768, S0=0, SK=-1):
        self._h = C.c_void_p()
        p = capi.ViterbiParams(constellation, hierarchy, coderate, bsize, S0, SK)
        check(lib().dvbt_b200_viterbi_create(C.byref(p), C.byref(self._h)))
        self.output_multiple = lib().dvbt_b200_viterbi_output_multiple(self._h)
        self.ntraceback = lib().dvbt_b200_viterbi_ntraceback(self._h)
        self.m = 2 * (constellation + 1)
        self.k = [1, 2, 3, 5, 7][coderate]
        self.n = self.k + 1

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().dvbt_b200_viterbi_destroy(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    __del__ = close

    def set_tuning(self, chunk_bytes=0, warmup_bytes=0, threads_per_block=0, ring_depth=0):
        t = capi.ViterbiTuning(chunk_bytes, warmup_bytes, threads_per_block, ring_depth)
        check(lib().dvbt_b200_viterbi_set_tuning(self._h, C.byref(t)))

    def reset(self):
        check(lib().dvbt_b200_viterbi_reset(self._h))

    def forecast(self, noutput_items):
        return lib().dvbt_b200_viterbi_forecast(self._h, noutput_items)

    def general_work(self, noutput_items, inp, tags=()):
        """One scheduler call on host arrays.  tags: iterable of (offset, key_name, value) relative
        to inp[0].  Returns (out array of the produced items, consumed, out_tags)."""
        inp = np.ascontiguousarray(inp, np.uint8)
        out = np.zeros(max(noutput_items, 1), np.uint8)
        tin = (capi.Tag * max(1, len(tags)))()
        for i, (off, key, val) in enumerate(tags):
            tin[i] = capi.Tag(off, capi.TAG_KEYS[key], val)
        tout = (capi.Tag * 4)()
        ntout = C.c_size_t(0)
        cons, prod = C.c_size_t(0), C.c_size_t(0)
        check(lib().dvbt_b200_viterbi_work(self._h, inp.ctypes.data, inp.size, out.ctypes.data, noutput_items,
                                           C.byref(cons), C.byref(prod), tin, len(tags), tout, 4, C.byref(ntout)))
        otags = [(int(tout[i].offset), capi.TAG_NAMES[tout[i].key], int(tout[i].value)) for i in range(ntout.value)]
        return out[: prod.value].copy(), int(cons.value), otags

    def set_soft(self, on=True):
        """soft-decision mode (beyond the reference, include/dvbt_b200.h): decode_soft() instead of decode() / general_work()"""
        check(lib().dvbt_b200_viterbi_set_soft(self._h, 1 if on else 0))

    def decode_soft(self, values):
        """one stream from a reset; values: one int8 per transmitted code bit, > 0 = "1", clamped to +-6"""
        v = np.ascontiguousarray(values, np.int8).reshape(-1)
        nbt = len(v) * self.k // (8 * self.n)
        out = np.zeros(max(nbt - self.ntraceback, 0), np.uint8)
        n_out = C.c_size_t(0)
        check(lib().dvbt_b200_viterbi_decode_soft_host(self._h, v.ctypes.data, len(v), out.ctypes.data, C.byref(n_out)))
        assert n_out.value == len(out)
        return out

    def decode(self, inp, nstreams=1):
        """Batch decode of nstreams equal-length streams (rows of inp), each from a reset (host arrays)."""
        inp = np.ascontiguousarray(inp, np.uint8).reshape(nstreams, -1)
        n_in = inp.shape[1]
        nbt = n_in * self.m * self.k // (8 * self.n)
        out = np.zeros((nstreams, max(nbt - self.ntraceback, 0)), np.uint8)
        n_out = C.c_size_t(0)
        check(lib().dvbt_b200_viterbi_decode_host(self._h, inp.ctypes.data, n_in, n_in, nstreams, out.ctypes.data,
                                                  out.shape[1], C.byref(n_out)))
        assert n_out.value == out.shape[1]
        return out

    def decode_dev(self, d_in, in_stride, n_in, nstreams, d_out, out_stride):
        n_out = C.c_size_t(0)
        check(lib().dvbt_b200_viterbi_decode_dev(self._h, _addr(d_in), in_stride, n_in, nstreams, _addr(d_out), out_stride,
                                                 C.byref(n_out)))
        return int(n_out.value)

    def decode_soft_dev(self, d_in, n_in, d_out):
        """device buffers: n_in int8 soft values (one per transmitted code bit) -> decoded bytes; returns their count"""
        n_out = C.c_size_t(0)
        check(lib().dvbt_b200_viterbi_decode_soft_dev(self._h, _addr(d_in), n_in, _addr(d_out), C.byref(n_out)))
        return int(n_out.value)

    def last_stats(self):
        a, b, ms = C.c_longlong(0), C.c_longlong(0), C.c_float(0)
        check(lib().dvbt_b200_viterbi_last_stats(self._h, C.byref(a), C.byref(b), C.byref(ms)))
        return dict(chunks=int(a.value), repaired=int(b.value), acs_kernel_ms=float(ms.value))


class _Handle:
    _destroy = None

    def close(self):
        if getattr(self, "_h", None):
            try:
                getattr(lib(), self._destroy)(self._h)
            except Exception:  # interpreter shutdown
                pass
            self._h = None

    __del__ = close


class reed_solomon_dec(_Handle):
    """dvbt.reed_solomon_dec(p, m, gfpoly, n, k, t, s, blocks) (include/dvbt/reed_solomon_dec.h:49)."""
    _destroy = "dvbt_b200_rsdec_destroy"

    def __init__(self, p=2, m=8, gfpoly=0x11D, n=255, k=239, t=8, s=51, blocks=8):
        self._h = C.c_void_p()
        par = capi.RsdecParams(p, m, gfpoly, n, k, t, s, blocks)
        check(lib().dvbt_b200_rsdec_create(C.byref(par), C.byref(self._h)))
        self.blocks = blocks

    def set_compat(self, as_built):
        check(lib().dvbt_b200_rsdec_set_compat(self._h, int(as_built)))

    def general_work(self, noutput_items, inp):
        inp = np.ascontiguousarray(inp, np.uint8)
        out = np.zeros(noutput_items * self.blocks * 188, np.uint8)
        cons, prod = C.c_size_t(0), C.c_size_t(0)
        check(lib().dvbt_b200_rsdec_work(self._h, inp.ctypes.data, inp.size // (self.blocks * 204), out.ctypes.data, noutput_items,
                                         C.byref(cons), C.byref(prod)))
        return out[: prod.value * self.blocks * 188], int(cons.value)

    def decode_dev(self, d_in, npackets, d_out, d_status=None):
        check(lib().dvbt_b200_rsdec_decode_dev(self._h, _addr(d_in), npackets, _addr(d_out), _addr(d_status) if d_status is not None else None))


class dvbt_demap(_Handle):
    """dvbt.dvbt_demap(nsize, constellation, hierarchy, transmission, gain) (include/dvbt/dvbt_demap.h:50)."""
    _destroy = "dvbt_b200_demap_destroy"

    def __init__(self, nsize, constellation, hierarchy, transmission, gain=1.0):
        self._h = C.c_void_p()
        par = capi.DemapParams(nsize, constellation, hierarchy, transmission, gain)
        check(lib().dvbt_b200_demap_create(C.byref(par), C.byref(self._h)))
        self.nsize = nsize

    def points(self):
        buf = np.zeros(128, np.float32)
        n = lib().dvbt_b200_demap_points(self._h, buf.ctypes.data, 64)
        if n < 0:
            check(n)
        return buf[: 2 * n].view(np.complex64).copy()

    def general_work(self, noutput_items, inp):
        inp = np.ascontiguousarray(inp, np.complex64)
        out = np.zeros(noutput_items * self.nsize, np.uint8)
        cons, prod = C.c_size_t(0), C.c_size_t(0)
        check(lib().dvbt_b200_demap_work(self._h, inp.ctypes.data, inp.size // self.nsize, out.ctypes.data, noutput_items,
                                         C.byref(cons), C.byref(prod)))
        return out[: prod.value * self.nsize], int(cons.value)

    def run_dev(self, d_in, ncells, d_out):
        check(lib().dvbt_b200_demap_run_dev(self._h, _addr(d_in), ncells, _addr(d_out)))


class demod_reference_signals(_Handle):
    """dvbt.demod_reference_signals(itemsize, ninput, noutput, constellation, hierarchy, code_rate_HP, code_rate_LP,
    guard_interval, transmission_mode, include_cell_id, cell_id) (include/dvbt/demod_reference_signals.h:50-54)."""
    _destroy = "dvbt_b200_demod_destroy"

    def __init__(self, itemsize, ninput, noutput, constellation, hierarchy, code_rate_HP, code_rate_LP, guard_interval,
                 transmission_mode, include_cell_id=0, cell_id=0):
        self._h = C.c_void_p()
        par = capi.DemodParams(itemsize, ninput, noutput, constellation, hierarchy, code_rate_HP, code_rate_LP, guard_interval,
                               transmission_mode, include_cell_id, cell_id)
        check(lib().dvbt_b200_demod_create(C.byref(par), C.byref(self._h)))
        self.N, self.P = ninput, noutput

    def general_work(self, inp, out_capacity=None, tags=()):
        """inp: (nsym, N) complex64.  Returns (cells (nout, P), consumed, out_tags)."""
        inp = np.ascontiguousarray(inp, np.complex64).reshape(-1, self.N)
        nsym = inp.shape[0]
        cap = out_capacity if out_capacity is not None else max(nsym - 1, 1)
        out = np.zeros((cap, self.P), np.complex64)
        tin = (capi.Tag * max(1, len(tags)))()
        for i, (off, key, val) in enumerate(tags):
            tin[i] = capi.Tag(off, capi.TAG_KEYS[key], val)
        tout = (capi.Tag * (cap + 4))()
        ntout = C.c_size_t(0)
        cons, prod = C.c_size_t(0), C.c_size_t(0)
        check(lib().dvbt_b200_demod_work(self._h, inp.ctypes.data, nsym, out.ctypes.data, cap, C.byref(cons), C.byref(prod),
                                         tin, len(tags), tout, cap + 4, C.byref(ntout)))
        otags = [(int(tout[i].offset), capi.TAG_NAMES[tout[i].key], int(tout[i].value)) for i in range(ntout.value)]
        return out[: prod.value].copy(), int(cons.value), otags


class rx_chain(_Handle):
    """Fused device-resident receive chain (include/dvbt_b200.h: dvbt_b200_rx_*): what apps/dvbt_rx_demo*.grc
    does from the FFT output to the TS file."""
    _destroy = "dvbt_b200_rx_destroy"
    STAGES = dict(cells=(0, np.complex64), demap=(1, np.uint8), bitdeint=(2, np.uint8), viterbi=(3, np.uint8), rs=(4, np.uint8),
                  rs_status=(5, np.int32), symbol_index=(6, np.int32),
                  soft_cells=(7, np.uint32), soft_values=(8, np.int8))

    def __init__(self, constellation, hierarchy, code_rate, guard_interval, transmission_mode):
        self._h = C.c_void_p()
        par = capi.RxParams(constellation, hierarchy, code_rate, guard_interval, transmission_mode)
        check(lib().dvbt_b200_rx_create(C.byref(par), C.byref(self._h)))
        self.N = 2048 if transmission_mode == T2k else 8192
        self.P = 1512 if transmission_mode == T2k else 6048

    def set_rs_compat(self, as_built):
        check(lib().dvbt_b200_rx_set_rs_compat(self._h, int(as_built)))

    def set_soft_decision(self, on=True, scale=0.0):
        """soft-decision mode of the chain (beyond the reference; include/dvbt_b200.h: dvbt_b200_rx_set_soft_decision)"""
        check(lib().dvbt_b200_rx_set_soft_decision(self._h, 1 if on else 0, float(scale)))

    def run_freq(self, X):
        """X: (nsym, N) complex64 host array -> TS bytes (numpy)."""
        X = np.ascontiguousarray(X, np.complex64).reshape(-1, self.N)
        cap = X.shape[0] * self.P + 4096
        ts = np.zeros(cap, np.uint8)
        n = C.c_size_t(0)
        check(lib().dvbt_b200_rx_run_freq_host(self._h, X.ctypes.data, X.shape[0], ts.ctypes.data, cap, C.byref(n)))
        return ts[: n.value].copy()

    def run_baseband(self, samples):
        """samples: complex64 host array at the OFDM sample rate -> TS bytes."""
        x = np.ascontiguousarray(samples, np.complex64).reshape(-1)
        cap = len(x) + 4096
        ts = np.zeros(cap, np.uint8)
        n = C.c_size_t(0)
        check(lib().dvbt_b200_rx_run_baseband_host(self._h, x.ctypes.data, len(x), ts.ctypes.data, cap, C.byref(n)))
        return ts[: n.value].copy()

    def run_file(self, samples, gain):
        """samples: complex64 host array at 10 Msps (the capture file of the flowgraphs) -> TS bytes."""
        x = np.ascontiguousarray(samples, np.complex64).reshape(-1)
        cap = len(x) + 4096
        ts = np.zeros(cap, np.uint8)
        n = C.c_size_t(0)
        check(lib().dvbt_b200_rx_run_file_host(self._h, x.ctypes.data, len(x), gain, ts.ctypes.data, cap, C.byref(n)))
        return ts[: n.value].copy()

    LEVELS = dict(file=0, baseband=1, freq=2)

    def stream_reset(self):
        check(lib().dvbt_b200_rx_stream_reset(self._h))

    def stream_push(self, level, data, gain=1.0, end=False):
        """one piece of a stream (include/dvbt_b200.h: dvbt_b200_rx_stream_push_host).  level: 'file' (10 Msps capture
        samples), 'baseband' (OFDM-rate samples) or 'freq' ((nsym, N) post-FFT symbols).  Returns the TS bytes that
        became available."""
        x = np.ascontiguousarray(data, np.complex64).reshape(-1)
        lv = self.LEVELS[level]
        count = len(x) // self.N if lv == 2 else len(x)
        cap = len(x) + 2 * 16 * 204 * 64 + 4096
        ts = np.zeros(cap, np.uint8)
        n = C.c_size_t(0)
        check(lib().dvbt_b200_rx_stream_push_host(self._h, lv, x.ctypes.data, count, gain, int(end), ts.ctypes.data, cap, C.byref(n)))
        return ts[: n.value].copy()

    def run_file_dev(self, d_x, nsamples, gain, d_ts, ts_capacity):
        n = C.c_size_t(0)
        check(lib().dvbt_b200_rx_run_file_dev(self._h, _addr(d_x), nsamples, gain, _addr(d_ts), ts_capacity, C.byref(n)))
        return int(n.value)

    def run_baseband_dev(self, d_x, nsamples, d_ts, ts_capacity):
        n = C.c_size_t(0)
        check(lib().dvbt_b200_rx_run_baseband_dev(self._h, _addr(d_x), nsamples, _addr(d_ts), ts_capacity, C.byref(n)))
        return int(n.value)

    def run_freq_dev(self, d_X, nsym, d_ts, ts_capacity):
        n = C.c_size_t(0)
        check(lib().dvbt_b200_rx_run_freq_dev(self._h, _addr(d_X), nsym, _addr(d_ts), ts_capacity, C.byref(n)))
        return int(n.value)

    def info(self):
        i = capi.RxInfo()
        check(lib().dvbt_b200_rx_last_info(self._h, C.byref(i)))
        return {n: getattr(i, n) for n, _ in capi.RxInfo._fields_}

    def stage(self, name):
        sid, dt = self.STAGES[name]
        inf = self.info()
        cap = max(inf["symbols_parsed"], 1) * self.P * 8 + 4096
        buf = np.zeros(cap, np.uint8)
        n = C.c_size_t(0)
        check(lib().dvbt_b200_rx_read_stage(self._h, sid, buf.ctypes.data, cap, C.byref(n)))
        return buf[: n.value].view(dt).copy()


class tx_chain(_Handle):
    """The transmit flowgraph on the device (include/dvbt_b200.h: dvbt_b200_tx_*): TS packets -> frequency-domain symbols /
    baseband / 10 Msps capture.  A synthetic-input generator for the receive path."""
    _destroy = "dvbt_b200_tx_destroy"
    LEVELS = dict(file=0, baseband=1, freq=2)
    STAGES = dict(energy=0, rs=1, outer=2, inner_coder=3, bit_interleaver=4, symbol_interleaver=5)

    def __init__(self, constellation, hierarchy, code_rate, guard_interval, transmission_mode):
        self._h = C.c_void_p()
        par = capi.RxParams(constellation, hierarchy, code_rate, guard_interval, transmission_mode)
        check(lib().dvbt_b200_tx_create(C.byref(par), C.byref(self._h)))
        self.N = 2048 if transmission_mode == T2k else 8192
        self.P = 1512 if transmission_mode == T2k else 6048
        self.cp = self.N // (32 >> guard_interval)

    def run(self, ts, level="freq", gain=1.0):
        """ts: TS bytes (188-byte packets).  Returns complex64 (nsym, N) / samples and the symbol count."""
        ts = np.ascontiguousarray(ts, np.uint8).reshape(-1)
        npk = len(ts) // 188
        cap = (npk * 204 // 300 + 8) * (self.N + self.cp) * 35 // 32 + 4096   # >= any level's size (a symbol carries >= 378 bytes)
        out = np.zeros(cap, np.complex64)
        n, nsym = C.c_size_t(0), C.c_size_t(0)
        check(lib().dvbt_b200_tx_run_host(self._h, ts.ctypes.data, npk, self.LEVELS[level], gain, out.ctypes.data, cap, C.byref(n), C.byref(nsym)))
        res = out[: n.value].copy()
        return (res.reshape(-1, self.N) if level == "freq" else res), int(nsym.value)

    def run_dev(self, d_ts, npackets, level, gain, d_out, capacity):
        n, nsym = C.c_size_t(0), C.c_size_t(0)
        check(lib().dvbt_b200_tx_run_dev(self._h, _addr(d_ts), npackets, self.LEVELS[level], gain, _addr(d_out), capacity, C.byref(n), C.byref(nsym)))
        return int(n.value), int(nsym.value)

    def stage(self, name, capacity):
        buf = np.zeros(capacity, np.uint8)
        n = C.c_size_t(0)
        check(lib().dvbt_b200_tx_read_stage(self._h, self.STAGES[name], buf.ctypes.data, capacity, C.byref(n)))
        return buf[: n.value].copy()


class ofdm_sym_acquisition(_Handle):
    """dvbt.ofdm_sym_acquisition(blocks, fft_length, occupied_tones, cp_length, snr) (include/dvbt/ofdm_sym_acquisition.h:49)."""
    _destroy = "dvbt_b200_acq_destroy"

    def __init__(self, blocks, fft_length, occupied_tones, cp_length, snr):
        self._h = C.c_void_p()
        par = capi.AcqParams(blocks, fft_length, occupied_tones, cp_length, snr)
        check(lib().dvbt_b200_acq_create(C.byref(par), C.byref(self._h)))
        self.N, self.cp = fft_length, cp_length

    def general_work(self, samples, out_capacity=None, apply_fft=False):
        """samples: complex64 host array.  Returns (symbols (nout, N), consumed, out_tags)."""
        x = np.ascontiguousarray(samples, np.complex64).reshape(-1)
        cap = out_capacity if out_capacity is not None else len(x) // (self.N + self.cp) + 1
        out = np.zeros((cap, self.N), np.complex64)
        tout = (capi.Tag * 4)()
        ntout = C.c_size_t(0)
        cons, prod = C.c_size_t(0), C.c_size_t(0)
        check(lib().dvbt_b200_acq_work(self._h, x.ctypes.data, len(x), out.ctypes.data, cap, C.byref(cons), C.byref(prod),
                                       tout, 4, C.byref(ntout), int(apply_fft)))
        otags = [(int(tout[i].offset), capi.TAG_NAMES[tout[i].key], int(tout[i].value)) for i in range(ntout.value)]
        return out[: prod.value].copy(), int(cons.value), otags
