"""ctypes declarations for include/dvbt_b200.h."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdvbt_b200.so")


class DvbtError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("libdvbt_b200: %s (code %d)" % (text, code))
        self.code = code


class Tag(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("key", C.c_int32), ("value", C.c_int64)]


TAG_SYNC_START, TAG_SUPERFRAME_START, TAG_SYMBOL_INDEX = 1, 2, 3
TAG_NAMES = {1: "sync_start", 2: "superframe_start", 3: "symbol_index"}
TAG_KEYS = {v: k for k, v in TAG_NAMES.items()}


class ViterbiParams(C.Structure):
    _fields_ = [("constellation", C.c_int), ("hierarchy", C.c_int), ("code_rate", C.c_int),
                ("bsize", C.c_int), ("S0", C.c_int), ("SK", C.c_int)]


class RsdecParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("p", "m", "gfpoly", "n", "k", "t", "s", "blocks")]


class DemapParams(C.Structure):
    _fields_ = [("nsize", C.c_int), ("constellation", C.c_int), ("hierarchy", C.c_int), ("transmission", C.c_int), ("gain", C.c_float)]


class DemodParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("itemsize", "ninput", "noutput", "constellation", "hierarchy", "code_rate_HP", "code_rate_LP",
                                       "guard_interval", "transmission_mode", "include_cell_id", "cell_id")]


class RxParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("constellation", "hierarchy", "code_rate", "guard_interval", "transmission_mode")]


class RxInfo(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("symbols_parsed", "first_symbol", "symbols_out", "viterbi_bytes", "viterbi_repaired",
                                             "rs_packets", "first_packet", "ts_bytes", "acq_symbols", "acq_cp_start", "acq_lost_at", "acq_run_symbols", "acq_single_symbols", "acq_sequential_symbols")] + \
               [(n, C.c_float) for n in ("ms_resample", "ms_acq_fft", "ms_demod", "ms_inner", "ms_viterbi", "ms_viterbi_acs", "ms_rs", "ms_descramble", "ms_fft", "ms_equalise")] + \
               [(n, C.c_longlong) for n in ("n_sync_start", "n_superframe_start", "n_viterbi_runs", "ts_total")]


class AcqParams(C.Structure):
    _fields_ = [("blocks", C.c_int), ("fft_length", C.c_int), ("occupied_tones", C.c_int), ("cp_length", C.c_int), ("snr", C.c_float)]


class ViterbiTuning(C.Structure):
    _fields_ = [("chunk_bytes", C.c_int), ("warmup_bytes", C.c_int), ("threads_per_block", C.c_int), ("ring_depth", C.c_int)]


_lib = None


def lib():
    """Loads libdvbt_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("gr_dvbt_b200/libdvbt_b200.so is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a).  There is no CPU fallback.")
        _lib = declare(C.CDLL(LIB_PATH))
    return _lib


def declare(L):
    """argument types of every entry point of include/dvbt_b200.h on a loaded library object"""
    if True:
        vp = C.c_void_p
        L.dvbt_b200_last_error.restype = C.c_char_p
        L.dvbt_b200_kernel_launches.restype = C.c_ulonglong
        L.dvbt_b200_set_device.argtypes = [C.c_int]
        L.dvbt_b200_viterbi_create.argtypes = [C.POINTER(ViterbiParams), C.POINTER(vp)]
        L.dvbt_b200_viterbi_destroy.argtypes = [vp]
        L.dvbt_b200_viterbi_set_tuning.argtypes = [vp, C.POINTER(ViterbiTuning)]
        L.dvbt_b200_viterbi_reset.argtypes = [vp]
        L.dvbt_b200_viterbi_forecast.argtypes = [vp, C.c_int]
        L.dvbt_b200_viterbi_output_multiple.argtypes = [vp]
        L.dvbt_b200_viterbi_ntraceback.argtypes = [vp]
        L.dvbt_b200_viterbi_work.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                             C.POINTER(Tag), C.c_size_t, C.POINTER(Tag), C.c_size_t, C.POINTER(C.c_size_t)]
        for name in ("dvbt_b200_viterbi_decode_host", "dvbt_b200_viterbi_decode_dev"):
            getattr(L, name).argtypes = [vp, vp, C.c_size_t, C.c_size_t, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dvbt_b200_viterbi_set_soft.argtypes = [vp, C.c_int]
        for name in ("dvbt_b200_viterbi_decode_soft_host", "dvbt_b200_viterbi_decode_soft_dev"):
            getattr(L, name).argtypes = [vp, vp, C.c_size_t, vp, C.POINTER(C.c_size_t)]
        L.dvbt_b200_viterbi_last_stats.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_float)]
        L.dvbt_b200_rsdec_create.argtypes = [C.POINTER(RsdecParams), C.POINTER(vp)]
        L.dvbt_b200_rsdec_destroy.argtypes = [vp]
        L.dvbt_b200_rsdec_set_compat.argtypes = [vp, C.c_int]
        L.dvbt_b200_rsdec_work.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.dvbt_b200_rsdec_decode_dev.argtypes = [vp, vp, C.c_size_t, vp, vp]
        L.dvbt_b200_demap_create.argtypes = [C.POINTER(DemapParams), C.POINTER(vp)]
        L.dvbt_b200_demap_destroy.argtypes = [vp]
        L.dvbt_b200_demap_points.argtypes = [vp, vp, C.c_int]
        L.dvbt_b200_demap_work.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.dvbt_b200_demap_run_dev.argtypes = [vp, vp, C.c_size_t, vp]
        L.dvbt_b200_demod_create.argtypes = [C.POINTER(DemodParams), C.POINTER(vp)]
        L.dvbt_b200_demod_destroy.argtypes = [vp]
        L.dvbt_b200_demod_work.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                           C.POINTER(Tag), C.c_size_t, C.POINTER(Tag), C.c_size_t, C.POINTER(C.c_size_t)]
        L.dvbt_b200_rx_create.argtypes = [C.POINTER(RxParams), C.POINTER(vp)]
        L.dvbt_b200_rx_destroy.argtypes = [vp]
        L.dvbt_b200_set_blocking_wait.argtypes = [C.c_int]
        L.dvbt_b200_rx_set_rs_compat.argtypes = [vp, C.c_int]
        L.dvbt_b200_rx_set_soft_decision.argtypes = [vp, C.c_int, C.c_float]
        L.dvbt_b200_rx_run_freq_host.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dvbt_b200_rx_run_freq_dev.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dvbt_b200_rx_last_info.argtypes = [vp, C.POINTER(RxInfo)]
        L.dvbt_b200_rx_read_stage.argtypes = [vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dvbt_b200_acq_create.argtypes = [C.POINTER(AcqParams), C.POINTER(vp)]
        L.dvbt_b200_acq_destroy.argtypes = [vp]
        L.dvbt_b200_acq_work.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t),
                                         C.POINTER(Tag), C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
        L.dvbt_b200_rx_run_baseband_host.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dvbt_b200_rx_run_baseband_dev.argtypes = [vp, vp, C.c_size_t, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dvbt_b200_rx_run_file_host.argtypes = [vp, vp, C.c_size_t, C.c_float, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dvbt_b200_rx_run_file_dev.argtypes = [vp, vp, C.c_size_t, C.c_float, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dvbt_b200_rx_stream_reset.argtypes = [vp]
        for name in ("dvbt_b200_rx_stream_push_host", "dvbt_b200_rx_stream_push_dev"):
            getattr(L, name).argtypes = [vp, C.c_int, vp, C.c_size_t, C.c_float, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dvbt_b200_tx_create.argtypes = [C.POINTER(RxParams), C.POINTER(vp)]
        L.dvbt_b200_tx_destroy.argtypes = [vp]
        for name in ("dvbt_b200_tx_run_host", "dvbt_b200_tx_run_dev"):
            getattr(L, name).argtypes = [vp, vp, C.c_size_t, C.c_int, C.c_float, vp, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.dvbt_b200_tx_read_stage.argtypes = [vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.dvbt_b200_resampler_taps.argtypes = [vp, C.c_int]
        L.dvbt_b200_resample_host.argtypes = [vp, C.c_size_t, C.c_float, vp, C.c_size_t, C.POINTER(C.c_size_t), C.c_int]
    return L


def check(rc):
    if rc != 0:
        raise DvbtError(rc, lib().dvbt_b200_last_error().decode(errors="replace"))
