"""Every RX flowgraph the reference ships (apps/dvbt_rx_demo*.grc = BASELINE.json configs[0..3] + the 8k/QPSK one),
run the way the flowgraph runs it: a 10 Msps complex64 capture (what file_source reads from testBB.bin) ->
rational_resampler 64/70 -> multiply_const -> ofdm_sym_acquisition -> FFT -> ... -> energy_descramble, through
dvbt_b200_rx_run_file_host.  Pass = the transport stream is the transmitted one from the packet SURVEY A.6 derives for
the mode (504 / 1328 / 3976 / 2016 / 1768), and the reference chain (oracle/_ref blocks; scipy/numpy for the stock
GNU Radio resampler and FFT, which gr-dvbt does not contain) returns a prefix of the same bytes.

configs[0] transmits the head of the reference's own apps/test.ts (tests/golden/apps_test_ts_head.npz); the 8k modes
start later than that head is long, so they transmit a seeded random TS."""
import os

import numpy as np
import pytest

from oracle import refchain as R

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
FX = np.load(os.path.join(os.path.dirname(__file__), "golden", "apps_test_ts_head.npz"))

GAIN_2K, GAIN_8K = 0.0022097087, 0.00055242272   # multiply_const of apps/dvbt_rx_demo.grc / dvbt_rx_demo_8k*.grc

# (flowgraph, constellation, code rate, mode, multiply_const, OFDM symbols transmitted, first TS packet, source)
CASES = [
    ("dvbt_rx_demo.grc", R.QAM16, R.C1_2, R.T2k, GAIN_2K, 420, 504, "apps/test.ts"),
    ("dvbt_rx_demo_2k_QAM64_rate78.grc", R.QAM64, R.C7_8, R.T2k, GAIN_2K, 330, 1328, "random"),
    ("dvbt_rx_demo_8k_QAM64_rate78.grc", R.QAM64, R.C7_8, R.T8k, GAIN_8K, 236, 3976, "random"),
    ("dvbt_rx_demo_8k.grc", R.QAM16, R.C1_2, R.T8k, GAIN_8K, 300, 2016, "random"),
    ("dvbt_rx_demo_8k_QPSK_rate78.grc", R.QPSK, R.C7_8, R.T8k, GAIN_8K, 300, 1768, "random"),
]


def transmit(con, cr, tm, nsym, source, seed=21):
    """TS -> reference TX blocks -> IFFT/CP -> 35/32 resampler: the 10 Msps capture and the TS it carries"""
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate, to_capture_rate
    if source == "apps/test.ts":
        ts = FX["ts_head"]
        ed, rs, ci = R.tx_outer(ts)
        tx = R.tx_inner(ci, con, cr, tm, nsym=nsym)
    else:
        tx = tx_frequency_domain(con, cr, tm, nsym, seed)
        ts = tx["ts"]
    x = ofdm_modulate(tx["X"], tm, gain=1.0, offset=500, seed=4)   # the TX flowgraph's multiply_const is folded into the RX gain
    return to_capture_rate(x), ts


def reference_capture_rx(cap, gain, con, cr, tm):
    """the reference flowgraph on the CPU: stock blocks restated with scipy/numpy, gr-dvbt blocks from oracle/_ref"""
    from scipy.signal import resample_poly
    from test_rx_chain_gpu import reference_rx
    x = (resample_poly(cap.astype(np.complex128), 32, 35, window=("kaiser", 7.0)) * gain).astype(np.complex64)
    sym, cons, tags = R.rx_acquisition(x, tm)
    Xf = np.fft.fftshift(np.fft.fft(sym.astype(np.complex128), axis=1), axes=1).astype(np.complex64)
    return reference_rx(Xf, con, cr, tm), sym.shape[0]


@needs_ref
@pytest.mark.parametrize("grc,con,cr,tm,gain,nsym,first_packet,source", CASES, ids=[c[0] for c in CASES])
def test_flowgraph_from_capture_to_ts(grc, con, cr, tm, gain, nsym, first_packet, source):
    import gr_dvbt_b200 as g
    cap, src = transmit(con, cr, tm, nsym, source)
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    ts = rx.run_file(cap, gain)
    info = rx.info()
    # a peak lost right after the initial search restarts acquisition half a symbol later, in the reference as well
    # (ofdm_sym_acquisition_impl.cc:545-558); it must not happen once the receiver is in superframe sync
    assert info["acq_lost_at"] == -1 or info["acq_lost_at"] < info["first_symbol"], info
    assert info["acq_symbols"] >= nsym - 3
    assert info["first_packet"] >= 0 and len(ts) >= 2 * 1504
    # the transmitted stream, from the packet the mode's superframe alignment gives
    assert np.array_equal(ts, src[first_packet * 188: first_packet * 188 + len(ts)])
    # and the reference flowgraph delivers a (scheduler-dependent) prefix of the same bytes
    ref, nref = reference_capture_rx(cap, gain, con, cr, tm)
    assert abs(info["acq_symbols"] - nref) <= 1
    assert len(ref["ts"]) >= 1504 and np.array_equal(ts[: len(ref["ts"])], ref["ts"])
    assert info["symbols_out"] in (ref["Y"].shape[0], ref["Y"].shape[0] + 1)
