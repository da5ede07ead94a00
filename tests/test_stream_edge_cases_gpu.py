"""Edge cases of the stream interface of the fused chain (include/dvbt_b200.h: dvbt_b200_rx_stream_push_*): empty and
one-sample pieces, an empty stream, input shorter than a symbol, silence, a capacity that is too small, a level change
in mid-stream, and that an error or a mode switch leaves the handle ready for a new stream.  The reference blocks see
the same situations as scheduler calls with few or no items (forecast() not satisfied: general_work is simply not
called, e.g. lib/ofdm_sym_acquisition_impl.cc:468-486, lib/demod_reference_signals_impl.cc:84-94); the chain must
neither invent output nor lose state."""
import ctypes as C

import numpy as np
import pytest

from oracle import refchain as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capture():
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    tx = tx_frequency_domain(con, cr, tm, 420, 17)
    x = ofdm_modulate(tx["X"], tm, offset=333, cfo_bins=0.0, seed=2)
    return con, cr, tm, tx, x


def test_empty_and_one_sample_pieces_change_nothing(capture):
    import gr_dvbt_b200 as g
    con, cr, tm, tx, x = capture
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    one = rx.run_baseband(x)
    assert len(one) >= 1504 * 4
    rx.stream_reset()
    got = [rx.stream_push("baseband", x[:0])]                       # nothing at all, first call of the stream
    for i in range(40):                                             # forty one-sample pieces, an empty one in between
        got.append(rx.stream_push("baseband", x[i:i + 1]))
        if i % 7 == 0:
            got.append(rx.stream_push("baseband", x[:0]))
    cuts = [40, 2111, 2112, 2113, 50000, 50001, 400000, len(x)]
    for a, b in zip(cuts[:-1], cuts[1:]):
        got.append(rx.stream_push("baseband", x[a:b]))
        got.append(rx.stream_push("baseband", x[:0]))
    got.append(rx.stream_push("baseband", x[:0], end=True))          # the end of the stream arrives with no data
    ts = np.concatenate(got)
    assert np.array_equal(ts[: len(one)], one) and len(ts) >= len(one)
    assert all(len(p) % 1504 == 0 for p in got)


def test_empty_stream_and_input_shorter_than_a_symbol(capture):
    import gr_dvbt_b200 as g
    con, cr, tm, tx, x = capture
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    rx.stream_reset()
    assert len(rx.stream_push("baseband", x[:0], end=True)) == 0     # a stream that ends before it starts
    assert len(rx.run_baseband(x[:1000])) == 0                       # less than one symbol: acquisition never runs
    assert len(rx.run_baseband(x[:3 * 2112])) == 0                   # a few symbols: no superframe start yet
    assert len(rx.run_file(x[:500], 1.0)) == 0
    X = tx["X"]
    assert len(rx.run_freq(X[:1])) == 0 and len(rx.run_freq(X[:0])) == 0
    info = rx.info()
    assert info["ts_bytes"] == 0 and info["rs_packets"] == 0
    # and the handle still decodes a whole capture afterwards
    assert len(rx.run_baseband(x)) >= 1504 * 4


def test_silence_gives_no_output_and_no_lock(capture):
    import gr_dvbt_b200 as g
    con, cr, tm, tx, x = capture
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    z = np.zeros(200000, np.complex64)
    assert len(rx.run_baseband(z)) == 0
    assert rx.info()["symbols_out"] == 0
    # silence in front of a capture only delays it: same TS as the capture alone
    alone = rx.run_baseband(x)
    both = rx.run_baseband(np.concatenate([z, x]))
    assert len(alone) >= 1504 * 4 and np.array_equal(both[: len(alone)], alone[: len(both)]) and abs(len(both) - len(alone)) <= 2 * 1504


def test_capacity_too_small_is_an_error_and_the_next_stream_is_clean(capture):
    import gr_dvbt_b200 as g
    con, cr, tm, tx, x = capture
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    one = rx.run_baseband(x)
    xs = np.ascontiguousarray(x, np.complex64)
    ts = np.zeros(1504, np.uint8)
    n = C.c_size_t(0)
    rc = g.capi.lib().dvbt_b200_rx_run_baseband_host(rx._h, xs.ctypes.data, len(xs), ts.ctypes.data, len(ts), C.byref(n))
    assert rc != 0 and b"ts_capacity" in g.capi.lib().dvbt_b200_last_error()
    assert np.array_equal(rx.run_baseband(x), one)                   # nothing of the failed call leaks into the next stream


def test_level_change_in_mid_stream_is_refused(capture):
    import gr_dvbt_b200 as g
    con, cr, tm, tx, x = capture
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    one = rx.run_baseband(x)
    rx.stream_reset()
    rx.stream_push("baseband", x[:100000])
    with pytest.raises(Exception):
        rx.stream_push("freq", tx["X"][:4])
    # the error ended that stream: a new one starts from scratch
    assert np.array_equal(rx.run_baseband(x), one)


def test_mode_switch_resets_the_stream(capture):
    import gr_dvbt_b200 as g
    con, cr, tm, tx, x = capture
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    one = rx.run_baseband(x)
    rx.stream_reset()
    rx.stream_push("baseband", x[:300000])
    rx.set_soft_decision(True)                                       # drops the half-fed stream
    parts = [rx.stream_push("baseband", x[:300000]), rx.stream_push("baseband", x[300000:], end=True)]
    soft = np.concatenate(parts)
    assert np.array_equal(soft[: len(one)], one)                     # noise-free: soft and hard decisions agree
    rx.set_soft_decision(False)
    assert np.array_equal(rx.run_baseband(x), one)
