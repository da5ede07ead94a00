"""Guard intervals 1/16, 1/8 and 1/4 (lib/dvbt_config.cc:194-208) - every shipped flowgraph uses 1/32, but the blocks
take the guard interval as a parameter: cp_length of ofdm_sym_acquisition (N/16, N/8, N/4), and guard_interval of
demod_reference_signals, where it enters the carrier-frequency estimate (reference_signals_impl.cc:753, :792-819).
Parity against the reference blocks (oracle/_ref) built with the same parameters."""
import numpy as np
import pytest

from oracle import refchain as R

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")

GIS = [(R.G1_16, "1/16"), (R.G1_8, "1/8"), (R.G1_4, "1/4")]


@needs_ref
@pytest.mark.parametrize("gi,name", GIS, ids=[g[1] for g in GIS])
@pytest.mark.parametrize("tm", [R.T2k, R.T8k], ids=["2k", "8k"])
def test_acquisition_with_other_guard_intervals(gi, name, tm):
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate
    con, cr = R.QAM16, R.C1_2
    N, P, K, cp = R.mode_dims(tm, gi)
    nsym = 40 if tm == R.T2k else 16
    tx = tx_frequency_domain(con, cr, tm, nsym, 2, gi=gi)
    x = ofdm_modulate(tx["X"][:nsym], tm, offset=911, cfo_bins=0.0, seed=1, gi=gi)
    ref, cons_ref, tags_ref = R.rx_acquisition(x, tm, gi=gi)
    acq = g.ofdm_sym_acquisition(1, N, K, cp, 30.0)
    out, cons, tags = acq.general_work(x)
    n = min(len(out), len(ref))
    assert n >= nsym - 4 and abs(len(out) - len(ref)) <= 1
    assert cons >= cons_ref and (cons - cons_ref) % (N + cp) == 0
    assert sorted(set(t[0] for t in tags)) == sorted(set(t[0] for t in tags_ref if t[0] < len(out)))
    err = np.abs(out[:n] - ref[:n]).max() / np.abs(ref[:n]).max()
    assert err < 2e-5, err


@needs_ref
@pytest.mark.parametrize("gi,name", GIS, ids=[g[1] for g in GIS])
def test_demod_cells_bit_exact_with_other_guard_intervals(gi, name):
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, channel
    con, cr, tm = R.QAM64, R.C3_4, R.T2k
    N, P, K, cp = R.mode_dims(tm, gi)
    tx = tx_frequency_domain(con, cr, tm, 300, 5, gi=gi)
    X = channel(tx["X"], noise=0.02, seed=2)
    Y, tags = R.rx_demod(X, con, cr, tm, gi=gi)
    dem = g.demod_reference_signals(8, N, P, con, g.NH, cr, cr, gi, tm, 0, 0)
    got, consumed, gtags = dem.general_work(X, tags=[(0, "sync_start", 1)])
    assert got.shape == Y.shape and len(Y) > 10
    assert np.array_equal(got.view(np.uint32), Y.view(np.uint32))
    assert gtags == tags


@needs_ref
@pytest.mark.parametrize("gi,name", GIS, ids=[g[1] for g in GIS])
def test_chain_from_baseband_with_other_guard_intervals(gi, name):
    """time-domain loopback with a longer cyclic prefix, a timing offset and a carrier offset: the transmitted TS, and the
    reference chain's TS (reference acquisition + numpy FFT + reference blocks with the same guard interval)"""
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    N, P, K, cp = R.mode_dims(tm, gi)
    tx = tx_frequency_domain(con, cr, tm, 420, 11, gi=gi)
    x = ofdm_modulate(tx["X"], tm, offset=777, cfo_bins=0.1, seed=4, gi=gi)
    rx = g.rx_chain(con, g.NH, cr, gi, tm)
    ts = rx.run_baseband(x)
    info = rx.info()
    assert info["acq_symbols"] >= 416
    src = tx["ts"]
    assert len(ts) > 1504 * 4 and np.array_equal(ts, src[504 * 188: 504 * 188 + len(ts)])
    sym, cons, atags = R.rx_acquisition(x, tm, gi=gi)
    Xf = np.fft.fftshift(np.fft.fft(sym.astype(np.complex128), axis=1), axes=1).astype(np.complex64)
    sync = sorted(set(o for o, k, v in atags if k == "sync_start"))
    Y, tags = R.rx_demod(Xf, con, cr, tm, sync_offsets=sync, gi=gi)
    dm = R.rx_demap(Y, con, tm)
    sd, bd = R.rx_deinterleave(dm, tags, con, tm)
    sf = [t[0] for t in tags if t[1] == "superframe_start"]
    vo, vtags = R.rx_viterbi(bd, con, cr, [o * P for o in sf], blocks_per_call=1)
    cd, rd, ref_ts = R.rx_outer(vo, vtags, min_calls=True)
    assert len(ref_ts) >= 1504 * 4 and np.array_equal(ts[: len(ref_ts)], ref_ts)
