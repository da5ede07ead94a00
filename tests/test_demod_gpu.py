"""GPU parity: CUDA demod_reference_signals vs the reference block itself (oracle/_ref, the
reference sources compiled verbatim) on reference-TX-generated OFDM symbols."""
import numpy as np
import pytest

from oracle import refchain as R

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (needs /root/reference at build time)")


def run_case(con, cr, tm, nsym, noise, shift, seed):
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, channel
    N, P, _, _ = R.mode_dims(tm)
    tx = tx_frequency_domain(con, cr, tm, nsym, seed)
    X = channel(tx["X"], noise=noise, bin_shift=shift, seed=seed)
    Yref, tags_ref = R.rx_demod(X, con, cr, tm)
    d = g.demod_reference_signals(8, N, P, con, g.NH, cr, cr, g.G1_32, tm, 0, 0)
    Y, cons, tags = d.general_work(X, tags=[(0, "sync_start", 1)])
    assert cons == X.shape[0] - 1
    assert sorted(tags) == sorted(tags_ref), (tags[:3], tags_ref[:3])
    assert Y.shape == Yref.shape
    same = np.array_equal(Y.view(np.uint32), Yref.view(np.uint32))
    if not same:
        err = np.abs(Y - Yref) / (np.abs(Yref) + 1e-12)
        bad = np.argwhere(Y.view(np.uint64) != Yref.view(np.uint64))
        print("max rel err %.3e, %d of %d cells differ, first at %s" % (err.max(), len(bad), Y.size, bad[:3].tolist()))
    return Y, Yref, same


@needs_ref
@pytest.mark.parametrize("con,cr,tm,nsym", [(R.QAM16, R.C1_2, R.T2k, 300), (R.QAM64, R.C7_8, R.T2k, 292), (R.QAM16, R.C1_2, R.T8k, 284)])
def test_clean_channel_bit_exact(con, cr, tm, nsym):
    Y, Yref, same = run_case(con, cr, tm, nsym, 0.0, 0, 1)
    assert Y.shape[0] >= 8
    assert same


@needs_ref
def test_noise_and_integer_offset():
    """AWGN (decisions near ties, TPS votes) and a +3 bin carrier offset (integer CFO path)."""
    Y, Yref, same = run_case(R.QAM64, R.C7_8, R.T2k, 292, 0.05, 3, 7)
    # sincosf/atan2f of the rotor may differ in the last ulp between glibc and CUDA: tolerance 1e-5 relative
    err = np.abs(Y - Yref) / (np.abs(Yref) + 1e-9)
    assert err.max() < 1e-5


@needs_ref
def test_streaming_calls_carry_state():
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, channel
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    N, P, _, _ = R.mode_dims(tm)
    X = channel(tx_frequency_domain(con, cr, tm, 300, 3)["X"])
    Yref, tags_ref = R.rx_demod(X, con, cr, tm)
    d = g.demod_reference_signals(8, N, P, con, g.NH, cr, cr, g.G1_32, tm, 0, 0)
    pos, outs, tags, nwritten = 0, [], [], 0
    sizes = [5, 17, 64, 3, 129, 40]
    i = 0
    first = True
    while pos < X.shape[0] - 1:
        n = min(sizes[i % len(sizes)], X.shape[0] - 1 - pos)
        i += 1
        Y, cons, t = d.general_work(X[pos: pos + n + 1], tags=[(0, "sync_start", 1)] if first else [])
        first = False
        for off, key, val in t:
            tags.append((nwritten + off, key, val))
        nwritten += Y.shape[0]
        outs.append(Y)
        pos += cons
    Y = np.concatenate(outs)
    assert np.array_equal(Y.view(np.uint32), Yref.view(np.uint32))
    assert sorted(tags) == sorted(tags_ref)


@needs_ref
@pytest.mark.parametrize("damage", ["tps_frame", "dropped_symbol", "zero_symbols"])
def test_lock_loss_in_the_middle_matches_reference(damage):
    """The scan kernel validates whole TPS frames in parallel once in lock; a frame that fails (corrupted TPS
    carriers -> BCH error, a missing symbol -> scattered-pilot phase jump, blank symbols) must hand control
    back to the per-symbol machine exactly where the reference loses and regains lock."""
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, channel
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    N, P, _, _ = R.mode_dims(tm)
    X = channel(tx_frequency_domain(con, cr, tm, 68 * 11 + 20, 5)["X"]).copy()
    rng = np.random.default_rng(9)
    if damage == "tps_frame":
        # flip the sign of a dozen symbols' worth of TPS carriers inside frame 6 (DBPSK votes flip -> BCH fails)
        tps = np.array([34, 50, 209, 346, 413, 569, 595, 688, 790, 901, 1073, 1219, 1262, 1286, 1469, 1594, 1687])
        zl = (N - 1705 + 1) // 2
        for s in range(68 * 6 + 20, 68 * 6 + 32, 2):
            X[s, zl + tps] *= -1
    elif damage == "dropped_symbol":
        X = np.delete(X, 68 * 5 + 33, axis=0)
    else:
        X[68 * 7 + 10: 68 * 7 + 13] = 0
    Yref, tags_ref = R.rx_demod(X, con, cr, tm)
    d = g.demod_reference_signals(8, N, P, con, g.NH, cr, cr, g.G1_32, tm, 0, 0)
    Y, cons, tags = d.general_work(X, tags=[(0, "sync_start", 1)])
    assert cons == X.shape[0] - 1
    assert sorted(tags) == sorted(tags_ref), (len(tags), len(tags_ref))
    assert Y.shape == Yref.shape and Y.shape[0] > 68 * 3
    # blank symbols equalise to NaN (0 * inf); x86 and CUDA produce different NaN bit patterns, so NaNs only have
    # to sit in the same places - every other cell is compared bit for bit
    a, b = Y.view(np.float32), Yref.view(np.float32)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), (int(nan_a.sum()), int(nan_b.sum()))
    assert np.array_equal(a.view(np.uint32)[~nan_a], b.view(np.uint32)[~nan_b])
