"""The fused chain as a STREAM (include/dvbt_b200.h: dvbt_b200_rx_stream_*) and across a loss of lock.

1. A capture fed in uneven pieces - the way the GNU Radio scheduler feeds the flowgraph - must give, concatenated, the
   transport stream of the one-shot run: every block's state (resampler history, acquisition input buffer and tracking
   state, the symbol demod is still waiting to see the successor of, the cells short of a 768-block, the Viterbi
   decoder, the outer deinterleaver's delay lines, the descrambler's NSYNC index) is carried from call to call.
2. A capture that loses lock in the middle: the reference restarts acquisition after one missed peak
   (lib/ofdm_sym_acquisition_impl.cc:545-558), re-sends sync_start (:507), demod_reference_signals re-arms on it and
   waits for the next superframe start (lib/demod_reference_signals_impl.cc:112-116), whose tag resets the Viterbi
   decoder (lib/viterbi_decoder_impl.cc:213-229) and re-aligns the outer deinterleaver
   (lib/convolutional_deinterleaver_impl.cc:109-120, delay lines not cleared); energy_descramble finds NSYNC again
   (lib/energy_descramble_impl.cc:121-141).  The CUDA chain must deliver the transport stream of the reference chain
   (oracle/_ref blocks driven with the smallest scheduler calls, where the result is well defined) byte for byte."""
import numpy as np
import pytest

from oracle import refchain as R

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")


def reference_stream_rx(x, con, cr, tm, fixed_rs=True):
    """the reference RX chain on baseband samples with every tag carried: acquisition (one symbol per call) -> numpy FFT
    -> demod (sync_start tags where acquisition sent them) -> demap -> deinterleavers -> Viterbi (one 768-block per
    call, every superframe_start tag) -> outer deinterleaver (2 items per call) -> RS -> descrambler (smallest calls)"""
    N, P, K, cp = R.mode_dims(tm)
    sym, cons, tags = R.rx_acquisition(x, tm)
    Xf = np.fft.fftshift(np.fft.fft(sym.astype(np.complex128), axis=1), axes=1).astype(np.complex64)
    sync = sorted(set(o for o, k, v in tags if k == "sync_start"))
    Y, dtags = R.rx_demod(Xf, con, cr, tm, sync_offsets=sync)
    sf = [t[0] for t in dtags if t[1] == "superframe_start"]
    dm = R.rx_demap(Y, con, tm)
    sd, bd = R.rx_deinterleave(dm, dtags, con, tm)
    vo, vtags = R.rx_viterbi(bd, con, cr, [o * P for o in sf], blocks_per_call=1)
    cd, rd, ts = R.rx_outer(vo, vtags, fixed_rs=fixed_rs, min_calls=True)
    return dict(nsym=sym.shape[0], sync=sync, sf=sf, Y=Y, vo=vo, vtags=vtags, rd=rd, ts=ts)


def pieces_of(n, sizes):
    """cut [0, n) into pieces of the given sizes (cycled) - deliberately uneven, some tiny"""
    out, pos, i = [], 0, 0
    while pos < n:
        s = min(sizes[i % len(sizes)], n - pos)
        out.append((pos, pos + s))
        pos += s
        i += 1
    return out


STREAM_CASES = [
    # level, constellation, rate, mode, symbols, piece sizes (in samples / symbols of that level)
    ("file", R.QAM64, R.C7_8, R.T2k, 330, [200001, 7, 123457, 35, 300000, 1, 99999]),
    ("baseband", R.QAM16, R.C1_2, R.T2k, 420, [150000, 2111, 1, 333333, 4096, 77777]),
    ("freq", R.QPSK, R.C7_8, R.T2k, 480, [100, 1, 1, 57, 2, 140, 3]),
    ("file", R.QAM16, R.C1_2, R.T8k, 300, [700001, 9239, 500000, 64, 1234567]),
]


@needs_ref
@pytest.mark.parametrize("level,con,cr,tm,nsym,sizes", STREAM_CASES, ids=["%s-%d-%d-%d" % c[:4] for c in STREAM_CASES])
def test_pieces_give_the_one_shot_transport_stream(level, con, cr, tm, nsym, sizes):
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate, to_capture_rate, channel
    N, P, K, cp = R.mode_dims(tm)
    tx = tx_frequency_domain(con, cr, tm, nsym, 17)
    gain = 0.0022097087 if tm == R.T2k else 0.00055242272
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    if level == "freq":
        data = channel(tx["X"])
        whole = rx.run_freq(data)
        unit = N
    elif level == "baseband":
        data = ofdm_modulate(tx["X"], tm, offset=777, cfo_bins=0.1, seed=4)
        whole = rx.run_baseband(data)
        unit = 1
    else:
        data = to_capture_rate(ofdm_modulate(tx["X"], tm, gain=1.0, offset=500, seed=4))
        whole = rx.run_file(data, gain)
        unit = 1
    assert len(whole) >= 4 * 1504
    flat = np.ascontiguousarray(data).reshape(-1)
    cuts = pieces_of(len(flat) // unit, sizes)
    assert len(cuts) >= 5
    rx2 = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    rx2.stream_reset()
    got = []
    for i, (a, b) in enumerate(cuts):
        got.append(rx2.stream_push(level, flat[a * unit: b * unit], gain=gain, end=(i == len(cuts) - 1)))
    ts = np.concatenate(got)
    assert len(ts) == len(whole) and np.array_equal(ts, whole)
    assert rx2.info()["ts_total"] == len(whole)
    # the same handle again after the stream ended: a new stream, same result (no state leaks)
    again = np.concatenate([rx2.stream_push(level, flat[a * unit: b * unit], gain=gain, end=(i == len(cuts) - 1)) for i, (a, b) in enumerate(cuts)])
    assert np.array_equal(again, whole)


def lossy_capture(con, cr, tm, nsym, zero_at, nzero=3, seed=11):
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate
    N, P, K, cp = R.mode_dims(tm)
    tx = tx_frequency_domain(con, cr, tm, nsym, seed)
    x = ofdm_modulate(tx["X"], tm, offset=777, cfo_bins=0.1, seed=4)
    z0 = 777 + zero_at * (N + cp) + 100
    x[z0: z0 + nzero * (N + cp)] = 0
    return x, tx["ts"]


@needs_ref
@pytest.mark.parametrize("con,cr,tm,nsym,zero_at", [(R.QAM16, R.C1_2, R.T2k, 900, 400), (R.QAM64, R.C7_8, R.T2k, 900, 431)])
def test_lost_lock_mid_capture_gives_the_reference_transport_stream(con, cr, tm, nsym, zero_at):
    import gr_dvbt_b200 as g
    x, src = lossy_capture(con, cr, tm, nsym, zero_at)
    ref = reference_stream_rx(x, con, cr, tm)
    assert len(ref["sync"]) >= 2 and len(ref["sf"]) == 2, "the test capture must make the reference re-synchronise"
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    ts = rx.run_baseband(x)
    info = rx.info()
    assert info["acq_lost_at"] == ref["sync"][1]            # the symbol count at which the peak was missed
    assert info["n_sync_start"] == len(ref["sync"]) and info["n_superframe_start"] == 2
    assert abs(info["acq_symbols"] - ref["nsym"]) <= 1
    vit = rx.stage("viterbi")
    assert len(vit) >= len(ref["vo"]) and np.array_equal(vit[: len(ref["vo"])], ref["vo"])
    # the reference's file is a prefix of ours (the scheduler never calls energy_descramble for the last items)
    assert len(ref["ts"]) >= 100 * 188 and len(ts) >= len(ref["ts"])
    assert np.array_equal(ts[: len(ref["ts"])], ref["ts"])
    # what it means: the transmitted stream before the gap, a few packets of mixed delay-line contents / descrambled with
    # the NSYNC phase of the old alignment, the transmitted stream again from the next superframe
    srcp = src[: len(src) // 188 * 188].reshape(-1, 188)
    index = {p.tobytes(): i for i, p in enumerate(srcp)}
    loc = np.array([index.get(p.tobytes(), -1) for p in ts.reshape(-1, 188)])
    bad = np.flatnonzero(loc < 0)
    assert 0 < len(bad) <= 32 and bad[-1] - bad[0] < 32
    assert np.all(np.diff(loc[: bad[0]]) == 1) and np.all(np.diff(loc[bad[-1] + 1:]) == 1)
    assert loc[bad[-1] + 1] > loc[bad[0] - 1] + 12

    # and the same capture as a stream, cut inside the gap and around the re-acquisition
    N, P, K, cp = R.mode_dims(tm)
    z = 777 + zero_at * (N + cp)
    cuts = [0, 100000, z + 500, z + 2 * (N + cp), z + 5 * (N + cp) + 17, z + 40 * (N + cp), len(x) - 12345, len(x)]
    rx2 = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    got = [rx2.stream_push("baseband", x[a:b], end=(b == len(x))) for a, b in zip(cuts[:-1], cuts[1:])]
    assert np.array_equal(np.concatenate(got), ts)


@needs_ref
def test_acquisition_block_tags_every_reacquisition():
    """dvbt_b200_acq_work sends sync_start on the first item after EVERY (re)acquisition: same offsets as the
    reference block (send_sync_start at nitems_written on every attempt, ofdm_sym_acquisition_impl.cc:507)"""
    import gr_dvbt_b200 as g
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    N, P, K, cp = R.mode_dims(tm)
    x, _ = lossy_capture(con, cr, tm, 120, 50)
    ref, cons_ref, tags_ref = R.rx_acquisition(x, tm)
    want = sorted(set(o for o, k, v in tags_ref if k == "sync_start"))
    assert len(want) == 2
    acq = g.ofdm_sym_acquisition(1, N, K, cp, 30.0)
    out, cons, tags = acq.general_work(x)
    assert [t[0] for t in tags] == want and all(t[1] == "sync_start" for t in tags)
    assert abs(len(out) - len(ref)) <= 1
    # in two calls, split inside the gap: the second call's tag is on the item it belongs to
    acq2 = g.ofdm_sym_acquisition(1, N, K, cp, 30.0)
    cut = 777 + 51 * (N + cp)
    o1, c1, t1 = acq2.general_work(x[:cut])
    o2, c2, t2 = acq2.general_work(x[c1:])
    offs = [t[0] for t in t1] + [len(o1) + t[0] for t in t2]
    assert sorted(set(offs)) == want


@needs_ref
def test_two_lock_losses_under_noise_in_pieces():
    """25 dB AWGN (Viterbi and RS both correcting), two gaps - the second one only a symbol and a half long - and the capture
    fed in pieces: the reference chain's transport stream, byte for byte"""
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate
    con, cr, tm = R.QAM16, R.C2_3, R.T2k
    N, P, K, cp = R.mode_dims(tm)
    tx = tx_frequency_domain(con, cr, tm, 1300, 23)
    x = ofdm_modulate(tx["X"], tm, offset=911, cfo_bins=0.17, noise=10 ** (-25 / 20), seed=8)
    for at, n in ((380, 3.0), (830, 1.5)):
        z0 = 911 + at * (N + cp) + 333
        x[z0: z0 + int(n * (N + cp))] = 0
    ref = reference_stream_rx(x, con, cr, tm, fixed_rs=True)
    assert len(ref["sf"]) == 3 and len(ref["ts"]) > 300 * 188
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    whole = rx.run_baseband(x)
    assert rx.info()["n_superframe_start"] == 3
    assert len(whole) >= len(ref["ts"]) and np.array_equal(whole[: len(ref["ts"])], ref["ts"])
    cuts = pieces_of(len(x), [400001, 2112, 77, 650000, 31, 123456])
    rx2 = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    got = [rx2.stream_push("baseband", x[a:b], end=(b == len(x))) for a, b in cuts]
    assert np.array_equal(np.concatenate(got), whole)
