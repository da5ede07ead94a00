"""Soft-decision mode of the fused receive chain (dvbt_b200_rx_set_soft_decision) - BEYOND the reference, which demaps
and decodes hard decisions only (lib/dvbt_demap_impl.cc:167-203, lib/d_metrics.c:57-74 a stub, TODO.txt:25).  No reference
output exists for it; what the tests hold on to:
  * noise-free input: the TS of the soft chain is the TS of the hard chain (= the reference's);
  * the sign of every non-zero soft value is the hard decision of the reference demapper;
  * the soft values against a float64 model of the documented rule (tolerance: one quantisation step, stated below);
  * from the soft values on, integer work: the chain's Viterbi output is oracle/port's scalar soft decoder on the same
    values, bit for bit;
  * the point of the mode: under noise, and under a frequency-selective channel, fewer wrong TS packets."""
import numpy as np
import pytest

from oracle import port as O
from oracle import refchain as R

pytestmark = pytest.mark.gpu


def soft_model(cells, con, scale=4.0):
    """float64 model of demap_cell_soft for a flat channel (channel-state weight 1): nibble values - 8, shape (ncell, m)"""
    m = 2 * (con + 1)
    H = m // 2
    L = 1 << H
    norm = {2: 1 / np.sqrt(2), 4: 1 / np.sqrt(10), 6: 1 / np.sqrt(42)}[m]
    # DVB-T Gray mapping along one axis (dvbt_demap_impl.cc:117-165): first bit = sign (1 = negative), then the folds
    lv = np.arange(-(L - 1), L, 2, dtype=np.float64) * norm
    bits = np.zeros((L, H), int)
    for k, n in enumerate(range(-(L - 1), L, 2)):
        a = abs(n)
        bits[k, 0] = 1 if n < 0 else 0
        if H >= 2:
            bits[k, 1] = 1 if a < L // 2 else 0 if H == 2 else (1 if a < 4 else 0)
        if H == 3:
            bits[k, 2] = 1 if 2 < a < 6 else 0
    out = np.zeros((len(cells), m))
    for a, x in enumerate((cells.real.astype(np.float64), cells.imag.astype(np.float64))):
        d = (x[:, None] - lv[None, :]) ** 2
        for j in range(H):
            one = bits[:, j] == 1
            d0 = d[:, ~one].min(axis=1)
            d1 = d[:, one].min(axis=1)
            out[:, 2 * j + a] = (d0 - d1) * scale / (4 * norm * norm)
    return np.clip(np.rint(out), -6, 6)


def nibbles(words, m):
    return ((words[:, None] >> (4 * np.arange(m))) & 15).astype(int) - 8


MODES = [(R.QAM16, R.C1_2, R.T2k), (R.QAM64, R.C7_8, R.T2k), (R.QPSK, R.C2_3, R.T2k), (R.QAM64, R.C3_4, R.T8k)]


@pytest.mark.parametrize("con,cr,tm", MODES, ids=["2k-qam16-1/2", "2k-qam64-7/8", "2k-qpsk-2/3", "8k-qam64-3/4"])
def test_noise_free_soft_chain_gives_the_hard_chain_ts(con, cr, tm):
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, channel
    nsym = 420 if tm == R.T2k else 300
    tx = tx_frequency_domain(con, cr, tm, nsym, 3)
    X = channel(tx["X"], noise=0.0)
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    hard = rx.run_freq(X)
    hard_dm = rx.stage("demap")
    rx.set_soft_decision(True)
    soft = rx.run_freq(X)
    assert len(hard) >= 1504 * 4 and np.array_equal(soft, hard)
    m = 2 * (con + 1)
    v = nibbles(rx.stage("soft_cells"), m)
    hb = ((hard_dm[:, None] >> np.arange(m - 1, -1, -1)) & 1).astype(int)
    assert v.shape == hb.shape and np.all(v != 0) and np.array_equal(v > 0, hb == 1)
    with pytest.raises(Exception):
        rx.stage("demap")
    rx.set_soft_decision(False)
    assert np.array_equal(rx.run_freq(X), hard)


@pytest.mark.parametrize("con,cr", [(R.QAM16, R.C1_2), (R.QAM64, R.C3_4), (R.QPSK, R.C7_8)], ids=["qam16", "qam64", "qpsk"])
def test_soft_values_follow_the_documented_rule_and_the_decoder_is_exact(con, cr):
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, channel
    tm = R.T2k
    m = 2 * (con + 1)
    tx = tx_frequency_domain(con, cr, tm, 300, 5)
    # noise on every carrier but the boosted pilots: the channel estimate is then exact, the channel-state weight of every
    # cell is 1 and the float64 model below has everything it needs
    X0 = channel(tx["X"], noise=0.0)
    pil = (tx["X"].imag == 0) & (np.abs(np.abs(tx["X"].real) - 4.0 / 3.0) < 1e-5)
    Xn = channel(tx["X"], noise=0.06, seed=9)
    X = np.where(pil, X0, Xn).astype(np.complex64)
    # hard chain on the same input: equalised cells and the reference demapper's decisions
    rxh = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    rxh.run_freq(X)
    cells = rxh.stage("cells")
    hard_dm = rxh.stage("demap")
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    rx.set_soft_decision(True)
    ts = rx.run_freq(X)
    v = nibbles(rx.stage("soft_cells"), m)
    # (a) signs = the reference demapper's decisions
    hb = ((hard_dm[:, None] >> np.arange(m - 1, -1, -1)) & 1).astype(int)
    nz = v != 0
    assert np.array_equal((v > 0)[nz], (hb == 1)[nz])
    # (b) values = the documented rule.  Tolerance: the kernel computes in float32 (the weight of a cell is 1 only up to
    # rounding), the model in float64, so a value may land on the other side of a rounding boundary: at most one step, on
    # fewer than 0.1 % of the bits
    assert len(cells) == len(v)
    model = soft_model(cells, con)
    diff = np.abs(model - v)
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3, (diff.max(), (diff > 0).mean())
    assert (np.abs(v) < 6).mean() > 0.2          # and the noise makes the test mean something: many unsaturated values
    # (c) integer work from the values on: deinterleaved values -> scalar soft decoder == the chain's Viterbi output
    vals = rx.stage("soft_values")
    vit = rx.stage("viterbi")
    k, nn = O.RATE_KN[cr]
    nblocks = len(vals) // (768 * nn)
    ref = O.viterbi_soft(vals[: nblocks * 768 * nn], cr)
    assert len(vit) == len(ref) and len(ref) > 1000 and np.array_equal(vit, ref)
    # (d) and the deinterleaved values are the cells' values moved by the reference's index maps: the hard bits of the
    # same positions, pushed through the reference deinterleavers, agree in sign
    info = rx.info()
    assert info["ts_bytes"] == len(ts)


def lost_packets(ts, clean, src):
    """packets of the noise-free run that the run `ts` did not deliver intact (the descrambler drops whole groups when it
    loses the inverted sync byte, so positions do not line up: packets are looked up by content)"""
    srcset = {bytes(p) for p in src[: len(src) // 188 * 188].reshape(-1, 188)}
    good = sum(1 for p in ts[: len(ts) // 188 * 188].reshape(-1, 188) if bytes(p) in srcset)
    return len(clean) // 188 - good, len(clean) // 188


@pytest.mark.parametrize("con,cr,noise", [(R.QAM16, R.C1_2, 0.21), (R.QAM64, R.C3_4, 0.08), (R.QAM64, R.C7_8, 0.062)],
                         ids=["qam16-1/2", "qam64-3/4", "qam64-7/8"])
def test_soft_decisions_lose_fewer_packets_under_noise(con, cr, noise):
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, channel
    tm = R.T2k
    tx = tx_frequency_domain(con, cr, tm, 600, 12)
    X = channel(tx["X"], noise=noise, seed=21)
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    hard = rx.run_freq(X)
    rx.set_soft_decision(True)
    soft = rx.run_freq(X)
    clean = g.rx_chain(con, g.NH, cr, g.G1_32, tm).run_freq(channel(tx["X"], noise=0.0))
    lh, n = lost_packets(hard, clean, tx["ts"])
    ls, _ = lost_packets(soft, clean, tx["ts"])
    assert n > 200 and lh >= 20, (lh, n)    # the hard chain loses packets at this noise level ...
    assert ls * 4 <= lh, (ls, lh, n)        # ... the soft chain at most a quarter of them


def test_channel_state_weight_helps_on_a_frequency_selective_channel():
    """two-path channel with deep notches + noise: cells in a notch are weak, the weight |H|^2 tells the decoder"""
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, channel
    con, cr, tm = R.QAM16, R.C2_3, R.T2k
    tx = tx_frequency_domain(con, cr, tm, 600, 14)
    N = 2048
    k = np.arange(N) - N // 2
    # echo 3 samples late at -1.4 dB: notches 16 dB deep every 683 carriers (slow enough for the reference's linear
    # interpolation between pilots 12 carriers apart; with no noise both modes deliver every packet)
    Hf = (1.0 + 0.85 * np.exp(-2j * np.pi * k * 3 / N)).astype(np.complex64)
    X = channel((tx["X"] * Hf[None, :]).astype(np.complex64), noise=0.13, seed=5)
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    hard = rx.run_freq(X)
    rx.set_soft_decision(True)
    soft = rx.run_freq(X)
    clean = g.rx_chain(con, g.NH, cr, g.G1_32, tm).run_freq(channel((tx["X"] * Hf[None, :]).astype(np.complex64), noise=0.0))
    assert lost_packets(clean, clean, tx["ts"])[0] == 0
    lh, n = lost_packets(hard, clean, tx["ts"])
    ls, _ = lost_packets(soft, clean, tx["ts"])
    assert n > 200 and lh >= 20 and ls * 4 <= lh, (ls, lh, n)


def test_soft_stream_in_pieces_equals_one_shot():
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, channel
    con, cr, tm = R.QAM64, R.C5_6, R.T2k
    tx = tx_frequency_domain(con, cr, tm, 420, 8)
    X = channel(tx["X"], noise=0.05, seed=3)
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    rx.set_soft_decision(True)
    one = rx.run_freq(X)
    rx.stream_reset()
    cuts = [0, 37, 38, 150, 151, 290, len(X)]
    parts = [rx.stream_push("freq", X[a:b], end=(b == len(X))) for a, b in zip(cuts[:-1], cuts[1:])]
    got = np.concatenate(parts)
    assert len(one) > 1504 * 4 and np.array_equal(got[: len(one)], one) and len(got) >= len(one)
