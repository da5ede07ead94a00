"""CPU tests: the C restatements (oracle/port) against the committed golden fixtures that were
generated from the reference itself (tests/golden/make_golden.py), and — where oracle/_ref was built —
against the reference build directly."""
import os

import numpy as np
import pytest

from oracle import port as O, refchain as R

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hotpath_golden.npz"))
CON = {2: 0, 4: 1, 6: 2}


@pytest.mark.parametrize("rate", range(5))
@pytest.mark.parametrize("m", [2, 4, 6])
@pytest.mark.parametrize("ber", [0, 2])
def test_viterbi_port_matches_golden(rate, m, ber):
    rx = G["vit_in_r%d_m%d_b%d" % (rate, m, ber)]
    want = G["vit_out_r%d_m%d_b%d" % (rate, m, ber)]
    got = O.Viterbi(m, rate).work(rx)
    assert np.array_equal(got, want)


def test_viterbi_port_block_calls_and_reset():
    """general_work semantics: ntraceback bytes are withheld once after each reset"""
    rate, m = 2, 4
    k, n = O.RATE_KN[rate]
    data = np.random.default_rng(5).integers(0, 256, 6 * 96 * k, dtype=np.uint8)
    rx = O.conv_encode(data, m, rate)
    v = O.Viterbi(m, rate)
    a = v.work(rx[: 2 * v.in_per_block])
    b = v.work(rx[2 * v.in_per_block:])
    assert len(a) == 2 * v.out_per_block - v.ntb and len(b) == 4 * v.out_per_block
    whole = O.Viterbi(m, rate).work(rx)
    assert np.array_equal(np.concatenate([a, b]), whole)
    assert np.array_equal(whole, data[: len(whole)])
    v.reset()
    c = v.work(rx[: v.in_per_block])
    assert np.array_equal(c, data[: len(c)])


@pytest.mark.parametrize("key,rate,con", [("r0_c1", 0, 1), ("r4_c2", 4, 2), ("r2_c0", 2, 0)])
def test_conv_encoder_matches_reference_inner_coder(key, rate, con):
    m = R.BITS_PER_CELL[con]
    assert np.array_equal(O.conv_encode(G["ic_in_" + key], m, rate), G["ic_out_" + key])


def test_rs_port_matches_golden_both_builds():
    rx = G["rs_rx"]
    out_fixed, st = O.rs_decode(rx, as_built=False)
    out_asb, _ = O.rs_decode(rx, as_built=True)
    assert np.array_equal(out_fixed, G["rs_out_fixed"])
    assert np.array_equal(out_asb, G["rs_out_asbuilt"])
    nerr = np.arange(len(rx)) % 12
    good = nerr <= 8
    assert np.array_equal(out_fixed[good], G["rs_data"][good])
    assert (st[good] == nerr[good]).all() and (st[~good] == -1).all()
    assert np.array_equal(O.rs_encode(G["rs_data"]), G["rs_codewords"])


@pytest.mark.parametrize("con", [0, 1, 2])
def test_demap_port_matches_golden(con):
    assert np.array_equal(O.demap(G["demap_in_c%d" % con], con), G["demap_out_c%d" % con])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_ports_match_the_reference_build_on_fresh_inputs():
    rate, m = 4, 6
    k, n = O.RATE_KN[rate]
    data = np.random.default_rng(99).integers(0, 256, 4 * 96 * k, dtype=np.uint8)
    rx = O.flip_bits(O.conv_encode(data, m, rate), m, 0.006, 2)
    ref, _ = R.rx_viterbi(rx, CON[m], rate, None, blocks_per_call=3)
    assert np.array_equal(O.Viterbi(m, rate).work(rx), ref)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_reference_round_trip_on_random_ts():
    """the reference chain itself (frequency-domain loopback) reproduces the transmitted TS"""
    from dvbt_testlib import tx_frequency_domain, channel
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    tx = tx_frequency_domain(con, cr, tm, 300, 1)
    X = channel(tx["X"])
    Y, tags = R.rx_demod(X, con, cr, tm)
    dm = R.rx_demap(Y, con, tm)
    sd, bd = R.rx_deinterleave(dm, tags, con, tm)
    sf = [t for t in tags if t[1] == "superframe_start"][0][0]
    vo, vt = R.rx_viterbi(bd, con, cr, sf * 1512)
    cd, rd, ts = R.rx_outer(vo, vt)
    assert len(ts) >= 1504
    assert np.array_equal(ts, tx["ts"][504 * 188: 504 * 188 + len(ts)])


# ---- the rest of the chain: restatements against the committed reference outputs -----------------
CH = np.load(os.path.join(os.path.dirname(__file__), "golden", "chain_2k_qam16_r12.npz"))


def test_demod_port_matches_reference_fixture():
    Y, si, tag = O.demod(CH["X"], 1, 0)
    assert Y.shape[0] == int(CH["n_out"]) and tag == 0
    assert np.array_equal(si, CH["symbol_index"])
    assert np.array_equal(Y[:3].view(np.uint32), CH["cells_head"].view(np.uint32))   # bit-exact floats
    assert np.array_equal(O.demap(Y, 1).reshape(Y.shape[0], -1), CH["demap"])


def test_glue_ports_match_reference_fixture():
    sd = O.symbol_deinterleave(CH["demap"], 0, CH["symbol_index"])
    assert np.array_equal(sd[:4], CH["sym_deint_head"])
    bd = O.bit_deinterleave(sd, 4)
    assert np.array_equal(bd, CH["bit_deint"].reshape(-1))
    vo = O.Viterbi(4, 0).work(bd)
    assert np.array_equal(vo, CH["viterbi"][: len(vo)]) and len(vo) >= len(CH["viterbi"]) - 96
    cd = O.conv_deinterleave(CH["viterbi"])
    assert np.array_equal(cd[: len(CH["conv_deint_head"])], CH["conv_deint_head"])
    rs, st = O.rs_decode(cd[: len(cd) // 204 * 204].reshape(-1, 204))
    assert np.array_equal(rs.reshape(-1)[: len(CH["rs"])], CH["rs"])
    ts, first = O.descramble(rs)
    assert first == 11 and np.array_equal(ts[: len(CH["ts"])], CH["ts"])
    assert np.array_equal(ts, CH["ts_source"][: len(ts)])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_acquisition_port_matches_the_reference_build():
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate
    tx = tx_frequency_domain(R.QAM16, R.C1_2, R.T2k, 24, 2)
    x = ofdm_modulate(tx["X"][:24], R.T2k, offset=1301, cfo_bins=0.11, seed=1)
    ref, cons_ref, tags = R.rx_acquisition(x, R.T2k)
    got, cons, tag = O.acquisition(x, 2048, 64)
    assert cons == cons_ref and tag and np.array_equal(ref.view(np.uint32), got.view(np.uint32))


# ---- the whole restated chain against the reference build, mode by mode --------------------------
# (the fixture above pins 2k / QAM16 / 1/2 only; here the reference itself is run, where oracle/_ref is built)
PORT_CHAIN_CASES = [
    (R.QPSK, R.C2_3, R.T2k, 300, 0.0),
    (R.QAM64, R.C7_8, R.T2k, 300, 0.0),      # BASELINE configs[1]
    (R.QAM16, R.C3_4, R.T2k, 300, 0.12),     # Viterbi and RS correcting
    (R.QAM64, R.C7_8, R.T8k, 280, 0.0),      # configs[2]
    (R.QAM16, R.C1_2, R.T8k, 280, 0.0),      # configs[3]
    (R.QPSK, R.C5_6, R.T8k, 280, 0.05),
]


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("con,cr,tm,nsym,noise", PORT_CHAIN_CASES)
def test_port_chain_matches_the_reference_build_stage_by_stage(con, cr, tm, nsym, noise):
    """every restatement in oracle/port against the reference block it restates, on the same input, every stage
    bit for bit (floats included), in both transmission modes, all three constellations, with and without noise"""
    from dvbt_testlib import tx_frequency_domain, channel
    N, P, _, _ = R.mode_dims(tm)
    m = R.BITS_PER_CELL[con]
    tx = tx_frequency_domain(con, cr, tm, nsym, 21)
    X = channel(tx["X"], noise=noise, seed=4)
    Yr, tags = R.rx_demod(X, con, cr, tm)
    Yp, si, tag = O.demod(X, con, tm)
    assert Yp.shape == Yr.shape and np.array_equal(Yp.view(np.uint32), Yr.view(np.uint32))
    assert np.array_equal(si, np.array([t[2] for t in tags if t[1] == "symbol_index"], si.dtype))
    sf = [t for t in tags if t[1] == "superframe_start"][0][0]
    assert tag == sf
    dmr = R.rx_demap(Yr, con, tm)
    dmp = O.demap(Yp, con).reshape(Yp.shape[0], -1)
    assert np.array_equal(dmp, dmr)
    sdr, bdr = R.rx_deinterleave(dmr, tags, con, tm)
    bdp = O.bit_deinterleave(O.symbol_deinterleave(dmp, tm, si), m)
    assert np.array_equal(bdp.reshape(-1), bdr.reshape(-1))
    vor, vtags = R.rx_viterbi(bdr, con, cr, sf * P)
    vop = O.Viterbi(m, cr).work(bdp.reshape(-1)[sf * P:])
    nv = min(len(vor), len(vop))
    assert nv > 2000 and np.array_equal(vop[:nv], vor[:nv])
    for fixed in (False, True):
        cdr, rdr, tsr = R.rx_outer(vor, vtags, fixed_rs=fixed)
        cdp = O.conv_deinterleave(vor)
        rsp, st = O.rs_decode(cdp[: len(cdp) // 204 * 204].reshape(-1, 204), as_built=not fixed)
        tsp, first = O.descramble(rsp)
        assert len(tsr) >= 1504 and np.array_equal(tsp[: len(tsr)], tsr)
        if noise == 0.0:
            src = tx["ts"]
            k0 = [c for c in range(0, len(src) // 188, 8) if np.array_equal(src[c * 188: c * 188 + 1504], tsr[:1504])]
            assert k0 and np.array_equal(tsr, src[k0[0] * 188: k0[0] * 188 + len(tsr)])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("con,hier,alpha", [(1, 2, 2), (1, 3, 4), (2, 2, 2), (2, 3, 4), (1, 1, 1)])
def test_demap_port_matches_the_reference_build_on_hierarchical_constellations(con, hier, alpha):
    """dvbt_demap with hierarchy ALPHA1 / ALPHA2 / ALPHA4 (non-uniform constellations): the reference block itself against
    the restatement, which in turn checks the CUDA table and kernel (tests/test_demap_host_emul_cpu.py)"""
    pts = O.constellation_points(con, alpha, 1.0)
    rng = np.random.default_rng(con * 10 + alpha)
    n = 1512 * 2
    c = (pts[rng.integers(0, len(pts), n)] + (rng.normal(0, 0.2, n) + 1j * rng.normal(0, 0.2, n))).astype(np.complex64)
    mids = ((pts[:, None] + pts[None, :]) / 2).reshape(-1).astype(np.complex64)
    c[: min(len(mids), n)] = mids[:n]
    ref = np.zeros(n, np.uint8)
    R.RefBlock("dvbt_demap", 1512, con, hier, R.T2k, 1.0).work(2, 2, np.ascontiguousarray(c), ref)
    assert np.array_equal(O.demap(c, con, alpha, 1.0), ref)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_descrambler_call_by_call_matches_reference_with_broken_nsync():
    """energy_descramble keeps one piece of state (d_index) and re-checks NSYNC once per call (energy_descramble_impl.cc:
    121-141): the restatement at the scheduler's smallest call size against the reference block driven the same way, on a
    packet stream whose NSYNC bytes are partly destroyed and whose 8-packet phase jumps in the middle (what a mid-stream
    re-synchronisation of the outer deinterleaver looks like from here)"""
    rng = np.random.default_rng(5)
    src = CH["rs"].reshape(-1, 188)[:160].copy()
    assert src[0, 0] == 0xB8 or (src[:, 0] == 0xB8).any()
    a = int(np.flatnonzero(src[:, 0] == 0xB8)[0])
    stream = np.concatenate([rng.integers(0, 0xB0, (5, 188), dtype=np.uint8), src[a: a + 64], src[a + 67: a + 67 + 77]])   # lead junk, then a phase jump of 3
    nsync = np.flatnonzero(stream[:, 0] == 0xB8)
    for kill in ([], [nsync[1]], [nsync[2], nsync[3]], list(nsync[4:7])):
        pk = stream.copy()
        pk[kill, 0] = 0x00
        n = len(pk) // 8 * 8
        items = n // 8
        out = np.zeros(n * 188 + 16, np.uint8)
        b = R.RefBlock("energy_descramble", 8)
        oo = 0
        while items - b.nread >= 4:
            r, cons = b.work(4 * 1504, items - b.nread, pk.ctypes.data + b.nread * 1504, out.ctypes.data + oo)
            oo += max(r, 0)
            if cons == 0:
                break
        got, used, pkidx, first = O.descramble_calls(pk[:n])
        assert used == b.nread and len(got) == oo and np.array_equal(got, out[:oo])
        # in two pieces, carrying d_index: the same bytes
        cut = 40
        g1, u1, p1, f1 = O.descramble_calls(pk[:cut])
        g2, u2, p2, f2 = O.descramble_calls(pk[8 * u1: n], pk=p1)
        assert np.array_equal(np.concatenate([g1, g2]), got)
