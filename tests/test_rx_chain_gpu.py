"""GPU parity of the fused receive chain against the reference chain (oracle/_ref: the
reference's own blocks, compiled verbatim, driven block by block), stage by stage, on a
frequency-domain loopback of reference-TX-generated symbols."""
import numpy as np
import pytest

from oracle import refchain as R

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")


def reference_rx(X, con, cr, tm, fixed_rs=False):
    N, P, _, _ = R.mode_dims(tm)
    Y, tags = R.rx_demod(X, con, cr, tm)
    dm = R.rx_demap(Y, con, tm)
    sd, bd = R.rx_deinterleave(dm, tags, con, tm)
    sf = [t for t in tags if t[1] == "superframe_start"][0][0]
    vo, vtags = R.rx_viterbi(bd, con, cr, sf * P)
    cd, rd, ts = R.rx_outer(vo, vtags, fixed_rs=fixed_rs)
    return dict(Y=Y, tags=tags, dm=dm, bd=bd, vo=vo, rd=rd, ts=ts)


CASES = [
    (R.QAM16, R.C1_2, R.T2k, 420, 504),
    (R.QAM64, R.C7_8, R.T2k, 330, 1328),
    (R.QPSK, R.C2_3, R.T2k, 480, None),
    (R.QAM64, R.C3_4, R.T8k, 300, None),
    (R.QAM64, R.C7_8, R.T8k, 290, None),
    (R.QAM16, R.C5_6, R.T8k, 290, None),
]


@needs_ref
@pytest.mark.parametrize("con,cr,tm,nsym,first_ts_packet", CASES)
def test_chain_matches_reference_stage_by_stage(con, cr, tm, nsym, first_ts_packet):
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, channel
    tx = tx_frequency_domain(con, cr, tm, nsym, 11)
    X = channel(tx["X"])
    ref = reference_rx(X, con, cr, tm)
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    ts = rx.run_freq(X)
    info = rx.info()
    N, P, _, _ = R.mode_dims(tm)
    assert info["symbols_out"] == ref["Y"].shape[0]
    assert np.array_equal(rx.stage("cells").view(np.uint32), ref["Y"].reshape(-1).view(np.uint32))
    assert np.array_equal(rx.stage("demap"), ref["dm"].reshape(-1))
    assert np.array_equal(rx.stage("bitdeint"), ref["bd"].reshape(-1))
    sidx = [t[2] for t in ref["tags"] if t[1] == "symbol_index"]
    assert np.array_equal(rx.stage("symbol_index"), np.array(sidx, np.int32))
    vit = rx.stage("viterbi")
    nv = min(len(vit), len(ref["vo"]))
    assert nv > 1000 and np.array_equal(vit[:nv], ref["vo"][:nv])
    rs = rx.stage("rs")
    nr = min(len(rs), len(ref["rd"]))
    assert nr > 0 and np.array_equal(rs[:nr], ref["rd"][:nr])
    # the reference's TS is a (scheduler dependent) prefix of ours
    assert len(ref["ts"]) > 0 and len(ts) >= len(ref["ts"])
    assert np.array_equal(ts[: len(ref["ts"])], ref["ts"])
    # and it is the transmitted TS from the expected packet on
    src = tx["ts"]
    k0 = None
    for cand in range(0, len(src) // 188 - len(ts) // 188 + 1, 8):
        if np.array_equal(ts[:1504], src[cand * 188: cand * 188 + 1504]):
            k0 = cand
            break
    assert k0 is not None
    assert np.array_equal(ts, src[k0 * 188: k0 * 188 + len(ts)])
    if first_ts_packet is not None:
        assert k0 == first_ts_packet
    assert info["viterbi_repaired"] == 0


@needs_ref
def test_chain_with_noise_uses_rs_and_matches_both_reference_builds():
    """AWGN so that Viterbi leaves byte errors and RS has work to do; as-built vs fixed RS."""
    import gr_dvbt_b200 as g
    from dvbt_testlib import tx_frequency_domain, channel
    con, cr, tm = R.QAM16, R.C3_4, R.T2k
    tx = tx_frequency_domain(con, cr, tm, 400, 5)
    X = channel(tx["X"], noise=0.12, seed=3)
    for as_built in (0, 1):
        ref = reference_rx(X, con, cr, tm, fixed_rs=not as_built)
        rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
        rx.set_rs_compat(as_built)
        ts = rx.run_freq(X)
        vit = rx.stage("viterbi")
        nv = min(len(vit), len(ref["vo"]))
        assert np.array_equal(vit[:nv], ref["vo"][:nv])
        st = rx.stage("rs_status")
        assert (st > 0).sum() > 3, "test should exercise the RS correction path"
        rs = rx.stage("rs")
        nr = min(len(rs), len(ref["rd"]))
        assert np.array_equal(rs[:nr], ref["rd"][:nr])
        assert np.array_equal(ts[: len(ref["ts"])], ref["ts"])
