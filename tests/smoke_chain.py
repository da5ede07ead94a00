"""The chain half of __graft_entry__.smoke(): one small frequency-domain capture (the committed fixture, made from the
reference itself) through the fused CUDA receive chain, compared stage by stage with the oracle restatements
(oracle/port) run here on the same input, and with the reference's own outputs stored in the fixture."""
import os

import numpy as np


def run():
    import gr_dvbt_b200 as g
    from oracle import port as O
    CH = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chain_2k_qam16_r12.npz"))
    X = CH["X"]
    rx = g.rx_chain(g.QAM16, g.NH, g.C1_2, g.G1_32, g.T2k)
    ts = rx.run_freq(X)
    # the oracle on the same input (demod_reference_signals, dvbt_demap, deinterleavers, viterbi_decoder, outer chain)
    Y, si, tag = O.demod(X, 1, 0)
    dm = O.demap(Y, 1).reshape(Y.shape[0], -1)
    bd = O.bit_deinterleave(O.symbol_deinterleave(dm, 0, si), 4)
    vo = O.Viterbi(4, 0).work(bd)
    rsd, st = O.rs_decode(O.conv_deinterleave(vo)[: len(vo) // 204 * 204].reshape(-1, 204))
    ts_o, first = O.descramble(rsd)
    assert rx.info()["symbols_out"] == Y.shape[0] == int(CH["n_out"])
    assert np.array_equal(rx.stage("cells").view(np.uint32), Y.reshape(-1).view(np.uint32)), "equalised cells differ from the oracle"
    assert np.array_equal(rx.stage("demap"), dm.reshape(-1)), "demapped cells differ from the oracle"
    assert np.array_equal(rx.stage("bitdeint"), bd.reshape(-1)), "inner deinterleavers differ from the oracle"
    vit = rx.stage("viterbi")
    n = min(len(vit), len(vo))
    assert n > 30000 and np.array_equal(vit[:n], vo[:n]), "Viterbi output differs from the oracle"
    n = min(len(ts), len(ts_o))
    assert n >= 1504 and np.array_equal(ts[:n], ts_o[:n]), "transport stream differs from the oracle"
    # and the reference's own outputs for this input (tests/golden/make_golden.py)
    assert len(ts) >= len(CH["ts"]) and np.array_equal(ts[: len(CH["ts"])], CH["ts"])
    assert np.array_equal(ts, CH["ts_source"][: len(ts)]), "transport stream is not the transmitted one"
    return len(ts)
