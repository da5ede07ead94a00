"""GPU parity against the committed golden fixtures (generated from the reference itself by
tests/golden/make_golden.py), so that parity does not depend on oracle/_ref being present."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GDIR = os.path.join(os.path.dirname(__file__), "golden")


def test_chain_stage_by_stage_against_golden():
    import gr_dvbt_b200 as g
    CH = np.load(os.path.join(GDIR, "chain_2k_qam16_r12.npz"))
    rx = g.rx_chain(g.QAM16, g.NH, g.C1_2, g.G1_32, g.T2k)
    ts = rx.run_freq(CH["X"])
    assert rx.info()["symbols_out"] == int(CH["n_out"])
    assert np.array_equal(rx.stage("symbol_index"), CH["symbol_index"])
    cells = rx.stage("cells").reshape(-1, 1512)
    assert np.array_equal(cells[:3].view(np.uint32), CH["cells_head"].view(np.uint32))
    assert np.array_equal(rx.stage("demap").reshape(-1, 1512), CH["demap"])
    assert np.array_equal(rx.stage("bitdeint"), CH["bit_deint"].reshape(-1))
    vit = rx.stage("viterbi")
    n = min(len(vit), len(CH["viterbi"]))
    assert n > 30000 and np.array_equal(vit[:n], CH["viterbi"][:n])
    rs = rx.stage("rs")
    assert np.array_equal(rs[: len(CH["rs"])], CH["rs"])
    assert len(ts) >= len(CH["ts"]) and np.array_equal(ts[: len(CH["ts"])], CH["ts"])
    assert np.array_equal(ts, CH["ts_source"][: len(ts)])


def test_blocks_against_golden():
    import gr_dvbt_b200 as g
    G = np.load(os.path.join(GDIR, "hotpath_golden.npz"))
    for rate in range(5):
        for m in (2, 4, 6):
            dec = g.viterbi_decoder({2: 0, 4: 1, 6: 2}[m], g.NH, rate)
            for ber in (0, 2):
                out = dec.decode(G["vit_in_r%d_m%d_b%d" % (rate, m, ber)])[0]
                assert np.array_equal(out, G["vit_out_r%d_m%d_b%d" % (rate, m, ber)]), (rate, m, ber)
    rs = g.reed_solomon_dec()
    out, _ = rs.general_work(len(G["rs_rx"]) // 8, G["rs_rx"].reshape(-1))
    assert np.array_equal(out.reshape(-1, 188), G["rs_out_fixed"])
    rs.set_compat(1)
    out, _ = rs.general_work(len(G["rs_rx"]) // 8, G["rs_rx"].reshape(-1))
    assert np.array_equal(out.reshape(-1, 188), G["rs_out_asbuilt"])
    for con in (0, 1, 2):
        d = g.dvbt_demap(1512, con, g.NH, g.T2k, 1.0)
        out, _ = d.general_work(2, G["demap_in_c%d" % con])
        assert np.array_equal(out, G["demap_out_c%d" % con])
