"""BASELINE.json configs[0] on the GPU: the reference's own apps/test.ts (head committed as
tests/golden/apps_test_ts_head.npz, see tests/test_apps_test_ts_cpu.py), transmitted by the reference's TX blocks
(oracle/_ref), decoded by the fused CUDA receive chain: the transport stream must be apps/test.ts again from packet
504 on, and at least as long as what the reference RX chain returns for the same symbols."""
import os

import numpy as np
import pytest

from oracle import refchain as R

pytestmark = pytest.mark.gpu
FX = np.load(os.path.join(os.path.dirname(__file__), "golden", "apps_test_ts_head.npz"))


@pytest.mark.skipif(not R.available(), reason="oracle/_ref (reference TX blocks) not built")
def test_cuda_chain_reproduces_test_ts_from_packet_504():
    import gr_dvbt_b200 as g
    from dvbt_testlib import channel
    head = FX["ts_head"]
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    ed, rs, ci = R.tx_outer(head)
    X = channel(R.tx_inner(ci, con, cr, tm, nsym=None)["X"])
    rx = g.rx_chain(con, g.NH, cr, g.G1_32, tm)
    ts = rx.run_freq(X)
    k0 = int(FX["first_packet"])
    assert len(ts) >= int(FX["reference_rx_bytes"])
    assert np.array_equal(ts, head[k0 * 188: k0 * 188 + len(ts)])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not built")
def test_cuda_chain_config3_8k_qam16_rate12_stage_by_stage():
    """BASELINE.json configs[3] (8k / QAM16 / rate 1/2) through the stage-by-stage comparison of tests/test_rx_chain_gpu.py;
    the reference chain starts its TS at packet 2016 in this mode (re-derived with oracle/_ref on the CPU)."""
    import test_rx_chain_gpu as T
    T.test_chain_matches_reference_stage_by_stage(R.QAM16, R.C1_2, R.T8k, 290, 2016)
