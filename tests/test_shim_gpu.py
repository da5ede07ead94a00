"""Drop-in test of the gr::block shims (gr_dvbt_b200/shim): the reference RX chain is driven block
by block exactly as for the reference, but the five hot blocks are created through their public
make() from the shim library (B200 behind the C ABI); every other block is the reference's own.
The transport stream must not change."""
import numpy as np
import pytest

from oracle import refchain as R

pytestmark = pytest.mark.gpu
needs = pytest.mark.skipif(not (R.available() and R.shim_available()), reason="needs oracle/_ref and the shim test build")


@needs
def test_flowgraph_with_hot_blocks_swapped_gives_the_same_ts():
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate
    from test_rx_chain_gpu import reference_rx
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    N, P, K, cp = R.mode_dims(tm)
    tx = tx_frequency_domain(con, cr, tm, 420, 21)
    x = ofdm_modulate(tx["X"], tm, offset=640, seed=2)

    def run():
        sym, cons, tags = R.rx_acquisition(x, tm)
        Xf = np.fft.fftshift(np.fft.fft(sym.astype(np.complex128), axis=1), axes=1).astype(np.complex64)
        return sym, reference_rx(Xf, con, cr, tm, fixed_rs=True)

    sym_ref, ref = run()
    R.SHIM_BLOCKS.update(R.HOT_BLOCKS)
    try:
        sym_b200, got = run()
    finally:
        R.SHIM_BLOCKS.clear()
    assert sym_b200.shape == sym_ref.shape
    assert np.abs(sym_b200 - sym_ref).max() / np.abs(sym_ref).max() < 2e-5
    assert np.array_equal(got["vo"], ref["vo"])
    assert len(ref["ts"]) > 0 and np.array_equal(got["ts"], ref["ts"])
    assert np.array_equal(got["ts"], tx["ts"][504 * 188: 504 * 188 + len(got["ts"])])
    assert [t for t in got["tags"] if t[1] != "symbol_index"] == [t for t in ref["tags"] if t[1] != "symbol_index"]
