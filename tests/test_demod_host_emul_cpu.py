"""demod_reference_signals on the CPU from the kernels' own source (gr_dvbt_b200/csrc/demod.cu: stage 1, equalise + fused
demap, vote, scan, and the host-side mode tables), compiled for the host (tests/emul/: every float / double operation
rounded on its own, libm's atan2f / sincosf standing in for CUDA's) against the oracle restatement of
reference_signals_impl.cc - equalised cells bit for bit, demapped cells, symbol_index tags, superframe gating."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from oracle import port as O, refchain as R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import build_vit_emul  # noqa: E402

# a kernel that is not warp-converged would dead-lock the lock-step emulation: never hang the suite (the host threads sit
# inside a C call, so only the thread method of pytest-timeout can end the run)
pytestmark = pytest.mark.timeout(900, method="thread")

CH = np.load(os.path.join(os.path.dirname(__file__), "golden", "chain_2k_qam16_r12.npz"))


@pytest.fixture(scope="module", params=[0, 1], ids=["stage1+symbol", "fused-front"])
def demod(request):
    """both launch forms of demod_run: demod_stage1_kernel + demod_symbol_kernel<false> (default) and the opt-in
    demod_symbol_kernel<true> that does stage 1 inside the block (DVBT_B200_DEMOD_FUSED=1)"""
    os.environ["DVBT_EMUL_DEMOD_FUSED"] = str(request.param)
    lib = C.CDLL(build_vit_emul.build_demod())

    def run(X, con, tm):
        N, P = (2048, 1512) if tm == 0 else (8192, 6048)
        X = np.ascontiguousarray(X, np.complex64).reshape(-1, N)
        nsym = X.shape[0]
        Xp = np.zeros((nsym + 1) * N + 64, np.complex64)     # the kernels read a few bins around an item, like the reference
        Xp[32: 32 + X.size] = X.reshape(-1)
        Y = np.zeros((nsym, P), np.complex64)
        dm = np.zeros((nsym, P), np.uint8)
        si, src = np.zeros(nsym, np.int32), np.zeros(nsym, np.int32)
        n_out, first, sf = C.c_int(0), C.c_int(0), C.c_int(0)
        fi_start = 2 if (con == 2 and tm == 1) else 3
        os.environ["DVBT_EMUL_DEMOD_FUSED"] = str(request.param)
        rc = lib.emul_demod(C.c_void_p(Xp.ctypes.data + 32 * 8), nsym, con, tm, fi_start, 1, C.c_void_p(Y.ctypes.data), C.c_void_p(dm.ctypes.data),
                            C.c_void_p(si.ctypes.data), C.c_void_p(src.ctypes.data), C.byref(n_out), C.byref(first), C.byref(sf))
        assert rc == 0
        n = n_out.value
        return Y[src[:n]], dm[src[:n]], si[:n].copy(), sf.value
    return run


def check(demod, X, con, tm):
    Yo, sio, tag = O.demod(X, con, tm)
    Y, dm, si, sf = demod(X, con, tm)
    assert Y.shape == Yo.shape and Yo.shape[0] > 0
    assert np.array_equal(si, sio)
    assert np.array_equal(Y.view(np.uint32), Yo.view(np.uint32)), "equalised cells differ from the oracle"
    assert np.array_equal(dm, O.demap(Yo, con).reshape(Yo.shape[0], -1))
    assert sf == tag


def test_reference_fixture(demod):
    check(demod, CH["X"], 1, 0)
    Y, dm, si, sf = demod(CH["X"], 1, 0)
    assert np.array_equal(Y[:3].view(np.uint32), CH["cells_head"].view(np.uint32)) and np.array_equal(dm, CH["demap"])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref (reference TX blocks) not built")
@pytest.mark.parametrize("con,cr,tm,nsym,noise,shift", [(2, 4, 0, 330, 0.0, 0), (0, 1, 0, 480, 0.1, 3), (1, 2, 0, 420, 0.05, -5), (2, 4, 1, 290, 0.0, 0)])
def test_reference_tx_symbols(demod, con, cr, tm, nsym, noise, shift):
    """reference-TX symbols through a flat channel with noise and an integer carrier offset, both transmission modes"""
    from dvbt_testlib import tx_frequency_domain, channel
    tx = tx_frequency_domain(con, cr, tm, nsym, 9)
    check(demod, channel(tx["X"], noise=noise, bin_shift=shift, seed=2), con, tm)
