"""ofdm_sym_acquisition on the CPU from gr_dvbt_b200/csrc/acq.cu AS A WHOLE - the kernels (lambda tables, detector passes,
speculative state maps, compose / walk / finish, derotation) and the host orchestration of dvbt_b200_acq_work - compiled
for the host on a stand-in CUDA runtime (tests/emul/fake_cuda/) and driven through the same C ABI, against the oracle
restatement of ofdm_sym_acquisition_impl.cc.  Same assertions as tests/test_acq_gpu.py: timing decisions identical,
magnitudes to rounding, phase to the reference's own accumulated float rounding (closed-form derotation, DESIGN K2)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from gr_dvbt_b200 import capi
from oracle import port as O, refchain as R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import build_vit_emul  # noqa: E402

# a kernel that is not warp-converged would dead-lock the lock-step emulation: never hang the suite (the host threads sit
# inside a C call, so only the thread method of pytest-timeout can end the run)
pytestmark = pytest.mark.timeout(900, method="thread")


@pytest.fixture(scope="module")
def acq_work():
    lib = C.CDLL(build_vit_emul.build_acq())

    def run(x, N, K, cp):
        x = np.ascontiguousarray(x, np.complex64).reshape(-1)
        h = C.c_void_p()
        par = capi.AcqParams(1, N, K, cp, 30.0)
        assert lib.dvbt_b200_acq_create(C.byref(par), C.byref(h)) == 0
        cap = len(x) // (N + cp) + 1
        out = np.zeros((cap, N), np.complex64)
        tout = (capi.Tag * 4)()
        ntout, cons, prod = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
        rc = lib.dvbt_b200_acq_work(h, C.c_void_p(x.ctypes.data), C.c_size_t(len(x)), C.c_void_p(out.ctypes.data), C.c_size_t(cap), C.byref(cons),
                                    C.byref(prod), tout, C.c_size_t(4), C.byref(ntout), 0)
        assert rc == 0
        tags = [(int(tout[i].offset), capi.TAG_NAMES[tout[i].key], int(tout[i].value)) for i in range(ntout.value)]
        lib.dvbt_b200_acq_destroy(h)
        return out[: prod.value].copy(), int(cons.value), tags
    return run


@pytest.mark.skipif(not R.available(), reason="oracle/_ref (reference TX blocks) not built")
@pytest.mark.parametrize("tm,offset,cfo,noise", [(0, 777, 0.0, 0.0), (0, 1301, 0.11, 0.0), (0, 40, -0.2, 0.05), (1, 4000, -0.07, 0.0)])
def test_acquisition_matches_oracle(acq_work, tm, offset, cfo, noise):
    from dvbt_testlib import tx_frequency_domain, ofdm_modulate
    N, P, K, cp = R.mode_dims(tm)
    nsym = 60 if tm == 0 else 20
    tx = tx_frequency_domain(R.QAM16, R.C1_2, tm, nsym, 2)
    x = ofdm_modulate(tx["X"][:nsym], tm, offset=offset, cfo_bins=cfo, noise=noise, seed=1)
    ref, cons_ref, tag_ref = O.acquisition(x, N, cp)
    out, cons, tags = acq_work(x, N, K, cp)
    n = min(len(out), len(ref))
    assert n >= nsym - 4 and abs(len(out) - len(ref)) <= 1      # the batch may see one more complete symbol
    assert cons >= cons_ref and (cons - cons_ref) % (N + cp) == 0
    assert tags and tags[0] == (0, "sync_start", 1) and tag_ref
    mag = np.abs(np.abs(out[:n]) - np.abs(ref[:n])).max() / np.abs(ref[:n]).max()
    assert mag < 2e-6, mag
    err = np.abs(out[:n] - ref[:n]).max() / np.abs(ref[:n]).max()
    assert err < (2e-5 if cfo == 0.0 else 2e-3), err
