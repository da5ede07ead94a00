"""The Viterbi DEVICE code of gr_dvbt_b200/csrc/viterbi.cu - depuncture kernel, the one-lane ACS kernel of every
schedule (byte-SWAR, h16, h16b) with whole and split survivor ring, verify and repair kernels - compiled for the host
(tests/emul/) and run against the oracle on the CPU.  It is the kernels' own source text: what the GPU parity tests prove
on a B200, this proves (for everything that does not depend on the hardware itself) where no GPU is available - which is
how the opt-in h16b schedule was checked before its first GPU run."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from oracle import port as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import build_vit_emul  # noqa: E402

# a kernel that is not warp-converged would dead-lock the lock-step emulation: never hang the suite (the host threads sit
# inside a C call, so only the thread method of pytest-timeout can end the run)
pytestmark = pytest.mark.timeout(900, method="thread")

VARIANTS = {"swar": 0, "h16": 1, "h16b": 2}


@pytest.fixture(scope="module")
def emul():
    lib = C.CDLL(build_vit_emul.build())
    lib.emul_viterbi.restype = C.c_int
    lib.emul_viterbi.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_uint)]

    def run(rx, rate, m, variant, L, W, bd, depth=0):
        rx = np.ascontiguousarray(rx, np.uint8)
        out = np.zeros(len(rx) * m + 64, np.uint8)
        n_out = C.c_longlong(0)
        counters = (C.c_uint * 4)()
        rc = lib.emul_viterbi(rx.ctypes.data, len(rx), rate, m, VARIANTS[variant], L, W, bd, depth, out.ctypes.data, C.byref(n_out), counters)
        assert rc == 0
        return out[: n_out.value].copy(), dict(flagged=int(counters[0]), repaired=int(counters[1]))
    return run


def make_case(rate, m, nblocks, ber, seed):
    k, n = O.RATE_KN[rate]
    data = np.random.default_rng(seed).integers(0, 256, nblocks * 96 * k, dtype=np.uint8)
    rx = O.conv_encode(data, m, rate)
    if ber > 0:
        rx = O.flip_bits(rx, m, ber, seed + 1)
    return data, rx


@pytest.mark.parametrize("variant", ["swar", "h16", "h16b"])
@pytest.mark.parametrize("rate,m,ber", [(0, 4, 0.03), (1, 2, 0.02), (2, 6, 0.02), (3, 4, 0.01), (4, 6, 0.005), (4, 2, 0.0)])
def test_kernels_match_oracle_whole_ring(emul, variant, rate, m, ber):
    data, rx = make_case(rate, m, 6, ber, 40 + rate)
    ref = O.Viterbi(m, rate).work(rx)
    out, st = emul(rx, rate, m, variant, L=96, W=40, bd=16)
    assert np.array_equal(out, ref)
    if ber == 0.0:
        assert np.array_equal(out, data[: len(out)]) and st["repaired"] == 0


@pytest.mark.parametrize("variant", ["swar", "h16", "h16b"])
@pytest.mark.parametrize("rate,m,ber", [(0, 4, 0.04), (4, 6, 0.006), (2, 2, 0.03)])
def test_repair_path_is_exact(emul, variant, rate, m, ber):
    """a warm-up of one byte time cannot converge: verify flags the boundaries, repair re-decodes from the true state"""
    data, rx = make_case(rate, m, 5, ber, 7)
    ref = O.Viterbi(m, rate).work(rx)
    out, st = emul(rx, rate, m, variant, L=72, W=1, bd=8)
    assert st["flagged"] > 0 and st["repaired"] > 0
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("variant", ["swar", "h16", "h16b"])
@pytest.mark.parametrize("rate,m,ber,depth", [(4, 6, 0.008, 2), (4, 6, 0.0, 1), (3, 4, 0.012, 3), (4, 2, 0.004, 12)])
def test_split_survivor_ring_is_exact(emul, variant, rate, m, ber, depth):
    """ring depth < ntraceback: older rows come from the global ring when a traceback has not merged"""
    data, rx = make_case(rate, m, 6, ber, 31 + depth)
    ref = O.Viterbi(m, rate).work(rx)
    out, st = emul(rx, rate, m, variant, L=80, W=40, bd=16, depth=depth)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("rounds", [0, 1, 2])
def test_parallel_repair_rounds_equal_sequential_repair(emul, rounds, monkeypatch):
    """vit_repair_round_kernel re-decodes all bad chunks whose predecessor is settled at once; whatever the number of
    rounds, the sequential kernel finishes the job: same bytes as the oracle with 0, 1 or 2 parallel rounds - with every
    boundary bad (warm-up of one byte time: runs of consecutive bad chunks, worst case for the rounds) and with sparse
    bad chunks (noisy input, short warm-up)"""
    monkeypatch.setenv("DVBT_EMUL_REPAIR_ROUNDS", str(rounds))
    rate, m = 4, 6
    k, n = O.RATE_KN[rate]
    data = np.random.default_rng(3).integers(0, 256, 40 * 96 * k, dtype=np.uint8)
    for ber, W in ((0.0, 1), (0.01, 1), (0.012, 12), (0.02, 20)):
        rx = O.flip_bits(O.conv_encode(data, m, rate), m, ber, 5)
        want = O.Viterbi(m, rate).work(rx)
        out, st = emul(rx, rate, m, "h16", L=96, W=W, bd=32)
        assert np.array_equal(out, want), (rounds, ber, W)
        assert st["flagged"] > 0 and st["repaired"] > 0, (rounds, ber, W, st)
