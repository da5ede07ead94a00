"""rs_decode_kernel of gr_dvbt_b200/csrc/rs.cu (clean-packet division, warp-cooperative Berlekamp-Massey / Chien / Forney,
both load paths: packed packets and the Forney-deinterleaving gather from the Viterbi stream) compiled for the host
(tests/emul/, warp shuffles and ballots as lock-step exchanges) against the reference's golden outputs and the oracle."""
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

GDIR = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def rs():
    lib = C.CDLL(build_vit_emul.build_rs())

    def run(inp, npk, as_built, gather_bytes=-1):
        inp = np.ascontiguousarray(inp, np.uint8).reshape(-1)
        out = np.zeros(npk * 188, np.uint8)
        st = np.full(npk, -99, np.int32)
        rc = lib.emul_rs(C.c_void_p(inp.ctypes.data), C.c_void_p(out.ctypes.data), C.c_void_p(st.ctypes.data), C.c_longlong(npk), int(as_built),
                         C.c_longlong(gather_bytes))
        assert rc == 0
        return out.reshape(-1, 188), st
    return run


@pytest.mark.parametrize("as_built", [0, 1])
def test_golden_packets_both_reference_builds(rs, as_built):
    """0..11 byte errors per packet: the reference's own outputs, source-intended and as-built"""
    G = np.load(os.path.join(GDIR, "hotpath_golden.npz"))
    out, st = rs(G["rs_rx"], len(G["rs_rx"]), as_built)
    assert np.array_equal(out, G["rs_out_asbuilt" if as_built else "rs_out_fixed"])
    ref, ref_st = O.rs_decode(G["rs_rx"], as_built=bool(as_built))
    assert np.array_equal(out, ref.reshape(-1, 188))
    assert (st > 0).sum() >= 8 * 7      # corrected packets report the number of corrections


@pytest.mark.parametrize("as_built", [0, 1])
def test_random_corruption_more_than_one_tile(rs, as_built):
    rng = np.random.default_rng(17 + as_built)
    npk = 256 + 77                       # a full tile and a partial one
    cw = O.rs_encode(rng.integers(0, 256, (npk, 188), dtype=np.uint8)).reshape(npk, 204).copy()
    for p in range(npk):
        ne = int(rng.integers(0, 11)) if p % 3 else 0
        pos = rng.choice(204, ne, replace=False)
        cw[p, pos] ^= rng.integers(1, 256, ne, dtype=np.uint8)
    out, st = rs(cw, npk, as_built)
    ref, _ = O.rs_decode(cw, as_built=bool(as_built))
    assert np.array_equal(out, ref.reshape(-1, 188))


def test_gather_path_applies_the_outer_deinterleaver(rs):
    """input = the Viterbi output stream of the reference fixture: convolutional_deinterleaver as an index map in the load"""
    CH = np.load(os.path.join(GDIR, "chain_2k_qam16_r12.npz"))
    stream = np.ascontiguousarray(CH["viterbi"])
    npk = len(stream) // 204
    out, st = rs(stream, npk, 0, gather_bytes=len(stream))
    cd = O.conv_deinterleave(stream)
    ref, _ = O.rs_decode(cd[: len(cd) // 204 * 204].reshape(-1, 204))
    n = min(len(ref.reshape(-1, 188)), npk)
    assert n > 200 and np.array_equal(out[:n], ref.reshape(-1, 188)[:n])
    assert np.array_equal(out.reshape(-1)[: len(CH["rs"])], CH["rs"])
