"""The glue kernels of gr_dvbt_b200/csrc/rx_chain.cu compiled for the host (tests/emul/) against the oracle on the CPU:
rx_inner_codes_kernel (symbol deinterleaver + bit deinterleaver + vector_to_stream + unpack/depuncture as one index map
from demapped cells to Viterbi step codes) and rx_descramble_kernel (NSYNC search + energy descrambler).  The kernels'
own source text; the GPU parity tests prove the same on a B200."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import viterbi_model as VM
from oracle import port as O, refchain as R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import build_vit_emul  # noqa: E402

# a kernel that is not warp-converged would dead-lock the lock-step emulation: never hang the suite (the host threads sit
# inside a C call, so only the thread method of pytest-timeout can end the run)
pytestmark = pytest.mark.timeout(900, method="thread")

CH = np.load(os.path.join(os.path.dirname(__file__), "golden", "chain_2k_qam16_r12.npz"))


@pytest.fixture(scope="module")
def lib():
    return C.CDLL(build_vit_emul.build_rx())


def permutation_tables(tm):
    """H and H^-1 of the symbol interleaver, read off the oracle: an even symbol is out[q] = in[H(q)], an odd one
    out[H(q)] = in[q] (symbol_inner_interleaver_impl.cc:202-208)"""
    P = 1512 if tm == 0 else 6048
    idx = np.arange(P)
    tabs = []
    for parity in (0, 1):
        lo = O.symbol_deinterleave((idx & 0xFF).astype(np.uint8), tm, [parity]).reshape(-1).astype(np.int32)
        hi = O.symbol_deinterleave((idx >> 8).astype(np.uint8), tm, [parity]).reshape(-1).astype(np.int32)
        tabs.append((lo | (hi << 8)).astype(np.int16))
    H, Hinv = tabs
    assert np.array_equal(np.sort(H), idx) and np.array_equal(H[Hinv], idx)
    return np.ascontiguousarray(H), np.ascontiguousarray(Hinv)


def run_inner(lib, dm, symidx, tm, m, rate):
    P = 1512 if tm == 0 else 6048
    k, n = O.RATE_KN[rate]
    dm = np.ascontiguousarray(dm, np.uint8).reshape(-1, P)
    n_out = dm.shape[0]
    nblocks = n_out * P // (768 * n // m)
    nbt = nblocks * 96 * k
    H, Hinv = permutation_tables(tm)
    src = np.arange(n_out, dtype=np.int32)
    si = np.ascontiguousarray(symidx, np.int32)
    codes = np.zeros(nbt + 8, np.uint32)
    rc = lib.emul_inner_codes(C.c_void_p(dm.ctypes.data), C.c_void_p(src.ctypes.data), C.c_void_p(si.ctypes.data), C.c_void_p(H.ctypes.data),
                              C.c_void_p(Hinv.ctypes.data), P, m, n_out, rate, C.c_void_p(codes.ctypes.data), nbt)
    assert rc == 0
    bd = O.bit_deinterleave(O.symbol_deinterleave(dm, tm, si), m)
    want = VM.depuncture_codes(bd[: nblocks * (768 * n // m)], m, rate, nbt)
    return codes[:nbt], want


def test_inner_codes_kernel_on_the_reference_fixture(lib):
    got, want = run_inner(lib, CH["demap"], CH["symbol_index"], 0, 4, 0)
    assert len(want) > 40000 and np.array_equal(got, want)


@pytest.mark.parametrize("tm,m,rate,nsym,seed", [(0, 6, 4, 41, 1), (0, 2, 1, 30, 2), (1, 6, 2, 9, 3), (1, 4, 3, 7, 4), (0, 4, 4, 13, 5)])
def test_inner_codes_kernel_random_cells(lib, tm, m, rate, nsym, seed):
    """random demapped cells, symbol_index starting at an arbitrary value (both directions of the symbol interleaver,
    every constellation, tiles that end inside a byte time, a last partial tile)"""
    P = 1512 if tm == 0 else 6048
    rng = np.random.default_rng(seed)
    dm = rng.integers(0, 1 << m, (nsym, P), dtype=np.uint8)
    si = (np.arange(nsym) + seed) % 68
    got, want = run_inner(lib, dm, si, tm, m, rate)
    assert len(want) > 1000 and np.array_equal(got, want)


def prbs_table():
    """the 8-packet PRBS of energy_descramble, read off the oracle: descrambling zero payloads returns it"""
    z = np.zeros((16, 188), np.uint8)
    z[0, 0] = z[8, 0] = 0xB8
    ts, first = O.descramble(z)
    assert first == 0 and len(ts) >= 1504
    return np.ascontiguousarray(ts[:1504]).view(np.uint32).copy()


@pytest.mark.parametrize("lead", [0, 3, 8, 13])
def test_descramble_kernel_matches_oracle(lib, lead):
    """the fixture's RS output behind `lead` junk packets (NSYNC search), word path and byte path (odd output address)"""
    rs = np.concatenate([np.random.default_rng(lead).integers(0, 0xB0, (lead, 188), dtype=np.uint8), CH["rs"].reshape(-1, 188)])
    want, first = O.descramble(rs)
    prbs = prbs_table()
    for misalign in (0, 1):
        buf = np.zeros(rs.size + 8, np.uint8)
        ts = buf[misalign:]
        p0, ng = C.c_int(-2), C.c_longlong(0)
        rc = lib.emul_descramble(C.c_void_p(np.ascontiguousarray(rs).ctypes.data), C.c_longlong(rs.shape[0]), C.c_void_p(prbs.ctypes.data),
                                 C.c_void_p(ts.ctypes.data), C.c_longlong(rs.size), 3, C.byref(p0), C.byref(ng))
        assert rc == 0 and p0.value == first
        n = min(ng.value * 1504, len(want))
        assert n >= 1504 * 10 and np.array_equal(ts[:n], want[:n])


@pytest.mark.parametrize("pieces", [1, 3])
def test_descrambler_state_machine_matches_oracle_with_broken_nsync(lib, pieces):
    """rx_descr_plan_kernel replays energy_descramble's per-call NSYNC check (energy_descramble_impl.cc:121-141): destroyed
    NSYNC bytes, a jump of the 8-packet phase, the index carried across pieces - against the oracle restatement
    (dvbt_oracle_descramble_calls, itself pinned to the reference block in tests/test_oracle_cpu.py)"""
    rng = np.random.default_rng(9)
    src = CH["rs"].reshape(-1, 188)
    a = int(np.flatnonzero(src[:, 0] == 0xB8)[0])
    stream = np.concatenate([rng.integers(0, 0xB0, (5, 188), dtype=np.uint8), src[a: a + 64], src[a + 67: a + 67 + 93]])
    nsync = np.flatnonzero(stream[:, 0] == 0xB8)
    prbs = prbs_table()
    for kill in ([], [nsync[1]], [nsync[2], nsync[3]], list(nsync[4:7]), [nsync[-2]]):
        pk = stream.copy()
        pk[kill, 0] = 0
        cuts = [0, len(pk)] if pieces == 1 else [0, 37, 90, len(pk)]
        want_all, got_all = [], []
        pend_o = np.zeros((0, 188), np.uint8)
        pend_k = np.zeros((0, 188), np.uint8)
        pk_o, pk_k = 0, 0
        for ci in range(len(cuts) - 1):
            end = ci == len(cuts) - 2
            new = pk[cuts[ci]: cuts[ci + 1]]
            pend_o = np.concatenate([pend_o, new])
            want, used, pk_o, first = O.descramble_calls(pend_o, pk=pk_o, flush=end)
            want_all.append(want)
            pend_o = pend_o[8 * used:]
            pend_k = np.ascontiguousarray(np.concatenate([pend_k, new]))
            ts = np.zeros(pend_k.size + 3008, np.uint8)
            pkc, fp, ng, iu = C.c_int(pk_k), C.c_longlong(-1), C.c_longlong(0), C.c_longlong(0)
            rc = lib.emul_descramble_stream(C.c_void_p(pend_k.ctypes.data), C.c_longlong(pend_k.shape[0]), C.c_void_p(prbs.ctypes.data),
                                            C.c_void_p(ts.ctypes.data), C.c_longlong(ts.size), 2, int(end), C.byref(pkc), C.byref(fp), C.byref(ng), C.byref(iu))
            assert rc == 0
            pk_k = pkc.value
            got_all.append(ts[: ng.value * 1504].copy())
            assert iu.value == used and pk_k == pk_o and fp.value == first
            pend_k = pend_k[8 * iu.value:]
        assert np.array_equal(np.concatenate(got_all), np.concatenate(want_all))
        assert len(np.concatenate(want_all)) >= 1504 * 8
