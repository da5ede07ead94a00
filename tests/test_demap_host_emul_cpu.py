"""demap_kernel of gr_dvbt_b200/csrc/demap.cu with the cell decision of demod.cuh (exact separable search + clear-case
shortcut) and make_demap_table, compiled for the host (tests/emul/, every float operation rounded on its own) against the
reference's golden outputs and the oracle restatement of dvbt_demap_impl.cc - hierarchical constellations included."""
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

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "hotpath_golden.npz"))
HIER = {1: 0, 2: 2, 4: 3}   # alpha -> dvbt_hierarchy_t (NH and ALPHA1 both mean alpha = 1)


@pytest.fixture(scope="module")
def demap():
    lib = C.CDLL(build_vit_emul.build_demap())
    lib.emul_demap.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]

    def run(cells, con, alpha=1, gain=1.0):
        c = np.ascontiguousarray(cells, np.complex64).reshape(-1)
        out = np.zeros(len(c), np.uint8)
        pts = np.zeros(128, np.float32)
        near_ok = lib.emul_demap(c.ctypes.data, len(c), con, HIER[alpha], gain, out.ctypes.data, pts.ctypes.data)
        assert near_ok >= 0
        return out, pts[: 2 << (2 * (con + 1))].view(np.complex64).copy(), near_ok
    return run


@pytest.mark.parametrize("con", [0, 1, 2])
def test_golden_cells(demap, con):
    """noisy cells and every midpoint between two constellation points (exact ties): the reference's own decisions"""
    out, pts, near_ok = demap(G["demap_in_c%d" % con], con)
    assert near_ok == 1
    assert np.array_equal(pts.view(np.uint32), O.constellation_points(con).astype(np.complex64).view(np.uint32))
    assert np.array_equal(out, G["demap_out_c%d" % con])


@pytest.mark.parametrize("con,alpha,gain", [(0, 1, 1.0), (1, 1, 0.37), (2, 1, 1.0), (2, 1, 3.5), (1, 2, 1.0), (1, 4, 1.0), (2, 2, 1.0), (2, 4, 0.8)])
def test_dense_sweep_matches_oracle(demap, con, alpha, gain):
    """a dense grid over the constellation and beyond it, decision boundaries +- a few ulps, non-finite cells; ragged length"""
    pts = O.constellation_points(con, alpha, gain).astype(np.complex64)
    top = float(np.abs(pts.real).max()) * 1.6
    ax = np.linspace(-top, top, 301, dtype=np.float32)
    grid = (ax[:, None] + 1j * ax[None, :]).astype(np.complex64).reshape(-1)
    lv = np.unique(pts.real)
    mids = ((lv[:-1] + lv[1:]) / 2).astype(np.float32)
    near = np.concatenate([np.nextafter(mids, np.float32(9), dtype=np.float32), mids, np.nextafter(mids, np.float32(-9), dtype=np.float32)])
    edge = (near[:, None] + 1j * near[None, :]).astype(np.complex64).reshape(-1)
    odd = np.array([np.inf, -np.inf + 1j, np.nan + 0j, 1j * np.inf, 0, 1e30 - 1e30j], np.complex64)
    cells = np.concatenate([grid, edge, odd])[:-1]          # length not a multiple of 4: the scalar tail of the kernel
    out, kpts, near_ok = demap(cells, con, alpha, gain)
    assert np.array_equal(kpts.view(np.uint32), pts.view(np.uint32))
    assert near_ok == (1 if alpha == 1 else 0)
    assert np.array_equal(out, O.demap(cells, con, alpha, gain))
