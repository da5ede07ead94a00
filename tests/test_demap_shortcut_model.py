"""CPU model check of the demapper's clear-case shortcut (gr_dvbt_b200/csrc/demod.cuh, demap_cell_near).

The reference (dvbt_demap_impl.cc:167-203) scans all 2^m constellation points for the first strictly smallest
fl(fl(dr^2) + fl(di^2)).  The kernel guesses the nearest level of each axis by division and, when the cell lies within
0.49 level spacings of that level on both axes (on the open side of an edge level: up to 16 spacings out), returns
the guessed point without looking at any other.  This file restates that decision in float32 numpy and requires that
every cell it accepts is demapped by the oracle restatement of the reference (oracle/port/demap_port.c) to exactly
the guessed point - on dense sweeps across the acceptance limit, the 16-spacing cut-off, huge and tiny magnitudes and
non-finite values.  It pins the algorithm; the GPU suite pins the CUDA build."""
import numpy as np
import pytest

from oracle import port as O

F = np.float32


def shortcut(cells, con, gain=1.0):
    """returns (accepted mask, guessed constellation index) the way demap_cell_near's clear-case test decides"""
    pts = O.constellation_points(con, 1, gain)
    m = {4: 2, 16: 4, 64: 6}[len(pts)]
    top = F((1 << (m // 2)) - 1)
    lv = np.unique(pts.real.astype(np.float32))
    g = F(lv[lv > 0][0])                               # level n = 1: g * 1
    inv_step = F(1.0) / (F(2.0) * g)
    lim, far = F(0.98) * g, F(32.0) * g
    with np.errstate(all="ignore"):
        def axis(v):
            n = np.minimum(np.maximum(F(2.0) * np.floor(v * inv_step) + F(1.0), -top), top).astype(np.float32)
            n = np.where(np.isnan(v), -top, n)         # fmaxf(NaN, -TOP) = -TOP, then fminf(-TOP, TOP)
            d = (v - g * n).astype(np.float32)
            ok = ((d > -lim) | (n == -top)) & ((d < lim) | (n == top)) & (np.abs(d) < far)
            return ok, (g * n).astype(np.float32)
        okx, lx = axis(cells.real.astype(np.float32))
        oky, ly = axis(cells.imag.astype(np.float32))
    # index of the point with exactly these two levels
    key = {(F(p.real).tobytes(), F(p.imag).tobytes()): i for i, p in enumerate(pts)}
    acc = okx & oky
    guess = np.full(len(cells), -1, np.int64)
    for i in np.nonzero(acc)[0]:
        guess[i] = key[(lx[i].tobytes(), ly[i].tobytes())]
    return acc, guess


def sweep(con, rng, n):
    pts = O.constellation_points(con)
    lv = np.unique(pts.real)
    step = float(lv[1] - lv[0]) if len(lv) > 1 else 2.0 * float(abs(lv[0]))
    parts = []
    # uniformly over and around the constellation
    parts.append(rng.uniform(lv[0] - 2 * step, lv[-1] + 2 * step, n) + 1j * rng.uniform(lv[0] - 2 * step, lv[-1] + 2 * step, n))
    # dense around the acceptance limit on one axis, anything on the other
    fr = rng.uniform(0.485, 0.495, n) * rng.choice([-1.0, 1.0], n)
    a = rng.choice(lv, n) + fr * step
    b = rng.choice(lv, n) + rng.uniform(-0.7, 0.7, n) * step
    parts += [a + 1j * b, b + 1j * a]
    # the decision boundary itself and one float either side
    mid = (rng.choice(lv[:-1], n) + step / 2).astype(np.float32) if len(lv) > 1 else np.zeros(n, np.float32)
    parts.append(np.nextafter(mid, np.float32(rng.choice([-9, 9]))) + 1j * b)
    # far outside: around the 16-spacing cut-off, and absurd magnitudes
    out = lv[-1] + rng.uniform(15.5, 16.5, n) * step
    parts += [out + 1j * b, b - 1j * out, rng.choice([1e-30, 1e6, 3e38, -3e38], n) + 1j * b]
    c = np.concatenate(parts).astype(np.complex64)
    c[:6] = np.array([complex(np.nan, 0.1), complex(0.1, np.nan), complex(np.inf, 0.1), complex(0.1, -np.inf), complex(np.inf, np.inf), 0], np.complex64)
    return c


@pytest.mark.parametrize("con", [0, 1, 2])
def test_clear_case_shortcut_agrees_with_the_reference_scan(con):
    rng = np.random.default_rng(40 + con)
    c = sweep(con, rng, 6000)
    acc, guess = shortcut(c, con)
    ref = O.demap(c, con)
    assert acc.sum() > len(c) // 4                      # the shortcut is the common case ...
    assert (~acc).sum() > len(c) // 20                  # ... and the sweep does reach the other paths
    bad = np.nonzero(acc & (guess != ref))[0]
    assert len(bad) == 0, (c[bad[:5]], guess[bad[:5]], ref[bad[:5]])


def test_shortcut_with_gain():
    rng = np.random.default_rng(7)
    with np.errstate(all="ignore"):
        c = (sweep(2, rng, 3000) * np.float32(0.37)).astype(np.complex64)
    acc, guess = shortcut(c, 2, 0.37)
    ref = O.demap(c, 2, 1, 0.37)
    assert acc.sum() > 1000 and np.array_equal(guess[acc], ref[acc])
