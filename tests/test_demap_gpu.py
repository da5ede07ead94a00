"""GPU parity: CUDA dvbt_demap vs the oracle restatement of dvbt_demap_impl.cc (bit-exact:
the output is an integer decision; every float op is individually rounded on both sides)."""
import numpy as np
import pytest

from oracle import port as O

pytestmark = pytest.mark.gpu


def cells_for(con, n, seed, sigma=0.12):
    rng = np.random.default_rng(seed)
    pts = O.constellation_points(con)
    idx = rng.integers(0, len(pts), n)
    c = (pts[idx] + (rng.normal(0, sigma, n) + 1j * rng.normal(0, sigma, n))).astype(np.complex64)
    c[: len(pts)] = pts
    mids = ((pts[:, None] + pts[None, :]) / 2).reshape(-1).astype(np.complex64)  # exact ties
    k = min(len(mids), n - 100)
    c[100: 100 + k] = mids[:k]
    c[-4:] = np.array([0, 1e6 + 1e6j, -1e-30, np.complex64(complex(3.0, -7.5))], np.complex64)
    # one ulp either side of the ties, and non-finite cells (every comparison false in the reference's scan)
    near = mids[:k].copy()
    re = near.real.copy()
    im = near.imag.copy()
    re[0::2] = np.nextafter(re[0::2], np.float32(10))
    im[1::3] = np.nextafter(im[1::3], np.float32(-10))
    j = 100 + k
    m = min(k, n - j - 20)
    if m > 0:
        c[j: j + m] = (re + 1j * im).astype(np.complex64)[:m]
    # around the 0.49-spacing limit of the kernel's clear-case shortcut, on either axis and beyond the edge levels
    # (placed in the random region between the tie blocks and the non-finite cells)
    lv = np.unique(pts.real)
    step = float(lv[1] - lv[0]) if len(lv) > 1 else 2.0 * float(abs(lv[0]))
    start = j + max(m, 0)
    q = min(600, max(0, (n - 12 - start) // 2 - 2))
    if q > 4:
        fr = rng.uniform(0.47, 0.53, q) * rng.choice([-1.0, 1.0], q)
        a = rng.choice(lv, q) + fr * step
        b = rng.choice(lv, q) + rng.uniform(-0.6, 0.6, q) * step
        c[start: start + q] = (a + 1j * b).astype(np.complex64)
        c[start + q: start + 2 * q] = (b + 1j * a).astype(np.complex64)
        c[start: start + 4] = np.array([complex(lv[-1] + 15.9 * step, 0.1), complex(lv[-1] + 16.1 * step, 0.1), complex(0.1, lv[0] - 15.9 * step),
                                        complex(0.1, lv[0] - 16.1 * step)], np.complex64)
    c[-12:-4] = np.array([complex(np.nan, 0.3), complex(0.2, np.nan), complex(np.inf, 1), complex(-1, -np.inf), complex(np.inf, np.inf),
                          complex(np.nan, np.nan), complex(3.4e38, 3.4e38), complex(-3.4e38, 1e-38)], np.complex64)
    return c


@pytest.mark.parametrize("con", [0, 1, 2])
@pytest.mark.parametrize("P", [1512, 6048])
def test_demap_matches_oracle(con, P):
    import gr_dvbt_b200 as g
    nsym = 9
    c = cells_for(con, nsym * P, 10 * con + (P > 2000))
    d = g.dvbt_demap(P, con, g.NH, g.T2k if P == 1512 else g.T8k, 1.0)
    assert np.array_equal(d.points(), O.constellation_points(con))
    out, cons = d.general_work(nsym, c)
    assert cons == nsym
    assert np.array_equal(out, O.demap(c, con))


def test_demap_ragged_and_gain():
    import torch
    import gr_dvbt_b200 as g
    con = 2
    for n in (1, 2, 3, 5, 4099):
        c = cells_for(con, max(n, 5000), n)[:n].copy() * np.float32(0.5)
        d = g.dvbt_demap(1512, con, g.NH, g.T2k, 0.5)
        d_in = torch.from_numpy(c).cuda()
        d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
        d.run_dev(d_in.data_ptr(), n, d_out.data_ptr())
        assert np.array_equal(d_out.cpu().numpy(), O.demap(c, con, 1, 0.5))
