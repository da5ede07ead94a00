"""Front end of apps/dvbt_rx_demo*.grc on the GPU: rational_resampler_ccc(64, 70) + multiply_const.

Stock GNU Radio blocks (no source under the reference tree, SURVEY §8c: parity unpinned), so the oracle here is the
documented polyphase formula evaluated in float64 with the taps the library reports:
    y[m] = gain * sum_j h[(35 m mod 32) + 32 j] * x[floor(35 m / 32) - j]      (zero history)
Tolerance: 5e-6 of the output RMS (float32 accumulation of 36 products).  All kernel variants add the taps in the
same order, so they must agree with each other bit for bit."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def taps():
    import gr_dvbt_b200 as g
    lib = g.capi.lib()
    n = lib.dvbt_b200_resampler_taps(None, 0)
    t = np.zeros(n, np.float32)
    assert lib.dvbt_b200_resampler_taps(t.ctypes.data, n) == n
    return t


def run(x, gain, variant):
    import gr_dvbt_b200 as g
    lib = g.capi.lib()
    x = np.ascontiguousarray(x, np.complex64)
    cap = len(x) * 32 // 35 + 8
    y = np.zeros(cap, np.complex64)
    n = C.c_size_t(0)
    g.capi.check(lib.dvbt_b200_resample_host(x.ctypes.data, len(x), gain, y.ctypes.data, cap, C.byref(n), variant))
    return y[: n.value].copy()


def polyphase_f64(x, h, gain):
    nout = ((len(x) - 1) * 32) // 35 + 1 if len(x) else 0
    per_arm = len(h) // 32
    xp = np.concatenate([np.zeros(per_arm, np.complex128), x.astype(np.complex128)])
    m = np.arange(nout)
    a = (35 * m) // 32
    ph = (35 * m) % 32
    y = np.zeros(nout, np.complex128)
    hh = h.astype(np.float64)
    for j in range(per_arm):
        y += hh[ph + 32 * j] * xp[a - j + per_arm]
    return y * gain


@pytest.mark.parametrize("n", [1, 35, 1119, 1120, 1121, 2240 * 3 + 17, 200_003])
def test_resampler_matches_the_polyphase_formula(n):
    rng = np.random.default_rng(n)
    x = (rng.normal(size=n) + 1j * rng.normal(size=n)).astype(np.complex64)
    h = taps()
    assert len(h) % 32 == 0 and len(h) // 32 == 36
    gain = 0.0022097087
    ref = polyphase_f64(x, h, np.float32(gain))
    outs = {v: run(x, gain, v) for v in (-1, 0, 1, 2, 4)}
    for v, y in outs.items():
        assert len(y) == len(ref), v
        rms = max(float(np.sqrt(np.mean(np.abs(ref) ** 2))), 1e-30)
        assert float(np.max(np.abs(y - ref))) <= 5e-6 * rms, v
    for v in (-1, 1, 2, 4):
        assert np.array_equal(outs[v].view(np.uint32), outs[0].view(np.uint32)), v


def test_resampler_empty_and_capacity():
    import gr_dvbt_b200 as g
    lib = g.capi.lib()
    y = np.zeros(4, np.complex64)
    n = C.c_size_t(7)
    g.capi.check(lib.dvbt_b200_resample_host(None, 0, 1.0, y.ctypes.data, 4, C.byref(n), -1))
    assert n.value == 0
    x = np.ones(100, np.complex64)
    assert lib.dvbt_b200_resample_host(x.ctypes.data, 100, 1.0, y.ctypes.data, 4, C.byref(n), -1) != 0
