"""CPU checks of the front end's host-side pieces and numerical models (no GPU needed).

* dvbt_b200_resampler_taps: the 32/35 low-pass prototype must be what GNU Radio 3.7's
  filter.rational_resampler_ccc(64, 70) designs (python/rational_resampler.py design_filter ->
  firdes.low_pass(interp, interp, mid, tw, WIN_KAISER, 7.0), gr-filter/lib/firdes.cc): restated here in float64.
  Stock GNU Radio code is not under the reference tree (parity unpinned), so the formula is the pin.
* the derotation rotors of the fused FFT kernel (acq_fftd_kernel): two geometric sequences per thread instead of a
  sincos per sample.  Float32 model of the recurrence against float64 phases: the error must stay well inside the
  2e-5 the acquisition tests allow against the reference's own float accumulation."""
import numpy as np


def gnuradio_taps():
    interp, decim, beta = 32, 35, 7.0
    rate = interp / decim
    tw = rate * (0.5 - 0.4)
    mid = rate * 0.5 - tw / 2.0
    fs, gain = float(interp), float(interp)
    a = beta / 0.1102 + 8.7                                  # firdes::max_attenuation(WIN_KAISER, beta)
    ntaps = int(a * fs / (22.0 * tw))
    if ntaps % 2 == 0:
        ntaps += 1
    n = np.arange(ntaps)
    w = np.i0(beta * np.sqrt(1.0 - (2.0 * n / (ntaps - 1) - 1.0) ** 2)) / np.i0(beta)
    w = w.astype(np.float32).astype(np.float64)              # the window is stored as float
    M = (ntaps - 1) // 2
    k = n - M
    fw = 2 * np.pi * mid / fs
    with np.errstate(all="ignore"):
        t = np.where(k == 0, fw / np.pi * w, np.sin(k * fw) / (k * np.pi) * w)
    t = t.astype(np.float32).astype(np.float64)
    fmax = t[M] + 2 * t[M + 1:].sum()
    return (t * (gain / fmax)).astype(np.float32), ntaps


def test_resampler_taps_follow_gnuradio_design():
    import gr_dvbt_b200 as g
    lib = g.capi.lib()
    n = lib.dvbt_b200_resampler_taps(None, 0)
    got = np.zeros(n, np.float32)
    assert lib.dvbt_b200_resampler_taps(got.ctypes.data, n) == n
    ref, ntaps = gnuradio_taps()
    assert ntaps == 1149 and n == 1152 and n % 32 == 0       # install_taps pads to a multiple of the arm count: 36 per arm
    assert np.all(got[ntaps:] == 0)
    assert np.max(np.abs(got[:ntaps] - ref)) <= 2e-7 * np.max(np.abs(ref))
    assert np.array_equal(got[:ntaps], got[:ntaps][::-1])    # linear phase
    assert abs(float(got.astype(np.float64).sum()) - 32.0) < 1e-3   # unity gain per polyphase arm set


def test_derotation_rotor_recurrence_accuracy():
    F = np.float32
    rng = np.random.default_rng(3)
    worst = 0.0
    for N, T in ((2048, 128), (8192, 512)):
        for _ in range(40):
            phase0 = rng.uniform(-np.pi, np.pi)
            inc0, inc1 = rng.uniform(-3e-3, 3e-3, 2)          # up to half a carrier spacing of offset at 2k
            sw = int(rng.choice([0, 1, rng.integers(1, N), N + 300, 1 << 30]))
            for t in (0, 1, T // 2, T - 1):
                def rotor(ph):
                    ph = ph - 2 * np.pi * np.rint(ph / (2 * np.pi))
                    return complex(F(np.cos(F(ph))), F(np.sin(F(ph))))
                ra = rotor(phase0 + (t + 1) * inc0)
                rb = rotor(phase0 + sw * inc0 + (t + 1 - sw) * inc1)
                sa, sb = rotor(T * inc0), rotor(T * inc1)

                def cmul(a, b):                              # float32 complex multiply, each product and sum rounded
                    return complex(F(F(a.real) * F(b.real)) - F(F(a.imag) * F(b.imag)), F(F(a.real) * F(b.imag)) + F(F(a.imag) * F(b.real)))
                for r in range(16):
                    j = t + r * T
                    steps = j + 1
                    ph = phase0 + (steps * inc0 if steps <= sw else sw * inc0 + (steps - sw) * inc1)
                    exact = np.exp(1j * ph)
                    used = ra if steps <= sw else rb
                    worst = max(worst, abs(used - exact))
                    ra, rb = cmul(ra, sa), cmul(rb, sb)
                    ra, rb = complex(F(ra.real), F(ra.imag)), complex(F(rb.real), F(rb.imag))
    assert worst < 3e-6, worst
