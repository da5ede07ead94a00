"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from oracle import refchain as R


def random_ts(npackets, seed):
    """188-byte TS packets: sync 0x47 + seeded random payload."""
    rng = np.random.default_rng(seed)
    ts = rng.integers(0, 256, (npackets, 188), dtype=np.uint8)
    ts[:, 0] = 0x47
    return ts.reshape(-1)


def tx_frequency_domain(con, cr, tm, nsym_min, seed, gi=0):
    """Reference TX chain (verbatim reference blocks through oracle/_ref) up to the pilot insertion.
    Returns dict(ts, ci, X (nsym, N) complex64, plus the inner-chain intermediates)."""
    N, P, _, _ = R.mode_dims(tm)
    k, n = R.RATE_KN[cr]
    m = R.BITS_PER_CELL[con]
    per_item = P * k * m // (8 * n)
    npk = ((nsym_min + 8) * per_item // 204 // 8 + 3) * 8
    ts = random_ts(npk, seed)
    ed, rs, ci = R.tx_outer(ts)
    tx = R.tx_inner(ci, con, cr, tm, nsym=(nsym_min + 3) // 4 * 4, gi=gi)
    tx["ts"] = ts
    tx["ci"] = ci
    return tx


def channel(X, scale=0.01, noise=0.0, bin_shift=0, seed=0):
    """flat channel as in SURVEY B.4: scale, optional AWGN, optional integer carrier offset"""
    rng = np.random.default_rng(seed)
    Y = (X * np.float32(scale)).astype(np.complex64)
    if bin_shift:
        Y = np.roll(Y, bin_shift, axis=1)
    if noise > 0:
        Y = (Y + (rng.normal(0, noise * scale, Y.shape) + 1j * rng.normal(0, noise * scale, Y.shape))).astype(np.complex64)
    return np.ascontiguousarray(Y)


def ofdm_modulate(X, tm, gain=None, offset=0, cfo_bins=0.0, noise=0.0, seed=0, gi=0):
    """What the TX flowgraph does after reference_signals (apps/dvbt_tx_demo*.grc): fft_vxx(reverse,
    shift=True) = unnormalised inverse DFT of the half-swapped vector, cyclic prefix N/32,
    multiply_const.  Plus a test channel: `offset` leading samples, carrier offset in bins, AWGN."""
    N, P, K, cp = R.mode_dims(tm, gi)
    if gain is None:
        gain = 0.0022097087 if tm == R.T2k else 0.00055242272
    t = np.fft.ifft(np.fft.ifftshift(X.astype(np.complex128), axes=1), axis=1) * N
    t = np.concatenate([t[:, N - cp:], t], axis=1).reshape(-1) * gain
    rng = np.random.default_rng(seed)
    lead = (rng.normal(0, 1e-4, offset) + 1j * rng.normal(0, 1e-4, offset)) if offset else np.zeros(0)
    t = np.concatenate([lead, t])
    if cfo_bins:
        t = t * np.exp(2j * np.pi * cfo_bins * np.arange(len(t)) / N)
    if noise > 0:
        rms = np.sqrt(np.mean(np.abs(t) ** 2))
        t = t + (rng.normal(0, noise * rms, len(t)) + 1j * rng.normal(0, noise * rms, len(t)))
    return t.astype(np.complex64)


def to_capture_rate(x):
    """64/7 Msps -> 10 Msps, the job of rational_resampler(70, 64) in the TX flowgraphs
    (polyphase 35/32 with a Kaiser-windowed low-pass; scipy's design, not GNU Radio's taps)."""
    from scipy.signal import resample_poly
    y = resample_poly(x.astype(np.complex128), 35, 32, window=("kaiser", 7.0))
    return y.astype(np.complex64)
