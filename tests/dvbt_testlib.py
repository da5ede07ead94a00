"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from oracle import refchain as R


def random_ts(npackets, seed):
    """188-byte TS packets: sync 0x47 + seeded random payload."""
    rng = np.random.default_rng(seed)
    ts = rng.integers(0, 256, (npackets, 188), dtype=np.uint8)
    ts[:, 0] = 0x47
    return ts.reshape(-1)


def tx_frequency_domain(con, cr, tm, nsym_min, seed):
    """Reference TX chain (verbatim reference blocks through oracle/_ref) up to the pilot insertion.
    Returns dict(ts, ci, X (nsym, N) complex64, plus the inner-chain intermediates)."""
    N, P, _, _ = R.mode_dims(tm)
    k, n = R.RATE_KN[cr]
    m = R.BITS_PER_CELL[con]
    per_item = P * k * m // (8 * n)
    npk = ((nsym_min + 8) * per_item // 204 // 8 + 3) * 8
    ts = random_ts(npk, seed)
    ed, rs, ci = R.tx_outer(ts)
    tx = R.tx_inner(ci, con, cr, tm, nsym=(nsym_min + 3) // 4 * 4)
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
