"""Soft-decision mode of the Viterbi decoder and of the fused chain - BEYOND the reference (lib/d_metrics.c:57-74 is a
stub, TODO.txt:25), so there is no reference output to match.  What pins it:
  * soft values of +-1 are the reference's hard metric: decode_soft(+-1) == the hard decoder == oracle/_ref, bit for bit;
  * any other values: bit for bit against oracle/port's scalar restatement of the same rule (viterbi_port.c);
  * the point of the mode: under noise it makes fewer errors than the hard decoder on the same received signal."""
import numpy as np
import pytest

from oracle import port as O

pytestmark = pytest.mark.gpu


def bpsk_case(rate, nbytes, sigma, seed):
    """information bytes -> code bits (punctured, reference order) -> +-1 + Gaussian noise -> soft values in [-6, 6]"""
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, nbytes, dtype=np.uint8)
    enc = O.conv_encode(data, 2, rate)                           # m = 2: two code bits per byte, MSB first
    bits = ((enc[:, None] >> np.array([1, 0])) & 1).reshape(-1).astype(np.float64)
    r = (2.0 * bits - 1.0) + sigma * rng.standard_normal(len(bits))
    soft = np.clip(np.rint(r * 4.0), -6, 6).astype(np.int8)       # saturates at 1.5 x the nominal amplitude
    hard = (r > 0).astype(np.uint8)
    return data, soft, hard


@pytest.mark.parametrize("rate", [0, 1, 2, 3, 4])
def test_plus_minus_one_is_the_hard_decoder(rate):
    import gr_dvbt_b200 as g
    k, n = O.RATE_KN[rate]
    data, soft, hard = bpsk_case(rate, 96 * k * 40, 0.45, 10 + rate)
    hard_in = (hard.reshape(-1, 2) @ np.array([2, 1])).astype(np.uint8)
    ref = O.Viterbi(2, rate).work(hard_in)
    v = g.viterbi_decoder(g.QPSK, g.NH, rate)
    got_hard = v.decode(hard_in)[0]
    v.set_soft(True)
    got = v.decode_soft((2 * hard.astype(np.int8) - 1))
    assert np.array_equal(got_hard, ref[: len(got_hard)])
    assert np.array_equal(got, got_hard)
    with pytest.raises(Exception):
        v.decode(hard_in)                                        # the hard entry points refuse a soft handle
    v.set_soft(False)
    assert np.array_equal(v.decode(hard_in)[0], got_hard)


@pytest.mark.parametrize("rate,sigma", [(0, 0.7), (1, 0.55), (2, 0.5), (3, 0.42), (4, 0.38), (4, 0.0)])
def test_soft_values_match_the_scalar_restatement(rate, sigma):
    import gr_dvbt_b200 as g
    k, n = O.RATE_KN[rate]
    data, soft, hard = bpsk_case(rate, 96 * k * 60, sigma, 20 + rate)
    ref = O.viterbi_soft(soft, rate)
    v = g.viterbi_decoder(g.QPSK, g.NH, rate)
    v.set_soft(True)
    got = v.decode_soft(soft)
    assert len(got) == len(ref) and np.array_equal(got, ref)
    if sigma == 0.0:
        assert np.array_equal(got, data[: len(got)])


def test_soft_repair_path_is_exact():
    """a warm-up of one byte time cannot converge: the verify / repair kernels run on soft step codes"""
    import gr_dvbt_b200 as g
    rate = 4
    data, soft, hard = bpsk_case(rate, 96 * 7 * 30, 0.4, 77)
    ref = O.viterbi_soft(soft, rate)
    v = g.viterbi_decoder(g.QPSK, g.NH, rate)
    v.set_soft(True)
    v.set_tuning(chunk_bytes=96, warmup_bytes=1, threads_per_block=32)
    got = v.decode_soft(soft)
    assert np.array_equal(got, ref)
    assert v.last_stats()["repaired"] > 0
    v.set_tuning(chunk_bytes=120, warmup_bytes=40, threads_per_block=64, ring_depth=3)   # split survivor ring
    assert np.array_equal(v.decode_soft(soft), ref)


def test_extreme_values_do_not_overflow_the_metrics():
    """every value at +-6 (clamped from +-127) and adversarial: the 8-bit metric range holds (spread 72 + 96 per byte time)"""
    import gr_dvbt_b200 as g
    rng = np.random.default_rng(3)
    for rate in (0, 4):
        k, n = O.RATE_KN[rate]
        nb = 96 * k * 20
        soft = rng.choice(np.array([-127, -6, 6, 127], np.int8), size=nb * 8 * n // k)
        ref = O.viterbi_soft(soft, rate)
        v = g.viterbi_decoder(g.QPSK, g.NH, rate)
        v.set_soft(True)
        assert np.array_equal(v.decode_soft(soft), ref)


@pytest.mark.parametrize("rate,sigma", [(0, 0.75), (2, 0.5), (4, 0.38)])
def test_soft_decisions_make_fewer_errors(rate, sigma):
    import gr_dvbt_b200 as g
    k, n = O.RATE_KN[rate]
    data, soft, hard = bpsk_case(rate, 96 * k * 200, sigma, 50 + rate)
    v = g.viterbi_decoder(g.QPSK, g.NH, rate)
    hard_in = (hard.reshape(-1, 2) @ np.array([2, 1])).astype(np.uint8)
    got_hard = v.decode(hard_in)[0]
    v.set_soft(True)
    got_soft = v.decode_soft(soft)
    eh = int(np.unpackbits(got_hard ^ data[: len(got_hard)]).sum())
    es = int(np.unpackbits(got_soft ^ data[: len(got_soft)]).sum())
    assert eh > 50, eh                      # the hard decoder is in trouble at this noise level ...
    assert es * 4 < eh, (es, eh)            # ... and the soft one makes less than a quarter of its errors
