"""GPU parity: CUDA reed_solomon_dec vs the oracle restatement of reed_solomon.cc (both the
source-intended decoder and the as-built gcc behaviour)."""
import numpy as np
import pytest

from oracle import port as O

pytestmark = pytest.mark.gpu


def corrupted(npk, seed, max_err=12):
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, (npk, 188), dtype=np.uint8)
    rx = O.rs_encode(data)
    nerr = rng.integers(0, max_err + 1, npk)
    for p in range(npk):
        pos = rng.choice(204, nerr[p], replace=False)
        rx[p, pos] ^= rng.integers(1, 256, nerr[p], dtype=np.uint8)
    return data, rx, nerr


@pytest.mark.parametrize("as_built", [0, 1])
def test_general_work_matches_oracle(as_built):
    import gr_dvbt_b200 as g
    data, rx, nerr = corrupted(8 * 125, 5)
    ref, st = O.rs_decode(rx, bool(as_built))
    dec = g.reed_solomon_dec(2, 8, 0x11D, 255, 239, 8, 51, 8)
    dec.set_compat(as_built)
    out, cons = dec.general_work(125, rx.reshape(-1))
    assert cons == 125
    assert np.array_equal(out.reshape(-1, 188), ref)
    if not as_built:
        good = nerr <= 8
        assert np.array_equal(out.reshape(-1, 188)[good], data[good])


def test_status_and_edge_patterns():
    import torch
    import gr_dvbt_b200 as g
    rng = np.random.default_rng(9)
    data = rng.integers(0, 256, (64, 188), dtype=np.uint8)
    rx = O.rs_encode(data)
    rx[0, 0] ^= 1                    # first byte
    rx[1, 203] ^= 0x80               # last parity byte
    rx[2, :8] ^= 0xFF                # burst of 8 at the start
    rx[3, 196:204] ^= 0x55           # burst of 8 in the parity
    rx[4, :9] ^= 0xFF                # 9 errors: uncorrectable
    rx[5] = 0                        # all-zero packet is a code word
    rx[6] = 0xFF                     # all-ones
    rx[7, ::25] ^= 3                 # spread
    rx[8:16] = rng.integers(0, 256, (8, 204), dtype=np.uint8)  # random garbage: miscorrection / failure paths
    ref, st = O.rs_decode(rx, False)
    dec = g.reed_solomon_dec()
    d_in = torch.from_numpy(rx).cuda()
    d_out = torch.zeros((64, 188), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(64, dtype=torch.int32, device="cuda")
    dec.decode_dev(d_in.data_ptr(), 64, d_out.data_ptr(), d_st.data_ptr())
    assert np.array_equal(d_out.cpu().numpy(), ref)
    assert np.array_equal(d_st.cpu().numpy(), st)


def test_large_batch_property():
    """full-size property: decode(encode(x) + <=8 errors) == x for 200k packets, status = error count"""
    import torch
    import gr_dvbt_b200 as g
    npk = 200_000
    rng = np.random.default_rng(1)
    data = rng.integers(0, 256, (npk, 188), dtype=np.uint8)
    rx = O.rs_encode(data)
    nerr = rng.integers(0, 9, npk)
    for e in range(1, 9):
        idx = np.where(nerr >= e)[0]
        # distinct positions: stride pattern per error index
        pos = (rng.integers(0, 25, len(idx)) + 25 * (e - 1)) % 204
        rx[idx, pos] ^= rng.integers(1, 256, len(idx), dtype=np.uint8)
    dec = g.reed_solomon_dec()
    d_in = torch.from_numpy(rx).cuda()
    d_out = torch.zeros((npk, 188), dtype=torch.uint8, device="cuda")
    d_st = torch.zeros(npk, dtype=torch.int32, device="cuda")
    dec.decode_dev(d_in.data_ptr(), npk, d_out.data_ptr(), d_st.data_ptr())
    assert np.array_equal(d_out.cpu().numpy(), data)
    assert np.array_equal(d_st.cpu().numpy(), nerr)
