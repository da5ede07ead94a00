"""The one known answer the reference ships (BASELINE.json configs[0]): apps/test.ts through apps/dvbt_tx_demo.grc and
back through apps/dvbt_rx_demo.grc (2k / QAM16 / rate 1/2) is apps/test.ts again, from TS packet 504 on.

tests/golden/apps_test_ts_head.npz holds the first 2016 packets of that file (made by tests/golden/make_golden.py in
the build container, with the sha256 of the whole file).  Here, on the CPU:
  * the fixture is the head of the reference's file (where /root/reference exists);
  * the oracle restatements (oracle/port: demod, demap, deinterleavers, Viterbi, RS, descrambler) - the checker every
    GPU parity test relies on - reproduce the identity on it, fed by the reference's own TX blocks (oracle/_ref)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import port as O, refchain as R

FX = np.load(os.path.join(os.path.dirname(__file__), "golden", "apps_test_ts_head.npz"))
REF_TS = "/root/reference/apps/test.ts"


@pytest.mark.skipif(not os.path.exists(REF_TS), reason="reference tree not present")
def test_fixture_is_the_head_of_the_reference_file():
    raw = np.fromfile(REF_TS, np.uint8)
    assert len(raw) == int(FX["whole_file_bytes"]) == 70000 * 188
    assert hashlib.sha256(raw.tobytes()).digest() == bytes(FX["sha256_whole_file"])
    assert np.array_equal(raw[: len(FX["ts_head"])], FX["ts_head"])


@pytest.mark.skipif(not R.available(), reason="oracle/_ref (reference TX blocks) not built")
def test_oracle_port_chain_reproduces_test_ts_from_packet_504():
    from dvbt_testlib import channel
    head = FX["ts_head"]
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    ed, rs, ci = R.tx_outer(head)
    X = channel(R.tx_inner(ci, con, cr, tm, nsym=None)["X"])
    assert X.shape[0] == int(FX["nsym"])
    Y, si, tag = O.demod(X, con, tm)
    dm = O.demap(Y, con).reshape(Y.shape[0], -1)
    bd = O.bit_deinterleave(O.symbol_deinterleave(dm, tm, si), R.BITS_PER_CELL[con])
    vo = O.Viterbi(R.BITS_PER_CELL[con], cr).work(bd)
    cd = O.conv_deinterleave(vo)
    rsd, st = O.rs_decode(cd[: len(cd) // 204 * 204].reshape(-1, 204))
    ts, first = O.descramble(rsd)
    k0 = int(FX["first_packet"])
    assert len(ts) >= int(FX["reference_rx_bytes"]) - 1504 * 2 and len(ts) >= 1504 * 20
    assert np.array_equal(ts, head[k0 * 188: k0 * 188 + len(ts)])
