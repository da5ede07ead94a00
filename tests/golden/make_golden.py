#!/usr/bin/env python3
"""Generates the committed golden fixtures from the reference itself (oracle/_ref: the
reference's sources compiled verbatim).  Run in the build container, where /root/reference
exists:  python tests/golden/make_golden.py
The fixtures pin the C restatements in oracle/port (tests/test_oracle_cpu.py) and are also
used by the GPU tests, so parity does not depend on oracle/_ref being present at test time."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refchain as R, port as O  # noqa: E402


def viterbi_vectors():
    out = {}
    for rate in range(5):
        for m in (2, 4, 6):
            k, n = O.RATE_KN[rate]
            data = np.random.default_rng(1000 + 10 * rate + m).integers(0, 256, 3 * 96 * k, dtype=np.uint8)
            # the encoder of the port is itself pinned here against the reference's inner_coder block
            rx = O.conv_encode(data, m, rate)
            for ber in (0.0, 0.02):
                rxn = O.flip_bits(rx, m, ber, 7) if ber else rx
                ref, _ = R.rx_viterbi(rxn, {2: 0, 4: 1, 6: 2}[m], rate, None, blocks_per_call=2)
                out["vit_in_r%d_m%d_b%d" % (rate, m, int(ber * 100))] = rxn
                out["vit_out_r%d_m%d_b%d" % (rate, m, int(ber * 100))] = ref
    return out


def rs_vectors():
    rng = np.random.default_rng(42)
    npk = 8 * 12
    data = rng.integers(0, 256, (npk, 188), dtype=np.uint8)
    cw = np.zeros(npk * 204, np.uint8)
    R.RefBlock("reed_solomon_enc", 2, 8, 0x11D, 255, 239, 8, 51, 8).work(npk // 8, npk // 8, data.reshape(-1).copy(), cw)
    rx = cw.reshape(npk, 204).copy()
    nerr = np.arange(npk) % 12
    for p in range(npk):
        pos = rng.choice(204, nerr[p], replace=False)
        rx[p, pos] ^= rng.integers(1, 256, nerr[p], dtype=np.uint8)
    res = {"rs_data": data, "rs_codewords": cw.reshape(npk, 204), "rs_rx": rx}
    for fixed in (False, True):
        out = np.zeros(npk * 188, np.uint8)
        R.RefBlock("reed_solomon_dec", 2, 8, 0x11D, 255, 239, 8, 51, 8, fixed_rs=fixed).work(npk // 8, npk // 8, rx.reshape(-1).copy(), out)
        res["rs_out_fixed" if fixed else "rs_out_asbuilt"] = out.reshape(npk, 188)
    return res


def demap_vectors():
    res = {}
    rng = np.random.default_rng(3)
    for con in (0, 1, 2):
        pts = O.constellation_points(con)
        n = 1512 * 2
        idx = rng.integers(0, len(pts), n)
        c = (pts[idx] + (rng.normal(0, 0.15, n) + 1j * rng.normal(0, 0.15, n))).astype(np.complex64)
        mids = ((pts[:, None] + pts[None, :]) / 2).reshape(-1).astype(np.complex64)
        c[: min(len(mids), n)] = mids[:n]
        res["demap_in_c%d" % con] = c
        res["demap_out_c%d" % con] = R.rx_demap(c.reshape(2, 1512), con, R.T2k).reshape(-1)
    return res


def inner_coder_vectors():
    """reference inner_coder output for a known input: pins oracle.port.conv_encode"""
    res = {}
    for rate, con in ((0, 1), (4, 2), (2, 0)):
        k, n = O.RATE_KN[rate]
        m = R.BITS_PER_CELL[con]
        nbytes = 4 * 1512 * k * m // (8 * n)  # inner_coder consumes this for 4 items (inner_coder_impl.cc:254)
        data = np.random.default_rng(77 + rate).integers(0, 256, nbytes, dtype=np.uint8)
        ic = np.zeros(4 * 1512, np.uint8)
        R.RefBlock("inner_coder", 1, 1512, con, R.NH, rate).work(4, 0, data, ic)
        res["ic_in_r%d_c%d" % (rate, con)] = data
        res["ic_out_r%d_c%d" % (rate, con)] = ic
    return res


def chain_fixture():
    """frequency-domain loopback, 2k/QAM16/rate 1/2: input symbols + every reference stage output"""
    from dvbt_testlib import tx_frequency_domain, channel
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    tx = tx_frequency_domain(con, cr, tm, 400, 5)
    X = channel(tx["X"][:400])
    Y, tags = R.rx_demod(X, con, cr, tm)
    dm = R.rx_demap(Y, con, tm)
    sd, bd = R.rx_deinterleave(dm, tags, con, tm)
    sf = [t for t in tags if t[1] == "superframe_start"][0][0]
    vo, vtags = R.rx_viterbi(bd, con, cr, sf * 1512)
    cd, rd, ts = R.rx_outer(vo, vtags, fixed_rs=True)
    assert np.array_equal(ts, tx["ts"][504 * 188: 504 * 188 + len(ts)]) and len(ts) >= 1504
    d = dict(X=X, cells_head=Y[:3], symbol_index=np.array([t[2] for t in tags if t[1] == "symbol_index"], np.int32),
             n_out=np.int64(Y.shape[0]), demap=dm, sym_deint_head=sd[:4], bit_deint=bd, viterbi=vo, conv_deint_head=cd[: 204 * 40], rs=rd, ts=ts,
             ts_source=tx["ts"][504 * 188: 504 * 188 + len(ts) + 1504 * 4])
    path = os.path.join(HERE, "chain_2k_qam16_r12.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path), "bytes")


def apps_test_ts_fixture():
    """The one known-answer the reference ships: apps/test.ts is what apps/dvbt_tx_demo.grc transmits and what
    apps/dvbt_rx_demo.grc must give back (BASELINE.json configs[0], 2k/QAM16/rate 1/2; from TS packet 504 on: SURVEY §8c).
    The head of the file is committed (data, not source) together with the sha256 of the whole file and the reference
    RX chain's own output for it, so that the identity can be re-checked where /root/reference does not exist."""
    import hashlib
    path = "/root/reference/apps/test.ts"
    raw = np.fromfile(path, np.uint8)
    npk = 2016                                               # 4 superframes of 2k/QAM16/1-2 carry 504 packets each
    head = raw[: npk * 188].copy()
    assert np.all(head.reshape(-1, 188)[:, 0] == 0x47)
    con, cr, tm = R.QAM16, R.C1_2, R.T2k
    ed, rs, ci = R.tx_outer(head)
    tx = R.tx_inner(ci, con, cr, tm, nsym=None)
    nsym = tx["X"].shape[0]
    from dvbt_testlib import channel
    X = channel(tx["X"])
    Y, tags = R.rx_demod(X, con, cr, tm)
    dm = R.rx_demap(Y, con, tm)
    sd, bd = R.rx_deinterleave(dm, tags, con, tm)
    sf = [t for t in tags if t[1] == "superframe_start"][0][0]
    vo, vtags = R.rx_viterbi(bd, con, cr, sf * 1512)
    cd, rd, ts = R.rx_outer(vo, vtags, fixed_rs=True)
    assert len(ts) >= 1504 * 20 and np.array_equal(ts, head[504 * 188: 504 * 188 + len(ts)])
    out = os.path.join(HERE, "apps_test_ts_head.npz")
    np.savez_compressed(out, ts_head=head, sha256_whole_file=np.frombuffer(hashlib.sha256(raw.tobytes()).digest(), np.uint8),
                        whole_file_bytes=np.int64(len(raw)), first_packet=np.int64(504), nsym=np.int64(nsym),
                        reference_rx_bytes=np.int64(len(ts)))
    print("wrote", out, os.path.getsize(out), "bytes; reference RX returned", len(ts) // 188, "packets of", nsym, "symbols")


def main():
    assert R.available() and R.available(True), "build oracle/_ref first (make -C oracle ref)"
    d = {}
    d.update(viterbi_vectors())
    d.update(rs_vectors())
    d.update(demap_vectors())
    d.update(inner_coder_vectors())
    path = os.path.join(HERE, "hotpath_golden.npz")
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path), "bytes,", len(d), "arrays")
    chain_fixture()
    apps_test_ts_fixture()


if __name__ == "__main__":
    main()
