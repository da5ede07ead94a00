"""Executable numpy model of the CUDA Viterbi path (gr_dvbt_b200/csrc/viterbi.cu).

It interprets the *same* generated ACS instruction list the kernel compiles
(gen_viterbi_acs.py: build()/run_ops) and restates the hand-written event logic (survivor
ring, argmax, merge-shortcut traceback, chunking with warm-up and boundary verification),
so that the whole algorithm can be compared with the oracle on a CPU-only box.  It is a
test aid: the GPU tests compare the real kernel with the oracle directly.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr_dvbt_b200", "csrc"))
import gen_viterbi_acs as G  # noqa: E402
import gen_viterbi_acs_h16 as H  # noqa: E402

RATE_K = [1, 2, 3, 5, 7]
RATE_N = [2, 3, 4, 6, 8]
RATE_NTB = [5, 9, 10, 15, 24]
PUNCT = [[1, 1], [1, 1, 0, 1], [1, 1, 0, 1, 1, 0], [1, 1, 0, 1, 1, 0, 0, 1, 1, 0],
         [1, 1, 0, 1, 0, 1, 0, 1, 1, 0, 0, 1, 1, 0]]

_SCHED = None


def sched():
    global _SCHED
    if _SCHED is None:
        _SCHED = G.build()
    return _SCHED


def apk_lut():
    lut = np.zeros(16, np.uint32)
    for nib in range(16):
        s0, v0, s1, v1 = nib & 1, (nib >> 1) & 1, (nib >> 2) & 1, (nib >> 3) & 1
        lut[nib] = v0 * (0x01010000 if s0 else 0x00000101) + v1 * (0x01000100 if s1 else 0x00010001)
    return lut


def depuncture_codes(inp, m, rate, nbytetimes):
    """reference-format bytes (m bits each, MSB first) -> one u32 of 8 step nibbles per byte time.
    nibble i (bits 4i..4i+3) = sym0 | valid0<<1 | sym1<<2 | valid1<<3 for step 8j+i
    (viterbi_decoder_impl.cc:241-256 depuncture order, puncture tables :61-65)."""
    k = RATE_K[rate]
    p = PUNCT[rate]
    inp = np.asarray(inp, np.uint8)
    bits = ((inp[:, None] >> np.arange(m - 1, -1, -1)[None, :]) & 1).reshape(-1)
    codes = np.zeros(nbytetimes, np.uint32)
    pos = 0
    for j in range(nbytetimes):
        w = 0
        for i in range(8):
            ph = (8 * j + i) % k
            nib = 0
            if p[2 * ph]:
                nib |= int(bits[pos]) | 2
                pos += 1
            if p[2 * ph + 1]:
                nib |= (int(bits[pos]) << 2) | 8
                pos += 1
            w |= nib << (4 * i)
        codes[j] = w
    return codes


def event_byte_index(s):
    return ((s & 7) << 2) | (s & 32) | ((s >> 3) & 3)


def decode_chunk(codes, jstart, jend, first_real, ntb, out, save_at=(), init_metrics=None):
    """One thread of the kernel: byte times jstart..jend-1; real outputs for j >= first_real
    go to out[j-ntb].  save_at: byte times at which the post-event metrics are returned.
    init_metrics: None (zero state at step 8*jstart) or (16 words) event-layout metrics to
    resume from *after* the event of byte time jstart-1."""
    S = sched()
    lut = apk_lut()
    one = lambda v: np.array([v], np.uint32)
    M = [one(0) for _ in range(16)]
    P = [one(0) for _ in range(16)]
    ring = np.zeros((ntb, 64), np.uint8)
    trace = np.zeros(ntb, np.int64)
    have_trace = False
    saved = {}
    resume = init_metrics is not None
    j = jstart
    if resume:
        # enter at "after the event of byte time jstart-1": run part2 of that byte time first
        j = jstart - 1
    while j < jend:
        code = int(codes[j]) if j < len(codes) else 0
        apk = [one(lut[(code >> (4 * i)) & 15]) for i in range(8)]
        if resume:
            Mev = [one(x) for x in init_metrics]
            resume = False
        else:
            env = {"ZERO": one(0)}
            for i in range(16):
                env["M[%d]" % i] = M[i]
                env["P[%d]" % i] = P[i]
            for i in range(6):
                env["apk%d" % i] = apk[i]
            G.run_ops(S["part1"], env)
            Mev = [env["M_ev[%d]" % i] for i in range(16)]
            Pev = [env["P_ev[%d]" % i] for i in range(16)]
            # ---------------- event of byte time j
            slot = j % ntb
            row = np.zeros(64, np.uint8)
            for w in range(16):
                for b in range(4):
                    row[4 * w + b] = (int(Pev[w][0]) >> (8 * b)) & 0xFF
            ring[slot] = row
            met = np.zeros(64, np.int64)  # by state
            for s in range(64):
                bi = event_byte_index(s)
                met[s] = (int(Mev[bi >> 2][0]) >> (8 * (bi & 3))) & 0xFF
            assert met.max() < 128
            if j >= first_real:
                s = int(np.argmax(met))  # first maximum, d_viterbi.c:699-711
                merged = False
                for h in range(ntb - 1):
                    q = (j - h) % ntb
                    if h > 0 and have_trace and trace[q] == s:
                        merged = True
                        break
                    trace[q] = s
                    s = int(ring[q][event_byte_index(s)]) >> 2
                qf = (j - (ntb - 1)) % ntb
                if merged:
                    s = int(trace[qf])
                else:
                    trace[qf] = s
                have_trace = True
                out[j - ntb] = ring[qf][event_byte_index(s)]
            # renormalise with a lower bound of the minimum (spread <= 12)
            x = int(Mev[0][0]) & 0xFF
            sub = max(x, 12) - 12
            Mev = [one((int(v[0]) - sub * 0x01010101) & 0xFFFFFFFF) for v in Mev]
            if j in save_at:
                saved[j] = np.array([int(v[0]) for v in Mev], np.uint32)
        env = {"ZERO": one(0)}
        for i in range(16):
            env["M[%d]" % i] = Mev[i]
            env["P[%d]" % i] = one(0)
        env["apk6"], env["apk7"] = apk[6], apk[7]
        G.run_ops(S["part2"], env)
        M = [env["M_nx[%d]" % i] for i in range(16)]
        P = [env["P_nx[%d]" % i] for i in range(16)]
        j += 1
    return saved


_SCHED_H = None


def sched_h16():
    global _SCHED_H
    if _SCHED_H is None:
        _SCHED_H = H.build()
    return _SCHED_H


def brev8(b):
    return int("{:08b}".format(int(b))[::-1], 2)


def decode_chunk_h16(codes, jstart, jend, first_real, ntb, out, save_at=(), init_metrics=None):
    """Same contract as decode_chunk for the halfword schedule (gen_viterbi_acs_h16.py): 32 registers of
    (metric << 8 | path) halfwords, path bytes bit reversed, event layout of that generator."""
    S = sched_h16()
    lut = apk_lut()
    one = lambda v: np.array([v], np.uint32)
    V = [one(0) for _ in range(32)]
    ring = np.zeros((ntb, 64), np.uint8)
    trace = np.zeros(ntb, np.int64)
    have_trace = False
    saved = {}
    resume = init_metrics is not None
    j = jstart
    if resume:
        j = jstart - 1
    while j < jend:
        code = int(codes[j]) if j < len(codes) else 0
        apk = {}
        for i in range(8):
            q = H.addend_words((code >> (4 * i)) & 15, H.lane_xor(H.STEP_LANE_POS[i]))
            for c, v in zip("xyzw", q):
                apk["apk%d.%s" % (i, c)] = one(v)
        if resume:
            Mev = [one(x) for x in init_metrics]
            resume = False
        else:
            env = {"ZERO": one(0)}
            for i in range(32):
                env["V[%d]" % i] = V[i]
            env.update(apk)
            H.run_ops(S["part1"], env)
            Mev = [env["M_ev[%d]" % i] for i in range(16)]
            Pev = [env["P_ev[%d]" % i] for i in range(16)]
            slot = j % ntb
            row = np.zeros(64, np.uint8)
            for w in range(16):
                for b in range(4):
                    row[4 * w + b] = (int(Pev[w][0]) >> (8 * b)) & 0xFF
            ring[slot] = row
            met = np.zeros(64, np.int64)
            for s in range(64):
                bi = H.event_byte_index(s)
                met[s] = (int(Mev[bi >> 2][0]) >> (8 * (bi & 3))) & 0xFF
            assert met.max() < 128
            if j >= first_real:
                s = int(np.argmax(met))
                # the generated best-state search (VITH_ARGMAX) must agree with the reference's scan
                H.run_ops(S["argmax"], env)
                bw = int(env["BEST"][0])
                assert 63 - (max(bw & 0xFFFF, bw >> 16) & 63) == s
                merged = False
                for h in range(ntb - 1):
                    q = (j - h) % ntb
                    if h > 0 and have_trace and trace[q] == s:
                        merged = True
                        break
                    trace[q] = s
                    s = brev8(ring[q][H.event_byte_index(s)]) >> 2
                qf = (j - (ntb - 1)) % ntb
                if merged:
                    s = int(trace[qf])
                else:
                    trace[qf] = s
                have_trace = True
                out[j - ntb] = brev8(ring[qf][H.event_byte_index(s)])
            x = int(Mev[0][0]) & 0xFF
            sub = max(x, 12) - 12
            Mev = [one((int(v[0]) - sub * 0x01010101) & 0xFFFFFFFF) for v in Mev]
            if j in save_at:
                saved[j] = np.array([int(v[0]) for v in Mev], np.uint32)
        env = {"ZERO": one(0)}
        for i in range(16):
            env["M[%d]" % i] = Mev[i]
        env.update(apk)
        H.run_ops(S["part2"], env)
        V = [env["V_nx[%d]" % i] for i in range(32)]
        j += 1
    return saved


_SCHED_HB = None


def sched_h16b():
    global _SCHED_HB
    if _SCHED_HB is None:
        import gen_viterbi_acs_h16b as HB
        _SCHED_HB = (HB, HB.build())
    return _SCHED_HB


def decode_chunk_h16b(codes, jstart, jend, first_real, ntb, out, save_at=(), init_metrics=None):
    """Same contract for the second halfword schedule (gen_viterbi_acs_h16b.py): registers zipped directly at the
    event (no metric words), best state on the zipped registers, renormalisation folded into step 7.  The saved
    boundary vectors are the 32 un-renormalised zipped registers."""
    HB, S = sched_h16b()
    one = lambda v: np.array([v & 0xFFFFFFFF], np.uint32)
    V = [one(0) for _ in range(32)]
    ring = np.zeros((ntb, 64), np.uint8)
    trace = np.zeros(ntb, np.int64)
    have_trace = False
    saved = {}
    resume = init_metrics is not None
    j = jstart
    if resume:
        j = jstart - 1
    while j < jend:
        code = int(codes[j]) if j < len(codes) else 0
        apk = {}
        for i in range(8):
            q = H.addend_words((code >> (4 * i)) & 15, H.lane_xor(H.STEP_LANE_POS[i]))
            for c, v in zip("xyzw", q):
                apk["apk%d.%s" % (i, c)] = one(v)
        if resume:
            Z = [one(int(x)) for x in init_metrics]
            resume = False
        else:
            env = {"ZERO": one(0)}
            for i in range(32):
                env["V[%d]" % i] = V[i]
            env.update(apk)
            H.run_ops(S["part1"], env)
            Z = [env["Z[%d]" % i] for i in range(32)]
            Pev = [env["P_ev[%d]" % i] for i in range(16)]
            slot = j % ntb
            row = np.zeros(64, np.uint8)
            for w in range(16):
                for b in range(4):
                    row[4 * w + b] = (int(Pev[w][0]) >> (8 * b)) & 0xFF
            ring[slot] = row
            met = np.zeros(64, np.int64)
            for s in range(64):
                r, h = HB.z_position(s)
                assert (int(Z[r][0]) >> (16 * h)) & 0xFF == 0          # empty path bytes
                met[s] = (int(Z[r][0]) >> (8 + 16 * h)) & 0xFF
            assert met.max() < 128                                      # what the sign-replicating zip relies on
            if j >= first_real:
                s = int(np.argmax(met))
                H.run_ops(S["argmax"], env)
                bw = int(env["BEST"][0])
                assert 63 - (max(bw & 0xFFFF, bw >> 16) & 63) == s
                merged = False
                for h in range(ntb - 1):
                    q = (j - h) % ntb
                    if h > 0 and have_trace and trace[q] == s:
                        merged = True
                        break
                    trace[q] = s
                    s = brev8(ring[q][H.event_byte_index(s)]) >> 2
                qf = (j - (ntb - 1)) % ntb
                if merged:
                    s = int(trace[qf])
                else:
                    trace[qf] = s
                have_trace = True
                out[j - ntb] = brev8(ring[qf][H.event_byte_index(s)])
            if j in save_at:
                saved[j] = np.array([int(v[0]) for v in Z], np.uint32)
        x = (int(Z[0][0]) >> 8) & 0xFF                                  # metric of state 0
        sub = max(x, 12) - 12
        env = {"ZERO": one(0), "NEG2": one(((0 - sub) & 0xFF) * 0x01000100), "BITSUB": one(0x00010001 - sub * 0x01000100)}
        for i in range(32):
            env["Z[%d]" % i] = Z[i]
        env.update(apk)
        H.run_ops(S["part2"], env)
        V = [env["V_nx[%d]" % i] for i in range(32)]
        j += 1
    return saved


def normalised_h16b(words):
    b = np.array([(int(w) >> sh) & 0xFF for w in words for sh in (8, 24)], np.int64)
    return b - b.min()


def normalised(words):
    b = np.array([(int(w) >> (8 * i)) & 0xFF for w in words for i in range(4)], np.int64)
    return b - b.min()


def decode_stream(inp, m, rate, chunk_bytes, warm, force_fixup=False, schedule="swar"):
    """Whole stream from a reset, chunked like the kernel; returns (out bytes, n_fixups)."""
    decode_chunk = globals()[{"h16": "decode_chunk_h16", "h16b": "decode_chunk_h16b"}.get(schedule, "decode_chunk")]
    norm = normalised_h16b if schedule == "h16b" else normalised
    k, n, ntb = RATE_K[rate], RATE_N[rate], RATE_NTB[rate]
    nbt = len(inp) * m * k // (8 * n)  # byte times available
    codes = depuncture_codes(inp, m, rate, nbt)
    nout = nbt - ntb
    out = np.zeros(max(nout, 0), np.uint8)
    bounds = list(range(0, nout, chunk_bytes)) + [nout]
    G_, F_ = {}, {}
    for c in range(len(bounds) - 1):
        A, B = bounds[c], bounds[c + 1]
        js = max(0, A - warm)
        sv = decode_chunk(codes, js, B + ntb, A + ntb, ntb, out, save_at=(A, B))
        if js > 0:
            G_[c] = sv[A]
        if B in sv:
            F_[c] = sv[B]
    fix = 0
    for c in range(1, len(bounds) - 1):
        A, B = bounds[c], bounds[c + 1]
        if c in G_ and (force_fixup or not np.array_equal(norm(G_[c]), norm(F_[c - 1]))):
            fix += 1
            sv = decode_chunk(codes, A + 1, B + ntb, A + ntb, ntb, out, save_at=(B,), init_metrics=F_[c - 1])
            if B in sv:
                F_[c] = sv[B]
    return out, fix
