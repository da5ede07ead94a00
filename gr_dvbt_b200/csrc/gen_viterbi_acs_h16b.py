#!/usr/bin/env python3
"""Generator for the second halfword ACS schedule ("h16b"): the trellis steps of gen_viterbi_acs_h16.py with a
cheaper event.  Writes gr_dvbt_b200/csrc/viterbi_acs_h16b_gen.cuh (committed; regenerate with
`python gr_dvbt_b200/csrc/gen_viterbi_acs_h16b.py`).  Selected with DVBT_B200_VIT_ACS=h16b; the default stays the
h16 schedule until this one has been measured on a B200 (it was written without GPU access: validated against the
oracle with the numpy interpreter only, tests/test_viterbi_schedule.py).

Why.  The h16 kernel is bound by the ALU pipe (LOP3 / PRMT / VIADDMNMX share it; ncu 74 % active) with the issue
rate right behind (72 %): per byte time 904 SASS instructions, ~470 of them on the ALU pipe, of which only 256 are
the add-compare-selects.  The event of h16 (after step 6) unzips the 32 halfword registers into 16 metric words +
16 path words, renormalises the metric words, tags a copy of every register with 63 - state for the best-state
search, and zips the metric words back: 64 PRMT + 32 LOP3 + 16 subtractions + 32 PRMT.

What changes.
  * No metric words.  The registers are zipped straight from the step-6 layout (lane = state bit 5) to the step-7
    layout (lane = state bit 0) with EMPTY path bytes by one PRMT each: a selector nibble with bit 3 set replicates
    the sign bit of the selected byte, and a metric byte is < 128 at the event (the model asserts it: metrics are
    renormalised every byte time and spread by at most 12 + 16), so "sign of the metric byte" is the zero byte the
    path field needs.  The path words for the ring row are extracted as before (16 PRMT).  Event PRMT: 48 instead of 96.
  * Best state on the zipped registers: their path bytes are zero, so tagging is an ADDITION of the constant
    (63 - s_lo) | (63 - s_hi) << 16, which the kernel issues as IMAD on the FMA pipe (VITH_ADDC) instead of LOP3 on
    the ALU pipe; then the same VIMNMX3.U16x2 tournament.
  * Renormalisation folded into step 7.  Subtracting `sub` from all 64 metrics is the same as subtracting it from
    the branch-metric addends of the next step (4 + 4 words instead of 32 registers):
        plain'[L]   = viaddmax(Q[L], NEG2, 0)          per-halfword (A - sub) << 8 mod 2^16        (4 ALU)
        withbit'[L] = Q[L] + (bit - sub * 0x01000100)  32-bit two's complement: the 32-bit sum of a register and
                                                       this word is exact because both halfword results are
                                                       non-negative (metric >= sub)                (4 FMA)
    (__viaddmax_u16x2 adds modulo 2^16 per halfword: add.u16x2 + max.u16x2.)
    The boundary metrics G / F are therefore the un-renormalised zipped registers (32 words); the verify kernel
    compares min-normalised vectors, so any common offset is immaterial.
Net per byte time (before ptxas): ALU pipe -16 PRMT -32 LOP3 +4 VIADDMNMX = -44 of ~470, issue -26 of ~904.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_viterbi_acs_h16 as H  # noqa: E402


def zip_states(w):
    """states of the two zipped registers made from step-6 registers 2w and 2w+1 (lane = bit 5 there):
    za = (2w, 2w+1), zb = (2w+32, 2w+33) -- lane = bit 0"""
    return (2 * w, 2 * w + 1), (2 * w + 32, 2 * w + 33)


def z_index():
    """order of the zipped registers Z[0..31] (also the G / F vector): Z[2w] = za(w), Z[2w+1] = zb(w)"""
    out = []
    for w in range(16):
        za, zb = zip_states(w)
        out += [za, zb]
    return out


def build():
    L_start = H.layout(2)
    # ---- part 1: steps 1..6 with the lane move after step 3 (as h16), path words, zip
    g = H.Gen()
    V = [("V[%d]" % i, st) for i, st in enumerate(L_start)]
    V = g.butterfly_step(V, "apk0", 2, H.STEP_LANE_POS[0]); H.check_layout(V, 3)
    V = g.butterfly_step(V, "apk1", 3, H.STEP_LANE_POS[1]); H.check_layout(V, 4)
    V = g.butterfly_step(V, "apk2", 4, H.STEP_LANE_POS[2]); H.check_layout(V, 5)
    V = g.swap(V, 5, 2); H.check_layout(V, 2)
    V = g.butterfly_step(V, "apk3", 5, H.STEP_LANE_POS[3]); H.check_layout(V, 3)
    V = g.butterfly_step(V, "apk4", 6, H.STEP_LANE_POS[4]); H.check_layout(V, 4)
    V = g.butterfly_step(V, "apk5", 7, H.STEP_LANE_POS[5]); H.check_layout(V, 5)
    idx = {w[1]: w[0] for w in V}
    for w in range(16):
        s0, s1, s2, s3 = H.event_word_states(w)       # the ring row keeps the h16 event layout
        va, vb = idx[(s0, s1)], idx[(s2, s3)]
        g.emit("prmt", "P_ev[%d]" % w, va, vb, 0x6420)
    for w in range(16):
        ra, rb = idx[(2 * w, 2 * w + 32)], idx[(2 * w + 1, 2 * w + 33)]
        # byte 1 / 3 of a register = metric of its lo / hi state; nibble 8|i = sign of byte i = 0x00 (metric < 128)
        g.emit("prmt", "Z[%d]" % (2 * w), ra, rb, 0x5919)       # (0, m(2w), 0, m(2w+1))
        g.emit("prmt", "Z[%d]" % (2 * w + 1), ra, rb, 0x7B3B)   # (0, m(2w+32), 0, m(2w+33))
    part1 = g.ops

    # ---- best state: Z + tag (path bytes are zero), 16-bit maximum over all halfwords
    g = H.Gen()
    cur = []
    for r, (slo, shi) in enumerate(z_index()):
        t = g.new("q")
        g.emit("addc", t, "Z[%d]" % r, (63 - slo) | ((63 - shi) << 16))
        cur.append(t)
    while len(cur) > 1:
        nxt = []
        while len(cur) >= 3:
            t = g.new("q")
            g.emit("vimax3", t, cur[0], cur[1], cur[2])
            nxt.append(t)
            cur = cur[3:]
        if len(cur) == 2 and not nxt:
            t = g.new("q")
            g.emit("vimax3", t, cur[0], cur[1], cur[1])
            nxt.append(t)
            cur = []
        nxt += cur
        cur = nxt
    g.emit("mov", "BEST", cur[0])
    argmax = g.ops

    # ---- part 2: step 7 with the renormalisation folded into its addends, step 8
    g = H.Gen()
    for c in "xyzw":
        g.emit("viaddmax", "n6%s" % c, "apk6.%s" % c, "NEG2", "ZERO")    # plain'[L]
        g.emit("add", "b6%s" % c, "apk6.%s" % c, "BITSUB")               # withbit'[L], BITSUB = bit(0) - sub * 0x01000100
    V = [("Z[%d]" % r, st) for r, st in enumerate(z_index())]
    H.check_layout(V, 0)
    V = butterfly_step_folded(g, V, 0); H.check_layout(V, 1)
    V = g.butterfly_step(V, "apk7", 1, H.STEP_LANE_POS[7]); H.check_layout(V, 2)
    idx = {w[1]: w[0] for w in V}
    for i, st in enumerate(L_start):
        g.emit("mov", "V_nx[%d]" % i, idx[st])
    part2 = g.ops
    return dict(part1=part1, part2=part2, argmax=argmax, L_start=L_start)


def butterfly_step_folded(g, V, lane_pos):
    """step 7: as Gen.butterfly_step, with the prepared addends n6? (plain, renormalising) and b6? (with the
    decision bit, renormalising)"""
    idx = {w[1]: w[0] for w in V}
    out = []
    f = H.lane_xor(lane_pos)
    for name, st in sorted(V, key=lambda w: w[1]):
        if st[0] >= 32:
            continue
        hi = (st[0] + 32, st[1] + 32)
        v0, v1 = name, idx[hi]
        la = (H.label(st[0]), H.label(st[1]))
        lb = (3 - la[0], 3 - la[1])
        assert la[1] == la[0] ^ f and lb[1] == lb[0] ^ f
        c1, c2 = g.new("c"), g.new("c")
        e, o = g.new("V"), g.new("V")
        g.emit("add", c1, v1, "b6" + "xyzw"[lb[0]])
        g.emit("viaddmax", e, v0, "n6" + "xyzw"[la[0]], c1)
        g.emit("add", c2, v1, "b6" + "xyzw"[la[0]])
        g.emit("viaddmax", o, v0, "n6" + "xyzw"[lb[0]], c2)
        out.append((e, tuple((2 * s) & 63 for s in st)))
        out.append((o, tuple((2 * s + 1) & 63 for s in st)))
    assert len(out) == len(V)
    return out


def z_position(s):
    """(register, halfword) of state s in the zipped layout"""
    return 2 * ((s & 31) >> 1) + (s >> 5), s & 1


HEADER = """// GENERATED by gr_dvbt_b200/csrc/gen_viterbi_acs_h16b.py -- do not edit by hand.
// Second halfword ACS schedule (DVBT_B200_VIT_ACS=h16b): the steps of viterbi_acs_h16_gen.cuh with an event that
// zips the registers directly (empty path bytes by sign replication of the metric bytes), searches the best state
// on the zipped registers with FMA-pipe tagging, and folds the renormalisation into the addends of step 7.
// See the generator's docstring; reference semantics: /root/reference/lib/d_viterbi.c:461-576, 680-735.
// Requires vit_prmt (viterbi_acs_gen.cuh) and the VITH_* macros of viterbi_acs_h16_gen.cuh.
#pragma once
#include <stdint.h>

"""


def main():
    res = build()
    out = [HEADER]
    out.append("// Steps 1..6 of a byte time.  In: V[32] in start layout.  Out: P_ev[16] = path bytes in the h16 event layout\n"
               "// (the ring row), Z[32] = the registers zipped for step 7: Z[2w] = states (2w, 2w+1), Z[2w+1] = (2w+32, 2w+33),\n"
               "// metric << 8 per halfword, path bytes empty, NOT renormalised.\n"
               "#define VITB_ACS_PART1(V, Z, P_ev, apk0, apk1, apk2, apk3, apk4, apk5) \\\n")
    body = H.emit_cuda(res["part1"])
    out.append("  do { \\\n" + "\n".join(l + " \\" for l in body.split("\n")) + "\n  } while (0)\n\n")
    out.append("// Steps 7, 8 from the zipped registers.  NEG2 = ((0 - sub) & 0xff) * 0x01000100, BITSUB = 0x00010001 - sub * 0x01000100\n"
               "// (sub = renormalisation amount, at most the smallest metric).  Out: V_nx[32] in start layout.\n"
               "#define VITB_ACS_PART2(Z, V_nx, apk6, apk7, NEG2, BITSUB) \\\n")
    body = H.emit_cuda(res["part2"]).replace(", ZERO)", ", 0u)")
    out.append("  do { \\\n" + "\n".join(l + " \\" for l in body.split("\n")) + "\n  } while (0)\n\n")
    out.append("// Best state at the event from the zipped registers: BEST = 16-bit maximum of (metric << 8 | 63 - state).\n"
               "#define VITB_ARGMAX(Z, BEST) \\\n")
    body = H.emit_cuda(res["argmax"])
    out.append("  do { \\\n" + "\n".join(l + " \\" for l in body.split("\n")) + "\n  } while (0)\n\n")
    for s in range(64):
        r, h = z_position(s)
        assert z_index()[r][h] == s
    out.append("// metric of state s in a zipped vector Z[32] (G / F format of this schedule)\n"
               "__device__ __forceinline__ uint32_t vitb_metric(const uint32_t *Z, uint32_t s) {\n"
               "  return (Z[2u * ((s & 31u) >> 1) + (s >> 5)] >> (8u + 16u * (s & 1u))) & 0xffu;\n}\n")
    cnt = {}
    for part in ("part1", "part2", "argmax"):
        for o in res[part]:
            if o[0] != "mov":
                cnt[o[0]] = cnt.get(o[0], 0) + 1
    out.append("// op counts per byte time (before ptxas): %s\n" % ", ".join("%s %d" % kv for kv in sorted(cnt.items())))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "viterbi_acs_h16b_gen.cuh")
    with open(path, "w") as f:
        f.write("".join(out))
    print("wrote", path, cnt)


if __name__ == "__main__":
    main()
