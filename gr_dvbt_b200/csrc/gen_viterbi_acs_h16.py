#!/usr/bin/env python3
"""Generator for the halfword ("h16") ACS schedule of the sm_100a Viterbi kernel.

Writes gr_dvbt_b200/csrc/viterbi_acs_h16_gen.cuh (committed; regenerate with
`python gr_dvbt_b200/csrc/gen_viterbi_acs_h16.py`).  Like gen_viterbi_acs.py the instruction list can be
interpreted with numpy (run_ops), so the schedule is checked on a CPU against the oracle
(tests/test_viterbi_schedule.py) before any GPU time is spent.

Idea.  sm_100a has the 16x2 dynamic-programming instruction VIADDMNMX.U16x2 (`__viaddmax_u16x2(a, b, c)` =
max(a + b, c) per halfword), issued on the ALU pipe at the rate of LOP3/PRMT, next to IMAD on the FMA pipe
(tools/ubench/dpx_rate.cu, profiles/r01_dpx_rate.txt).  A state is kept as ONE halfword

        metric << 8 | path byte

so that one unsigned 16-bit maximum selects the survivor's metric AND its path at once:

        cand1 = V[i+32] + (bm1 << 8 | 1 << j)          IMAD.IADD  (FMA pipe)
        new   = max(V[i] + (bm0 << 8), cand1)          VIADDMNMX  (ALU pipe)

i.e. one ALU-pipe instruction per two states and step, where the byte-SWAR schedule (gen_viterbi_acs.py) needs
four per four states (compare word, sign mask, two selects).

Tie rule.  The reference takes the i+32 predecessor when the metrics are equal (decision = (int8)(m0-m1) > 0,
d_viterbi.c:508-511).  The decision bit of step j after the last path clear goes to bit j of the path byte
(first decision in bit 0), so cand1's path field has bit j set while cand0's has bit j clear and neither has a
higher bit set: with equal metrics cand1 is the larger halfword, and the maximum picks it.  The path byte is
therefore the bit reversal of the reference's (which shifts left, d_viterbi.c:513-515); the traceback reads
state = brev8(byte) >> 2 and the output byte is brev8(byte) (d_viterbi.c:719,724).

Layout.  A state index is 6 bits; a step is s' = ((s << 1) | u) & 63.  Five bits select the register, one the
halfword lane.  A butterfly needs bit 5 to be a register bit, and every step moves the lane bit up by one, so
between two lane moves there are at most 5 steps.  Per byte time (positions of the lane bit):
   start 2 | s1 3 | s2 4 | s3 5 | swap 5->2 | s4 3 | s5 4 | s6 5 | EVENT: unzip the 32 halfword registers into
   16 path words (ring row) + 16 metric words (4 states per word: argmax tournament, renormalisation, G/F), then
   zip the metric words back with the lane at bit 0 and empty path bytes | s7 1 | s8 2.
The unzip/zip pair costs what one lane move plus the two extractions would, clears the paths for free and
gives the event code the same 4-states-per-word format as the byte-SWAR kernel.

Branch metrics.  A[L], L = 2*c0 + c1, is the number of agreeing, non-erased code bits of a branch with label L
(see gen_viterbi_acs.py).  The two lanes of a register differ in one state bit, so their labels differ by a fixed
XOR f that depends only on the lane position: per step there are 4 distinct addend words
Q[L] = A[L] << 8 | A[L ^ f] << 24.  The kernel reads all four with one 16-byte shared-memory load from a table
indexed by (f, step code nibble); the words that also carry the decision bit cost one addition each.
"""
import os

POLYA, POLYB = 0x4F, 0x6D


def parity(v):
    return bin(v).count("1") & 1


def label(i):
    """index into APK of butterfly i (0..31): L = 2*c0 + c1"""
    return 2 * parity((2 * i) & POLYA) + parity((2 * i) & POLYB)


class Gen:
    def __init__(self):
        self.ops = []
        self.tmp = 0

    def new(self, prefix="t"):
        self.tmp += 1
        return "%s%d" % (prefix, self.tmp)

    def emit(self, *op):
        self.ops.append(op)

    def butterfly_step(self, V, apk, jbit, lane_pos):
        """V: list of (name, (s_lo, s_hi)).  One trellis step; the decision goes to path bit `jbit`.
        apk: name of the uint4 holding the addend words Q[0..3] for this step and lane position."""
        idx = {w[1]: w[0] for w in V}
        out = []
        plain, withbit = {}, {}
        bitconst = 0x00010001 << jbit
        f = lane_xor(lane_pos)

        def addend(labels, bit):
            assert labels[1] == labels[0] ^ f
            plain[labels] = "%s.%s" % (apk, "xyzw"[labels[0]])
            if not bit:
                return plain[labels]
            if labels not in withbit:
                r = self.new("bb")
                self.emit("addc", r, plain[labels], bitconst)
                withbit[labels] = r
            return withbit[labels]

        for name, st in sorted(V, key=lambda w: w[1]):
            if st[0] >= 32:
                continue
            assert st[1] < 32
            hi = (st[0] + 32, st[1] + 32)
            v0, v1 = name, idx[hi]
            la = (label(st[0]), label(st[1]))
            lb = (3 - la[0], 3 - la[1])
            c1, c2 = self.new("c"), self.new("c")
            e, o = self.new("V"), self.new("V")
            self.emit("add", c1, v1, addend(lb, True))
            self.emit("viaddmax", e, v0, addend(la, False), c1)   # state 2s:   max(V[s] + A[L], V[s+32] + A[3-L] + bit)
            self.emit("add", c2, v1, addend(la, True))
            self.emit("viaddmax", o, v0, addend(lb, False), c2)   # state 2s+1: max(V[s] + A[3-L], V[s+32] + A[L] + bit)
            out.append((e, tuple((2 * s) & 63 for s in st)))
            out.append((o, tuple((2 * s + 1) & 63 for s in st)))
        assert len(out) == len(V)
        return out

    def swap(self, V, lane_pos, word_pos):
        """exchange the lane bit (state bit lane_pos) with the register bit at state position word_pos"""
        idx = {w[1]: w[0] for w in V}
        out = []
        for name, st in sorted(V, key=lambda w: w[1]):
            if (st[0] >> word_pos) & 1:
                continue
            st1 = tuple(s | (1 << word_pos) for s in st)
            r0, r1 = name, idx[st1]
            a, b = self.new("x"), self.new("x")
            self.emit("prmt", a, r0, r1, 0x5410)   # (r0.lo, r1.lo): lane bit = word_pos, bit lane_pos = 0
            self.emit("prmt", b, r0, r1, 0x7632)   # (r0.hi, r1.hi)
            out.append((a, (st[0], st1[0])))
            out.append((b, (st[1], st1[1])))
        assert len(out) == len(V)
        return out


def lane_xor(lane_pos):
    """label(state with bit lane_pos set) = label(state) ^ lane_xor(lane_pos)"""
    return 2 * ((POLYA >> (lane_pos + 1)) & 1) + ((POLYB >> (lane_pos + 1)) & 1)


# lane position before each step, in macro-argument order (steps 1..6 of part 1, then 7, 8 of part 2)
STEP_LANE_POS = [2, 3, 4, 2, 3, 4, 0, 1]


def layout(lane_pos):
    """canonical 32 registers for the lane bit at state position lane_pos, ordered by the other 5 bits"""
    rest = [p for p in range(6) if p != lane_pos]
    regs = []
    for w in range(32):
        base = sum(((w >> i) & 1) << rest[i] for i in range(5))
        regs.append((base, base | (1 << lane_pos)))
    return regs


def check_layout(V, lane_pos):
    assert sorted(w[1] for w in V) == sorted(layout(lane_pos)), lane_pos


# event layout (4 states per word): word w = state bits (4,3,2,1), byte = 2*bit0 + bit5
def event_word_states(w):
    base = w << 1
    return (base, base | 32, base | 1, base | 33)


def build():
    L_start = layout(2)
    # ---- part 1: steps 1..6, lane move after step 3, unzip
    g = Gen()
    V = [("V[%d]" % i, st) for i, st in enumerate(L_start)]
    V = g.butterfly_step(V, "apk0", 2, STEP_LANE_POS[0]); check_layout(V, 3)
    V = g.butterfly_step(V, "apk1", 3, STEP_LANE_POS[1]); check_layout(V, 4)
    V = g.butterfly_step(V, "apk2", 4, STEP_LANE_POS[2]); check_layout(V, 5)
    V = g.swap(V, 5, 2); check_layout(V, 2)
    V = g.butterfly_step(V, "apk3", 5, STEP_LANE_POS[3]); check_layout(V, 3)
    V = g.butterfly_step(V, "apk4", 6, STEP_LANE_POS[4]); check_layout(V, 4)
    V = g.butterfly_step(V, "apk5", 7, STEP_LANE_POS[5]); check_layout(V, 5)
    idx = {w[1]: w[0] for w in V}
    for w in range(16):
        s0, s1, s2, s3 = event_word_states(w)
        va, vb = idx[(s0, s1)], idx[(s2, s3)]
        g.emit("prmt", "P_ev[%d]" % w, va, vb, 0x6420)   # path bytes of (va.lo, va.hi, vb.lo, vb.hi)
        g.emit("prmt", "M_ev[%d]" % w, va, vb, 0x7531)   # their metrics
    # the registers themselves, for the best-state search: V_ev[r] holds states (r, r + 32)
    for r in range(32):
        g.emit("mov", "V_ev[%d]" % r, idx[(r, r + 32)])
    part1 = g.ops

    # ---- best state (d_viterbi.c:699-711: lowest index with strictly greatest metric): replace the path byte of
    # every halfword by 63 - state and take the 16-bit maximum of all 64 halfwords
    g = Gen()
    cur = []
    for r in range(32):
        t = g.new("q")
        g.emit("lop3or", t, "V_ev[%d]" % r, 0xFF00FF00, (63 - r) | ((63 - (r + 32)) << 16))
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

    # ---- part 2: zip the metric words (lane = bit 0, paths empty), steps 7, 8
    g = Gen()
    V = []
    for w in range(16):
        s0, s1, s2, s3 = event_word_states(w)
        a, b = g.new("z"), g.new("z")
        g.emit("prmt", a, "M[%d]" % w, "ZERO", 0x2404)   # (0, m(s0), 0, m(s2)): states s0 (bit0=0), s2 (bit0=1), bit 5 = 0
        g.emit("prmt", b, "M[%d]" % w, "ZERO", 0x3414)   # (0, m(s1), 0, m(s3)): bit 5 = 1
        V += [(a, (s0, s2)), (b, (s1, s3))]
    check_layout(V, 0)
    V = g.butterfly_step(V, "apk6", 0, STEP_LANE_POS[6]); check_layout(V, 1)
    V = g.butterfly_step(V, "apk7", 1, STEP_LANE_POS[7]); check_layout(V, 2)
    idx = {w[1]: w[0] for w in V}
    for i, st in enumerate(L_start):
        g.emit("mov", "V_nx[%d]" % i, idx[st])
    part2 = g.ops
    return dict(part1=part1, part2=part2, argmax=argmax, L_start=L_start)


def addend_words(nib, f):
    """Q[0..3] for step code nibble nib = sym0 | valid0<<1 | sym1<<2 | valid1<<3 and lane XOR f"""
    s0, v0, s1, v1 = nib & 1, (nib >> 1) & 1, (nib >> 2) & 1, (nib >> 3) & 1
    A = [v0 * int((L >> 1) == s0) + v1 * int((L & 1) == s1) for L in range(4)]
    return [(A[L] << 8) | (A[L ^ f] << 24) for L in range(4)]


def event_byte_index(s):
    """ring-row / metric-vector byte index of state s in the event layout"""
    return (((s >> 1) & 15) << 2) | ((s & 1) << 1) | (s >> 5)


def event_state(w, b):
    """state held by byte b of event word w"""
    return (w << 1) | (b >> 1) | ((b & 1) << 5)


# --------------------------------------------------------------------------------------
# numpy interpreter (CPU validation of the schedule)
# --------------------------------------------------------------------------------------
def run_ops(ops, env):
    import numpy as np
    u32 = np.uint32

    def prmt(a, b, sel):
        src = [(a >> u32(8 * i)) & u32(0xFF) for i in range(4)] + [(b >> u32(8 * i)) & u32(0xFF) for i in range(4)]
        out = np.zeros_like(a)
        for lane in range(4):
            nib = (sel >> (4 * lane)) & 0xF
            byte = src[nib & 7]
            if nib & 8:   # prmt.b32 default mode: bit 3 of a selector nibble replicates the sign bit of the byte
                byte = (byte >> u32(7)) * u32(0xFF)
            out |= byte << u32(8 * lane)
        return out

    for op in ops:
        k, d = op[0], op[1]
        if k == "prmt":
            env[d] = prmt(env[op[2]], env[op[3]], op[4])
        elif k == "add":
            env[d] = (env[op[2]] + env[op[3]]).astype(u32)
        elif k == "addc":
            env[d] = (env[op[2]] + u32(op[3])).astype(u32)
        elif k == "viaddmax":
            a, b, c = env[op[2]], env[op[3]], env[op[4]]
            lo = np.maximum((a + b) & u32(0xFFFF), c & u32(0xFFFF))
            hi = np.maximum(((a >> u32(16)) + (b >> u32(16))) & u32(0xFFFF), c >> u32(16))
            env[d] = (lo | (hi << u32(16))).astype(u32)
        elif k == "lop3or":
            env[d] = ((env[op[2]] & u32(op[3])) | u32(op[4])).astype(u32)
        elif k == "vimax3":
            a, b, c = env[op[2]], env[op[3]], env[op[4]]
            lo = np.maximum(np.maximum(a & u32(0xFFFF), b & u32(0xFFFF)), c & u32(0xFFFF))
            hi = np.maximum(np.maximum(a >> u32(16), b >> u32(16)), c >> u32(16))
            env[d] = (lo | (hi << u32(16))).astype(u32)
        elif k == "mov":
            env[d] = env[op[2]]
        else:
            raise ValueError(k)
    return env


# --------------------------------------------------------------------------------------
# CUDA emitter
# --------------------------------------------------------------------------------------
def emit_cuda(ops, indent="  "):
    lines = []
    declared = set()

    def dst(x):
        if "[" in x or x in declared or x == "BEST":   # arrays and macro parameters are declared by the caller
            return x
        declared.add(x)
        return "uint32_t " + x

    for op in ops:
        k, d = op[0], op[1]
        if k == "prmt":
            b = "0u" if op[3] == "ZERO" else op[3]
            lines.append("%s = vit_prmt(%s, %s, 0x%04xu);" % (dst(d), op[2], b, op[4]))
        elif k == "add":
            lines.append("%s = VITH_ADD(%s, %s);" % (dst(d), op[2], op[3]))
        elif k == "addc":
            lines.append("%s = VITH_ADDC(%s, 0x%08xu);" % (dst(d), op[2], op[3]))
        elif k == "viaddmax":
            lines.append("%s = __viaddmax_u16x2(%s, %s, %s);" % (dst(d), op[2], op[3], op[4]))
        elif k == "lop3or":
            assert op[3] == 0xFF00FF00
            lines.append("%s = (%s & VITH_HIMASK) | 0x%08xu;" % (dst(d), op[2], op[4]))
        elif k == "vimax3":
            lines.append("%s = __vimax3_u16x2(%s, %s, %s);" % (dst(d), op[2], op[3], op[4]))
        elif k == "mov":
            lines.append("%s = %s;" % (dst(d), op[2]))
        else:
            raise ValueError(k)
    return "\n".join(indent + l for l in lines)


HEADER = """// GENERATED by gr_dvbt_b200/csrc/gen_viterbi_acs_h16.py -- do not edit by hand.
// Register-resident 64-state ACS schedule on halfwords (metric << 8 | path byte, 2 states per 32-bit register):
// one VIADDMNMX.U16x2 selects a survivor's metric and path at once.  See the generator's docstring and DESIGN.md
// (K1); reference semantics: /root/reference/lib/d_viterbi.c:461-576 (d_viterbi_butterfly2_sse2).
// Requires vit_prmt from viterbi_acs_gen.cuh.
#pragma once
#include <stdint.h>

// the predecessor-side addition of a butterfly (plain: ptxas issues it as IMAD.IADD on the FMA pipe)
#ifndef VITH_ADD
#define VITH_ADD(a, b) ((a) + (b))
#endif
// branch-metric word + decision-bit constant (4 per step).  Plain by default; the kernel may force an IMAD.
#ifndef VITH_ADDC
#define VITH_ADDC(a, c) ((a) + (c))
#endif
// 0xff00ff00 (the metric bytes of a register).  The including kernel may define it as a run-time value held in a
// register: with two immediates ptxas splits (x & mask) | k into two LOP3.
#ifndef VITH_HIMASK
#define VITH_HIMASK 0xff00ff00u
#endif

"""


def main():
    res = build()
    out = [HEADER]
    out.append("// Steps 1..6 of a byte time.  In: V[32] in start layout (register r, lane l hold the state whose bit 2 is l and\n"
               "// whose other bits, low to high, are r).  Out: M_ev[16], P_ev[16] = metrics / path bytes, 4 states per word, in\n"
               "// event layout: word w = state bits (4,3,2,1), byte = 2*bit0 + bit5; V_ev[32] = the halfword registers at the\n"
               "// event, V_ev[r] = states (r, r+32).  Path bytes are bit reversed with respect to the reference's (first\n"
               "// decision after the clear in bit 0).\n"
               "// apk0..apk5: uint4 of addend words Q[0..3] of steps 1..6 (table row f = VITH_STEP_F(step)).\n"
               "#define VITH_ACS_PART1(V, M_ev, P_ev, V_ev, apk0, apk1, apk2, apk3, apk4, apk5) \\\n")
    body = emit_cuda(res["part1"])
    out.append("  do { \\\n" + "\n".join(l + " \\" for l in body.split("\n")) + "\n  } while (0)\n\n")
    out.append("// Event-layout metric words M[16] (after renormalisation) -> halfword registers with empty paths, then steps 7, 8.\n"
               "// Out: V_nx[32] in start layout.\n"
               "#define VITH_ACS_PART2(M, V_nx, apk6, apk7) \\\n")
    body = emit_cuda(res["part2"])
    out.append("  do { \\\n" + "\n".join(l + " \\" for l in body.split("\n")) + "\n  } while (0)\n\n")
    out.append("// Best state at the event: BEST = 16-bit maximum over all states of (metric << 8 | 63 - state), in both halfwords\n"
               "// of the result word (take the larger).  63 - (max & 63) is the lowest state with the greatest metric.\n"
               "#define VITH_ARGMAX(V_ev, BEST) \\\n")
    body = emit_cuda(res["argmax"])
    out.append("  do { \\\n" + "\n".join(l + " \\" for l in body.split("\n")) + "\n  } while (0)\n\n")
    out.append("// lane XOR f of step i (0..5: steps 1..6, 6..7: steps 7, 8): row of the addend table to read\n"
               "#define VITH_STEP_F(i) (%s)\n\n" % " : ".join(["(i) == %d ? %d" % (i, lane_xor(p)) for i, p in enumerate(STEP_LANE_POS[:-1])]
                                                              + ["%d" % lane_xor(STEP_LANE_POS[-1])]))
    for s in range(64):
        bi = event_byte_index(s)
        assert event_word_states(bi >> 2)[bi & 3] == s and event_state(bi >> 2, bi & 3) == s
    # for a fixed byte lane the state grows with the word index (the argmax tournament keeps the lower word on ties)
    for b in range(4):
        assert all(event_state(w, b) < event_state(w + 1, b) for w in range(15))
    out.append("// ring-row byte index of state s in the event layout: 4*((s>>1)&15) + 2*(s&1) + (s>>5)\n"
               "__device__ __forceinline__ uint32_t vith_event_byte_index(uint32_t s) {\n"
               "  return ((s & 30u) << 1) | ((s & 1u) << 1) | (s >> 5);\n}\n"
               "// state held by byte b of event word w\n"
               "__device__ __forceinline__ uint32_t vith_event_state(uint32_t w, uint32_t b) {\n"
               "  return (w << 1) | (b >> 1) | ((b & 1u) << 5);\n}\n")
    cnt = {}
    for part in ("part1", "part2", "argmax"):
        for o in res[part]:
            if o[0] != "mov":
                cnt[o[0]] = cnt.get(o[0], 0) + 1
    out.append("// op counts per byte time (before ptxas): %s\n" % ", ".join("%s %d" % kv for kv in sorted(cnt.items())))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "viterbi_acs_h16_gen.cuh")
    with open(path, "w") as f:
        f.write("".join(out))
    print("wrote", path, cnt)


if __name__ == "__main__":
    main()
