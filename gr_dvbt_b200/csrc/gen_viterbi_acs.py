#!/usr/bin/env python3
"""Generator for the register-resident K=7 ACS schedule of the sm_100a Viterbi kernel.

Writes gr_dvbt_b200/csrc/viterbi_acs_gen.cuh (committed; regenerate with
`python gr_dvbt_b200/csrc/gen_viterbi_acs.py`).  The same instruction list can be
interpreted with numpy (run_ops) so the schedule is checked on a CPU against the oracle
(tests/test_viterbi_schedule.py) before any GPU time is spent.

Design.  One thread decodes one chunk.  The 64 path metrics live in 16
registers, 4 states per register as unsigned bytes (SWAR); the 64 eight-bit survivor
"path bytes" of the reference (d_viterbi.c:513-524) live in 16 more.  Metrics never reach
128 (spread <= 12, renormalised every 8 steps), so
   t = m1 + 0x80808080 - m0        has bit 7 of every byte = (m1 >= m0)
   mask = prmt(t, sign-replicate)  turns that into 0x00/0xff per byte
   new = (m1 & mask) | (m0 & ~mask)
is an exact 4-wide add-compare-select with the reference's tie rule
(decision = (int8)(m0-m1) > 0, d_viterbi.c:508-511): 5 integer ops per 4 states for the
metrics and 2 for the paths.

Layout.  A state index is 6 bits; a trellis step is s' = ((s << 1) | u) & 63.  Four of the
six bits select the register ("word bits"; free, because registers are renamed by the
straight-line code), two select the byte lane ("lane bits" A and B).  A butterfly pairs
states that differ in bit 5 and produces states that differ in bit 0, so it needs bit 5
to be a word bit; every step moves the lane bits up one position.  Per 8-step byte time
each lane bit therefore has to be moved down by 8 positions in total, which costs two
byte-transposes (one PRMT per register each).  Schedule (positions of lane bits A,B):
   start (2,3) | s1 (3,4) | s2 (4,5) swapB 5->0 | s3 (5,1) swapA 5->0 | s4 (1,2) | s5 (2,3)
   | s6 (3,4) | EVENT: ring store, argmax/traceback, renormalise, paths := 0,
   swap A 3->0 and B 4->1 on the metrics only (paths are zero) | s7 (1,2) | s8 (2,3).

Branch metrics.  Per step the kernel looks up APK = bytes A[L], L = 2*c0 + c1, where
A[L] = v0*[c0==sym0] + v1*[c1==sym1] is the number of agreeing, non-erased code bits
(d_viterbi.c:487-501 computes 2-metsvm / 1-metsvm).  The per-register addend words are
PRMT(APK, sel) with a compile-time selector derived from the state labels
c0(i) = parity(2i & 0x4f), c1(i) = parity(2i & 0x6d) (d_viterbi.c:38-39,273-277).
"""
import os
import sys

POLYA, POLYB = 0x4F, 0x6D
H = 0x80808080
# The compare (t = m1 + H - m0) and the path update (2p + 1) are emitted as macros (VIT_CMP / VIT_PATH2) so that
# the kernel can choose, per instantiation, between one ALU-pipe IADD3 and an FMA-pipe form (IMAD with run-time
# multipliers, one extra addition per compare).  With one warp per scheduler the FMA form measured 3 % slower
# (profiles/r01_bench_rx_v4.json vs v3); it is meant for the multi-warp configurations (split survivor ring).


def parity(v):
    return bin(v).count("1") & 1


def label(i):
    """index into APK of butterfly i (0..31): L = 2*c0 + c1"""
    return 2 * parity((2 * i) & POLYA) + parity((2 * i) & POLYB)


class Gen:
    def __init__(self):
        self.ops = []  # (op, dst, srcs...)
        self.tmp = 0

    def new(self, prefix="t"):
        self.tmp += 1
        return "%s%d" % (prefix, self.tmp)

    def emit(self, *op):
        self.ops.append(op)

    # ---- word-level helpers; a "word" is (name, (s0,s1,s2,s3)) with lane states ----
    def butterfly_step(self, M, P, apk):
        """M, P: lists of (name, states). Returns new (M, P) after one trellis step."""
        idxM = {w[1]: w[0] for w in M}
        idxP = {w[1]: w[0] for w in P} if P is not None else None
        newM, newP = [], []
        sel_cache = {}
        for name, st in sorted(M, key=lambda w: w[1]):
            if any(s >= 32 for s in st):
                continue
            assert all(s < 32 for s in st)
            hi = tuple(s + 32 for s in st)
            mi, mj = name, idxM[hi]
            sel_a = sum(label(st[b]) << (4 * b) for b in range(4))
            sel_b = sum((3 - label(st[b])) << (4 * b) for b in range(4))
            for sel in (sel_a, sel_b):
                if sel not in sel_cache:
                    r = self.new("bm")
                    self.emit("prmt", r, apk, "ZERO", sel)
                    rh = self.new("bh")
                    self.emit("addc", rh, r, H)  # addend + 0x80 per byte, shared by every pair of the step (dead code in the ALU form)
                    sel_cache[sel] = (r, rh)
            (a, ah), (b, bh) = sel_cache[sel_a], sel_cache[sel_b]
            m0, m1, m2, m3 = self.new("m"), self.new("m"), self.new("m"), self.new("m")
            # the four branch-metric additions of a butterfly pair are tagged so that the kernel can route two or all
            # four of them to the FMA pipe (VIT_ADDF / VIT_ADDG, see the header text below)
            self.emit("addf", m0, mi, a)
            self.emit("addf", m1, mj, b)
            self.emit("addg", m2, mi, b)
            self.emit("addg", m3, mj, a)
            # compare: t = (mj + b + H) - m0, both steps on the FMA pipe (IMAD) instead of one IADD3 on the ALU pipe
            t0, t1 = self.new("c"), self.new("c")
            self.emit("cmpx", t0, m1, mj, bh, m0)   # t0 = m1 + H - m0 = (mj + bh) - m0
            self.emit("cmpx", t1, m3, mj, ah, m2)
            k0, k1 = self.new("k"), self.new("k")
            self.emit("signmask", k0, t0)
            self.emit("signmask", k1, t1)
            e, o = self.new("M"), self.new("M")
            self.emit("sel", e, m0, m1, k0)  # mask ? m1 : m0
            self.emit("sel", o, m2, m3, k1)
            est = tuple((2 * s) & 63 for s in st)
            ost = tuple((2 * s + 1) & 63 for s in st)
            newM += [(e, est), (o, ost)]
            if P is not None:
                pi, pj = idxP[st], idxP[hi]
                si, sj = self.new("p"), self.new("p")
                self.emit("add", si, pi, pi)
                self.emit("path2", sj, pj)   # 2*pj + 0x01010101
                pe, po = self.new("P"), self.new("P")
                self.emit("sel", pe, si, sj, k0)
                self.emit("sel", po, si, sj, k1)
                newP += [(pe, est), (po, ost)]
        return newM, (newP if P is not None else None)

    def swap(self, W, lane_bit, word_pos, lane_pos):
        """Exchange lane bit `lane_bit` (currently state bit lane_pos) with the word bit
        at state position word_pos.  One PRMT per register."""
        idx = {w[1]: w[0] for w in W}
        out = []
        done = set()
        for name, st in sorted(W, key=lambda w: w[1]):
            if st in done:
                continue
            assert not (st[0] >> word_pos) & 1 or True
            if (st[0] >> word_pos) & 1:
                continue
            st1 = tuple(s | (1 << word_pos) for s in st)
            r0, r1 = name, idx[st1]
            done.add(st)
            done.add(st1)
            if lane_bit == 0:  # lanes (0,1) and (2,3) differ in this bit
                sel0, sel1 = 0x6240, 0x7351
                n0 = (st[0], st1[0], st[2], st1[2])
                n1 = (st[1], st1[1], st[3], st1[3])
            else:  # lanes (0,2) and (1,3)
                sel0, sel1 = 0x5410, 0x7632
                n0 = (st[0], st[1], st1[0], st1[1])
                n1 = (st[2], st[3], st1[2], st1[3])
            a, b = self.new("x"), self.new("x")
            self.emit("prmt", a, r0, r1, sel0)
            self.emit("prmt", b, r0, r1, sel1)
            out += [(a, n0), (b, n1)]
        assert len(out) == len(W)
        return out


def layout(pa, pb):
    """canonical list of 16 lane-state tuples for lane bits at state positions pa (A, lane
    bit 0) and pb (B, lane bit 1); ordered by the remaining 4 bits (ascending)."""
    rest = [p for p in range(6) if p not in (pa, pb)]
    words = []
    for w in range(16):
        base = sum(((w >> i) & 1) << rest[i] for i in range(4))
        words.append(tuple(base | ((b & 1) << pa) | ((b >> 1) << pb) for b in range(4)))
    return words


def check_layout(W, pa, pb):
    got = sorted(w[1] for w in W)
    want = sorted(layout(pa, pb))
    assert got == want, (pa, pb, got[:3], want[:3])


def build():
    """Returns dict with op lists for part1 (steps 1-6 incl. swaps) and part2 (event swap +
    steps 7,8), and the canonical layouts at the three cut points."""
    L_start = layout(2, 3)
    L_event = layout(3, 4)

    # ---- part 1
    g = Gen()
    M = [("M[%d]" % i, st) for i, st in enumerate(L_start)]
    P = [("P[%d]" % i, st) for i, st in enumerate(L_start)]
    M, P = g.butterfly_step(M, P, "apk0"); check_layout(M, 3, 4)
    M, P = g.butterfly_step(M, P, "apk1"); check_layout(M, 4, 5)
    M = g.swap(M, 1, 0, 5); P = g.swap(P, 1, 0, 5); check_layout(M, 4, 0)   # B: 5 -> 0
    M, P = g.butterfly_step(M, P, "apk2"); check_layout(M, 5, 1)
    M = g.swap(M, 0, 0, 5); P = g.swap(P, 0, 0, 5); check_layout(M, 0, 1)   # A: 5 -> 0
    M, P = g.butterfly_step(M, P, "apk3"); check_layout(M, 1, 2)
    M, P = g.butterfly_step(M, P, "apk4"); check_layout(M, 2, 3)
    M, P = g.butterfly_step(M, P, "apk5"); check_layout(M, 3, 4)
    for arr, W in (("M", M), ("P", P)):
        idx = {w[1]: w[0] for w in W}
        for i, st in enumerate(L_event):
            g.emit("mov", "%s_ev[%d]" % (arr, i), idx[st])
    part1 = g.ops

    # ---- part 2 (paths are zero on entry: only the metrics are transposed)
    g = Gen()
    M = [("M[%d]" % i, st) for i, st in enumerate(L_event)]
    M = g.swap(M, 0, 0, 3); check_layout(M, 0, 4)   # A: 3 -> 0
    M = g.swap(M, 1, 1, 4); check_layout(M, 0, 1)   # B: 4 -> 1
    P = [("P[%d]" % i, st) for i, st in enumerate(layout(0, 1))]  # all zero: any naming
    M, P = g.butterfly_step(M, P, "apk6"); check_layout(M, 1, 2)
    M, P = g.butterfly_step(M, P, "apk7"); check_layout(M, 2, 3)
    for arr, W in (("M", M), ("P", P)):
        idx = {w[1]: w[0] for w in W}
        for i, st in enumerate(L_start):
            g.emit("mov", "%s_nx[%d]" % (arr, i), idx[st])
    part2 = g.ops
    return dict(part1=part1, part2=part2, L_start=L_start, L_event=L_event)


# --------------------------------------------------------------------------------------
# numpy interpreter (CPU validation of the schedule)
# --------------------------------------------------------------------------------------

def run_ops(ops, env):
    import numpy as np

    def prmt(a, b, sel):
        src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
        out = np.zeros_like(a)
        for lane in range(4):
            nib = (sel >> (4 * lane)) & 0xF
            byte = src[nib & 7]
            if nib & 8:
                byte = np.where(byte & 0x80, 0xFF, 0).astype(a.dtype)
            out |= byte << (8 * lane)
        return out

    u32 = np.uint32
    for op in ops:
        k, d = op[0], op[1]
        if k == "prmt":
            env[d] = prmt(env[op[2]], env[op[3]], op[4])
        elif k in ("add", "addf", "addg"):
            env[d] = (env[op[2]] + env[op[3]]).astype(u32)
        elif k == "add3c":
            env[d] = (env[op[2]] + env[op[3]] + u32(op[4])).astype(u32)
        elif k == "cmp":
            env[d] = (env[op[2]] + u32(H) - env[op[3]]).astype(u32)
        elif k == "cmpx":
            env[d] = (env[op[2]] + u32(H) - env[op[5]]).astype(u32)
            assert np.array_equal(env[d], (env[op[3]] + env[op[4]] - env[op[5]]).astype(u32))
        elif k == "path2":
            env[d] = (env[op[2]] + env[op[2]] + u32(0x01010101)).astype(u32)
        elif k == "addc":
            env[d] = (env[op[2]] + u32(op[3])).astype(u32)
        elif k == "subm":
            env[d] = (env[op[2]] - env[op[3]]).astype(u32)
        elif k == "mad2c":
            env[d] = (env[op[2]] * u32(2) + u32(op[3])).astype(u32)
        elif k == "signmask":
            env[d] = prmt(env[op[2]], env["ZERO"], 0xBA98)
        elif k == "sel":
            m = env[op[4]]
            env[d] = ((env[op[3]] & m) | (env[op[2]] & ~m)).astype(u32)
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

    def ref(x):
        return x

    def dst(x):
        if "[" in x or x in declared:
            return x
        declared.add(x)
        return "uint32_t " + x

    for op in ops:
        k, d = op[0], op[1]
        if k == "prmt":
            b = "0u" if op[3] == "ZERO" else ref(op[3])
            lines.append("%s = vit_prmt(%s, %s, 0x%04xu);" % (dst(d), ref(op[2]), b, op[4]))
        elif k == "add":
            lines.append("%s = %s + %s;" % (dst(d), ref(op[2]), ref(op[3])))
        elif k == "addf":
            lines.append("%s = VIT_ADDF(%s, %s);" % (dst(d), ref(op[2]), ref(op[3])))
        elif k == "addg":
            lines.append("%s = VIT_ADDG(%s, %s);" % (dst(d), ref(op[2]), ref(op[3])))
        elif k == "add3c":
            lines.append("%s = %s + %s + 0x%08xu;" % (dst(d), ref(op[2]), ref(op[3]), op[4]))
        elif k == "cmp":
            lines.append("%s = %s + 0x80808080u - %s;" % (dst(d), ref(op[2]), ref(op[3])))
        elif k == "cmpx":
            lines.append("%s = VIT_CMP(%s, %s, %s, %s);" % (dst(d), ref(op[2]), ref(op[3]), ref(op[4]), ref(op[5])))
        elif k == "path2":
            lines.append("%s = VIT_PATH2(%s);" % (dst(d), ref(op[2])))
        elif k == "addc":
            lines.append("%s = %s + 0x%08xu;" % (dst(d), ref(op[2]), op[3]))
        elif k == "subm":
            lines.append("%s = vit_mad(%s, vit_neg1, %s);" % (dst(d), ref(op[3]), ref(op[2])))
        elif k == "mad2c":
            lines.append("%s = vit_mad(%s, vit_two, 0x%08xu);" % (dst(d), ref(op[2]), op[3]))
        elif k == "signmask":
            lines.append("%s = vit_prmt(%s, 0u, 0xba98u);" % (dst(d), ref(op[2])))
        elif k == "sel":
            lines.append("%s = vit_sel(%s, %s, %s);" % (dst(d), ref(op[2]), ref(op[3]), ref(op[4])))
        elif k == "mov":
            lines.append("%s = %s;" % (dst(d), ref(op[2])))
    return "\n".join(indent + l for l in lines)


HEADER = """// GENERATED by gr_dvbt_b200/csrc/gen_viterbi_acs.py -- do not edit by hand.
// Register-resident 64-state ACS schedule (4 states per 32-bit register, unsigned-byte SWAR)
// for one byte time (8 trellis steps) of the K=7 (0x4f,0x6d) code.  See the generator's
// docstring and DESIGN.md (K1) for the layout algebra; reference semantics:
// /root/reference/lib/d_viterbi.c:461-576 (d_viterbi_butterfly2_sse2).
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t vit_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}
// a * b + c on the FMA pipe (IMAD).  b comes from a kernel parameter (vit_neg1 = 0xffffffff, vit_two = 2)
// so that ptxas cannot fold it back into an ALU-pipe IADD3: the ALU pipe (LOP3/PRMT) is the bottleneck
// of this kernel and the FMA pipe has twice its width (profiles/r01_viterbi_v1_ncu_summary.txt).
__device__ __forceinline__ uint32_t vit_mad(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
// The branch-metric additions of the schedule.  The including kernel may define VIT_ADDF / VIT_ADDG as
// vit_mad(a, vit_one, b) to issue half or all of them on the FMA pipe; the default is a plain addition.
#ifndef VIT_ADDF
#define VIT_ADDF(a, b) ((a) + (b))
#endif
#ifndef VIT_ADDG
#define VIT_ADDG(a, b) ((a) + (b))
#endif
// Compare word t = m1 + 0x80808080 - m0, where m1 = mj + b and bh = b + 0x80808080: bit 7 of a byte = (m1 >= m0).
// Default: one 3-input addition.  FMA form: vit_mad(m0, vit_neg1, mj + bh).
#ifndef VIT_CMP
#define VIT_CMP(m1, mj, bh, m0) ((m1) + 0x80808080u - (m0))
#endif
// Path of the odd predecessor: 2 p + 1 per byte.  FMA form: vit_mad(p, vit_two, 0x01010101).
#ifndef VIT_PATH2
#define VIT_PATH2(p) ((p) + (p) + 0x01010101u)
#endif
// per byte: mask ? y : x   (mask bytes are 0x00 or 0xff)
__device__ __forceinline__ uint32_t vit_sel(uint32_t x, uint32_t y, uint32_t mask) {
  return (y & mask) | (x & ~mask);
}

"""


def state_index_tables(res):
    """ring-row byte index of state s at the event: word w (canonical event order) * 4 + lane"""
    tbl = [0] * 64
    for w, st in enumerate(res["L_event"]):
        for b, s in enumerate(st):
            tbl[s] = 4 * w + b
    return tbl


def main():
    res = build()
    out = [HEADER]
    out.append("// Steps 1..6 of a byte time.  In: M[16], P[16] in start layout (lane bits = state bits 2,3).\n"
               "// Out: M_ev[16], P_ev[16] in event layout: word w = (s5,s2,s1,s0), lane = (s4,s3).\n"
               "// (expects uint32_t vit_neg1, vit_two in scope)\n#define VIT_ACS_PART1(M, P, M_ev, P_ev, apk0, apk1, apk2, apk3, apk4, apk5) \\\n")
    body = emit_cuda(res["part1"], indent="  ")
    out.append("  do { \\\n" + "\n".join(l + " \\" for l in body.split("\n")) + "\n  } while (0)\n\n")
    out.append("// Event-layout metrics -> transposed to lane bits (0,1), then steps 7,8.  P must be all zero on entry\n"
               "// (it is re-created here).  Out: M_nx[16], P_nx[16] in start layout.\n"
               "#define VIT_ACS_PART2(M, P, M_nx, P_nx, apk6, apk7) \\\n")
    body = emit_cuda(res["part2"], indent="  ")
    out.append("  do { \\\n" + "\n".join(l + " \\" for l in body.split("\n")) + "\n  } while (0)\n\n")
    tbl = state_index_tables(res)
    # closed form used by the kernel; verified here against the table
    for s in range(64):
        w = ((s >> 5) << 3) | (s & 7)
        lane = (s >> 3) & 3
        assert tbl[s] == 4 * w + lane, (s, tbl[s], w, lane)
    out.append("// ring-row byte index of state s in event layout: 4*((s5<<3)|(s&7)) + ((s>>3)&3)\n"
               "__device__ __forceinline__ uint32_t vit_event_byte_index(uint32_t s) {\n"
               "  return ((s & 7u) << 2) | (s & 32u) | ((s >> 3) & 3u);\n}\n")
    n1 = sum(1 for o in res["part1"] if o[0] != "mov")
    n2 = sum(1 for o in res["part2"] if o[0] != "mov")
    out.append("// op counts: part1 %d, part2 %d (per byte time, before ptxas)\n" % (n1, n2))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "viterbi_acs_gen.cuh")
    with open(path, "w") as f:
        f.write("".join(out))
    print("wrote", path, "ops:", n1, n2)


if __name__ == "__main__":
    main()
