#!/usr/bin/env python3
"""Generator for the TWO-LANES-PER-CHUNK variant of the K=7 ACS schedule (viterbi_acs2_gen.cuh).

Same arithmetic as gen_viterbi_acs.py (unsigned-byte SWAR, exact reference tie rule), but the 64
states of a chunk are split over a pair of adjacent lanes: each lane holds 8 metric and 8 path
registers.  The survivor ring in shared memory is per chunk, so at rate 7/8 an SM now runs 256
threads (two warps per scheduler) instead of 128, which is what the first kernel lacked
(profiles/r01_viterbi_v1_ncu_summary.txt: issue 51 %, stall reason "wait").

Three of the six state bits are "physical": the lane-pair bit T and the two byte-lane bits A, B; the
other three select the register.  Every trellis step moves them up by one position and none of them
may sit at position 5 when a butterfly runs.  Schedule per byte time (positions of A, B, T):
   start (2,3,4) | s1 (3,4,5) swapT 5->0 | s2 (4,5,1) swapB 5->0 | s3 (5,1,2) swapA 5->0
   | s4 (1,2,3) | s5 (2,3,4) | s6 (3,4,5) | EVENT (lane t holds states with s5 = t; word = s&7,
   byte lane = (s4,s3): the ring row format of the one-lane kernel) ; metrics only: A 3->0, B 4->1,
   T 5->2 | s7 (1,2,3) | s8 (2,3,4).
A byte-lane swap is one PRMT per register; a lane-pair swap is select + shfl.xor(1) + 2 selects per
register pair.  The branch labels of lane 1 differ from lane 0's by a flip of label bits that depends
only on T's position, i.e. a fixed byte permutation of APK per step (one PRMT with a per-lane selector).
"""
import os

import gen_viterbi_acs as G1

H = G1.H
label = G1.label


def tmask(pT):
    """label-index flip seen by lane 1 when the pair bit sits at state position pT of the butterfly's low state"""
    return 2 * (pT in (0, 1, 2)) + (pT in (1, 2, 4))


class Gen2(G1.Gen):
    def butterfly_step2(self, M, P, apk, pT):
        """M, P: lists of (name, lane-0 states).  apk is permuted for lane 1 first."""
        mask = tmask(pT)
        if mask:
            a2 = self.new("ak")
            self.emit("prmtr", a2, apk, "ZERO", "tsel%d" % mask)  # selector register: lane 0 identity, lane 1 the flip
            apk = a2
        return self.butterfly_step(M, P, apk)

    def swap_thread(self, W, word_pos, pT):
        """exchange the pair bit (at state position pT) with the word bit at position word_pos"""
        idx = {w[1]: w[0] for w in W}
        out, done = [], set()
        for name, st in sorted(W, key=lambda w: w[1]):
            if st in done or (st[0] >> word_pos) & 1:
                continue
            st1 = tuple(s | (1 << word_pos) for s in st)
            w0, w1 = name, idx[st1]
            done.add(st)
            done.add(st1)
            send, recv, lo, hi = self.new("sd"), self.new("rv"), self.new("y"), self.new("y")
            self.emit("tsel", send, w0, w1)      # T ? w0 : w1
            self.emit("shfl", recv, send)
            self.emit("tsel", lo, recv, w0)      # T ? recv : w0
            self.emit("tsel", hi, w1, recv)      # T ? w1 : recv
            # lane-0 view: lo = own w0 (pT bit 0), hi = partner's w0 (pT bit 1); the pair bit is now word_pos
            out += [(lo, st), (hi, tuple(s | (1 << pT) for s in st))]
        assert len(out) == len(W)
        return out


def layout8(pa, pb, pt):
    """lane-0 words (pair bit = 0): 8 tuples ordered by the remaining 3 bits"""
    rest = [p for p in range(6) if p not in (pa, pb, pt)]
    words = []
    for w in range(8):
        base = sum(((w >> i) & 1) << rest[i] for i in range(3))
        words.append(tuple(base | ((b & 1) << pa) | ((b >> 1) << pb) for b in range(4)))
    return words


def check8(W, pa, pb, pt):
    assert sorted(w[1] for w in W) == sorted(layout8(pa, pb, pt)), (pa, pb, pt)


def build():
    L_start = layout8(2, 3, 4)
    L_event = layout8(3, 4, 5)
    g = Gen2()
    M = [("M[%d]" % i, st) for i, st in enumerate(L_start)]
    P = [("P[%d]" % i, st) for i, st in enumerate(L_start)]
    M, P = g.butterfly_step2(M, P, "apk0", 4); check8(M, 3, 4, 5)
    M = g.swap_thread(M, 0, 5); P = g.swap_thread(P, 0, 5); check8(M, 3, 4, 0)
    M, P = g.butterfly_step2(M, P, "apk1", 0); check8(M, 4, 5, 1)
    M = g.swap(M, 1, 0, 5); P = g.swap(P, 1, 0, 5); check8(M, 4, 0, 1)
    M, P = g.butterfly_step2(M, P, "apk2", 1); check8(M, 5, 1, 2)
    M = g.swap(M, 0, 0, 5); P = g.swap(P, 0, 0, 5); check8(M, 0, 1, 2)
    M, P = g.butterfly_step2(M, P, "apk3", 2); check8(M, 1, 2, 3)
    M, P = g.butterfly_step2(M, P, "apk4", 3); check8(M, 2, 3, 4)
    M, P = g.butterfly_step2(M, P, "apk5", 4); check8(M, 3, 4, 5)
    for arr, W in (("M", M), ("P", P)):
        idx = {w[1]: w[0] for w in W}
        for i, st in enumerate(L_event):
            g.emit("mov", "%s_ev[%d]" % (arr, i), idx[st])
    part1 = g.ops

    g = Gen2()
    M = [("M[%d]" % i, st) for i, st in enumerate(L_event)]
    M = g.swap(M, 0, 0, 3); check8(M, 0, 4, 5)
    M = g.swap(M, 1, 1, 4); check8(M, 0, 1, 5)
    M = g.swap_thread(M, 2, 5); check8(M, 0, 1, 2)
    P = [("P[%d]" % i, st) for i, st in enumerate(layout8(0, 1, 2))]
    M, P = g.butterfly_step2(M, P, "apk6", 2); check8(M, 1, 2, 3)
    M, P = g.butterfly_step2(M, P, "apk7", 3); check8(M, 2, 3, 4)
    for arr, W in (("M", M), ("P", P)):
        idx = {w[1]: w[0] for w in W}
        for i, st in enumerate(L_start):
            g.emit("mov", "%s_nx[%d]" % (arr, i), idx[st])
    part2 = g.ops
    return dict(part1=part1, part2=part2, L_start=L_start, L_event=L_event)


TSEL = {1: 0x2301, 2: 0x1032, 3: 0x0123}


def run_ops(ops, env):
    """numpy interpreter; every value is an array [..., 2] whose last axis is the lane of the pair"""
    import numpy as np
    u32 = np.uint32
    T = np.array([0, 1], dtype=bool)

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

    for op in ops:
        k, d = op[0], op[1]
        if k == "prmt":
            env[d] = prmt(env[op[2]], env[op[3]], op[4])
        elif k == "prmtr":
            m = int(op[4][4:])
            a = env[op[2]]
            env[d] = np.where(T, prmt(a, env[op[3]], TSEL[m]), a).astype(u32)
        elif k == "add":
            env[d] = (env[op[2]] + env[op[3]]).astype(u32)
        elif k == "add3c":
            env[d] = (env[op[2]] + env[op[3]] + u32(op[4])).astype(u32)
        elif k == "cmp":
            env[d] = (env[op[2]] + u32(H) - env[op[3]]).astype(u32)
        elif k == "signmask":
            env[d] = prmt(env[op[2]], env["ZERO"], 0xBA98)
        elif k == "sel":
            m = env[op[4]]
            env[d] = ((env[op[3]] & m) | (env[op[2]] & ~m)).astype(u32)
        elif k == "tsel":
            env[d] = np.where(T, env[op[2]], env[op[3]]).astype(u32)
        elif k == "shfl":
            env[d] = env[op[2]][..., ::-1].copy()
        elif k == "mov":
            env[d] = env[op[2]]
        else:
            raise ValueError(k)
    return env


def emit_cuda(ops, indent="  "):
    lines, declared = [], set()

    def dst(x):
        if "[" in x or x in declared:
            return x
        declared.add(x)
        return "uint32_t " + x

    for op in ops:
        k, d = op[0], op[1]
        if k == "prmt":
            b = "0u" if op[3] == "ZERO" else op[3]
            lines.append("%s = vit_prmt(%s, %s, 0x%04xu);" % (dst(d), op[2], b, op[4]))
        elif k == "prmtr":
            lines.append("%s = vit_prmt(%s, 0u, vit_%s);" % (dst(d), op[2], op[4]))
        elif k == "add":
            lines.append("%s = %s + %s;" % (dst(d), op[2], op[3]))
        elif k == "add3c":
            lines.append("%s = %s + %s + 0x%08xu;" % (dst(d), op[2], op[3], op[4]))
        elif k == "cmp":
            lines.append("%s = %s + 0x80808080u - %s;" % (dst(d), op[2], op[3]))
        elif k == "signmask":
            lines.append("%s = vit_prmt(%s, 0u, 0xba98u);" % (dst(d), op[2]))
        elif k == "sel":
            lines.append("%s = vit_sel(%s, %s, %s);" % (dst(d), op[2], op[3], op[4]))
        elif k == "tsel":
            lines.append("%s = vit_t ? %s : %s;" % (dst(d), op[2], op[3]))
        elif k == "shfl":
            lines.append("%s = __shfl_xor_sync(vit_pairmask, %s, 1);" % (dst(d), op[2]))
        elif k == "mov":
            lines.append("%s = %s;" % (dst(d), op[2]))
    return "\n".join(indent + l for l in lines)


HEADER = """// GENERATED by gr_dvbt_b200/csrc/gen_viterbi_acs2.py -- do not edit by hand.
// Two-lanes-per-chunk ACS schedule: lane t of a pair holds 32 of the 64 states (8 metric + 8 path
// registers, 4 states per register).  Expects in scope: bool vit_t (lane parity), unsigned vit_pairmask
// (the two lanes of the pair) and
// uint32_t vit_tsel1, vit_tsel2, vit_tsel3 (= vit_t ? 0x2301/0x1032/0x0123 : 0x3210).
// Reference semantics: /root/reference/lib/d_viterbi.c:461-576.
#pragma once
#include "viterbi_acs_gen.cuh"

"""


def main():
    res = build()
    out = [HEADER]
    out.append("#define VIT2_ACS_PART1(M, P, M_ev, P_ev, apk0, apk1, apk2, apk3, apk4, apk5) \\\n  do { \\\n")
    out.append("\n".join(l + " \\" for l in emit_cuda(res["part1"]).split("\n")) + "\n  } while (0)\n\n")
    out.append("#define VIT2_ACS_PART2(M, P, M_nx, P_nx, apk6, apk7) \\\n  do { \\\n")
    out.append("\n".join(l + " \\" for l in emit_cuda(res["part2"]).split("\n")) + "\n  } while (0)\n\n")
    # event layout check: lane-0 word w holds states s5=0, (s&7)=w, byte lane (s>>3)&3
    for w, st in enumerate(res["L_event"]):
        for b, s in enumerate(st):
            assert s < 32 and (s & 7) == w and ((s >> 3) & 3) == b
    n1 = sum(1 for o in res["part1"] if o[0] != "mov")
    n2 = sum(1 for o in res["part2"] if o[0] != "mov")
    out.append("// op counts per lane: part1 %d, part2 %d\n" % (n1, n2))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "viterbi_acs2_gen.cuh")
    with open(path, "w") as f:
        f.write("".join(out))
    print("wrote", path, "ops per lane:", n1, n2)


if __name__ == "__main__":
    main()
