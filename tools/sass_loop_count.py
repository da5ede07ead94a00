#!/usr/bin/env python3
"""Instruction mix of a kernel's loops, read from the SASS of gr_dvbt_b200/libdvbt_b200.so (no GPU needed).

    python tools/sass_loop_count.py 'vit_acs_kernelILb1ELi0ELi1ELi384' [rank]

Loops are the backward branches of the function, largest span first; `rank` picks one (default: every loop with more
than 100 instructions).  Instructions are attributed to the pipe they issue on for the purpose of the ALU-pipe /
issue-slot budget of DESIGN.md K1: ALU = integer logic, shifts, permutes, compares, the 16x2 DPX min/max; FMA = IMAD*
and FP32; everything else (LSU, branch, uniform datapath) only costs an issue slot."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gr_dvbt_b200", "libdvbt_b200.so")
ALU = {"VIADDMNMX", "PRMT", "LOP3", "VIMNMX3", "VIMNMX", "SHF", "VIADD", "ISETP", "SEL", "IADD3", "LEA", "PLOP3", "BREV",
       "IABS", "FLO", "POPC", "SGXT", "BMSK", "FSETP", "FSEL", "FMNMX", "MOV"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2"}


def functions(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if cur and m:
            out[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return out


def mnemonic(ins):
    return re.sub(r"^@!?U?P\w+\s+", "", ins).split()[0].split(".")[0]


def main():
    pat = sys.argv[1]
    rank = int(sys.argv[2]) if len(sys.argv) > 2 else None
    for name, ins in functions(LIB).items():
        if not re.search(pat, name):
            continue
        loops = []
        for addr, text in ins:
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\w+,\s*)?(0x[0-9a-f]+)", text)
            if m and int(m.group(1), 16) < addr:
                loops.append((addr - int(m.group(1), 16), int(m.group(1), 16), addr))
        loops.sort(reverse=True)
        print("%s: %d instructions, %d loops" % (name, len(ins), len(loops)))
        for r, (span, t, a) in enumerate(loops):
            body = [text for addr, text in ins if t <= addr <= a]
            if (rank is None and len(body) <= 100) or (rank is not None and r != rank):
                continue
            ops = collections.Counter(mnemonic(x) for x in body)
            alu = sum(v for k, v in ops.items() if k in ALU)
            fma = sum(v for k, v in ops.items() if k in FMA)
            print("  loop %d [%#x, %#x]: %d instructions, ALU pipe %d, FMA pipe %d, other %d" % (r, t, a, len(body), alu, fma, len(body) - alu - fma))
            print("    " + ", ".join("%s %d" % kv for kv in ops.most_common(18)))


if __name__ == "__main__":
    main()
