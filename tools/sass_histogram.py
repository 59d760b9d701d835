"""Instruction mix of the field-product bodies and of the accumulation kernels, read from the built objects with
cuobjdump (no GPU needed).  The multiplier pipe (IMAD*, 32 lanes / clk / SM for IMAD.WIDE) bounds these kernels, so
what counts is how many NON-multiply instructions ptxas placed on it (IMAD.MOV*, IMAD.X, IMAD.IADD).
usage: python tools/sass_histogram.py [object ...]   (default: the in-tree msm_g1.o / msm_g2.o)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
objs = sys.argv[1:] or [os.path.join(ROOT, "interactive_zkp_study_b200", "_build", f) for f in ("msm_g1.o", "msm_g2.o")]
PIPE = ("IMAD.WIDE.U32.X", "IMAD.WIDE.U32", "IMAD.HI.U32", "IMAD", "IMAD.MOV.U32", "IMAD.MOV", "IMAD.X", "IMAD.IADD")
ALU = ("IADD3.X", "IADD3", "MOV", "SEL", "LOP3.LUT", "VIADD")


def functions(path):
    text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, out = None, {}
    for line in text.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            out[cur] = []
            continue
        m = re.match(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P[0-9T]\s+)?([A-Z][A-Z0-9_.]*)(.*)", line)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(2), m.group(3)))
    return out


def show(name, rows):
    c = collections.Counter(op for _, op, _ in rows)
    pipe = " ".join("%s=%d" % (k, c[k]) for k in PIPE if c[k])
    alu = " ".join("%s=%d" % (k, c[k]) for k in ALU if c[k])
    print("  %-34s %5d instr | multiplier pipe: %s | ALU: %s" % (name, len(rows), pipe, alu))


for path in objs:
    print(os.path.basename(path))
    for fn, rows in functions(path).items():
        if "msm_accumulate_kernel" not in fn:
            continue
        # out-of-line bodies are emitted behind the kernel's own code: split at the call targets
        targets = sorted(set(int(re.search(r"0x([0-9a-f]+)", r[2]).group(1), 16) for r in rows if r[1].startswith("CALL.REL")))
        bounds = [0] + targets + [1 << 30]
        label = "G2" if "Fp2" in fn else "G1"
        for i, (lo, hi) in enumerate(zip(bounds[:-1], bounds[1:])):
            part = [r for r in rows if lo <= r[0] < hi and r[1] not in ("NOP", "BRA")]
            show("%s accumulate: %s" % (label, "loop body" if i == 0 else "out-of-line body %d" % i), part)
