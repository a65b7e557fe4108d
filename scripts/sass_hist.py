"""Static SASS view of a kernel in the built library: total instruction count and an opcode histogram of the rotation
loop of the afterstate kernels (the smallest loop holding the 10 unrolled slots).

    python scripts/sass_hist.py step_observe_kernelILi0 [--dump]
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "lib", "libtetris_piclim_sm100.so")
pat = sys.argv[1]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ins, on = [], False
for line in out.splitlines():
    if "Function :" in line:
        on = pat in line
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*;", line)
    if on and m:
        ins.append((int(m.group(1), 16), m.group(2)))
print(f"{pat}: {len(ins)} instructions")
best = None
for k, (addr, text) in enumerate(ins):
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", text)
    if m and int(m.group(1), 16) < addr:
        t = int(m.group(1), 16)
        body = [x for x in ins if t <= x[0] <= addr]
        if "--loops" in sys.argv:
            print(f"  loop {t:#x}..{addr:#x}: {len(body)} instructions")
        # the rotation loop: the smallest loop that holds the ten unrolled slots (two VABSDIFF4 each)
        if sum("VABSDIFF4" in x[1] for x in body) >= 10 and (best is None or len(body) < len(best)):
            best = body
ALU = ("LOP3", "SEL", "VIADDMNMX", "VABSDIFF", "SHF", "VIADD", "VIMNMX", "LEA", "ISETP", "PRMT", "IADD3", "PLOP3", "IABS", "MOV", "P2R", "R2P", "BMSK", "SGXT")
FMA = ("IMAD", "FFMA", "FMUL", "FADD")
h = collections.Counter()
pipes = collections.Counter()
for _, text in best:
    op = re.sub(r"^@!?U?P\d+\s+", "", text).split()[0]
    base = op.split(".")[0]
    h[op] += 1
    pipes["alu" if base in ALU else "fma" if base in FMA else "other"] += 1
print(f"rotation loop: {len(best)} instructions; alu {pipes['alu']} fma {pipes['fma']} other {pipes['other']}")
for op, c in h.most_common():
    print(f"  {c:4d} {op}")
if "--dump" in sys.argv:
    for a, t in best:
        print(f"/*{a:04x}*/ {t}")
