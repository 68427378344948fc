"""Opcode histogram per kernel of the built library (cuobjdump -sass), written to profiles/sass_summary.txt:
the evidence that the hot kernels use what DESIGN.md says (DMMA = FP64 tensor cores, UBLKCP = TMA bulk copies,
SYNCS = mbarrier, LDGSTS = cp.async, no library kernels).  Runs without a GPU.
    python scripts/sass_summary.py [out]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "saddlepoint_b200", "lib", "libsaddlepoint_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_summary.txt")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern = None
hist = collections.OrderedDict()
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        op = m.group(1)
        hist[kern][op.split(".")[0]] += 1
        if op.startswith(("DMMA", "UBLKCP", "SYNCS", "LDGSTS", "UTMA", "UTC", "RED", "ATOM", "MEMBAR", "STL", "LDL")):
            hist[kern]["*" + op] += 1
interesting = ("DMMA", "UBLKCP", "SYNCS", "LDGSTS", "UTMA", "UTC", "RED", "ATOM", "MEMBAR", "STL", "LDL", "DFMA", "DADD", "DMUL", "LDG", "STG", "LDS", "STS",
               "SHFL", "BAR")
with open(out, "w") as f:
    f.write("# cuobjdump -sass of saddlepoint_b200/lib/libsaddlepoint_b200.so (sm_100a): instructions per kernel\n")
    f.write("# columns: total, then the opcodes that carry the design (family counts; '*' = exact mnemonic)\n")
    for k, c in hist.items():
        total = sum(v for o, v in c.items() if not o.startswith("*"))
        fam = " ".join("%s=%d" % (o, c[o]) for o in interesting if c.get(o))
        exact = " ".join("%s=%d" % (o[1:], v) for o, v in sorted(c.items()) if o.startswith("*"))
        f.write("%s\n    total=%d %s\n    %s\n" % (k, total, fam, exact))
print(out, len(hist), "kernels")
