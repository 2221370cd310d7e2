"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass), written to profiles/sass_opcodes.txt.
The mnemonics that prove the sm_100a paths: UBLKCP (cp.async.bulk), UTMALDG (TMA tensor copies), UTCHMMA (tcgen05.mma),
LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), SYNCS (mbarrier), DFMA / FFMA2 (the arithmetic), RED / ATOM (scatter).
usage: python tools/sass_histogram.py [lib.so] [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "spark-agd_b200", "libagd_b200.so")
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "sass_opcodes.txt")
KEY = ["UBLKCP", "UTMALDG", "UTCHMMA", "UTCBAR", "LDTM", "SYNCS", "DFMA", "DADD", "DMUL", "FFMA2", "FFMA", "F2F", "LDS", "STS",
       "LDG", "STG", "RED", "ATOM", "ATOMG", "SHFL", "FSEL", "BAR", "MUFU"]

txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
kernels, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        kernels[cur][m.group(1)] += 1
total = collections.Counter()
lines = [f"# SASS opcode histogram of {os.path.relpath(lib, ROOT)} (static instruction counts, all code paths; cuobjdump -sass)",
         f"# {len(kernels)} kernels; columns: " + " ".join(KEY) + " | total"]
for name, c in kernels.items():
    total.update(c)
    short = re.sub(r"agd::\(anonymous namespace\)::", "", demangle(name))
    short = re.sub(r"\(agd::.*", "", short)[:110]
    lines.append(f"{short:112s} " + " ".join(f"{k}={c[k]}" for k in KEY if c[k]) + f" | {sum(c.values())}")
lines.append("# whole library: " + " ".join(f"{k}={total[k]}" for k in KEY if total[k]) + f" | {sum(total.values())}")
os.makedirs(os.path.dirname(out), exist_ok=True)
open(out, "w").write("\n".join(lines) + "\n")
print(lines[-1])
