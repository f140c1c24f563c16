#!/usr/bin/env python
"""Static SASS opcode histogram per kernel of libpyfe3d_b200.so (cuobjdump -sass), written to profiles/<tag>_sass_mix.txt.
Shows what the sm_100a build consists of: FP64 pipe (DFMA/DADD/DMUL), TMA bulk copies (UBLKCP), async copies (LDGSTS),
L2 prefetch (UBLKPF), no tensor-core instructions (the path is FP64 3x3 / 6x6 arithmetic: north_star).
usage: python scripts/sass_static_mix.py <tag>"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "pyfe3d_b200", "lib", "libpyfe3d_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
parts = re.split(r"\n\s*Function : ", txt)
rows = []
for p in parts[1:]:
    name = p.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"pf3::\(anonymous namespace\)::", "", dem)
    short = re.sub(r"\(.*", "", short)[:70]
    c = collections.Counter()
    n = 0
    for l in p.split("\n"):
        m = re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            c[m.group(1)] += 1
            n += 1
    rows.append((short, n, c))
rows.sort(key=lambda r: -r[1])
keys = ["DFMA", "DADD", "DMUL", "MUFU", "LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKCP", "UBLKPF", "SHFL", "MATCH", "ATOM",
        "ATOMG", "RED", "HMMA", "UTCHMMA", "UTCMMA", "BAR"]
out = os.path.join(ROOT, "profiles", "%s_sass_mix.txt" % sys.argv[1])
with open(out, "w") as f:
    f.write("# static SASS opcode counts per kernel, %s, cubins: %s\n" % (os.path.basename(lib), ", ".join(arch)))
    f.write("%-72s %6s " % ("kernel", "instrs") + " ".join("%6s" % k for k in keys) + "\n")
    tot = collections.Counter()
    for short, n, c in rows:
        f.write("%-72s %6d " % (short, n) + " ".join("%6d" % c.get(k, 0) for k in keys) + "\n")
        tot.update(c)
    f.write("%-72s %6d " % ("TOTAL (%d kernels)" % len(rows), sum(r[1] for r in rows)) + " ".join("%6d" % tot.get(k, 0) for k in keys) + "\n")
    f.write("# top opcodes overall: " + ", ".join("%s %d" % kv for kv in tot.most_common(25)) + "\n")
print(open(out).read()[:6000])
