#!/usr/bin/env python
"""Per-source-line stall samples / executed instructions of one kernel from an ncu report captured with
--import-source on.  ncu's csv source page is SASS-only, so the SASS stream is aligned (instruction by instruction) with
`nvdisasm --print-line-info` of the same kernel in the in-tree library.
usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel name substring in ncu> <mangled substring in the cubin> [cu file] [top]
env NCU_LINES_CUBIN=<file.cubin>: align against this cubin instead of the in-tree library.
env NCU_LINES_SHARED=1: rank the lines by shared-memory wavefronts (actual / ideal / excess = bank conflicts) instead."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kname, mangled = sys.argv[1:4]
cu = sys.argv[4] if len(sys.argv) > 4 else None
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
blk = [b for b in blocks if kname in b["name"]][0]
hdr, data = blk["rows"][0], blk["rows"][1:]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")

tmp = tempfile.mkdtemp()
if os.environ.get("NCU_LINES_CUBIN"):     # a cubin of the profiled version of the source (when the tree has moved on)
    import shutil
    shutil.copy(os.environ["NCU_LINES_CUBIN"], tmp)
else:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "pyfe3d_b200", "lib", "libpyfe3d_b200.so")], cwd=tmp,
                   capture_output=True)
seq = None
for f in sorted(os.listdir(tmp)):
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    lines = txt.splitlines()
    starts = [i for i, ln in enumerate(lines) if ln.startswith(".text.") and mangled in ln]
    if not starts:
        continue
    i0 = starts[0]
    i1 = next((i for i in range(i0 + 1, len(lines)) if lines[i].startswith("//-------")), len(lines))
    seq, where = [], None
    for ln in lines[i0:i1]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            where = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            seq.append((where, m.group(2).strip()))
    break
assert seq is not None, "kernel not found in the library"


def opcode(t):
    p = t.split()
    return (p[1] if p[0].startswith("@") else p[0]).split(".")[0]


mism = sum(opcode(t) != opcode(r[ia].strip()) for (_, t), r in zip(seq, data))
print("# %s: %d SASS instructions in the report, %d in the library, %d opcode mismatches" % (blk["name"][:60], len(data), len(seq), mism))
if os.environ.get("NCU_LINES_SHARED"):
    iw, ii, ix = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Ideal"), hdr.index("L1 Wavefronts Shared Excessive")
    agg = defaultdict(lambda: [0, 0, 0, 0])
    for (w, t), r in zip(seq, data):
        if int(r[iw] or 0):
            a = agg[(w, t.split()[1] if t.startswith("@") else t.split()[0])]
            a[0] += int(r[iw] or 0); a[1] += int(r[ii] or 0); a[2] += int(r[ix] or 0); a[3] += int(r[iex])
    tot = sum(a[0] for a in agg.values()) or 1
    print("# shared-memory wavefronts: %d, of which %d excess (bank conflicts)" % (tot, sum(a[2] for a in agg.values())))
    for (w, op), a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print("%-26s %-12s wavefronts %10d %5.1f%%  ideal %10d  excess %10d  executed %9d" % ("%s:%d" % w if w else "?", op, a[0], 100 * a[0] / tot, a[1], a[2], a[3]))
    sys.exit(0)
bys, bye = defaultdict(int), defaultdict(int)
for (w, _), r in zip(seq, data):
    bys[w] += int(r[isamp])
    bye[w] += int(r[iex])
tot, tote = sum(bys.values()) or 1, sum(bye.values()) or 1
src = {}
for w, v in sorted(bys.items(), key=lambda x: -x[1])[:top]:
    text = ""
    if w:
        path = os.path.join(ROOT, "pyfe3d_b200", "csrc", w[0])
        if os.path.exists(path):
            src.setdefault(w[0], open(path).read().splitlines())
            text = src[w[0]][w[1] - 1].strip()[:90]
    print("%-26s samples %6d %5.1f%%  executed %5.1f%%  %s" % ("%s:%d" % w if w else "?", v, 100 * v / tot, 100 * bye[w] / tote, text))
