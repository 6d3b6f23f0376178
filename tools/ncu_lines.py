#!/usr/bin/env python3
"""Join an ncu SASS-level source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm line info
of the same library build: per source line, executed warp-instructions and stall samples.
Usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep [kernel-substring] [top-n]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "step_kernelIfE"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
lib = os.environ.get("F16_B200_LIB", "f16_jsb_b200/libf16b200.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Warp Stall Sampling (All Samples)")}
sass = []
for r in rows[2:]:
    if len(r) <= ci["Instructions Executed"]:
        continue
    try:
        sass.append((r[ci["Source"]].strip(), float(r[ci["Instructions Executed"]]), float(r[ci["# Samples"]])))
    except ValueError:
        pass
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
txt = ""
for cubin in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):      # one cubin per .cu of the library
    t = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    if pat in t:
        txt = t
        break
in_fn, cur, lines = False, None, []
for line in txt.splitlines():
    if line.startswith("//--------------------- .text."):
        in_fn = pat in line
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line):
        lines.append(cur)
if len(lines) != len(sass):
    print("warning: %d SASS rows in the report vs %d in the library (different build?)" % (len(sass), len(lines)))
n = min(len(lines), len(sass))
inst, samp = collections.Counter(), collections.Counter()
for i in range(n):
    inst[lines[i]] += sass[i][1]
    samp[lines[i]] += sass[i][2]
ti, ts = sum(inst.values()), sum(samp.values())
cache = {}
def src(f, ln):
    for base in ("f16_jsb_b200/csrc", "include"):
        p = os.path.join(base, f)
        if os.path.exists(p):
            cache.setdefault(p, open(p).read().splitlines())
            return cache[p][ln - 1].strip()[:80] if ln <= len(cache[p]) else ""
    return ""
print("%d warp-instructions, %d samples" % (ti, ts))
print("--- by stall samples")
for k, v in samp.most_common(topn):
    print("%5.1f%% samples %5.1f%% inst  %s:%d  %s" % (100 * v / ts, 100 * inst[k] / ti, k[0], k[1], src(*k)))
