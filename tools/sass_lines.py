#!/usr/bin/env python3
"""Attribute the static SASS of one kernel to source lines (nvdisasm -g on the extracted cubin).
Usage: python tools/sass_lines.py [kernel-substring] [top-n]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

lib = os.environ.get("F16_B200_LIB", "f16_jsb_b200/libf16b200.so")
pat = sys.argv[1] if len(sys.argv) > 1 else "step_kernelIfE"
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
txt = ""
for cubin in sorted(f for f in os.listdir(d) if f.endswith(".cubin")):   # one cubin per .cu file of the library
    t = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
    if pat in t:
        txt = t
        break
in_fn, cur, counts, ops_by_line = False, None, collections.Counter(), collections.defaultdict(collections.Counter)
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
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur:
        ins = re.sub(r"^@!?U?P[0-9T]+\s+", "", m.group(1).strip()).split()[0].split(".")[0]
        counts[cur] += 1
        ops_by_line[cur][ins] += 1
total = sum(counts.values())
print("kernel %s: %d instructions with line info" % (pat, total))
src_cache = {}
def src(f, n):
    for base in ("f16_jsb_b200/csrc", "include"):
        p = os.path.join(base, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[n - 1].strip()[:90] if n <= len(L) else ""
    return ""
for (f, n), c in counts.most_common(topn):
    print("%5d %5.1f%%  %s:%d  %s   [%s]" % (c, 100.0 * c / total, f, n, src(f, n), " ".join("%s%d" % kv for kv in ops_by_line[(f, n)].most_common(4))))
byfile = collections.Counter()
for (f, n), c in counts.items():
    byfile[f] += c
print(dict(byfile))
