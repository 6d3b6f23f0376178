#!/usr/bin/env python3
"""Static SASS statistics per kernel of libf16b200.so (cuobjdump -sass): instruction count and opcode mix."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "f16_jsb_b200/libf16b200.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "step_kernel"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur:
        ins = re.sub(r"^@!?U?P[0-9T]+\s+", "", m.group(1).strip())
        funcs[cur].append(ins.split()[0].split(".")[0])
for name, ops in funcs.items():
    if pat not in name:
        continue
    c = collections.Counter(ops)
    print("%s: %d instructions" % (name, len(ops)))
    print("   " + "  ".join("%s %d" % kv for kv in c.most_common(28)))
