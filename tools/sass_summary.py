#!/usr/bin/env python
"""Per-kernel SASS census of libunfazed_sm100.so (cuobjdump -sass): instruction count and the opcodes that say what
a kernel is made of -- asynchronous copies (LDGSTS = cp.async, UBLKCP = TMA bulk, UTMALDG/UTMASTG = TMA tensor),
barriers (BAR, SYNCS = mbarrier), warp collectives, atomics, tensor-core ops (must be 0: nothing here is a
contraction).      python tools/sass_summary.py > profiles/sass_r2.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "unfazed_b200", "libunfazed_sm100.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WATCH = ["LDG", "STG", "LDS", "STS", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "BAR", "SHFL", "VOTE", "POPC", "ATOMG", "ATOMS",
         "REDG", "REDUX", "MATCH", "DFMA", "MUFU", "HMMA", "IMMA", "UTCHMMA", "UTCIMMA", "LDL", "STL"]
arch = set(re.findall(r"arch = (sm_\w+)", out))
kern, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        kern[cur]["_n"] += 1
        kern[cur][op] += 1
print("# cuobjdump -sass %s" % os.path.basename(lib))
print("# cubin architectures:", ", ".join(sorted(arch)))
print("# %-44s %6s  %s" % ("kernel", "instrs", "watched opcodes (count)"))
for name, c in kern.items():
    short = re.sub(r"^_ZN\d+_GLOBAL__N__\w+?_cu_\w{8}\d+", "", name)
    short = re.sub(r"E(P|N|1|v|i|l).*$", "", short)[:44]
    ops = "  ".join("%s=%d" % (w, sum(v for k, v in c.items() if k == w or k.startswith(w + "."))) for w in WATCH
                    if any(k == w or k.startswith(w + ".") for k in c))
    print("  %-44s %6d  %s" % (short, c["_n"], ops))
tens = sum(v for c in kern.values() for k, v in c.items() if re.match(r"(HMMA|IMMA|DMMA|QMMA|UTC\w*MMA|WGMMA)", k))
print("# tensor-core instructions in the library: %d" % tens)
