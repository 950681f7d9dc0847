"""Per-source-line stall samples and instruction counts of one kernel from an ncu report
(needs -lineinfo and --import-source on; joins `ncu --page source --print-source sass` with
`nvdisasm -g` line markers of the same cubin).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTRING path/to/lib.so SOURCE.cu [top]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, kern, lib, srcfile = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    base = os.path.basename(srcfile).split(".")[0]
    cubin = [f for f in os.listdir(tmp) if f.startswith(base + ".")][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = end = None
    for i, l in enumerate(dis):
        if l.startswith(".text.") and kern in l:
            start = i
        elif start is not None and end is None and l.startswith(".text.") and i > start:
            end = i
    insts, cur = [], None
    for l in dis[start:end]:
        m = re.search(r'//## File ".*?/([^/"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
        elif re.match(r"\s+/\*[0-9a-f]{4}\*/", l):
            insts.append((cur, l.strip()))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name",
                          "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    h = [i for i, r in enumerate(rows) if "Source" in r and any("Sampling" in c for c in r)][0]
    idx = {c: i for i, c in enumerate(rows[h])}
    data = []
    for r in rows[h + 1:]:
        try:
            data.append((int(r[idx["Warp Stall Sampling (All Samples)"]] or 0), int(r[idx["Instructions Executed"]] or 0)))
        except Exception:
            pass
        if len(data) == len(insts):
            break
    by = {}
    for (loc, _), (st, ie) in zip(insts, data):
        a = by.setdefault(loc, [0, 0])
        a[0] += st
        a[1] += ie
    ts, ti = sum(v[0] for v in by.values()), sum(v[1] for v in by.values())
    src = open(srcfile).read().split("\n")
    name = os.path.basename(srcfile)
    print("# %s: %d SASS instructions, %d stall samples, %d warp-instructions executed" % (kern, len(insts), ts, ti))
    for title, key in (("stall samples", 0), ("instructions executed", 1)):
        print("## by %s" % title)
        for loc, v in sorted(by.items(), key=lambda x: -x[1][key])[:top]:
            t = src[loc[1] - 1].strip()[:80] if loc and loc[0] == name else str(loc)
            print("%5.1f%% stall %5.1f%% inst  L%-4s %s" % (100 * v[0] / max(ts, 1), 100 * v[1] / max(ti, 1), loc[1] if loc else "?", t))


if __name__ == "__main__":
    main()
