#!/usr/bin/env python
"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv) and/or a --set full report into
the short text summaries kept under profiles/.

    python tools/ncu_summary.py launches gpurun_out/launches_r1.csv > profiles/r1_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof_r1.ncu-rep > profiles/r1_kernels.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        agg[row["Kernel Name"].split("(")[0]].append(float(row["Metric Value"]))
    tot = sum(sum(v) for v in agg.values())
    print("# ncu launch list: per-kernel device time (cold-cache, serialised: compare SHARES)")
    print("%-72s %5s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-72s %5d %12.1f %10.1f %6.1f%%" % (k[-72:], len(v), sum(v) / 1e3, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full, one block per captured launch")
    for row in rows[2:]:
        print("\n== %s  grid=%s block=%s" % (row[hdr.index("Kernel Name")][:110], row[hdr.index("launch__grid_size")], row[hdr.index("launch__block_size")]))
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("   %-66s %16s %s" % (m, row[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
