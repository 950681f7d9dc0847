"""Debug helper (build with EXTRA=-DCH_DEBUG): cycles per phase of the chaining kernel (thread 0 of every CTA)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench.py", "--dnms", "4000", "--steps", "5", "--warmup", "3"]
import bench  # noqa: E402

bench.main()
from unfazed_b200 import _lib  # noqa: E402

lib = ctypes.CDLL(_lib.LIB_PATH)
out = (ctypes.c_ulonglong * 16)()
lib.unfz_debug_chain(out)
names = ["init", "seeds(1)", "2ab ranges", "2c candidates", "2d offsets/Q18", "3 seed reg", "3.5 alleles", "4 BFS", "5 evidence", "6 tally"]
tot = sum(out[:10])
for n, v in zip(names, out):
    print("%-16s %6.1f%%" % (n, 100.0 * v / max(tot, 1)))
