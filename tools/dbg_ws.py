"""Debug helper (build with EXTRA=-DWS_DEBUG): per-phase cycle counters of the warp-specialised read scan."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = ["bench.py", "--dnms", "4000", "--steps", "5", "--warmup", "3"]
import bench  # noqa: E402

bench.main()
from unfazed_b200 import _lib  # noqa: E402

lib = ctypes.CDLL(_lib.LIB_PATH)
out = (ctypes.c_ulonglong * 8)()
lib.unfz_debug_ws(out)
v = list(out)
tiles = max(v[7], 1)
print("tiles", tiles)
print("consumer warp0: full-wait %.0f cyc/tile, compute %.0f cyc/tile" % (v[0] / tiles, v[1] / tiles))
print("fill latency when waited: %.0f cyc (n=%d of %d)" % (v[2] / max(v[3], 1), v[3], tiles))
print("producer: prep %.0f cyc/tile, empty-wait %.0f cyc/tile" % (v[5] / tiles, v[4] / tiles))
