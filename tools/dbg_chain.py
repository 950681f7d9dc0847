"""Debug helper (library built with -DCH_DEBUG, selected through UNFZ_LIB): cycles per phase of the chaining
kernel (thread 0 of every CTA) and the kernel's time, on a 4000-DNM trio of the headline shape."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unfazed_b200 import _lib  # noqa: E402
from unfazed_b200.engine import Engine, make_params  # noqa: E402
from unfazed_b200.phaser import BatchPhaser  # noqa: E402
from unfazed_b200.synth import SynthConfig, make_dataset  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
sd = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
cov = float(sys.argv[3]) if len(sys.argv) > 3 else 30.0
import pickle  # noqa: E402
cache = "/tmp/dbg_chain_%d_%d_%g.pkl" % (n, sd, cov)     # several library variants are timed on the same data set
if os.path.exists(cache):
    ds = pickle.load(open(cache, "rb"))
else:
    ds = make_dataset(SynthConfig(dnms_per_trio=n, seed=1, search_dist=sd, coverage=cov))
    try:
        pickle.dump(ds, open(cache, "wb"), protocol=5)
    except Exception:
        pass
eng = Engine(0)
bp = BatchPhaser(eng, ds.sites, ds.reads, ds.pedigrees)
params = make_params(readlen=151)
cul = bp.cul(151, 1000000, 3)
plan, layout = bp.plan_batch(ds.dnms, [], threads=1, multiread_proc_min=10 ** 9, search_dist=sd)
for _ in range(3):
    eng.run(bp.dsites, bp.dreads, plan, params, blk_cul=cul, keep_device=False)
lib = ctypes.CDLL(_lib.LIB_PATH)
out = (ctypes.c_ulonglong * 16)()
if hasattr(lib, "unfz_debug_chain"):
    lib.unfz_debug_chain(out)          # reset
ms = {}
for _ in range(5):
    r = eng.run(bp.dsites, bp.dreads, plan, params, blk_cul=cul, time_stages=True, keep_device=False)
    for k, v in r.timings_ms.items():
        ms[k] = ms.get(k, 0.0) + v / 5
print("lib", os.path.basename(_lib.LIB_PATH), "dnms", n, "sd", sd, {k: round(v, 4) for k, v in ms.items() if k in ("read_scan", "read_site_alleles", "chain_tally", "chain_size")})
if hasattr(lib, "unfz_debug_chain"):
    lib.unfz_debug_chain(out)
    names = (["init", "seeds(1)", "2ab ranges", "2c candidates", "2d offsets/ids", "3 seed passes", "3.5 alleles+adj", "4 BFS", "5 evidence", "6 tally"]
             if os.environ.get("CH_OLD") is None else
             ["init", "seeds(1)", "2ab ranges", "2c candidates", "2d offsets/Q18", "3 seed reg", "3.5 alleles", "4 BFS", "5 evidence", "6 tally"])
    tot = sum(out[:10])
    print("  ".join("%s %.1f%%" % (n_, 100.0 * v / max(tot, 1)) for n_, v in zip(names, out)))
