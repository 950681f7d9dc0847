"""Saturating classifier measurement alone (2^25 pairs generated on the device): prints ms and GB/s."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from unfazed_b200.engine import Engine  # noqa: E402

r = bench.saturating_classify(Engine(0), 5)
r["gbps"] = r["bytes"] / (r["classify_ms"] * 1e6)
print(json.dumps(r))
