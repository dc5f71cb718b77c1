"""Achieved HBM bandwidth of the byte/integer kernels at BASELINE sizes (5000x5000, 169 x 512 x 512): prints the table that
bench.py reports as `roofline_hbm` (snb_b200/hbm_bench.py: CUDA events, L2 flush between iterations, algorithmic bytes)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import snb_b200  # noqa: E402,F401
from snb_b200 import hbm_bench  # noqa: E402

peak = 6549.1
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
only = sys.argv[1:] or None
for name, r in hbm_bench.measure(peak, reps=6, only=only).items():
    print("%-12s %8.1f us  %8.1f MB  %8.1f GB/s  %5.1f%% of measured %.0f GB/s   %s" % (
        name, r["us"], r["bytes"] / 1e6, r["GBps"], 100 * r["frac"], peak, r["what"]), flush=True)
