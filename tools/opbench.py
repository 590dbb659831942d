#!/usr/bin/env python
"""Per-operator device times (the 'kernels' table of bench.py) without the
whole-job run."""
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import util  # noqa: E402

P = util.pkg()
k, me = bench.kernel_rooflines(P, P.load(), 6547.8)
for r in k:
    print("%-62s %9.4f ms  %8.1f GB/s  frac %.4f  launches %d" % (r["kernel"][:62], r["ms_per_frame"], r["achieved_gbs"],
                                                                 r["frac"], r["launches_per_frame"]))
