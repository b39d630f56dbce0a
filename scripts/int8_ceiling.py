#!/usr/bin/env python
"""Measured int8 tensor ceiling of the box (cuBLASLt IGEMM through torch._int_mm): burst (best of 10) and sustained (>= 2 s
back to back), with the SM clock sampled during the sustained loop.  MEASURED_PEAKS.json carries no int8 figure; bench.py runs
the same measurement live (bench.measure_int8_ceiling) -- this script is the stand-alone version for profiles/."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


if __name__ == "__main__":
    D = bench.Dist(1)
    out = {"shapes": {}}
    for n in (8192, 16384):
        out["shapes"][str(n)] = bench.measure_int8_ceiling(D, n=n)
    print(json.dumps(out, indent=1))
