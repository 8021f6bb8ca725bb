#!/usr/bin/env python
"""Random-gather ceiling (orph_bench_gather_peak) over footprints from L2-resident to HBM-resident."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from orpheum_b200 import gpu

ctx = gpu.default_context()
for mb in (12.5, 50, 100, 125, 250, 500, 2000):
    fb = int(mb * 1e6)
    print(f"{mb:7.1f} MB: plain {ctx.gather_peak(fb, 512, False) / 1e9:7.1f} G sectors/s, persisting window {ctx.gather_peak(fb, 512, True) / 1e9:7.1f}", flush=True)
