"""Developer tool for ncu launch lists: a few BiMocq2D steps at n x n (default 1024) through the handle API."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
print(bench.measure_2d(torch, n=n, steps=steps, warm=2, cpu_leg=False))
