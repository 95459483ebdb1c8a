"""Developer tool: a few 1024^2 BiMocq2D steps (for ncu launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

print(bench.measure_2d(torch, n=int(sys.argv[1]) if len(sys.argv) > 1 else 1024, steps=3, warm=2))
