"""Developer tool for ncu captures: a few whole steps of the handle API at n^3.
Usage: python tools/one_step.py [n=256] [steps=3] [variant=1] [L=1.0]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpufluidsimulation_b200 import load_library, scenes  # noqa: E402
from gpufluidsimulation_b200.solver3d import BimocqAdvection3D  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 1
L = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
lib = load_library()
lib.bmq_set_gather_variant(variant)
dt, h = 0.02, L / n
u, v, w, rho, T = scenes.smoke_plume(n, n, n, L, xp=torch, device=torch.device("cuda:0"))
u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
s = BimocqAdvection3D(n, n, n, h, 1.0)
s.set_initial_device(u, v, w, rho, T)
for f in range(steps):
    s.advect(f, dt); s.apply_buoyancy(1e-2, dt); s.accumulate(f, dt)
torch.cuda.synchronize()
print("one_step ok", n, steps, variant, s.stats())
s.close()
