"""Developer tool: run the gather stages one by one with both gather variants on the same state and report
where they differ.  Usage: python tools/march_diff.py ni nj nk L"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpufluidsimulation_b200 import load_library, scenes  # noqa: E402
from gpufluidsimulation_b200.solver3d import BimocqAdvection3D  # noqa: E402

ni, nj, nk = (int(a) for a in sys.argv[1:4])
L = float(sys.argv[4])
lib = load_library()
dt, h = 0.02, L / ni
u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, L)
u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
s = BimocqAdvection3D(ni, nj, nk, h, 1.0)
s.set_initial(u, v, w, rho, T)
m = C.c_float()
s.stage("maxvel", C.byref(m)); s.stage("set_cfl", 0, m)
cfldt = s.stats()["cfldt"]
t = 0.0
while t < dt:
    sub = min(cfldt, dt - t); s.stage("dmc_substep", C.c_float(sub)); t += sub
s.stage("forward", C.c_float(dt))
OUT = {"advect": ("U_ADV", "V_ADV", "W_ADV", "RHO_ADV", "T_ADV"), "error": ("U_ERR", "V_ERR", "W_ERR", "RHO_ERR", "T_ERR"),
       "apply": ("U", "V", "W", "RHO", "T"), "accumulate": ("U_INIT", "V_INIT", "W_INIT", "RHO_INIT", "T_INIT")}
for stage in ("advect", "error", "apply", "accumulate"):
    if stage == "accumulate":
        s.apply_buoyancy(0.2, dt)
        s.field("DU_PROJ").copy_(0.01 * s.field("U")); s.field("DRHO_EXT").copy_(0.01 * s.field("RHO"))
    before = {n: s.field(n).clone() for n in OUT[stage]}
    res = {}
    for variant in (0, 1):
        for n in OUT[stage]:
            s.field(n).copy_(before[n])
        lib.bmq_set_gather_variant(variant)
        s.stage(stage, 2)
        torch.cuda.synchronize()
        res[variant] = {n: s.field(n).cpu().numpy().copy() for n in OUT[stage]}
    for n in OUT[stage]:
        a, b = res[0][n], res[1][n]
        d = np.abs(a - b)
        if d.max() > 0:
            k, j, i = np.unravel_index(np.argmax(d), d.shape)
            bad = np.argwhere(d > 0)
            print(f"{stage:10s} {n:8s} DIFF max {d.max():.3e} at (i,j,k)=({i},{j},{k}) of {a.shape[::-1]}; {len(bad)} cells differ; "
                  f"i range {bad[:,2].min()}..{bad[:,2].max()} j {bad[:,1].min()}..{bad[:,1].max()} k {bad[:,0].min()}..{bad[:,0].max()}; "
                  f"values {a[k,j,i]!r} vs {b[k,j,i]!r}")
        else:
            print(f"{stage:10s} {n:8s} identical")
lib.bmq_set_gather_variant(1)
s.close()
