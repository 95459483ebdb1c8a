"""Developer tool: per-stage CUDA-event timings of the handle API at one size for both gather variants
(0 = windowed, 1 = z-marching) and for a power-of-two and a general cell size.
Usage: python tools/stage_ab.py [n=512]"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpufluidsimulation_b200 import load_library, scenes  # noqa: E402
from gpufluidsimulation_b200.solver3d import BimocqAdvection3D  # noqa: E402


def run(n, L, variant, steps=4, warm=3, tolerance=0):
    lib = load_library()
    lib.bmq_set_gather_variant(variant)
    lib.bmq_set_tolerance_mode(tolerance)
    dt, h = 0.02, L / n
    dev = torch.device("cuda:0")
    u, v, w, rho, T = scenes.smoke_plume(n, n, n, L, xp=torch, device=dev)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    s = BimocqAdvection3D(n, n, n, h, 1.0)
    s.set_initial_device(u, v, w, rho, T)
    del u, v, w, rho, T
    for f in range(warm):
        s.advect(f, dt); s.apply_buoyancy(1e-2, dt); s.accumulate(f, dt)
    torch.cuda.synchronize()
    s.timing_enable(True); s.timing_read()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in range(warm, warm + steps):
        s.advect(f, dt); s.apply_buoyancy(1e-2, dt); s.accumulate(f, dt)
    e1.record(); e1.synchronize()
    st = s.timing_read()
    s.close()
    lib.bmq_set_gather_variant(1)
    lib.bmq_set_tolerance_mode(0)
    return {"n": n, "L": L, "variant": variant, "ms_per_step": round(e0.elapsed_time(e1) / steps, 3),
            "stage_ms": {k: round(v[0] / steps, 3) for k, v in st.items() if v[1] > 0}}


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    combos = sys.argv[2] if len(sys.argv) > 2 else "p0,p1,g0,g1"     # p/g = power-of-two / general h, 0/1 = gather variant
    for c in combos.split(","):
        r = run(n, 1.0 if c[0] == "p" else 0.2, int(c[1]), tolerance=1 if c[0] == "t" else 0)   # t = general h, tolerance mode
        r["tolerance_mode"] = c[0] == "t"
        r["lib"] = os.path.basename(os.environ.get("BMQ_LIB", "default"))
        print(json.dumps(r), flush=True)
