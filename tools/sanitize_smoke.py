"""Developer tool: a small end-to-end run of every 3D and 2D stage (both h paths, blend on) for
compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpufluidsimulation_b200 import scenes, zslab  # noqa: E402
from gpufluidsimulation_b200.solver2d import BimocqAdvection2D  # noqa: E402
from gpufluidsimulation_b200.solver3d import BimocqAdvection3D  # noqa: E402

for L in (1.0, 0.2):
    ni, nj, nk, dt = 24, 20, 28, 0.02
    h = L / ni
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, L)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    s = BimocqAdvection3D(ni, nj, nk, h, 0.5)
    s.set_initial(u, v, w, rho, T)
    for f in range(3):
        s.advect(f, dt, with_semilag=True)
        s.apply_buoyancy(0.2, dt)
        s.accumulate(f, dt)
    s.close()
    ranks = [zslab.CudaSlabRank(ni, nj, nk, h, 0.5, r, 2, 8) for r in range(2)]
    for r in ranks:
        for name, a in zip(zslab.CUR, (u, v, w, rho, T)):
            _, p0, npl, _, _ = r.solver.field_info(name)
            r.solver.upload(name, a[p0:p0 + npl])
        r.solver.reset()
    st = zslab.ZSlabStepper(ranks, zslab.LocalComm(2), 0.5)
    for f in range(3):
        st.advect(f, dt)
        st.accumulate(f, dt)
    for r in ranks:
        r.close()
n = 40
g = BimocqAdvection2D(n, n + 8, 1.0 / n, 0.5)
x = np.linspace(0, 1, n + 1)
g.upload("U", np.sin(np.pi * x)[None, :].repeat(n + 8, 0).astype(np.float32))
g.upload("V", np.zeros((n + 9, n), np.float32))
for f in range(3):
    g.advect(f, 0.01)
    g.field("U_FORCED").copy_(g.field("U")); g.field("V_FORCED").copy_(g.field("V"))
    g.accumulate(f, 0.01)
g.close()
print("sanitize_smoke done")
