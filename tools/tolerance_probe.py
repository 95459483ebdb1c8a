"""Developer tool: relative L-inf of the tolerance mode against the exact path, per frame (general h)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpufluidsimulation_b200 import load_library
from test_tolerance_mode_gpu import make, rel, FIELDS

lib = load_library()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
exact, dt = make(lib, n, n - 8, n + 8, 0.2, 0)
fast, _ = make(lib, n, n - 8, n + 8, 0.2, 1)
names = FIELDS + ("VBWD_X", "VBWD_Z", "VFWD_X", "SBWD_X", "U_INIT")
for frame in range(12):
    for s in (exact, fast):
        s.advect(frame, dt)
    mid = {k: rel(fast.download(k), exact.download(k)) for k in names}
    for s in (exact, fast):
        s.apply_buoyancy(0.2, dt); s.accumulate(frame, dt)
    print(frame, {k: f"{v:.1e}" for k, v in mid.items()}, "reinit", exact.stats()["vel_reinit"], fast.stats()["vel_reinit"], flush=True)
lib.bmq_set_tolerance_mode(0)
