"""Generates tests/golden/ref3d_kernels.npz by EXECUTING THE REFERENCE'S OWN CUDA KERNELS
(oracle/_ref/libref3d.so = /root/reference/src/bimocq3D/GPU_kernel.cu compiled unmodified by
oracle/Makefile) on a B200:  gpurun -- python tests/golden/make_golden_3d.py
The file holds, for every hot-path extern "C" gpu_* symbol, the seeded inputs and the reference's
outputs on a 20 x 18 x 22 grid with the reference scene's cell size h = 0.2/ni (non power of two).
tests/test_golden_cpu.py pins the CPU oracle against it; tests/test_golden_gpu.py pins the CUDA
library against it.  Also writes tests/golden/ref3d_projection.npz (pressure projection, see
projection_golden below); copy both from gpurun_out/ into tests/golden/."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import Case3D, load_reference_lib, run_gpu_symbol  # noqa: E402

NI, NJ, NK = 20, 18, 22
H = 0.2 / NI


def calls(c):
    """name -> (args, output indices into the array-only list)"""
    f, g = c.fields, c.fields2
    z = np.zeros_like
    return {
        "gpu_solve_forward": ([c.u, c.v, c.w, *c.fwd, c.h, c.ni, c.nj, c.nk, c.cfldt, c.dt], [3, 4, 5]),
        "gpu_solve_backwardDMC": ([c.u, c.v, c.w, *c.bwd, *[z(m) for m in c.bwd], c.h, c.ni, c.nj, c.nk, 0.7 * c.cfldt], [6, 7, 8]),
        "gpu_semilag": ([z(f["v"]), f["v"], c.u, c.v, c.w, 0, 1, 0, c.h, c.ni, c.nj, c.nk, c.cfldt, -c.dt], [0]),
        "gpu_advect_velocity": ([z(f["u"]), z(f["v"]), z(f["w"]), f["u"], f["v"], f["w"], *c.bwd, c.h, c.ni, c.nj, c.nk, False], [0, 1, 2]),
        "gpu_advect_field": ([z(f["c"]), f["c"], *c.bwd, c.h, c.ni, c.nj, c.nk, False], [0]),
        "gpu_compensate_velocity": ([f["u"], f["v"], f["w"], g["u"], g["v"], g["w"], z(f["u"]), z(f["v"]), z(f["w"]), *c.fwd, *c.bwd, c.h, c.ni, c.nj, c.nk, False], list(range(9))),
        "gpu_compensate_field": ([f["c"], g["c"], z(f["c"]), *c.fwd, *c.bwd, c.h, c.ni, c.nj, c.nk, False], [0, 1, 2]),
        "gpu_accumulate_velocity": ([f["u"], f["v"], f["w"], g["u"], g["v"], g["w"], *c.fwd, c.h, c.ni, c.nj, c.nk, False, 2.0], [3, 4, 5]),
        "gpu_accumulate_field": ([f["c"], g["c"], *c.fwd, c.h, c.ni, c.nj, c.nk, False, 1.0], [1]),
        "gpu_advect_vel_double": ([f["u"], f["v"], f["w"], g["u"], g["v"], g["w"], *c.bwd, *c.bwd_prev, c.h, c.ni, c.nj, c.nk, False, 0.5], [0, 1, 2]),
        "gpu_advect_field_double": ([f["c"], g["c"], *c.bwd, *c.bwd_prev, c.h, c.ni, c.nj, c.nk, False, 0.25], [0]),
        "gpu_estimate_distortion": ([z(f["c"]), *c.bwd, *c.fwd, c.h, c.ni, c.nj, c.nk], [0]),
        # source terms
        "gpu_emit_smoke": ([f["u"], f["v"], f["w"], f["c"], g["c"], c.h, c.ni, c.nj, c.nk, 0.45 * c.h * c.ni, 0.4 * c.h * c.nj,
                            0.5 * c.h * c.nk, 0.22 * c.h * c.ni, 1.0, 50.0, 1.0], [0, 1, 2, 3, 4]),
        "gpu_add_buoyancy": ([f["v"], f["c"], g["c"], c.ni, c.nj, c.nk, 0.3, 0.7, 0.02], [0]),
        "gpu_diffuse_field": ([f["c"], z(f["c"]), z(f["c"]), c.ni, c.nj, c.nk, 4, 0.37], [0]),
        "gpu_mad": ([z(f["u"]), f["u"], g["u"], 0.75, -1.25, int(f["u"].size)], [0]),
    }


def main():
    lib = load_reference_lib()
    assert lib is not None, "build oracle/_ref/libref3d.so first (make -C oracle)"
    c = Case3D(NI, NJ, NK, H, seed=11)
    store = {}
    for name, (args, outs) in calls(c).items():
        got = run_gpu_symbol(lib, name, args)
        for q in outs:
            store[f"{name}:out{q}"] = got[q]
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref3d_kernels.npz"), **store)
    print("wrote", len(store), "arrays")
    projection_golden(lib)


PROJECTION_CASE = (24, 20, 28, 2, 4)   # ni, nj, nk, levels, iterations


def projection_golden(lib):
    """tests/golden/ref3d_projection.npz: the reference's gpu_multi_grid_conjugate_gradient on the
    seeded velocity of tests/test_projection_gpu.py (inputs are regenerated from the seed there)."""
    from test_projection_gpu import run_legacy

    out = run_legacy(lib, *PROJECTION_CASE)
    keep = {k: out[k] for k in ("u", "v", "w", "p", "result")}
    keep["result"] = np.concatenate([out["result"][:16], np.zeros(2000 - 16), out["result"][2000:2016], np.zeros(4096 - 2016)])
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref3d_projection.npz"), case=np.array(PROJECTION_CASE), **keep)
    print("wrote projection golden", PROJECTION_CASE)


if __name__ == "__main__":
    main()
