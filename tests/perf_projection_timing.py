"""TEST INFRASTRUCTURE (performance comparison; loads oracle/_ref with --ref).  Times the pressure projection (gpu_multi_grid_conjugate_gradient) on one GPU with CUDA events:
libbimocq_b200.so, and with --ref the reference's own kernels (oracle/_ref/libref3d.so) on the
same inputs.  Prints one JSON object.

  python tests/perf_projection_timing.py --n 256 --iters 10 --ref

Algorithmic bytes per PCG iteration (fp64 arrays of N0 = n^3 cells; a Jacobi sweep counted as
read x, read b, write x; level l has N_l cells):
  CG step            q = A dir + dot 2, x += a dir 3, r = b - A x 3                  =  8 N0
  V-cycle, per level 32 + 4 sweeps 108, residual 3, restrict 1 (+ coarse write), prolong 2
                     (coarsest: 32 sweeps only)
  after the cycle    x += e 3, r = b - A x 3 (+ max), dot(r,r) 1, dir update 3       = 10 N0
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def alg_bytes_per_iteration(dims):
    n = [a * b * c for a, b, c in dims]
    total = 18 * n[0]
    for l, cells in enumerate(n):
        if l == len(n) - 1:
            total += 96 * cells
        else:
            total += (108 + 3 + 1 + 2) * cells + n[l + 1]
    return 8 * total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--levels", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--ref", action="store_true")
    a = ap.parse_args()

    import torch

    from gpufluidsimulation_b200 import projection as pj
    from gpufluidsimulation_b200.solver3d import alloc_field
    from helpers import load_reference_lib

    n = a.n
    levels = a.levels or pj.max_levels(n, n, n)
    dims = pj.level_dims(n, n, n, levels)
    g = torch.Generator(device="cuda").manual_seed(1)
    shapes = [(n, n, n + 1), (n, n + 1, n), (n + 1, n, n)]
    vel0 = [torch.randn(s, device="cuda", generator=g) for s in shapes]
    vel = [alloc_field(s) for s in shapes]
    cells = n ** 3
    bufs = [pj.alloc_double(cells, n * n + n + 2) for _ in range(6)]
    result = pj.alloc_double(4096)
    lv, keep = pj.make_levels(n, n, n, levels)
    out = {"n": n, "levels": levels, "level_dims": dims, "iters": a.iters,
           "alg_bytes_per_iteration": alg_bytes_per_iteration(dims)}

    def run(lib, tag):
        times = []
        for rep in range(a.reps + 1):
            for t, s in zip(vel, vel0):
                t.copy_(s)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            pj.projection_multi_grid(*vel, *bufs, result, lv, a.iters, 0.5, lib=lib)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                times.append(e0.elapsed_time(e1))
        ms = sorted(times)[len(times) // 2]
        hist = result[2000:2001 + a.iters].cpu().numpy()
        out[tag] = {"ms_per_solve": ms, "ms_per_iteration": ms / a.iters,
                    "alg_GBps": out["alg_bytes_per_iteration"] * a.iters / ms / 1e6,
                    "residual_first_last": [float(hist[0]), float(hist[-1])]}
        return [v.clone() for v in vel]

    mine = run(None, "libbimocq_b200")
    if a.ref:
        ref = load_reference_lib()
        assert ref is not None, "oracle/_ref/libref3d.so missing"
        theirs = run(ref, "reference_kernels")
        out["bit_identical_velocity"] = all(torch.equal(x, y) for x, y in zip(mine, theirs))
        out["speedup"] = out["reference_kernels"]["ms_per_solve"] / out["libbimocq_b200"]["ms_per_solve"]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
