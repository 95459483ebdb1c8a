"""TEST INFRASTRUCTURE (performance comparison; loads oracle/_ref with --ref).  CUDA-event timing of every stage of one BiMocq^2 advection step (handle API),
and of the reference's own kernels (oracle/_ref/libref3d.so) on the same data for comparison.
Usage: python tests/perf_stage_timing.py [n=256] [--ref]"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpufluidsimulation_b200 import scenes  # noqa: E402
from gpufluidsimulation_b200.solver3d import BimocqAdvection3D  # noqa: E402


def timed(fn, reps=3):
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
    with_ref = "--ref" in sys.argv
    dt = 0.02
    h = 1.0 / n
    dev = torch.device("cuda:0")
    u, v, w, rho, T = scenes.smoke_plume(n, n, n, 1.0, xp=torch, device=dev)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    s = BimocqAdvection3D(n, n, n, h, 1.0)
    for name, a in zip(("U", "V", "W", "RHO", "T"), (u, v, w, rho, T)):
        s.field(name).copy_(a)
    s.reset()
    # a few whole steps so that the maps are displaced
    for f in range(3):
        s.advect(f, dt); s.accumulate(f, dt)
    cells = n ** 3
    res = {}
    m = C.c_float()
    res["maxvel"] = timed(lambda: s.stage("maxvel", C.byref(m)))
    s.stage("set_cfl", 3, m)
    st = s.stats()
    sub = st["cfldt"]
    res["dmc_substep(x1)"] = timed(lambda: s.stage("dmc_substep", C.c_float(min(sub, dt))), reps=2)
    res["forward"] = timed(lambda: s.stage("forward", C.c_float(dt)), reps=1)
    for which, nm in ((0, "vel"), (1, "sca")):
        res[f"advect_{nm}"] = timed(lambda: s.stage("advect", which))
        res[f"error_{nm}"] = timed(lambda: s.stage("error", which))
        res[f"apply_{nm}"] = timed(lambda: s.stage("apply", which))
        res[f"accumulate_{nm}"] = timed(lambda: s.stage("accumulate", which), reps=1)
    a_, b_, c_ = C.c_float(), C.c_float(), C.c_float()
    res["distortion"] = timed(lambda: s.stage("distortion", C.byref(a_), C.byref(b_), C.byref(c_)))
    res["semilag"] = timed(lambda: s.stage("semilag", C.c_float(dt)), reps=1)
    nsub = st["n_substeps"]
    total = (res["maxvel"] + nsub * res["dmc_substep(x1)"] + res["forward"] + sum(
        res[k] for k in res if k.split("_")[0] in ("advect", "error", "apply", "accumulate")) + res["distortion"])
    out = {"n": n, "n_sub": nsub, "max_disp_z_cells": c_.value, "stage_ms": {k: round(v, 3) for k, v in res.items()},
           "step_ms_sum": round(total, 3), "cell_updates_per_s": cells / (total * 1e-3),
           "alg_bytes_per_cell": 448 + 60 * nsub,
           "roofline_frac_6450GBs": (448 + 60 * nsub) * cells / (total * 1e-3) / 6450.9e9}
    # whole-step timing through the public calls
    def whole():
        s.advect(5, dt); s.accumulate(5, dt)
    out["step_ms_api"] = round(timed(whole, reps=2), 3)
    print(json.dumps(out, indent=1))

    if with_ref:
        ref = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref3d.so"))
        F = C.POINTER(C.c_float)
        p = lambda t: C.cast(C.c_void_p(t.data_ptr()), F)
        fu, fv, fw = s.field("U"), s.field("V"), s.field("W")
        bx, by, bz = s.field("VBWD_X"), s.field("VBWD_Y"), s.field("VBWD_Z")
        fx, fy, fz = s.field("VFWD_X"), s.field("VFWD_Y"), s.field("VFWD_Z")
        ui, vi, wi = s.field("U_INIT"), s.field("V_INIT"), s.field("W_INIT")
        ou, ov, ow = s.field("U_ADV"), s.field("V_ADV"), s.field("W_ADV")
        tx, ty, tz = s.field("SBWDP_X").clone(), s.field("SBWDP_Y").clone(), s.field("SBWDP_Z").clone()
        r = {}
        hh = C.c_float(h); ni = C.c_int(n)
        r["ref_advect_velocity"] = timed(lambda: ref.gpu_advect_velocity(p(ou), p(ov), p(ow), p(ui), p(vi), p(wi), p(bx), p(by), p(bz), hh, ni, ni, ni, C.c_bool(False)), reps=2)
        r["ref_dmc(1 mapper)"] = timed(lambda: ref.gpu_solve_backwardDMC(p(fu), p(fv), p(fw), p(bx), p(by), p(bz), p(tx), p(ty), p(tz), hh, ni, ni, ni, C.c_float(min(sub, dt))), reps=2)
        cx, cy, cz = fx.clone(), fy.clone(), fz.clone()
        r["ref_forward(1 mapper)"] = timed(lambda: ref.gpu_solve_forward(p(fu), p(fv), p(fw), p(cx), p(cy), p(cz), hh, ni, ni, ni, C.c_float(sub), C.c_float(dt)), reps=1)
        r["ref_accumulate_velocity(1 change)"] = timed(lambda: ref.gpu_accumulate_velocity(p(ou), p(ov), p(ow), p(s.field("U_ERR")), p(s.field("V_ERR")), p(s.field("W_ERR")), p(fx), p(fy), p(fz), hh, ni, ni, ni, C.c_bool(False), C.c_float(1.0)), reps=1)
        d = s.field("RHO_ERR")
        r["ref_estimate(1 mapper)"] = timed(lambda: ref.gpu_estimate_distortion(p(d), p(bx), p(by), p(bz), p(fx), p(fy), p(fz), hh, ni, ni, ni), reps=2)
        print(json.dumps({"reference_kernels_ms": {k: round(v, 3) for k, v in r.items()}}, indent=1))


if __name__ == "__main__" and "--sources" not in sys.argv:
    main()


def source_terms(n=512):
    """GB/s of the streaming source-term symbols (gpu_add_buoyancy, gpu_diffuse_field, gpu_mad)."""
    from gpufluidsimulation_b200 import load_library
    from gpufluidsimulation_b200.solver3d import alloc_field
    lib = load_library()
    F = C.POINTER(C.c_float)
    p = lambda t: C.cast(C.c_void_p(t.data_ptr()), F)
    v = alloc_field((n, n + 1, n)); rho = alloc_field((n + 1, n + 1, n)); T = alloc_field((n + 1, n + 1, n))
    a = alloc_field((n, n, n)); b = alloc_field((n, n, n)); c = alloc_field((n, n, n))
    a.normal_(); T.normal_()
    cells = n ** 3
    out = {}
    t = timed(lambda: lib.gpu_add_buoyancy(p(v), p(rho), p(T), n, n, n, 0.1, 0.2, 0.02))
    out["gpu_add_buoyancy"] = {"ms": t, "alg_GBps": 4 * 4 * cells / t / 1e6}          # R rho,T,v + W v
    t = timed(lambda: lib.gpu_diffuse_field(p(a), p(b), p(c), n, n, n, 20, 0.1), reps=2)
    out["gpu_diffuse_field(20 sweeps)"] = {"ms": t, "alg_GBps": (20 * 3 + 4) * 4 * cells / t / 1e6}   # per sweep R field,in + W out
    t = timed(lambda: lib.gpu_mad(p(c), p(a), p(b), 0.5, 0.25, cells))
    out["gpu_mad"] = {"ms": t, "alg_GBps": 3 * 4 * cells / t / 1e6}
    # 8^3-blocked host container layout <-> dense: one read + one write per element
    nb = int(lib.bmq_blocked_elems(n + 1, n, n))
    import torch
    blk = torch.zeros(nb, dtype=torch.float32, device="cuda")
    u = alloc_field((n, n, n + 1)); u.normal_()
    t = timed(lambda: lib.bmq_linear_to_blocked(p(u), p(blk), n + 1, n, n, None))
    out["bmq_linear_to_blocked(u faces)"] = {"ms": t, "alg_GBps": 4 * (u.numel() + nb) / t / 1e6}
    t = timed(lambda: lib.bmq_blocked_to_linear(p(blk), p(u), n + 1, n, n, None))
    out["bmq_blocked_to_linear(u faces)"] = {"ms": t, "alg_GBps": 4 * (u.numel() + nb) / t / 1e6}
    print(json.dumps({"source_terms_512": out}, indent=1))


if "--sources" in sys.argv:
    source_terms()
