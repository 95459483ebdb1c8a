"""Parity at the BENCHMARKED shapes, against the reference's own CUDA kernels (oracle/_ref/libref3d.so =
bimocq3D/GPU_kernel.cu compiled unmodified) on the same B200 -- not against this library's own generic
kernels, and not only on 40^3 grids ("What's weak" 1):

* whole steps, per-step parity, at 128^3 and 256^3 (the pitch-specialised kernels: compile-time 128 / 256
  plane extents) for a power-of-two cell size and for the reference scene's h = 0.2 / ni;
* every pitch-specialised 512 x 512 kernel, symbol by symbol, on 512 x 512 x 20 grids (the kernels the
  512^3 benchmark runs; a full 512^3 state of the reference side does not fit next to ours in the test);
* the leapfrogging-rings scene of BASELINE configs[3] (scenes.leapfrog_rings);
* the 2D path at 256^2 (BASELINE configs[0]) and 1024^2 (configs[1]) against the reference's own 2D code
  (oracle/_ref/libref2d.so).
Tolerance: relative L-inf <= 1e-5 per step / per call (helpers.TOL_STEP); observed: 0 on the smooth scenes, a
last-ulp difference in ~1e-4 of the cells on the noisy fields of the symbol-by-symbol test."""
import ctypes as C
import os

import numpy as np
import pytest

from gpufluidsimulation_b200 import scenes
from helpers import TOL_STEP, DeviceStepper, load_reference_lib, rel_linf, run_gpu_symbol
from test_solver_vs_reference_gpu import INITS, NAMES, PREVS, _forcing, _maps_err, _sync_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def reflib(cuda):
    lib = load_reference_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libref3d.so not built")
    return lib


def _per_step(cuda, reflib, n, L, frames, scene, blend=1.0):
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    h, dt = L / n, 0.02
    u, v, w, rho, T = scene(n, n, n, L)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    ref = DeviceStepper(n, n, n, h, blend, lib=reflib)
    ref.set_initial(u, v, w, rho, T)
    sg = BimocqAdvection3D(n, n, n, h, blend)
    sg.set_initial(u, v, w, rho, T)
    worst = 0.0
    for frame in range(frames):
        _sync_state(ref, sg)
        ref.advect(frame, dt)
        sg.advect(frame, dt)
        cur = [t.cpu().numpy() for t in ref.cur]
        e = max(rel_linf(sg.download(nm), c) for nm, c in zip(NAMES, cur))
        em = _maps_err(ref, sg)
        assert e <= TOL_STEP and em <= TOL_STEP, (n, L, frame, e, em)
        forced, final = _forcing(cur, dt, n)
        ref.accumulate(frame, dt, forced, final)
        sg.accumulate_host(frame, dt, forced, final)
        st = sg.stats()
        assert (bool(st["vel_reinit"]), bool(st["scalar_reinit"])) == (ref.stats["vel_reinit"], ref.stats["scalar_reinit"])
        ei = max(rel_linf(sg.download(nm), t.cpu().numpy()) for nm, t in zip(INITS + PREVS, ref.init + ref.prev))
        assert ei <= 2 * TOL_STEP, (n, L, frame, "init/prev", ei)
        worst = max(worst, e, em, ei)
    sg.close()
    del ref
    cuda.cuda.empty_cache()
    return worst


@pytest.mark.parametrize("n,L,frames", [(128, 1.0, 3), (128, 0.2, 3), (256, 1.0, 2), (256, 0.2, 1)])
def test_whole_steps_at_benchmark_sizes_match_the_reference_kernels(cuda, reflib, n, L, frames):
    worst = _per_step(cuda, reflib, n, L, frames, scenes.smoke_plume)
    print(f"plume {n}^3, h = {L}/{n}: worst rel Linf over {frames} steps vs the reference kernels {worst:.2e}")


@pytest.mark.parametrize("n,L", [(96, 1.5), (128, 0.2)])
def test_leapfrogging_rings_match_the_reference_kernels(cuda, reflib, n, L):
    worst = _per_step(cuda, reflib, n, L, 3, scenes.leapfrog_rings, blend=0.5 if n == 96 else 1.0)
    print(f"rings {n}^3, h = {L}/{n}: worst rel Linf vs the reference kernels {worst:.2e}")


# ---- the 512 x 512 pitch-specialised kernels, symbol by symbol --------------------------------------------
def _slab_case(ni, nj, nk, h, seed):
    """Fields and displaced maps for an ni x nj x nk grid, built with numpy only (no oracle at this size)."""
    rng = np.random.default_rng(seed)
    shp = {"u": (nk, nj, ni + 1), "v": (nk, nj + 1, ni), "w": (nk + 1, nj, ni), "c": (nk, nj, ni)}
    z, y, x = np.meshgrid(np.arange(nk, dtype=np.float32), np.arange(nj, dtype=np.float32), np.arange(ni, dtype=np.float32), indexing="ij")
    def smooth(shape, s):
        zz, yy, xx = np.meshgrid(*[np.linspace(0, 1, m, dtype=np.float32) for m in shape], indexing="ij")
        r = np.random.default_rng(s)
        f = np.sin(2 * np.pi * (3 * xx + r.uniform()) + 1.3 * yy) * np.cos(2 * np.pi * (2 * yy + r.uniform())) * np.sin(2 * np.pi * zz + r.uniform())
        return (f + 0.05 * r.standard_normal(shape)).astype(np.float32)
    fields = {k: smooth(s, seed + q) for q, (k, s) in enumerate(shp.items())}
    fields2 = {k: smooth(s, seed + 10 + q) for q, (k, s) in enumerate(shp.items())}
    def disp(s):   # a few cells of smooth displacement
        return 2.5 * h * smooth(shp["c"], s)
    ident = [x * np.float32(h), y * np.float32(h), z * np.float32(h)]
    # maps stay inside [h, (n-1) h] like the maps the tracing kernels produce (traceRK3's clamp, GPU_kernel.cu:87-88):
    # the reference samples the velocity at the raw map value, so an out-of-domain map is an out-of-bounds read
    top = [np.float32(h) * np.float32(m - 1) for m in (ni, nj, nk)]
    fwd = [np.ascontiguousarray(np.clip(ident[c] + disp(seed + 20 + c), np.float32(h), top[c]), dtype=np.float32) for c in range(3)]
    bwd = [np.ascontiguousarray(np.clip(ident[c] - disp(seed + 30 + c), np.float32(h), top[c]), dtype=np.float32) for c in range(3)]
    vel = [(0.8 * h / 0.02) * smooth(shp[k], seed + 40 + q) for q, k in enumerate("uvw")]
    _ = rng
    return fields, fields2, fwd, bwd, vel


@pytest.mark.parametrize("L", [1.0, 0.2])
def test_512_plane_kernels_match_the_reference_symbol_by_symbol(cuda, reflib, L):
    from gpufluidsimulation_b200 import load_library
    ours = C.CDLL(load_library()._name)
    ni = nj = 512
    nk = 20
    h = float(np.float32(L) / np.float32(ni))
    f, f2, fwd, bwd, vel = _slab_case(ni, nj, nk, h, 11)
    cfldt = float(np.float32(h) / np.float32(max(np.abs(a).max() for a in vel)))

    stats = []

    def both(name, args, outputs):
        a = run_gpu_symbol(ours, name, args)
        b = run_gpu_symbol(reflib, name, args)
        for q in outputs:
            assert np.isfinite(a[q]).all()
            # bit-identical except for rare last-ulp cases: the fp32 fmaf lerp rounds once, the reference's
            # double-promoted lerp (GPU_kernel.cu:22-25) twice; the noisy fields of this
            # test provoke ~1e-4 of the cells, smooth flows none (the whole-step tests above)
            e = rel_linf(a[q], b[q])
            assert e <= 1e-6 <= TOL_STEP, (name, L, q, e)
            assert float((a[q] != b[q]).mean()) <= 1e-3, (name, L, q, float((a[q] != b[q]).mean()))
            stats.append((name, q, e, float((a[q] != b[q]).mean())))

    z = lambda k: np.zeros_like(f[k])
    both("gpu_advect_velocity", [z("u"), z("v"), z("w"), f["u"], f["v"], f["w"], *bwd, h, ni, nj, nk, False], (0, 1, 2))
    both("gpu_advect_field", [z("c"), f["c"], *bwd, h, ni, nj, nk, False], (0,))
    both("gpu_accumulate_velocity", [f["u"], f["v"], f["w"], f2["u"], f2["v"], f2["w"], *fwd, h, ni, nj, nk, False, 2.0], (3, 4, 5))
    both("gpu_accumulate_field", [f["c"], f2["c"], *fwd, h, ni, nj, nk, False, 1.0], (1,))
    both("gpu_compensate_velocity", [f["u"], f["v"], f["w"], f2["u"], f2["v"], f2["w"], z("u"), z("v"), z("w"), *fwd, *bwd, h, ni, nj, nk, False],
         (0, 1, 2, 6, 7, 8))
    both("gpu_compensate_field", [f["c"], f2["c"], z("c"), *fwd, *bwd, h, ni, nj, nk, False], (0, 2))
    both("gpu_solve_forward", [*vel, *[a.copy() for a in fwd], h, ni, nj, nk, cfldt, 0.02], (3, 4, 5))
    both("gpu_solve_backwardDMC", [*vel, *bwd, z("c"), z("c"), z("c"), h, ni, nj, nk, min(cfldt, 0.02)], (6, 7, 8))
    both("gpu_estimate_distortion", [z("c"), *bwd, *fwd, h, ni, nj, nk], (0,))
    print(f"512 x 512 x {nk} kernels vs the reference, h = {L}/512: worst rel Linf {max(s[2] for s in stats):.2e}, "
          f"largest fraction of cells differing in the last ulp {max(s[3] for s in stats):.1e}")


# ---- 2D at the BASELINE sizes -----------------------------------------------------------------------------
@pytest.mark.parametrize("n,frames", [(256, 3), (1024, 1)])
def test_2d_steps_at_benchmark_sizes_match_the_reference(cuda, n, frames):
    import ref2d
    from gpufluidsimulation_b200.solver2d import BimocqAdvection2D
    from test_solver2d_gpu import _blob, _inner, _sync, _vortex_field
    if not ref2d.available():
        pytest.skip("oracle/_ref/libref2d.so not built")
    os.environ.setdefault("BMQ_SHIM_THREADS", str(os.cpu_count() or 1))
    L, blend = 1.0, 1.0
    ref = ref2d.Ref2D(n, n, L, blend)
    dt = 0.5 * ref.h           # max |vel| = 2 -> CFL_frame ~ 1
    u, v = _vortex_field(n, n, L, strength=2.0 * L)
    rho = _blob(n, n, L, 0.5, 0.7, 0.12); T = _blob(n, n, L, 0.4, 0.3, 0.1)
    for mem, a in (("u", u), ("v", v), ("u_init", u), ("v_init", v), ("rho", rho), ("temperature", T), ("rho_init", rho), ("T_init", T)):
        ref.field(mem)[...] = a
    g = BimocqAdvection2D(n, n, ref.h, blend)
    worst = 0.0
    for frame in range(frames):
        _sync(ref, g)
        ref.phase_a(dt, frame)
        g.advect(frame, dt)
        for mem, name in (("u", "U"), ("v", "V"), ("rho", "RHO"), ("temperature", "T")):
            e = rel_linf(_inner(g.download(name)), _inner(ref.field(mem)))
            worst = max(worst, e)
            assert e <= TOL_STEP, (n, frame, name, e)
        for mem in ("forward_x", "forward_y", "backward_x", "backward_y", "forward_scalar_x", "backward_scalar_y"):
            e = rel_linf(g.download(ref2d.MEMBERS[mem]), ref.field(mem))
            worst = max(worst, e)
            assert e <= TOL_STEP, (n, frame, mem, e)
        adv = [ref.field(m).copy() for m in ("u", "v", "rho", "temperature")]
        for nme, a in zip(("U", "V", "RHO", "T", "U_SAVE", "V_SAVE", "RHO_SAVE", "T_SAVE"), adv + adv):
            g.upload(nme, a)
        ref.phase_b(dt, frame, adv[0], adv[1], adv[0], adv[1], adv[2], adv[3])
        g.accumulate_host(frame, dt, adv[0], adv[1], adv[0], adv[1], adv[2], adv[3])
        rc, gs = ref.counters(), g.stats()
        assert (gs["vel_remap"], gs["scalar_remap"]) == (rc["vel_remap"], rc["scalar_remap"])
        for mem in ("du", "dv", "drho", "dT", "u_init", "v_init", "rho_init", "u", "v"):
            e = rel_linf(g.download(ref2d.MEMBERS[mem]), ref.field(mem))
            worst = max(worst, e)
            assert e <= TOL_STEP, (n, frame, mem, e)
    print(f"2D {n}x{n}: worst rel Linf over {frames} step(s) vs the reference's own 2D code {worst:.2e}")
    g.close(); ref.close()
