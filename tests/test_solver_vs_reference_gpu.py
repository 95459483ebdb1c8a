"""The parity gate proper: the handle API (fused sm_100a kernels + scheduler) against the
REFERENCE'S OWN CUDA KERNELS (oracle/_ref/libref3d.so, GPU_kernel.cu compiled unmodified) driven
through the reference's call sequence (helpers.DeviceStepper = BimocqSolver::advanceBimocq with
MapperBase semantics), on the same B200, same inputs.

* per-step parity: both sides start every step from the identical state (the reference side's)
  -> relative L-inf <= 1e-5 on u, v, w, rho, T, the init buffers and all maps, and identical
  reinitialisation decisions;
* free-running drift over 100 steps is measured and bounded loosely: the reference's DMC formula
  1 - exp(-a s) (GPU_kernel.cu:194-196) amplifies last-ulp differences of the previous step's
  velocity by up to 6e-8/(a s), so trajectories separate at velocity extrema no matter how the
  arithmetic is arranged; the fraction of cells within 1e-5 is reported next to the L-inf."""
import numpy as np
import pytest

from gpufluidsimulation_b200 import scenes
from helpers import TOL_STEP, DeviceStepper, load_reference_lib, rel_linf

pytestmark = pytest.mark.gpu
NAMES = ("U", "V", "W", "RHO", "T")
INITS = ("U_INIT", "V_INIT", "W_INIT", "RHO_INIT", "T_INIT")
PREVS = ("U_PREV", "V_PREV", "W_PREV", "RHO_PREV", "T_PREV")


def _forcing(cur, dt, nj):
    forced = [a.copy() for a in cur[:3]]
    forced[1] = (forced[1] + scenes.buoyancy_increment(cur[3], cur[4], 0.0, 0.2, dt, nj + 1)).astype(np.float32)
    final = [(0.98 * a).astype(np.float32) for a in forced]
    final += [cur[3].copy(), (0.995 * cur[4]).astype(np.float32)]
    return forced, final


def _sync_state(ref, sg):
    """Copy the reference side's complete advection state into the handle."""
    for c in range(5):
        sg.field(NAMES[c]).copy_(ref.cur[c]); sg.field(INITS[c]).copy_(ref.init[c]); sg.field(PREVS[c]).copy_(ref.prev[c])
    for pre, m in (("V", ref.vel), ("S", ref.sca)):
        for ax in "XYZ":
            sg.field(f"{pre}FWD_{ax}").copy_(getattr(m, "Forward" + ax))
            sg.field(f"{pre}BWD_{ax}").copy_(getattr(m, "Backward" + ax))
            sg.field(f"{pre}BWDP_{ax}").copy_(getattr(m, "Backward" + ax + "Prev"))


def _maps_err(ref, sg):
    e = 0.0
    for pre, m in (("V", ref.vel), ("S", ref.sca)):
        for ax in "XYZ":
            e = max(e, rel_linf(sg.field(f"{pre}FWD_{ax}").cpu().numpy(), getattr(m, "Forward" + ax).cpu().numpy()))
            e = max(e, rel_linf(sg.field(f"{pre}BWD_{ax}").cpu().numpy(), getattr(m, "Backward" + ax).cpu().numpy()))
    return e


@pytest.fixture(scope="module")
def reflib(cuda):
    lib = load_reference_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libref3d.so not built")
    return lib


@pytest.mark.parametrize("L,blend", [(1.0, 1.0), (0.2, 1.0), (0.2, 0.5)])
def test_per_step_parity_with_reference_kernels(cuda, reflib, L, blend):
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    ni, nj, nk = 40, 48, 36
    h = L / ni
    dt = 0.02
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, L)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    ref = DeviceStepper(ni, nj, nk, h, blend, lib=reflib)
    ref.set_initial(u, v, w, rho, T)
    sg = BimocqAdvection3D(ni, nj, nk, h, blend)
    sg.set_initial(u, v, w, rho, T)
    worst_f = worst_m = worst_i = 0.0
    reinit_frames = []
    for frame in range(14):
        _sync_state(ref, sg)
        ref.advect(frame, dt)
        sg.advect(frame, dt)
        cur = [t.cpu().numpy() for t in ref.cur]
        e = max(rel_linf(sg.download(n), c) for n, c in zip(NAMES, cur))
        worst_f = max(worst_f, e)
        assert e <= TOL_STEP, (frame, "fields", e)
        em = _maps_err(ref, sg)
        worst_m = max(worst_m, em)
        assert em <= TOL_STEP, (frame, "maps", em)
        forced, final = _forcing(cur, dt, nj)
        ref.accumulate(frame, dt, forced, final)
        sg.accumulate_host(frame, dt, forced, final)
        st = sg.stats()
        assert bool(st["vel_reinit"]) == ref.stats["vel_reinit"], (frame, st, ref.stats)
        assert bool(st["scalar_reinit"]) == ref.stats["scalar_reinit"], (frame, st, ref.stats)
        assert abs(st["vel_distortion"] - ref.stats["vel_distortion"]) <= 1e-3 * max(1.0, ref.stats["vel_distortion"])
        if st["vel_reinit"]: reinit_frames.append(("v", frame))
        if st["scalar_reinit"]: reinit_frames.append(("s", frame))
        ei = max(rel_linf(sg.download(n), t.cpu().numpy()) for n, t in zip(INITS, ref.init))
        ei = max(ei, max(rel_linf(sg.download(n), t.cpu().numpy()) for n, t in zip(PREVS, ref.prev)))
        worst_i = max(worst_i, ei)
        assert ei <= 2 * TOL_STEP, (frame, "init/prev", ei)
    print(f"per-step parity vs reference kernels (L={L}, blend={blend}): fields {worst_f:.2e}, maps {worst_m:.2e}, "
          f"init/prev {worst_i:.2e}; reinit frames {reinit_frames}")
    assert len(reinit_frames) >= 2
    sg.close()


def test_free_running_drift_vs_reference_kernels(cuda, reflib):
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    ni, nj, nk, L, dt = 40, 48, 36, 1.0, 0.02
    h = L / ni
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, L)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    ref = DeviceStepper(ni, nj, nk, h, 1.0, lib=reflib)
    ref.set_initial(u, v, w, rho, T)
    sg = BimocqAdvection3D(ni, nj, nk, h, 1.0)
    sg.set_initial(u, v, w, rho, T)
    rf, gf = [], []
    report = []
    for frame in range(100):
        ref.advect(frame, dt); sg.advect(frame, dt)
        cur_r = [t.cpu().numpy() for t in ref.cur]
        cur_g = [sg.download(n) for n in NAMES]
        ref.accumulate(frame, dt, *_forcing(cur_r, dt, nj))
        sg.accumulate_host(frame, dt, *_forcing(cur_g, dt, nj))
        st = sg.stats()
        if ref.stats["vel_reinit"]: rf.append(("v", frame))
        if ref.stats["scalar_reinit"]: rf.append(("s", frame))
        if st["vel_reinit"]: gf.append(("v", frame))
        if st["scalar_reinit"]: gf.append(("s", frame))
        errs = [rel_linf(g, r) for g, r in zip(cur_g, cur_r)]
        within = min(float((np.abs(g - r) <= 1e-5 * np.abs(r).max()).mean()) for g, r in zip(cur_g, cur_r))
        report.append((frame, max(errs), within))
    for frame, e, wi in report[::20] + report[-1:]:
        print(f"free-running frame {frame:2d}: rel Linf {e:.2e}, cells within 1e-5: {100 * wi:.3f}%")
    print("reinit frames: ours", gf, "reference", rf)
    # the first ten reinitialisations must coincide; later ones may shift by a frame when the
    # distortion ratio sits at its threshold (free-running trajectories separate, see module doc)
    assert gf[:10] == rf[:10]
    assert report[-1][2] >= 0.95
    sg.close()
