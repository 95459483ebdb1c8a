"""GPU parity of the handle API (bmq3d_*) against the oracle's restatement of
BimocqSolver::advanceBimocq (oracle/oracle3d.py:Solver): per step and over a run that crosses
velocity and scalar re-initialisations, with identical forcing.  Checks fields, maps, the
per-step scalars the reference prints, and the exact sequence of reinit frames."""
import numpy as np
import pytest

from gpufluidsimulation_b200 import scenes
from helpers import rel_linf

# Against the CPU ORACLE the per-step bound is looser than against the reference kernels
# (tests/test_solver_vs_reference_gpu.py holds the 1e-5 gate): glibc's expf and CUDA's differ in
# the last ulp and the reference's DMC formula 1 - exp(-a s) amplifies that (see test_kernels_gpu.py).
TOL_STEP = 5e-4

pytestmark = pytest.mark.gpu

NAMES = ("U", "V", "W", "RHO", "T")


def _forcing(cur, dt, nj):
    """Synthetic 'forces + projection' stand-in shared by both sides: buoyancy on v (reference
    formula, GPU_kernel.cu:804-823) and a smooth damping as the 'projection' change."""
    forced = [a.copy() for a in cur[:3]]
    forced[1] = (forced[1] + scenes.buoyancy_increment(cur[3], cur[4], 0.0, 0.2, dt, nj + 1)).astype(np.float32)
    final = [(0.98 * a).astype(np.float32) for a in forced]
    final += [cur[3].copy(), (0.995 * cur[4]).astype(np.float32)]
    return forced, final


def _setup(oracle, ni, nj, nk, L, dt, cfl, blend):
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    h = L / ni
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, L)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, cfl)
    so = oracle.Solver(ni, nj, nk, h, blend)
    so.set_initial(u, v, w, rho, T)
    sg = BimocqAdvection3D(ni, nj, nk, h, blend)
    sg.set_initial(u, v, w, rho, T)
    return so, sg


@pytest.mark.parametrize("blend", [1.0, 0.5])
def test_single_step_parity(cuda, oracle, blend):
    ni, nj, nk, dt = 40, 48, 36, 0.02
    so, sg = _setup(oracle, ni, nj, nk, 1.0, dt, 1.5, blend)
    # two steps so that the second one runs after the frame-0 reinitialisation (two-level blend active)
    for frame in range(2):
        so.advect(frame, dt, with_semilag=True)
        sg.advect(frame, dt, with_semilag=True)
        st = sg.stats()
        assert st["n_substeps"] == so.stats["n_substeps"] == 2
        assert abs(st["cfldt"] - so.stats["cfldt"]) <= 1e-6 * so.stats["cfldt"]
        got = [sg.download(n) for n in NAMES]
        for n, g, w_ in zip(NAMES, got, so.cur):
            assert rel_linf(g, w_) <= TOL_STEP, (frame, n, rel_linf(g, w_))
        for c, n in enumerate(("U_SEMI", "V_SEMI", "W_SEMI", "RHO_SEMI", "T_SEMI")):
            assert rel_linf(sg.download(n), so.semi[c]) <= TOL_STEP, n
        for c, ax in enumerate("XYZ"):
            assert rel_linf(sg.download("VFWD_" + ax), so.vel.fwd[c]) <= TOL_STEP
            assert rel_linf(sg.download("VBWD_" + ax), so.vel.bwd[c]) <= TOL_STEP
            assert rel_linf(sg.download("SBWD_" + ax), so.sca.bwd[c]) <= TOL_STEP
        # identical forcing on both sides, computed from the oracle's fields
        forced, final = _forcing(so.cur, dt, nj)
        so.accumulate(frame, dt, forced, final)
        # device path: hand the GPU the same fields but let it difference against ITS advected state
        sg.accumulate_host(frame, dt, forced, final)
        st = sg.stats()
        assert bool(st["vel_reinit"]) == so.stats["vel_reinit"]
        assert bool(st["scalar_reinit"]) == so.stats["scalar_reinit"]
        assert abs(st["vel_distortion"] - so.stats["vel_distortion"]) <= 1e-3 * max(1.0, so.stats["vel_distortion"])
        for c, n in enumerate(("U_INIT", "V_INIT", "W_INIT", "RHO_INIT", "T_INIT")):
            assert rel_linf(sg.download(n), so.init[c]) <= 2 * TOL_STEP, (frame, n, rel_linf(sg.download(n), so.init[c]))
        for c, n in enumerate(("U_PREV", "V_PREV", "W_PREV", "RHO_PREV", "T_PREV")):
            assert rel_linf(sg.download(n), so.prev[c]) <= 2 * TOL_STEP, (frame, n)
    sg.close()


def test_multi_step_parity_and_reinit_sequence(cuda, oracle):
    """16 free-running steps (each side advects its own state): identical reinit frames, and the
    drift stays small.  Per-step tolerance is 1e-5; over the run the bound asserted is 1e-3 and
    the measured drift is printed (-s)."""
    ni, nj, nk, dt = 32, 40, 32, 0.02
    so, sg = _setup(oracle, ni, nj, nk, 1.0, dt, 1.2, 1.0)
    o_frames, g_frames = [], []
    worst = 0.0
    for frame in range(16):
        so.advect(frame, dt)
        sg.advect(frame, dt)
        forced, final = _forcing(so.cur, dt, nj)
        so.accumulate(frame, dt, forced, final)
        cur_g = [sg.download(n) for n in NAMES]
        forced_g, final_g = _forcing(cur_g, dt, nj)
        sg.accumulate_host(frame, dt, forced_g, final_g)
        st = sg.stats()
        if so.stats["vel_reinit"]: o_frames.append(("v", frame))
        if so.stats["scalar_reinit"]: o_frames.append(("s", frame))
        if st["vel_reinit"]: g_frames.append(("v", frame))
        if st["scalar_reinit"]: g_frames.append(("s", frame))
        err = max(rel_linf(sg.download(n), w_) for n, w_ in zip(NAMES, so.cur))
        worst = max(worst, err)
    print(f"multi-step drift: worst rel Linf over 16 steps = {worst:.3e}; reinit frames {g_frames}")
    assert g_frames == o_frames
    assert worst <= 1e-3
    sg.close()


def test_device_change_fields_path(cuda, oracle):
    """bmq3d_accumulate with change fields written on the device (the drop-in for
    BimocqGPUSolver, where du_extern / du_proj are device buffers)."""
    import torch
    ni, nj, nk, dt = 32, 32, 32, 0.02
    so, sg = _setup(oracle, ni, nj, nk, 1.0, dt, 1.5, 1.0)
    so.advect(0, dt)
    sg.advect(0, dt)
    forced, final = _forcing(so.cur, dt, nj)
    d_ext = [np.ascontiguousarray(forced[c] - so.cur[c]) for c in range(3)]
    d_proj = [np.ascontiguousarray(final[c] - forced[c]) for c in range(3)]
    d_sca = [np.ascontiguousarray(final[c] - so.cur[c]) for c in (3, 4)]
    for c, n in enumerate(("DU_EXT", "DV_EXT", "DW_EXT")):
        sg.field(n).copy_(torch.from_numpy(d_ext[c]))
    for c, n in enumerate(("DU_PROJ", "DV_PROJ", "DW_PROJ")):
        sg.field(n).copy_(torch.from_numpy(d_proj[c]))
    sg.field("DRHO_EXT").copy_(torch.from_numpy(d_sca[0]))
    sg.field("DT_EXT").copy_(torch.from_numpy(d_sca[1]))
    for c, n in enumerate(NAMES):
        sg.field(n).copy_(torch.from_numpy(final[c]))
    torch.cuda.synchronize()
    so.accumulate(0, dt, forced, final)
    sg.accumulate(0, dt)
    for c, n in enumerate(("U_INIT", "V_INIT", "W_INIT", "RHO_INIT", "T_INIT")):
        assert rel_linf(sg.download(n), so.init[c]) <= 2 * TOL_STEP, n
    sg.close()


def test_error_paths(cuda):
    from gpufluidsimulation_b200 import BimocqLibraryError
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    with pytest.raises(BimocqLibraryError):
        BimocqAdvection3D(4, 4, 4, 0.25)
    s = BimocqAdvection3D(16, 16, 16, 1 / 16)
    with pytest.raises(BimocqLibraryError):
        s.advect(0, -1.0)
    with pytest.raises(KeyError):
        s.field("NOPE")
    s.close()
