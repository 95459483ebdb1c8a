"""2D parity on the GPU: bmq2d_* against THE REFERENCE'S OWN 2D CODE (oracle/_ref/libref2d.so,
BimocqSolver2D.cpp compiled unmodified; its hot-path methods are called in advanceBIMOCQ's order by
oracle/ref2d_wrapper.cpp).  Per step, from identical state: u, v, rho, T, the change buffers, all
maps, the remapping conditions and decisions.  Tolerance 1e-5 relative L-inf, measured values are
printed.  Rows j=0 and j=nj-1 are excluded for fields that went through clampExtrema2, where the
reference reads outside its arrays (undefined, see solver2d.cu:k2_clamp_extrema)."""
import numpy as np
import pytest

import ref2d
from helpers import rel_linf

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _vortex_field(ni, nj, L, strength=1.0):
    """Analytic single-vortex-in-a-box stream function psi = sin^2(pi x) sin^2(pi y) / pi,
    sampled as a discretely divergence-free MAC velocity (u = dpsi/dy, v = -dpsi/dx)."""
    h = L / ni
    xn = np.arange(ni + 1) * h / L
    yn = np.arange(nj + 1) * h / (h * nj)
    psi = strength * (np.sin(np.pi * xn)[None, :] ** 2) * (np.sin(np.pi * yn)[:, None] ** 2) * L / np.pi
    u = (psi[1:, :] - psi[:-1, :]) / h            # (nj, ni+1)
    v = -(psi[:, 1:] - psi[:, :-1]) / h           # (nj+1, ni)
    return u.astype(np.float32), v.astype(np.float32)


def _blob(ni, nj, L, cx, cy, r):
    h = L / ni
    x = (np.arange(ni) + 0.5) * h
    y = (np.arange(nj) + 0.5) * h
    d2 = (x[None, :] - cx * L) ** 2 + (y[:, None] - cy * h * nj) ** 2
    return np.exp(-d2 / (r * L) ** 2).astype(np.float32)


def _sync(ref, gpu):
    for member, name in ref2d.MEMBERS.items():
        gpu.upload(name, ref.field(member))
    c = ref.counters()
    gpu.set_counters(c["last_remesh"], c["last_scalar_remesh"])


def _inner(a):
    return a[1:-1, :]


@pytest.mark.parametrize("ni,nj,L,blend", [(64, 64, 1.0, 1.0), (48, 80, 0.2, 0.5)])
def test_per_step_parity_with_reference_2d(cuda, ni, nj, L, blend):
    if not ref2d.available():
        pytest.skip("oracle/_ref/libref2d.so not built")
    from gpufluidsimulation_b200.solver2d import BimocqAdvection2D
    dt = 0.02 * L
    ref = ref2d.Ref2D(ni, nj, L, blend)
    u, v = _vortex_field(ni, nj, L, strength=2.0 * L)
    ref.field("u")[...] = u; ref.field("v")[...] = v
    ref.field("u_init")[...] = u; ref.field("v_init")[...] = v
    rho = _blob(ni, nj, L, 0.5, 0.7, 0.12); T = _blob(ni, nj, L, 0.4, 0.3, 0.1)
    ref.field("rho")[...] = rho; ref.field("temperature")[...] = T
    ref.field("rho_init")[...] = rho; ref.field("T_init")[...] = T
    gpu = BimocqAdvection2D(ni, nj, ref.h, blend)
    worst = {}
    remaps = []
    for frame in range(12):
        _sync(ref, gpu)
        ref.phase_a(dt, frame)
        gpu.advect(frame, dt)
        assert abs(gpu.stats()["cfl"] - ref.scalars()["cfl"]) <= 1e-6 * ref.scalars()["cfl"]
        for member, name in (("u", "U"), ("v", "V"), ("rho", "RHO"), ("temperature", "T")):
            e = rel_linf(_inner(gpu.download(name)), _inner(ref.field(member)))
            worst[name] = max(worst.get(name, 0.0), e)
            assert e <= TOL, (frame, name, e)
        for member in ("forward_x", "forward_y", "backward_x", "backward_y", "forward_scalar_x", "backward_scalar_y"):
            e = rel_linf(gpu.download(ref2d.MEMBERS[member]), ref.field(member))
            worst["maps"] = max(worst.get("maps", 0.0), e)
            assert e <= TOL, (frame, member, e)
        # caller stand-in: a buoyancy-like force on v and a damping "projection"
        adv = [ref.field(m).copy() for m in ("u", "v", "rho", "temperature")]
        v_forced = adv[1].copy()
        v_forced[1:-1, :] += np.float32(0.5 * dt) * (adv[3][1:, :] + adv[3][:-1, :])
        u_final = (0.99 * adv[0]).astype(np.float32); v_final = (0.99 * v_forced).astype(np.float32)
        # same forced / final fields for both sides (the reference's advected state is the common input)
        for nme, a in zip(("U", "V", "RHO", "T", "U_SAVE", "V_SAVE", "RHO_SAVE", "T_SAVE"), adv + adv):
            gpu.upload(nme, a)
        ref.phase_b(dt, frame, adv[0], v_forced, u_final, v_final, adv[2], adv[3])
        gpu.accumulate_host(frame, dt, adv[0], v_forced, u_final, v_final, adv[2], adv[3])
        rs, rc, gs = ref.scalars(), ref.counters(), gpu.stats()
        assert (gs["vel_remap"], gs["scalar_remap"]) == (rc["vel_remap"], rc["scalar_remap"]), (frame, gs, rs, rc)
        assert abs(gs["vel_condition"] - rs["vel_condition"]) <= 1e-4 * max(1.0, abs(rs["vel_condition"]))
        if rc["vel_remap"]: remaps.append(("v", frame))
        if rc["scalar_remap"]: remaps.append(("s", frame))
        for member in ("du", "dv", "drho", "dT", "u_init", "v_init", "rho_init", "u_origin", "du_prev", "u", "v", "u_temp",
                       "backward_xprev", "forward_x", "backward_scalar_x"):
            e = rel_linf(gpu.download(ref2d.MEMBERS[member]), ref.field(member))
            worst["phaseB"] = max(worst.get("phaseB", 0.0), e)
            assert e <= TOL, (frame, member, e)
    print(f"2D per-step parity vs the reference ({ni}x{nj}, L={L}, blend={blend}): worst rel Linf {worst}; remaps {remaps}")
    assert len(remaps) >= 1
    ref.close(); gpu.close()


def test_levelset_mode_rigid_rotation_matches_reference(cuda):
    """advect_levelset = true (bimocq2D/main.cpp:135-223, Zalesak's disk): prescribed rigid rotation,
    only the scalar maps are evolved, no compensation, no accumulation (BimocqSolver2D.cpp:405-435,
    467, 481).  Per-step parity with the reference's own code, plus the analytic check the scene
    exists for: after the rotation the notched disk is still where rigid rotation puts it."""
    if not ref2d.available():
        pytest.skip("oracle/_ref/libref2d.so not built")
    from gpufluidsimulation_b200.solver2d import BimocqAdvection2D
    n, L = 64, 1.0
    h = L / n
    # rigid rotation about the centre, u = -omega (y - 1/2), v = omega (x - 1/2) on the MAC faces
    omega = 2.0
    yu = (np.arange(n) + 0.5) * h
    xv = (np.arange(n) + 0.5) * h
    u = np.repeat((-omega * (yu - 0.5))[:, None], n + 1, axis=1).astype(np.float32)
    v = np.repeat((omega * (xv - 0.5))[None, :], n + 1, axis=0).astype(np.float32)
    x = (np.arange(n) + 0.5) * h
    X, Y = np.meshgrid(x, x)
    disk = (np.sqrt((X - 0.5) ** 2 + (Y - 0.72) ** 2) - 0.15)
    notch = np.maximum(np.abs(X - 0.5) - 0.025, Y - 0.80) * -1.0
    rho = np.maximum(disk, notch).astype(np.float32)          # signed distance-like level set
    ref = ref2d.Ref2D(n, n, L, 1.0)
    ref.set_levelset(True)
    gpu = BimocqAdvection2D(n, n, ref.h, 1.0)
    gpu.set_levelset(True)
    for m, a in (("u", u), ("v", v), ("rho", rho), ("rho_init", rho)):
        ref.field(m)[...] = a
    dt = 0.01
    worst = 0.0
    for frame in range(10):
        _sync(ref, gpu)
        ref.phase_a(dt, frame)
        gpu.advect(frame, dt)
        e = rel_linf(gpu.download("RHO")[1:-1], ref.field("rho")[1:-1])
        worst = max(worst, e)
        assert e <= TOL, (frame, e)
        assert np.array_equal(gpu.download("U"), ref.field("u"))       # velocity untouched in this mode
        adv = [ref.field(m).copy() for m in ("u", "v", "rho", "temperature")]
        for nme, a in zip(("U", "V", "RHO", "T", "U_SAVE", "V_SAVE", "RHO_SAVE", "T_SAVE"), adv + adv):
            gpu.upload(nme, a)
        ref.phase_b(dt, frame, adv[0], adv[1], adv[0], adv[1], adv[2], adv[3])
        gpu.accumulate_host(frame, dt, adv[0], adv[1], adv[0], adv[1], adv[2], adv[3])
        rc, gs = ref.counters(), gpu.stats()
        assert (gs["vel_remap"], gs["scalar_remap"]) == (rc["vel_remap"], rc["scalar_remap"])
        for member in ("rho_init", "drho", "backward_scalar_x", "forward_scalar_y", "u"):
            assert rel_linf(gpu.download(ref2d.MEMBERS[member]), ref.field(member)) <= TOL, (frame, member)
    # analytic: the zero level set has been rotated by omega * 10 dt about the centre
    ang = omega * 10 * dt
    got = gpu.download("RHO")
    Xr = 0.5 + (X - 0.5) * np.cos(-ang) - (Y - 0.5) * np.sin(-ang)
    Yr = 0.5 + (X - 0.5) * np.sin(-ang) + (Y - 0.5) * np.cos(-ang)
    disk_r = np.sqrt((Xr - 0.5) ** 2 + (Yr - 0.72) ** 2) - 0.15
    inside_far = disk_r < -0.06
    outside_far = disk_r > 0.06
    notch_r = (np.abs(Xr - 0.5) < 0.03) & (Yr < 0.81)
    assert (got[inside_far & ~notch_r & (np.abs(Xr - 0.5) > 0.06)] < 0).all()
    assert (got[outside_far] > 0).all()
    print(f"2D level-set mode: worst per-step rel Linf vs the reference {worst:.2e}")
    ref.close(); gpu.close()
