"""CPU checks of the oracle itself (test infrastructure): analytic known answers, invariants and
the scheduler restatement.  Golden-vector pinning lives in test_golden_cpu.py."""
import numpy as np
import pytest

from gpufluidsimulation_b200 import scenes


def _linear(shape, h, kind, o3):
    dx, dy, dz = o3.DIMS[kind]
    nz, ny, nx = shape
    z, y, x = np.meshgrid((np.arange(nz) - 0.5 * dz) * h, (np.arange(ny) - 0.5 * dy) * h,
                          (np.arange(nx) - 0.5 * dx) * h, indexing="ij")
    return o3.padded_copy((0.3 * x - 0.7 * y + 1.1 * z + 0.25).astype(np.float32))


@pytest.mark.parametrize("kind", ["u", "v", "w", "c"])
def test_identity_maps_reproduce_linear_fields(oracle, kind):
    o3 = oracle
    ni, nj, nk, h = 14, 12, 13, 0.125
    f0 = _linear(o3.shape_of(ni, nj, nk, kind), h, kind, o3)
    ident = o3.identity_maps(ni, nj, nk, h)
    f = o3.padded(f0.shape)
    o3.advect(f, f0, *ident, h, ni, nj, nk, kind)
    dx, dy, dz = o3.DIMS[kind]
    inner = (slice(3 + dz, f.shape[0] - 3), slice(3 + dy, f.shape[1] - 3), slice(3 + dx, f.shape[2] - 3))
    assert np.abs(f[inner] - f0[inner]).max() < 2e-6
    outer = f.copy()
    outer[inner] = 0
    assert not outer.any(), "advect_kernel must leave the guard ring untouched (GPU_kernel.cu:341)"


def test_uniform_translation_maps(oracle):
    """Constant velocity U: psi = x + U dt and chi = x - U dt in the interior (a = 0 -> Euler branch
    of the DMC update, GPU_kernel.cu:194), and the round-trip distortion vanishes."""
    o3 = oracle
    ni, nj, nk, h = 18, 16, 17, 0.0625
    U = (0.4, -0.3, 0.2)
    u = o3.padded(o3.shape_of(ni, nj, nk, "u"), U[0])
    v = o3.padded(o3.shape_of(ni, nj, nk, "v"), U[1])
    w = o3.padded(o3.shape_of(ni, nj, nk, "w"), U[2])
    dt = 0.1
    cfldt = h / 0.4
    m = o3.Mapper(ni, nj, nk, h, 1.0)
    n = m.update_mapping(u, v, w, cfldt, dt)
    assert n == 1
    ident = o3.identity_maps(ni, nj, nk, h)
    s = (slice(4, -4),) * 3
    for c in range(3):
        assert np.abs(m.fwd[c][s] - (ident[c][s] + U[c] * dt)).max() < 1e-6
        assert np.abs(m.bwd[c][s] - (ident[c][s] - U[c] * dt)).max() < 1e-6
    d = o3.padded((nk, nj, ni))
    o3.estimate(d, m.bwd, m.fwd, h, ni, nj, nk)
    assert np.sqrt(d[s].max()) < 1e-6


def test_substep_sequence_matches_reference_loop(oracle):
    """MapperBase::updateBackward (Mapping.cpp:13-20) in float arithmetic: cfldt sub-steps, then
    the remainder."""
    o3 = oracle
    ni = nj = nk = 12
    h = 1.0 / 12
    m = o3.Mapper(ni, nj, nk, h, 1.0)
    z = lambda k: o3.padded(o3.shape_of(ni, nj, nk, k))
    assert m.update_backward(z("u"), z("v"), z("w"), 0.04, 0.1) == 3
    assert m.update_backward(z("u"), z("v"), z("w"), 0.05, 0.1) == 2
    assert m.update_backward(z("u"), z("v"), z("w"), 1.0, 0.1) == 1


def test_clamp_extrema_bounds(oracle):
    o3 = oracle
    rng = np.random.default_rng(3)
    before = o3.padded_copy(rng.standard_normal((9, 10, 11)).astype(np.float32))
    after = o3.padded_copy((before + 5 * rng.standard_normal(before.shape)).astype(np.float32))
    ref = after.copy()
    o3.clamp_extrema(before, after)
    for (k, j, i) in [(1, 1, 1), (4, 5, 6), (7, 8, 9)]:
        nb = before[k - 1:k + 2, j - 1:j + 2, i - 1:i + 2]
        assert after[k, j, i] == np.float32(min(max(nb.min(), ref[k, j, i]), nb.max()))
    assert np.array_equal(after[0], ref[0]) and np.array_equal(after[:, 0], ref[:, 0])


def _run_oracle(o3, s, frames, dt, nj, beta=0.1):
    log = []
    for frame in range(frames):
        s.advect(frame, dt)
        forced = [a.copy() for a in s.cur[:3]]
        forced[1] = forced[1] + scenes.buoyancy_increment(s.cur[3], s.cur[4], 0.0, beta, dt, nj + 1)
        final = forced + [s.cur[3].copy(), s.cur[4].copy()]
        s.accumulate(frame, dt, forced, final)
        log.append(dict(s.stats))
        for a in s.cur:
            assert np.isfinite(a).all()
    return log


def test_scheduler_frame_cap(oracle):
    """BimocqSolver.cpp:175-185 with a distortion-free flow (fluid at rest): the velocity
    maps are reinitialised by the frame cap alone (framenum - last > 10 -> frames 11 and 22) and
    the scalar maps (cap 30) never within 24 frames."""
    o3 = oracle
    ni, nj, nk = 16, 16, 16
    h = 1.0 / ni
    shp = lambda k: o3.shape_of(ni, nj, nk, k)
    u = np.zeros(shp("u"), np.float32); v = np.zeros(shp("v"), np.float32); w = np.zeros(shp("w"), np.float32)
    rho = scenes.smooth_ball(ni, nj, nk, 1.0, (0.5, 0.5, 0.5), 0.2)
    s = o3.Solver(ni, nj, nk, h)
    s.set_initial(u, v, w, rho, rho.copy())
    log = _run_oracle(o3, s, 24, 0.05, nj, beta=0.0)
    assert [f for f, st in enumerate(log) if st["vel_reinit"]] == [11, 22]
    assert not any(st["scalar_reinit"] for st in log)
    assert log[-1]["vel_reinit_count"] == 2


def test_scheduler_distortion_trigger(oracle):
    """Plume scene on a coarse grid: reinitialisations follow the reference's rule exactly --
    frame 0 always trips it because max_v is forced to h there (BimocqSolver.cpp:94), afterwards
    velocity reinit <=> distortion/(max_v dt) > 1 or 11 frames since the last one; scalars use 5 / 31."""
    o3 = oracle
    ni, nj, nk = 20, 24, 20
    h = 1.0 / ni
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, 1.0)
    dt = 0.02
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 0.4)
    s = o3.Solver(ni, nj, nk, h)
    s.set_initial(u, v, w, rho, T)
    log = _run_oracle(o3, s, 8, dt, nj)
    assert log[0]["vel_reinit"] and log[0]["scalar_reinit"]
    last_v = last_s = 0
    for f, st in enumerate(log):
        assert st["vel_reinit"] == (st["vel_distortion"] > 1.0 or f - last_v > 10)
        assert st["scalar_reinit"] == (st["scalar_distortion"] > 5.0 or f - last_s > 30)
        last_v = f if st["vel_reinit"] else last_v
        last_s = f if st["scalar_reinit"] else last_s
    assert sum(st["vel_reinit"] for st in log) >= 2


def test_dmc_formula_amplifies_last_ulp_velocity_differences(oracle):
    """Why only bit-identical upstream arithmetic reproduces the reference's maps ("What exactness
    costs"): the DMC update evaluates 1 - exp(-a s) in fp32 (GPU_kernel.cu:194-196), which cancels for the usual
    a s << 1.  Here the oracle's DMC kernel runs twice on the benchmark scene (h = 0.2 / 64, one CFL sub-step from
    identity maps), the second time with every velocity value moved by +-1 ulp: the back-traced points move by tens
    of micro-cells -- hundreds of times what a 1.2e-7 relative change of a <= 1 cell displacement would explain."""
    from gpufluidsimulation_b200 import scenes
    ni, nj, nk, L, dt = 48, 40, 56, 0.2, 0.02
    h = float(np.float32(L / ni))
    u, v, w, _, _ = scenes.smoke_plume(ni, nj, nk, L)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    sub = float(np.float32(h) / np.float32(max(float(np.abs(a).max()) for a in (u, v, w))))

    def run(vel):
        U, V, W = (oracle.padded_copy(a) for a in vel)
        ident = oracle.identity_maps(ni, nj, nk, h)
        out = [oracle.padded(a.shape) for a in ident]
        oracle.gpu_solve_backwardDMC(U, V, W, *ident, *out, h, ni, nj, nk, sub)
        return out

    rng = np.random.default_rng(7)

    def one_ulp(a):
        up = rng.integers(0, 2, a.shape).astype(bool)
        return np.where(up, np.nextafter(a, np.float32(np.inf)), np.nextafter(a, np.float32(-np.inf))).astype(np.float32)

    base, moved = run((u, v, w)), run((one_ulp(u), one_ulp(v), one_ulp(w)))
    worst_cells = max(float(np.abs(a - b).max()) for a, b in zip(base, moved)) / h
    plain = 1.2e-7 * 1.0          # a displacement of at most one cell, changed by one ulp of the velocity
    assert worst_cells > 50 * plain, worst_cells          # amplified (measured: ~2e-5 .. 5e-5 cells)
    assert worst_cells < 1e-3, worst_cells                # ... but not broken
