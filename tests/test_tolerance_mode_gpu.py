"""Opt-in tolerance mode (bmq_set_tolerance_mode): with a cell size that is not a power of two every kernel takes the
path written for a power of two (one multiplication by RN(1/h) instead of the correctly rounded division, grid-unit
positions, node shortcuts, fp32 lerps in the DMC update).  Results are no longer bit-identical to the reference; this
test states by how much they differ -- per step from identical state and over a free-running sequence -- and that the
mode changes nothing when h IS a power of two."""
import numpy as np
import pytest

from gpufluidsimulation_b200 import load_library, scenes

pytestmark = pytest.mark.gpu
FIELDS = ("U", "V", "W", "RHO", "T")
# Stated tolerance of the mode (relative L-inf against the exact path), measured on B200:
#   one step from identical state: 3e-6 .. 5e-6 in the fields, 5e-7 in the maps -- rounding noise, as intended;
#   from the second step on: up to 1.3e-3 in the velocity, 5e-5 in the backward maps (3e-3 cells).  That is not the
#   mode's arithmetic but the reference's DMC formula, 1 - exp(-a s) in fp32 (GPU_kernel.cu:194-196): it cancels 3-4
#   digits for the usual a s << 1 and amplifies relative differences of its velocity input several hundred times
#   (tests/test_oracle_cpu.py measures +-1 ulp in -> up to 4.5e-5 cells out), so the 3-5e-6 differences after the
#   first step become ~1e-3 cells in the second step's maps.  Only bit-identical upstream arithmetic reproduces the
#   reference to 1e-5 -- which is what the default (exact) path pays for.
TOL_PER_STEP = 2e-5
TOL_RUN = 5e-3


@pytest.fixture
def tolerance_lib(cuda):
    lib = load_library()
    yield lib
    lib.bmq_set_tolerance_mode(0)


def make(lib, ni, nj, nk, L, mode, blend=1.0):
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    lib.bmq_set_tolerance_mode(mode)
    h, dt = L / ni, 0.02
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, L)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    s = BimocqAdvection3D(ni, nj, nk, h, blend)
    s.set_initial(u, v, w, rho, T)
    lib.bmq_set_tolerance_mode(0)
    return s, dt


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_power_of_two_cell_size_is_untouched(tolerance_lib):
    a, dt = make(tolerance_lib, 40, 36, 44, 1.25, 0)       # h = 2^-5
    b, _ = make(tolerance_lib, 40, 36, 44, 1.25, 1)
    for frame in range(4):
        for s in (a, b):
            s.advect(frame, dt); s.apply_buoyancy(0.2, dt); s.accumulate(frame, dt)
    for n in FIELDS:
        assert np.array_equal(a.download(n), b.download(n)), n
    a.close(); b.close()


def test_general_cell_size_within_stated_tolerance(tolerance_lib):
    ni, nj, nk, L = 64, 56, 72, 0.2                        # the reference scene's h = 0.2 / ni
    exact, dt = make(tolerance_lib, ni, nj, nk, L, 0)
    fast, _ = make(tolerance_lib, ni, nj, nk, L, 1)
    worst_run, first_step = 0.0, None
    for frame in range(12):
        for s in (exact, fast):
            s.advect(frame, dt); s.apply_buoyancy(0.2, dt); s.accumulate(frame, dt)
        errs = {n: rel(fast.download(n), exact.download(n)) for n in FIELDS}
        if frame == 0:
            first_step = max(errs.values())          # one step from identical state
            assert first_step <= TOL_PER_STEP, errs
        worst_run = max(worst_run, max(errs.values()))
        assert max(errs.values()) <= TOL_RUN, (frame, errs)
        assert exact.stats()["vel_reinit"] == fast.stats()["vel_reinit"], frame
    print(f"tolerance mode, h = 0.2/64: rel L-inf after one step {first_step:.2e}, worst over 12 free-running steps {worst_run:.2e}")
    assert worst_run > 0.0          # the mode did take the other path
    exact.close(); fast.close()
