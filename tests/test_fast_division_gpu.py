"""Division by a cell size that is not a power of two (the reference's shipped scene: h = 0.2 / ni,
bimocq3D/main.cpp:36-38).  The kernels replace the IEEE division pos / h by a three-instruction sequence that
the library verifies exhaustively per h on the device (include/bimocq_b200.h: bmq_set_fast_division).  Here:
the verification accepts the cell sizes the tests and the benchmark use, and whole steps with the fast
sequence are bit-identical to the same steps with IEEE division."""
import numpy as np
import pytest

from gpufluidsimulation_b200 import load_library, scenes, zslab

pytestmark = pytest.mark.gpu
CHECK = zslab.CUR + zslab.INIT + zslab.MAPS_BWD + zslab.MAPS_FWD


def test_verification_accepts_the_usual_cell_sizes(cuda):
    lib = load_library()
    for n in (37, 48, 128, 256, 512):
        h = float(np.float32(0.2) / np.float32(n))
        assert lib.bmq_division_is_fast(h, n) == 1, (n, h)
    assert lib.bmq_division_is_fast(1.0 / 64, 64) == 1          # power of two: exact multiplication, no division
    lib.bmq_set_fast_division(0)
    try:
        assert lib.bmq_division_is_fast(float(np.float32(0.2) / np.float32(37)), 37) == 0
        assert lib.bmq_division_is_fast(1.0 / 64, 64) == 1
    finally:
        lib.bmq_set_fast_division(1)


def _run(fast, dims, L, frames, dt, variant):
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    lib = load_library()
    lib.bmq_set_fast_division(fast)
    lib.bmq_set_gather_variant(variant)
    try:
        ni, nj, nk = dims
        h = L / ni
        u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, L)
        u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
        s = BimocqAdvection3D(ni, nj, nk, h, 0.5)        # grid (and its division mode) fixed at creation
        s.set_initial(u, v, w, rho, T)
        for frame in range(frames):
            s.advect(frame, dt, with_semilag=True)
            s.apply_buoyancy(0.2, dt)
            s.accumulate(frame, dt)
        out = {n: s.download(n) for n in CHECK + ("U_SEMI", "RHO_SEMI")}
        out["stats"] = s.stats()
        s.close()
        return out
    finally:
        lib.bmq_set_fast_division(1)
        lib.bmq_set_gather_variant(1)


@pytest.mark.parametrize("dims,L,variant", [((37, 41, 35), 0.2, 1), ((37, 41, 35), 0.2, 0), ((128, 128, 24), 0.2, 1)])
def test_fast_division_is_bit_identical_to_ieee_division(cuda, dims, L, variant):
    a = _run(0, dims, L, 12, 0.02, variant)
    b = _run(1, dims, L, 12, 0.02, variant)
    assert a["stats"]["vel_reinit_count"] == b["stats"]["vel_reinit_count"] >= 1     # DMC, forward trace, gathers, reinit all ran
    for name in CHECK + ("U_SEMI", "RHO_SEMI"):
        assert np.array_equal(a[name], b[name]), name
