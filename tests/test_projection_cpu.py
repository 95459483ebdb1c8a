"""CPU-side checks of the projection row: host logic of
gpufluidsimulation_b200.projection, the ABI struct layout, and properties of the CPU oracle
(oracle/projection_oracle.c; its bit-level pin against the reference's GPU output is
tests/test_golden_cpu.py::test_projection_oracle_matches_reference_golden)."""
import ctypes as C

import numpy as np
import pytest

from gpufluidsimulation_b200 import capi, projection


def test_level_table_follows_the_reference_rule():
    # BimocqGPUSolver.cpp:77-82: n_{l+1} = (n_l - 1) / 2, LEVEL_COUNT = 6
    assert projection.level_dims(512, 512, 512) == [(512,) * 3, (255,) * 3, (127,) * 3, (63,) * 3, (31,) * 3, (15,) * 3]
    assert projection.level_dims(48, 64, 56, 4) == [(48, 64, 56), (23, 31, 27), (11, 15, 13), (5, 7, 6)]
    assert projection.max_levels(48, 64, 56) == 4          # the next level would be 2 x 3 x 2
    assert projection.max_levels(512, 512, 512) == 6       # capped at LEVEL_COUNT
    assert projection.max_levels(5, 5, 5) == 1


def test_coarse_level_struct_has_the_reference_layout():
    # SCoarseLevelInfo (GPU_Advection.h:13-24): 4 ints, 2 doubles, 3 pointers, natural alignment
    assert C.sizeof(capi.CoarseLevel) == 56
    offs = {n: getattr(capi.CoarseLevel, n).offset for n, _ in capi.CoarseLevel._fields_}
    assert offs == {"ni": 0, "nj": 4, "nk": 8, "number": 12, "alpha": 16, "beta": 24, "b": 32, "x": 40, "r": 48}


def _velocity(ni, nj, nk, seed=1):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal(s).astype(np.float32) for s in ((nk, nj, ni + 1), (nk, nj + 1, ni), (nk + 1, nj, ni))]


def test_oracle_zero_velocity_gives_zero_pressure(oracle):
    ni, nj, nk = 17, 15, 19
    u, v, w = [np.zeros_like(a) for a in _velocity(ni, nj, nk)]
    out = oracle.gpu_multi_grid_conjugate_gradient(u, v, w, levels=2, iters=0)
    assert not out["p"].any() and not out["div"].any() and not u.any()


def test_oracle_pressure_equation_and_convergence(oracle):
    """After the solve: residual = div - lap p on interior cells (the quantity the solver tracks), the
    residual shrinks, and the velocity update is u -= halfrdx * grad p on the cells the reference's
    gradient kernel touches (GPU_kernel.cu:1003-1021: indices 2 .. n-1)."""
    ni, nj, nk, iters = 31, 27, 23, 6
    u, v, w = _velocity(ni, nj, nk)
    u0, v0, w0 = u.copy(), v.copy(), w.copy()
    out = oracle.gpu_multi_grid_conjugate_gradient(u, v, w, levels=3, iters=iters)
    p, div, r = out["p"], out["div"], out["residual"]
    lap = (p[1:-1, 1:-1, :-2] + p[1:-1, 1:-1, 2:] + p[1:-1, :-2, 1:-1] + p[1:-1, 2:, 1:-1] + p[:-2, 1:-1, 1:-1] + p[2:, 1:-1, 1:-1]
           - 6 * p[1:-1, 1:-1, 1:-1])
    assert np.allclose(r[1:-1, 1:-1, 1:-1], div[1:-1, 1:-1, 1:-1] - lap, rtol=0, atol=1e-12)
    hist = out["result"][2000:2001 + iters]
    assert hist[-1] < 0.05 * hist[0], hist
    want = u0.copy()
    want[2:, 2:, 2:ni] -= (0.5 * (p[2:, 2:, 2:ni] - p[2:, 2:, 1:ni - 1])).astype(np.float32)
    assert np.array_equal(u, want)
    assert np.array_equal(v[:, :2, :], v0[:, :2, :]) and np.array_equal(w[:2], w0[:2])      # untouched rings


def test_oracle_rejects_too_many_levels(oracle):
    u, v, w = _velocity(9, 9, 9)
    with pytest.raises(ValueError):
        oracle.gpu_multi_grid_conjugate_gradient(u, v, w, levels=3, iters=1)     # 9 -> 4 -> 1
