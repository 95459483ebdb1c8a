"""Pins the CPU oracle against golden outputs of THE REFERENCE'S OWN CUDA KERNELS
(tests/golden/ref3d_kernels.npz, produced on a B200 by tests/golden/make_golden_3d.py from
oracle/_ref/libref3d.so).  The inputs are regenerated from their seed (the generator's Case3D),
the outputs must match: bit-for-bit tolerance 2e-7 (rel L-inf) for every kernel except the DMC
update, where glibc's and CUDA's expf differ in the last ulp (see tests/test_kernels_gpu.py)."""
import importlib.util
import os

import numpy as np
import pytest

from helpers import Case3D, rel_linf

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "ref3d_kernels.npz")


def _gen():
    spec = importlib.util.spec_from_file_location("make_golden_3d", os.path.join(HERE, "golden", "make_golden_3d.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def setup(oracle):
    if not os.path.exists(GOLD):
        pytest.skip("golden file not generated yet")
    m = _gen()
    c = Case3D(m.NI, m.NJ, m.NK, m.H, seed=11)
    return m, c, np.load(GOLD)


def f32_(x):
    return float(np.float32(x))


def _run_oracle(o3, name, arrays, c):
    A = [o3.padded_copy(a) for a in arrays]
    d = (c.h, c.ni, c.nj, c.nk)
    if name == "gpu_solve_forward":
        o3.gpu_solve_forward(*A, *d, c.cfldt, c.dt)
    elif name == "gpu_solve_backwardDMC":
        o3.gpu_solve_backwardDMC(*A, *d, 0.7 * c.cfldt)
    elif name == "gpu_semilag":
        o3.gpu_semilag(*A, 0, 1, 0, *d, c.cfldt, -c.dt)
    elif name == "gpu_advect_velocity":
        for q, k in enumerate("uvw"):
            o3.advect(A[q], A[3 + q], A[6], A[7], A[8], *d, k)
    elif name == "gpu_advect_field":
        o3.advect(A[0], A[1], A[2], A[3], A[4], *d, "c")
    elif name == "gpu_compensate_velocity":
        for q, k in enumerate("uvw"):
            o3.gpu_compensate(A[q], A[3 + q], A[6 + q], A[9:12], A[12:15], *d, k)
    elif name == "gpu_compensate_field":
        o3.gpu_compensate(A[0], A[1], A[2], A[3:6], A[6:9], *d, "c")
    elif name == "gpu_accumulate_velocity":
        for q, k in enumerate("uvw"):
            o3.cumulate(A[q], A[3 + q], A[6:9], *d, k, 2.0)
    elif name == "gpu_accumulate_field":
        o3.cumulate(A[0], A[1], A[2:5], *d, "c", 1.0)
    elif name == "gpu_advect_vel_double":
        for q, k in enumerate("uvw"):
            o3.double_advect(A[q], A[3 + q], A[6:9], A[9:12], *d, k, 0.5)
    elif name == "gpu_advect_field_double":
        o3.double_advect(A[0], A[1], A[2:5], A[5:8], *d, "c", 0.25)
    elif name == "gpu_estimate_distortion":
        o3.estimate(A[0], A[1:4], A[4:7], *d)
    elif name == "gpu_emit_smoke":
        f32 = lambda x: float(np.float32(x))
        o3.gpu_emit_smoke(*A[:5], c.h, c.ni, c.nj, c.nk, f32(0.45 * c.h * c.ni), f32(0.4 * c.h * c.nj), f32(0.5 * c.h * c.nk),
                          f32(0.22 * c.h * c.ni), 1.0, 50.0, 1.0)
    elif name == "gpu_add_buoyancy":
        big = [o3.padded((c.nk, c.nj + 1, c.ni)) for _ in range(2)]   # indexed with the v-face index by the reference
        for b, src in zip(big, A[1:3]):
            b.reshape(-1)[:src.size] = src.reshape(-1)
        o3.gpu_add_buoyancy(A[0], big[0], big[1], c.ni, c.nj, c.nk, f32_(0.3), f32_(0.7), f32_(0.02))
    elif name == "gpu_diffuse_field":
        o3.gpu_diffuse_field(A[0], A[1], A[2], c.ni, c.nj, c.nk, 4, f32_(0.37))
    elif name == "gpu_mad":
        o3.gpu_mad(A[0], A[1], A[2], 0.75, -1.25)
    else:
        raise KeyError(name)
    return A


def test_oracle_matches_reference_kernel_golden_vectors(setup, oracle):
    m, c, gold = setup
    seen = 0
    for name, (args, outs) in m.calls(c).items():
        arrays = [a for a in args if isinstance(a, np.ndarray)]
        A = _run_oracle(oracle, name, arrays, c)
        # DMC: expf ulp amplification; emitter: acosf/cosf/hypotf of glibc vs CUDA
        tol = 5e-4 if name == "gpu_solve_backwardDMC" else 2e-7
        for q in outs:
            want = gold[f"{name}:out{q}"]
            err = rel_linf(A[q], want)
            assert err <= tol, (name, q, err)
            seen += 1
    assert seen == len(gold.files)


def test_projection_oracle_matches_reference_golden(oracle):
    """oracle/projection_oracle.c against the reference's gpu_multi_grid_conjugate_gradient run on a
    B200 (tests/golden/ref3d_projection.npz): fp64/fp32 results bit for bit."""
    path = os.path.join(HERE, "golden", "ref3d_projection.npz")
    if not os.path.exists(path):
        pytest.skip("golden projection vectors not generated yet")
    from test_projection_gpu import velocity

    g = np.load(path)
    ni, nj, nk, levels, iters = (int(x) for x in g["case"])
    u, v, w = velocity(ni, nj, nk)
    out = oracle.gpu_multi_grid_conjugate_gradient(u, v, w, levels=levels, iters=iters, halfrdx=0.5)
    assert np.array_equal(out["p"], g["p"])
    for name, a in (("u", u), ("v", v), ("w", w)):
        assert np.array_equal(a, g[name]), name
    assert np.array_equal(out["result"][: 2 * iters + 3], g["result"][: 2 * iters + 3])
    assert np.array_equal(out["result"][2000:2001 + iters], g["result"][2000:2001 + iters])


def test_blocked_layout_restatement_matches_reference_container(oracle):
    """numpy restatement of Buffer3D's block indexing against storage written by the reference's own
    class (tests/golden/ref_blocked_layout.npz, made on CPU by tests/golden/make_golden_blocked.py)."""
    g = np.load(os.path.join(HERE, "golden", "ref_blocked_layout.npz"))
    for nx, ny, nz in g["shapes"]:
        nx, ny, nz = int(nx), int(ny), int(nz)
        lin = (np.arange(nx * ny * nz, dtype=np.float32) + 1).reshape(nz, ny, nx)
        want = g[f"blocked_{nx}x{ny}x{nz}"]
        assert np.array_equal(oracle.linear_to_blocked(lin), want)
        assert np.array_equal(oracle.blocked_to_linear(want, nx, ny, nz), lin)
