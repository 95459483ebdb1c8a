"""the reference's 8^3-blocked host containers (buffer3Df) accepted as they are.
Device relayout kernels against golden storage produced by the reference's own Buffer3D class
(tests/golden/ref_blocked_layout.npz) and against the numpy restatement; handle API with blocked
host buffers against the same run with dense buffers.  Pure permutations: bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_blocked_layout.npz")


def relayout(lib, fn, src, n_out, nx, ny, nz):
    import torch

    F = C.POINTER(C.c_float)
    s = torch.from_numpy(np.ascontiguousarray(src, dtype=np.float32).ravel()).cuda()
    d = torch.full((n_out,), -7.0, dtype=torch.float32, device="cuda")
    st = getattr(lib, fn)(C.cast(C.c_void_p(s.data_ptr()), F), C.cast(C.c_void_p(d.data_ptr()), F), nx, ny, nz, None)
    assert st == 0
    torch.cuda.synchronize()
    return d.cpu().numpy()


def test_relayout_kernels_match_reference_container():
    from gpufluidsimulation_b200 import capi

    lib = capi.load_library()
    g = np.load(GOLD)
    for nx, ny, nz in g["shapes"]:
        nx, ny, nz = int(nx), int(ny), int(nz)
        want = g[f"blocked_{nx}x{ny}x{nz}"]
        assert lib.bmq_blocked_elems(nx, ny, nz) == want.size
        lin = (np.arange(nx * ny * nz, dtype=np.float32) + 1).reshape(nz, ny, nx)
        got = relayout(lib, "bmq_linear_to_blocked", lin, want.size, nx, ny, nz)
        assert np.array_equal(got, want), (nx, ny, nz)
        back = relayout(lib, "bmq_blocked_to_linear", want, lin.size, nx, ny, nz)
        assert np.array_equal(back.reshape(nz, ny, nx), lin), (nx, ny, nz)


@pytest.mark.parametrize("shape", [(65, 33, 40), (128, 128, 129)])
def test_relayout_kernels_match_numpy_restatement(shape):
    from gpufluidsimulation_b200 import capi
    from oracle import oracle3d

    lib = capi.load_library()
    nx, ny, nz = shape
    lin = np.random.default_rng(5).standard_normal((nz, ny, nx)).astype(np.float32)
    want = oracle3d.linear_to_blocked(lin)
    got = relayout(lib, "bmq_linear_to_blocked", lin, want.size, nx, ny, nz)
    assert np.array_equal(got, want)
    back = relayout(lib, "bmq_blocked_to_linear", want, lin.size, nx, ny, nz).reshape(nz, ny, nx)
    assert np.array_equal(back, lin)


def test_host_path_with_blocked_buffers_equals_dense_path():
    """bmq3d_advect_host / bmq3d_accumulate_host fed with buffer3Df storage give the same fields, in the
    same storage, as the dense-buffer path (BimocqSolver's host-orchestrated step)."""
    from gpufluidsimulation_b200 import scenes
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    from oracle import oracle3d

    ni, nj, nk, dt = 40, 36, 44, 0.02
    h = 1.0 / ni
    u, v, w, rho, T = [np.ascontiguousarray(a, dtype=np.float32) for a in scenes.smoke_plume(ni, nj, nk, 1.0)]
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    dense, blocked = BimocqAdvection3D(ni, nj, nk, h, 1.0), BimocqAdvection3D(ni, nj, nk, h, 1.0)
    dense.set_initial(u, v, w, rho, T)
    blocked.set_host_layout(True)
    blocked.set_initial(*[oracle3d.linear_to_blocked(a) for a in (u, v, w, rho, T)])
    fd = [a.copy() for a in (u, v, w, rho, T)]
    fb = [oracle3d.linear_to_blocked(a) for a in (u, v, w, rho, T)]
    shapes = [a.shape for a in fd]
    for frame in range(4):
        dense.advect_host(frame, dt, *fd)
        blocked.advect_host(frame, dt, *fb)
        for a, b, shp in zip(fd, fb, shapes):
            assert np.array_equal(oracle3d.blocked_to_linear(b, shp[2], shp[1], shp[0]), a), f"frame {frame} after advect"
        forced = [a.copy() for a in fd[:3]]
        forced[1] += scenes.buoyancy_increment(fd[3], fd[4], 0.1, 0.3, dt, nj + 1)
        final = [0.97 * f for f in forced] + [fd[3], fd[4]]
        dense.accumulate_host(frame, dt, forced, final)
        blocked.accumulate_host(frame, dt, [oracle3d.linear_to_blocked(a) for a in forced],
                                [oracle3d.linear_to_blocked(np.ascontiguousarray(a)) for a in final])
        for a, f in zip(fd, final):
            a[...] = f
        for b, f in zip(fb, final):
            b[...] = oracle3d.linear_to_blocked(np.ascontiguousarray(f))
        for name in ("U_INIT", "W_INIT", "RHO_INIT", "DV_EXT", "V", "VBWD_X", "SFWD_Z"):
            a, b = dense.download(name), blocked.download(name)
            nz_, ny_, nx_ = a.shape
            assert np.array_equal(oracle3d.blocked_to_linear(b, nx_, ny_, nz_), a), f"frame {frame}: {name}"
        assert dense.stats() == blocked.stats()
    dense.close(); blocked.close()
