"""TEST INFRASTRUCTURE ONLY -- Python driver of the CPU oracle for the 3D BiMocq^2 hot path.

Binds oracle/liboracle3d.so (the plain-C restatement of /root/reference/src/bimocq3D/GPU_kernel.cu,
see bimocq3d_oracle.c) and restates, on numpy arrays, the call sequences of

* gpuMapper / the extern "C" wrappers        -- bimocq3D/GPU_kernel.cu:567-734, GPU_Advection.h:328-424
* MapperBase (host-orchestrated advector)    -- bimocq3D/Mapping.cpp:7-271
* BimocqSolver::advanceBimocq, advection part -- bimocq3D/BimocqSolver.cpp:88-230, getCFL :1067-1118,
  velocityReinitialize / scalarReinitialize   :1433-1451

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (gpufluidsimulation_b200) never does.

Parity status: the reference has no tests or golden vectors; this oracle is pinned against the
reference's own CUDA kernels executed on a B200 (tests/test_kernels_gpu.py, oracle/_ref/libref3d.so).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_F = C.POINTER(C.c_float)


def build(force: bool = False) -> str:
    """Compile liboracle3d.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle3d.so")
    srcs = [os.path.join(_HERE, n) for n in ("bimocq3d_oracle.c", "projection_oracle.c")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(x) for x in srcs):
        try:
            subprocess.check_call(["make", "-C", _HERE, "liboracle3d.so"], stdout=subprocess.DEVNULL)
        except (OSError, subprocess.CalledProcessError):
            if not os.path.exists(so):      # a prebuilt library (shipped by build()) is good enough
                raise
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle3d.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        i, f = C.c_int, C.c_float
        L.o3_forward.argtypes = [_F] * 6 + [f, i, i, i, f, f, i, i]
        L.o3_clamp_extrema.argtypes = [_F, _F, i, i, i, i, i]
        L.o3_dmc_backward.argtypes = [_F] * 9 + [f, i, i, i, f, i, i]
        L.o3_semilag.argtypes = [_F] * 5 + [i, i, i, f, i, i, i, f, f, i, i]
        L.o3_advect.argtypes = [_F] * 5 + [f, i, i, i, i, i, i, i, i, i]
        L.o3_double_advect.argtypes = [_F] * 8 + [f, i, i, i, i, i, i, i, f, i, i]
        L.o3_cumulate.argtypes = [_F] * 5 + [f, i, i, i, i, i, i, i, f, i, i]
        L.o3_compensate.argtypes = [_F] * 6 + [f, i, i, i, i, i, i, i, i, i]
        L.o3_estimate.argtypes = [_F] * 7 + [f, i, i, i, i, i]
        L.o3_add.argtypes = [_F, _F, f, C.c_long]
        L.o3_add_field.argtypes = [_F, _F, _F, f, C.c_long]
        L.o3_maxabs.argtypes = [_F, C.c_long, f]
        L.o3_maxabs.restype = f
        L.o3_max_dist.argtypes = [_F, C.c_void_p, i, i, i]
        L.o3_max_dist.restype = f
        L.o3_emit_velocity.argtypes = [_F, f, i, i, i, f, f, f, f, f]
        L.o3_emit_field.argtypes = [_F, _F, f, i, i, i, f, f, f, f, f, f]
        L.o3_add_buoyancy.argtypes = [_F, _F, _F, i, i, i, f, f, f]
        L.o3_diffuse_sweep.argtypes = [_F, _F, _F, i, i, i, f]
        L.o3_mad.argtypes = [_F, _F, _F, f, f, C.c_long]
        _D = C.POINTER(C.c_double)
        L.o3_mgpcg.argtypes = [_F, _F, _F, _D, _D, _D, _D, _D, i, i, i, i, i, C.c_double]
        L.o3_mgpcg.restype = i
        _LIB = L
    return _LIB


class VirtualArray:
    """A float* that points `offset_elems` floats BEFORE a numpy array: a z-slab rank hands the
    oracle the address its plane 0 would have, so that all indices stay global (the CUDA
    kernels get the same kind of virtual base, gpufluidsimulation_b200/csrc/kernels3d.cu)."""

    def __init__(self, arr: np.ndarray, offset_elems: int):
        assert arr.dtype == np.float32 and arr.flags["C_CONTIGUOUS"]
        self.arr = arr
        self.vaddr = arr.ctypes.data - 4 * int(offset_elems)


def _p(a):
    if isinstance(a, VirtualArray):
        return C.cast(C.c_void_p(a.vaddr), _F)
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_F)


DIMS = {"u": (1, 0, 0), "v": (0, 1, 0), "w": (0, 0, 1), "c": (0, 0, 0)}


def shape_of(ni, nj, nk, kind):
    """numpy shape (z, y, x) of a field; memory order is the reference's idx = i + nx*j + nx*ny*k."""
    dx, dy, dz = DIMS[kind]
    return (nk + dz, nj + dy, ni + dx)


def padded(shape, fill=0.0):
    """A zeroed array with one plane + one row + 2 floats of slack after it: the reference's
    sampler reads one node past its clamp bound with weight 0 (GPU_kernel.cu:53-61)."""
    n = int(np.prod(shape))
    slack = shape[1] * shape[2] + shape[2] + 2
    buf = np.zeros(n + slack, dtype=np.float32)
    a = buf[:n].reshape(shape)
    if fill:
        a[...] = fill
    return a


def padded_copy(src):
    a = padded(src.shape)
    a[...] = src
    return a


# ----------------------------------------------------------------------------------------------
# extern "C" gpu_* equivalents on host arrays (GPU_kernel.cu:567-734)
# ----------------------------------------------------------------------------------------------
def gpu_solve_forward(u, v, w, xf, yf, zf, h, ni, nj, nk, cfldt, dt, krange=None):
    kb, ke = krange or (0, nk)
    lib().o3_forward(_p(u), _p(v), _p(w), _p(xf), _p(yf), _p(zf), h, ni, nj, nk, cfldt, dt, kb, ke)


def gpu_solve_backwardDMC(u, v, w, xi, yi, zi, xo, yo, zo, h, ni, nj, nk, substep, krange=None):
    kb, ke = krange or (0, nk)
    lib().o3_dmc_backward(_p(u), _p(v), _p(w), _p(xi), _p(yi), _p(zi), _p(xo), _p(yo), _p(zo), h, ni, nj, nk,
                          substep, kb, ke)


def gpu_semilag(field, src, u, v, w, dx, dy, dz, h, ni, nj, nk, cfldt, dt, krange=None):
    kb, ke = krange or (0, nk + dz)
    lib().o3_semilag(_p(field), _p(src), _p(u), _p(v), _p(w), dx, dy, dz, h, ni, nj, nk, cfldt, dt, kb, ke)


def advect(field, init, bx, by, bz, h, ni, nj, nk, kind, is_point=False, krange=None):
    dx, dy, dz = DIMS[kind]
    kb, ke = krange or (0, nk + dz)
    lib().o3_advect(_p(field), _p(init), _p(bx), _p(by), _p(bz), h, ni, nj, nk, dx, dy, dz, int(is_point), kb, ke)


def double_advect(field, prev, b, bp, h, ni, nj, nk, kind, blend, is_point=False, krange=None):
    dx, dy, dz = DIMS[kind]
    kb, ke = krange or (0, nk + dz)
    lib().o3_double_advect(_p(field), _p(prev), _p(b[0]), _p(b[1]), _p(b[2]), _p(bp[0]), _p(bp[1]), _p(bp[2]),
                           h, ni, nj, nk, dx, dy, dz, int(is_point), blend, kb, ke)


def cumulate(dfield, target, m, h, ni, nj, nk, kind, coeff, is_point=False, krange=None):
    dx, dy, dz = DIMS[kind]
    kb, ke = krange or (0, nk + dz)
    lib().o3_cumulate(_p(dfield), _p(target), _p(m[0]), _p(m[1]), _p(m[2]), h, ni, nj, nk, dx, dy, dz,
                      int(is_point), coeff, kb, ke)


def compensate_kernel(src, temp, test, m, h, ni, nj, nk, kind, is_point=False, krange=None):
    dx, dy, dz = DIMS[kind]
    kb, ke = krange or (0, nk + dz)
    lib().o3_compensate(_p(src), _p(temp), _p(test), _p(m[0]), _p(m[1]), _p(m[2]), h, ni, nj, nk, dx, dy, dz,
                        int(is_point), kb, ke)


def clamp_extrema(before, after, krange=None, dims=None):
    nz, ny, nx = dims or before.shape
    kb, ke = krange or (0, nz)
    lib().o3_clamp_extrema(_p(before), _p(after), nx, ny, nz, kb, ke)


def gpu_compensate(f, df, f_src, fwd, bwd, h, ni, nj, nk, kind, is_point=False):
    """One component of gpu_compensate_velocity / gpu_compensate_field (GPU_kernel.cu:640-682):
    f_src <- error at time 0, df <- pre-correction f, f <- compensated + clamped."""
    compensate_kernel(f, df, f_src, fwd, h, ni, nj, nk, kind, is_point)
    df[...] = f
    cumulate(f_src, f, bwd, h, ni, nj, nk, kind, -0.5, is_point)
    clamp_extrema(df, f)


def estimate(dist, bwd, fwd, h, ni, nj, nk, krange=None):
    kb, ke = krange or (0, nk)
    lib().o3_estimate(_p(dist), _p(bwd[0]), _p(bwd[1]), _p(bwd[2]), _p(fwd[0]), _p(fwd[1]), _p(fwd[2]), h, ni,
                      nj, nk, kb, ke)


def max_dist(dist, boundary=None):
    nz, ny, nx = dist.shape
    bp = None if boundary is None else boundary.ctypes.data_as(C.c_void_p)
    return float(lib().o3_max_dist(_p(dist), bp, nx, ny, nz))


def maxabs(a, start=0.0):
    return float(lib().o3_maxabs(_p(a), a.size, start))


# ---- source terms (GPU_kernel.cu:736-876, 952-964) -----------------------------------------
def gpu_emit_smoke(u, v, w, rho, T, h, ni, nj, nk, cx, cy, cz, radius, density, temperature, emiter):
    L = lib()
    L.o3_emit_velocity(_p(u), h, ni + 1, nj, nk, cx, cy, cz, radius, emiter)
    L.o3_emit_velocity(_p(v), h, ni, nj + 1, nk, cx, cy, cz, radius, 0.0)
    L.o3_emit_velocity(_p(w), h, ni, nj, nk + 1, cx, cy, cz, radius, 0.0)
    L.o3_emit_field(_p(rho), _p(T), h, ni, nj, nk, cx, cy, cz, radius, density, temperature)


def gpu_add_buoyancy(field, density, temperature, ni, nj, nk, alpha, beta, dt):
    lib().o3_add_buoyancy(_p(field), _p(density), _p(temperature), ni, nj + 1, nk, alpha, beta, dt)


def gpu_diffuse_field(field, tmp0, tmp1, ni, nj, nk, iters, coef):
    """gpu_diffuse_field, GPU_kernel.cu:855-876, including which buffer is copied back (:875)."""
    a, b = tmp0, tmp1
    a[...] = field
    for _ in range(iters):
        lib().o3_diffuse_sweep(_p(field), _p(a), _p(b), ni, nj, nk, coef)
        a, b = b, a
    field[...] = b


def gpu_mad(field, f1, f2, c1, c2):
    lib().o3_mad(_p(field), _p(f1), _p(f2), c1, c2, field.size)


def gpu_multi_grid_conjugate_gradient(u, v, w, levels=6, iters=50, halfrdx=0.5, div=None, residual=None, dir=None):
    """gpu_multi_grid_conjugate_gradient (GPU_kernel.cu:1784-1828) with the level table of
    BimocqGPUSolver.cpp:68-90.  u, v, w (float32 face fields, shapes (nk,nj,ni+1), (nk,nj+1,ni),
    (nk+1,nj,ni)) are projected in place; returns dict(p, div, residual, dir, result) of float64."""
    nk, nj, ni = u.shape[0], u.shape[1], u.shape[2] - 1
    assert v.shape == (nk, nj + 1, ni) and w.shape == (nk + 1, nj, ni)
    for a in (u, v, w):
        assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    z = lambda a: np.zeros((nk, nj, ni), np.float64) if a is None else np.ascontiguousarray(a, np.float64)
    out = dict(p=np.zeros((nk, nj, ni)), div=z(div), residual=z(residual), dir=z(dir), result=np.zeros(4096))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = lib().o3_mgpcg(_p(u), _p(v), _p(w), dp(out["div"]), dp(out["p"]), dp(out["dir"]), dp(out["residual"]),
                        dp(out["result"]), ni, nj, nk, levels, iters, halfrdx)
    if rc != 0:
        raise ValueError(f"o3_mgpcg: {levels} levels do not fit a {ni}x{nj}x{nk} grid")
    return out


def blocked_index(nx, ny, nz):
    """Storage index of cell (i,j,k) in the reference's 8^3-blocked Buffer3D (include/fluid_buffer3D.h:173-189),
    as an int64 array of shape (nz, ny, nx), and the physical element count (:56-75)."""
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    bx, by, bz = (nx + 7) // 8, (ny + 7) // 8, (nz + 7) // 8
    idx = ((((k >> 3) * bx * by + (j >> 3) * bx + (i >> 3)) << 9) + ((k & 7) << 6) + ((j & 7) << 3) + (i & 7)).astype(np.int64)
    return idx, bx * by * bz * 512


def linear_to_blocked(linear):
    """Buffer3D storage (padding = 0) holding the dense (nz, ny, nx) array `linear`."""
    nz, ny, nx = linear.shape
    idx, n = blocked_index(nx, ny, nz)
    out = np.zeros(n, dtype=linear.dtype)
    out[idx.ravel()] = linear.ravel()
    return out


def blocked_to_linear(blocked, nx, ny, nz):
    idx, n = blocked_index(nx, ny, nz)
    assert blocked.size == n
    return blocked[idx.ravel()].reshape(nz, ny, nx)


def identity_maps(ni, nj, nk, h):
    """Mapping.cpp:310-324: x = (float)i * h etc."""
    h32 = np.float32(h)
    x = padded((nk, nj, ni)); y = padded((nk, nj, ni)); z = padded((nk, nj, ni))
    x[...] = (np.arange(ni, dtype=np.float32) * h32)[None, None, :]
    y[...] = (np.arange(nj, dtype=np.float32) * h32)[None, :, None]
    z[...] = (np.arange(nk, dtype=np.float32) * h32)[:, None, None]
    return [x, y, z]


# ----------------------------------------------------------------------------------------------
# MapperBase (Mapping.cpp:7-271) on host arrays
# ----------------------------------------------------------------------------------------------
class Mapper:
    def __init__(self, ni, nj, nk, h, blend):
        self.ni, self.nj, self.nk, self.h, self.blend = ni, nj, nk, float(np.float32(h)), float(blend)
        self.total_reinit_count = 0
        self.fwd = identity_maps(ni, nj, nk, h)
        self.bwd = identity_maps(ni, nj, nk, h)
        self.bwd_prev = identity_maps(ni, nj, nk, h)
        self._tmp = identity_maps(ni, nj, nk, h)

    # Mapping.cpp:7-24 (+ gpuMapper::solveBackwardDMC, GPU_Advection.h:333-342)
    def update_backward(self, u, v, w, cfldt, dt):
        T = np.float32(0.0); substep = np.float32(cfldt); dt = np.float32(dt)
        n = 0
        while T < dt:
            if T + substep > dt:
                substep = np.float32(dt - T)
            gpu_solve_backwardDMC(u, v, w, *self.bwd, *self._tmp, self.h, self.ni, self.nj, self.nk, float(substep))
            for c in range(3):      # the reference copies x_out back into x_in; whole-array copy
                self.bwd[c][...] = self._tmp[c]
            T = np.float32(T + substep)
            n += 1
        return n

    # Mapping.cpp:26-36
    def update_forward(self, u, v, w, cfldt, dt):
        gpu_solve_forward(u, v, w, *self.fwd, self.h, self.ni, self.nj, self.nk, float(np.float32(cfldt)),
                          float(np.float32(dt)))

    def update_mapping(self, u, v, w, cfldt, dt):
        n = self.update_backward(u, v, w, cfldt, dt)
        self.update_forward(u, v, w, cfldt, dt)
        return n

    # Mapping.cpp:169-236 for one field; kind in u,v,w,c.  Returns the new field (zero outer ring,
    # from the cudaMemset in gpuMapper::advectVelocity, GPU_Advection.h:346-348).
    def advect_field(self, init, prev, kind):
        ni, nj, nk, h = self.ni, self.nj, self.nk, self.h
        f = padded(init.shape)
        advect(f, init, *self.bwd, h, ni, nj, nk, kind)
        df = padded_copy(init)           # gpu.du holds the init buffer (Mapping.cpp:182-184)
        f_src = padded(init.shape)       # cudaMemset in compensateVelocity (GPU_Advection.h:381-383)
        gpu_compensate(f, df, f_src, self.fwd, self.bwd, h, ni, nj, nk, kind)
        blend = self.blend if self.total_reinit_count != 0 else 1.0
        double_advect(f, prev, self.bwd, self.bwd_prev, h, ni, nj, nk, kind, blend)
        return f

    # Mapping.cpp:52-89
    def accumulate(self, init, change, kind, coeff):
        cumulate(change, init, self.fwd, self.h, self.ni, self.nj, self.nk, kind, coeff)

    # Mapping.cpp:91-118
    def estimate_distortion(self, boundary=None):
        d = padded((self.nk, self.nj, self.ni))
        estimate(d, self.bwd, self.fwd, self.h, self.ni, self.nj, self.nk)
        return max_dist(d, boundary)

    # Mapping.cpp:238-271
    def reinitialize(self):
        self.total_reinit_count += 1
        ident = identity_maps(self.ni, self.nj, self.nk, self.h)
        for c in range(3):
            self.bwd_prev[c][...] = self.bwd[c]
            self.bwd[c][...] = ident[c]
            self.fwd[c][...] = ident[c]


# ----------------------------------------------------------------------------------------------
# BimocqSolver::advanceBimocq, advection part (BimocqSolver.cpp:88-230)
# ----------------------------------------------------------------------------------------------
class Solver:
    """Host fields + two Mappers + the reinitialisation scheduler.  Forces and the projection are
    outside the hot path: the caller changes u,v,w,rho,T between advect() and accumulate()."""

    KINDS = ("u", "v", "w", "c", "c")
    NAMES = ("u", "v", "w", "rho", "T")

    def __init__(self, ni, nj, nk, h, blend=1.0):
        self.ni, self.nj, self.nk = ni, nj, nk
        self.h = float(np.float32(h))
        self.vel = Mapper(ni, nj, nk, h, blend)
        self.sca = Mapper(ni, nj, nk, h, blend)
        mk = lambda kind: padded(shape_of(ni, nj, nk, kind))
        self.cur = [mk(k) for k in self.KINDS]
        self.init = [mk(k) for k in self.KINDS]
        self.prev = [mk(k) for k in self.KINDS]
        self.semi = [mk(k) for k in self.KINDS]
        self.max_v = 0.0
        self.cfldt = 0.0
        self.vel_last_reinit = 0
        self.scalar_last_reinit = 0
        self.stats = {}
        self._adv = None

    def set_initial(self, u, v, w, rho, T):
        for dst, src in zip(self.cur, (u, v, w, rho, T)):
            dst[...] = src
        for c in range(5):
            self.init[c][...] = self.cur[c]
            self.prev[c][...] = self.cur[c]

    # BimocqSolver.cpp:1067-1118
    def get_cfl(self):
        m = 1e-4
        for c in range(3):
            m = maxabs(self.cur[c], m)
        self.max_v = float(np.float32(m))
        return float(np.float32(self.h) / np.float32(m))

    # BimocqSolver.cpp:90-126
    def advect(self, framenum, dt, with_semilag=False):
        ni, nj, nk, h = self.ni, self.nj, self.nk, self.h
        u, v, w, rho, T = self.cur
        cfldt = self.get_cfl()
        self.cfldt = cfldt
        if framenum == 0:
            self.max_v = h
        n = self.vel.update_mapping(u, v, w, cfldt, dt)
        self.sca.update_mapping(u, v, w, cfldt, dt)
        if with_semilag:   # BimocqSolver.cpp:645-668, traced with -dt
            for c, kind in enumerate(self.KINDS):
                dx, dy, dz = DIMS[kind]
                self.semi[c][...] = 0
                gpu_semilag(self.semi[c], self.cur[c], u, v, w, dx, dy, dz, h, ni, nj, nk, cfldt, -float(np.float32(dt)))
        new = [self.vel.advect_field(self.init[c], self.prev[c], self.KINDS[c]) for c in range(3)]
        new += [self.sca.advect_field(self.init[c], self.prev[c], "c") for c in (3, 4)]
        for c in range(5):
            self.cur[c][...] = new[c]
        self._adv = [padded_copy(a) for a in self.cur]
        self.stats.update(max_v=self.max_v, cfldt=cfldt, n_substeps=n)

    # BimocqSolver.cpp:149-229 given the caller's velocity after forces and final fields
    def accumulate(self, framenum, dt, forced, final):
        """forced = [u,v,w] after external forces; final = [u,v,w,rho,T] after projection."""
        d_ext = [np.ascontiguousarray(forced[c] - self._adv[c]) for c in range(3)]
        d_proj = [np.ascontiguousarray(final[c] - forced[c]) for c in range(3)]
        d_sca = [np.ascontiguousarray(final[c] - self._adv[c]) for c in (3, 4)]
        for c in range(5):
            self.cur[c][...] = final[c]
        self.accumulate_changes(framenum, dt, [padded_copy(a) for a in d_ext], [padded_copy(a) for a in d_proj],
                                [padded_copy(a) for a in d_sca])

    def accumulate_changes(self, framenum, dt, d_ext, d_proj, d_sca):
        dt32 = np.float32(dt)
        proj_coeff = 2.0
        vd = np.float32(self.vel.estimate_distortion()) / (np.float32(self.max_v) * dt32)
        sd = np.float32(self.sca.estimate_distortion()) / (np.float32(self.max_v) * dt32)
        vel_reinit = scalar_reinit = False
        if vd > 1.0 or framenum - self.vel_last_reinit > 10:
            vel_reinit = True
            self.vel_last_reinit = framenum
            proj_coeff = 1.0
        if sd > 5.0 or framenum - self.scalar_last_reinit > 30:
            scalar_reinit = True
            self.scalar_last_reinit = framenum
        for c in range(3):
            self.vel.accumulate(self.init[c], d_ext[c], self.KINDS[c], 1.0)
        for c in range(3):
            self.vel.accumulate(self.init[c], d_proj[c], self.KINDS[c], proj_coeff)
        self.sca.accumulate(self.init[3], d_sca[0], "c", 1.0)
        self.sca.accumulate(self.init[4], d_sca[1], "c", 1.0)
        if vel_reinit:
            self.vel.reinitialize()
            for c in range(3):          # velocityReinitialize, :1433-1442
                self.prev[c][...] = self.init[c]
                self.init[c][...] = self.cur[c]
            for c in range(3):
                self.vel.accumulate(self.init[c], d_proj[c], self.KINDS[c], 1.0)
        if scalar_reinit:
            self.sca.reinitialize()
            for c in (3, 4):            # scalarReinitialize, :1444-1451
                self.prev[c][...] = self.init[c]
                self.init[c][...] = self.cur[c]
        self.stats.update(vel_distortion=float(vd), scalar_distortion=float(sd), vel_reinit=vel_reinit,
                          scalar_reinit=scalar_reinit, vel_reinit_count=self.vel.total_reinit_count,
                          scalar_reinit_count=self.sca.total_reinit_count)
