"""Shared builders of seeded test inputs (tests only)."""
from __future__ import annotations

import numpy as np

from gpufluidsimulation_b200 import scenes
from oracle import oracle3d as o3

# relative L-infinity tolerance of the fp32 CUDA path against the oracle / the reference kernels,
# per kernel call and per solver step (BASELINE.json north_star: "for example 1e-5 in fp32").
TOL_STEP = 1e-5


def rel_linf(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


class Case3D:
    """A small 3D problem: grid, smooth random velocity scaled to a target CFL, smooth random
    fields, and forward/backward maps that have been evolved for a few steps by the oracle so that
    they are realistic (near identity, displaced by a few cells)."""

    def __init__(self, ni, nj, nk, h, seed=0, cfl=1.5, dt=0.05, evolve=2):
        self.ni, self.nj, self.nk = ni, nj, nk
        self.h = float(np.float32(h))
        self.dt = float(np.float32(dt))
        kinds = ("u", "v", "w")
        vel = [scenes.smooth_random(o3.shape_of(ni, nj, nk, k), seed + 10 + c) for c, k in enumerate(kinds)]
        vel = scenes.scale_to_cfl(*vel, self.h, self.dt, cfl)
        self.u, self.v, self.w = [o3.padded_copy(a.astype(np.float32)) for a in vel]
        self.max_v = max(float(np.abs(a).max()) for a in (self.u, self.v, self.w))
        self.cfldt = float(np.float32(self.h) / np.float32(self.max_v))
        L = self.h * ni
        self.fields = {k: o3.padded_copy(scenes.smooth_random(o3.shape_of(ni, nj, nk, k), seed + 20 + c))
                       for c, k in enumerate(("u", "v", "w", "c"))}
        self.fields2 = {k: o3.padded_copy(scenes.smooth_random(o3.shape_of(ni, nj, nk, k), seed + 30 + c))
                        for c, k in enumerate(("u", "v", "w", "c"))}
        m = o3.Mapper(ni, nj, nk, self.h, 1.0)
        for _ in range(evolve):
            m.update_mapping(self.u, self.v, self.w, self.cfldt, self.dt)
        self.fwd, self.bwd = m.fwd, m.bwd
        m2 = o3.Mapper(ni, nj, nk, self.h, 1.0)
        m2.update_mapping(self.u, self.v, self.w, self.cfldt, 0.5 * self.dt)
        self.bwd_prev = m2.bwd
        _ = L


# ---------------------------------------------------------------------------------------------
# running one legacy extern "C" gpu_* symbol on a device library (ours or oracle/_ref/libref3d.so)
# ---------------------------------------------------------------------------------------------
def run_gpu_symbol(lib, name, args):
    """args: list of numpy float32 arrays (copied to padded device buffers, copied back after the
    call) and python scalars.  Returns the list of arrays after the call (same order, arrays only)."""
    import ctypes as C

    import torch

    from gpufluidsimulation_b200.solver3d import alloc_field

    F = C.POINTER(C.c_float)
    dev, cargs = [], []
    for a in args:
        if isinstance(a, np.ndarray):
            # 2x allocation: the reference's gpu_compensate_field copies (ni+1)*nj*nk floats on
            # ni*nj*nk buffers (GPU_kernel.cu:678); with tight buffers that overflow corrupts a neighbour
            t = torch.zeros(2 * a.size + 64, dtype=torch.float32, device="cuda")[: a.size].view(*a.shape)
            t.copy_(torch.from_numpy(np.ascontiguousarray(a)))
            dev.append(t)
            cargs.append(C.cast(C.c_void_p(t.data_ptr()), F))
        elif isinstance(a, bool):
            cargs.append(C.c_bool(a))
        elif isinstance(a, int):
            cargs.append(C.c_int(a))
        else:
            cargs.append(C.c_float(a))
    torch.cuda.synchronize()
    fn = getattr(lib, name)
    fn.restype = None
    fn(*cargs)
    torch.cuda.synchronize()
    return [t.cpu().numpy() for t in dev]


def load_reference_lib():
    """oracle/_ref/libref3d.so = the reference's GPU_kernel.cu compiled unmodified (oracle/Makefile)."""
    import ctypes as C
    import os

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref3d.so")
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


class DeviceStepper:
    """BimocqSolver::advanceBimocq (BimocqSolver.cpp:88-230) with MapperBase buffer semantics
    (Mapping.cpp:169-236: the init buffers are COPIED into scratch before the compensation step
    clobbers them), executed on device tensors through the legacy gpu_* symbols of `lib` --
    either libbimocq_b200.so or the reference's own kernels (oracle/_ref/libref3d.so)."""

    KINDS = ("u", "v", "w", "c", "c")

    def __init__(self, ni, nj, nk, h, blend=1.0, lib=None):
        from gpufluidsimulation_b200.solver3d import GpuMapper, MapperBaseGPU, alloc_field, field_shape
        self.ni, self.nj, self.nk, self.h = ni, nj, nk, float(np.float32(h))
        self.scratch = GpuMapper(ni, nj, nk, h)
        self.vel = MapperBaseGPU().init(ni, nj, nk, h, blend, self.scratch, lib)
        self.sca = MapperBaseGPU().init(ni, nj, nk, h, blend, self.scratch, lib)
        # MapperBase semantics: the DMC output scratch carries identity values on the ring the
        # kernel does not write (in the host solver x_out holds a map whenever DMC runs, except on
        # frame 0 for the velocity mapper -- a scratch-reuse accident we do not reproduce)
        self.scratch.x_out.copy_(self.vel.InitX); self.scratch.y_out.copy_(self.vel.InitY)
        self.scratch.z_out.copy_(self.vel.InitZ)
        mk = lambda k: alloc_field(field_shape(ni, nj, nk, k))
        self.cur = [mk(k) for k in self.KINDS]
        self.init = [mk(k) for k in self.KINDS]
        self.prev = [mk(k) for k in self.KINDS]
        self._alloc = mk
        self.max_v = 0.0
        self.vel_last = self.sca_last = 0
        self.stats = {}

    def set_initial(self, u, v, w, rho, T):
        import torch
        for dst, src in zip(self.cur, (u, v, w, rho, T)):
            dst.copy_(torch.from_numpy(np.ascontiguousarray(src)))
        for c in range(5):
            self.init[c].copy_(self.cur[c]); self.prev[c].copy_(self.cur[c])

    def advect(self, frame, dt):
        m = max(1e-4, max(float(a.abs().max().item()) for a in self.cur[:3]))
        self.max_v = float(np.float32(m))
        cfldt = float(np.float32(self.h) / np.float32(m))
        if frame == 0:
            self.max_v = self.h
        u, v, w = self.cur[:3]
        self.vel.updateMapping(u, v, w, cfldt, dt)
        self.sca.updateMapping(u, v, w, cfldt, dt)
        nu, nv, nw = [self._alloc(k) for k in "uvw"]
        iu, iv, iw = [t.clone() for t in self.init[:3]]
        iu, iv, iw = [self._pad(t, k) for t, k in zip((iu, iv, iw), "uvw")]
        self.vel.advectVelocity(nu, nv, nw, iu, iv, iw, *self.prev[:3])
        out = [nu, nv, nw]
        for c in (3, 4):
            f = self._alloc("c")
            self.sca.advectField(f, self._pad(self.init[c].clone(), "c"), self.prev[c])
            out.append(f)
        for c in range(5):
            self.cur[c].copy_(out[c])
        self._adv = [t.clone() for t in self.cur]
        self.stats.update(max_v=self.max_v, cfldt=cfldt)

    def _pad(self, t, kind):
        p = self._alloc(kind)
        p.copy_(t)
        return p

    def accumulate(self, frame, dt, forced, final):
        import torch
        dev = self.cur[0].device
        forced = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in forced]
        final = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in final]
        d_ext = [self._pad(forced[c] - self._adv[c], self.KINDS[c]) for c in range(3)]
        d_proj = [self._pad(final[c] - forced[c], self.KINDS[c]) for c in range(3)]
        d_sca = [self._pad(final[c] - self._adv[c], "c") for c in (3, 4)]
        for c in range(5):
            self.cur[c].copy_(final[c])
        dt32 = np.float32(dt)
        vd = np.float32(self.vel.estimateDistortion()) / (np.float32(self.max_v) * dt32)
        sd = np.float32(self.sca.estimateDistortion()) / (np.float32(self.max_v) * dt32)
        proj_coeff = 2.0
        vre = sre = False
        if vd > 1.0 or frame - self.vel_last > 10:
            vre, self.vel_last, proj_coeff = True, frame, 1.0
        if sd > 5.0 or frame - self.sca_last > 30:
            sre, self.sca_last = True, frame
        self.vel.accumulateVelocity(*self.init[:3], *d_ext, 1.0)
        self.vel.accumulateVelocity(*self.init[:3], *d_proj, proj_coeff)
        self.sca.accumulateField(self.init[3], d_sca[0])
        self.sca.accumulateField(self.init[4], d_sca[1])
        if vre:
            self.vel.reinitializeMapping()
            for c in range(3):
                self.prev[c].copy_(self.init[c]); self.init[c].copy_(self.cur[c])
            self.vel.accumulateVelocity(*self.init[:3], *d_proj, 1.0)
        if sre:
            self.sca.reinitializeMapping()
            for c in (3, 4):
                self.prev[c].copy_(self.init[c]); self.init[c].copy_(self.cur[c])
        self.stats.update(vel_distortion=float(vd), scalar_distortion=float(sd), vel_reinit=vre, scalar_reinit=sre)
