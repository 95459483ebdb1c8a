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
            t = alloc_field(a.shape)
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
