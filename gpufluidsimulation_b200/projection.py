"""Host-side mirror of the reference's pressure projection entry point on libbimocq_b200.so.

``PressureProjection3D`` = ``BimocqGPUSolver::projection`` (bimocq3D/BimocqGPUSolver.cpp:406-467,
live branch: ``gpuMapper::projectionMultiGrid`` with LEVEL_COUNT = 6 levels, 50 iterations,
halfrdx = 0.5) on the handle API ``bmq_mgpcg_*``; ``make_levels`` / ``projection_multi_grid`` drive
the legacy symbol ``gpu_multi_grid_conjugate_gradient`` with caller-owned buffers the way
``gpuMapper`` does (GPU_Advection.h:622-626).  torch supplies device memory only.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import CoarseLevel, check, check_legacy, load_library
from .solver3d import _dp, _torch

LEVEL_COUNT = 6          # GPU_Advection.h:12
MG_BUFFERS = {"div": 0, "p": 1, "dir": 2, "residual": 3, "result": 4}


def level_dims(ni, nj, nk, levels=LEVEL_COUNT):
    """BimocqGPUSolver.cpp:68-90: n_{l+1} = (n_l - 1) / 2."""
    dims = [(ni, nj, nk)]
    for _ in range(1, levels):
        a, b, c = dims[-1]
        dims.append(((a - 1) // 2, (b - 1) // 2, (c - 1) // 2))
    return dims


def max_levels(ni, nj, nk, cap=LEVEL_COUNT):
    n = 1
    while n < cap and min(level_dims(ni, nj, nk, n + 1)[-1]) >= 3:
        n += 1
    return n


def alloc_double(n, pad=0):
    """Zeroed float64 device buffer of n elements with `pad` zero elements of slack behind it."""
    torch = _torch()
    return torch.zeros(n + pad, dtype=torch.float64, device="cuda")[:n]


def make_levels(ni, nj, nk, levels=LEVEL_COUNT, pad_planes=1):
    """Level table with caller-owned buffers.  Returns (ctypes array, list of tensors kept alive).
    `pad_planes` zero planes follow every buffer: the reference's prolongation reads one plane past
    levels[l+1].x when level l has an even size (GPU_kernel.cu:1620)."""
    arr = (CoarseLevel * levels)()
    keep = []
    for l, (a, b, c) in enumerate(level_dims(ni, nj, nk, levels)):
        if min(a, b, c) < 3:
            raise ValueError(f"level {l} would be {a}x{b}x{c}")
        n = a * b * c
        bufs = [alloc_double(n, pad_planes * a * b + a + 2) for _ in range(3)]
        keep.append(bufs)
        arr[l].ni, arr[l].nj, arr[l].nk, arr[l].number = a, b, c, n
        arr[l].alpha, arr[l].beta = -1.0, 1.0 / 6.0
        arr[l].b, arr[l].x, arr[l].r = (t.data_ptr() for t in bufs)
    return arr, keep


def projection_multi_grid(u, v, w, div, p, dir, residual, temp0, temp1, temp_result, levels, iters, halfrdx=0.5, lib=None):
    """gpuMapper::projectionMultiGrid (GPU_Advection.h:622-626) on device tensors."""
    own = lib is None
    lib = lib or load_library()
    fn = lib.gpu_multi_grid_conjugate_gradient
    if not own:
        fn.restype = None
        fn.argtypes = capi._PROTOS["gpu_multi_grid_conjugate_gradient"][1]
    d = lambda t: C.c_void_p(t.data_ptr())
    fn(_dp(u), _dp(v), _dp(w), d(div), d(p), d(dir), d(residual), d(temp0), d(temp1), d(temp_result), levels, len(levels), iters, halfrdx)
    if own:
        check_legacy("gpu_multi_grid_conjugate_gradient")


class PressureProjection3D:
    """Owns the fp64 work buffers (BimocqGPUSolver.cpp:56-90) and runs the reference's projection."""

    def __init__(self, ni, nj, nk, levels=None):
        _torch()
        self.lib = load_library()
        self.ni, self.nj, self.nk = ni, nj, nk
        self.levels = max_levels(ni, nj, nk) if levels is None else levels
        h = C.c_void_p()
        check(self.lib.bmq_mgpcg_create(ni, nj, nk, self.levels, C.byref(h)), "bmq_mgpcg_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.bmq_mgpcg_destroy(self.h)
            self.h = None

    __del__ = close

    def set_stream(self, stream):
        check(self.lib.bmq_mgpcg_set_stream(self.h, C.c_void_p(stream.cuda_stream if stream is not None else 0)))

    def project(self, u, v, w, iters=50, halfrdx=0.5):
        """u, v, w: float32 CUDA face tensors (nk,nj,ni+1), (nk,nj+1,ni), (nk+1,nj,ni); in place."""
        assert tuple(u.shape) == (self.nk, self.nj, self.ni + 1) and tuple(v.shape) == (self.nk, self.nj + 1, self.ni)
        assert tuple(w.shape) == (self.nk + 1, self.nj, self.ni)
        check(self.lib.bmq_mgpcg_solve(self.h, _dp(u), _dp(v), _dp(w), iters, halfrdx), "bmq_mgpcg_solve")

    def buffer(self, name):
        """Zero-copy float64 view of div / p / dir / residual (nk,nj,ni) or result (4096)."""
        torch = _torch()
        ptr, cnt = C.c_void_p(), C.c_longlong()
        check(self.lib.bmq_mgpcg_buffer(self.h, MG_BUFFERS[name], C.byref(ptr), C.byref(cnt)), "bmq_mgpcg_buffer")
        n = cnt.value
        iface = {"shape": (n,), "typestr": "<f8", "data": (ptr.value, False), "version": 3}
        holder = type("_View", (), {"__cuda_array_interface__": iface})()
        t = torch.as_tensor(holder, device="cuda")
        return t if name == "result" else t.view(self.nk, self.nj, self.ni)

    def residual_history(self, iters):
        """tempResult[2000 .. 2000+iters]: max residual before the first and after every iteration."""
        return self.buffer("result")[2000:2001 + iters].cpu().numpy()
