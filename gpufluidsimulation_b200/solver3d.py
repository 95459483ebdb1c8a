"""Host-side mirror of the reference's 3D advection interface on top of libbimocq_b200.so.

* ``MapperBaseGPU`` -- same methods, argument order and buffer contract as the reference class
  (bimocq3D/Mapping.h:47-105, Mapping.cpp:276-447), driving the legacy ``gpu_*`` symbols the way
  ``gpuMapper`` does (GPU_Advection.h:453-600).  Buffers are torch CUDA tensors (device memory only).
* ``BimocqAdvection3D`` -- the handle API: device-resident state, fused stages and the
  reinitialisation scheduler of ``BimocqSolver::advanceBimocq`` (BimocqSolver.cpp:88-230).

PyTorch is used for device memory and streams only; all arithmetic happens in the library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import FIELD, Stats3D, check, check_legacy, load_library

_F = C.POINTER(C.c_float)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise capi.BimocqLibraryError("gpufluidsimulation_b200 needs a CUDA device; there is no CPU fallback")
    return torch


def _dp(t):
    """Device pointer of a contiguous float32 CUDA tensor as float*."""
    assert t.is_cuda and t.dtype == _torch().float32 and t.is_contiguous()
    return C.cast(C.c_void_p(t.data_ptr()), _F)


def alloc_field(shape, device="cuda"):
    """Zeroed float32 device field with one plane + one row of slack behind it (the reference's
    sampler reads one node past its clamp bound with weight zero, GPU_kernel.cu:53-61)."""
    torch = _torch()
    n = int(np.prod(shape))
    slack = shape[-1] * shape[-2] + shape[-1] + 2
    buf = torch.zeros(n + slack, dtype=torch.float32, device=device)
    return buf[:n].view(*shape)


def field_shape(ni, nj, nk, kind):
    dx, dy, dz = {"u": (1, 0, 0), "v": (0, 1, 0), "w": (0, 0, 1), "c": (0, 0, 0)}[kind]
    return (nk + dz, nj + dy, ni + dx)


class GpuMapper:
    """The scratch buffers of the reference's gpuMapper (GPU_Advection.h:118-136) that the
    MapperBaseGPU methods borrow: u_src/v_src/w_src for compensation and x_out/y_out/z_out for the
    DMC ping-pong."""

    def __init__(self, ni, nj, nk, h):
        self.ni, self.nj, self.nk, self.h = ni, nj, nk, float(np.float32(h))
        self.u_src = alloc_field(field_shape(ni, nj, nk, "u"))
        self.v_src = alloc_field(field_shape(ni, nj, nk, "v"))
        self.w_src = alloc_field(field_shape(ni, nj, nk, "w"))
        self.x_out = alloc_field(field_shape(ni, nj, nk, "c"))
        self.y_out = alloc_field(field_shape(ni, nj, nk, "c"))
        self.z_out = alloc_field(field_shape(ni, nj, nk, "c"))
        self.du = alloc_field(field_shape(ni, nj, nk, "u"))


class MapperBaseGPU:
    """Drop-in mirror of the reference MapperBaseGPU (Mapping.h:47-105)."""

    def init(self, ni, nj, nk, h, coeff, mymapper: GpuMapper, lib=None):
        """``lib``: a ctypes handle exporting the legacy gpu_* symbols; defaults to libbimocq_b200.so.
        (The parity tests pass the reference's own kernels, oracle/_ref/libref3d.so, to drive the
        identical call sequence through both.)"""
        torch = _torch()
        self.CellNumberX, self.CellNumberY, self.CellNumberZ = ni, nj, nk
        self.CellSize = float(np.float32(h))
        self.BlendCoeff = float(coeff)
        self.TotalReinitCount = 0
        self.gpuSolver = mymapper
        self._ours = lib is None
        self.lib = load_library() if lib is None else _with_legacy_prototypes(lib)
        shp = field_shape(ni, nj, nk, "c")
        h32 = np.float32(h)
        # Mapping.cpp:310-324: Init = (float)i * CellSize
        ix = torch.from_numpy(np.arange(ni, dtype=np.float32) * h32).cuda()
        iy = torch.from_numpy(np.arange(nj, dtype=np.float32) * h32).cuda()
        iz = torch.from_numpy(np.arange(nk, dtype=np.float32) * h32).cuda()
        self.InitX = alloc_field(shp); self.InitX[...] = ix[None, None, :]
        self.InitY = alloc_field(shp); self.InitY[...] = iy[None, :, None]
        self.InitZ = alloc_field(shp); self.InitZ[...] = iz[:, None, None]
        for name in ("Forward", "Backward"):
            for ax, src in zip("XYZ", (self.InitX, self.InitY, self.InitZ)):
                t = alloc_field(shp); t.copy_(src)
                setattr(self, name + ax, t)
        for ax, src in zip("XYZ", (self.InitX, self.InitY, self.InitZ)):
            t = alloc_field(shp); t.copy_(src)
            setattr(self, "Backward" + ax + "Prev", t)
        return self

    def _dims(self):
        return self.CellSize, self.CellNumberX, self.CellNumberY, self.CellNumberZ

    def _check(self, what):
        if self._ours:
            check_legacy(what)

    # Mapping.cpp:354-368 + gpuMapper::solveBackwardDMC (GPU_Advection.h:460-470)
    def updateBackward(self, velocityU, velocityV, velocityW, cfldt, dt):
        g = self.gpuSolver
        h, ni, nj, nk = self._dims()
        T = np.float32(0.0); substep = np.float32(cfldt); dt = np.float32(dt)
        while T < dt:
            if T + substep > dt:
                substep = np.float32(dt - T)
            self.lib.gpu_solve_backwardDMC(_dp(velocityU), _dp(velocityV), _dp(velocityW), _dp(self.BackwardX),
                                           _dp(self.BackwardY), _dp(self.BackwardZ), _dp(g.x_out), _dp(g.y_out),
                                           _dp(g.z_out), h, ni, nj, nk, float(substep))
            self.BackwardX.copy_(g.x_out); self.BackwardY.copy_(g.y_out); self.BackwardZ.copy_(g.z_out)
            T = np.float32(T + substep)
        self._check("gpu_solve_backwardDMC")

    # Mapping.cpp:370-373
    def updateForward(self, velocityU, velocityV, velocityW, cfldt, dt):
        h, ni, nj, nk = self._dims()
        self.lib.gpu_solve_forward(_dp(velocityU), _dp(velocityV), _dp(velocityW), _dp(self.ForwardX),
                                   _dp(self.ForwardY), _dp(self.ForwardZ), h, ni, nj, nk, float(np.float32(cfldt)),
                                   float(np.float32(dt)))
        self._check("gpu_solve_forward")

    # Mapping.cpp:347-352
    def updateMapping(self, velocityU, velocityV, velocityW, cfldt, dt):
        self.updateBackward(velocityU, velocityV, velocityW, cfldt, dt)
        self.updateForward(velocityU, velocityV, velocityW, cfldt, dt)

    # Mapping.cpp:375-391.  Like the reference, velocityUInit.. are overwritten with the
    # pre-correction velocity by the compensation step (GPU_kernel.cu:656-658).
    def advectVelocity(self, velocityU, velocityV, velocityW, velocityUInit, velocityVInit, velocityWInit,
                       velocityUPrev, velocityVPrev, velocityWPrev):
        g = self.gpuSolver
        h, ni, nj, nk = self._dims()
        L = self.lib
        for t in (velocityU, velocityV, velocityW):
            t.zero_()                                    # GPU_Advection.h:477-479
        L.gpu_advect_velocity(_dp(velocityU), _dp(velocityV), _dp(velocityW), _dp(velocityUInit), _dp(velocityVInit),
                              _dp(velocityWInit), _dp(self.BackwardX), _dp(self.BackwardY), _dp(self.BackwardZ), h,
                              ni, nj, nk, False)
        for t in (g.u_src, g.v_src, g.w_src):
            t.zero_()                                    # GPU_Advection.h:499-501
        L.gpu_compensate_velocity(_dp(velocityU), _dp(velocityV), _dp(velocityW), _dp(velocityUInit),
                                  _dp(velocityVInit), _dp(velocityWInit), _dp(g.u_src), _dp(g.v_src), _dp(g.w_src),
                                  _dp(self.ForwardX), _dp(self.ForwardY), _dp(self.ForwardZ), _dp(self.BackwardX),
                                  _dp(self.BackwardY), _dp(self.BackwardZ), h, ni, nj, nk, False)
        blend = self.BlendCoeff if self.TotalReinitCount != 0 else 1.0
        L.gpu_advect_vel_double(_dp(velocityU), _dp(velocityV), _dp(velocityW), _dp(velocityUPrev), _dp(velocityVPrev),
                                _dp(velocityWPrev), _dp(self.BackwardX), _dp(self.BackwardY), _dp(self.BackwardZ),
                                _dp(self.BackwardXPrev), _dp(self.BackwardYPrev), _dp(self.BackwardZPrev), h, ni, nj,
                                nk, False, blend)
        self._check("advectVelocity")

    # Mapping.cpp:393-407
    def advectField(self, field, fieldInit, fieldPrev):
        g = self.gpuSolver
        h, ni, nj, nk = self._dims()
        L = self.lib
        field.zero_()
        L.gpu_advect_field(_dp(field), _dp(fieldInit), _dp(self.BackwardX), _dp(self.BackwardY), _dp(self.BackwardZ),
                           h, ni, nj, nk, False)
        g.u_src.zero_()
        L.gpu_compensate_field(_dp(field), _dp(fieldInit), _dp(g.u_src), _dp(self.ForwardX), _dp(self.ForwardY),
                               _dp(self.ForwardZ), _dp(self.BackwardX), _dp(self.BackwardY), _dp(self.BackwardZ), h,
                               ni, nj, nk, False)
        blend = self.BlendCoeff if self.TotalReinitCount != 0 else 1.0
        L.gpu_advect_field_double(_dp(field), _dp(fieldPrev), _dp(self.BackwardX), _dp(self.BackwardY),
                                  _dp(self.BackwardZ), _dp(self.BackwardXPrev), _dp(self.BackwardYPrev),
                                  _dp(self.BackwardZPrev), h, ni, nj, nk, False, blend)
        self._check("advectField")

    # Mapping.cpp:420-423 (argument order of the definition: init buffers first)
    def accumulateVelocity(self, duInit, dvInit, dwInit, uChange, vChange, wChange, coeff):
        h, ni, nj, nk = self._dims()
        self.lib.gpu_accumulate_velocity(_dp(uChange), _dp(vChange), _dp(wChange), _dp(duInit), _dp(dvInit),
                                         _dp(dwInit), _dp(self.ForwardX), _dp(self.ForwardY), _dp(self.ForwardZ), h,
                                         ni, nj, nk, False, float(coeff))
        self._check("accumulateVelocity")

    # Mapping.cpp:425-428
    def accumulateField(self, dfieldInit, fieldChange):
        h, ni, nj, nk = self._dims()
        self.lib.gpu_accumulate_field(_dp(fieldChange), _dp(dfieldInit), _dp(self.ForwardX), _dp(self.ForwardY),
                                      _dp(self.ForwardZ), h, ni, nj, nk, False, 1.0)
        self._check("accumulateField")

    # Mapping.cpp:495-519 (boundary: optional int8 tensor, cells == 2 are skipped)
    def estimateDistortion(self, boundary=None):
        torch = _torch()
        g = self.gpuSolver
        h, ni, nj, nk = self._dims()
        g.du.zero_()
        self.lib.gpu_estimate_distortion(_dp(g.du), _dp(self.BackwardX), _dp(self.BackwardY), _dp(self.BackwardZ),
                                         _dp(self.ForwardX), _dp(self.ForwardY), _dp(self.ForwardZ), h, ni, nj, nk)
        self._check("gpu_estimate_distortion")
        d = g.du.reshape(-1)[: ni * nj * nk].view(nk, nj, ni)
        if boundary is not None:
            d = torch.where(boundary == 2, torch.zeros_like(d), d)
        return float(torch.sqrt(d.max()).item())

    # Mapping.cpp:430-447
    def reinitializeMapping(self):
        self.TotalReinitCount += 1
        for ax, init in zip("XYZ", (self.InitX, self.InitY, self.InitZ)):
            getattr(self, "Backward" + ax + "Prev").copy_(getattr(self, "Backward" + ax))
            getattr(self, "Backward" + ax).copy_(init)
            getattr(self, "Forward" + ax).copy_(init)


def _with_legacy_prototypes(lib):
    for name, (res, args) in capi._PROTOS.items():
        if name.startswith("gpu_"):
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
    return lib


class _DevView:
    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class BimocqAdvection3D:
    """Handle API wrapper (include/bimocq_b200.h, bmq3d_*).  One instance per GPU (or per z-slab)."""

    CURRENT = ("U", "V", "W", "RHO", "T")

    def __init__(self, ni, nj, nk, h, blend_coeff=1.0, slab=None, halo=0, borrowed_handle=None):
        """borrowed_handle: wrap an existing bmq3d_solver* (e.g. the slab handle a bmq3d_mg owns) instead of
        creating one; close() then leaves it alone."""
        _torch()
        self.lib = load_library()
        self.ni, self.nj, self.nk = ni, nj, nk
        self.h = float(np.float32(h))
        self._owned = borrowed_handle is None
        if borrowed_handle is not None:
            self._h = borrowed_handle
            return
        self._h = C.c_void_p()
        k0, k1 = slab if slab is not None else (0, nk)
        check(self.lib.bmq3d_create_slab(ni, nj, nk, self.h, float(blend_coeff), k0, k1, halo, C.byref(self._h)),
              "bmq3d_create_slab")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            if self._owned:
                self.lib.bmq3d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def grow_halo(self, new_halo):
        """Slab handles: re-allocate every field with a wider halo (device pointers change)."""
        check(self.lib.bmq3d_grow_halo(self._h, int(new_halo)), "bmq3d_grow_halo")

    def set_stream(self, torch_stream):
        check(self.lib.bmq3d_set_stream(self._h, C.c_void_p(torch_stream.cuda_stream if torch_stream else 0)))

    def field_info(self, name):
        ptr = C.c_void_p(); p0 = C.c_int(); npl = C.c_int(); nx = C.c_int(); ny = C.c_int()
        check(self.lib.bmq3d_field_ptr(self._h, FIELD[name], C.byref(ptr), C.byref(p0), C.byref(npl), C.byref(nx),
                                       C.byref(ny)), "bmq3d_field_ptr")
        return ptr.value, p0.value, npl.value, nx.value, ny.value

    def field(self, name):
        """Zero-copy torch view (planes, ny, nx) of a device field's stored planes.  Map and
        init/prev pointers rotate between calls: fetch a fresh view after every stage."""
        torch = _torch()
        ptr, p0, npl, nx, ny = self.field_info(name)
        return torch.as_tensor(_DevView(ptr, (npl, ny, nx)), device="cuda")

    def set_host_layout(self, blocked: bool):
        """Host buffers of upload / download / *_host are dense x-fastest (default) or the reference's
        8^3-blocked buffer3Df storage (include/fluid_buffer3D.h:173-189), relaid out on the device."""
        check(self.lib.bmq3d_set_host_layout(self._h, 1 if blocked else 0), "bmq3d_set_host_layout")
        self._blocked = bool(blocked)

    def host_elems(self, name):
        _, _, npl, nx, ny = self.field_info(name)
        return int(self.lib.bmq_blocked_elems(nx, ny, npl)) if getattr(self, "_blocked", False) else npl * nx * ny

    def upload(self, name, host):
        host = np.ascontiguousarray(host, dtype=np.float32)
        assert host.size == self.host_elems(name), (name, host.shape, self.host_elems(name))
        check(self.lib.bmq3d_upload(self._h, FIELD[name], host.ctypes.data_as(C.c_void_p)), "bmq3d_upload")

    def download(self, name):
        _, _, npl, nx, ny = self.field_info(name)
        out = (np.empty(self.host_elems(name), dtype=np.float32) if getattr(self, "_blocked", False)
               else np.empty((npl, ny, nx), dtype=np.float32))
        check(self.lib.bmq3d_download(self._h, FIELD[name], out.ctypes.data_as(C.c_void_p)), "bmq3d_download")
        return out

    def set_initial(self, u, v, w, rho, T):
        for n, a in zip(self.CURRENT, (u, v, w, rho, T)):
            self.upload(n, a)
        check(self.lib.bmq3d_reset(self._h), "bmq3d_reset")

    def reset(self):
        check(self.lib.bmq3d_reset(self._h), "bmq3d_reset")

    def set_initial_device(self, u, v, w, rho, T):
        """Initial fields given as device tensors (whole grid)."""
        for n, a in zip(self.CURRENT, (u, v, w, rho, T)):
            self.field(n).copy_(a)
        _torch().cuda.synchronize()
        self.reset()

    def apply_buoyancy(self, beta, dt, alpha=0.0):
        """Caller stand-in used by the benchmark between the two phases (NOT part of the hot
        path): the reference's buoyancy (GPU_kernel.cu:804-823), dv = dt*(-alpha*rho + beta*T)
        averaged onto the v faces, written to DV_EXT and added to V; the other change fields
        stay zero."""
        torch = _torch()
        T, dv, V = self.field("T"), self.field("DV_EXT"), self.field("V")
        inner = dv[:, 1:-1, :]
        torch.add(T[:, 1:, :], T[:, :-1, :], out=inner)
        inner.mul_(0.5 * dt * beta)
        if alpha:
            rho = self.field("RHO")
            inner.add_(rho[:, 1:, :] + rho[:, :-1, :], alpha=-0.5 * dt * alpha)
        V.add_(dv)

    def advect(self, framenum, dt, with_semilag=False):
        check(self.lib.bmq3d_advect(self._h, int(framenum), float(np.float32(dt)), int(with_semilag)), "bmq3d_advect")

    def accumulate(self, framenum, dt):
        check(self.lib.bmq3d_accumulate(self._h, int(framenum), float(np.float32(dt))), "bmq3d_accumulate")

    def advect_host(self, framenum, dt, u, v, w, rho, T):
        ptrs = [a.ctypes.data_as(C.c_void_p) if isinstance(a, np.ndarray) else C.c_void_p(a.data_ptr())
                for a in (u, v, w, rho, T)]
        check(self.lib.bmq3d_advect_host(self._h, int(framenum), float(np.float32(dt)), *ptrs), "bmq3d_advect_host")

    def accumulate_host(self, framenum, dt, forced, final):
        arrs = list(forced) + list(final)
        ptrs = [a.ctypes.data_as(C.c_void_p) if isinstance(a, np.ndarray) else C.c_void_p(a.data_ptr()) for a in arrs]
        check(self.lib.bmq3d_accumulate_host(self._h, int(framenum), float(np.float32(dt)), *ptrs),
              "bmq3d_accumulate_host")

    def stats(self):
        st = Stats3D()
        check(self.lib.bmq3d_get_stats(self._h, C.byref(st)), "bmq3d_get_stats")
        return st.as_dict()

    def timing_enable(self, on=True):
        check(self.lib.bmq3d_timing_enable(self._h, int(on)), "bmq3d_timing_enable")

    def timing_read(self):
        """{stage name: (milliseconds, spans)} since the last read (synchronises the stream)."""
        n = capi.N_TIMING_SLOTS
        ms = (C.c_float * n)(); cnt = (C.c_int * n)()
        check(self.lib.bmq3d_timing_read(self._h, ms, cnt, n), "bmq3d_timing_read")
        return {self.lib.bmq3d_timing_slot_name(q).decode(): (ms[q], cnt[q]) for q in range(n)}

    def timing_read_gaps(self):
        """{stage name: (milliseconds, spans, milliseconds the stream idled before the stage)} since the last read."""
        n = capi.N_TIMING_SLOTS
        ms = (C.c_float * n)(); cnt = (C.c_int * n)(); gap = (C.c_float * n)()
        check(self.lib.bmq3d_timing_read_gaps(self._h, ms, cnt, gap, n), "bmq3d_timing_read_gaps")
        return {self.lib.bmq3d_timing_slot_name(q).decode(): (ms[q], cnt[q], gap[q]) for q in range(n)}

    # fine-grained stages (z-slab driver)
    def stage(self, name, *args):
        fn = getattr(self.lib, "bmq3d_stage_" + name)
        check(fn(self._h, *args), "bmq3d_stage_" + name)
