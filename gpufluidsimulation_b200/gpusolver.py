"""Host-side mirror of the reference's device-resident smoke solver, ``BimocqGPUSolver``
(bimocq3D/BimocqGPUSolver.{h,cpp}), both of its schemes: same members, same call sequences
(``advanceBimocq``, BimocqGPUSolver.cpp:128-232; ``advanceReflection``, the MacCormack + reflection scheme,
:232-335; ``semilagAdvect``, :337-344), every device operation going through the legacy
``gpu_*`` symbols exactly as ``gpuMapper`` forwards them (GPU_Advection.h:328-626).

It exists to show -- and test -- that libbimocq_b200.so replaces the WHOLE frame of that solver
(advection, smoke emission, buoyancy, diffusion, projection, accumulation, reinitialisation), not
only the advection kernels: ``tests/test_gpusolver_frame_gpu.py`` runs it once on
libbimocq_b200.so and once on the reference's own kernels (``lib=`` oracle/_ref/libref3d.so) and
compares every field bit for bit.  torch supplies device memory and D2D copies only.

Differences from the reference, both deliberate:
* ``getCFL`` (BimocqGPUSolver.cpp:348-373) walks host copies of u, v, w that the reference only
  refreshes when it writes output files; here the maximum is taken from the device fields
  (``bmq_max_abs3``), which is what the host-orchestrated solver does (BimocqSolver.cpp:1067-1118).
* OpenVDB output and the emitter objects' time-dependent ``update`` are not mirrored; the two
  emitters are the fixed spheres hard-coded in ``emitSmoke`` (BimocqGPUSolver.cpp:386-389).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi, projection
from .capi import check, check_legacy, load_library
from .solver3d import GpuMapper, MapperBaseGPU, _dp, _torch, _with_legacy_prototypes, alloc_field, field_shape

LEVEL_COUNT = projection.LEVEL_COUNT


class BimocqGPUSolver:
    def __init__(self, nx, ny, nz, L, vis_coeff=0.0, blend_coeff=1.0, mymapper: GpuMapper | None = None, lib=None,
                 levels=LEVEL_COUNT):
        torch = _torch()
        self.CellNumberX, self.CellNumberY, self.CellNumberZ = nx, ny, nz
        self.CellSize = float(np.float32(L) / np.float32(nx))
        self.MaxVelocity = 0.0
        self.Viscosity = float(vis_coeff)
        self._ours = lib is None
        self.ours = load_library()                      # bmq_max_abs3 always comes from here
        self.lib = self.ours if lib is None else _with_legacy_prototypes(lib)
        if lib is not None:
            fn = self.lib.gpu_multi_grid_conjugate_gradient
            fn.restype, fn.argtypes = capi._PROTOS["gpu_multi_grid_conjugate_gradient"]
        self.GpuSolver = mymapper or GpuMapper(nx, ny, nz, self.CellSize)
        mk = lambda kind: alloc_field(field_shape(nx, ny, nz, kind))
        for comp, kind in (("U", "u"), ("V", "v"), ("W", "w")):
            for suffix in ("", "Init", "Prev", "Temp"):
                setattr(self, f"Velocity{comp}{suffix}", mk(kind))
            setattr(self, f"d{comp.lower()}Proj", mk(kind))
            setattr(self, f"d{comp.lower()}Extern", mk(kind))
            setattr(self, f"TempSrc{comp}", mk(kind))
        for name in ("Density", "Temperature"):
            for suffix in ("", "Init", "Prev", "Temp", "Extern"):
                setattr(self, name + suffix, mk("c"))
        n = nx * ny * nz
        pad = nx * ny + nx + 2
        for name in ("p", "dir", "residual", "div", "temp0", "temp1"):
            setattr(self, name, projection.alloc_double(n, pad))
        self.tempResult = projection.alloc_double(4096)
        self.LevelCount = levels
        self.levels, self._level_buffers = projection.make_levels(nx, ny, nz, levels)
        self.VelocityAdvector = MapperBaseGPU().init(nx, ny, nz, self.CellSize, blend_coeff, self.GpuSolver, lib)
        self.ScalarAdvector = MapperBaseGPU().init(nx, ny, nz, self.CellSize, blend_coeff, self.GpuSolver, lib)
        self.vel_lastReinit = 0
        self.scalar_lastReinit = 0
        self._alpha, self._beta = 0.0, 0.0
        self.emitFrame = 0
        self.emit_density, self.emit_temperature = 1.0, 1.0
        self.ProjectionIterations = 50     # BimocqGPUSolver.cpp:444
        torch.cuda.synchronize()

    # BimocqGPUSolver.cpp:529-540
    def setSmoke(self, drop, raise_, emit_frames=0, emit_density=1.0, emit_temperature=1.0):
        self._alpha, self._beta = float(drop), float(raise_)
        self.emitFrame = int(emit_frames)
        self.emit_density, self.emit_temperature = float(emit_density), float(emit_temperature)

    def _check(self, what):
        if self._ours:
            check_legacy(what)

    # BimocqGPUSolver.cpp:348-373 (see the module docstring)
    def getCFL(self):
        m = C.c_float()
        u, v, w = self.VelocityU, self.VelocityV, self.VelocityW
        check(self.ours.bmq_max_abs3(_dp(u), u.numel(), _dp(v), v.numel(), _dp(w), w.numel(), C.byref(m)), "bmq_max_abs3")
        self.MaxVelocity = max(float(np.float32(1e-4)), m.value)
        return float(np.float32(self.CellSize) / np.float32(self.MaxVelocity))

    # BimocqGPUSolver.cpp:375-391
    def emitSmoke(self, framenum, dt):
        if framenum < self.emitFrame:
            dims = (self.CellSize, self.CellNumberX, self.CellNumberY, self.CellNumberZ)
            f = [_dp(t) for t in (self.VelocityU, self.VelocityV, self.VelocityW, self.Density, self.Temperature)]
            self.lib.gpu_emit_smoke(*f, *dims, 0.04, 0.2, 0.2, 0.015, self.emit_density, self.emit_temperature, 1.0)
            self.lib.gpu_emit_smoke(*f, *dims, 0.16, 0.201, 0.2, 0.015, self.emit_density, self.emit_temperature, -1.0)
            self._check("gpu_emit_smoke")

    # BimocqGPUSolver.cpp:393-396
    def addBuoyancy(self, dt):
        self.lib.gpu_add_buoyancy(_dp(self.VelocityV), _dp(self.Density), _dp(self.Temperature), self.CellNumberX,
                                  self.CellNumberY, self.CellNumberZ, self._alpha, self._beta, dt)
        self._check("gpu_add_buoyancy")

    # BimocqGPUSolver.cpp:398-403
    def diffuseField(self, field, fieldTemp0, fieldTemp1, ni, nj, nk, iters, nu, dt):
        coef = float(np.float32(nu) * (np.float32(dt) / (np.float32(self.CellSize) * np.float32(self.CellSize))))
        self.lib.gpu_diffuse_field(_dp(field), _dp(fieldTemp0), _dp(fieldTemp1), ni, nj, nk, iters, coef)
        self._check("gpu_diffuse_field")

    # BimocqGPUSolver.cpp:406-467, live branch
    def projection(self):
        d = lambda t: C.c_void_p(t.data_ptr())
        self.lib.gpu_multi_grid_conjugate_gradient(
            _dp(self.VelocityU), _dp(self.VelocityV), _dp(self.VelocityW), d(self.div), d(self.p), d(self.dir),
            d(self.residual), d(self.temp0), d(self.temp1), d(self.tempResult), self.levels, self.LevelCount,
            self.ProjectionIterations, 0.5)
        self._check("gpu_multi_grid_conjugate_gradient")

    def _add_fields(self, out, f1, f2, coeff):
        self.lib.gpu_add_field(_dp(out), _dp(f1), _dp(f2), coeff, out.numel())

    def _add(self, f1, f2, coeff):
        self.lib.gpu_add(_dp(f1), _dp(f2), coeff, f1.numel())

    # BimocqGPUSolver.cpp:503-527
    def velocityReinitialize(self):
        for c in "UVW":
            getattr(self, f"Velocity{c}Prev").copy_(getattr(self, f"Velocity{c}Init"))
            getattr(self, f"Velocity{c}Init").copy_(getattr(self, f"Velocity{c}"))

    def scalarReinitialize(self):
        for n in ("Density", "Temperature"):
            getattr(self, n + "Prev").copy_(getattr(self, n + "Init"))
            getattr(self, n + "Init").copy_(getattr(self, n))

    # BimocqGPUSolver.cpp:109-126
    def advance(self, framenum, dt, scheme="BIMOCQ"):
        if scheme == "BIMOCQ":
            self.advanceBimocq(framenum, dt)
        elif scheme == "MAC_REFLECTION":
            self.advanceReflection(framenum, dt)
        else:
            raise ValueError(f"unknown scheme {scheme!r} (the reference's GPU solver has BIMOCQ and MAC_REFLECTION)")

    # ---- gpuMapper's forwarding members used by the non-BiMocq schemes (GPU_Advection.h:530-551, 610-613)
    def _semilagAdvectField(self, field, field_src, cfldt, dt):
        nx, ny, nz = self.CellNumberX, self.CellNumberY, self.CellNumberZ
        field.zero_()      # the reference clears (ni+1)*nj*nk floats of a centred buffer; the in-range part is what matters
        self.lib.gpu_semilag(_dp(field), _dp(field_src), _dp(self.VelocityU), _dp(self.VelocityV), _dp(self.VelocityW),
                             0, 0, 0, self.CellSize, nx, ny, nz, cfldt, dt)

    def _semilagAdvectVelocity(self, outs, srcs, cfldt, dt):
        nx, ny, nz = self.CellNumberX, self.CellNumberY, self.CellNumberZ
        vel = (_dp(self.VelocityU), _dp(self.VelocityV), _dp(self.VelocityW))
        for o in outs:
            o.zero_()
        for o, src, dims in zip(outs, srcs, ((1, 0, 0), (0, 1, 0), (0, 0, 1))):
            self.lib.gpu_semilag(_dp(o), _dp(src), *vel, *dims, self.CellSize, nx, ny, nz, cfldt, dt)

    def _clampExtrema(self, field, fieldTemp, dims, origin, dt):
        nx, ny, nz = self.CellNumberX, self.CellNumberY, self.CellNumberZ
        self.lib.gpu_clamp_extrema(_dp(field), _dp(fieldTemp), _dp(self.VelocityU), _dp(self.VelocityV), _dp(self.VelocityW),
                                   nx + dims[0], ny + dims[1], nz + dims[2], *dims, *origin, self.CellSize, dt)

    def _mad(self, field, f1, f2, c1, c2):
        self.lib.gpu_mad(_dp(field), _dp(f1), _dp(f2), c1, c2, field.numel())

    def _maccormack_velocity(self, srcs, cfldt, half_dt):
        """srcs advected back over half_dt with the MacCormack correction and the extrema clamp, result copied into
        the velocity (the block that appears twice in advanceReflection, BimocqGPUSolver.cpp:264-283 and :312-333)."""
        vel = (self.VelocityU, self.VelocityV, self.VelocityW)
        tmp = (self.VelocityUTemp, self.VelocityVTemp, self.VelocityWTemp)
        back = (self.TempSrcU, self.TempSrcV, self.TempSrcW)
        self._semilagAdvectVelocity(tmp, srcs, cfldt, -half_dt)
        self._semilagAdvectVelocity(back, tmp, cfldt, half_dt)
        for t, b in zip(tmp, back):
            self._add(t, b, -0.5)
        for t, s_ in zip(tmp, srcs):
            self._add(t, s_, 0.5)
        for v_, t, dims, org in zip(vel, tmp, ((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((0.5, 0.0, 0.0), (0.0, 0.5, 0.0), (0.0, 0.0, 0.5))):
            self._clampExtrema(v_, t, dims, org, half_dt)
        for v_, t in zip(vel, tmp):
            v_.copy_(t)
        self._check("MacCormack velocity advection")

    # BimocqGPUSolver.cpp:232-335
    def advanceReflection(self, framenum, dt):
        nx, ny, nz = self.CellNumberX, self.CellNumberY, self.CellNumberZ
        cfldt = self.getCFL()
        vel = (self.VelocityU, self.VelocityV, self.VelocityW)
        tmp = (self.VelocityUTemp, self.VelocityVTemp, self.VelocityWTemp)
        proj = (self.duProj, self.dvProj, self.dwProj)
        for field, ftemp in ((self.Density, self.DensityTemp), (self.Temperature, self.TemperatureTemp)):
            self._semilagAdvectField(ftemp, field, cfldt, -dt)
            self._semilagAdvectField(self.TempSrcU, ftemp, cfldt, dt)     # the reference borrows TempSrcU (u-sized) here
            self._add_n(ftemp, self.TempSrcU, -0.5, nx * ny * nz)
            self._add_n(ftemp, field, 0.5, nx * ny * nz)
            self._clampExtrema(field, ftemp, (0, 0, 0), (0.0, 0.0, 0.0), dt)
            field.copy_(ftemp)
        self._maccormack_velocity(vel, cfldt, 0.5 * dt)
        self.emitSmoke(framenum, dt)
        self.addBuoyancy(0.5 * dt)
        if self.Viscosity:
            self.diffuseField(vel[0], tmp[0], self.TempSrcU, nx + 1, ny, nz, 20, self.Viscosity, 0.5 * dt)
            self.diffuseField(vel[1], tmp[1], self.TempSrcV, nx, ny + 1, nz, 20, self.Viscosity, 0.5 * dt)
            self.diffuseField(vel[2], tmp[2], self.TempSrcW, nx, ny, nz + 1, 20, self.Viscosity, 0.5 * dt)
        for t, v_ in zip(tmp, vel):
            t.copy_(v_)
        self.projection()
        for p_, v_, t in zip(proj, vel, tmp):
            self._mad(p_, v_, t, 2.0, -1.0)                 # the reflected velocity 2 u_projected - u_before
        self._maccormack_velocity(proj, cfldt, 0.5 * dt)
        self.addBuoyancy(0.5 * dt)
        if self.Viscosity:
            self.diffuseField(vel[0], tmp[0], self.TempSrcU, nx + 1, ny, nz, 20, self.Viscosity, 0.5 * dt)
            self.diffuseField(vel[1], tmp[1], self.TempSrcV, nx, ny + 1, nz, 20, self.Viscosity, 0.5 * dt)
            self.diffuseField(vel[2], tmp[2], self.TempSrcW, nx, ny, nz + 1, 20, self.Viscosity, 0.5 * dt)
        self.projection()

    # BimocqGPUSolver.cpp:337-344
    def semilagAdvect(self, cfldt, dt):
        self._semilagAdvectVelocity((self.VelocityUTemp, self.VelocityVTemp, self.VelocityWTemp),
                                    (self.VelocityU, self.VelocityV, self.VelocityW), cfldt, dt)
        self._semilagAdvectField(self.DensityTemp, self.Density, cfldt, dt)
        self._semilagAdvectField(self.TemperatureTemp, self.Temperature, cfldt, dt)
        self._check("gpu_semilag")

    def _add_n(self, f1, f2, coeff, number):
        self.lib.gpu_add(_dp(f1), _dp(f2), coeff, number)

    # BimocqGPUSolver.cpp:128-232
    def advanceBimocq(self, framenum, dt):
        nx, ny, nz = self.CellNumberX, self.CellNumberY, self.CellNumberZ
        if framenum == 0:
            self.MaxVelocity = self.CellSize
        proj_coeff = 2.0
        cfldt = self.getCFL()
        U, V, W = self.VelocityU, self.VelocityV, self.VelocityW
        self.VelocityAdvector.updateMapping(U, V, W, cfldt, dt)
        self.ScalarAdvector.updateMapping(U, V, W, cfldt, dt)
        self.VelocityAdvector.advectVelocity(U, V, W, self.VelocityUInit, self.VelocityVInit, self.VelocityWInit,
                                             self.VelocityUPrev, self.VelocityVPrev, self.VelocityWPrev)
        self.ScalarAdvector.advectField(self.Density, self.DensityInit, self.DensityPrev)
        self.ScalarAdvector.advectField(self.Temperature, self.TemperatureInit, self.TemperaturePrev)

        self.VelocityUTemp.copy_(U); self.VelocityVTemp.copy_(V); self.VelocityWTemp.copy_(W)
        self.emitSmoke(framenum, dt)
        self.addBuoyancy(dt)
        if self.Viscosity:
            self.diffuseField(U, self.VelocityUTemp, self.TempSrcU, nx + 1, ny, nz, 20, self.Viscosity, dt)
            self.diffuseField(V, self.VelocityVTemp, self.TempSrcV, nx, ny + 1, nz, 20, self.Viscosity, dt)
            self.diffuseField(W, self.VelocityWTemp, self.TempSrcW, nx, ny, nz + 1, 20, self.Viscosity, dt)
        self._add_fields(self.duExtern, U, self.VelocityUTemp, -1.0)
        self._add_fields(self.dvExtern, V, self.VelocityVTemp, -1.0)
        self._add_fields(self.dwExtern, W, self.VelocityWTemp, -1.0)
        self.VelocityUTemp.copy_(U); self.VelocityVTemp.copy_(V); self.VelocityWTemp.copy_(W)

        self.projection()

        self.DensityTemp.copy_(self.Density); self.TemperatureTemp.copy_(self.Temperature)
        self.duProj.copy_(U); self.dvProj.copy_(V); self.dwProj.copy_(W)
        self._add(self.duProj, self.VelocityUTemp, -1.0)
        self._add(self.dvProj, self.VelocityVTemp, -1.0)
        self._add(self.dwProj, self.VelocityWTemp, -1.0)
        self.DensityExtern.copy_(self.Density); self.TemperatureExtern.copy_(self.Temperature)
        self._add(self.DensityExtern, self.DensityTemp, -1.0)
        self._add(self.TemperatureExtern, self.TemperatureTemp, -1.0)
        self._check("gpu_add")

        if framenum - self.vel_lastReinit > 10:
            self.vel_lastReinit = framenum
            proj_coeff = 1.0
        if framenum - self.scalar_lastReinit > 30:
            self.scalar_lastReinit = framenum

        va, sa = self.VelocityAdvector, self.ScalarAdvector
        init = (self.VelocityUInit, self.VelocityVInit, self.VelocityWInit)
        va.accumulateVelocity(*init, self.duExtern, self.dvExtern, self.dwExtern, 1.0)
        va.accumulateVelocity(*init, self.duProj, self.dvProj, self.dwProj, proj_coeff)
        sa.accumulateField(self.DensityInit, self.DensityExtern)
        sa.accumulateField(self.TemperatureInit, self.TemperatureExtern)
        # the reference re-initialises both mappers every frame (`if (1)`, :216-229)
        va.reinitializeMapping()
        self.velocityReinitialize()
        va.accumulateVelocity(*init, self.duProj, self.dvProj, self.dwProj, 1.0)
        sa.reinitializeMapping()
        self.scalarReinitialize()

    FIELDS = ("VelocityU", "VelocityV", "VelocityW", "Density", "Temperature", "VelocityUInit", "VelocityVInit",
              "VelocityWInit", "VelocityUPrev", "VelocityVPrev", "VelocityWPrev", "DensityInit", "TemperatureInit",
              "DensityPrev", "TemperaturePrev", "duExtern", "dvExtern", "dwExtern", "duProj", "dvProj", "dwProj", "p",
              "VelocityUTemp", "VelocityVTemp", "VelocityWTemp", "DensityTemp", "TemperatureTemp")

    def snapshot(self):
        """Host copies of every state field (tests)."""
        _torch().cuda.synchronize()
        out = {n: getattr(self, n).cpu().numpy() for n in self.FIELDS}
        for who, m in (("vel", self.VelocityAdvector), ("sca", self.ScalarAdvector)):
            for n in ("ForwardX", "ForwardY", "ForwardZ", "BackwardX", "BackwardY", "BackwardZ", "BackwardXPrev"):
                out[f"{who}.{n}"] = getattr(m, n).cpu().numpy()
        out["tempResult"] = self.tempResult.cpu().numpy()
        return out
