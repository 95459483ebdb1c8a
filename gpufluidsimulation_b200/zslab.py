"""z-slab domain decomposition of the 3D BiMocq^2 advection path over the GPUs of one box.

The reference has no multi-GPU support at all; this module is new.  One process
per GPU (torch.distributed, NCCL over NVLink/NVSwitch).  Rank r owns the global planes
[k0, k1) of every cell-centred field and of u, v; w faces [k0, k1) plus the top face nk on the
last rank.  Every rank stores its planes plus `halo` planes on both sides (bmq3d_create_slab) and
the kernels work on global indices, so a rank computes bit-for-bit what a single GPU computes.

All kernels are gathers (no scatter), so the only communication is read-side halos plus three
scalar max-reductions per step:

* CFL-local stencils (DMC sub-step, clamp)           -> fixed narrow halo (NARROW planes)
* map-indirected gathers (advect / error / apply / accumulate / forward trace / distortion)
  read at MAP VALUES, whose displacement accumulates since the last re-initialisation
                                       -> halo = ceil(D + CFL_frame) + 3 planes,
  where D = max |map_z - z| measured by the distortion kernel of the previous step and
  all-reduced.  If that exceeds the allocated halo the step raises HaloTooNarrow (loudly;
  nothing is computed from stale data).

The step is written once against a small rank/communicator interface, so the same code drives
(a) real ranks over NCCL, (b) N logical ranks on ONE GPU (LocalComm; used by the GPU tests, where
a multi-GPU box is not needed to exercise the decomposition) and (c) CPU ranks over gloo backed by
the oracle (tests only).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

NARROW = 5          # planes: reach of one DMC sub-step (<= ~3.2 cells) + stencil
VEL = ("U", "V", "W")
CUR = ("U", "V", "W", "RHO", "T")
INIT = ("U_INIT", "V_INIT", "W_INIT", "RHO_INIT", "T_INIT")
PREV = ("U_PREV", "V_PREV", "W_PREV", "RHO_PREV", "T_PREV")
ADV = ("U_ADV", "V_ADV", "W_ADV", "RHO_ADV", "T_ADV")
ERR = ("U_ERR", "V_ERR", "W_ERR", "RHO_ERR", "T_ERR")
CHANGE = ("DU_EXT", "DV_EXT", "DW_EXT", "DRHO_EXT", "DT_EXT", "DU_PROJ", "DV_PROJ", "DW_PROJ")
MAPS_BWD = tuple(f"{p}BWD_{a}" for p in "VS" for a in "XYZ")
MAPS_FWD = tuple(f"{p}FWD_{a}" for p in "VS" for a in "XYZ")
MAPS_BWDP = tuple(f"{p}BWDP_{a}" for p in "VS" for a in "XYZ")
W_TYPE = {"W", "W_INIT", "W_PREV", "W_ADV", "W_ERR", "DW_EXT", "DW_PROJ", "W_SEMI"}


class HaloTooNarrow(RuntimeError):
    pass


def slab_bounds(nk: int, world: int, rank: int):
    """Owned planes [k0, k1) of rank `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(nk, world)
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


def halo_planes(name: str, nk: int, k0: int, k1: int, world: int, rank: int, width: int):
    """Global plane ranges (lower_recv, upper_recv, send_down, send_up) of one field for an
    exchange of `width` planes; None where there is no neighbour.  The upper halo carries one
    extra plane (the +1 node of the trilinear stencil)."""
    nz = nk + (1 if name in W_TYPE else 0)
    lower_recv = upper_recv = send_down = send_up = None
    if rank > 0:
        lower_recv = (k0 - width, k0)
        send_down = (k0, min(k0 + width + 1, nz))      # becomes rank-1's upper halo
    if rank < world - 1:
        upper_recv = (k1, min(k1 + width + 1, nz))
        send_up = (k1 - width, k1)                      # becomes rank+1's lower halo
    return lower_recv, upper_recv, send_down, send_up


# ----------------------------------------------------------------------------------------------
# communicators
# ----------------------------------------------------------------------------------------------
class LocalComm:
    """All ranks live in this process (one device): halo exchange = tensor copies."""

    def __init__(self, world):
        self.world = world

    def exchange(self, ranks, names, width):
        for r in ranks:
            for name in names:
                t, p0 = r.field_with_origin(name)
                lo, up, _, _ = halo_planes(name, r.nk, r.k0, r.k1, self.world, r.rank, width)
                if lo is not None:
                    src, q0 = ranks[r.rank - 1].field_with_origin(name)
                    t[lo[0] - p0:lo[1] - p0].copy_(src[lo[0] - q0:lo[1] - q0])
                if up is not None:
                    src, q0 = ranks[r.rank + 1].field_with_origin(name)
                    t[up[0] - p0:up[1] - p0].copy_(src[up[0] - q0:up[1] - q0])

    def allreduce_max(self, per_rank_values):
        return [max(v) for v in zip(*per_rank_values)]


class DistComm:
    """One rank per process: neighbour send/recv batched per exchange, max all-reduce for scalars."""

    def __init__(self, world, rank, device):
        import torch.distributed as dist
        self.dist = dist
        self.world, self.rank, self.device = world, rank, device

    def exchange(self, ranks, names, width):
        dist = self.dist
        (r,) = ranks
        ops = []
        for name in names:
            t, p0 = r.field_with_origin(name)
            lo, up, down, upsend = halo_planes(name, r.nk, r.k0, r.k1, self.world, self.rank, width)
            if down is not None:
                ops.append(dist.P2POp(dist.isend, t[down[0] - p0:down[1] - p0], self.rank - 1))
            if upsend is not None:
                ops.append(dist.P2POp(dist.isend, t[upsend[0] - p0:upsend[1] - p0], self.rank + 1))
            if lo is not None:
                ops.append(dist.P2POp(dist.irecv, t[lo[0] - p0:lo[1] - p0], self.rank - 1))
            if up is not None:
                ops.append(dist.P2POp(dist.irecv, t[up[0] - p0:up[1] - p0], self.rank + 1))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def allreduce_max(self, per_rank_values):
        import torch
        (vals,) = per_rank_values
        t = torch.tensor(list(vals), dtype=torch.float32, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]


# ----------------------------------------------------------------------------------------------
# a rank backed by the CUDA library
# ----------------------------------------------------------------------------------------------
class CudaSlabRank:
    def __init__(self, ni, nj, nk, h, blend, rank, world, halo):
        from .solver3d import BimocqAdvection3D
        self.ni, self.nj, self.nk, self.rank, self.world, self.halo = ni, nj, nk, rank, world, halo
        self.k0, self.k1 = slab_bounds(nk, world, rank)
        self.h = float(np.float32(h))
        self.solver = BimocqAdvection3D(ni, nj, nk, h, blend, slab=(self.k0, self.k1), halo=halo)

    # -- data access
    def field_with_origin(self, name):
        _, p0, _, _, _ = self.solver.field_info(name)
        return self.solver.field(name), p0

    # -- stages (one kernel family each, owned planes only)
    def maxvel(self):
        m = C.c_float()
        self.solver.stage("maxvel", C.byref(m))
        return m.value

    def set_cfl(self, frame, gmax):
        self.solver.stage("set_cfl", int(frame), C.c_float(gmax))
        return self.solver.stats()["cfldt"]

    def dmc_substep(self, substep):
        self.solver.stage("dmc_substep", C.c_float(substep))

    def forward(self, dt):
        self.solver.stage("forward", C.c_float(dt))

    def advect(self, which):
        self.solver.stage("advect", which)

    def error(self, which):
        self.solver.stage("error", which)

    def apply(self, which):
        self.solver.stage("apply", which)

    def blend(self, which):
        self.solver.stage("blend", which)

    def distortion(self):
        a, b, c = C.c_float(), C.c_float(), C.c_float()
        self.solver.stage("distortion", C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def decide(self, frame, dt, vd2, sd2):
        self.solver.stage("decide", int(frame), C.c_float(dt), C.c_float(vd2), C.c_float(sd2))
        st = self.solver.stats()
        return bool(st["vel_reinit"]), bool(st["scalar_reinit"])

    def accumulate(self, which):
        self.solver.stage("accumulate", which)

    def reinit(self, which, phase):
        self.solver.stage("reinit", which, phase)

    def close(self):
        self.solver.close()


# ----------------------------------------------------------------------------------------------
# the step, written once
# ----------------------------------------------------------------------------------------------
class ZSlabStepper:
    """BimocqSolver::advanceBimocq's advection part (BimocqSolver.cpp:88-230) over z-slabs.
    `ranks`: the rank objects living in this process (one with DistComm, all with LocalComm)."""

    def __init__(self, ranks, comm, blend=1.0):
        self.ranks, self.comm, self.blend = ranks, comm, float(blend)
        self.disp = 0.0            # max |map_z - z| in cells, from the previous distortion stage
        self.halo = ranks[0].halo
        self.nk = ranks[0].nk
        self.h = ranks[0].h
        self.min_slab = min(slab_bounds(self.nk, comm.world, r)[1] - slab_bounds(self.nk, comm.world, r)[0]
                            for r in range(comm.world))
        self.stats = {}
        self.reinit_count = [0, 0]

    def _width(self, want):
        w = int(want)
        if w > self.halo or w + 1 > self.min_slab:
            raise HaloTooNarrow(f"need {w} halo planes (map displacement {self.disp:.2f} cells), allocated {self.halo}, "
                                f"smallest slab {self.min_slab}")
        return w

    def _each(self, fn):
        return [fn(r) for r in self.ranks]

    def advect(self, frame, dt):
        comm, ranks = self.comm, self.ranks
        dt = float(np.float32(dt))
        (gmax,) = comm.allreduce_max(self._each(lambda r: (r.maxvel(),)))
        cfldt = self._each(lambda r: r.set_cfl(frame, gmax))[0]
        cfl_frame = dt * max(gmax, 1e-4) / self.h
        wide = self._width(math.ceil(self.disp + cfl_frame) + 3)
        self.stats.update(max_abs_vel=gmax, cfldt=cfldt, halo_used=wide)
        comm.exchange(ranks, VEL, wide)
        # updateBackward (Mapping.cpp:354-368): the reference's float sub-step loop
        T = np.float32(0.0); sub = np.float32(cfldt); dt32 = np.float32(dt)
        n = 0
        comm.exchange(ranks, MAPS_BWD, self._width(NARROW))
        while T < dt32:
            if T + sub > dt32:
                sub = np.float32(dt32 - T)
            self._each(lambda r: r.dmc_substep(float(sub)))
            T = np.float32(T + sub)
            n += 1
            last = not (T < dt32)
            comm.exchange(ranks, MAPS_BWD, wide if last else self._width(NARROW))
        self.stats["n_substeps"] = n
        self._each(lambda r: r.forward(dt))          # psi is read at the own cell only: no halo needed
        comm.exchange(ranks, MAPS_FWD, wide)
        comm.exchange(ranks, INIT, wide)
        for which, sl in ((0, slice(0, 3)), (1, slice(3, 5))):
            self._each(lambda r: r.advect(which))
            comm.exchange(ranks, ADV[sl], wide)
            self._each(lambda r: r.error(which))
            comm.exchange(ranks, ERR[sl], wide)
            self._each(lambda r: r.apply(which))
            if self.blend != 1.0 and self.reinit_count[which] > 0:
                full = self._width(self.halo)
                comm.exchange(ranks, PREV[sl], full)
                comm.exchange(ranks, MAPS_BWDP[which * 3:which * 3 + 3], full)
                self._each(lambda r: r.blend(which))

    def accumulate(self, frame, dt):
        comm, ranks = self.comm, self.ranks
        dt = float(np.float32(dt))
        wide = self.stats.get("halo_used", self._width(3))
        vd2, sd2, disp = comm.allreduce_max(self._each(lambda r: r.distortion()))
        self.disp = disp
        dec = self._each(lambda r: r.decide(frame, dt, vd2, sd2))
        vel_reinit, sca_reinit = dec[0]
        comm.exchange(ranks, CHANGE, wide)
        self._each(lambda r: r.accumulate(0))
        self._each(lambda r: r.accumulate(1))
        if vel_reinit:
            self._each(lambda r: r.reinit(0, 0))
            self._each(lambda r: r.reinit(0, 1))
            self.reinit_count[0] += 1
        if sca_reinit:
            self._each(lambda r: r.reinit(1, 0))
            self.reinit_count[1] += 1
        if vel_reinit and sca_reinit:
            self.disp = 0.0
        self.stats.update(vel_reinit=vel_reinit, scalar_reinit=sca_reinit, max_disp_z=disp,
                          vel_d2=vd2, scalar_d2=sd2)


# ----------------------------------------------------------------------------------------------
# user-facing wrapper for one process per GPU (bench.py, multi-GPU tests)
# ----------------------------------------------------------------------------------------------
class ZSlabAdvection3D:
    def __init__(self, ni, nj, nk, h, blend_coeff=1.0, rank=0, world=1, halo=24):
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        self.r = CudaSlabRank(ni, nj, nk, h, blend_coeff, rank, world, halo)
        self.comm = DistComm(world, rank, torch.device("cuda", torch.cuda.current_device()))
        self.stepper = ZSlabStepper([self.r], self.comm, blend_coeff)
        self.lib = self.r.solver.lib

    def _own_slice(self, name, full):
        """The planes of a whole-grid tensor that this rank stores."""
        _, p0, npl, _, _ = self.r.solver.field_info(name)
        return full[p0:p0 + npl]

    def set_initial_device(self, u, v, w, rho, T):
        for n, a in zip(CUR, (u, v, w, rho, T)):
            self.r.solver.field(n).copy_(self._own_slice(n, a))
        self.torch.cuda.synchronize()
        self.r.solver.reset()

    def field(self, name):
        return self.r.solver.field(name)

    def advect(self, frame, dt):
        self.stepper.advect(frame, dt)

    def accumulate(self, frame, dt):
        self.stepper.accumulate(frame, dt)

    def apply_buoyancy(self, beta, dt, alpha=0.0):
        # local operation on the stored planes (owned planes are what matters; halos are refreshed
        # by the exchange that precedes every consumer)
        self.r.solver.apply_buoyancy(beta, dt, alpha)

    def stats(self):
        st = self.r.solver.stats()
        st.update(self.stepper.stats)
        return st

    def timing_enable(self, on=True):
        self.r.solver.timing_enable(on)

    def timing_read(self):
        return self.r.solver.timing_read()

    def close(self):
        self.r.close()
