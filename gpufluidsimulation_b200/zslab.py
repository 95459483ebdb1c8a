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
  where D = max |map_z - z| of THAT MAPPER's maps, measured by the distortion kernel of the
  previous step and all-reduced.  The velocity mapper is re-initialised at least every 10 frames,
  the scalar mapper every 30 (BimocqSolver.cpp:175-185), so the two sets of fields are exchanged
  with different widths.  A halo may be wider than the neighbouring slab: the planes then come
  from the ranks that own them (halo_segments).  If a width exceeds the allocated halo, every
  rank re-allocates its fields with a wider one (bmq3d_grow_halo; the decision is taken from
  all-reduced numbers, so all ranks take it in the same step) -- HaloTooNarrow is raised only by
  rank types that cannot grow.  Nothing is ever computed from stale data.

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
GROW_SLACK = 8      # extra planes allocated when a halo has to grow (so that it grows rarely)
VEL = ("U", "V", "W")
CUR = ("U", "V", "W", "RHO", "T")
INIT = ("U_INIT", "V_INIT", "W_INIT", "RHO_INIT", "T_INIT")
PREV = ("U_PREV", "V_PREV", "W_PREV", "RHO_PREV", "T_PREV")
ADV = ("U_ADV", "V_ADV", "W_ADV", "RHO_ADV", "T_ADV")
ERR = ("U_ERR", "V_ERR", "W_ERR", "RHO_ERR", "T_ERR")
CHANGE = ("DU_EXT", "DV_EXT", "DW_EXT", "DRHO_EXT", "DT_EXT", "DU_PROJ", "DV_PROJ", "DW_PROJ")
CHANGE_V = ("DU_EXT", "DV_EXT", "DW_EXT", "DU_PROJ", "DV_PROJ", "DW_PROJ")
CHANGE_S = ("DRHO_EXT", "DT_EXT")
MAPS_BWD = tuple(f"{p}BWD_{a}" for p in "VS" for a in "XYZ")
MAPS_FWD = tuple(f"{p}FWD_{a}" for p in "VS" for a in "XYZ")
MAPS_BWDP = tuple(f"{p}BWDP_{a}" for p in "VS" for a in "XYZ")
W_TYPE = {"W", "W_INIT", "W_PREV", "W_ADV", "W_ERR", "DW_EXT", "DW_PROJ", "W_SEMI"}


class HaloTooNarrow(RuntimeError):
    pass


class PeerUnavailable(RuntimeError):
    pass


def slab_bounds(nk: int, world: int, rank: int):
    """Owned planes [k0, k1) of rank `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(nk, world)
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


def default_halo(cfl_frame: float = 1.5) -> int:
    """Halo planes to allocate so that the step does not have to grow them in the common case: the
    scalar mapper is re-initialised after at most 31 frames (BimocqSolver.cpp:181), a map point
    moves at most CFL_frame cells per frame; plus this frame's reach and the stencil."""
    return int(math.ceil(32 * cfl_frame)) + 3


def halo_planes(name: str, nk: int, k0: int, k1: int, world: int, rank: int, width: int):
    """Global plane ranges (lower_recv, upper_recv, send_down, send_up) of one field for an
    exchange of `width` planes with the DIRECT neighbours; None where there is no neighbour.  The
    upper halo carries one extra plane (the +1 node of the trilinear stencil).  Geometry helper;
    the exchanges themselves use halo_segments, which also reaches past the direct neighbour."""
    nz = nk + (1 if name in W_TYPE else 0)
    lower_recv = upper_recv = send_down = send_up = None
    if rank > 0:
        lower_recv = (k0 - width, k0)
        send_down = (k0, min(k0 + width + 1, nz))      # becomes rank-1's upper halo
    if rank < world - 1:
        upper_recv = (k1, min(k1 + width + 1, nz))
        send_up = (k1 - width, k1)                      # becomes rank+1's lower halo
    return lower_recv, upper_recv, send_down, send_up


def owned_range(name: str, nk: int, world: int, rank: int):
    """Global planes [a, b) of field `name` that rank `rank` owns (w faces: the top face nk belongs
    to the last rank)."""
    k0, k1 = slab_bounds(nk, world, rank)
    return k0, k1 + (1 if name in W_TYPE and rank == world - 1 else 0)


def halo_segments(name: str, nk: int, world: int, rank: int, width: int):
    """What rank `rank` receives for a halo of `width` planes of field `name`:
    [(source rank, a, b)] with [a, b) global planes owned by the source.  The halo may be wider than
    the neighbouring slab: segments then come from ranks further away."""
    nz = nk + (1 if name in W_TYPE else 0)
    a0, b0 = owned_range(name, nk, world, rank)
    want = []
    if rank > 0:
        want.append((max(0, a0 - width), a0))
    if rank < world - 1:
        want.append((b0, min(b0 + width + 1, nz)))
    out = []
    for lo, hi in want:
        for q in range(world):
            if q == rank:
                continue
            qa, qb = owned_range(name, nk, world, q)
            a, b = max(lo, qa), min(hi, qb)
            if a < b:
                out.append((q, a, b))
    return out


# ----------------------------------------------------------------------------------------------
# communicators.  An exchange is a list of GROUPS [(field names, width), ...] moved under one
# synchronisation point (the two mappers' fields travel together but with their own widths).
# ----------------------------------------------------------------------------------------------
class LocalComm:
    """All ranks live in this process (one device): halo exchange = tensor copies."""

    def __init__(self, world):
        self.world = world

    def exchange(self, ranks, groups):
        for r in ranks:
            for names, width in groups:
                for name in names:
                    t, p0 = r.field_with_origin(name)
                    for q, a, b in halo_segments(name, r.nk, self.world, r.rank, width):
                        src, q0 = ranks[q].field_with_origin(name)
                        t[a - p0:b - p0].copy_(src[a - q0:b - q0])

    def exchange_async(self, ranks, groups):
        self.exchange(ranks, groups)      # one process, one stream: nothing to overlap
        return None

    def wait(self, handle):
        pass

    def allreduce_max(self, per_rank_values):
        return [max(v) for v in zip(*per_rank_values)]

    def rebuild(self, ranks, regrow):
        regrow()


class DistComm:
    """One rank per process: send/recv batched per exchange, max all-reduce for scalars."""

    def __init__(self, world, rank, device):
        import torch.distributed as dist
        self.dist = dist
        self.world, self.rank, self.device = world, rank, device

    def exchange_async(self, ranks, groups):
        """Posts the sends/receives of one halo exchange on the communicator's stream and returns
        the requests; the transfer overlaps whatever is launched before wait().  Every rank walks
        the same global list of (destination, source, planes), so sends and receives pair up."""
        dist = self.dist
        (r,) = ranks
        ops = []
        for names, width in groups:
            for name in names:
                t, p0 = r.field_with_origin(name)
                for dst in range(self.world):
                    for src, a, b in halo_segments(name, r.nk, self.world, dst, width):
                        if src == self.rank:
                            ops.append(dist.P2POp(dist.isend, t[a - p0:b - p0], dst))
                        elif dst == self.rank:
                            ops.append(dist.P2POp(dist.irecv, t[a - p0:b - p0], src))
        return dist.batch_isend_irecv(ops) if ops else []

    def exchange(self, ranks, groups):
        self.wait(self.exchange_async(ranks, groups))

    def wait(self, handle):
        for req in handle or ():
            req.wait()

    def allreduce_max(self, per_rank_values):
        import torch
        (vals,) = per_rank_values
        t = torch.tensor(list(vals), dtype=torch.float32, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def rebuild(self, ranks, regrow):
        regrow()


class PeerComm(DistComm):
    """Halo exchange by direct peer-to-peer copies over NVLink instead of NCCL send/recv.

    Every rank exports CUDA IPC handles of all its field allocations once; every other rank maps
    them.  An exchange is then (1) a stream-ordered barrier -- a one-element NCCL all-reduce on the
    COPY stream (which first waits for this rank's producer kernels), so every rank's producers have
    finished, without blocking any host or the compute stream -- and (2) each rank
    PULLING its halo planes straight out of the owners' planes with cudaMemcpyAsync on that
    copy stream (measured: NCCL send/recv of these 10 MB blocks reaches < 100 GB/s on
    this box, a peer copy ~700 GB/s).  The copy stream is joined to the compute stream by events,
    so an exchange posted with exchange_async overlaps the kernels launched before wait().

    Safety of reading another rank's buffer without a second barrier: a source buffer is only
    rewritten by its owner after at least one later exchange (barrier), and a rank's pulls are
    ordered before its own arrival at that barrier (zslab.ZSlabStepper schedule)."""

    # every allocation a rank owns, in an order all ranks share (buffers rotate in lock-step)
    ALLOC_NAMES = CUR + INIT + PREV + CHANGE + MAPS_FWD + MAPS_BWD + MAPS_BWDP + ADV + ERR + tuple(
        f"TMPMAP{i}" for i in range(6))

    def __init__(self, world, rank, device, slab_rank):
        super().__init__(world, rank, device)
        import torch
        self.torch = torch
        self.r = slab_rank
        self.lib = slab_rank.solver.lib
        self.copy_stream = torch.cuda.Stream(device=device)
        self.flag = torch.zeros(1, dtype=torch.float32, device=device)
        self.index_of_ptr = {}
        self.peer = {}
        self._connect()

    def _connect(self):
        """Collective: export this rank's allocations, map everybody else's.  Every rank takes part
        in every collective below even if a local step failed, and the outcome is agreed on with a
        MIN all-reduce, so that either all ranks use peer copies or all of them raise
        PeerUnavailable (the caller then falls back to NCCL send/recv)."""
        from .capi import BimocqLibraryError, check
        torch, world, rank, slab_rank = self.torch, self.world, self.rank, self.r
        n = len(self.ALLOC_NAMES)
        handles = torch.zeros((n, 64), dtype=torch.uint8)
        self.index_of_ptr = {}
        self.peer = {}
        ok = 1.0
        try:
            for i, name in enumerate(self.ALLOC_NAMES):
                ptr = slab_rank.solver.field_info(name)[0]
                buf = (C.c_ubyte * 64)()
                check(self.lib.bmq_ipc_export(C.c_void_p(ptr), buf), "bmq_ipc_export")
                handles[i] = torch.frombuffer(bytearray(buf), dtype=torch.uint8)
                self.index_of_ptr[ptr] = i
        except BimocqLibraryError:
            ok = 0.0
        mine = handles.to(self.device)
        everyone = [torch.zeros_like(mine) for _ in range(world)]
        self.dist.all_gather(everyone, mine)
        gathered = [g.cpu() for g in everyone]
        try:
            for nb in range(world):
                if nb != rank and ok:
                    ptrs = []
                    self.peer[nb] = ptrs
                    for i in range(n):
                        raw = (C.c_ubyte * 64).from_buffer_copy(bytes(gathered[nb][i].tolist()))
                        out = C.c_void_p()
                        check(self.lib.bmq_ipc_open(raw, C.byref(out)), "bmq_ipc_open")   # lazy peer access
                        ptrs.append(out.value)
        except BimocqLibraryError:
            ok = 0.0
        agreed = torch.tensor([ok], dtype=torch.float32, device=self.device)
        self.dist.all_reduce(agreed, op=self.dist.ReduceOp.MIN)
        if float(agreed.item()) < 1.0:
            self.close()
            self.lib.bmq_clear_error()
            raise PeerUnavailable("CUDA IPC peer mapping failed on at least one rank")

    def rebuild(self, ranks, regrow):
        """bmq3d_grow_halo moves every allocation: finish all pulls, unmap the peers' buffers, barrier
        (nobody frees a buffer that a peer still has mapped or in use), re-allocate, export and map again."""
        self.torch.cuda.synchronize()
        self.close()
        self.dist.barrier()
        regrow()
        self._connect()

    @staticmethod
    def _check(status, what):
        from .capi import check
        check(status, what)

    def _stored_origin(self, rank):
        k0, _ = slab_bounds(self.r.nk, self.world, rank)
        return max(0, k0 - self.r.halo)

    def exchange_async(self, ranks, groups):
        torch = self.torch
        (r,) = ranks
        # Everything below runs on the copy stream, so the compute stream is never held up by the barrier:
        # (1) the copy stream waits for this rank's producers (all work queued on the compute stream so far),
        # (2) barrier in stream order -- a one-element all-reduce: afterwards every rank's producers are complete,
        # (3) pull the halo planes out of their owners' memory.
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.copy_stream.wait_event(ev)
        with torch.cuda.stream(self.copy_stream):
            self.dist.all_reduce(self.flag)
        cs = C.c_void_p(self.copy_stream.cuda_stream)
        for names, width in groups:
            for name in names:
                ptr, p0, npl, nx, ny = r.solver.field_info(name)
                idx = self.index_of_ptr[ptr]
                plane = nx * ny
                for q, a, b in halo_segments(name, r.nk, self.world, self.rank, width):
                    q0 = self._stored_origin(q)
                    src = self.peer[q][idx] + 4 * plane * (a - q0)
                    dst = ptr + 4 * plane * (a - p0)
                    st = self.lib.bmq_copy_async(C.c_void_p(dst), C.c_void_p(src), 4 * plane * (b - a), cs)
                    if st != 0:
                        self._check(st, "bmq_copy_async")
        done = torch.cuda.Event()
        done.record(self.copy_stream)
        return done

    def wait(self, handle):
        if handle is not None:
            self.torch.cuda.current_stream().wait_event(handle)

    def close(self):
        for ptrs in self.peer.values():
            for p in ptrs:
                if p:
                    self.lib.bmq_ipc_close(C.c_void_p(p))
        self.peer = {}


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
        self._views = {}

    # -- data access.  Buffers rotate (ping-pong, re-initialisation) but their set is fixed, so
    # the torch views are cached by device pointer instead of being rebuilt on every exchange.
    def field_with_origin(self, name):
        ptr, p0, npl, nx, ny = self.solver.field_info(name)
        key = (ptr, npl, ny, nx)
        t = self._views.get(key)
        if t is None:
            t = self.solver.field(name)
            self._views[key] = t
        return t, p0

    def grow_halo(self, new_halo):
        """Re-allocate every field with a wider halo (owned planes and old halos are kept)."""
        self._views = {}
        self.solver.grow_halo(new_halo)
        self.halo = new_halo

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
        """(velocity d^2, scalar d^2, velocity-map z displacement, scalar-map z displacement)"""
        a, b, c, d = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        self.solver.stage("distortion2", C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return a.value, b.value, c.value, d.value

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
        # max |map_z - z| in cells per mapper (velocity, scalar): current maps (from the previous
        # distortion stage) and the maps frozen into chi_prev at the last re-initialisation
        self.disp = [0.0, 0.0]
        self.disp_prev = [0.0, 0.0]
        self.halo = ranks[0].halo
        self.nk = ranks[0].nk
        self.h = ranks[0].h
        self.stats = {}
        self.reinit_count = [0, 0]
        self.grow_count = 0
        import os
        self.profile = bool(os.environ.get("BMQ_ZSLAB_PROFILE"))
        self.prof = {}
        if self.profile:
            ex, red, exa = comm.exchange, comm.allreduce_max, comm.exchange_async
            # keyed by the first field of the exchange, so that the profile names the exchange that costs most
            comm.exchange = lambda *a: self._timed("exchange(blocking):" + a[1][0][0][0], lambda: ex(*a))
            comm.exchange_async = lambda *a: self._timed("exchange(posted, then synchronised by the profiler):" + a[1][0][0][0], lambda: exa(*a))
            comm.allreduce_max = lambda *a: self._timed("allreduce", lambda: red(*a))
            each = self._each
            self._each = lambda fn: self._timed("stages", lambda: each(fn))

    def _width(self, want):
        """Halo planes for a reach of `want`; grows every rank's halo when the allocation is too
        narrow.  `want` derives from all-reduced numbers only, so all ranks grow in the same call."""
        w = int(want)
        if w > self.halo:
            if not all(hasattr(r, "grow_halo") for r in self.ranks):
                raise HaloTooNarrow(f"need {w} halo planes (map displacement {max(self.disp):.2f} cells), "
                                    f"allocated {self.halo}")
            new = w + GROW_SLACK
            self.comm.rebuild(self.ranks, lambda: [r.grow_halo(new) for r in self.ranks])
            self.halo = new
            self.grow_count += 1
        return w

    def _each(self, fn):
        return [fn(r) for r in self.ranks]

    # optional host-side profile (BMQ_ZSLAB_PROFILE=1): wall time per category with a device
    # synchronisation around every item -- for finding overheads, never for reported numbers
    def _timed(self, kind, fn):
        if not self.profile:
            return fn()
        import time

        import torch
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        self.prof[kind] = self.prof.get(kind, 0.0) + (time.perf_counter() - t0)
        return out

    def advect(self, frame, dt):
        """Phase A.  Halo exchanges are posted as soon as their producer has been launched and
        waited for right before their first consumer, so that only the velocity halo and the
        per-sub-step chi halo are exposed; everything else overlaps the next stage's kernel."""
        comm, ranks = self.comm, self.ranks
        dt = float(np.float32(dt))
        (gmax,) = comm.allreduce_max(self._each(lambda r: (r.maxvel(),)))
        cfldt = self._each(lambda r: r.set_cfl(frame, gmax))[0]
        cfl_frame = dt * max(gmax, 1e-4) / self.h
        # widest first, so that a growth happens before anything is in flight
        need = [math.ceil(self.disp[m] + cfl_frame) + 3 for m in (0, 1)]
        blend_on = [self.blend != 1.0 and self.reinit_count[m] > 0 for m in (0, 1)]
        need_b = [math.ceil(self.disp_prev[m] + self.disp[m] + cfl_frame) + 3 if blend_on[m] else 0 for m in (0, 1)]
        self._width(max(need + need_b + [NARROW]))
        wv, ws = self._width(need[0]), self._width(need[1])
        wmax = max(wv, ws)
        narrow = self._width(NARROW)
        self.stats.update(max_abs_vel=gmax, cfldt=cfldt, halo_used=wmax, halo_vel=wv, halo_scalar=ws,
                          halo_allocated=self.halo)
        both = lambda names5: [(names5[0:3], wv), (names5[3:5], ws)]
        maps = lambda names6, a, b: [(names6[0:3], a), (names6[3:6], b)]
        comm.exchange(ranks, [(VEL, wmax)])                   # consumers: DMC (next kernel), forward
        comm.exchange(ranks, [(MAPS_BWD, narrow)])
        h_init = comm.exchange_async(ranks, both(INIT))       # consumer: advect; overlaps DMC + forward
        # updateBackward (Mapping.cpp:354-368): the reference's float sub-step loop
        T = np.float32(0.0); sub = np.float32(cfldt); dt32 = np.float32(dt)
        n = 0
        h_bwd = None
        while T < dt32:
            if T + sub > dt32:
                sub = np.float32(dt32 - T)
            self._each(lambda r: r.dmc_substep(float(sub)))
            T = np.float32(T + sub)
            n += 1
            if T < dt32:
                comm.exchange(ranks, [(MAPS_BWD, narrow)])    # consumer: the next sub-step
            else:
                h_bwd = comm.exchange_async(ranks, maps(MAPS_BWD, wv, ws))   # overlaps forward
        self.stats["n_substeps"] = n
        self._each(lambda r: r.forward(dt))                   # psi is read at the own cell only
        h_fwd = comm.exchange_async(ranks, maps(MAPS_FWD, wv, ws))   # consumer: error; overlaps advect
        comm.wait(h_bwd)
        comm.wait(h_init)
        self._each(lambda r: r.advect(0))
        h_av = comm.exchange_async(ranks, [(ADV[0:3], wv)])   # overlaps advect(scalars)
        self._each(lambda r: r.advect(1))
        h_as = comm.exchange_async(ranks, [(ADV[3:5], ws)])   # overlaps error(velocity)
        comm.wait(h_fwd)
        comm.wait(h_av)
        self._each(lambda r: r.error(0))
        h_ev = comm.exchange_async(ranks, [(ERR[0:3], wv)])   # overlaps error(scalars)
        comm.wait(h_as)
        self._each(lambda r: r.error(1))
        h_es = comm.exchange_async(ranks, [(ERR[3:5], ws)])   # overlaps apply(velocity)
        comm.wait(h_ev)
        self._each(lambda r: r.apply(0))
        comm.wait(h_es)
        self._each(lambda r: r.apply(1))
        for which, sl in ((0, slice(0, 3)), (1, slice(3, 5))):
            if blend_on[which]:
                # chi_prev(chi(x)) reaches the displacement frozen at the last reinit on top of the
                # current one
                wb = self._width(need_b[which])
                comm.exchange(ranks, [(PREV[sl], wb), (MAPS_BWDP[which * 3:which * 3 + 3], wb)])
                self._each(lambda r: r.blend(which))

    def accumulate(self, frame, dt):
        comm, ranks = self.comm, self.ranks
        dt = float(np.float32(dt))
        wv = self.stats.get("halo_vel", 3)
        ws = self.stats.get("halo_scalar", 3)
        h_ch = comm.exchange_async(ranks, [(CHANGE_V, wv), (CHANGE_S, ws)])   # overlaps the distortion kernel
        vd2, sd2, disp_v, disp_s = comm.allreduce_max(self._each(lambda r: r.distortion()))
        self.disp = [disp_v, disp_s]
        dec = self._each(lambda r: r.decide(frame, dt, vd2, sd2))
        vel_reinit, sca_reinit = dec[0]
        comm.wait(h_ch)
        self._each(lambda r: r.accumulate(0))
        self._each(lambda r: r.accumulate(1))
        if vel_reinit:
            self._each(lambda r: r.reinit(0, 0))
            self._each(lambda r: r.reinit(0, 1))
            self.reinit_count[0] += 1
            self.disp_prev[0], self.disp[0] = self.disp[0], 0.0
        if sca_reinit:
            self._each(lambda r: r.reinit(1, 0))
            self.reinit_count[1] += 1
            self.disp_prev[1], self.disp[1] = self.disp[1], 0.0
        self.stats.update(vel_reinit=vel_reinit, scalar_reinit=sca_reinit, max_disp_z=max(disp_v, disp_s),
                          disp_z_vel=disp_v, disp_z_scalar=disp_s, vel_d2=vd2, scalar_d2=sd2,
                          halo_grown=self.grow_count)


# ----------------------------------------------------------------------------------------------
# the same step inside the library: bmq3d_mg_* (include/bimocq_b200.h), driven from here with
# torch.distributed supplying the two collectives the C layer asks its host for
# ----------------------------------------------------------------------------------------------
class NativeSlab:
    """One rank of the C-level z-slab driver.  `collectives` = (allreduce_max(list[float]) -> list[float],
    stream_barrier(cuda_stream_ptr), host_barrier(), all_gather_bytes(bytes) -> list[bytes])."""

    def __init__(self, ni, nj, nk, h, blend, rank, world, halo, collectives):
        from . import capi
        from .solver3d import BimocqAdvection3D
        self.capi, self.lib = capi, capi.load_library()
        self.ni, self.nj, self.nk, self.rank, self.world = ni, nj, nk, rank, world
        self.k0, self.k1 = slab_bounds(nk, world, rank)
        self.h = float(np.float32(h))
        self.allreduce_max, self.stream_barrier, self.host_barrier, self.all_gather_bytes = collectives
        self._mg = C.c_void_p()
        capi.check(self.lib.bmq3d_mg_create(ni, nj, nk, self.h, float(blend), rank, world, int(halo), C.byref(self._mg)), "bmq3d_mg_create")
        sh = C.c_void_p()
        capi.check(self.lib.bmq3d_mg_solver(self._mg, C.byref(sh)), "bmq3d_mg_solver")
        self.solver = BimocqAdvection3D(ni, nj, nk, h, blend, borrowed_handle=sh)
        self.error = None

        def _reduce(vals, n, _ctx):
            try:
                out = self.allreduce_max([vals[q] for q in range(n)])
                for q in range(n):
                    vals[q] = out[q]
                return 0
            except BaseException as exc:   # noqa: BLE001 -- must not unwind through C
                self.error = exc
                return 1

        def _barrier(stream, _ctx):
            try:
                self.stream_barrier(stream)
                return 0
            except BaseException as exc:   # noqa: BLE001
                self.error = exc
                return 1

        self._cb = (capi.ALLREDUCE_MAX_FN(_reduce), capi.STREAM_BARRIER_FN(_barrier))     # keep them alive
        capi.check(self.lib.bmq3d_mg_set_collectives(self._mg, self._cb[0], self._cb[1], None), "bmq3d_mg_set_collectives")
        self.connect()

    @property
    def halo(self):
        return self.mg_stats()["halo_allocated"]

    def connect(self):
        if self.world == 1:
            return
        n = C.c_size_t()
        self.capi.check(self.lib.bmq3d_mg_export_size(self._mg, C.byref(n)), "bmq3d_mg_export_size")
        blob = (C.c_ubyte * n.value)()
        self.capi.check(self.lib.bmq3d_mg_export(self._mg, blob), "bmq3d_mg_export")
        everyone = b"".join(self.all_gather_bytes(bytes(blob)))
        assert len(everyone) == n.value * self.world
        self.capi.check(self.lib.bmq3d_mg_connect(self._mg, everyone), "bmq3d_mg_connect")

    def _check(self, status, what):
        if status != 0 and self.error is not None:
            err, self.error = self.error, None
            raise err
        self.capi.check(status, what)

    def advect(self, frame, dt):
        dt = float(np.float32(dt))
        st = self.lib.bmq3d_mg_advect(self._mg, int(frame), dt)
        if st == 3:      # BMQ_ERR_HALO: every rank sees it in the same call (the width comes from all-reduced numbers)
            self.lib.bmq_clear_error()
            need = self.mg_stats()["halo_needed"]
            self.capi.check(self.lib.bmq3d_mg_disconnect(self._mg), "bmq3d_mg_disconnect")
            self.host_barrier()      # nobody frees a buffer that a peer still has mapped
            self.capi.check(self.lib.bmq3d_mg_grow_halo(self._mg, need + GROW_SLACK), "bmq3d_mg_grow_halo")
            self.connect()
            st = self.lib.bmq3d_mg_advect(self._mg, int(frame), dt)
        self._check(st, "bmq3d_mg_advect")

    def accumulate(self, frame, dt):
        self._check(self.lib.bmq3d_mg_accumulate(self._mg, int(frame), float(np.float32(dt))), "bmq3d_mg_accumulate")

    def mg_stats(self):
        st = self.capi.MgStats()
        self.capi.check(self.lib.bmq3d_mg_get_stats(self._mg, C.byref(st)), "bmq3d_mg_get_stats")
        return st.as_dict()

    def field_with_origin(self, name):
        _, p0, _, _, _ = self.solver.field_info(name)
        return self.solver.field(name), p0

    def close(self):
        if self._mg:
            self.solver.close()
            self.lib.bmq3d_mg_destroy(self._mg)
            self._mg = None


class _NativeStepper:
    """What bench.py and the tests read off a stepper, for the native driver."""

    def __init__(self, slab):
        self.slab, self.prof, self.profile, self.stats = slab, {}, False, {}

    @property
    def halo(self):
        return self.slab.halo

    @property
    def grow_count(self):
        return self.slab.mg_stats()["halo_grown"]

    def advect(self, frame, dt):
        self.slab.advect(frame, dt)
        st = self.slab.mg_stats()
        self.stats.update(halo_used=max(st["halo_vel"], st["halo_scalar"]), halo_vel=st["halo_vel"], halo_scalar=st["halo_scalar"],
                          halo_allocated=st["halo_allocated"], n_substeps=self.slab.solver.stats()["n_substeps"],
                          signalling={0: "host callbacks + copy engines", 1: "device flags + pull kernel",
                                      2: "device flags + copy engines"}.get(st["signalling"], st["signalling"]))

    def accumulate(self, frame, dt):
        self.slab.accumulate(frame, dt)
        st = self.slab.solver.stats()
        self.stats.update(vel_reinit=bool(st["vel_reinit"]), scalar_reinit=bool(st["scalar_reinit"]), max_disp_z=st["max_disp_z"],
                          disp_z_vel=st["max_disp_z_vel"], disp_z_scalar=st["max_disp_z_scalar"], halo_grown=self.grow_count)


def torch_collectives(device):
    """The collectives NativeSlab needs, on torch.distributed (NCCL)."""
    import torch
    import torch.distributed as dist
    flag = torch.zeros(1, dtype=torch.float32, device=device)
    streams = {}

    def allreduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def stream_barrier(stream_ptr):
        st = streams.get(stream_ptr)
        if st is None:
            st = streams[stream_ptr] = torch.cuda.ExternalStream(stream_ptr, device=device)
        with torch.cuda.stream(st):
            dist.all_reduce(flag)

    def host_barrier():
        torch.cuda.synchronize()
        dist.barrier()

    def all_gather_bytes(b):
        mine = torch.frombuffer(bytearray(b), dtype=torch.uint8).to(device)
        out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(out, mine)
        return [bytes(o.cpu().numpy().tobytes()) for o in out]

    return allreduce_max, stream_barrier, host_barrier, all_gather_bytes


# ----------------------------------------------------------------------------------------------
# user-facing wrapper for one process per GPU (bench.py, multi-GPU tests)
# ----------------------------------------------------------------------------------------------
class ZSlabAdvection3D:
    def __init__(self, ni, nj, nk, h, blend_coeff=1.0, rank=0, world=1, halo=None, transport="peer",
                 cfl_frame=1.5):
        """halo: planes allocated on both sides of the slab; None = default_halo(cfl_frame), the
        reach of the scalar mapper's 30-frame re-initialisation cap (it grows on demand anyway).
        transport: "native" = the C-level driver bmq3d_mg_* (peer copies over NVLink issued by the library;
        torch.distributed supplies the scalar all-reduce and the stream barrier), "peer" = direct P2P copies of peer-mapped memory over NVLink (PeerComm; if the
        peer mapping cannot be set up on every rank, all ranks fall back to NCCL and say so on
        stderr), "nccl" = NCCL send/recv (DistComm)."""
        import torch
        self.torch = torch
        self.rank, self.world = rank, world
        if halo is None:
            halo = default_halo(cfl_frame)
        halo = min(int(halo), nk)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.transport = transport
        if transport == "native":
            # the whole z-slab step inside the library (bmq3d_mg_*); torch.distributed only supplies the collectives
            self.r = NativeSlab(ni, nj, nk, h, blend_coeff, rank, world, halo, torch_collectives(dev))
            self.comm = None
            self.stepper = _NativeStepper(self.r)
            self.lib = self.r.solver.lib
            return
        self.r = CudaSlabRank(ni, nj, nk, h, blend_coeff, rank, world, halo)
        if transport == "peer":
            try:
                self.comm = PeerComm(world, rank, dev, self.r)
            except PeerUnavailable as exc:
                import sys
                if rank == 0:
                    print(f"zslab: {exc}; using NCCL send/recv for the halo exchange", file=sys.stderr)
                self.comm = DistComm(world, rank, dev)
                self.transport = "nccl"
        else:
            self.comm = DistComm(world, rank, dev)
        self.stepper = ZSlabStepper([self.r], self.comm, blend_coeff)
        self.lib = self.r.solver.lib

    @property
    def halo(self):
        return self.stepper.halo

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

    # ---- host-buffer step (what bmq3d_advect_host / bmq3d_accumulate_host are on one GPU): every rank
    # moves only the planes it owns across its own PCIe link
    def owned(self, name):
        """torch view of the planes of `name` this rank owns (w-type fields: the top face belongs to
        the last rank)."""
        _, p0, _, _, _ = self.r.solver.field_info(name)
        kb, ke = self.r.k0, self.r.k1 + (1 if name in W_TYPE and self.r.k1 == self.r.nk else 0)
        return self.r.solver.field(name)[kb - p0:ke - p0]

    def alloc_host(self):
        """Pinned host buffers for the owned planes of u, v, w, rho, T."""
        return [self.torch.empty(tuple(self.owned(n).shape), dtype=self.torch.float32).pin_memory() for n in CUR]

    def advect_host(self, frame, dt, host):
        """host: five pinned tensors (owned planes); u, v, w are uploaded, all five come back advected."""
        for n, hb in zip(CUR[:3], host[:3]):
            self.owned(n).copy_(hb, non_blocking=True)
        self.stepper.advect(frame, dt)
        for n, hb in zip(CUR, host):
            hb.copy_(self.owned(n), non_blocking=True)
        self.torch.cuda.synchronize()

    def accumulate_host(self, frame, dt, forced, final):
        """forced: u, v, w after the external forces; final: u, v, w, rho, T after projection (owned
        planes, pinned).  The change fields are formed on the device as the reference forms them on the
        host (BimocqSolver.cpp:149-162): d_ext = forced - advected, d_proj = final - forced,
        d_scalar = final - advected; the current fields become `final`."""
        from .solver3d import _dp
        lib, torch = self.lib, self.torch
        if getattr(self, "_stage", None) is None:
            self._stage = torch.empty(max(self.owned(n).numel() for n in CUR), dtype=torch.float32, device=self.owned("U").device)
        for c, n in enumerate(CUR):
            cur, ext = self.owned(n), self.owned(CHANGE[c])
            stage = self._stage[:cur.numel()].view(cur.shape)
            stage.copy_(final[c], non_blocking=True)
            if c < 3:
                proj = self.owned(CHANGE[5 + c])
                proj.copy_(forced[c], non_blocking=True)
                lib.gpu_add_field(_dp(ext), _dp(proj), _dp(cur), -1.0, cur.numel())      # forced - advected
                lib.gpu_add_field(_dp(proj), _dp(stage), _dp(proj), -1.0, cur.numel())   # final - forced
            else:
                lib.gpu_add_field(_dp(ext), _dp(stage), _dp(cur), -1.0, cur.numel())     # final - advected
            cur.copy_(stage)
        self.stepper.accumulate(frame, dt)
        torch.cuda.synchronize()

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

    def timing_read_gaps(self):
        return self.r.solver.timing_read_gaps()

    def close(self):
        if self.comm is None:                 # native driver: unmap the peers before anybody frees
            self.torch.cuda.synchronize()
            self.dist_barrier()
            self.r.close()
            return
        if hasattr(self.comm, "close"):
            self.torch.cuda.synchronize()
            self.dist_barrier()
            self.comm.close()
        self.r.close()

    def dist_barrier(self):
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
