"""TESTS ONLY: the exchange schedule of the C-level z-slab driver (bmq3d_mg_advect / bmq3d_mg_accumulate,
gpufluidsimulation_b200/csrc/solver3d.cu) restated on the Python stepper's stage interface, so that its LOGIC can be
checked on CPU with oracle-backed ranks (tests/test_zslab_cpu.py).  Against ZSlabStepper's plain schedule it
  * posts the velocity halo before the max-velocity reduction, with last step's width + 2, and tops it up when the
    true width is larger;
  * skips the first chi exchange of a step for a mapper whose halo still holds >= NARROW valid planes (last step's wide
    exchange, or the identity fill of a re-initialisation);
  * moves the change fields as two exchanges and accumulates the scalars first.
A wrong skip or a width that is too small leaves stale planes in a halo, and the owned planes stop matching the
single-domain oracle."""
import math

import numpy as np

from gpufluidsimulation_b200 import zslab
from gpufluidsimulation_b200.zslab import ADV, CHANGE_S, CHANGE_V, ERR, INIT, MAPS_BWD, MAPS_BWDP, MAPS_FWD, NARROW, PREV, VEL


class FastScheduleStepper(zslab.ZSlabStepper):
    def __init__(self, ranks, comm, blend=1.0):
        super().__init__(ranks, comm, blend)
        self.chi_valid = [0, 0]
        self.wv = self.ws = 3
        self.skipped_chi = 0
        self.topped_up = 0

    def advect(self, frame, dt):
        comm, ranks = self.comm, self.ranks
        dt = float(np.float32(dt))
        w_guess = min(self.halo, max(self.wv, self.ws) + 2)
        h_vel = comm.exchange_async(ranks, [(VEL, w_guess)])
        (gmax,) = comm.allreduce_max(self._each(lambda r: (r.maxvel(),)))
        cfldt = self._each(lambda r: r.set_cfl(frame, gmax))[0]
        cfl_frame = dt * max(gmax, 1e-4) / self.h
        need = [math.ceil(self.disp[m] + cfl_frame) + 3 for m in (0, 1)]
        blend_on = [self.blend != 1.0 and self.reinit_count[m] > 0 for m in (0, 1)]
        need_b = [math.ceil(self.disp_prev[m] + self.disp[m] + cfl_frame) + 3 if blend_on[m] else 0 for m in (0, 1)]
        widest = max(need + need_b + [NARROW])
        if widest > self.halo:
            # BMQ_ERR_HALO: the host lets the posted exchange finish, grows, and calls again
            comm.wait(h_vel)
            self._width(widest)
            return self.advect(frame, dt)
        wv, ws = need
        self.wv, self.ws = wv, ws
        wmax = max(wv, ws)
        self.stats.update(max_abs_vel=gmax, cfldt=cfldt, halo_used=wmax, halo_vel=wv, halo_scalar=ws, halo_allocated=self.halo)
        both = lambda names5: [(names5[0:3], wv), (names5[3:5], ws)]
        maps = lambda names6, a, b: [(names6[0:3], a), (names6[3:6], b)]
        comm.wait(h_vel)
        if wmax > w_guess:
            comm.exchange(ranks, [(VEL, wmax)])
            self.topped_up += 1
        groups = [(MAPS_BWD[3 * m:3 * m + 3], NARROW) for m in (0, 1) if self.chi_valid[m] < NARROW]
        self.skipped_chi += 2 - len(groups)
        if groups:
            comm.exchange(ranks, groups)
        h_init = comm.exchange_async(ranks, both(INIT))
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
                comm.exchange(ranks, [(MAPS_BWD, NARROW)])
            else:
                h_bwd = comm.exchange_async(ranks, maps(MAPS_BWD, wv, ws))
                self.chi_valid = [wv, ws]
        self.stats["n_substeps"] = n
        self._each(lambda r: r.forward(dt))
        h_fwd = comm.exchange_async(ranks, maps(MAPS_FWD, wv, ws))
        comm.wait(h_bwd)
        comm.wait(h_init)
        self._each(lambda r: r.advect(0))
        h_av = comm.exchange_async(ranks, [(ADV[0:3], wv)])
        self._each(lambda r: r.advect(1))
        h_as = comm.exchange_async(ranks, [(ADV[3:5], ws)])
        comm.wait(h_fwd)
        comm.wait(h_av)
        self._each(lambda r: r.error(0))
        h_ev = comm.exchange_async(ranks, [(ERR[0:3], wv)])
        comm.wait(h_as)
        self._each(lambda r: r.error(1))
        h_es = comm.exchange_async(ranks, [(ERR[3:5], ws)])
        comm.wait(h_ev)
        self._each(lambda r: r.apply(0))
        comm.wait(h_es)
        self._each(lambda r: r.apply(1))
        for which, sl in ((0, slice(0, 3)), (1, slice(3, 5))):
            if blend_on[which]:
                wb = need_b[which]
                comm.exchange(ranks, [(PREV[sl], wb), (MAPS_BWDP[which * 3:which * 3 + 3], wb)])
                self._each(lambda r: r.blend(which))

    def accumulate(self, frame, dt):
        comm, ranks = self.comm, self.ranks
        dt = float(np.float32(dt))
        h_s = comm.exchange_async(ranks, [(CHANGE_S, self.ws)])
        h_v = comm.exchange_async(ranks, [(CHANGE_V, self.wv)])
        vd2, sd2, disp_v, disp_s = comm.allreduce_max(self._each(lambda r: r.distortion()))
        comm.wait(h_s)
        self._each(lambda r: r.accumulate(1))          # before the decision: it does not depend on it
        self.disp = [disp_v, disp_s]
        vel_reinit, sca_reinit = self._each(lambda r: r.decide(frame, dt, vd2, sd2))[0]
        comm.wait(h_v)
        self._each(lambda r: r.accumulate(0))
        if vel_reinit:
            self._each(lambda r: r.reinit(0, 0))
            self._each(lambda r: r.reinit(0, 1))
            self.reinit_count[0] += 1
            self.disp_prev[0], self.disp[0] = self.disp[0], 0.0
            self.chi_valid[0] = self.halo
        if sca_reinit:
            self._each(lambda r: r.reinit(1, 0))
            self.reinit_count[1] += 1
            self.disp_prev[1], self.disp[1] = self.disp[1], 0.0
            self.chi_valid[1] = self.halo
        self.stats.update(vel_reinit=vel_reinit, scalar_reinit=sca_reinit, max_disp_z=max(disp_v, disp_s),
                          disp_z_vel=disp_v, disp_z_scalar=disp_s, vel_d2=vd2, scalar_d2=sd2, halo_grown=self.grow_count)
