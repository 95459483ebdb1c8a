"""TESTS ONLY: a z-slab rank whose stages run the CPU oracle (oracle/bimocq3d_oracle.c) on numpy
arrays, with the same stage interface and the same storage convention as the CUDA rank
(gpufluidsimulation_b200.zslab.CudaSlabRank / csrc/solver3d.cu).  Lets the decomposition logic
(partition, halo widths, exchange order, reductions) be checked on CPU over gloo."""
from __future__ import annotations

import numpy as np
import torch

from gpufluidsimulation_b200 import zslab
from oracle import oracle3d as o3

KIND = {}
for _names, _kinds in ((zslab.CUR, "uvwcc"), (zslab.INIT, "uvwcc"), (zslab.PREV, "uvwcc"), (zslab.ADV, "uvwcc"),
                       (zslab.ERR, "uvwcc"), (zslab.CHANGE, "uvwccuvw")):
    for _n, _k in zip(_names, _kinds):
        KIND[_n] = _k
for _n in zslab.MAPS_BWD + zslab.MAPS_FWD + zslab.MAPS_BWDP:
    KIND[_n] = "c"


class OracleSlabRank:
    def __init__(self, ni, nj, nk, h, blend, rank, world, halo):
        self.ni, self.nj, self.nk, self.rank, self.world, self.halo = ni, nj, nk, rank, world, halo
        self.h = float(np.float32(h))
        self.blend_coeff = float(blend)
        self.k0, self.k1 = zslab.slab_bounds(nk, world, rank)
        self.f = {}
        self.p0 = {}
        for name, kind in KIND.items():
            self._alloc(name, kind)
        for t in range(6):
            self._alloc(f"TMP{t}", "c")
        self.reset_maps()
        self.max_v = self.cfldt = 0.0
        self.proj_coeff = 2.0
        self.vel_last = self.sca_last = 0
        self.count = [0, 0]

    # storage: planes [k0-halo, k1+halo+1) clipped, like alloc_field in solver3d.cu
    def _alloc(self, name, kind):
        nz, ny, nx = o3.shape_of(self.ni, self.nj, self.nk, kind)
        p0 = max(0, self.k0 - self.halo)
        p1 = min(nz, self.k1 + self.halo + 1)
        self.f[name] = o3.padded((p1 - p0, ny, nx))
        self.p0[name] = p0

    def grow_halo(self, new_halo):
        """Same contract as bmq3d_grow_halo: wider storage, stored planes kept, new planes zero."""
        old_f, old_p0 = self.f, self.p0
        self.f, self.p0, self.halo = {}, {}, new_halo
        for name, a in old_f.items():
            kind = KIND.get(name, "c")
            self._alloc(name, kind)
            off = old_p0[name] - self.p0[name]
            self.f[name][off:off + a.shape[0]] = a

    def field_with_origin(self, name):
        return torch.from_numpy(self.f[name]), self.p0[name]

    def v(self, name):
        a = self.f[name]
        return o3.VirtualArray(a, self.p0[name] * a.shape[1] * a.shape[2])

    def own(self, dz):
        return (self.k0, self.k1 + (1 if dz and self.k1 == self.nk else 0))

    def _identity(self, names):
        h32 = np.float32(self.h)
        for name, axis in zip(names, range(3)):
            a = self.f[name]
            p0 = self.p0[name]
            if axis == 0:
                a[...] = (np.arange(self.ni, dtype=np.float32) * h32)[None, None, :]
            elif axis == 1:
                a[...] = (np.arange(self.nj, dtype=np.float32) * h32)[None, :, None]
            else:
                a[...] = (np.arange(p0, p0 + a.shape[0], dtype=np.float32) * h32)[:, None, None]

    def reset_maps(self):
        for group in (zslab.MAPS_BWD, zslab.MAPS_FWD, zslab.MAPS_BWDP):
            self._identity(group[:3]); self._identity(group[3:])
        self._identity(["TMP0", "TMP1", "TMP2"]); self._identity(["TMP3", "TMP4", "TMP5"])

    def set_initial(self, full):
        for name, a in zip(zslab.CUR, full):
            p0 = self.p0[name]
            self.f[name][...] = a[p0:p0 + self.f[name].shape[0]]
        for c in range(5):
            self.f[zslab.INIT[c]][...] = self.f[zslab.CUR[c]]
            self.f[zslab.PREV[c]][...] = self.f[zslab.CUR[c]]

    # ---- stages
    def maxvel(self):
        m = 0.0
        for name, dz in (("U", 0), ("V", 0), ("W", 1)):
            kb, ke = self.own(dz)
            p0 = self.p0[name]
            m = max(m, float(np.abs(self.f[name][kb - p0:ke - p0]).max()))
        return m

    def set_cfl(self, frame, gmax):
        mv = max(np.float32(1e-4), np.float32(gmax))
        self.cfldt = float(np.float32(self.h) / mv)
        self.max_v = self.h if frame == 0 else float(mv)
        return self.cfldt

    def _dims(self):
        return self.h, self.ni, self.nj, self.nk

    def dmc_substep(self, substep):
        u, v, w = self.v("U"), self.v("V"), self.v("W")
        for m, names in enumerate((zslab.MAPS_BWD[:3], zslab.MAPS_BWD[3:])):
            tmp = [f"TMP{3 * m + c}" for c in range(3)]
            o3.gpu_solve_backwardDMC(u, v, w, *[self.v(n) for n in names], *[self.v(t) for t in tmp], *self._dims(),
                                     substep, krange=self.own(0))
            for n, t in zip(names, tmp):        # ping-pong like stage_dmc in solver3d.cu
                self.f[n], self.f[t] = self.f[t], self.f[n]

    def forward(self, dt):
        u, v, w = self.v("U"), self.v("V"), self.v("W")
        for names in (zslab.MAPS_FWD[:3], zslab.MAPS_FWD[3:]):
            o3.gpu_solve_forward(u, v, w, *[self.v(n) for n in names], *self._dims(), self.cfldt, dt, krange=self.own(0))

    def _comps(self, which):
        return ((0, "u"), (1, "v"), (2, "w")) if which == 0 else ((3, "c"), (4, "c"))

    def _maps(self, group, which):
        return [self.v(n) for n in group[which * 3:which * 3 + 3]]

    def advect(self, which):
        for c, kind in self._comps(which):
            o3.advect(self.v(zslab.ADV[c]), self.v(zslab.INIT[c]), *self._maps(zslab.MAPS_BWD, which), *self._dims(),
                      kind, krange=self.own(kind == "w"))

    def error(self, which):
        for c, kind in self._comps(which):
            o3.compensate_kernel(self.v(zslab.ADV[c]), self.v(zslab.INIT[c]), self.v(zslab.ERR[c]),
                                 self._maps(zslab.MAPS_FWD, which), *self._dims(), kind, krange=self.own(kind == "w"))

    def apply(self, which):
        for c, kind in self._comps(which):
            kb, ke = self.own(kind == "w")
            cur, adv = self.f[zslab.CUR[c]], self.f[zslab.ADV[c]]
            p0 = self.p0[zslab.CUR[c]]
            cur[kb - p0:ke - p0] = adv[kb - p0:ke - p0]
            o3.cumulate(self.v(zslab.ERR[c]), self.v(zslab.CUR[c]), self._maps(zslab.MAPS_BWD, which), *self._dims(),
                        kind, -0.5, krange=(kb, ke))
            o3.clamp_extrema(self.v(zslab.ADV[c]), self.v(zslab.CUR[c]), krange=(kb, ke),
                             dims=o3.shape_of(self.ni, self.nj, self.nk, kind))

    def blend(self, which):
        if self.count[which] == 0 or self.blend_coeff == 1.0:
            return
        for c, kind in self._comps(which):
            o3.double_advect(self.v(zslab.CUR[c]), self.v(zslab.PREV[c]), self._maps(zslab.MAPS_BWD, which),
                             self._maps(zslab.MAPS_BWDP, which), *self._dims(), kind, self.blend_coeff,
                             krange=self.own(kind == "w"))

    def distortion(self):
        out = []
        disps = []
        kb, ke = self.own(0)
        for which in (0, 1):
            disp = 0.0
            d = o3.padded(self.f["RHO"].shape)
            dv = o3.VirtualArray(d, self.p0["RHO"] * d.shape[1] * d.shape[2])
            bw, fw = self._maps(zslab.MAPS_BWD, which), self._maps(zslab.MAPS_FWD, which)
            o3.estimate(dv, bw, fw, *self._dims(), krange=(kb, ke))
            p0 = self.p0["RHO"]
            out.append(float(d[kb - p0:ke - p0].max()))
            zs = (np.arange(kb, ke, dtype=np.float32) * np.float32(self.h))[:, None, None]
            for group in (zslab.MAPS_BWD, zslab.MAPS_FWD):
                mz = self.f[group[which * 3 + 2]]
                q0 = self.p0[group[which * 3 + 2]]
                disp = max(disp, float(np.abs(mz[kb - q0:ke - q0] - zs).max()) / self.h)
            disps.append(disp)
        return out[0], out[1], disps[0], disps[1]

    def decide(self, frame, dt, vd2, sd2):
        dt32 = np.float32(dt)
        vd = np.float32(np.sqrt(np.float32(vd2))) / (np.float32(self.max_v) * dt32)
        sd = np.float32(np.sqrt(np.float32(sd2))) / (np.float32(self.max_v) * dt32)
        self.proj_coeff = 2.0
        vre = sre = False
        if vd > 1.0 or frame - self.vel_last > 10:
            vre, self.vel_last, self.proj_coeff = True, frame, 1.0
        if sd > 5.0 or frame - self.sca_last > 30:
            sre, self.sca_last = True, frame
        return vre, sre

    def accumulate(self, which):
        psi = self._maps(zslab.MAPS_FWD, which)
        if which == 0:
            for c, kind in self._comps(0):
                kr = self.own(kind == "w")
                o3.cumulate(self.v(zslab.CHANGE[c]), self.v(zslab.INIT[c]), psi, *self._dims(), kind, 1.0, krange=kr)
                o3.cumulate(self.v(zslab.CHANGE[5 + c]), self.v(zslab.INIT[c]), psi, *self._dims(), kind, self.proj_coeff, krange=kr)
        else:
            for c in (3, 4):
                o3.cumulate(self.v(zslab.CHANGE[c]), self.v(zslab.INIT[c]), psi, *self._dims(), "c", 1.0, krange=self.own(0))

    def reinit(self, which, phase):
        comps = self._comps(which)
        if phase == 0:
            self.count[which] += 1
            for c3 in range(3):
                a, b = zslab.MAPS_BWDP[which * 3 + c3], zslab.MAPS_BWD[which * 3 + c3]
                self.f[a], self.f[b] = self.f[b], self.f[a]
            self._identity(zslab.MAPS_BWD[which * 3:which * 3 + 3])
            self._identity(zslab.MAPS_FWD[which * 3:which * 3 + 3])
            for c, _ in comps:
                self.f[zslab.PREV[c]], self.f[zslab.INIT[c]] = self.f[zslab.INIT[c]], self.f[zslab.PREV[c]]
                self.f[zslab.INIT[c]][...] = self.f[zslab.CUR[c]]
        elif which == 0:
            psi = self._maps(zslab.MAPS_FWD, 0)
            for c, kind in comps:
                o3.cumulate(self.v(zslab.CHANGE[5 + c]), self.v(zslab.INIT[c]), psi, *self._dims(), kind, 1.0,
                            krange=self.own(kind == "w"))

    def close(self):
        pass


