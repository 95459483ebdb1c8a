"""Host-side mirror of the 2D advection interface (bmq2d_*, include/bimocq_b200.h).

The 2D reference keeps everything in public Array2f members of BimocqSolver2D and steps with
``advance(float dt, int frame)`` (bimocq2D/BimocqSolver2D.h:149).  ``BimocqAdvection2D`` exposes
the same step split at the two non-advection calls of advanceBIMOCQ (applyBuoyancyForce :447,
projection :454): ``advect`` = lines 394-445, ``accumulate`` = lines 449-507.  Arrays are numpy
(nj, ni) views of the reference's row-major a[i + ni*j]."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import FIELD2, Stats2D, check, load_library


class BimocqAdvection2D:
    def __init__(self, ni, nj, h, blend_coeff=1.0):
        import torch
        if not torch.cuda.is_available():
            raise capi.BimocqLibraryError("gpufluidsimulation_b200 needs a CUDA device; there is no CPU fallback")
        self.lib = load_library()
        self.ni, self.nj, self.h = ni, nj, float(np.float32(h))
        self._h = C.c_void_p()
        check(self.lib.bmq2d_create(ni, nj, self.h, float(blend_coeff), C.byref(self._h)), "bmq2d_create")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.bmq2d_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shape(self, name):
        ptr = C.c_void_p(); a = C.c_int(); b = C.c_int()
        check(self.lib.bmq2d_field_ptr(self._h, FIELD2[name], C.byref(ptr), C.byref(a), C.byref(b)), "bmq2d_field_ptr")
        return (b.value, a.value)

    def field(self, name):
        """Zero-copy torch view (nj, ni) of a device field (pointers rotate on remeshing: re-fetch)."""
        import torch
        from .solver3d import _DevView
        ptr = C.c_void_p(); a = C.c_int(); b = C.c_int()
        check(self.lib.bmq2d_field_ptr(self._h, FIELD2[name], C.byref(ptr), C.byref(a), C.byref(b)), "bmq2d_field_ptr")
        return torch.as_tensor(_DevView(ptr.value, (b.value, a.value)), device="cuda")

    def upload(self, name, host):
        host = np.ascontiguousarray(host, dtype=np.float32)
        assert host.shape == self.shape(name), (name, host.shape, self.shape(name))
        check(self.lib.bmq2d_upload(self._h, FIELD2[name], host.ctypes.data_as(C.c_void_p)), "bmq2d_upload")

    def download(self, name):
        out = np.empty(self.shape(name), dtype=np.float32)
        check(self.lib.bmq2d_download(self._h, FIELD2[name], out.ctypes.data_as(C.c_void_p)), "bmq2d_download")
        return out

    def reset(self):
        check(self.lib.bmq2d_reset(self._h), "bmq2d_reset")

    def set_levelset(self, on):
        check(self.lib.bmq2d_set_levelset(self._h, int(on)), "bmq2d_set_levelset")

    def set_counters(self, lastremeshing, rho_lastremeshing):
        check(self.lib.bmq2d_set_counters(self._h, int(lastremeshing), int(rho_lastremeshing)), "bmq2d_set_counters")

    def advect(self, frame, dt):
        check(self.lib.bmq2d_advect(self._h, int(frame), float(np.float32(dt))), "bmq2d_advect")

    def accumulate(self, frame, dt):
        check(self.lib.bmq2d_accumulate(self._h, int(frame), float(np.float32(dt))), "bmq2d_accumulate")

    def accumulate_host(self, frame, dt, u_forced, v_forced, u_final, v_final, rho_final, T_final):
        arrs = [np.array(a, dtype=np.float32, order="C") for a in (u_forced, v_forced, u_final, v_final, rho_final, T_final)]
        check(self.lib.bmq2d_accumulate_host(self._h, int(frame), float(np.float32(dt)),
                                             *[a.ctypes.data_as(C.c_void_p) for a in arrs]), "bmq2d_accumulate_host")
        return arrs[2], arrs[3]     # the time-averaged u, v (BimocqSolver2D.cpp:497-506)

    def advect_host(self, frame, dt, u, v, rho, T):
        arrs = [np.array(a, dtype=np.float32, order="C") for a in (u, v, rho, T)]
        check(self.lib.bmq2d_advect_host(self._h, int(frame), float(np.float32(dt)),
                                         *[a.ctypes.data_as(C.c_void_p) for a in arrs]), "bmq2d_advect_host")
        return arrs

    def stats(self):
        st = Stats2D()
        check(self.lib.bmq2d_get_stats(self._h, C.byref(st)), "bmq2d_get_stats")
        return st.as_dict()

    def deferred_counts(self):
        """Cells of the last step whose solveODE went past its first round, per work list (diagnostic; synchronises)."""
        c = (C.c_int * 6)()
        check(self.lib.bmq2d_deferred_counts(self._h, c), "bmq2d_deferred_counts")
        return list(c)

    def deferred_round_counts(self):
        """[list][round - 1]: cells that entered round 1..6 of solveODE in the last step (diagnostic; synchronises)."""
        c = (C.c_int * 36)()
        check(self.lib.bmq2d_deferred_round_counts(self._h, c), "bmq2d_deferred_round_counts")
        return [list(c[6 * w:6 * w + 6]) for w in range(6)]

    def launches(self):
        return int(self.lib.bmq2d_kernel_launch_count(self._h))
