"""TESTS ONLY: ctypes binding of oracle/_ref/libref2d.so = the reference's own 2D solver
(bimocq2D/BimocqSolver2D.cpp compiled unmodified) behind oracle/ref2d_wrapper.cpp."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref2d.so")

# reference member name -> our field name
MEMBERS = {
    "u": "U", "v": "V", "rho": "RHO", "temperature": "T", "u_temp": "U_TEMP", "v_temp": "V_TEMP",
    "u_init": "U_INIT", "v_init": "V_INIT", "rho_init": "RHO_INIT", "T_init": "T_INIT",
    "u_origin": "U_ORIG", "v_origin": "V_ORIG", "rho_orig": "RHO_ORIG", "T_orig": "T_ORIG",
    "du": "DU", "dv": "DV", "drho": "DRHO", "dT": "DT", "du_prev": "DU_PREV", "dv_prev": "DV_PREV",
    "drho_prev": "DRHO_PREV", "dT_prev": "DT_PREV",
    "forward_x": "FWD_X", "forward_y": "FWD_Y", "backward_x": "BWD_X", "backward_y": "BWD_Y",
    "backward_xprev": "BWDP_X", "backward_yprev": "BWDP_Y",
    "forward_scalar_x": "SFWD_X", "forward_scalar_y": "SFWD_Y", "backward_scalar_x": "SBWD_X",
    "backward_scalar_y": "SBWD_Y", "backward_scalar_xprev": "SBWDP_X", "backward_scalar_yprev": "SBWDP_Y",
}


def available():
    return os.path.exists(_PATH)


class Ref2D:
    def __init__(self, ni, nj, L, blend):
        self.lib = C.CDLL(_PATH)
        L_ = self.lib
        L_.ref2d_create.restype = C.c_void_p
        L_.ref2d_create.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float]
        L_.ref2d_field.restype = C.POINTER(C.c_float)
        L_.ref2d_field.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L_.ref2d_h.restype = C.c_float
        L_.ref2d_h.argtypes = [C.c_void_p]
        L_.ref2d_phase_a.argtypes = [C.c_void_p, C.c_float, C.c_int]
        L_.ref2d_phase_b.argtypes = [C.c_void_p, C.c_float, C.c_int] + [C.c_void_p] * 6
        L_.ref2d_get_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L_.ref2d_get_scalars.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L_.ref2d_set_levelset.argtypes = [C.c_void_p, C.c_int]
        L_.ref2d_destroy.argtypes = [C.c_void_p]
        L_.ref2d_max_vel.restype = C.c_float
        L_.ref2d_max_vel.argtypes = [C.c_void_p]
        self.p = C.c_void_p(L_.ref2d_create(ni, nj, L, blend))
        self.ni, self.nj = ni, nj
        self.h = float(L_.ref2d_h(self.p))

    def field(self, member):
        """numpy view (nj, ni) INTO the reference object's storage (writes go to the reference)."""
        a = C.c_int(); b = C.c_int()
        ptr = self.lib.ref2d_field(self.p, member.encode(), C.byref(a), C.byref(b))
        assert ptr, member
        return np.ctypeslib.as_array(ptr, shape=(b.value, a.value))

    def phase_a(self, dt, frame):
        self.lib.ref2d_phase_a(self.p, C.c_float(dt), frame)

    def phase_b(self, dt, frame, u_forced, v_forced, u_final, v_final, rho_final, T_final):
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (u_forced, v_forced, u_final, v_final, rho_final, T_final)]
        self.lib.ref2d_phase_b(self.p, C.c_float(dt), frame, *[a.ctypes.data_as(C.c_void_p) for a in arrs])

    def counters(self):
        out = (C.c_int * 6)()
        self.lib.ref2d_get_counters(self.p, out)
        return dict(zip(("last_remesh", "last_scalar_remesh", "total_remesh", "total_scalar_remesh", "vel_remap", "scalar_remap"), out))

    def scalars(self):
        out = (C.c_float * 4)()
        self.lib.ref2d_get_scalars(self.p, out)
        return dict(zip(("cfl", "vel_condition", "scalar_condition", "max_vel"), out))

    def set_levelset(self, on):
        self.lib.ref2d_set_levelset(self.p, int(on))

    def close(self):
        if self.p:
            self.lib.ref2d_destroy(self.p)
            self.p = None
