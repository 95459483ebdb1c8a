"""ctypes binding of include/bimocq_b200.h.  Argument order and meaning are the header's."""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_LIB = None


class BimocqLibraryError(RuntimeError):
    pass


def library_path() -> str:
    # BMQ_LIB: developer override for A/B-testing kernel variants (tools/); default is the in-tree build
    return os.environ.get("BMQ_LIB") or os.path.join(_PKG, "lib", "libbimocq_b200.so")


def header_path() -> str:
    return os.path.join(_ROOT, "include", "bimocq_b200.h")


def build_library(clean: bool = False) -> str:
    """nvcc-compile csrc/*.cu for sm_100a into lib/libbimocq_b200.so (in-tree)."""
    csrc = os.path.join(_PKG, "csrc")
    if clean:
        subprocess.check_call(["make", "-C", csrc, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", csrc, "-j4"], stdout=subprocess.DEVNULL)
    return library_path()


def declared_symbols() -> list[str]:
    """Every function name include/bimocq_b200.h declares."""
    text = open(header_path()).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:gpu|bmq3d|bmq2d|bmq)_[A-Za-z0-9_]+)\s*\(", text)))


_F = C.POINTER(C.c_float)
_I = C.c_int
_f = C.c_float
_H = C.c_void_p  # bmq3d_solver*


class Stats3D(C.Structure):
    _fields_ = [("max_v", _f), ("cfldt", _f), ("n_substeps", _I), ("vel_distortion", _f),
                ("scalar_distortion", _f), ("vel_reinit", _I), ("scalar_reinit", _I),
                ("vel_reinit_count", _I), ("scalar_reinit_count", _I), ("max_disp_z", _f),
                ("max_disp_z_vel", _f), ("max_disp_z_scalar", _f)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# field ids (keep in step with the enum in include/bimocq_b200.h; tests/test_capi_symbols.py checks)
FIELD_NAMES = [
    "U", "V", "W", "RHO", "T", "U_INIT", "V_INIT", "W_INIT", "RHO_INIT", "T_INIT",
    "U_PREV", "V_PREV", "W_PREV", "RHO_PREV", "T_PREV",
    "DU_EXT", "DV_EXT", "DW_EXT", "DRHO_EXT", "DT_EXT", "DU_PROJ", "DV_PROJ", "DW_PROJ",
    "VFWD_X", "VFWD_Y", "VFWD_Z", "VBWD_X", "VBWD_Y", "VBWD_Z", "VBWDP_X", "VBWDP_Y", "VBWDP_Z",
    "SFWD_X", "SFWD_Y", "SFWD_Z", "SBWD_X", "SBWD_Y", "SBWD_Z", "SBWDP_X", "SBWDP_Y", "SBWDP_Z",
    "U_SEMI", "V_SEMI", "W_SEMI", "RHO_SEMI", "T_SEMI",
]
FIELD = {n: i for i, n in enumerate(FIELD_NAMES)}
FIELD.update({n: 64 + i for i, n in enumerate(
    ["U_ADV", "V_ADV", "W_ADV", "RHO_ADV", "T_ADV", "U_ERR", "V_ERR", "W_ERR", "RHO_ERR", "T_ERR"])})
FIELD.update({f"TMPMAP{i}": 80 + i for i in range(6)})

N_TIMING_SLOTS = 16

FIELD2_NAMES = [
    "U", "V", "RHO", "T", "U_TEMP", "V_TEMP", "U_INIT", "V_INIT", "RHO_INIT", "T_INIT",
    "U_ORIG", "V_ORIG", "RHO_ORIG", "T_ORIG", "DU", "DV", "DRHO", "DT", "DU_PREV", "DV_PREV", "DRHO_PREV", "DT_PREV",
    "DU_EXT", "DV_EXT", "DRHO_EXT", "DT_EXT", "DU_PROJ", "DV_PROJ", "U_FORCED", "V_FORCED",
    "FWD_X", "FWD_Y", "BWD_X", "BWD_Y", "BWDP_X", "BWDP_Y", "SFWD_X", "SFWD_Y", "SBWD_X", "SBWD_Y", "SBWDP_X", "SBWDP_Y",
    "MAP_TMPX", "MAP_TMPY", "U_PRESAVE", "V_PRESAVE", "U_SAVE", "V_SAVE", "RHO_SAVE", "T_SAVE",
    "U_SEMI", "V_SEMI", "RHO_SEMI", "T_SEMI",
    "U_SCRATCH", "U_SCRATCH2", "V_SCRATCH", "V_SCRATCH2", "C_SCRATCH", "C_SCRATCH2",
]
FIELD2 = {n: i for i, n in enumerate(FIELD2_NAMES)}


class MgStats(C.Structure):
    _fields_ = [("halo_allocated", _I), ("halo_needed", _I), ("halo_vel", _I), ("halo_scalar", _I), ("halo_grown", _I),
                ("exchanges", C.c_longlong), ("bytes_exchanged", C.c_longlong), ("signalling", _I)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


ALLREDUCE_MAX_FN = C.CFUNCTYPE(_I, C.POINTER(_f), _I, C.c_void_p)
STREAM_BARRIER_FN = C.CFUNCTYPE(_I, C.c_void_p, C.c_void_p)


class Stats2D(C.Structure):
    _fields_ = [("cfl", _f), ("max_vel_pre", _f), ("n_substeps", _I), ("max_vel", _f), ("vel_condition", _f),
                ("scalar_condition", _f), ("vel_remap", _I), ("scalar_remap", _I), ("last_remesh", _I),
                ("last_scalar_remesh", _I), ("total_remesh", _I), ("total_scalar_remesh", _I)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class CoarseLevel(C.Structure):
    """bmq_coarse_level == the reference's SCoarseLevelInfo (GPU_Advection.h:13-24)."""
    _fields_ = [("ni", _I), ("nj", _I), ("nk", _I), ("number", _I), ("alpha", C.c_double), ("beta", C.c_double),
                ("b", C.c_void_p), ("x", C.c_void_p), ("r", C.c_void_p)]


_D = C.c_void_p   # device double*

_PROTOS = {
    "bmq_last_error": (C.c_char_p, []),
    "bmq_clear_error": (_I, []),
    "bmq_version": (C.c_char_p, []),
    "bmq_kernel_launch_count": (C.c_ulonglong, []),
    "bmq_set_pitch_specialisation": (_I, [_I]),
    "bmq_set_gather_variant": (_I, [_I]),
    "bmq_set_fast_division": (_I, [_I]),
    "bmq_division_is_fast": (_I, [_f, _I]),
    "bmq_set_tolerance_mode": (_I, [_I]),
    "bmq_tolerance_mode": (_I, []),
    "bmq_ipc_export": (_I, [C.c_void_p, C.c_void_p]),
    "bmq_ipc_open": (_I, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "bmq_ipc_close": (_I, [C.c_void_p]),
    "bmq_copy_async": (_I, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "gpu_solve_forward": (None, [_F] * 6 + [_f, _I, _I, _I, _f, _f]),
    "gpu_solve_backwardDMC": (None, [_F] * 9 + [_f, _I, _I, _I, _f]),
    "gpu_advect_velocity": (None, [_F] * 9 + [_f, _I, _I, _I, C.c_bool]),
    "gpu_advect_vel_double": (None, [_F] * 12 + [_f, _I, _I, _I, C.c_bool, _f]),
    "gpu_advect_field": (None, [_F] * 5 + [_f, _I, _I, _I, C.c_bool]),
    "gpu_advect_field_double": (None, [_F] * 8 + [_f, _I, _I, _I, C.c_bool, _f]),
    "gpu_accumulate_velocity": (None, [_F] * 9 + [_f, _I, _I, _I, C.c_bool, _f]),
    "gpu_accumulate_field": (None, [_F] * 5 + [_f, _I, _I, _I, C.c_bool, _f]),
    "gpu_estimate_distortion": (None, [_F] * 7 + [_f, _I, _I, _I]),
    "gpu_add": (None, [_F, _F, _f, _I]),
    "gpu_compensate_velocity": (None, [_F] * 15 + [_f, _I, _I, _I, C.c_bool]),
    "gpu_compensate_field": (None, [_F] * 9 + [_f, _I, _I, _I, C.c_bool]),
    "gpu_semilag": (None, [_F] * 5 + [_I, _I, _I, _f, _I, _I, _I, _f, _f]),
    "gpu_add_field": (None, [_F, _F, _F, _f, _I]),
    "gpu_emit_smoke": (None, [_F] * 5 + [_f, _I, _I, _I] + [_f] * 7),
    "gpu_add_buoyancy": (None, [_F] * 3 + [_I, _I, _I, _f, _f, _f]),
    "gpu_diffuse_field": (None, [_F] * 3 + [_I, _I, _I, _I, _f]),
    "gpu_mad": (None, [_F] * 3 + [_f, _f, _I]),
    "gpu_clamp_extrema": (None, [_F] * 5 + [_I] * 6 + [_f] * 5),
    "bmq_blocked_elems": (C.c_longlong, [_I, _I, _I]),
    "bmq_blocked_to_linear": (_I, [_F, _F, _I, _I, _I, C.c_void_p]),
    "bmq_linear_to_blocked": (_I, [_F, _F, _I, _I, _I, C.c_void_p]),
    "bmq3d_set_host_layout": (_I, [_H, _I]),
    "bmq_max_abs3": (_I, [_F, C.c_longlong, _F, C.c_longlong, _F, C.c_longlong, C.POINTER(_f)]),
    "gpu_multi_grid_conjugate_gradient": (None, [_F] * 3 + [_D] * 7 + [C.POINTER(CoarseLevel), _I, _I, C.c_double]),
    "gpu_conjugate_gradient": (None, [_F] * 8 + [_I, _I, _I, _I, _f]),
    "gpu_projection_jacobi": (None, [_F] * 7 + [_I, _I, _I, _I, _f, _f, _f]),
    "bmq_mgpcg_create": (_I, [_I, _I, _I, _I, C.POINTER(_H)]),
    "bmq_mgpcg_destroy": (None, [_H]),
    "bmq_mgpcg_set_stream": (_I, [_H, C.c_void_p]),
    "bmq_mgpcg_solve": (_I, [_H, _F, _F, _F, _I, C.c_double]),
    "bmq_mgpcg_buffer": (_I, [_H, _I, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]),
    "bmq_mgpcg_levels": (_I, [_H, C.POINTER(CoarseLevel), _I]),
    "bmq3d_create": (_I, [_I, _I, _I, _f, _f, C.POINTER(_H)]),
    "bmq3d_create_slab": (_I, [_I, _I, _I, _f, _f, _I, _I, _I, C.POINTER(_H)]),
    "bmq3d_destroy": (_I, [_H]),
    "bmq3d_set_stream": (_I, [_H, C.c_void_p]),
    "bmq3d_field_ptr": (_I, [_H, _I, C.POINTER(C.c_void_p), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "bmq3d_upload": (_I, [_H, _I, C.c_void_p]),
    "bmq3d_download": (_I, [_H, _I, C.c_void_p]),
    "bmq3d_reset": (_I, [_H]),
    "bmq3d_advect": (_I, [_H, _I, _f, _I]),
    "bmq3d_accumulate": (_I, [_H, _I, _f]),
    "bmq3d_get_stats": (_I, [_H, C.POINTER(Stats3D)]),
    "bmq3d_timing_enable": (_I, [_H, _I]),
    "bmq3d_timing_read": (_I, [_H, _F, C.POINTER(_I), _I]),
    "bmq3d_timing_read_gaps": (_I, [_H, _F, C.POINTER(_I), _F, _I]),
    "bmq3d_timing_slot_name": (C.c_char_p, [_I]),
    "bmq3d_advect_host": (_I, [_H, _I, _f] + [C.c_void_p] * 5),
    "bmq3d_accumulate_host": (_I, [_H, _I, _f] + [C.c_void_p] * 8),
    "bmq2d_create": (_I, [_I, _I, _f, _f, C.POINTER(_H)]),
    "bmq2d_destroy": (_I, [_H]),
    "bmq2d_reset": (_I, [_H]),
    "bmq2d_set_levelset": (_I, [_H, _I]),
    "bmq2d_set_counters": (_I, [_H, _I, _I]),
    "bmq2d_field_ptr": (_I, [_H, _I, C.POINTER(C.c_void_p), C.POINTER(_I), C.POINTER(_I)]),
    "bmq2d_upload": (_I, [_H, _I, C.c_void_p]),
    "bmq2d_download": (_I, [_H, _I, C.c_void_p]),
    "bmq2d_advect": (_I, [_H, _I, _f]),
    "bmq2d_accumulate": (_I, [_H, _I, _f]),
    "bmq2d_get_stats": (_I, [_H, C.POINTER(Stats2D)]),
    "bmq2d_deferred_counts": (_I, [_H, C.POINTER(_I)]),
    "bmq2d_deferred_round_counts": (_I, [_H, C.POINTER(_I)]),
    "bmq2d_advect_host": (_I, [_H, _I, _f] + [C.c_void_p] * 4),
    "bmq2d_accumulate_host": (_I, [_H, _I, _f] + [C.c_void_p] * 6),
    "bmq2d_kernel_launch_count": (C.c_ulonglong, [_H]),
    "bmq3d_mg_create": (_I, [_I, _I, _I, _f, _f, _I, _I, _I, C.POINTER(_H)]),
    "bmq3d_mg_destroy": (_I, [_H]),
    "bmq3d_mg_solver": (_I, [_H, C.POINTER(_H)]),
    "bmq3d_mg_set_collectives": (_I, [_H, ALLREDUCE_MAX_FN, STREAM_BARRIER_FN, C.c_void_p]),
    "bmq3d_mg_set_signalling": (_I, [_H, _I]),
    "bmq3d_mg_export_size": (_I, [_H, C.POINTER(C.c_size_t)]),
    "bmq3d_mg_export": (_I, [_H, C.c_void_p]),
    "bmq3d_mg_connect": (_I, [_H, C.c_void_p]),
    "bmq3d_mg_disconnect": (_I, [_H]),
    "bmq3d_mg_grow_halo": (_I, [_H, _I]),
    "bmq3d_mg_advect": (_I, [_H, _I, _f]),
    "bmq3d_mg_accumulate": (_I, [_H, _I, _f]),
    "bmq3d_mg_get_stats": (_I, [_H, C.POINTER(MgStats)]),
    "bmq3d_stage_maxvel": (_I, [_H, C.POINTER(_f)]),
    "bmq3d_stage_set_cfl": (_I, [_H, _I, _f]),
    "bmq3d_stage_dmc_substep": (_I, [_H, _f]),
    "bmq3d_stage_forward": (_I, [_H, _f]),
    "bmq3d_stage_semilag": (_I, [_H, _f]),
    "bmq3d_stage_advect": (_I, [_H, _I]),
    "bmq3d_stage_error": (_I, [_H, _I]),
    "bmq3d_stage_apply": (_I, [_H, _I]),
    "bmq3d_stage_blend": (_I, [_H, _I]),
    "bmq3d_stage_distortion": (_I, [_H, C.POINTER(_f), C.POINTER(_f), C.POINTER(_f)]),
    "bmq3d_stage_distortion2": (_I, [_H, C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), C.POINTER(_f)]),
    "bmq3d_grow_halo": (_I, [_H, _I]),
    "bmq3d_stage_decide": (_I, [_H, _I, _f, _f, _f]),
    "bmq3d_stage_accumulate": (_I, [_H, _I]),
    "bmq3d_stage_reinit": (_I, [_H, _I, _I]),
}


def load_library():
    """dlopen lib/libbimocq_b200.so and attach prototypes.  Raises (never falls back) if missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise BimocqLibraryError(
            f"{path} has not been built; run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or make -C gpufluidsimulation_b200/csrc).  There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)   # AttributeError here = header/library mismatch, which must be loud
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load_library().bmq_last_error().decode(errors="replace")
        raise BimocqLibraryError(f"{what or 'libbimocq_b200'} failed with status {status}: {msg}")


def check_legacy(what: str = "") -> None:
    """The legacy gpu_* symbols return void; surface a latched error as an exception."""
    lib = load_library()
    msg = lib.bmq_last_error().decode(errors="replace")
    code = lib.bmq_clear_error()
    if code != 0:
        raise BimocqLibraryError(f"{what or 'gpu_*'} latched error {code}: {msg}")
