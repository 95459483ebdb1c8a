"""CPU-side checks of the drop-in boundary: the C-ABI library is built, loads, and exports every
symbol include/bimocq_b200.h declares (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

from gpufluidsimulation_b200 import capi


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(capi.library_path()):
        capi.build_library()
    return capi.load_library()


def test_library_exports_every_declared_symbol(lib):
    names = capi.declared_symbols()
    assert len(names) >= 40
    raw = ctypes.CDLL(capi.library_path())
    missing = [n for n in names if not hasattr(raw, n)]
    assert not missing, f"declared in include/bimocq_b200.h but not exported: {missing}"


def test_every_declared_symbol_has_a_ctypes_prototype(lib):
    names = set(capi.declared_symbols())
    assert names == set(capi._PROTOS), names ^ set(capi._PROTOS)


def test_legacy_prototypes_match_reference_header_shapes():
    """All 22 `extern "C"` functions of the reference (bimocq3D/GPU_Advection.h:26-108) are declared
    with the reference's argument counts: the 14 hot-path symbols and the 8 of SURVEY 8(f)."""
    expected = {"gpu_solve_forward": 12, "gpu_solve_backwardDMC": 14, "gpu_advect_velocity": 14,
                "gpu_advect_vel_double": 18, "gpu_advect_field": 10, "gpu_advect_field_double": 14,
                "gpu_accumulate_velocity": 15, "gpu_accumulate_field": 11, "gpu_estimate_distortion": 11,
                "gpu_add": 4, "gpu_compensate_velocity": 20, "gpu_compensate_field": 14, "gpu_semilag": 14,
                "gpu_add_field": 5,
                "gpu_emit_smoke": 16, "gpu_add_buoyancy": 9, "gpu_diffuse_field": 8, "gpu_projection_jacobi": 14,
                "gpu_clamp_extrema": 16, "gpu_mad": 6, "gpu_conjugate_gradient": 13,
                "gpu_multi_grid_conjugate_gradient": 14}
    assert len(expected) == 22
    text = re.sub(r"/\*.*?\*/", "", open(capi.header_path()).read(), flags=re.S)
    for name, nargs in expected.items():
        m = re.search(r"void\s+%s\s*\(([^)]*)\)" % name, text)
        assert m, name
        assert len(m.group(1).split(",")) == nargs, name
        assert len(capi._PROTOS[name][1]) == nargs, name


def test_field_enum_matches_header():
    text = open(capi.header_path()).read()
    body = text[text.index("BMQ_F_U = 0"):text.index("BMQ_F_COUNT")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    ids = re.findall(r"BMQ_F_([A-Z_]+)", body)
    assert ids == capi.FIELD_NAMES


def test_version_and_error_latch_work_without_a_gpu(lib):
    assert b"sm_100a" in lib.bmq_version()
    lib.bmq_clear_error()
    assert lib.bmq_last_error() == b""


def test_no_cpu_fallback_without_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    h = ctypes.c_void_p()
    st = lib.bmq3d_create(16, 16, 16, 0.0625, 1.0, ctypes.byref(h))
    assert st != 0 and not h.value
    assert b"no CUDA device" in lib.bmq_last_error() or b"CUDA" in lib.bmq_last_error()
    lib.bmq_clear_error()


def test_argument_errors_of_the_round_two_entry_points(lib):
    """Null handles and bad modes are reported through the status code and the error latch (no GPU needed)."""
    for call in (lambda: lib.bmq3d_mg_set_signalling(None, 1),
                 lambda: lib.bmq3d_timing_read_gaps(None, None, None, None, 16),
                 lambda: lib.bmq2d_deferred_counts(None, None),
                 lambda: lib.bmq3d_mg_advect(None, 0, 0.01),
                 lambda: lib.bmq3d_mg_accumulate(None, 0, 0.01)):
        lib.bmq_clear_error()
        assert call() != 0
        assert lib.bmq_last_error() != b""
    lib.bmq_clear_error()


def test_signalling_modes_match_header():
    text = open(capi.header_path()).read()
    m = re.search(r"BMQ_MG_SIGNAL_AUTO = (-?\d+), BMQ_MG_SIGNAL_HOST = (\d+), BMQ_MG_SIGNAL_DEVICE = (\d+), BMQ_MG_SIGNAL_DEVICE_CE = (\d+)", text)
    assert m and [int(x) for x in m.groups()] == [-1, 0, 1, 2]
