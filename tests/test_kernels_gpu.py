"""GPU parity, kernel by kernel: every legacy extern "C" gpu_* symbol of libbimocq_b200.so against
(a) the CPU oracle and (b) the reference's own CUDA kernels (oracle/_ref/libref3d.so, compiled
unmodified from bimocq3D/GPU_kernel.cu) on identical seeded inputs.  Also pins the oracle against
the reference kernels.  Grids are deliberately not multiples of the 32x8 thread tile; one case has
a power-of-two h (exact-multiply fast path), one the reference scene's h = 0.2/ni (IEEE division).
Tolerance: relative L-inf <= 1e-5 per call (helpers.TOL_STEP)."""
import ctypes as C

import numpy as np
import pytest

from helpers import TOL_STEP, Case3D, load_reference_lib, rel_linf, run_gpu_symbol

pytestmark = pytest.mark.gpu

CASES = {"pow2": (40, 36, 44, 1.0 / 32), "general": (37, 41, 35, 0.2 / 37)}


@pytest.fixture(scope="module", params=list(CASES))
def case(request, oracle):
    ni, nj, nk, h = CASES[request.param]
    return Case3D(ni, nj, nk, h, seed=7)


@pytest.fixture(scope="module")
def ours(cuda):
    from gpufluidsimulation_b200 import load_library
    lib = load_library()
    return C.CDLL(lib._name)     # raw handle: helpers set the prototypes per call


@pytest.fixture(scope="module")
def ref(cuda):
    lib = load_reference_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libref3d.so not built")
    return lib


def _errs(got, want):
    out = []
    for g, w in zip(got, want):
        assert np.isfinite(g).all(), "non-finite values in an output"
        out.append(rel_linf(g, w))
    return max(out)


# The reference's DMC update evaluates 1 - exp(-a*s) in fp32 (GPU_kernel.cu:194-196): for small a*s
# one ulp of expf is a relative error of 6e-8/(a*s) in the back-traced displacement.  CUDA's expf
# (MUFU.EX2 based) and glibc's differ in that last ulp, so a CPU oracle cannot track the reference's
# DMC kernel to 1e-5; the GPU library can, and must (same expf, bit-identical velocity samples).
TOL_ORACLE_DMC = 5e-4


def _three_way(name, args, oracle_fn, ours, ref, outputs, tol_oracle=TOL_STEP):
    """Run on the oracle, our library and the reference library; compare the arrays at `outputs`
    (indices into the array-only list).  Parity bar: ours vs the reference kernels <= TOL_STEP."""
    arrays = [a for a in args if isinstance(a, np.ndarray)]
    from oracle import oracle3d as o3
    o_arrays = [o3.padded_copy(a) for a in arrays]
    oracle_fn(o_arrays)
    mine = run_gpu_symbol(ours, name, args)
    e_mo = _errs([mine[i] for i in outputs], [o_arrays[i] for i in outputs])
    msg = f"{name}: ours-vs-oracle {e_mo:.2e}"
    if ref is not None:
        theirs = run_gpu_symbol(ref, name, args)
        e_mr = _errs([mine[i] for i in outputs], [theirs[i] for i in outputs])
        e_or = _errs([o_arrays[i] for i in outputs], [theirs[i] for i in outputs])
        msg += f", ours-vs-reference-kernel {e_mr:.2e}, oracle-vs-reference-kernel {e_or:.2e}"
        print(msg)
        assert e_mr <= TOL_STEP, msg
        assert e_or <= tol_oracle, msg
    else:
        print(msg)
    assert e_mo <= tol_oracle, msg


def test_solve_forward(case, oracle, ours, ref):
    c = case
    args = [c.u, c.v, c.w, *c.fwd, c.h, c.ni, c.nj, c.nk, c.cfldt, c.dt]
    _three_way("gpu_solve_forward", args,
               lambda a: oracle.gpu_solve_forward(*a, c.h, c.ni, c.nj, c.nk, c.cfldt, c.dt), ours, ref, [3, 4, 5])


def test_solve_backward_dmc(case, oracle, ours, ref):
    c = case
    out = [np.zeros_like(m) for m in c.bwd]
    args = [c.u, c.v, c.w, *c.bwd, *out, c.h, c.ni, c.nj, c.nk, 0.7 * c.cfldt]
    _three_way("gpu_solve_backwardDMC", args,
               lambda a: oracle.gpu_solve_backwardDMC(*a, c.h, c.ni, c.nj, c.nk, 0.7 * c.cfldt), ours, ref, [6, 7, 8],
               tol_oracle=TOL_ORACLE_DMC)


@pytest.mark.parametrize("kind", ["u", "w", "c"])
def test_semilag(case, oracle, ours, ref, kind):
    c = case
    dx, dy, dz = oracle.DIMS[kind]
    src = c.fields[kind]
    args = [np.zeros_like(src), src, c.u, c.v, c.w, dx, dy, dz, c.h, c.ni, c.nj, c.nk, c.cfldt, -c.dt]
    _three_way("gpu_semilag", args,
               lambda a: oracle.gpu_semilag(*a, dx, dy, dz, c.h, c.ni, c.nj, c.nk, c.cfldt, -c.dt), ours, ref, [0])


@pytest.mark.parametrize("is_point", [False, True])
def test_advect_velocity(case, oracle, ours, ref, is_point):
    c = case
    f = c.fields
    args = [np.zeros_like(f["u"]), np.zeros_like(f["v"]), np.zeros_like(f["w"]), f["u"], f["v"], f["w"], *c.bwd,
            c.h, c.ni, c.nj, c.nk, is_point]

    def run(a):
        for q, kind in enumerate("uvw"):
            oracle.advect(a[q], a[3 + q], a[6], a[7], a[8], c.h, c.ni, c.nj, c.nk, kind, is_point)

    _three_way("gpu_advect_velocity", args, run, ours, ref, [0, 1, 2])


def test_advect_field(case, oracle, ours, ref):
    c = case
    f = c.fields["c"]
    args = [np.zeros_like(f), f, *c.bwd, c.h, c.ni, c.nj, c.nk, False]
    _three_way("gpu_advect_field", args,
               lambda a: oracle.advect(a[0], a[1], a[2], a[3], a[4], c.h, c.ni, c.nj, c.nk, "c"), ours, ref, [0])


def test_compensate_velocity(case, oracle, ours, ref):
    """All three outputs of the reference's contract: u (result), du (pre-correction copy), u_src (error)."""
    c = case
    f, g = c.fields, c.fields2
    args = [f["u"], f["v"], f["w"], g["u"], g["v"], g["w"], np.zeros_like(f["u"]), np.zeros_like(f["v"]),
            np.zeros_like(f["w"]), *c.fwd, *c.bwd, c.h, c.ni, c.nj, c.nk, False]

    def run(a):
        for q, kind in enumerate("uvw"):
            oracle.gpu_compensate(a[q], a[3 + q], a[6 + q], a[9:12], a[12:15], c.h, c.ni, c.nj, c.nk, kind)

    _three_way("gpu_compensate_velocity", args, run, ours, ref, list(range(9)))


def test_compensate_field(case, oracle, ours, ref):
    c = case
    f, g = c.fields["c"], c.fields2["c"]
    args = [f, g, np.zeros_like(f), *c.fwd, *c.bwd, c.h, c.ni, c.nj, c.nk, False]
    _three_way("gpu_compensate_field", args,
               lambda a: oracle.gpu_compensate(a[0], a[1], a[2], a[3:6], a[6:9], c.h, c.ni, c.nj, c.nk, "c"),
               ours, ref, [0, 1, 2])


@pytest.mark.parametrize("coeff", [1.0, 2.0])
def test_accumulate_velocity(case, oracle, ours, ref, coeff):
    c = case
    f, g = c.fields, c.fields2
    args = [f["u"], f["v"], f["w"], g["u"], g["v"], g["w"], *c.fwd, c.h, c.ni, c.nj, c.nk, False, coeff]

    def run(a):
        for q, kind in enumerate("uvw"):
            oracle.cumulate(a[q], a[3 + q], a[6:9], c.h, c.ni, c.nj, c.nk, kind, coeff)

    _three_way("gpu_accumulate_velocity", args, run, ours, ref, [3, 4, 5])


def test_accumulate_field(case, oracle, ours, ref):
    c = case
    f, g = c.fields["c"], c.fields2["c"]
    args = [f, g, *c.fwd, c.h, c.ni, c.nj, c.nk, False, 1.0]
    _three_way("gpu_accumulate_field", args,
               lambda a: oracle.cumulate(a[0], a[1], a[2:5], c.h, c.ni, c.nj, c.nk, "c", 1.0), ours, ref, [1])


def test_advect_vel_double(case, oracle, ours, ref):
    c = case
    f, g = c.fields, c.fields2
    args = [f["u"], f["v"], f["w"], g["u"], g["v"], g["w"], *c.bwd, *c.bwd_prev, c.h, c.ni, c.nj, c.nk, False, 0.5]

    def run(a):
        for q, kind in enumerate("uvw"):
            oracle.double_advect(a[q], a[3 + q], a[6:9], a[9:12], c.h, c.ni, c.nj, c.nk, kind, 0.5)

    _three_way("gpu_advect_vel_double", args, run, ours, ref, [0, 1, 2])


def test_advect_field_double(case, oracle, ours, ref):
    c = case
    f, g = c.fields["c"], c.fields2["c"]
    args = [f, g, *c.bwd, *c.bwd_prev, c.h, c.ni, c.nj, c.nk, False, 0.25]
    _three_way("gpu_advect_field_double", args,
               lambda a: oracle.double_advect(a[0], a[1], a[2:5], a[5:8], c.h, c.ni, c.nj, c.nk, "c", 0.25),
               ours, ref, [0])


def test_estimate_distortion(case, oracle, ours, ref):
    c = case
    d = np.zeros_like(c.fields["c"])
    args = [d, *c.bwd, *c.fwd, c.h, c.ni, c.nj, c.nk]
    _three_way("gpu_estimate_distortion", args,
               lambda a: oracle.estimate(a[0], a[1:4], a[4:7], c.h, c.ni, c.nj, c.nk), ours, ref, [0])


def test_add_and_add_field(case, oracle, ours, ref):
    c = case
    a, b = c.fields["c"], c.fields2["c"]
    n = a.size
    want = a - 0.5 * b
    got = run_gpu_symbol(ours, "gpu_add", [a, b, -0.5, n])
    assert rel_linf(got[0], want) <= 1e-6
    got = run_gpu_symbol(ours, "gpu_add_field", [np.zeros_like(a), a, b, -1.0, n])
    assert np.array_equal(got[0], a - b)


# ---- source terms ------------------------------------------------------
def test_emit_smoke(case, oracle, ours, ref):
    c = case
    f = c.fields
    L = c.h * c.ni
    args = [f["u"], f["v"], f["w"], f["c"], c.fields2["c"], c.h, c.ni, c.nj, c.nk, 0.45 * L, 0.4 * c.h * c.nj,
            0.5 * c.h * c.nk, 0.22 * L, 1.0, 50.0, 1.0]
    # the emitter's angle goes through acosf / cosf / hypotf: CUDA's and glibc's differ in the last ulps
    _three_way("gpu_emit_smoke", args,
               lambda a: oracle.gpu_emit_smoke(a[0], a[1], a[2], a[3], a[4], *[float(np.float32(x)) if isinstance(x, float) else x for x in args[5:]]),
               ours, ref, [0, 1, 2, 3, 4], tol_oracle=1e-6)


@pytest.mark.parametrize("alpha,beta", [(0.0, 0.5), (0.3, 0.7)])
def test_add_buoyancy(case, oracle, ours, ref, alpha, beta):
    c = case
    args = [c.fields["v"], c.fields["c"], c.fields2["c"], c.ni, c.nj, c.nk, alpha, beta, 0.02]

    def run(a):
        # the reference indexes the cell-centred inputs with the v-face index (nj+1 rows): give the oracle room
        from oracle import oracle3d as o3
        big = [o3.padded((c.nk, c.nj + 1, c.ni)) for _ in range(2)]
        for b, src in zip(big, a[1:3]):
            b.reshape(-1)[:src.size] = src.reshape(-1)
        oracle.gpu_add_buoyancy(a[0], big[0], big[1], c.ni, c.nj, c.nk, alpha, beta, 0.02)

    _three_way("gpu_add_buoyancy", args, run, ours, ref, [0])


@pytest.mark.parametrize("iters", [1, 2, 4, 7, 20])
def test_diffuse_field(case, oracle, ours, ref, iters):
    """field AND both scratch arrays (the reference's GPU solver reads the first one afterwards,
    BimocqGPUSolver.cpp:167-175); the second scratch array arrives with a non-zero ring, which the
    sweeps read on every other level and the final copy puts into `field` (GPU_kernel.cu:862-875)."""
    c = case
    f = c.fields["c"]
    stale = (0.5 * c.fields2["c"]).astype(np.float32)
    args = [f, np.zeros_like(f), stale, c.ni, c.nj, c.nk, iters, 0.37]
    _three_way("gpu_diffuse_field", args,
               lambda a: oracle.gpu_diffuse_field(a[0], a[1], a[2], c.ni, c.nj, c.nk, iters, float(np.float32(0.37))),
               ours, ref, [0, 1, 2])


def test_mad(case, oracle, ours, ref):
    c = case
    a, b = c.fields["u"], c.fields2["u"]
    args = [np.zeros_like(a), a, b, 0.75, -1.25, a.size]
    _three_way("gpu_mad", args, lambda x: oracle.gpu_mad(x[0], x[1], x[2], 0.75, -1.25), ours, ref, [0])


@pytest.mark.parametrize("stag", [(0, 0, 0, 0.0, 0.0, 0.0), (1, 0, 0, 0.5, 0.0, 0.0), (0, 0, 1, 0.0, 0.0, 0.5)], ids=["centred", "u", "w"])
def test_clamp_extrema_macCormack_matches_reference(stag):
    """gpu_clamp_extrema (GPU_kernel.cu:892-950) scatters to the back-traced cell, so two threads may
    hit one cell (a race in the reference).  With a uniform velocity every thread lands on its own
    cell and the kernel is deterministic; h = 1 makes the reference's floor(position) a cell index."""
    ref = load_reference_lib()
    if ref is None:
        pytest.skip("oracle/_ref/libref3d.so not built")
    from gpufluidsimulation_b200 import capi

    ours = capi.load_library()
    ni, nj, nk = 24, 20, 28
    dx, dy, dz, ox, oy, oz = stag
    rng = np.random.default_rng(17)
    field = rng.standard_normal((nk + dz, nj + dy, ni + dx)).astype(np.float32)
    temp = (field + 0.8 * rng.standard_normal(field.shape)).astype(np.float32)
    u = np.full((nk, nj, ni + 1), 0.5, np.float32)
    v = np.full((nk, nj + 1, ni), -0.25, np.float32)
    w = np.full((nk + 1, nj, ni), 1.75, np.float32)
    import torch

    def run(lib):
        # the kernel samples the velocity up to two cells outside the grid (no bounds checks in the
        # reference): every array sits in the middle of a zeroed allocation three times its size
        dev = []
        for a in (field, temp, u, v, w):
            big = torch.zeros(3 * a.size, dtype=torch.float32, device="cuda")
            t = big[a.size:2 * a.size].view(*a.shape)
            t.copy_(torch.from_numpy(a))
            dev.append((big, t))
        F = C.POINTER(C.c_float)
        fn = lib.gpu_clamp_extrema
        fn.restype, fn.argtypes = capi._PROTOS["gpu_clamp_extrema"]
        torch.cuda.synchronize()
        fn(*[C.cast(C.c_void_p(t.data_ptr()), F) for _, t in dev], ni + dx, nj + dy, nk + dz, dx, dy, dz, ox, oy, oz, 1.0, 1.0)
        torch.cuda.synchronize()
        return dev[1][1].cpu().numpy()

    got, want = run(ours), run(ref)
    assert np.array_equal(got, want)
    assert (got != temp).sum() > 100          # the kernel did replace values
