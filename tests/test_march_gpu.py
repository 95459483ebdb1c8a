"""The z-marching gather kernels (csrc/march3d.cuh, the default) against the one-cell-per-thread windowed
kernels (bmq_set_gather_variant(0)): same arithmetic, so every field, map and init buffer must be
bit-identical after free-running steps -- on odd shapes, with a general and a power-of-two cell size, on
the pitch-specialised 128 x 128 planes, and on z-slab plane ranges that cut columns into short chunks.
(Parity with the REFERENCE kernels is tested in test_kernels_gpu.py / test_solver_vs_reference_gpu.py,
which run the default variant.)"""
import numpy as np
import pytest

from gpufluidsimulation_b200 import load_library, scenes, zslab

pytestmark = pytest.mark.gpu
CHECK = zslab.CUR + zslab.INIT + zslab.PREV + zslab.MAPS_BWD + zslab.MAPS_FWD + zslab.ADV + zslab.ERR


def _run(variant, dims, L, blend, frames, dt):
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    lib = load_library()
    lib.bmq_set_gather_variant(variant)
    try:
        ni, nj, nk = dims
        h = L / ni
        u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, L)
        u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
        s = BimocqAdvection3D(ni, nj, nk, h, blend)
        s.set_initial(u, v, w, rho, T)
        out = []
        for frame in range(frames):
            s.advect(frame, dt)
            s.apply_buoyancy(0.2, dt)
            s.accumulate(frame, dt)
            out.append({n: s.download(n) for n in CHECK})
        s.close()
        return out
    finally:
        lib.bmq_set_gather_variant(1)


@pytest.mark.parametrize("dims,L,blend", [((40, 36, 44), 1.25, 1.0), ((37, 41, 35), 0.2, 0.5), ((128, 128, 24), 1.0, 1.0),
                                          ((128, 128, 40), 0.2, 1.0), ((33, 9, 70), 0.5, 1.0)])
def test_march_is_bit_identical_to_windowed(cuda, dims, L, blend):
    frames, dt = 4, 0.02
    a = _run(0, dims, L, blend, frames, dt)
    # 1: clamp fused into the apply kernel; 2: the clamp as its own shared-memory tiled kernel (clamp27.cu)
    for variant in (1, 2):
        b = _run(variant, dims, L, blend, frames, dt)
        for frame in range(frames):
            for name in CHECK:
                assert np.array_equal(a[frame][name], b[frame][name]), (variant, frame, name, float(np.abs(a[frame][name] - b[frame][name]).max()))
    # the scene really moves: the comparison is not one of untouched buffers
    assert np.abs(a[-1]["U"] - a[0]["U"]).max() > 0


@pytest.mark.parametrize("variant", [1, 2])
def test_march_on_slab_ranges(cuda, variant):
    """Columns cut by slab boundaries (owned ranges of 5-7 planes): slab ranks on the marching kernels (clamp fused /
    clamp as its own tiled kernel) vs a single windowed solver."""
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    lib = load_library()
    ni, nj, nk, dt, world, halo = 32, 28, 26, 0.02, 4, 10
    h = 1.0 / ni
    full = list(scenes.smoke_plume(ni, nj, nk, 1.0))
    full[:3] = scenes.scale_to_cfl(*full[:3], h, dt, 1.5)
    lib.bmq_set_gather_variant(0)
    single = BimocqAdvection3D(ni, nj, nk, h, 1.0)
    single.set_initial(*full)
    ranks = [zslab.CudaSlabRank(ni, nj, nk, h, 1.0, r, world, halo) for r in range(world)]
    for r in ranks:
        for name, a in zip(zslab.CUR, full):
            _, p0, npl, _, _ = r.solver.field_info(name)
            r.solver.upload(name, a[p0:p0 + npl])
        r.solver.reset()
    st = zslab.ZSlabStepper(ranks, zslab.LocalComm(world))
    try:
        for frame in range(3):
            lib.bmq_set_gather_variant(0)
            single.advect(frame, dt); single.apply_buoyancy(0.2, dt); single.accumulate(frame, dt)
            lib.bmq_set_gather_variant(variant)
            st.advect(frame, dt)
            for r in ranks:
                r.solver.apply_buoyancy(0.2, dt)
            st.accumulate(frame, dt)
            for name in zslab.CUR + zslab.INIT + zslab.MAPS_BWD + zslab.MAPS_FWD:
                want = single.download(name)
                dz = 1 if name in zslab.W_TYPE else 0
                for r in ranks:
                    kb, ke = r.k0, r.k1 + (1 if dz and r.k1 == nk else 0)
                    got, p0 = r.field_with_origin(name)
                    assert np.array_equal(got[kb - p0:ke - p0].cpu().numpy(), want[kb:ke]), (frame, name, r.rank)
    finally:
        lib.bmq_set_gather_variant(1)
        for r in ranks:
            r.close()
        single.close()
