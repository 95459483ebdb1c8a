"""GPU tests of the z-slab path: N logical ranks on ONE device (LocalComm) must reproduce the
single-GPU solver bit for bit -- owned planes of every field and map, the reinit sequence -- and,
when the box has >= 2 GPUs, the same over NCCL with one process per GPU (tests/run_zslab_nccl.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from gpufluidsimulation_b200 import scenes, zslab

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
CHECK = zslab.CUR + zslab.INIT + zslab.PREV + zslab.MAPS_BWD + zslab.MAPS_FWD + zslab.MAPS_BWDP


# the last case has 128 x 128 planes and a power-of-two cell size: the pitch-specialised kernels on slabs
@pytest.mark.parametrize("world,blend,L,dims", [(2, 1.0, 1.0, (32, 28, 48)), (3, 0.5, 0.2, (32, 28, 48)), (4, 1.0, 1.0, (32, 28, 48)),
                                                (2, 0.5, 1.0, (128, 128, 40))])
def test_logical_slabs_on_one_gpu_match_single_solver(cuda, world, blend, L, dims):
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    (ni, nj, nk), halo, frames, dt = dims, 11, 6, 0.02
    if ni == 128:
        frames, dt = 4, 0.005
    h = L / ni
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, L)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    full = [u, v, w, rho, T]
    single = BimocqAdvection3D(ni, nj, nk, h, blend)
    single.set_initial(*full)
    ranks = [zslab.CudaSlabRank(ni, nj, nk, h, blend, r, world, halo) for r in range(world)]
    for r in ranks:
        for name, a in zip(zslab.CUR, full):
            _, p0, npl, _, _ = r.solver.field_info(name)
            r.solver.upload(name, a[p0:p0 + npl])
        r.solver.reset()
    st = zslab.ZSlabStepper(ranks, zslab.LocalComm(world), blend)
    for frame in range(frames):
        single.advect(frame, dt)
        st.advect(frame, dt)
        single.apply_buoyancy(0.2, dt)
        for r in ranks:
            r.solver.apply_buoyancy(0.2, dt)
        single.accumulate(frame, dt)
        st.accumulate(frame, dt)
        sst = single.stats()
        assert (bool(sst["vel_reinit"]), bool(sst["scalar_reinit"])) == (st.stats["vel_reinit"], st.stats["scalar_reinit"])
        assert st.stats["halo_used"] <= st.halo
        for name in CHECK:
            want = single.download(name)
            dz = 1 if name in zslab.W_TYPE else 0
            for r in ranks:
                kb, ke = r.k0, r.k1 + (1 if dz and r.k1 == nk else 0)
                got, p0 = r.field_with_origin(name)
                assert np.array_equal(got[kb - p0:ke - p0].cpu().numpy(), want[kb:ke]), (frame, name, r.rank)
    for r in ranks:
        r.close()
    single.close()


def _load(ranks, full):
    for r in ranks:
        for name, a in zip(zslab.CUR, full):
            _, p0, npl, _, _ = r.solver.field_info(name)
            r.solver.upload(name, a[p0:p0 + npl])
        r.solver.reset()


def _assert_owned_equal(ranks, single, nk, names, tag):
    for name in names:
        want = single.download(name)
        dz = 1 if name in zslab.W_TYPE else 0
        for r in ranks:
            kb, ke = r.k0, r.k1 + (1 if dz and r.k1 == nk else 0)
            got, p0 = r.field_with_origin(name)
            assert np.array_equal(got[kb - p0:ke - p0].cpu().numpy(), want[kb:ke]), (tag, name, r.rank)


def test_halo_overflow_grows_the_halo(cuda):
    """A halo allocated too narrow for the first frame's reach is re-allocated (bmq3d_grow_halo) and the
    step goes on, bit-identical to a single GPU."""
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    ni, nj, nk, dt = 24, 24, 32, 0.02
    h = 1.0 / ni
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, 1.0)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 6.0)     # CFL 6 -> needs >= 9 halo planes
    single = BimocqAdvection3D(ni, nj, nk, h, 1.0)
    single.set_initial(u, v, w, rho, T)
    ranks = [zslab.CudaSlabRank(ni, nj, nk, h, 1.0, r, 2, 4) for r in range(2)]
    _load(ranks, (u, v, w, rho, T))
    st = zslab.ZSlabStepper(ranks, zslab.LocalComm(2))
    for frame in range(2):
        single.advect(frame, dt); st.advect(frame, dt)
        single.accumulate(frame, dt); st.accumulate(frame, dt)
        _assert_owned_equal(ranks, single, nk, CHECK, frame)
    assert st.grow_count >= 1 and st.halo >= 9 and all(r.halo == st.halo for r in ranks)
    for r in ranks:
        r.close()
    single.close()


@pytest.mark.parametrize("world,blend", [(8, 1.0), (3, 0.5)])
def test_forty_frames_with_growing_displacement(cuda, world, blend):
    """The failure of round 1 (SCALE_r01: HaloTooNarrow at frame ~20): 40 free-running frames at
    CFL 1.5, so that the scalar mapper's z-displacement grows for up to 30 frames.  The halo starts
    at 12 planes, narrower than what the run needs: it has to grow; with 8 ranks the slabs are 12 planes, so the
    halos reach past the neighbouring slab (planes pulled from ranks two away).  Owned
    planes stay bit-identical to a single GPU; the widths follow the per-mapper displacement."""
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    ni, nj, nk, dt, frames = 32, 28, 96, 0.02, 40
    h = 1.0 / ni
    # a z-directed jet w = sin^2(pi x) sin^2(pi y) (divergence free) on top of the plume's ring: map points
    # travel up to 1.5 planes per frame, so the scalar maps reach tens of planes before they are reinitialised
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, 1.0)
    x = (np.arange(ni) / ni)[None, None, :]; y = (np.arange(nj) / nj)[None, :, None]
    jet = (np.sin(np.pi * x) ** 2 * np.sin(np.pi * y) ** 2).astype(np.float32)
    w = 0.02 * w / max(float(np.abs(w).max()), 1e-30) + np.broadcast_to(jet, w.shape)
    u, v = 0.02 * u / float(np.abs(u).max()), 0.02 * v / float(np.abs(v).max())
    u, v, w = scenes.scale_to_cfl(u, v, np.ascontiguousarray(w, dtype=np.float32), h, dt, 1.5)
    u, v, w = [np.ascontiguousarray(a, dtype=np.float32) for a in (u, v, w)]
    single = BimocqAdvection3D(ni, nj, nk, h, blend)
    single.set_initial(u, v, w, rho, T)
    ranks = [zslab.CudaSlabRank(ni, nj, nk, h, blend, r, world, 12) for r in range(world)]
    _load(ranks, (u, v, w, rho, T))
    st = zslab.ZSlabStepper(ranks, zslab.LocalComm(world), blend)
    widest = 0
    for frame in range(frames):
        single.advect(frame, dt); st.advect(frame, dt)
        single.apply_buoyancy(0.2, dt)
        for r in ranks:
            r.solver.apply_buoyancy(0.2, dt)
        single.accumulate(frame, dt); st.accumulate(frame, dt)
        sst = single.stats()
        assert (bool(sst["vel_reinit"]), bool(sst["scalar_reinit"])) == (st.stats["vel_reinit"], st.stats["scalar_reinit"]), frame
        assert st.stats["halo_vel"] <= st.stats["halo_scalar"] + 16   # the velocity mapper is reinitialised more often
        widest = max(widest, st.stats["halo_used"])
        if frame % 8 == 7 or frame == frames - 1:
            _assert_owned_equal(ranks, single, nk, CHECK, frame)
    assert widest > 12 and st.grow_count >= 1, (widest, st.grow_count)
    if world == 8:
        assert {q for q, _, _ in zslab.halo_segments("RHO", nk, world, 3, widest)} - {2, 4}, "halo never reached past the neighbour"
    for r in ranks:
        r.close()
    single.close()


def test_native_driver_with_logical_ranks_in_threads(cuda):
    """bmq3d_mg_* (the z-slab driver inside the library) with three logical ranks on one GPU, one host thread per
    rank as a C++ host would run them: peers are mapped by raw pointer (same process), the two collectives the
    library asks for are thread barriers.  Halo starts too narrow, so BMQ_ERR_HALO -> disconnect / grow /
    reconnect is exercised.  Owned planes bit-identical to a single GPU."""
    import threading

    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    torch = cuda
    ni, nj, nk, dt, world, frames, blend = 32, 28, 45, 0.02, 3, 8, 0.5
    h = 1.0 / ni
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, 1.0)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    full = (u, v, w, rho, T)
    single = BimocqAdvection3D(ni, nj, nk, h, blend)
    single.set_initial(*full)
    want = []
    for frame in range(frames):
        single.advect(frame, dt); single.apply_buoyancy(0.2, dt); single.accumulate(frame, dt)
        want.append(({n: single.download(n) for n in CHECK}, single.stats()))
    single.close()

    bar = threading.Barrier(world)
    slots, blobs, errors, grown = [None] * world, [None] * world, [], [0] * world

    def collectives(rank):
        def allreduce_max(vals):
            slots[rank] = list(vals)
            bar.wait()
            out = [max(c) for c in zip(*slots)]
            bar.wait()
            return out

        def stream_barrier(_stream):
            torch.cuda.synchronize()
            bar.wait()

        def host_barrier():
            torch.cuda.synchronize()
            bar.wait()

        def all_gather_bytes(b):
            blobs[rank] = b
            bar.wait()
            out = list(blobs)
            bar.wait()
            return out
        return allreduce_max, stream_barrier, host_barrier, all_gather_bytes

    def run(rank):
        try:
            r = zslab.NativeSlab(ni, nj, nk, h, blend, rank, world, 4, collectives(rank))
            for name, a in zip(zslab.CUR, full):
                _, p0, npl, _, _ = r.solver.field_info(name)
                r.solver.upload(name, a[p0:p0 + npl])
            r.solver.reset()
            for frame in range(frames):
                r.advect(frame, dt)
                r.solver.apply_buoyancy(0.2, dt)
                r.accumulate(frame, dt)
                torch.cuda.synchronize()
                st = r.solver.stats()
                ref, rst = want[frame]
                assert (st["vel_reinit"], st["scalar_reinit"]) == (rst["vel_reinit"], rst["scalar_reinit"]), frame
                for name in CHECK:
                    dz = 1 if name in zslab.W_TYPE else 0
                    kb, ke = r.k0, r.k1 + (1 if dz and r.k1 == nk else 0)
                    got, p0 = r.field_with_origin(name)
                    assert np.array_equal(got[kb - p0:ke - p0].cpu().numpy(), ref[name][kb:ke]), (frame, name, rank)
                bar.wait()
            grown[rank] = r.mg_stats()["halo_grown"]
            bar.wait()       # nobody destroys buffers a peer may still be pulling from
            r.close()
        except BaseException as exc:   # noqa: BLE001
            errors.append((rank, exc))
            bar.abort()

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not errors, errors
    assert all(g >= 1 for g in grown), grown


def test_nccl_two_processes(cuda):
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29571", os.path.join(HERE, "run_zslab_nccl.py")],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "ZSLAB_NCCL_OK" in out.stdout
