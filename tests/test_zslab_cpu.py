"""CPU tests of the z-slab decomposition logic (gpufluidsimulation_b200/zslab.py): partition and
halo geometry, and -- over gloo with world_size 2 and 3 -- that the slab stepper driving
oracle-backed ranks reproduces the single-domain oracle bit for bit (fields, maps, reinit frames)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from gpufluidsimulation_b200 import scenes, zslab

HERE = os.path.dirname(os.path.abspath(__file__))


def test_slab_bounds_cover_domain():
    for nk in (16, 17, 50, 512):
        for world in (1, 2, 3, 4, 8):
            b = [zslab.slab_bounds(nk, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == nk
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [k1 - k0 for k0, k1 in b]
            assert max(sizes) - min(sizes) <= 1


def test_halo_plane_ranges_pair_up():
    """What rank r sends up/down is exactly what rank r+1 / r-1 expects to receive."""
    nk, world, width = 40, 4, 6
    for name in ("RHO", "W", "U"):
        for r in range(world):
            k0, k1 = zslab.slab_bounds(nk, world, r)
            lo, up, down, upsend = zslab.halo_planes(name, nk, k0, k1, world, r, width)
            if r + 1 < world:
                n0, n1 = zslab.slab_bounds(nk, world, r + 1)
                nlo, _, ndown, _ = zslab.halo_planes(name, nk, n0, n1, world, r + 1, width)
                assert upsend == nlo       # my top planes -> neighbour's lower halo
                assert up == ndown         # neighbour's bottom planes -> my upper halo
            else:
                assert up is None and upsend is None
            if r == 0:
                assert lo is None and down is None


def test_halo_segments_reach_past_the_neighbour():
    """A halo wider than the neighbouring slab is assembled from the ranks that own the planes; the
    segments tile the wanted ranges exactly and every plane comes from its owner."""
    nk, world = 40, 8       # 5-plane slabs
    for name in ("RHO", "W"):
        nz = nk + (name == "W")
        for r in range(world):
            a0, b0 = zslab.owned_range(name, nk, world, r)
            for width in (3, 5, 12, 40):
                got = sorted((a, b, q) for q, a, b in zslab.halo_segments(name, nk, world, r, width))
                planes = [p for a, b, _ in got for p in range(a, b)]
                want = list(range(max(0, a0 - width), a0)) if r > 0 else []
                want += list(range(b0, min(b0 + width + 1, nz))) if r < world - 1 else []
                assert planes == want, (name, r, width)
                for a, b, q in got:
                    qa, qb = zslab.owned_range(name, nk, world, q)
                    assert qa <= a < b <= qb and q != r
    # direct neighbours only when the halo fits into one slab
    assert {q for q, _, _ in zslab.halo_segments("RHO", 40, 8, 3, 4)} == {2, 4}
    assert {q for q, _, _ in zslab.halo_segments("RHO", 40, 8, 3, 12)} == {0, 1, 2, 4, 5, 6}


def test_default_halo_covers_the_scalar_reinit_cap():
    # 31 frames at CFL_frame cells per frame + this frame's reach + stencil
    assert zslab.default_halo(1.5) >= 31 * 1.5 + 1.5 + 3
    assert zslab.default_halo(0.5) >= 31 * 0.5 + 0.5 + 3


def test_halo_too_narrow_is_loud_when_a_rank_cannot_grow():
    class R:   # minimal stand-in without grow_halo
        halo, nk, h, rank, k0, k1 = 4, 32, 1 / 32, 0, 0, 16
    st = zslab.ZSlabStepper([R()], zslab.LocalComm(2))
    st.disp = [7.3, 2.0]
    with pytest.raises(zslab.HaloTooNarrow):
        st._width(11)
    assert st._width(4) == 4


def test_width_grows_the_halo_of_every_rank():
    class R:
        nk, h, rank, k0, k1 = 32, 1 / 32, 0, 0, 16
        def __init__(self):
            self.halo = 4
        def grow_halo(self, new):
            self.halo = new
    ranks = [R(), R()]
    st = zslab.ZSlabStepper(ranks, zslab.LocalComm(2))
    assert st._width(11) == 11
    assert st.halo == 11 + zslab.GROW_SLACK and all(r.halo == st.halo for r in ranks) and st.grow_count == 1
    assert st._width(12) == 12 and st.grow_count == 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ni, nj, nk, halo, frames, blend, out_q):
    sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_slab_rank import OracleSlabRank
    h = 1.0 / ni
    dt = 0.02
    full = list(scenes.smoke_plume(ni, nj, nk, 1.0))
    full[:3] = scenes.scale_to_cfl(*full[:3], h, dt, 1.5)
    full = [np.ascontiguousarray(a, dtype=np.float32) for a in full]
    r = OracleSlabRank(ni, nj, nk, h, blend, rank, world, halo)
    r.set_initial(full)
    st = zslab.ZSlabStepper([r], zslab.DistComm(world, rank, torch.device("cpu")), blend)
    log = []
    for frame in range(frames):
        st.advect(frame, dt)
        # caller stand-in (local): buoyancy on v from T, written to DV_EXT, zero other changes
        T, V, dV = r.f["T"], r.f["V"], r.f["DV_EXT"]
        dV[:, 1:-1, :] = np.float32(0.5 * dt * 0.2) * (T[:, 1:, :] + T[:, :-1, :])
        V += dV
        st.accumulate(frame, dt)
        log.append((st.stats["vel_reinit"], st.stats["scalar_reinit"], st.stats["halo_used"]))
    owned = {}
    for name in zslab.CUR + zslab.INIT + zslab.MAPS_BWD + zslab.MAPS_FWD:
        dz = 1 if name in zslab.W_TYPE else 0
        kb, ke = r.own(dz)
        p0 = r.p0[name]
        owned[name] = (kb, r.f[name][kb - p0:ke - p0].copy())
    out_q.put((rank, owned, log))
    dist.barrier()
    dist.destroy_process_group()


# last case: the allocated halo is too narrow from the first frame on, so every rank has to grow it
@pytest.mark.parametrize("world,blend,halo", [(2, 1.0, 10), (3, 0.5, 10), (2, 0.5, 3)])
def test_gloo_slab_ranks_match_single_domain_oracle(oracle, world, blend, halo):
    ni, nj, nk, frames = 16, 20, 36, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ni, nj, nk, halo, frames, blend, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-domain oracle with the same forcing
    h = 1.0 / ni
    dt = 0.02
    full = list(scenes.smoke_plume(ni, nj, nk, 1.0))
    full[:3] = scenes.scale_to_cfl(*full[:3], h, dt, 1.5)
    s = oracle.Solver(ni, nj, nk, h, blend)
    s.set_initial(*full)
    ref_log = []
    for frame in range(frames):
        s.advect(frame, dt)
        adv = [a.copy() for a in s.cur]
        dv = np.zeros_like(adv[1])
        dv[:, 1:-1, :] = np.float32(0.5 * dt * 0.2) * (adv[4][:, 1:, :] + adv[4][:, :-1, :])
        forced = [adv[0], adv[1] + dv, adv[2]]
        final = forced + [adv[3], adv[4]]
        # same change fields as the slab ranks form: d_ext = dv on v only, everything else zero
        z = lambda a: oracle.padded(a.shape)
        d_ext = [z(adv[0]), oracle.padded_copy(dv), z(adv[2])]
        d_proj = [z(adv[0]), z(adv[1]), z(adv[2])]
        for c in range(5):
            s.cur[c][...] = final[c]
        s.accumulate_changes(frame, dt, d_ext, d_proj, [z(adv[3]), z(adv[4])])
        ref_log.append((s.stats["vel_reinit"], s.stats["scalar_reinit"]))
    want = dict(zip(zslab.CUR, s.cur)); want.update(zip(zslab.INIT, s.init))
    want.update(zip(zslab.MAPS_BWD, s.vel.bwd + s.sca.bwd)); want.update(zip(zslab.MAPS_FWD, s.vel.fwd + s.sca.fwd))
    for rank, owned, log in results:
        assert [(a, b) for a, b, _ in log] == ref_log
        assert halo < 5 or all(w <= halo for _, _, w in log)
        for name, (kb, arr) in owned.items():
            assert np.array_equal(arr, want[name][kb:kb + arr.shape[0]]), (rank, name)


def _single_domain_reference(oracle, ni, nj, nk, frames, blend, cfl, dt=0.02):
    h = 1.0 / ni
    full = list(scenes.smoke_plume(ni, nj, nk, 1.0))
    full[:3] = scenes.scale_to_cfl(*full[:3], h, dt, cfl)
    s = oracle.Solver(ni, nj, nk, h, blend)
    s.set_initial(*full)
    log = []
    for frame in range(frames):
        s.advect(frame, dt)
        adv = [a.copy() for a in s.cur]
        dv = np.zeros_like(adv[1])
        dv[:, 1:-1, :] = np.float32(0.5 * dt * 0.2) * (adv[4][:, 1:, :] + adv[4][:, :-1, :])
        final = [adv[0], adv[1] + dv, adv[2], adv[3], adv[4]]
        z = lambda a: oracle.padded(a.shape)
        for c in range(5):
            s.cur[c][...] = final[c]
        s.accumulate_changes(frame, dt, [z(adv[0]), oracle.padded_copy(dv), z(adv[2])], [z(adv[0]), z(adv[1]), z(adv[2])],
                             [z(adv[3]), z(adv[4])])
        log.append((s.stats["vel_reinit"], s.stats["scalar_reinit"]))
    want = dict(zip(zslab.CUR, s.cur)); want.update(zip(zslab.INIT, s.init))
    want.update(zip(zslab.MAPS_BWD, s.vel.bwd + s.sca.bwd)); want.update(zip(zslab.MAPS_FWD, s.vel.fwd + s.sca.fwd))
    return want, log


# cfl 1.5: two DMC sub-steps, widths >= 5, so the first chi exchange is skipped from the second frame on; cfl 0.7: one
# sub-step, widths of 4 < NARROW, nothing may be skipped; halo 3: the allocation has to grow in the first frame
@pytest.mark.parametrize("world,blend,halo,cfl", [(3, 1.0, 10, 1.5), (4, 0.5, 10, 1.5), (3, 1.0, 10, 0.7), (2, 0.5, 3, 1.5)])
def test_c_driver_schedule_on_logical_oracle_ranks(oracle, world, blend, halo, cfl):
    """The exchange schedule of bmq3d_mg_* (velocity halo posted with a guessed width, chi exchange skipped while the
    halo is still valid, change fields in two exchanges) restated in tests/zslab_fast_schedule.py, on oracle-backed
    logical ranks: owned planes, maps and re-initialisation frames must match the single-domain oracle bit for bit."""
    sys.path.insert(0, HERE)
    from oracle_slab_rank import OracleSlabRank
    from zslab_fast_schedule import FastScheduleStepper
    ni, nj, nk, frames, dt = 16, 20, 36, 5, 0.02
    h = 1.0 / ni
    full = list(scenes.smoke_plume(ni, nj, nk, 1.0))
    full[:3] = scenes.scale_to_cfl(*full[:3], h, dt, cfl)
    full = [np.ascontiguousarray(a, dtype=np.float32) for a in full]
    ranks = [OracleSlabRank(ni, nj, nk, h, blend, r, world, halo) for r in range(world)]
    for r in ranks:
        r.set_initial(full)
    st = FastScheduleStepper(ranks, zslab.LocalComm(world), blend)
    log = []
    for frame in range(frames):
        st.advect(frame, dt)
        for r in ranks:
            T, V, dV = r.f["T"], r.f["V"], r.f["DV_EXT"]
            dV[:, 1:-1, :] = np.float32(0.5 * dt * 0.2) * (T[:, 1:, :] + T[:, :-1, :])
            V += dV
        st.accumulate(frame, dt)
        log.append((st.stats["vel_reinit"], st.stats["scalar_reinit"]))
    want, ref_log = _single_domain_reference(oracle, ni, nj, nk, frames, blend, cfl, dt)
    assert log == ref_log
    for r in ranks:
        for name in zslab.CUR + zslab.INIT + zslab.MAPS_BWD + zslab.MAPS_FWD:
            dz = 1 if name in zslab.W_TYPE else 0
            kb, ke = r.own(dz)
            p0 = r.p0[name]
            assert np.array_equal(r.f[name][kb - p0:ke - p0], want[name][kb:ke]), (r.rank, name)
    if cfl > 1.0 and halo >= 5:
        assert st.skipped_chi >= 2, st.skipped_chi       # the shortcut was actually taken (whenever the last widths were >= 5)
    if cfl < 1.0:
        assert st.skipped_chi == 0 or any(a or b for a, b in log), st.skipped_chi   # widths of 4: only a re-initialisation validates the halo
    if halo < 5:
        assert st.grow_count >= 1
