"""Launched by torchrun (one process per GPU): z-slab ranks over NCCL versus a single-GPU run of
the same problem on rank 0; owned planes must be bit-identical.  Prints ZSLAB_NCCL_OK."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpufluidsimulation_b200 import scenes, zslab  # noqa: E402
from gpufluidsimulation_b200.solver3d import BimocqAdvection3D  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, halo, frames, dt = 64, 8, 14, 0.02      # the halo has to grow on the way (bmq3d_grow_halo + IPC re-mapping)
    ni, nj, nk = n, n - 8, n + 8
    h = 1.0 / ni
    dev = torch.device("cuda", local)
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, 1.0, xp=torch, device=dev)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    # the C-level driver with its three ways of synchronising an exchange (device flags + pull kernel, device flags +
    # copy engines, the host's collectives), then the two Python transports
    for transport, signal in (("native", 1), ("native", 2), ("native", 0), ("peer", None), ("nccl", None)):
        if signal is not None:
            os.environ["BMQ_MG_SIGNAL"] = str(signal)
        run(transport, rank, world, ni, nj, nk, h, halo, frames, dt, u, v, w, rho, T, signal)
    dist.barrier()
    if rank == 0:
        print("ZSLAB_NCCL_OK", world, "ranks")
    dist.destroy_process_group()


def run(transport, rank, world, ni, nj, nk, h, halo, frames, dt, u, v, w, rho, T, signal=None):
    z = zslab.ZSlabAdvection3D(ni, nj, nk, h, 1.0, rank=rank, world=world, halo=halo, transport=transport)
    if signal is not None:
        assert z.r.mg_stats()["signalling"] == signal, z.r.mg_stats()
        transport = f"native/signal={signal}"
    z.set_initial_device(u, v, w, rho, T)
    single = BimocqAdvection3D(ni, nj, nk, h, 1.0)
    single.set_initial_device(u, v, w, rho, T)
    for frame in range(frames):
        z.advect(frame, dt); z.apply_buoyancy(0.2, dt); z.accumulate(frame, dt)
        single.advect(frame, dt); single.apply_buoyancy(0.2, dt); single.accumulate(frame, dt)
        for name in zslab.CUR + zslab.INIT + zslab.MAPS_BWD + zslab.MAPS_FWD:
            dz = 1 if name in zslab.W_TYPE else 0
            kb, ke = z.r.k0, z.r.k1 + (1 if dz and z.r.k1 == nk else 0)
            got, p0 = z.r.field_with_origin(name)
            want = single.field(name)
            if not torch.equal(got[kb - p0:ke - p0], want[kb:ke]):
                err = (got[kb - p0:ke - p0] - want[kb:ke]).abs().max().item()
                print(f"[{transport}] rank {rank} frame {frame} field {name}: MISMATCH max abs {err}", flush=True)
                sys.exit(1)
    # host-buffer step: owned planes through pinned host memory on every rank vs the single-GPU C path
    host = z.alloc_host()
    full = [torch.empty(tuple(single.field(n).shape), dtype=torch.float32).pin_memory() for n in zslab.CUR]
    for hb, fb, n in zip(host, full, zslab.CUR):
        hb.copy_(z.owned(n)); fb.copy_(single.field(n))
    torch.cuda.synchronize()
    for frame in range(frames, frames + 2):
        z.advect_host(frame, dt, host)
        single.advect_host(frame, dt, *full)
        forced_z = [h.clone().pin_memory() for h in host[:3]]; forced_s = [f.clone().pin_memory() for f in full[:3]]
        forced_z[1] *= 1.01; forced_s[1] *= 1.01
        final_z = [0.98 * f for f in forced_z] + host[3:]; final_s = [0.98 * f for f in forced_s] + full[3:]
        z.accumulate_host(frame, dt, forced_z, final_z)
        single.accumulate_host(frame, dt, forced_s, final_s)
        for hb, fz in zip(host[:3], final_z[:3]):
            hb.copy_(fz)
        for fb, fs in zip(full[:3], final_s[:3]):
            fb.copy_(fs)
        for name in zslab.CUR + zslab.INIT + zslab.CHANGE:
            _, p0, _, _, _ = single.field_info(name)
            kb, ke = z.r.k0, z.r.k1 + (1 if name in zslab.W_TYPE and z.r.k1 == nk else 0)
            if not torch.equal(z.owned(name), single.field(name)[kb:ke]):
                print(f"[{transport}] host path, rank {rank} frame {frame} field {name}: MISMATCH", flush=True)
                sys.exit(1)
    if rank == 0:
        print(f"[{transport}] bit-identical to the single-GPU run over {frames} frames (+2 through host buffers)", z.stats(), flush=True)
    assert z.stepper.grow_count >= 1, "the run was meant to exercise the halo growth"
    z.close(); single.close()


if __name__ == "__main__":
    main()
