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
    n, halo, frames, dt = 64, 14, 6, 0.02
    ni, nj, nk = n, n - 8, n + 8
    h = 1.0 / ni
    dev = torch.device("cuda", local)
    u, v, w, rho, T = scenes.smoke_plume(ni, nj, nk, 1.0, xp=torch, device=dev)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, dt, 1.5)
    for transport in ("peer", "nccl"):
        run(transport, rank, world, ni, nj, nk, h, halo, frames, dt, u, v, w, rho, T)
    dist.barrier()
    if rank == 0:
        print("ZSLAB_NCCL_OK", world, "ranks")
    dist.destroy_process_group()


def run(transport, rank, world, ni, nj, nk, h, halo, frames, dt, u, v, w, rho, T):
    z = zslab.ZSlabAdvection3D(ni, nj, nk, h, 1.0, rank=rank, world=world, halo=halo, transport=transport)
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
    if rank == 0:
        print(f"[{transport}] bit-identical to the single-GPU run over {frames} frames", z.stats(), flush=True)
    z.close(); single.close()


if __name__ == "__main__":
    main()
