#!/usr/bin/env python
"""bench.py -- BiMocq^2 3D advection throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--impl reference]

A "step" is one pass of the advection hot path over the whole grid: phase A (max velocity, DMC
backward + RK3 forward map update for both mappers, advect + compensate + clamp velocity, density
and temperature) and phase B (distortion estimate, reinit decision, accumulation of the change
fields, re-initialisation when triggered), i.e. bmq3d_advect + bmq3d_accumulate.  Between the two
phases a caller stand-in (torch, not counted as ours) applies the reference's buoyancy formula to
form the external change field; the projection change is zero.

Workload at N=1: BASELINE.json configs[4], the 512^3 smoke plume (fits one 180 GB B200; ~33 GB of
state).  Inputs are synthetic (closed-form vortex ring + smooth ball), resident in HBM before the
timed region; every field is 537 MB, larger than the 126 MB L2, so no L2 flush is needed between
iterations.  Timing: CUDA events on the launching stream, barrier + synchronize on both sides,
max over ranks.

Output: ONE JSON line (see the keys below).  `--impl reference` times the CPU oracle port of the
reference CUDA kernels (the 3D reference has no CPU advection path) on the host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "BiMocq2 3D advection cell-updates/sec"
UNIT = "cell-updates/s"
DT = 0.02
CFL = 1.5


def alg_bytes_per_cell(n_sub: float) -> float:
    """compulsory fp32 traffic of one cell-update = 448 + 60*n_sub bytes."""
    return 448.0 + 60.0 * n_sub


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nme, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# CPU oracle legs (test infrastructure used as the measured CPU baseline, never as the product)
# ----------------------------------------------------------------------------------------------
def cpu_oracle_rate(n: int, steps: int, warm: int = 0):
    """cell-updates/s of the CPU oracle (C restatement of the reference CUDA kernels, OpenMP over
    all host cores) on an n^3 plume with the benchmark's CFL and forcing."""
    from gpufluidsimulation_b200 import scenes
    from oracle import oracle3d as o3
    o3.build()
    h = 1.0 / n
    u, v, w, rho, T = scenes.smoke_plume(n, n, n, 1.0)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, DT, CFL)
    s = o3.Solver(n, n, n, h, 1.0)
    s.set_initial(u, v, w, rho, T)
    times = []
    for frame in range(warm + steps):
        t0 = time.perf_counter()
        s.advect(frame, DT)
        t1 = time.perf_counter()
        forced = [a for a in s.cur[:3]]
        dv = scenes.buoyancy_increment(s.cur[3], s.cur[4], 0.0, 1e-2, DT, n + 1)
        forced = [s.cur[0], (s.cur[1] + dv).astype(np.float32), s.cur[2]]
        final = forced + [s.cur[3], s.cur[4]]
        t2 = time.perf_counter()
        s.accumulate(frame, DT, forced, final)
        t3 = time.perf_counter()
        if frame >= warm:
            times.append((t1 - t0) + (t3 - t2))
    return n ** 3 * len(times) / sum(times), sum(times) / len(times)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_size
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    rate, sec = cpu_oracle_rate(n, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"BiMocq3D smoke plume {args.size}^3 (velocity + density + temperature)",
                   "timed_on": f"bounded sample: same scene at {n}^3, CFL {CFL}, dt {DT}"},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n}^3 plume, {args.steps} steps after {args.warmup} warm-up; the 3D reference "
                                   "has no CPU advection path, so this is the C restatement of its CUDA kernels (OpenMP)"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local: int) -> None:
    """Pin this rank to the CPU cores next to its GPU (NVML's ideal affinity), so that the pinned host
    buffers of the e2e leg are allocated on the GPU's own NUMA node and every rank uses its own
    PCIe root.  Best effort: any failure leaves the affinity alone."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception as exc:   # noqa: BLE001
        print(f"bench: NUMA binding skipped ({exc})", file=sys.stderr)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    from gpufluidsimulation_b200 import load_library, scenes
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D

    lib = load_library()
    n = args.size
    h = 1.0 / n
    cells = n ** 3

    if world > 1:
        from gpufluidsimulation_b200.zslab import ZSlabAdvection3D
        solver = ZSlabAdvection3D(n, n, n, h, 1.0, rank=rank, world=world, halo=args.halo, transport=args.transport,
                                  cfl_frame=CFL)
    else:
        solver = BimocqAdvection3D(n, n, n, h, 1.0)

    u, v, w, rho, T = scenes.smoke_plume(n, n, n, 1.0, xp=torch, device=dev)
    u, v, w = scenes.scale_to_cfl(u, v, w, h, DT, CFL)
    solver.set_initial_device(u, v, w, rho, T)
    del u, v, w, rho, T
    torch.cuda.empty_cache()

    beta = 1e-2

    def forcing():
        # caller stand-in between the phases: reference buoyancy (GPU_kernel.cu:804-823) on v
        solver.apply_buoyancy(beta, DT)

    def step(frame):
        solver.advect(frame, DT)
        forcing()
        solver.accumulate(frame, DT)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    frame = 0
    for _ in range(args.warmup):
        step(frame); frame += 1
    barrier()
    if getattr(solver, "stepper", None) is not None:
        solver.stepper.prof = {}
    solver.timing_enable(True)
    solver.timing_read()
    launches0 = lib.bmq_kernel_launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    nsub = []
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(frame); frame += 1
        nsub.append(solver.stats()["n_substeps"])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.bmq_kernel_launch_count() - launches0
    stage = solver.timing_read()
    solver.timing_enable(False)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = cells * args.steps / (ms * 1e-3)
    n_sub = float(np.mean(nsub))

    # ---- roofline of the dominant kernel (largest share of the step), live CUDA-event timing
    peak, peak_src = measured_peak()
    faces = {"u": (n + 1) * n * n, "c": n * n * n}
    own_frac = 1.0 / world
    # algorithmic bytes per LAUNCH (DESIGN.md "Kernels"): fp32 arrays a launch must read/write once
    alg = {
        "accumulate_velocity": ("k_cumulate_win<NF=1,NCH=2>", 7 * 4 * faces["u"] * own_frac),   # psi3 + d_ext + d_proj + init R/W
        "advect_velocity": ("k_advect_win<NF=1>", 5 * 4 * faces["u"] * own_frac),               # chi3 + init + out
        "error_velocity": ("k_error_win<NF=1>", 6 * 4 * faces["u"] * own_frac),                 # psi3 + f_adv + init + e0
        "apply_velocity": ("k_apply_clamp_win<NF=1>", 6 * 4 * faces["u"] * own_frac),           # chi3 + e0 + f_adv + out
        "advect_scalars": ("k_advect_win<NF=2>", 7 * 4 * faces["c"] * own_frac),
        "error_scalars": ("k_error_win<NF=2>", 9 * 4 * faces["c"] * own_frac),
        "apply_scalars": ("k_apply_clamp_win<NF=2>", 9 * 4 * faces["c"] * own_frac),
        "accumulate_scalars": ("k_cumulate_win<NF=2,NCH=1>", 9 * 4 * faces["c"] * own_frac),
        "dmc_backward": ("k_dmc<NMAP=2>", 15 * 4 * faces["c"] * own_frac),
        "forward": ("k_forward<NMAP=2>", 15 * 4 * faces["c"] * own_frac),
        "distortion": ("k_estimate<NMAP=2>", 12 * 4 * faces["c"] * own_frac),
    }
    launches_per_span = {"accumulate_velocity": 3, "advect_velocity": 3, "error_velocity": 3, "apply_velocity": 3}
    shares = {k: v[0] for k, v in stage.items() if v[1] > 0}
    total_stage_ms = sum(shares.values())
    top = max((k for k in shares if k in alg), key=lambda k: shares[k])
    top_launches = stage[top][1] * launches_per_span.get(top, 1)
    top_ms_per_launch = stage[top][0] / top_launches
    achieved = alg[top][1] / (top_ms_per_launch * 1e-3) / 1e9
    traffic = None
    tfile = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tfile):
        try:
            traffic = json.load(open(tfile)).get(alg[top][0])
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": alg[top][0], "stage": top, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "ms_per_launch": top_ms_per_launch, "share_of_step": stage[top][0] / total_stage_ms,
                "alg_bytes_per_launch": alg[top][1],
                "whole_step": {"alg_bytes_per_cell_update": alg_bytes_per_cell(n_sub), "n_sub": n_sub,
                               "achieved_gbs": alg_bytes_per_cell(n_sub) * value / 1e9,
                               "frac": alg_bytes_per_cell(n_sub) * value / 1e9 / peak},
                "stage_ms_per_step": {k: round(v / args.steps, 4) for k, v in shares.items()}}

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BiMocq3D smoke plume {n}^3 (velocity + density + temperature), maps + 3 velocity "
                                   "components + 2 scalars per step",
                       "grid": [n, n, n], "dt": DT, "cfl_frame": CFL, "n_sub_mean": n_sub, "blend_coeff": 1.0,
                       "parallelism": "single GPU" if world == 1 else f"z-slab x{world}, halo {solver.halo} planes allocated (grown {solver.stepper.grow_count}x), exchange={solver.transport}",
                       "l2_policy": "inputs larger than L2 (537 MB per field vs 126 MB L2), no flush needed"},
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        }

    # ---- e2e: the same metric through the C-ABI host-buffer calls (pinned host memory, copies timed)
    if world > 1 and not args.no_e2e:
        e2e = measure_e2e_slabs(solver, torch, dist, n, args, frame, world)
        if rank == 0:
            line["e2e"] = e2e
    elif world == 1 and not args.no_e2e:
        e2e = measure_e2e(solver, torch, n, args, frame)
        if line is not None:
            line["e2e"] = e2e
    elif line is not None:
        line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                       "note": "skipped (--no-e2e)"}

    if world == 1 and rank == 0 and not args.no_2d:
        try:
            line["bimocq2d"] = measure_2d(torch)
        except Exception as exc:   # secondary line: never lose the headline over it
            line["bimocq2d"] = {"error": repr(exc)}

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on a bounded sample
    if world == 1 and rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        rate, sec = cpu_oracle_rate(args.ref_size, 2, 0)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{args.ref_size}^3 plume, 2 steps ({sec:.1f} s/step): C restatement of the reference "
                                          "CUDA kernels with OpenMP (the 3D reference has no CPU advection path)"}
    if world > 1 and rank == 0 and getattr(solver, "stepper", None) is not None and solver.stepper.profile:
        print("ZSLAB_PROFILE", {k: round(v, 4) for k, v in solver.stepper.prof.items()}, file=sys.stderr)
    if line is not None:
        emit(line)
    solver.close()
    if world > 1:
        dist.destroy_process_group()


def measure_2d(torch, n=1024, steps=20, warm=5):
    """BASELINE configs[1]: BiMocq2D 1024x1024 on one B200 (device-resident bmq2d_* path), vortex-in-a-box
    flow with two scalar blobs; the caller stand-in between the phases is a device copy (zero forces)."""
    from gpufluidsimulation_b200.solver2d import BimocqAdvection2D
    L = 1.0
    h = L / n
    dt = 0.5 * h      # max |vel| = 2 -> CFL_frame ~ 1
    xn = torch.arange(n + 1, dtype=torch.float64, device="cuda") / n
    psi = 2.0 * L * (torch.sin(torch.pi * xn)[None, :] ** 2) * (torch.sin(torch.pi * xn)[:, None] ** 2) * L / torch.pi
    u = ((psi[1:, :] - psi[:-1, :]) / h).float()
    v = (-(psi[:, 1:] - psi[:, :-1]) / h).float()
    xc = (torch.arange(n, dtype=torch.float64, device="cuda") + 0.5) / n
    rho = torch.exp(-((xc[None, :] - 0.5) ** 2 + (xc[:, None] - 0.7) ** 2) / 0.12 ** 2).float()
    T = torch.exp(-((xc[None, :] - 0.4) ** 2 + (xc[:, None] - 0.3) ** 2) / 0.1 ** 2).float()
    s = BimocqAdvection2D(n, n, h, 1.0)
    for name, a in (("U", u), ("V", v), ("U_INIT", u), ("V_INIT", v), ("RHO", rho), ("RHO_INIT", rho), ("T", T), ("T_INIT", T)):
        s.field(name).copy_(a)

    def step(frame):
        s.advect(frame, dt)
        s.field("U_FORCED").copy_(s.field("U")); s.field("V_FORCED").copy_(s.field("V"))
        s.accumulate(frame, dt)

    for f in range(warm):
        step(f)
    torch.cuda.synchronize()
    l0 = s.launches()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in range(warm, warm + steps):
        step(f)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / steps
    nsub = s.stats()["n_substeps"]
    out = {"workload": f"BiMocq2D vortex-in-a-box {n}x{n}, velocity + 2 scalars", "ms_per_step": ms,
           "cell_updates_per_s": n * n / (ms * 1e-3), "n_sub": nsub, "kernel_launches_per_step": (s.launches() - l0) / steps,
           "alg_bytes_per_cell_update": 424 + 40 * nsub,
           "hbm_roofline_frac": (424 + 40 * nsub) * n * n / (ms * 1e-3) / 1e9 / measured_peak()[0],
           "note": "1 M cells x ~500 B = 0.5 GB/step: latency/launch bound, not HBM bound (BASELINE.md section 4)"}
    s.close()
    return out


def measure_e2e(solver, torch, n, args, frame):
    """bmq3d_advect_host + bmq3d_accumulate_host: every step uploads u,v,w (3 fields; phase A does
    not read the current scalars), downloads the advected fields (5), uploads the velocity after
    forces (3) and the final fields (5).  Host buffers are pinned; the timed region is the two
    calls (synchronous)."""
    names = ("U", "V", "W", "RHO", "T")
    host = [torch.empty(tuple(solver.field(nm).shape), dtype=torch.float32).pin_memory() for nm in names]
    for hbuf, nm in zip(host, names):
        hbuf.copy_(solver.field(nm))
    torch.cuda.synchronize()
    nbytes = [hb.numel() * 4 for hb in host]
    h2d = sum(nbytes[:3]) + sum(nbytes[:3]) + sum(nbytes)
    d2h = sum(nbytes)
    steps = max(2, min(args.steps, 4))
    total = 0.0
    for it in range(1 + steps):
        t0 = time.perf_counter()
        solver.advect_host(frame, DT, *host)
        # caller stand-in: no host-side forces (zero change fields); the device kernels still run
        solver.accumulate_host(frame, DT, host[:3], host)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        frame += 1
        if it > 0:
            total += t1 - t0
    return {"value": n ** 3 * steps / total, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "ms_per_step": total / steps * 1e3, "steps": steps,
            "path": "bmq3d_advect_host + bmq3d_accumulate_host (C ABI), pinned host buffers"}


def measure_e2e_slabs(solver, torch, dist, n, args, frame, world):
    """The same step through host buffers on N GPUs: every rank uploads / downloads the planes it owns
    over its own PCIe link (ZSlabAdvection3D.advect_host / accumulate_host), halos travel over NVLink
    as in the device-resident step.  Time = max over ranks between two barriers."""
    host = solver.alloc_host()
    for hb, nm in zip(host, ("U", "V", "W", "RHO", "T")):
        hb.copy_(solver.owned(nm))
    torch.cuda.synchronize()
    nbytes = [hb.numel() * 4 for hb in host]
    counts = torch.tensor([float(sum(nbytes[:3]) * 2 + sum(nbytes)), float(sum(nbytes))], device="cuda")
    dist.all_reduce(counts)
    steps = max(2, min(args.steps, 4))
    total = 0.0
    for it in range(1 + steps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        solver.advect_host(frame, DT, host)
        solver.accumulate_host(frame, DT, host[:3], host)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        frame += 1
        if it > 0:
            total += float(t.item())
    return {"value": n ** 3 * steps / total, "unit": UNIT, "h2d_bytes_per_step": int(counts[0].item()),
            "d2h_bytes_per_step": int(counts[1].item()), "ms_per_step": total / steps * 1e3, "steps": steps,
            "path": f"ZSlabAdvection3D.advect_host + accumulate_host on {world} ranks, pinned host buffers of the owned planes"}


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """Write the ONE JSON line to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries print on stdout behind our back (NCCL writes its version banner there); the driver
    # reads exactly one JSON line, so fd 1 is pointed at stderr for the whole run and only emit()
    # writes to the real stdout.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--halo", type=int, default=None,
                    help="z-slab halo planes to allocate; default: the reach of the scalar mapper's 30-frame reinit cap "
                         "(zslab.default_halo); grows on demand")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"], help="z-slab halo exchange: P2P copies or NCCL send/recv")
    ap.add_argument("--ref-size", type=int, default=64, dest="ref_size")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-2d", action="store_true", dest="no_2d")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    try:
        if args.impl == "reference":
            run_reference_arm(args)
        else:
            run_gpu(args)
    except BaseException as exc:   # noqa: BLE001 -- the driver keeps only the tail of the output: say why, in one line
        if not isinstance(exc, SystemExit) or exc.code not in (0, None):
            msg = f"BENCH_FAILED rank={os.environ.get('RANK', '0')} {type(exc).__name__}: {exc}".replace("\n", " | ")
            print(msg, file=sys.stderr, flush=True)
            os.write(_REAL_STDOUT, (msg + "\n").encode())
        raise


if __name__ == "__main__":
    main()
