#!/usr/bin/env python
"""bench.py -- BiMocq^2 3D advection throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload plume512|plume256|plume128|rings256]
                    [--domain-length 1.0|0.2] [--impl reference]

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

The headline line uses the power-of-two cell size h = 1/n (exact-multiplication path); the same run also
reports, as extra keys at N=1: `general_h` (the same workload at the reference scene's h = 0.2/n,
bimocq3D/main.cpp:36-38: division path), `workloads` (the other BASELINE configs: plume 128^3, rings 256^3),
`reference_gpu` (the reference's own kernels, oracle/_ref/libref3d.so, through its call sequence on the same
B200: the kernel to beat), `bimocq2d` (1024^2 on the GPU, plus the reference's own 2D CPU code at 256^2 over
100 steps) and `cpu_baseline` (the C restatement of the 3D kernels on the host cores).

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
# BASELINE.json configs[2..4] (+ a 256^3 plume for quick runs): grid size, scene, description
WORKLOADS = {
    "plume512": (512, "plume", "BiMocq3D smoke plume 512^3 (velocity + density + temperature)"),
    "plume256": (256, "plume", "BiMocq3D smoke plume 256^3 (velocity + density + temperature)"),
    "plume128": (128, "plume", "BiMocq3D smoke plume 128^3 (velocity + density + temperature)"),
    "rings256": (256, "rings", "BiMocq3D leapfrogging vortex rings 256^3 (velocity + two ring-core scalars)"),
}


def make_scene(kind, n, L, xp, device=None):
    from gpufluidsimulation_b200 import scenes
    f = scenes.smoke_plume if kind == "plume" else scenes.leapfrog_rings
    u, v, w, rho, T = f(n, n, n, L, xp=xp, device=device) if device is not None else f(n, n, n, L)
    u, v, w = scenes.scale_to_cfl(u, v, w, L / n, DT, CFL)
    return u, v, w, rho, T


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
        "config": {"workload": WORKLOADS[args.workload][2],
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


def alg_bytes_per_launch(n, world):
    """DESIGN.md "Kernels": fp32 arrays a launch of the stage's kernel must read or write once (per rank)."""
    fu, fc, own = (n + 1) * n * n, n * n * n, 1.0 / world
    return {
        "accumulate_velocity": ("k_march<cumulate,NF=1,NCH=2>", 7 * 4 * fu * own),   # psi3 + d_ext + d_proj + init R/W
        "advect_velocity": ("k_march<advect,NF=1>", 5 * 4 * fu * own),               # chi3 + init + out
        "error_velocity": ("k_march<error,NF=1>", 6 * 4 * fu * own),                 # psi3 + f_adv + init + e0
        "apply_velocity": ("k_march<apply,NF=1>", 6 * 4 * fu * own),                 # chi3 + e0 + f_adv + out
        "advect_scalars": ("k_march<advect,NF=2>", 7 * 4 * fc * own),
        "error_scalars": ("k_march<error,NF=2>", 9 * 4 * fc * own),
        "apply_scalars": ("k_march<apply,NF=2>", 9 * 4 * fc * own),
        "accumulate_scalars": ("k_march<cumulate,NF=2,NCH=1>", 9 * 4 * fc * own),
        "dmc_backward": ("k_dmc<NMAP=2>", 15 * 4 * fc * own),
        "forward": ("k_forward<NMAP=2>", 15 * 4 * fc * own),
        "distortion": ("k_estimate<NMAP=2>", 12 * 4 * fc * own),
    }


def make_solver(n, L, world, rank, args, blend=1.0):
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    h = L / n
    if world > 1:
        from gpufluidsimulation_b200.zslab import ZSlabAdvection3D
        return ZSlabAdvection3D(n, n, n, h, blend, rank=rank, world=world, halo=args.halo, transport=args.transport, cfl_frame=CFL)
    return BimocqAdvection3D(n, n, n, h, blend)


def timed_steps(solver, torch, dist, world, dev, steps, warmup, beta=1e-2, on_timed_start=None):
    """`warmup` untimed steps, then `steps` steps between two barriers + synchronisations, timed with CUDA
    events on the launching stream; returns (ms for all steps (max over ranks), mean n_sub, stage timings,
    next frame)."""
    def step(frame):
        solver.advect(frame, DT)
        solver.apply_buoyancy(beta, DT)      # caller stand-in between the phases (reference buoyancy on v)
        solver.accumulate(frame, DT)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    frame = 0
    for _ in range(warmup):
        step(frame); frame += 1
    barrier()
    solver.timing_enable(True)
    solver.timing_read()
    if on_timed_start:
        on_timed_start()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    nsub = []
    barrier()
    e0.record()
    for _ in range(steps):
        step(frame); frame += 1
        nsub.append(solver.stats()["n_substeps"])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # (ms, spans, ms the compute stream idled before the stage: halo exchanges, reductions, host round trips)
    stage = solver.timing_read_gaps() if hasattr(solver, "timing_read_gaps") else solver.timing_read()
    solver.timing_enable(False)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, float(np.mean(nsub)), stage, frame


def short_run(torch, dist, world, rank, dev, args, kind, n, L, steps=8, warmup=3, blend=1.0):
    """One of the other workloads, briefly: ms/step, cell-updates/s, whole-step roofline fraction."""
    solver = make_solver(n, L, world, rank, args, blend)
    solver.set_initial_device(*make_scene(kind, n, L, torch, dev))
    torch.cuda.empty_cache()
    ms, n_sub, stage, _ = timed_steps(solver, torch, dist, world, dev, steps, warmup)
    st = solver.stats()
    solver.close()
    torch.cuda.empty_cache()
    rate = n ** 3 * steps / (ms * 1e-3)
    peak, _ = measured_peak()
    out = {"grid": [n, n, n], "h": f"{L}/{n}", "steps": steps, "warmup": warmup, "ms_per_step": ms / steps, "cell_updates_per_s": rate,
           "n_sub_mean": n_sub, "hbm_roofline_frac": alg_bytes_per_cell(n_sub) * rate / 1e9 / (peak * world)}
    if world > 1:
        out["halo_vel_scalar_allocated"] = [st.get("halo_vel"), st.get("halo_scalar"), st.get("halo_allocated")]
    if blend != 1.0:
        out["blend_coeff"] = blend
        out["two_level_blend_ms_per_step"] = (stage.get("blend_velocity", (0, 0))[0] + stage.get("blend_scalars", (0, 0))[0]) / steps
        out["vel_reinit_count"] = st.get("vel_reinit_count")
    return out


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

    from gpufluidsimulation_b200 import load_library

    lib = load_library()
    n, kind, desc = WORKLOADS[args.workload]
    if args.size:
        n, desc = args.size, desc.replace(str(WORKLOADS[args.workload][0]) + "^3", f"{args.size}^3")
    L = args.domain_length
    cells = n ** 3
    solver = make_solver(n, L, world, rank, args)
    solver.set_initial_device(*make_scene(kind, n, L, torch, dev))
    torch.cuda.empty_cache()

    sampler = ClockSampler(local)
    launches0 = [0]

    def on_timed_start():
        if getattr(solver, "stepper", None) is not None:
            solver.stepper.prof = {}
        launches0[0] = lib.bmq_kernel_launch_count()
        if rank == 0:
            sampler.start()

    ms, n_sub, stage, frame = timed_steps(solver, torch, dist, world, dev, args.steps, args.warmup, on_timed_start=on_timed_start)
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.bmq_kernel_launch_count() - launches0[0]
    ms_per_step = ms / args.steps
    value = cells * args.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel (largest share of the step), live CUDA-event timing
    peak, peak_src = measured_peak()
    alg = alg_bytes_per_launch(n, world)
    launches_per_span = {"accumulate_velocity": 3, "advect_velocity": 3, "error_velocity": 3, "apply_velocity": 3}
    shares = {k: v[0] for k, v in stage.items() if v[1] > 0}
    total_stage_ms = sum(shares.values())
    top = max((k for k in shares if k in alg), key=lambda k: shares[k])
    top_launches = stage[top][1] * launches_per_span.get(top, 1)
    top_ms_per_launch = stage[top][0] / top_launches
    achieved = alg[top][1] / (top_ms_per_launch * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tfile = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if os.path.exists(tfile) and n == 512:
        try:
            rec = json.load(open(tfile))
            traffic = rec.get(alg[top][0])
            traffic_src = rec.get("source", "imported from an ncu --set full capture under profiles/ (not measured in this run)")
            if traffic is not None:
                traffic = traffic / world
        except Exception:
            traffic = None
    whole_gbs = alg_bytes_per_cell(n_sub) * value / 1e9
    roofline = {"bound": "hbm", "kernel": alg[top][0], "stage": top, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "ms_per_launch": top_ms_per_launch, "share_of_step": stage[top][0] / total_stage_ms,
                "alg_bytes_per_launch": alg[top][1],
                # whole job: all ranks' bytes over the max-over-ranks time, against N GPUs' worth of HBM
                "whole_step": {"alg_bytes_per_cell_update": alg_bytes_per_cell(n_sub), "n_sub": n_sub,
                               "achieved_gbs": whole_gbs, "peak_gbs": peak * world, "frac": whole_gbs / (peak * world)},
                "stage_ms_per_step": {k: round(v / args.steps, 4) for k, v in shares.items()}}

    # every rank's own kernel time and idle time: the step runs at the pace of the slowest slab (data-dependent: gathers
    # in the plume's slabs miss L1 more often than gathers through near-identity maps in still air)
    per_rank = None
    if world > 1:
        mine = {"stage_kernels_ms_per_step": round(total_stage_ms / args.steps, 4),
                "idle_ms_per_step": round(sum(v[2] for v in stage.values() if len(v) > 2) / args.steps, 4)}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    line = None
    if rank == 0:
        par = "single GPU"
        if world > 1:
            st = solver.stats()
            par = (f"z-slab x{world}, halo {solver.halo} planes allocated (grown {solver.stepper.grow_count}x; last step exchanged "
                   f"{st.get('halo_vel')} velocity-mapper / {st.get('halo_scalar')} scalar-mapper planes), exchange={solver.transport}" + (f" ({st.get('signalling')})" if st.get("signalling") is not None else ""))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{desc}, maps + 3 velocity components + 2 scalars per step",
                       "grid": [n, n, n], "h": f"{L}/{n}" + (" (power of two: exact-multiplication path)" if L == 1.0 else
                                                          " (the reference scene's cell size: division path)"),
                       "dt": DT, "cfl_frame": CFL, "n_sub_mean": n_sub, "blend_coeff": 1.0, "parallelism": par,
                       "l2_policy": f"inputs larger than L2 ({4 * n ** 3 / 1e6:.0f} MB per field vs 126 MB L2), no flush needed"
                                    if n >= 512 else f"{4 * n ** 3 / 1e6:.0f} MB per field, ~60 fields: working set larger than the 126 MB L2"},
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
        }
        if world > 1:
            # rank 0's view of where the step's fixed cost sits: time its compute stream idled before each stage
            waits = {k: round(v[2] / args.steps, 4) for k, v in stage.items() if len(v) > 2 and v[1] > 0 and v[2] / args.steps >= 0.005}
            line["multi_gpu"] = {"stage_kernels_ms_per_step": round(total_stage_ms / args.steps, 4),
                                 "fixed_cost_ms_per_step": round(ms_per_step - total_stage_ms / args.steps, 4),
                                 "idle_before_stage_ms_per_step": dict(sorted(waits.items(), key=lambda kv: -kv[1])),
                                 "per_rank_stage_kernels_ms_per_step": [r["stage_kernels_ms_per_step"] for r in per_rank],
                                 "per_rank_idle_ms_per_step": [r["idle_ms_per_step"] for r in per_rank],
                                 "note": "idle time before a stage = the halo exchange / reduction / host round trip it waited for "
                                         "(distortion also waits for the caller's buoyancy kernel)"}

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
    if world > 1 and rank == 0 and getattr(solver, "stepper", None) is not None and solver.stepper.profile:
        print("ZSLAB_PROFILE", {k: round(v, 4) for k, v in solver.stepper.prof.items()}, file=sys.stderr)
    solver.close()
    del solver
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs and the comparison legs (bounded: a few seconds each)
    if not args.no_extras:
        extras = {}

        def leg(name, fn):
            try:
                extras[name] = fn()
            except Exception as exc:   # noqa: BLE001 -- secondary lines: never lose the headline over them
                extras[name] = {"error": repr(exc)}

        if world == 1:
            leg("general_h", lambda: short_run(torch, dist, world, rank, dev, args, kind, n, 0.2 if L == 1.0 else 1.0, steps=min(args.steps, 6)))

            def tolerance_leg():
                # what bit-exactness costs at the reference scene's cell size: the same run with bmq_set_tolerance_mode(1)
                # (NOT a drop-in mode: see include/bimocq_b200.h; results deviate up to 1e-3 over a run)
                lib.bmq_set_tolerance_mode(1)
                try:
                    out = short_run(torch, dist, world, rank, dev, args, kind, n, 0.2, steps=min(args.steps, 6))
                finally:
                    lib.bmq_set_tolerance_mode(0)
                out["note"] = "opt-in tolerance mode: every cell size on the power-of-two kernels; not bit-exact"
                return out
            leg("general_h_tolerance_mode", tolerance_leg)
            others = {}
            for wname in ("plume128", "rings256"):
                if wname != args.workload:
                    wn, wkind, _ = WORKLOADS[wname]
                    others[wname] = short_run(torch, dist, world, rank, dev, args, wkind, wn, 1.0)
            extras["workloads"] = others
            # the two-level map blend (doubleAdvect_kernel, Mapping.cpp:196-201): off in every shipped scene (blend_coeff = 1),
            # timed here at 0.5 over enough steps for the velocity mapper to have been reinitialised
            leg("blend_0.5", lambda: short_run(torch, dist, world, rank, dev, args, "plume", 256, 1.0, steps=14, warmup=12, blend=0.5))
            leg("reference_gpu", lambda: measure_reference_gpu(torch))
            if not args.no_2d:
                leg("bimocq2d", lambda: measure_2d(torch))
        else:
            # BASELINE configs[3]: the rings at 256^3, z-slab sharded over the same ranks (all ranks take part)
            wn, wkind, _ = WORKLOADS["rings256"]
            try:
                r = short_run(torch, dist, world, rank, dev, args, wkind, wn, 1.0)
            except Exception as exc:   # noqa: BLE001
                r = {"error": repr(exc)}
            extras["workloads"] = {"rings256": r}
        if line is not None:
            line.update(extras)

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on bounded samples
    if world == 1 and rank == 0 and not args.no_cpu:
        cores = os.cpu_count() or 1
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        rate, sec = cpu_oracle_rate(args.ref_size, 1, 0)
        rate64, sec64 = cpu_oracle_rate(64, 2, 0)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{args.ref_size}^3 plume, 1 step ({sec:.1f} s): C restatement of the reference CUDA kernels "
                                          "with OpenMP (the 3D reference has no CPU advection path)",
                                "per_cell_scaling_check": {"64^3": {"value": rate64, "s_per_step": sec64},
                                                           f"{args.ref_size}^3": {"value": rate, "s_per_step": sec}}}
    if line is not None:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_reference_gpu(torch, n=256, steps=2):
    """The kernel to beat: the reference's own CUDA kernels (oracle/_ref/libref3d.so =
    GPU_kernel.cu compiled unmodified for sm_100a) driven through the reference's call sequence
    (tests/helpers.DeviceStepper = BimocqSolver::advanceBimocq with MapperBase buffer semantics, device
    resident) on the same B200 and the same 256^3 plume, next to this library's handle API."""
    import ctypes as C
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    path = os.path.join(ROOT, "oracle", "_ref", "libref3d.so")
    if not os.path.exists(path):
        return {"unavailable": "oracle/_ref/libref3d.so not built"}
    from helpers import DeviceStepper
    from gpufluidsimulation_b200.solver3d import BimocqAdvection3D
    L, h = 1.0, 1.0 / n
    u, v, w, rho, T = [a.cpu().numpy() for a in make_scene("plume", n, L, torch, torch.device("cuda"))]
    ref = DeviceStepper(n, n, n, h, 1.0, lib=C.CDLL(path))
    ref.set_initial(u, v, w, rho, T)
    ours = BimocqAdvection3D(n, n, n, h, 1.0)
    ours.set_initial(u, v, w, rho, T)

    def ref_step(frame):
        ref.advect(frame, DT)
        cur = [t.cpu().numpy() for t in ref.cur]
        return cur

    def timed(fn):
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        return a.elapsed_time(b)

    t_ref_a = t_ref_b = t_our = 0.0
    for frame in range(1 + steps):
        ta = timed(lambda: ref.advect(frame, DT))
        cur = [t.cpu().numpy() for t in ref.cur]
        tb = timed(lambda: ref.accumulate(frame, DT, cur[:3], cur))
        to = timed(lambda: (ours.advect(frame, DT), ours.accumulate(frame, DT)))
        if frame > 0:
            t_ref_a += ta; t_ref_b += tb; t_our += to
    ours.close()
    del ref
    torch.cuda.empty_cache()
    ref_ms = (t_ref_a + t_ref_b) / steps
    return {"workload": f"smoke plume {n}^3, h = 1/{n}, {steps} steps after 1 warm-up, device resident",
            "reference_kernels_ms_per_step": ref_ms, "reference_phase_a_ms": t_ref_a / steps, "reference_phase_b_ms": t_ref_b / steps,
            "ours_ms_per_step": t_our / steps, "speedup": ref_ms / (t_our / steps),
            "note": "reference side = GPU_kernel.cu unmodified (sm_100a) through MapperBaseGPU's call sequence incl. its "
                    "device-to-device copies and the host-side max reductions it does; phase B's upload of the change fields "
                    "from host arrays is inside its time (the reference forms them on the host, BimocqSolver.cpp:149-162)"}


def measure_2d(torch, n=1024, steps=20, warm=5, cpu_leg=True):
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
           "solveode_deferred_cells_last_step": s.deferred_counts(),
           "solveode_cells_entering_round_1_to_6": s.deferred_round_counts(),
           "cell_updates_per_s": n * n / (ms * 1e-3), "n_sub": nsub, "kernel_launches_per_step": (s.launches() - l0) / steps,
           "alg_bytes_per_cell_update": 424 + 40 * nsub,
           "hbm_roofline_frac": (424 + 40 * nsub) * n * n / (ms * 1e-3) / 1e9 / measured_peak()[0],
           "note": "1 M cells x ~500 B = 0.5 GB/step: latency/launch bound, not HBM bound (BASELINE.md section 4)"}
    s.close()
    if cpu_leg:
        try:
            out["cpu_reference_2d"] = measure_2d_cpu_reference()
        except Exception as exc:   # noqa: BLE001
            out["cpu_reference_2d"] = {"error": repr(exc)}
    return out


def measure_2d_cpu_reference(n=256, steps=100):
    """BASELINE configs[0]: the reference's own 2D code (bimocq2D/BimocqSolver2D.cpp compiled unmodified into
    oracle/_ref/libref2d.so) on the host cores: hot-path stages of advanceBIMOCQ (everything except forces and
    projection) over 100 steps at 256x256; tbb::parallel_for is a static-chunk std::thread shim (oracle/shim)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ref2d
    if not ref2d.available():
        return {"unavailable": "oracle/_ref/libref2d.so not built"}
    cores = os.cpu_count() or 1
    os.environ.setdefault("BMQ_SHIM_THREADS", str(cores))
    L = 1.0
    ref = ref2d.Ref2D(n, n, L, 1.0)
    h = ref.h
    dt = 0.5 * h
    xn = np.arange(n + 1, dtype=np.float64) / n
    psi = 2.0 * L * (np.sin(np.pi * xn)[None, :] ** 2) * (np.sin(np.pi * xn)[:, None] ** 2) * L / np.pi
    u = ((psi[1:, :] - psi[:-1, :]) / h).astype(np.float32)
    v = (-(psi[:, 1:] - psi[:, :-1]) / h).astype(np.float32)
    xc = (np.arange(n, dtype=np.float64) + 0.5) / n
    rho = np.exp(-((xc[None, :] - 0.5) ** 2 + (xc[:, None] - 0.7) ** 2) / 0.12 ** 2).astype(np.float32)
    T = np.exp(-((xc[None, :] - 0.4) ** 2 + (xc[:, None] - 0.3) ** 2) / 0.1 ** 2).astype(np.float32)
    for mem, a in (("u", u), ("v", v), ("u_init", u), ("v_init", v), ("rho", rho), ("temperature", T), ("rho_init", rho), ("T_init", T)):
        ref.field(mem)[...] = a
    t0 = time.perf_counter()
    for frame in range(steps):
        ref.phase_a(dt, frame)
        adv = [np.array(ref.field(m)) for m in ("u", "v", "rho", "temperature")]
        ref.phase_b(dt, frame, adv[0], adv[1], adv[0], adv[1], adv[2], adv[3])
    sec = time.perf_counter() - t0
    ref.close()
    return {"workload": f"BiMocq2D vortex-in-a-box {n}x{n}, {steps} steps, the reference's own C++ (hot-path stages)",
            "cores": int(os.environ["BMQ_SHIM_THREADS"]), "kind": "reference", "s_total": sec, "ms_per_step": sec / steps * 1e3,
            "cell_updates_per_s": n * n * steps / sec}


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
    as in the device-resident step.  Time = max over ranks between two barriers.  Also times the SAME
    transfers alone (all ranks at once, no kernels): the host-side bandwidth wall the e2e number sits on."""
    host = solver.alloc_host()
    for hb, nm in zip(host, ("U", "V", "W", "RHO", "T")):
        hb.copy_(solver.owned(nm))
    torch.cuda.synchronize()
    nbytes = [hb.numel() * 4 for hb in host]
    counts = torch.tensor([float(sum(nbytes[:3]) * 2 + sum(nbytes)), float(sum(nbytes))], device="cuda")
    dist.all_reduce(counts)
    steps = max(2, min(args.steps, 4))

    def timed(fn):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    total = 0.0
    for it in range(1 + steps):
        def one():
            solver.advect_host(frame, DT, host)
            solver.accumulate_host(frame, DT, host[:3], host)
        t = timed(one)
        frame += 1
        if it > 0:
            total += t
    # the transfers of one step alone: 3 + 3 + 5 uploads and 5 downloads of the owned planes, all ranks at once
    dev = [torch.empty_like(hb, device="cuda") for hb in host]

    def copies():
        for q in (0, 1, 2, 0, 1, 2, 0, 1, 2, 3, 4):
            dev[q].copy_(host[q], non_blocking=True)
        for q in range(5):
            host[q].copy_(dev[q], non_blocking=True)
    timed(copies)
    copy_s = min(timed(copies) for _ in range(2))
    moved = float(counts[0].item() + counts[1].item())
    return {"value": n ** 3 * steps / total, "unit": UNIT, "h2d_bytes_per_step": int(counts[0].item()),
            "d2h_bytes_per_step": int(counts[1].item()), "ms_per_step": total / steps * 1e3, "steps": steps,
            "transfers_alone_ms_per_step": copy_s * 1e3, "transfers_alone_aggregate_gbs": moved / copy_s / 1e9,
            "path": f"ZSlabAdvection3D.advect_host + accumulate_host on {world} ranks, pinned host buffers of the owned planes; "
                    "transfers_alone = the same H2D/D2H copies of one step issued by all ranks at once with no kernels "
                    "(the host-side bandwidth the e2e step cannot beat)"}


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
    ap.add_argument("--workload", default="plume512", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=0, help="override the workload's grid size (n^3)")
    ap.add_argument("--domain-length", type=float, default=1.0, dest="domain_length",
                    help="L: cell size h = L/n.  1.0 = power-of-two h (headline), 0.2 = the reference scene's cell size")
    ap.add_argument("--no-extras", action="store_true", dest="no_extras",
                    help="skip general_h / workloads / reference_gpu / bimocq2d (headline, e2e and cpu_baseline only)")
    ap.add_argument("--halo", type=int, default=None,
                    help="z-slab halo planes to allocate; default: the reach of the scalar mapper's 30-frame reinit cap "
                         "(zslab.default_halo); grows on demand")
    ap.add_argument("--transport", default="native", choices=["native", "peer", "nccl"],
                    help="z-slab driver: the library's own (bmq3d_mg_*, peer copies), the Python stepper with peer copies, or with NCCL send/recv")
    ap.add_argument("--ref-size", type=int, default=128, dest="ref_size", help="grid of the CPU oracle sample (n^3)")
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
