"""Generates tests/golden/ref2d_steps.npz by EXECUTING THE REFERENCE'S OWN 2D CODE
(oracle/_ref/libref2d.so = bimocq2D/BimocqSolver2D.cpp compiled unmodified; CPU, so it runs in
the build container):  python tests/golden/make_golden_2d.py
Per step the file holds the complete advection state before the step, the fields after phase A
(advanceBIMOCQ lines 394-445), the forcing, and the state after phase B (lines 449-507)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ref2d  # noqa: E402

NI, NJ, L, BLEND, DT, FRAMES = 40, 32, 0.2, 0.5, 0.004, 10
AFTER_A = ("u", "v", "rho", "temperature", "forward_x", "backward_x", "backward_scalar_y")
AFTER_B = ("u", "v", "u_temp", "du", "dv", "drho", "dT", "u_init", "rho_init", "u_origin", "du_prev", "backward_xprev",
           "forward_x", "backward_scalar_x")


def initial(ni, nj, L):
    h = L / ni
    xn = np.arange(ni + 1) / ni
    yn = np.arange(nj + 1) / nj
    psi = 2.0 * L * (np.sin(np.pi * xn)[None, :] ** 2) * (np.sin(np.pi * yn)[:, None] ** 2) * L / np.pi
    u = ((psi[1:, :] - psi[:-1, :]) / h).astype(np.float32)
    v = (-(psi[:, 1:] - psi[:, :-1]) / h).astype(np.float32)
    x = (np.arange(ni) + 0.5) / ni
    y = (np.arange(nj) + 0.5) / nj
    rho = np.exp(-((x[None, :] - 0.5) ** 2 + (y[:, None] - 0.7) ** 2) / 0.12 ** 2).astype(np.float32)
    T = np.exp(-((x[None, :] - 0.4) ** 2 + (y[:, None] - 0.3) ** 2) / 0.1 ** 2).astype(np.float32)
    return u, v, rho, T


def forcing(adv, dt):
    v_forced = adv[1].copy()
    v_forced[1:-1, :] += np.float32(0.5 * dt) * (adv[3][1:, :] + adv[3][:-1, :])
    return adv[0], v_forced, (0.99 * adv[0]).astype(np.float32), (0.99 * v_forced).astype(np.float32), adv[2], adv[3]


def main():
    ref = ref2d.Ref2D(NI, NJ, L, BLEND)
    u, v, rho, T = initial(NI, NJ, L)
    for m, a in (("u", u), ("v", v), ("u_init", u), ("v_init", v), ("rho", rho), ("rho_init", rho), ("temperature", T), ("T_init", T)):
        ref.field(m)[...] = a
    store = {}
    for frame in range(FRAMES):
        for m in ref2d.MEMBERS:
            store[f"f{frame}:pre:{m}"] = ref.field(m).copy()
        c = ref.counters()
        store[f"f{frame}:pre:counters"] = np.array([c["last_remesh"], c["last_scalar_remesh"]], dtype=np.int32)
        ref.phase_a(DT, frame)
        for m in AFTER_A:
            store[f"f{frame}:a:{m}"] = ref.field(m).copy()
        adv = [ref.field(m).copy() for m in ("u", "v", "rho", "temperature")]
        ref.phase_b(DT, frame, *forcing(adv, DT))
        for m in AFTER_B:
            store[f"f{frame}:b:{m}"] = ref.field(m).copy()
        c, s = ref.counters(), ref.scalars()
        store[f"f{frame}:b:flags"] = np.array([c["vel_remap"], c["scalar_remap"]], dtype=np.int32)
        store[f"f{frame}:b:scalars"] = np.array([s["cfl"], s["vel_condition"], s["scalar_condition"], s["max_vel"]], dtype=np.float32)
    out = os.path.join(HERE, "ref2d_steps.npz")
    np.savez_compressed(out, **store)
    print("wrote", out, len(store), "arrays", os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
