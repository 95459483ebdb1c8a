"""Deterministic synthetic initial conditions for the parity tests and the benchmark.

The reference ships one 3D scene (two colliding emitter spheres, bimocq3D/main.cpp:28-80) that
needs OpenVDB; BASELINE.json asks for a smoke plume and vortex rings instead.  Both are closed
form here so that every box generates identical inputs:

* velocity = curl of a vector potential A = a(s, y) e_phi with a Gaussian core around a circle
  of radius R in the plane y = y0 (a vortex ring whose vorticity is concentrated in the core);
  being a curl it is divergence free, so no projection is needed to make the input physical.
  It is evaluated on the staggered MAC faces: u at (i, j+1/2, k+1/2) h, etc. -- in the
  reference's convention cell centre (i,j,k) sits at world position (i,j,k) h and faces are
  shifted by -h/2 along their axis (GPU_kernel.cu:67-69, :212).
* density = temperature = smooth ball (plume) or the ring-core indicator (rings).

All functions take ``xp`` = numpy or torch-like module (torch is used on the GPU box for 512^3).
"""
from __future__ import annotations

import math

import numpy as np


def _axes(xp, ni, nj, nk, h, kind, device=None):
    dx, dy, dz = {"u": (1, 0, 0), "v": (0, 1, 0), "w": (0, 0, 1), "c": (0, 0, 0)}[kind]

    def ar(n, shift):
        if xp is np:
            a = np.arange(n, dtype=np.float64)
        else:
            a = xp.arange(n, dtype=xp.float64, device=device)
        return (a - 0.5 * shift) * h

    x = ar(ni + dx, dx)[None, None, :]
    y = ar(nj + dy, dy)[None, :, None]
    z = ar(nk + dz, dz)[:, None, None]
    return x, y, z


def _ring_velocity_component(xp, x, y, z, comp, centre, R, core, strength):
    """comp-th component of curl(a e_phi) with a = strength*core*(s/R)*exp(-((s-R)^2+(y-y0)^2)/(2 core^2)),
    ring axis +y, s = distance from the axis (the s/R factor makes a vanish on the axis, so the
    field is smooth everywhere).  In cylindrical (s, phi, y):
    curl(a e_phi) = -da/dy e_s + (1/s) d(s a)/ds e_y."""
    cx, cy, cz = centre
    X = x - cx; Y = y - cy; Z = z - cz
    s = xp.sqrt(X * X + Z * Z)
    e = xp.exp(-((s - R) ** 2 + Y * Y) / (2.0 * core * core))
    A0 = strength * core / R
    if comp == 1:
        # (1/s) d(s * A0 s e)/ds = A0 (2 e + s de/ds)
        return A0 * e * (2.0 - s * (s - R) / (core * core))
    # -da/dy * (X or Z)/s = A0 e Y/core^2 * (X or Z)
    g = A0 * e * Y / (core * core)
    return g * X if comp == 0 else g * Z


def vortex_rings(ni, nj, nk, L=1.0, rings=((0.5, 0.2, 0.5, 0.12, 0.03, 1.0),), xp=np, device=None,
                 dtype=None):
    """Staggered velocity (u, v, w) of one or more coaxial (+y) vortex rings.
    rings: tuples (cx, cy, cz, radius, core, strength) in units of L."""
    h = L / ni
    out = []
    for comp, kind in enumerate("uvw"):
        x, y, z = _axes(xp, ni, nj, nk, h, kind, device)
        acc = None
        for (cx, cy, cz, R, core, strength) in rings:
            f = _ring_velocity_component(xp, x, y, z, comp, (cx * L, cy * L, cz * L), R * L, core * L, strength)
            acc = f if acc is None else acc + f
        if xp is np:
            out.append(np.ascontiguousarray(np.broadcast_to(acc, (nk + (kind == "w"), nj + (kind == "v"), ni + (kind == "u"))), dtype=np.float32))
        else:
            out.append(acc.expand(nk + (kind == "w"), nj + (kind == "v"), ni + (kind == "u")).to(dtype or xp.float32).contiguous())
    return out


def smooth_ball(ni, nj, nk, L=1.0, centre=(0.5, 0.15, 0.5), radius=0.1, xp=np, device=None, dtype=None):
    """Smoothed indicator of a ball: 1 inside, cosine roll-off over 3 cells."""
    h = L / ni
    x, y, z = _axes(xp, ni, nj, nk, h, "c", device)
    r = xp.sqrt((x - centre[0] * L) ** 2 + (y - centre[1] * L) ** 2 + (z - centre[2] * L) ** 2)
    t = (r - radius * L) / (3.0 * h)
    if xp is np:
        t = np.clip(t, -1.0, 1.0)
        f = 0.5 - 0.5 * np.sin(0.5 * math.pi * t)
        return np.ascontiguousarray(f, dtype=np.float32)
    t = xp.clamp(t, -1.0, 1.0)
    f = 0.5 - 0.5 * xp.sin(0.5 * math.pi * t)
    return f.to(dtype or xp.float32).contiguous()


def smoke_plume(ni, nj, nk, L=1.0, xp=np, device=None):
    """BASELINE configs[2] / configs[4]: ring velocity + density = temperature = smooth ball."""
    u, v, w = vortex_rings(ni, nj, nk, L, ((0.5, 0.2, 0.5, 0.12, 0.03, 1.0),), xp, device)
    rho = smooth_ball(ni, nj, nk, L, (0.5, 0.15, 0.5), 0.1, xp, device)
    T = rho.copy() if xp is np else rho.clone()
    return u, v, w, rho, T


def leapfrog_rings(ni, nj, nk, L=1.0, xp=np, device=None):
    """BASELINE configs[3]: two coaxial rings (radii 0.10 and 0.14) that leapfrog; scalars mark
    the ring cores."""
    rings = ((0.5, 0.2, 0.5, 0.10, 0.03, 1.0), (0.5, 0.26, 0.5, 0.14, 0.03, 1.0))
    u, v, w = vortex_rings(ni, nj, nk, L, rings, xp, device)
    h = L / ni
    x, y, z = _axes(xp, ni, nj, nk, h, "c", device)
    acc = None
    for (cx, cy, cz, R, core, _) in rings:
        s = xp.sqrt((x - cx * L) ** 2 + (z - cz * L) ** 2 + 1e-30)
        d2 = (s - R * L) ** 2 + (y - cy * L) ** 2
        f = xp.exp(-d2 / (2.0 * (core * L) ** 2))
        acc = f if acc is None else acc + f
    if xp is np:
        rho = np.ascontiguousarray(acc, dtype=np.float32)
        return u, v, w, rho, rho.copy()
    rho = acc.to(xp.float32).contiguous()
    return u, v, w, rho, rho.clone()


def scale_to_cfl(u, v, w, h, dt, cfl):
    """Scale a velocity field so that dt * max|vel| / h == cfl (n_sub = ceil(cfl) CFL sub-steps)."""
    m = max(float(abs(a).max()) for a in (u, v, w))
    s = cfl * h / (dt * m)
    return u * s, v * s, w * s


def buoyancy_increment(rho, T, alpha, beta, dt, nj_faces):
    """Reference buoyancy (GPU_kernel.cu:804-823): dv = dt*(-alpha*rho + beta*T) averaged onto the
    v faces between cells j-1 and j; zero on the outermost faces.  Returns dv with nj+1 rows."""
    xp = np if isinstance(rho, np.ndarray) else None
    f = dt * (-alpha * rho + beta * T)
    if xp is np:
        dv = np.zeros((rho.shape[0], nj_faces, rho.shape[2]), dtype=np.float32)
        dv[:, 1:-1, :] = 0.5 * (f[:, 1:, :] + f[:, :-1, :])
        return dv
    import torch
    dv = torch.zeros((rho.shape[0], nj_faces, rho.shape[2]), dtype=torch.float32, device=rho.device)
    dv[:, 1:-1, :] = 0.5 * (f[:, 1:, :] + f[:, :-1, :])
    return dv


def smooth_random(shape, seed, amplitude=1.0, modes=3):
    """Seeded smooth random field (a few low-wavenumber sines) + 5% white noise: used by the
    kernel-level parity tests, where cell-scale variation matters more than physics."""
    rng = np.random.default_rng(seed)
    nz, ny, nx = shape
    z, y, x = np.meshgrid(np.arange(nz) / nz, np.arange(ny) / ny, np.arange(nx) / nx, indexing="ij")
    f = np.zeros(shape, dtype=np.float64)
    for _ in range(modes):
        kx, ky, kz = rng.integers(1, 4, size=3)
        ph = rng.uniform(0, 2 * np.pi, size=3)
        f += rng.uniform(0.3, 1.0) * np.sin(2 * np.pi * kx * x + ph[0]) * np.sin(2 * np.pi * ky * y + ph[1]) * \
            np.sin(2 * np.pi * kz * z + ph[2])
    f += 0.05 * rng.standard_normal(shape)
    f *= amplitude / np.abs(f).max()
    return np.ascontiguousarray(f, dtype=np.float32)
