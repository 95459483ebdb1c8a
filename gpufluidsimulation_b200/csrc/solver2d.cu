// 2D BiMocq^2 advection path for sm_100a: kernels + the handle API bmq2d_*.
//
// The 2D reference is pure CPU code (TBB lambdas over Array2f members,
// bimocq2D/BimocqSolver2D.cpp); there is no device seam to keep, so the seam is created behind
// BimocqSolver2D::advanceBIMOCQ (:390-508): bmq2d_advect = lines 394-445, bmq2d_accumulate =
// lines 455-507, the caller's forces + projection sit in between (like the 3D handle API).
//
// Numerics: the reference is compiled for the host without FMA, so every float operation rounds
// separately.  All arithmetic here uses __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn in the
// reference's operation order (no contraction), which makes every kernel bit-identical to the
// reference on the same inputs -- except exp() in the DMC update (glibc expf vs the device; a
// double-precision exp rounded to float is used, which agrees with glibc except in rare
// last-ulp cases).
//
// Layout: row-major a[i + ni*j] (array2.h:93-103).  Cell centre (i+1/2, j+1/2) h, u at
// (i, j+1/2) h, v at (i+1/2, j) h.
#include "common.h"
#include "launch3d.h"

#include <cmath>
#include <cstring>
#include <new>
#include <utility>

namespace {

struct G2 {
    int ni, nj;
    float h;
    float inv_h;   // RN(1/h)
    int div_mode;  // how pos / h is evaluated: 0 IEEE division, 1 exact multiplication (h a power of two), 2 the
                   // three-instruction sequence of device3d.cuh:div_h, verified for this h (kernels3d.cu)
};

#define FM(a, b) __fmul_rn((a), (b))
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
#define FD(a, b) __fdiv_rn((a), (b))

struct V2 {
    float x, y;
};

// pos / h as the reference's CPU code computes it (a correctly rounded float division), without the ~10
// instruction IEEE division sequence where it can be avoided bit for bit
__device__ __forceinline__ float div_cell(float p, const G2 &g)
{
    if (g.div_mode == 1) return FM(p, g.inv_h);
    if (g.div_mode == 2 && (fabsf(p) >= 7.888609052e-31f || p == 0.f)) {
        const float q0 = FM(p, g.inv_h);
        const float r = __fmaf_rn(-g.h, q0, p);
        return __fmaf_rn(r, g.inv_h, q0);
    }
    return FD(p, g.h);
}

// lerp / bilerp, BimocqSolver2D.cpp:71-79
__device__ __forceinline__ float lerp2(float v0, float v1, float c) { return FA(FM(FS(1.0f, c), v0), FM(c, v1)); }
__device__ __forceinline__ float bilerp2(float v00, float v01, float v10, float v11, float cx, float cy)
{
    return lerp2(lerp2(v00, v01, cx), lerp2(v10, v11, cx), cy);
}

// sampleField, :2328-2334 with Array2::boundedAt (array2.h:273-284)
__device__ __forceinline__ float sample_field(const float *__restrict__ f, int fni, int fnj, const G2 &g, float px, float py)
{
    const float qx = div_cell(px, g), qy = div_cell(py, g);
    const int i = (int)floorf(qx), j = (int)floorf(qy);
    const int i0 = min(max(i, 0), fni - 1), i1 = min(max(i + 1, 0), fni - 1);
    const int j0 = min(max(j, 0), fnj - 1), j1 = min(max(j + 1, 0), fnj - 1);
    return bilerp2(__ldg(f + i0 + fni * j0), __ldg(f + i1 + fni * j0), __ldg(f + i0 + fni * j1),
                   __ldg(f + i1 + fni * j1), FS(qx, (float)i), FS(qy, (float)j));
}

// getVelocity, :2307-2325: zero outside the sampled component's stencil
__device__ __forceinline__ V2 get_velocity(const G2 &g, const float *__restrict__ u, const float *__restrict__ v, V2 pos)
{
    const float hh = (float)(0.5 * (double)g.h);
    V2 r;
    {
        const float ux = FS(pos.x, 0.0f), uy = FS(pos.y, hh);
        const float qx = div_cell(ux, g), qy = div_cell(uy, g);
        const int i = (int)floorf(qx), j = (int)floorf(qy);
        if (!(i >= 0 && i <= g.ni - 1 && j >= 0 && j <= g.nj - 2)) r.x = 0.f;
        else {
            const int n = g.ni + 1;
            r.x = bilerp2(__ldg(u + i + n * j), __ldg(u + i + 1 + n * j), __ldg(u + i + n * (j + 1)),
                          __ldg(u + i + 1 + n * (j + 1)), FS(qx, (float)i), FS(qy, (float)j));
        }
    }
    {
        const float vx = FS(pos.x, hh), vy = FS(pos.y, 0.0f);
        const float qx = div_cell(vx, g), qy = div_cell(vy, g);
        const int i = (int)floorf(qx), j = (int)floorf(qy);
        if (!(i >= 0 && i <= g.ni - 2 && j >= 0 && j <= g.nj - 1)) r.y = 0.f;
        else {
            const int n = g.ni;
            r.y = bilerp2(__ldg(v + i + n * j), __ldg(v + i + 1 + n * j), __ldg(v + i + n * (j + 1)),
                          __ldg(v + i + 1 + n * (j + 1)), FS(qx, (float)i), FS(qy, (float)j));
        }
    }
    return r;
}

// traceRK3, :4-19
__device__ __forceinline__ V2 trace_rk3(const G2 &g, const float *u, const float *v, float dt, V2 pos)
{
    const float c1 = (float)(2.0 / 9.0 * (double)dt), c2 = (float)(3.0 / 9.0 * (double)dt), c3 = (float)(4.0 / 9.0 * (double)dt);
    const float hd = (float)(0.5 * (double)dt), qd = (float)(0.75 * (double)dt);
    const V2 v1 = get_velocity(g, u, v, pos);
    const V2 m1 = {FA(pos.x, FM(hd, v1.x)), FA(pos.y, FM(hd, v1.y))};
    const V2 v2 = get_velocity(g, u, v, m1);
    const V2 m2 = {FA(pos.x, FM(qd, v2.x)), FA(pos.y, FM(qd, v2.y))};
    const V2 v3 = get_velocity(g, u, v, m2);
    V2 o;
    o.x = FA(FA(FA(pos.x, FM(c1, v1.x)), FM(c2, v2.x)), FM(c3, v3.x));
    o.y = FA(FA(FA(pos.y, FM(c1, v1.y)), FM(c2, v2.y)), FM(c3, v3.y));
    const float e = FM(0.001f, g.h);
    o.x = fminf(fmaxf(e, o.x), FS(FM((float)g.ni, g.h), e));
    o.y = fminf(fmaxf(e, o.y), FS(FM((float)g.nj, g.h), e));
    return o;
}

// solveODE, :21-43: compare one step with two half steps, halve until they agree (<= 6 rounds)
__device__ V2 solve_ode(const G2 &g, const float *u, const float *v, float dt, V2 pos)
{
    float ddt = dt;
    V2 pos1 = trace_rk3(g, u, v, ddt, pos);
    ddt = (float)((double)ddt / 2.0);
    int substeps = 2;
    V2 pos2 = trace_rk3(g, u, v, ddt, pos);
    pos2 = trace_rk3(g, u, v, ddt, pos2);
    int iter = 0;
    const double thr = 0.0001 * (double)g.h;
    for (;;) {
        const float dx = FS(pos2.x, pos1.x), dy = FS(pos2.y, pos1.y);
        const float d = sqrtf(FA(FM(dx, dx), FM(dy, dy)));
        if (!((double)d > thr && iter < 6)) break;
        pos1 = pos2;
        ddt = (float)((double)ddt / 2.0);
        substeps *= 2;
        pos2 = pos;
        for (int j = 0; j < substeps; ++j) pos2 = trace_rk3(g, u, v, ddt, pos2);
        ++iter;
    }
    return pos2;
}

// First round of solveODE only: one full step against two half steps.  Returns true (and the
// result) when the reference's loop would stop right there, which is the case almost everywhere;
// the few cells that keep halving (next to walls, where traceRK3's clamp and getVelocity's
// zero-outside rule make the trace non-smooth; up to 254 RK3 steps) are deferred to a compacted
// second pass so that they do not hold whole warps of converged cells hostage.
__device__ __forceinline__ bool solve_ode_quick(const G2 &g, const float *u, const float *v, float dt, V2 pos, V2 &out)
{
    const V2 pos1 = trace_rk3(g, u, v, dt, pos);
    const float ddt = (float)((double)dt / 2.0);
    V2 pos2 = trace_rk3(g, u, v, ddt, pos);
    pos2 = trace_rk3(g, u, v, ddt, pos2);
    const float dx = FS(pos2.x, pos1.x), dy = FS(pos2.y, pos1.y);
    const float d = sqrtf(FA(FM(dx, dx), FM(dy, dy)));
    out = pos2;
    return !((double)d > 0.0001 * (double)g.h);
}

__device__ __forceinline__ void clamp_pos(const G2 &g, V2 &p)   // BimocqSolver2D.h:128-132
{
    p.x = fminf(fmaxf(g.h, p.x), FS(FM((float)g.ni, g.h), g.h));
    p.y = fminf(fmaxf(g.h, p.y), FS(FM((float)g.nj, g.h), g.h));
}

// exp(float) of the host reference is glibc's expf: (almost always) the correctly rounded value
__device__ __forceinline__ float exp_host_like(float x) { return (float)exp((double)x); }

// solveODEDMC = calculateA + traceDMC, :45-51, :58-69, :81-91
__device__ V2 solve_ode_dmc(const G2 &g, const float *u, const float *v, float dt, V2 pos)
{
    const V2 vel = get_velocity(g, u, v, pos);
    V2 np = {vel.x > 0.f ? FS(pos.x, g.h) : FA(pos.x, g.h), vel.y > 0.f ? FS(pos.y, g.h) : FA(pos.y, g.h)};
    const V2 nv = get_velocity(g, u, v, np);
    const float ax = FD(FS(vel.x, nv.x), FS(pos.x, np.x)), ay = FD(FS(vel.y, nv.y), FS(pos.y, np.y));
    const bool ex = (double)fabsf(ax) > 1e-4, ey = (double)fabsf(ay) > 1e-4;
    V2 fb = {0.f, 0.f};
    if (!ex || !ey) fb = solve_ode(g, u, v, -dt, pos);
    V2 o;
    o.x = ex ? FS(pos.x, FD(FM(FS(1.0f, exp_host_like(FM(-ax, dt))), vel.x), ax)) : fb.x;
    o.y = ey ? FS(pos.y, FD(FM(FS(1.0f, exp_host_like(FM(-ay, dt))), vel.y), ay)) : fb.y;
    return o;
}

// position of sample k of the 5-point quadrature for element (i,j) with face offset (offx,offy):
// h*Vec2f(i,j) + h*Vec2f(offx,offy) + h*dir[k]   (e.g. :749, :852)
__device__ __forceinline__ V2 quad_pos(const G2 &g, int i, int j, float offx, float offy, int k)
{
    const float dx = (k == 4) ? 0.0f : ((k & 1) ? 0.25f : -0.25f);
    const float dy = (k == 4) ? 0.0f : ((k & 2) ? 0.25f : -0.25f);
    V2 p;
    p.x = FA(FA(FM(g.h, (float)i), FM(g.h, offx)), FM(g.h, dx));
    p.y = FA(FA(FM(g.h, (float)j), FM(g.h, offy)), FM(g.h, dy));
    return p;
}
__device__ __forceinline__ float quad_w(int k) { return k == 4 ? 0.5f : 0.125f; }

// map a position through a cell-centred map pair: sampleField(pos - h*Vec2f(0.5), mx/my), clampPos
__device__ __forceinline__ V2 map_through(const G2 &g, const float *mx, const float *my, V2 pos)
{
    const float hh = FM(g.h, 0.5f);
    const float sx = FS(pos.x, hh), sy = FS(pos.y, hh);
    V2 r = {sample_field(mx, g.ni, g.nj, g, sx, sy), sample_field(my, g.ni, g.nj, g, sx, sy)};
    clamp_pos(g, r);
    return r;
}

#define IJ(fni, fnj)                                     \
    const int i = blockIdx.x * 32 + threadIdx.x;         \
    const int j = blockIdx.y * 8 + threadIdx.y;          \
    if (i >= (fni) || j >= (fnj)) return;                \
    const int idx = i + (fni) * j;

// ---------------------------------------------------------------- kernels
// Work list of the cells whose solveODE has not converged yet: the cell and the result of its last round (the
// reference's pos1 of the next comparison).
struct WorkList {
    int *count;   // device counter
    int *items;   // flat element indices
    float *px, *py;
};
// Warp-aggregated append: the lanes of a warp that defer take consecutive slots (one atomicAdd per warp).
__device__ __forceinline__ void defer(const WorkList &wl, int idx, V2 last)
{
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(wl.count, __popc(m));
    base = __shfl_sync(m, base, leader);
    const int e = base + __popc(m & ((1u << lane) - 1u));
    wl.items[e] = idx;
    wl.px[e] = last.x;
    wl.py[e] = last.y;
}

// Round r = 1..6 of solveODE (:30-41) for ONE cell whose previous round gave `last`: 2^(r+1) RK3 steps of dt / 2^(r+1)
// from the start position; returns true when the reference's loop stops after this round (the two last rounds agree
// to 1e-4 h, or six halvings are done).  Rounds are launched one after the other over compacted lists, so a warp
// never waits for one lane's 254 steps while its other lanes are done after 4 (the single slow pass of round 1 did).
__device__ __forceinline__ bool solve_ode_round(const G2 &g, const float *u, const float *v, float dt, int r, V2 pos, V2 last, V2 &out)
{
    float ddt = dt;
    int substeps = 1;
    for (int q = 0; q <= r; ++q) {
        ddt = (float)((double)ddt / 2.0);
        substeps *= 2;
    }
    V2 pos2 = pos;
    for (int j = 0; j < substeps; ++j) pos2 = trace_rk3(g, u, v, ddt, pos2);
    const float dx = FS(pos2.x, last.x), dy = FS(pos2.y, last.y);
    const float d = sqrtf(FA(FM(dx, dx), FM(dy, dy)));
    out = pos2;
    return !((double)d > 0.0001 * (double)g.h) || r >= 6;
}

// updateForward, :1228-1240 (all cells): every cell, first solveODE round only, non-converged cells deferred.
__global__ void __launch_bounds__(256) k2_forward(G2 g, const float *u, const float *v, float *fx, float *fy, float dt, WorkList wl)
{
    IJ(g.ni, g.nj)
    V2 p = {fx[idx], fy[idx]}, q;
    if (!solve_ode_quick(g, u, v, dt, p, q)) { defer(wl, idx, q); return; }
    clamp_pos(g, q);
    fx[idx] = q.x;
    fy[idx] = q.y;
}
// one later round over the deferred cells (in place: a cell's start position is read by its own thread only)
__global__ void __launch_bounds__(128) k2_forward_round(G2 g, const float *u, const float *v, float *fx, float *fy, float dt, int r,
                                                        WorkList in, WorkList out, int *entered)
{
    const int n = *in.count;
    if (blockIdx.x == 0 && threadIdx.x == 0) entered[r] = n;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int idx = in.items[e];
        const V2 p = {fx[idx], fy[idx]}, last = {in.px[e], in.py[e]};
        V2 q;
        if (!solve_ode_round(g, u, v, dt, r, p, last, q)) { defer(out, idx, q); continue; }
        clamp_pos(g, q);
        fx[idx] = q.x;
        fy[idx] = q.y;
    }
}

// one sub-step of updateBackward (:1242-1259): semiLagAdvectDMC for x and y maps from one back-trace
__global__ void __launch_bounds__(256)
k2_backward(G2 g, const float *u, const float *v, const float *bx, const float *by, float *ox, float *oy, float substep)
{
    IJ(g.ni, g.nj)
    const float hh = FM(g.h, 0.5f);
    V2 pos = {FA(FM(g.h, (float)i), hh), FA(FM(g.h, (float)j), hh)};
    V2 b = solve_ode_dmc(g, u, v, substep, pos);
    clamp_pos(g, b);
    const float sx = FS(b.x, hh), sy = FS(b.y, hh);
    ox[idx] = sample_field(bx, g.ni, g.nj, g, sx, sy);
    oy[idx] = sample_field(by, g.ni, g.nj, g, sx, sy);
}

// semiLagAdvect, :110-123 (same scheme as k2_forward: quick pass, then one launch per later round)
__global__ void __launch_bounds__(256)
k2_semilag(G2 g, const float *u, const float *v, const float *src, float *dst, int fni, int fnj, float offx, float offy,
           float dt, WorkList wl)
{
    const float ox = FM(g.h, offx), oy = FM(g.h, offy);
    IJ(fni, fnj)
    V2 pos = {FA(FM(g.h, (float)i), ox), FA(FM(g.h, (float)j), oy)}, b;
    if (!solve_ode_quick(g, u, v, -dt, pos, b)) { defer(wl, idx, b); return; }
    dst[idx] = sample_field(src, fni, fnj, g, FS(b.x, ox), FS(b.y, oy));
}
__global__ void __launch_bounds__(128)
k2_semilag_round(G2 g, const float *u, const float *v, const float *src, float *dst, int fni, int fnj, float offx, float offy,
                 float dt, int r, WorkList in, WorkList out, int *entered)
{
    const float ox = FM(g.h, offx), oy = FM(g.h, offy);
    const int n = *in.count;
    if (blockIdx.x == 0 && threadIdx.x == 0) entered[r] = n;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const int idx = in.items[e];
        const int i = idx % fni, j = idx / fni;
        const V2 pos = {FA(FM(g.h, (float)i), ox), FA(FM(g.h, (float)j), oy)}, last = {in.px[e], in.py[e]};
        V2 b;
        if (!solve_ode_round(g, u, v, -dt, r, pos, last, b)) { defer(out, idx, b); continue; }
        dst[idx] = sample_field(src, fni, fnj, g, FS(b.x, ox), FS(b.y, oy));
    }
}

struct AdvectArgs {
    const float *bx, *by, *bxp, *byp;        // chi, chi_prev
    const float *f_init, *f_orig, *d, *d_prev, *semi;
    float *out;
    int fni, fnj;
    float offx, offy, blend;
    int i_lo, i_hi, j_lo, j_hi;              // interior: i_lo < i < i_hi && j_lo < j < j_hi
};

// advectVelocity / advectScalars, :933-1077
__global__ void __launch_bounds__(256) k2_advect(G2 g, AdvectArgs a)
{
    IJ(a.fni, a.fnj)
    if (!(i > a.i_lo && i < a.i_hi && j > a.j_lo && j < a.j_hi)) { a.out[idx] = __ldg(a.semi + idx); return; }
    const float ox = FM(g.h, a.offx), oy = FM(g.h, a.offy);
    float acc = 0.f;
    const float omb = FS(1.f, a.blend);
#pragma unroll 1
    for (int k = 0; k < 5; ++k) {
        const V2 pos = quad_pos(g, i, j, a.offx, a.offy, k);
        const V2 p1 = map_through(g, a.bx, a.by, pos);
        const V2 p2 = map_through(g, a.bxp, a.byp, p1);
        const float w = quad_w(k);
        const float s_orig = sample_field(a.f_orig, a.fni, a.fnj, g, FS(p2.x, ox), FS(p2.y, oy));
        const float s_d1 = sample_field(a.d, a.fni, a.fnj, g, FS(p1.x, ox), FS(p1.y, oy));
        const float s_dp = sample_field(a.d_prev, a.fni, a.fnj, g, FS(p2.x, ox), FS(p2.y, oy));
        acc = FA(acc, FM(FM(omb, w), FA(FA(s_orig, s_d1), s_dp)));
        const float s_init = sample_field(a.f_init, a.fni, a.fnj, g, FS(p1.x, ox), FS(p1.y, oy));
        acc = FA(acc, FM(FM(a.blend, w), FA(s_init, s_d1)));
    }
    a.out[idx] = acc;
}

struct CorrectArgs {
    const float *mx, *my;     // forward map (pass 1) / backward map (pass 2)
    const float *src;         // pass 1: advected field; pass 2: temp
    const float *d, *f_init;  // pass 1 only
    float *out;               // pass 1: temp; pass 2: field (read-modify-write)
    int fni, fnj;
    float offx, offy;
    int i_lo, i_hi, j_lo, j_hi;
};

// first half of correctVelocity / correctScalars (:735-758, :839-861): temp = 0.5*(sum_k w(f(psi) - d) - f_init)
__global__ void __launch_bounds__(256) k2_correct_error(G2 g, CorrectArgs a)
{
    IJ(a.fni, a.fnj)
    float t = 0.f;
    if (i > a.i_lo && i < a.i_hi && j > a.j_lo && j < a.j_hi) {
        const float ox = FM(g.h, a.offx), oy = FM(g.h, a.offy);
        const float dv = __ldg(a.d + idx);
#pragma unroll 1
        for (int k = 0; k < 5; ++k) {
            const V2 p1 = map_through(g, a.mx, a.my, quad_pos(g, i, j, a.offx, a.offy, k));
            t = FA(t, FM(quad_w(k), FS(sample_field(a.src, a.fni, a.fnj, g, FS(p1.x, ox), FS(p1.y, oy)), dv)));
        }
    }
    a.out[idx] = FM(FS(t, __ldg(a.f_init + idx)), 0.5f);
}

// second half (:759-780, :862-883): f -= sum_k w * temp(chi(x_k))
__global__ void __launch_bounds__(256) k2_correct_apply(G2 g, CorrectArgs a)
{
    IJ(a.fni, a.fnj)
    if (!(i > a.i_lo && i < a.i_hi && j > a.j_lo && j < a.j_hi)) return;
    const float ox = FM(g.h, a.offx), oy = FM(g.h, a.offy);
    float f = a.out[idx];
#pragma unroll 1
    for (int k = 0; k < 5; ++k) {
        const V2 p1 = map_through(g, a.mx, a.my, quad_pos(g, i, j, a.offx, a.offy, k));
        f = FS(f, FM(quad_w(k), sample_field(a.src, a.fni, a.fnj, g, FS(p1.x, ox), FS(p1.y, oy))));
    }
    a.out[idx] = f;
}

// clampExtrema2, :1261-1274.  The reference indexes before.at(ii,jj) = a[ii + ni*jj] without
// bounds checks: in columns i=0 / i=ni-1 that wraps into the neighbouring row (reproduced here);
// in rows j=0 / j=nj-1 it reads outside the array (undefined) -- there the index is clamped.
__global__ void __launch_bounds__(256) k2_clamp_extrema(int fni, int fnj, const float *before, float *after)
{
    IJ(fni, fnj)
    float mn = 1e+6f, mx = 0.f;
    const int n = fni * fnj;
#pragma unroll
    for (int jj = -1; jj <= 1; ++jj)
#pragma unroll
        for (int ii = -1; ii <= 1; ++ii) {
            int q = (i + ii) + fni * (j + jj);
            if (q < 0 || q >= n) q = min(max(i + ii, 0), fni - 1) + fni * min(max(j + jj, 0), fnj - 1);
            const float b = __ldg(before + q);
            mx = fmaxf(mx, b);
            mn = fminf(mn, b);
        }
    after[idx] = fminf(fmaxf(after[idx], mn), mx);
}

// live part of accumulateVelocity / accumulateScalars (:1131-1157, :1391-1424):
// d += sum_k (w*coeff) * change(psi(x_k)); NCH change sets in the reference's call order
struct AccumArgs {
    const float *mx, *my;
    const float *change[2];
    float coeff[2];
    int nch;
    float *d;
    int fni, fnj;
    float offx, offy;
    int i_lo, i_hi, j_lo, j_hi;
    int scalar_form;   // scalars use w*sample (no coeff factor in the expression)
};
__global__ void __launch_bounds__(256) k2_accumulate(G2 g, AccumArgs a)
{
    IJ(a.fni, a.fnj)
    if (!(i > a.i_lo && i < a.i_hi && j > a.j_lo && j < a.j_hi)) return;
    const float ox = FM(g.h, a.offx), oy = FM(g.h, a.offy);
    V2 p[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) p[k] = map_through(g, a.mx, a.my, quad_pos(g, i, j, a.offx, a.offy, k));
    float d = a.d[idx];
    for (int c = 0; c < a.nch; ++c) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float s = sample_field(a.change[c], a.fni, a.fnj, g, FS(p[k].x, ox), FS(p[k].y, oy));
            d = a.scalar_form ? FA(d, FM(quad_w(k), s)) : FA(d, FM(FM(quad_w(k), a.coeff[c]), s));
        }
    }
    a.d[idx] = d;
}

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ void block_atomic_max_nonneg(float v, float *dst)
{
    __shared__ float sm[8];
    const int tid = threadIdx.x + 32 * threadIdx.y;
    v = warp_max(v);
    if ((tid & 31) == 0) sm[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        float r = tid < 8 ? sm[tid] : 0.f;
        r = warp_max(r);
        if (tid == 0 && r > 0.f) atomicMax(reinterpret_cast<int *>(dst), __float_as_int(r));
    }
    __syncthreads();
}

// estimateDistortion, :666-697, both map pairs, max-reduced on the device
__global__ void __launch_bounds__(256)
k2_distortion(G2 g, const float *bx, const float *by, const float *fx, const float *fy, const float *sbx,
              const float *sby, const float *sfx, const float *sfy, float *out2)
{
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 8 + threadIdx.y;
    float d[2] = {0.f, 0.f};
    if (i < g.ni && j < g.nj && i > 2 && i < g.ni - 3 && j > 2 && j < g.nj - 3) {
        const int idx = i + g.ni * j;
        const float hh = FM(0.5f, g.h);
        const float ipx = FM(g.h, FA((float)i, 0.5f)), ipy = FM(g.h, FA((float)j, 0.5f));
        const float *B[2][2] = {{bx, by}, {sbx, sby}}, *F[2][2] = {{fx, fy}, {sfx, sfy}};
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const float f0x = __ldg(F[m][0] + idx), f0y = __ldg(F[m][1] + idx);
            float qx = sample_field(B[m][0], g.ni, g.nj, g, FS(f0x, hh), FS(f0y, hh));
            float qy = sample_field(B[m][1], g.ni, g.nj, g, FS(f0x, hh), FS(f0y, hh));
            float ex = FS(qx, ipx), ey = FS(qy, ipy);
            float dd = sqrtf(FA(FM(ex, ex), FM(ey, ey)));
            const float b0x = __ldg(B[m][0] + idx), b0y = __ldg(B[m][1] + idx);
            qx = sample_field(F[m][0], g.ni, g.nj, g, FS(b0x, hh), FS(b0y, hh));
            qy = sample_field(F[m][1], g.ni, g.nj, g, FS(b0x, hh), FS(b0y, hh));
            ex = FS(qx, ipx); ey = FS(qy, ipy);
            dd = fmaxf(dd, sqrtf(FA(FM(ex, ex), FM(ey, ey))));
            d[m] = dd;
        }
    }
    block_atomic_max_nonneg(d[0], out2);
    block_atomic_max_nonneg(d[1], out2 + 1);
}

// maxVel, :699-725: SIGNED max of u then v, starting from 0
__global__ void __launch_bounds__(256) k2_maxvel(const float *u, int nu, const float *v, int nv, float *out)
{
    float m = 0.f;
    const int stride = gridDim.x * blockDim.x;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nu; e += stride) m = fmaxf(m, __ldg(u + e));
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nv; e += stride) m = fmaxf(m, __ldg(v + e));
    __shared__ float sm[8];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        float r = threadIdx.x < 8 ? sm[threadIdx.x] : 0.f;
        r = warp_max(r);
        if (threadIdx.x == 0 && r > 0.f) atomicMax(reinterpret_cast<int *>(out), __float_as_int(r));
    }
}

__global__ void __launch_bounds__(256) k2_identity(G2 g, float *x0, float *y0, float *x1, float *y1)
{
    IJ(g.ni, g.nj)
    const float x = (float)((double)g.h * ((double)(float)i + 0.5)), y = (float)((double)g.h * ((double)(float)j + 0.5));
    x0[idx] = x; y0[idx] = y;
    if (x1) { x1[idx] = x; y1[idx] = y; }
}

__global__ void __launch_bounds__(256) k2_sub(float *out, const float *a, const float *b, int n)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = FS(a[e], b[e]);
}
// u = 0.5*(u_presave + u), :497-506 (0.5 is a double literal: exact halving of the float sum)
__global__ void __launch_bounds__(256) k2_average(float *u, const float *pre, int n)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) u[e] = FM(0.5f, FA(pre[e], u[e]));
}

}  // namespace

// ================================================================================ handle
struct bmq2d_solver {
    int ni, nj;
    float h, blend;
    G2 g;
    cudaStream_t stream = 0;
    float *f[BMQ2_F_COUNT] = {};
    int fni[BMQ2_F_COUNT], fnj[BMQ2_F_COUNT];
    float *d_red = nullptr, *h_red = nullptr;
    // 6 work lists x 2 buffers (the rounds ping-pong) of wl_stride words: counter, then `wl_cap` cell indices, last x, last y
    int *d_wl = nullptr;
    size_t wl_stride = 0, wl_cap = 0;
    int *d_round_counts = nullptr;   // [6 lists][8]: cells entering round 1..6 of the last step (diagnostic)
    cudaStream_t side[6] = {};
    cudaEvent_t ev_quick = nullptr, ev_side[6] = {};
    int pending_slow = 0;      // bit w set: slow pass w has been recorded on side[w]
    int lastremeshing = 0, rho_lastremeshing = 0, total_resample = 0, total_scalar_resample = 0;
    bool levelset = false;
    float cfl = 0.f;
    bmq2d_stats stats;
    unsigned long long launches = 0;
};

namespace {

enum Kind { KU, KV, KC };
Kind kind_of(int id)
{
    switch (id) {
    case BMQ2_F_U: case BMQ2_F_U_TEMP: case BMQ2_F_U_INIT: case BMQ2_F_U_ORIG: case BMQ2_F_DU: case BMQ2_F_DU_PREV:
    case BMQ2_F_DU_EXT: case BMQ2_F_DU_PROJ: case BMQ2_F_U_FORCED: case BMQ2_F_U_PRESAVE: case BMQ2_F_U_SAVE: case BMQ2_F_U_SEMI:
    case BMQ2_F_U_SCRATCH: case BMQ2_F_U_SCRATCH2:
        return KU;
    case BMQ2_F_V: case BMQ2_F_V_TEMP: case BMQ2_F_V_INIT: case BMQ2_F_V_ORIG: case BMQ2_F_DV: case BMQ2_F_DV_PREV:
    case BMQ2_F_DV_EXT: case BMQ2_F_DV_PROJ: case BMQ2_F_V_FORCED: case BMQ2_F_V_PRESAVE: case BMQ2_F_V_SAVE: case BMQ2_F_V_SEMI:
    case BMQ2_F_V_SCRATCH: case BMQ2_F_V_SCRATCH2:
        return KV;
    default:
        return KC;
    }
}

dim3 blk() { return dim3(32, 8, 1); }
dim3 grd(int a, int b) { return dim3((a + 31) / 32, (b + 7) / 8, 1); }
#define L2D(s) ((s)->launches++)

int ident(bmq2d_solver *s, int x0, int y0, int x1, int y1)
{
    k2_identity<<<grd(s->ni, s->nj), blk(), 0, s->stream>>>(s->g, s->f[x0], s->f[y0], x1 >= 0 ? s->f[x1] : nullptr,
                                                          y1 >= 0 ? s->f[y1] : nullptr);
    L2D(s);
    BMQ_CK(cudaGetLastError());
    return BMQ_OK;
}
int copyf(bmq2d_solver *s, int dst, int src)
{
    BMQ_CK(cudaMemcpyAsync(s->f[dst], s->f[src], sizeof(float) * s->fni[src] * s->fnj[src], cudaMemcpyDeviceToDevice, s->stream));
    return BMQ_OK;
}
int zerof(bmq2d_solver *s, int id)
{
    BMQ_CK(cudaMemsetAsync(s->f[id], 0, sizeof(float) * s->fni[id] * s->fnj[id], s->stream));
    return BMQ_OK;
}

struct FieldGeom { int fni, fnj; float offx, offy; };
FieldGeom geom(const bmq2d_solver *s, Kind k)
{
    if (k == KU) return {s->ni + 1, s->nj, 0.0f, 0.5f};
    if (k == KV) return {s->ni, s->nj + 1, 0.5f, 0.0f};
    return {s->ni, s->nj, 0.5f, 0.5f};
}

// maxVel() + 1e-5 (double add, float result), getCFL :53-56
int max_vel(bmq2d_solver *s, float *out)
{
    BMQ_CK(cudaMemsetAsync(s->d_red, 0, sizeof(float), s->stream));
    k2_maxvel<<<148, 256, 0, s->stream>>>(s->f[BMQ2_F_U], (s->ni + 1) * s->nj, s->f[BMQ2_F_V], s->ni * (s->nj + 1), s->d_red);
    L2D(s);
    BMQ_CK(cudaGetLastError());
    BMQ_CK(cudaMemcpyAsync(s->h_red, s->d_red, sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    BMQ_CK(cudaStreamSynchronize(s->stream));
    *out = (float)((double)s->h_red[0] + 1e-5);
    return BMQ_OK;
}

WorkList worklist(bmq2d_solver *s, int w, int buf = 0)
{
    int *base = s->d_wl + (size_t)(2 * w + buf) * s->wl_stride;
    float *fl = reinterpret_cast<float *>(base);
    return WorkList{base, base + 4, fl + 4 + s->wl_cap, fl + 4 + 2 * s->wl_cap};
}
// the rounds of one work list, one launch each, on its side stream; the grid covers the list of round 1 and later
// rounds leave most CTAs idle at once (the counts stay on the device: no host round trip)
template <typename Launch> int slow_rounds(bmq2d_solver *s, int w, Launch launch)
{
    for (int r = 1; r <= 6; ++r) {
        WorkList in = worklist(s, w, (r - 1) & 1), out = worklist(s, w, r & 1);
        BMQ_CK(cudaMemsetAsync(out.count, 0, sizeof(int), s->side[w]));
        launch(r, in, out);
        L2D(s);
        BMQ_CK(cudaGetLastError());
    }
    return BMQ_OK;
}

// The solveODE-based kernels run in two passes (see solve_ode_quick).  The quick passes go on the
// solver's stream; the six compacted slow passes (2 forward maps, 4 semi-Lagrangian fields) are
// independent of each other and of the DMC backward update, and each is one long dependent chain
// of up to 254 RK3 steps, so they are launched side by side on six side streams and joined before
// the first consumer.
int forward_quick(bmq2d_solver *s, float dt, int fx, int fy, int w)
{
    WorkList wl = worklist(s, w);
    BMQ_CK(cudaMemsetAsync(wl.count, 0, sizeof(int), s->stream));
    k2_forward<<<grd(s->ni, s->nj), blk(), 0, s->stream>>>(s->g, s->f[BMQ2_F_U], s->f[BMQ2_F_V], s->f[fx], s->f[fy], dt, wl);
    L2D(s);
    BMQ_CK(cudaGetLastError());
    return BMQ_OK;
}
int forward_slow(bmq2d_solver *s, float dt, int fx, int fy, int w)
{
    BMQ_CK(cudaStreamWaitEvent(s->side[w], s->ev_quick, 0));
    {
        int st = slow_rounds(s, w, [&](int r, WorkList in, WorkList out) {
            k2_forward_round<<<148 * 4, 128, 0, s->side[w]>>>(s->g, s->f[BMQ2_F_U], s->f[BMQ2_F_V], s->f[fx], s->f[fy], dt, r, in, out, s->d_round_counts + 8 * w);
        });
        if (st != BMQ_OK) return st;
    }
    BMQ_CK(cudaEventRecord(s->ev_side[w], s->side[w]));
    s->pending_slow |= 1 << w;
    return BMQ_OK;
}
int semilag_quick(bmq2d_solver *s, int src, int dst, float dt, int w)
{
    FieldGeom q = geom(s, kind_of(src));
    WorkList wl = worklist(s, w);
    BMQ_CK(cudaMemsetAsync(wl.count, 0, sizeof(int), s->stream));
    k2_semilag<<<grd(q.fni, q.fnj), blk(), 0, s->stream>>>(s->g, s->f[BMQ2_F_U], s->f[BMQ2_F_V], s->f[src], s->f[dst],
                                                               q.fni, q.fnj, q.offx, q.offy, dt, wl);
    L2D(s);
    BMQ_CK(cudaGetLastError());
    return BMQ_OK;
}
int semilag_slow(bmq2d_solver *s, int src, int dst, float dt, int w)
{
    FieldGeom q = geom(s, kind_of(src));
    BMQ_CK(cudaStreamWaitEvent(s->side[w], s->ev_quick, 0));
    {
        int st = slow_rounds(s, w, [&](int r, WorkList in, WorkList out) {
            k2_semilag_round<<<148 * 4, 128, 0, s->side[w]>>>(s->g, s->f[BMQ2_F_U], s->f[BMQ2_F_V], s->f[src], s->f[dst], q.fni, q.fnj,
                                                             q.offx, q.offy, dt, r, in, out, s->d_round_counts + 8 * w);
        });
        if (st != BMQ_OK) return st;
    }
    BMQ_CK(cudaEventRecord(s->ev_side[w], s->side[w]));
    s->pending_slow |= 1 << w;
    return BMQ_OK;
}
int join_slow(bmq2d_solver *s)
{
    for (int w = 0; w < 6; ++w)
        if (s->pending_slow & (1 << w)) BMQ_CK(cudaStreamWaitEvent(s->stream, s->ev_side[w], 0));
    s->pending_slow = 0;
    return BMQ_OK;
}

// updateBackward, :1242-1259
int update_backward(bmq2d_solver *s, float dt, int bx, int by)
{
    const float *u = s->f[BMQ2_F_U], *v = s->f[BMQ2_F_V];
    float substep = s->cfl, T = dt, t = 0.f;
    int n = 0;
    while (t < T) {
        if (t + substep > T) substep = T - t;
        k2_backward<<<grd(s->ni, s->nj), blk(), 0, s->stream>>>(s->g, u, v, s->f[bx], s->f[by], s->f[BMQ2_F_MAP_TMPX],
                                                              s->f[BMQ2_F_MAP_TMPY], substep);
        L2D(s);
        BMQ_CK(cudaGetLastError());
        std::swap(s->f[bx], s->f[BMQ2_F_MAP_TMPX]);
        std::swap(s->f[by], s->f[BMQ2_F_MAP_TMPY]);
        t += substep;
        if (++n > 4096) return bmq::set_error(BMQ_ERR_ARG, "bmq2d: more than 4096 CFL sub-steps");
    }
    s->stats.n_substeps = n;
    return BMQ_OK;
}

void set_guard(Kind k, const bmq2d_solver *s, bool advect, int &ilo, int &ihi, int &jlo, int &jhi)
{
    const int ni = s->ni, nj = s->nj;
    if (advect) {
        // advectVelocity u: i>1&&i<ni-1&&j>1&&j<nj-2 ; v: j>1&&j<nj-1&&i>1&&i<ni-2 ; scalars: j>1&&j<nj-1&&i>0&&i<ni-1
        if (k == KU) { ilo = 1; ihi = ni - 1; jlo = 1; jhi = nj - 2; }
        else if (k == KV) { ilo = 1; ihi = ni - 2; jlo = 1; jhi = nj - 1; }
        else { ilo = 0; ihi = ni - 1; jlo = 1; jhi = nj - 1; }
    } else {
        // correct*/accumulate*: u and scalars: i>1&&i<ni-1&&j>0&&j<nj-1 ; v: j>1&&j<nj-1&&i>0&&i<ni-1
        if (k == KV) { ilo = 0; ihi = ni - 1; jlo = 1; jhi = nj - 1; }
        else { ilo = 1; ihi = ni - 1; jlo = 0; jhi = nj - 1; }
    }
}

int advect_one(bmq2d_solver *s, int cur, int init, int orig, int d, int dprev, int semi, bool scalar)
{
    Kind k = kind_of(cur);
    FieldGeom q = geom(s, k);
    AdvectArgs a;
    a.bx = s->f[scalar ? BMQ2_F_SBWD_X : BMQ2_F_BWD_X]; a.by = s->f[scalar ? BMQ2_F_SBWD_Y : BMQ2_F_BWD_Y];
    a.bxp = s->f[scalar ? BMQ2_F_SBWDP_X : BMQ2_F_BWDP_X]; a.byp = s->f[scalar ? BMQ2_F_SBWDP_Y : BMQ2_F_BWDP_Y];
    a.f_init = s->f[init]; a.f_orig = s->f[orig]; a.d = s->f[d]; a.d_prev = s->f[dprev]; a.semi = s->f[semi];
    a.out = s->f[cur];
    a.fni = q.fni; a.fnj = q.fnj; a.offx = q.offx; a.offy = q.offy; a.blend = s->blend;
    set_guard(k, s, true, a.i_lo, a.i_hi, a.j_lo, a.j_hi);
    k2_advect<<<grd(q.fni, q.fnj), blk(), 0, s->stream>>>(s->g, a);
    L2D(s);
    BMQ_CK(cudaGetLastError());
    return BMQ_OK;
}

// correctVelocity / correctScalars for one field: error via psi, apply via chi, clampExtrema2
int correct_one(bmq2d_solver *s, int cur, int init, int d, int scratch_curr, int scratch_temp, bool scalar)
{
    Kind k = kind_of(cur);
    FieldGeom q = geom(s, k);
    int st = copyf(s, scratch_curr, cur);       // u_curr = u
    if (st != BMQ_OK) return st;
    CorrectArgs a;
    a.fni = q.fni; a.fnj = q.fnj; a.offx = q.offx; a.offy = q.offy;
    set_guard(k, s, false, a.i_lo, a.i_hi, a.j_lo, a.j_hi);
    a.mx = s->f[scalar ? BMQ2_F_SFWD_X : BMQ2_F_FWD_X]; a.my = s->f[scalar ? BMQ2_F_SFWD_Y : BMQ2_F_FWD_Y];
    a.src = s->f[cur]; a.d = s->f[d]; a.f_init = s->f[init]; a.out = s->f[scratch_temp];
    k2_correct_error<<<grd(q.fni, q.fnj), blk(), 0, s->stream>>>(s->g, a);
    L2D(s);
    BMQ_CK(cudaGetLastError());
    a.mx = s->f[scalar ? BMQ2_F_SBWD_X : BMQ2_F_BWD_X]; a.my = s->f[scalar ? BMQ2_F_SBWD_Y : BMQ2_F_BWD_Y];
    a.src = s->f[scratch_temp]; a.out = s->f[cur];
    k2_correct_apply<<<grd(q.fni, q.fnj), blk(), 0, s->stream>>>(s->g, a);
    L2D(s);
    BMQ_CK(cudaGetLastError());
    k2_clamp_extrema<<<grd(q.fni, q.fnj), blk(), 0, s->stream>>>(q.fni, q.fnj, s->f[scratch_curr], s->f[cur]);
    L2D(s);
    BMQ_CK(cudaGetLastError());
    return BMQ_OK;
}

int accumulate_one(bmq2d_solver *s, int d, int nch, const int *change, const float *coeff, bool scalar)
{
    Kind k = kind_of(d);
    FieldGeom q = geom(s, k);
    AccumArgs a;
    a.mx = s->f[scalar ? BMQ2_F_SFWD_X : BMQ2_F_FWD_X]; a.my = s->f[scalar ? BMQ2_F_SFWD_Y : BMQ2_F_FWD_Y];
    a.nch = nch;
    for (int c = 0; c < 2; ++c) { a.change[c] = s->f[change[c < nch ? c : 0]]; a.coeff[c] = coeff[c < nch ? c : 0]; }
    a.d = s->f[d];
    a.fni = q.fni; a.fnj = q.fnj; a.offx = q.offx; a.offy = q.offy;
    a.scalar_form = scalar ? 1 : 0;
    set_guard(k, s, false, a.i_lo, a.i_hi, a.j_lo, a.j_hi);
    k2_accumulate<<<grd(q.fni, q.fnj), blk(), 0, s->stream>>>(s->g, a);
    L2D(s);
    BMQ_CK(cudaGetLastError());
    return BMQ_OK;
}

int sub_fields(bmq2d_solver *s, int out, int a, int b)
{
    const int n = s->fni[out] * s->fnj[out];
    k2_sub<<<(n + 255) / 256, 256, 0, s->stream>>>(s->f[out], s->f[a], s->f[b], n);
    L2D(s);
    BMQ_CK(cudaGetLastError());
    return BMQ_OK;
}

#define R2(x) do { int _s = (x); if (_s != BMQ_OK) return _s; } while (0)
#define NEED2(s) do { if (!(s)) return bmq::set_error(BMQ_ERR_ARG, "%s: null solver handle", __func__); } while (0)

}  // namespace

extern "C" {

int bmq2d_create(int ni, int nj, float h, float blend_coeff, bmq2d_solver **out)
{
    if (!out) return bmq::set_error(BMQ_ERR_ARG, "bmq2d_create: out is null");
    *out = nullptr;
    if (!bmq::require_device()) return BMQ_ERR_NODEVICE;
    if (ni < 8 || nj < 8 || !(h > 0.f)) return bmq::set_error(BMQ_ERR_ARG, "bmq2d_create: bad grid %dx%d h=%g", ni, nj, (double)h);
    bmq2d_solver *s = new (std::nothrow) bmq2d_solver();
    if (!s) return bmq::set_error(BMQ_ERR_ARG, "bmq2d_create: out of host memory");
    s->ni = ni; s->nj = nj; s->h = h; s->blend = blend_coeff;
    s->g = G2{ni, nj, h, 1.0f / h, 0};
    {
        int e;
        const int nmax = ni > nj ? ni : nj;
        if (frexpf(h, &e) == 0.5f) s->g.div_mode = 1;                                                  // power of two: exact
        else if (bmq::division_verified(h, 4.0f * (float)(nmax + 8) * h)) s->g.div_mode = 2;          // checked exhaustively
    }
    memset(&s->stats, 0, sizeof s->stats);
    int st = BMQ_OK;
    for (int id = 0; id < BMQ2_F_COUNT && st == BMQ_OK; ++id) {
        FieldGeom q = geom(s, kind_of(id));
        s->fni[id] = q.fni; s->fnj[id] = q.fnj;
        const size_t bytes = sizeof(float) * q.fni * q.fnj;
        st = bmq::check_cuda(cudaMalloc(&s->f[id], bytes), "cudaMalloc", __FILE__, __LINE__);
        if (st == BMQ_OK) st = bmq::check_cuda(cudaMemset(s->f[id], 0, bytes), "cudaMemset", __FILE__, __LINE__);
    }
    if (st == BMQ_OK) st = bmq::check_cuda(cudaMalloc(&s->d_red, 4 * sizeof(float)), "cudaMalloc", __FILE__, __LINE__);
    s->wl_cap = (size_t)(ni + 1) * (nj + 1);
    s->wl_stride = 4 + 3 * s->wl_cap;
    if (st == BMQ_OK) st = bmq::check_cuda(cudaMalloc(&s->d_wl, sizeof(int) * 12 * s->wl_stride), "cudaMalloc", __FILE__, __LINE__);
    if (st == BMQ_OK) st = bmq::check_cuda(cudaMalloc(&s->d_round_counts, sizeof(int) * 48), "cudaMalloc", __FILE__, __LINE__);
    if (st == BMQ_OK) st = bmq::check_cuda(cudaMemset(s->d_round_counts, 0, sizeof(int) * 48), "cudaMemset", __FILE__, __LINE__);
    for (int w = 0; w < 6 && st == BMQ_OK; ++w) {
        st = bmq::check_cuda(cudaStreamCreateWithFlags(&s->side[w], cudaStreamNonBlocking), "cudaStreamCreate", __FILE__, __LINE__);
        if (st == BMQ_OK) st = bmq::check_cuda(cudaEventCreateWithFlags(&s->ev_side[w], cudaEventDisableTiming), "cudaEventCreate", __FILE__, __LINE__);
    }
    if (st == BMQ_OK) st = bmq::check_cuda(cudaEventCreateWithFlags(&s->ev_quick, cudaEventDisableTiming), "cudaEventCreate", __FILE__, __LINE__);
    if (st == BMQ_OK) st = bmq::check_cuda(cudaMallocHost(&s->h_red, 4 * sizeof(float)), "cudaMallocHost", __FILE__, __LINE__);
    if (st == BMQ_OK) st = bmq2d_reset(s);
    if (st != BMQ_OK) { bmq2d_destroy(s); return st; }
    *out = s;
    return BMQ_OK;
}

int bmq2d_destroy(bmq2d_solver *s)
{
    if (!s) return BMQ_OK;
    for (float *p : s->f) if (p) cudaFree(p);
    if (s->d_red) cudaFree(s->d_red);
    if (s->d_wl) cudaFree(s->d_wl);
    if (s->d_round_counts) cudaFree(s->d_round_counts);
    for (int w = 0; w < 6; ++w) {
        if (s->side[w]) cudaStreamDestroy(s->side[w]);
        if (s->ev_side[w]) cudaEventDestroy(s->ev_side[w]);
    }
    if (s->ev_quick) cudaEventDestroy(s->ev_quick);
    if (s->h_red) cudaFreeHost(s->h_red);
    delete s;
    return BMQ_OK;
}

// constructor state of BimocqSolver2D (:156-270): identity maps, zero change buffers, counters
int bmq2d_reset(bmq2d_solver *s)
{
    NEED2(s);
    R2(ident(s, BMQ2_F_FWD_X, BMQ2_F_FWD_Y, BMQ2_F_BWD_X, BMQ2_F_BWD_Y));
    R2(ident(s, BMQ2_F_SFWD_X, BMQ2_F_SFWD_Y, BMQ2_F_SBWD_X, BMQ2_F_SBWD_Y));
    R2(ident(s, BMQ2_F_BWDP_X, BMQ2_F_BWDP_Y, BMQ2_F_SBWDP_X, BMQ2_F_SBWDP_Y));
    const int zero[] = {BMQ2_F_DU, BMQ2_F_DV, BMQ2_F_DRHO, BMQ2_F_DT, BMQ2_F_DU_PREV, BMQ2_F_DV_PREV, BMQ2_F_DRHO_PREV,
                        BMQ2_F_DT_PREV, BMQ2_F_U_ORIG, BMQ2_F_V_ORIG, BMQ2_F_RHO_ORIG, BMQ2_F_T_ORIG};
    for (int id : zero) R2(zerof(s, id));
    s->lastremeshing = s->rho_lastremeshing = 0;
    s->total_resample = s->total_scalar_resample = 0;
    BMQ_CK(cudaStreamSynchronize(s->stream));
    return BMQ_OK;
}

int bmq2d_set_levelset(bmq2d_solver *s, int on) { NEED2(s); s->levelset = on != 0; return BMQ_OK; }

int bmq2d_set_counters(bmq2d_solver *s, int lastremeshing, int rho_lastremeshing)
{
    NEED2(s);
    s->lastremeshing = lastremeshing;
    s->rho_lastremeshing = rho_lastremeshing;
    return BMQ_OK;
}

int bmq2d_field_ptr(bmq2d_solver *s, int field_id, float **dev_ptr, int *fni, int *fnj)
{
    NEED2(s);
    if (field_id < 0 || field_id >= BMQ2_F_COUNT) return bmq::set_error(BMQ_ERR_ARG, "bmq2d_field_ptr: bad field id %d", field_id);
    if (dev_ptr) *dev_ptr = s->f[field_id];
    if (fni) *fni = s->fni[field_id];
    if (fnj) *fnj = s->fnj[field_id];
    return BMQ_OK;
}

int bmq2d_upload(bmq2d_solver *s, int field_id, const float *host)
{
    NEED2(s);
    if (field_id < 0 || field_id >= BMQ2_F_COUNT || !host) return bmq::set_error(BMQ_ERR_ARG, "bmq2d_upload: bad argument");
    BMQ_CK(cudaMemcpyAsync(s->f[field_id], host, sizeof(float) * s->fni[field_id] * s->fnj[field_id], cudaMemcpyHostToDevice, s->stream));
    BMQ_CK(cudaStreamSynchronize(s->stream));
    return BMQ_OK;
}

int bmq2d_download(bmq2d_solver *s, int field_id, float *host)
{
    NEED2(s);
    if (field_id < 0 || field_id >= BMQ2_F_COUNT || !host) return bmq::set_error(BMQ_ERR_ARG, "bmq2d_download: bad argument");
    BMQ_CK(cudaMemcpyAsync(host, s->f[field_id], sizeof(float) * s->fni[field_id] * s->fnj[field_id], cudaMemcpyDeviceToHost, s->stream));
    BMQ_CK(cudaStreamSynchronize(s->stream));
    return BMQ_OK;
}

// advanceBIMOCQ lines 394-445
int bmq2d_advect(bmq2d_solver *s, int frame, float dt)
{
    NEED2(s);
    if (!(dt > 0.f)) return bmq::set_error(BMQ_ERR_ARG, "bmq2d_advect: dt must be positive");
    // getCFL() runs BEFORE the restore of u_temp in the reference (:395-400): it sees the
    // time-averaged velocity the previous step left in u, v
    {
        float m = 0.f;
        R2(max_vel(s, &m));
        s->stats.max_vel_pre = m;
    }
    if (frame != 0 && !s->levelset) {          // restore the un-averaged velocity, :396-400
        R2(copyf(s, BMQ2_F_U, BMQ2_F_U_TEMP));
        R2(copyf(s, BMQ2_F_V, BMQ2_F_V_TEMP));
    }
    s->cfl = s->h / fabsf(s->stats.max_vel_pre);
    s->stats.cfl = s->cfl;
    // quick passes of every solveODE-based kernel, then their slow passes side by side, overlapped
    // with the DMC backward updates (which depend on none of them)
    if (!s->levelset) R2(forward_quick(s, dt, BMQ2_F_FWD_X, BMQ2_F_FWD_Y, 0));
    R2(forward_quick(s, dt, BMQ2_F_SFWD_X, BMQ2_F_SFWD_Y, 1));
    R2(semilag_quick(s, BMQ2_F_RHO, BMQ2_F_RHO_SEMI, dt, 2));
    R2(semilag_quick(s, BMQ2_F_T, BMQ2_F_T_SEMI, dt, 3));
    R2(semilag_quick(s, BMQ2_F_U, BMQ2_F_U_SEMI, dt, 4));
    R2(semilag_quick(s, BMQ2_F_V, BMQ2_F_V_SEMI, dt, 5));
    BMQ_CK(cudaEventRecord(s->ev_quick, s->stream));
    if (!s->levelset) R2(forward_slow(s, dt, BMQ2_F_FWD_X, BMQ2_F_FWD_Y, 0));
    R2(forward_slow(s, dt, BMQ2_F_SFWD_X, BMQ2_F_SFWD_Y, 1));
    R2(semilag_slow(s, BMQ2_F_RHO, BMQ2_F_RHO_SEMI, dt, 2));
    R2(semilag_slow(s, BMQ2_F_T, BMQ2_F_T_SEMI, dt, 3));
    R2(semilag_slow(s, BMQ2_F_U, BMQ2_F_U_SEMI, dt, 4));
    R2(semilag_slow(s, BMQ2_F_V, BMQ2_F_V_SEMI, dt, 5));
    if (!s->levelset) R2(update_backward(s, dt, BMQ2_F_BWD_X, BMQ2_F_BWD_Y));
    R2(update_backward(s, dt, BMQ2_F_SBWD_X, BMQ2_F_SBWD_Y));
    R2(join_slow(s));
    R2(copyf(s, BMQ2_F_U_PRESAVE, BMQ2_F_U));
    R2(copyf(s, BMQ2_F_V_PRESAVE, BMQ2_F_V));
    if (!s->levelset) {
        // advectVelocity reads u_init/u_origin/du/du_prev only, so u and v can be written in place
        R2(advect_one(s, BMQ2_F_U, BMQ2_F_U_INIT, BMQ2_F_U_ORIG, BMQ2_F_DU, BMQ2_F_DU_PREV, BMQ2_F_U_SEMI, false));
        R2(advect_one(s, BMQ2_F_V, BMQ2_F_V_INIT, BMQ2_F_V_ORIG, BMQ2_F_DV, BMQ2_F_DV_PREV, BMQ2_F_V_SEMI, false));
        R2(correct_one(s, BMQ2_F_U, BMQ2_F_U_INIT, BMQ2_F_DU, BMQ2_F_U_SCRATCH, BMQ2_F_U_SCRATCH2, false));
        R2(correct_one(s, BMQ2_F_V, BMQ2_F_V_INIT, BMQ2_F_DV, BMQ2_F_V_SCRATCH, BMQ2_F_V_SCRATCH2, false));
    }
    R2(advect_one(s, BMQ2_F_RHO, BMQ2_F_RHO_INIT, BMQ2_F_RHO_ORIG, BMQ2_F_DRHO, BMQ2_F_DRHO_PREV, BMQ2_F_RHO_SEMI, true));
    R2(advect_one(s, BMQ2_F_T, BMQ2_F_T_INIT, BMQ2_F_T_ORIG, BMQ2_F_DT, BMQ2_F_DT_PREV, BMQ2_F_T_SEMI, true));
    if (!s->levelset) {
        R2(correct_one(s, BMQ2_F_RHO, BMQ2_F_RHO_INIT, BMQ2_F_DRHO, BMQ2_F_C_SCRATCH, BMQ2_F_C_SCRATCH2, true));
        R2(correct_one(s, BMQ2_F_T, BMQ2_F_T_INIT, BMQ2_F_DT, BMQ2_F_C_SCRATCH, BMQ2_F_C_SCRATCH2, true));
    }
    // u_save.. = u.. (:442-445)
    R2(copyf(s, BMQ2_F_U_SAVE, BMQ2_F_U));
    R2(copyf(s, BMQ2_F_V_SAVE, BMQ2_F_V));
    R2(copyf(s, BMQ2_F_RHO_SAVE, BMQ2_F_RHO));
    R2(copyf(s, BMQ2_F_T_SAVE, BMQ2_F_T));
    BMQ_CK(cudaStreamSynchronize(s->stream));
    return BMQ_OK;
}

// advanceBIMOCQ lines 449-507.  On entry U_FORCED/V_FORCED hold the velocity after the external
// forces (:447-448) and U, V, RHO, T the fields after the projection (:454).
int bmq2d_accumulate(bmq2d_solver *s, int frame, float dt)
{
    NEED2(s);
    float proj_coeff = 2.0f;
    R2(sub_fields(s, BMQ2_F_DU_EXT, BMQ2_F_U_FORCED, BMQ2_F_U_SAVE));     // du_temp = u - u_save
    R2(sub_fields(s, BMQ2_F_DV_EXT, BMQ2_F_V_FORCED, BMQ2_F_V_SAVE));
    BMQ_CK(cudaMemsetAsync(s->d_red, 0, 4 * sizeof(float), s->stream));
    k2_distortion<<<grd(s->ni, s->nj), blk(), 0, s->stream>>>(s->g, s->f[BMQ2_F_BWD_X], s->f[BMQ2_F_BWD_Y], s->f[BMQ2_F_FWD_X],
                                                            s->f[BMQ2_F_FWD_Y], s->f[BMQ2_F_SBWD_X], s->f[BMQ2_F_SBWD_Y],
                                                            s->f[BMQ2_F_SFWD_X], s->f[BMQ2_F_SFWD_Y], s->d_red + 1);
    L2D(s);
    BMQ_CK(cudaGetLastError());
    BMQ_CK(cudaMemcpyAsync(s->h_red, s->d_red, 4 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    float vel = 0.f;
    R2(max_vel(s, &vel));    // synchronises
    const float d_vel = s->h_red[1], d_scalar = s->h_red[2];
    const float cond_v = d_vel / (vel * dt), cond_s = d_scalar / (vel * dt);
    const bool vel_remap = (double)cond_v > 1.0 || frame - s->lastremeshing >= 8;
    const bool rho_remap = (double)cond_s > 1.0 || frame - s->rho_lastremeshing >= 20;
    if (vel_remap) proj_coeff = 1.0f;
    s->stats.max_vel = vel; s->stats.vel_condition = cond_v; s->stats.scalar_condition = cond_s;
    s->stats.vel_remap = 0; s->stats.scalar_remap = 0;
    if (!s->levelset) {
        R2(sub_fields(s, BMQ2_F_DU_PROJ, BMQ2_F_U, BMQ2_F_U_FORCED));
        R2(sub_fields(s, BMQ2_F_DV_PROJ, BMQ2_F_V, BMQ2_F_V_FORCED));
        R2(sub_fields(s, BMQ2_F_DRHO_EXT, BMQ2_F_RHO, BMQ2_F_RHO_SAVE));
        R2(sub_fields(s, BMQ2_F_DT_EXT, BMQ2_F_T, BMQ2_F_T_SAVE));
        const float coeff[2] = {1.0f, proj_coeff};
        const int cu[2] = {BMQ2_F_DU_EXT, BMQ2_F_DU_PROJ}, cv[2] = {BMQ2_F_DV_EXT, BMQ2_F_DV_PROJ};
        // accumulateVelocity(du_temp,dv_temp,1) then accumulateVelocity(du_proj,dv_proj,proj_coeff):
        // per component the two calls are consecutive read-modify-writes of du -> one pass, same order
        R2(accumulate_one(s, BMQ2_F_DU, 2, cu, coeff, false));
        R2(accumulate_one(s, BMQ2_F_DV, 2, cv, coeff, false));
        const int cr[1] = {BMQ2_F_DRHO_EXT}, ct[1] = {BMQ2_F_DT_EXT};
        R2(accumulate_one(s, BMQ2_F_DRHO, 1, cr, coeff, true));
        R2(accumulate_one(s, BMQ2_F_DT, 1, ct, coeff, true));
    }
    if (vel_remap && !s->levelset) {          // resampleVelBuffer, :1426-1449
        s->lastremeshing = frame;
        s->total_resample++;
        std::swap(s->f[BMQ2_F_U_ORIG], s->f[BMQ2_F_U_INIT]);      // u_origin = u_init
        std::swap(s->f[BMQ2_F_V_ORIG], s->f[BMQ2_F_V_INIT]);
        R2(copyf(s, BMQ2_F_U_INIT, BMQ2_F_U));                    // u_init = u
        R2(copyf(s, BMQ2_F_V_INIT, BMQ2_F_V));
        std::swap(s->f[BMQ2_F_DU_PREV], s->f[BMQ2_F_DU]);         // du_prev = du; du = 0
        std::swap(s->f[BMQ2_F_DV_PREV], s->f[BMQ2_F_DV]);
        R2(zerof(s, BMQ2_F_DU));
        R2(zerof(s, BMQ2_F_DV));
        std::swap(s->f[BMQ2_F_BWDP_X], s->f[BMQ2_F_BWD_X]);       // backward_prev = backward
        std::swap(s->f[BMQ2_F_BWDP_Y], s->f[BMQ2_F_BWD_Y]);
        R2(ident(s, BMQ2_F_FWD_X, BMQ2_F_FWD_Y, BMQ2_F_BWD_X, BMQ2_F_BWD_Y));
        const float c1[2] = {proj_coeff, 0.f};
        const int cu[2] = {BMQ2_F_DU_PROJ, 0}, cv[2] = {BMQ2_F_DV_PROJ, 0};
        R2(accumulate_one(s, BMQ2_F_DU, 1, cu, c1, false));       // :485
        R2(accumulate_one(s, BMQ2_F_DV, 1, cv, c1, false));
        s->stats.vel_remap = 1;
    }
    if (rho_remap) {                           // resampleRhoBuffer, :1451-1474
        s->rho_lastremeshing = frame;
        s->total_scalar_resample++;
        std::swap(s->f[BMQ2_F_RHO_ORIG], s->f[BMQ2_F_RHO_INIT]);
        std::swap(s->f[BMQ2_F_T_ORIG], s->f[BMQ2_F_T_INIT]);
        R2(copyf(s, BMQ2_F_RHO_INIT, BMQ2_F_RHO));
        R2(copyf(s, BMQ2_F_T_INIT, BMQ2_F_T));
        std::swap(s->f[BMQ2_F_DRHO_PREV], s->f[BMQ2_F_DRHO]);
        std::swap(s->f[BMQ2_F_DT_PREV], s->f[BMQ2_F_DT]);
        R2(zerof(s, BMQ2_F_DRHO));
        R2(zerof(s, BMQ2_F_DT));
        std::swap(s->f[BMQ2_F_SBWDP_X], s->f[BMQ2_F_SBWD_X]);
        std::swap(s->f[BMQ2_F_SBWDP_Y], s->f[BMQ2_F_SBWD_Y]);
        R2(ident(s, BMQ2_F_SFWD_X, BMQ2_F_SFWD_Y, BMQ2_F_SBWD_X, BMQ2_F_SBWD_Y));
        s->stats.scalar_remap = 1;
    }
    R2(copyf(s, BMQ2_F_U_TEMP, BMQ2_F_U));     // u_temp = u, :493-494
    R2(copyf(s, BMQ2_F_V_TEMP, BMQ2_F_V));
    if (frame != 0) {
        int n = (s->ni + 1) * s->nj;
        k2_average<<<(n + 255) / 256, 256, 0, s->stream>>>(s->f[BMQ2_F_U], s->f[BMQ2_F_U_PRESAVE], n);
        n = s->ni * (s->nj + 1);
        k2_average<<<(n + 255) / 256, 256, 0, s->stream>>>(s->f[BMQ2_F_V], s->f[BMQ2_F_V_PRESAVE], n);
        s->launches += 2;
        BMQ_CK(cudaGetLastError());
    }
    BMQ_CK(cudaStreamSynchronize(s->stream));
    s->stats.last_remesh = s->lastremeshing; s->stats.last_scalar_remesh = s->rho_lastremeshing;
    s->stats.total_remesh = s->total_resample; s->stats.total_scalar_remesh = s->total_scalar_resample;
    return BMQ_OK;
}

// Cells of the last step whose solveODE needed more than the first round, per work list: forward maps (velocity,
// scalar), semi-Lagrangian rho, T, u, v.  Synchronises; a diagnostic, not part of the step.
int bmq2d_deferred_counts(bmq2d_solver *s, int *counts6)
{
    NEED2(s);
    if (!counts6) return bmq::set_error(BMQ_ERR_ARG, "bmq2d_deferred_counts: counts6 is null");
    int all[48];
    BMQ_CK(cudaStreamSynchronize(s->stream));
    BMQ_CK(cudaMemcpy(all, s->d_round_counts, sizeof all, cudaMemcpyDeviceToHost));
    for (int w = 0; w < 6; ++w) counts6[w] = all[8 * w + 1];
    return BMQ_OK;
}
// the same per round: counts36[6 * w + (r - 1)] = cells of list w that entered round r = 1..6
int bmq2d_deferred_round_counts(bmq2d_solver *s, int *counts36)
{
    NEED2(s);
    if (!counts36) return bmq::set_error(BMQ_ERR_ARG, "bmq2d_deferred_round_counts: counts36 is null");
    int all[48];
    BMQ_CK(cudaStreamSynchronize(s->stream));
    BMQ_CK(cudaMemcpy(all, s->d_round_counts, sizeof all, cudaMemcpyDeviceToHost));
    for (int w = 0; w < 6; ++w)
        for (int r = 1; r <= 6; ++r) counts36[6 * w + r - 1] = all[8 * w + r];
    return BMQ_OK;
}

int bmq2d_get_stats(bmq2d_solver *s, bmq2d_stats *out)
{
    NEED2(s);
    if (!out) return bmq::set_error(BMQ_ERR_ARG, "bmq2d_get_stats: out is null");
    *out = s->stats;
    return BMQ_OK;
}

// host-buffer step: the reference's fields live on the host (Array2f)
int bmq2d_advect_host(bmq2d_solver *s, int frame, float dt, float *u, float *v, float *rho, float *T)
{
    NEED2(s);
    if (!u || !v || !rho || !T) return bmq::set_error(BMQ_ERR_ARG, "bmq2d_advect_host: null field");
    // the host's u,v are the time-averaged velocity (used for getCFL); for frame > 0 the solver
    // then restores its own un-averaged copy (u_temp), exactly like :395-400
    R2(bmq2d_upload(s, BMQ2_F_U, u));
    R2(bmq2d_upload(s, BMQ2_F_V, v));
    R2(bmq2d_upload(s, BMQ2_F_RHO, rho));
    R2(bmq2d_upload(s, BMQ2_F_T, T));
    R2(bmq2d_advect(s, frame, dt));
    R2(bmq2d_download(s, BMQ2_F_U, u));
    R2(bmq2d_download(s, BMQ2_F_V, v));
    R2(bmq2d_download(s, BMQ2_F_RHO, rho));
    R2(bmq2d_download(s, BMQ2_F_T, T));
    return BMQ_OK;
}

int bmq2d_accumulate_host(bmq2d_solver *s, int frame, float dt, const float *u_forced, const float *v_forced, float *u_final,
                          float *v_final, const float *rho_final, const float *T_final)
{
    NEED2(s);
    if (!u_forced || !v_forced || !u_final || !v_final || !rho_final || !T_final)
        return bmq::set_error(BMQ_ERR_ARG, "bmq2d_accumulate_host: null field");
    R2(bmq2d_upload(s, BMQ2_F_U_FORCED, u_forced));
    R2(bmq2d_upload(s, BMQ2_F_V_FORCED, v_forced));
    R2(bmq2d_upload(s, BMQ2_F_U, u_final));
    R2(bmq2d_upload(s, BMQ2_F_V, v_final));
    R2(bmq2d_upload(s, BMQ2_F_RHO, rho_final));
    R2(bmq2d_upload(s, BMQ2_F_T, T_final));
    R2(bmq2d_accumulate(s, frame, dt));
    // the time-averaged velocity the next step's projection-side code sees (:497-506)
    R2(bmq2d_download(s, BMQ2_F_U, u_final));
    R2(bmq2d_download(s, BMQ2_F_V, v_final));
    return BMQ_OK;
}

unsigned long long bmq2d_kernel_launch_count(bmq2d_solver *s) { return s ? s->launches : 0; }

}  // extern "C"
