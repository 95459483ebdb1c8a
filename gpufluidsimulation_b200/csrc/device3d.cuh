// Device-side building blocks of the 3D BiMocq^2 advection kernels (sm_100a).
//
// Numerics contract (DESIGN.md "Numerics"): fp32 storage and fp32 arithmetic with the
// interpolation nest in the reference's order x -> y -> z (GPU_kernel.cu:27-41).  The reference
// lerp (GPU_kernel.cu:22-25) evaluates (1.0-c)*a in double and c*b in float; here it is
// fmaf(1-c, a, c*b): the same float product c*b, one fused rounding of the sum, no FP64 and no
// F2F conversions.  Positions -> (cell, fraction) use the reference's expression
// floorf(p/h), p/h - i (GPU_kernel.cu:46-51) with an IEEE division; when h is a power of two
// the division is replaced by an exact multiplication (template parameter P2), which is
// bit-identical.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bmq {

struct Grid3 {
    int ni, nj, nk;   // global cell counts
    float h;          // cell size
    float inv_h;      // RN(1/h) (exact when h is a power of two); NEGATED when the three-instruction division
                      // div_h() has not been verified for this h (kernels3d.cu:make_grid) -> IEEE division
    int p2;           // host-side dispatch: take the kernels written for a power-of-two h (h is one, or the grid was
                      // made in tolerance mode, bmq_set_tolerance_mode)
};

// p / h, correctly rounded, for the one divisor a launch ever has.  The reference divides (pos / h,
// GPU_kernel.cu:46-51); an IEEE division is ~10 instructions plus a slow-path call.  With y = RN(1/h):
//   q0 = RN(p y);  r = p - h q0 (exact, one fma);  q = RN(q0 + r y)
// is the correctly rounded quotient (Markstein's division step) for all but pathological divisors, as long as
// the residual does not underflow; instead of relying on the theorem, make_grid() checks the sequence against
// __fdiv_rn for EVERY float p in {0} U [2^-100, largest position], once per h, on the device, and hands the
// kernels a negated inv_h if a single one differs.  |p| below 2^-100 (never a position in practice; the
// residual would be subnormal there) takes the IEEE division.
#define BMQ_DIV_TINY 7.888609052e-31f   /* 2^-100 */
__device__ __forceinline__ float div_h(float p, float h, float inv_h)
{
    if (inv_h > 0.f && (fabsf(p) >= BMQ_DIV_TINY || p == 0.f)) {
        const float q0 = __fmul_rn(p, inv_h);
        const float r = __fmaf_rn(-h, q0, p);
        return __fmaf_rn(r, inv_h, q0);
    }
    return __fdiv_rn(p, h);
}

// FIX != 0: the x and y extents are the compile-time constant FIX, so that row and plane pitches of
// every gather become immediates of the load instructions instead of 64-bit address arithmetic.
template <int FIX> __device__ __forceinline__ Grid3 fix_grid(const Grid3 &g)
{
    Grid3 r = g;
    if (FIX) { r.ni = FIX; r.nj = FIX; }
    return r;
}

struct Frac {
    int i;
    float f, omf;
};

template <bool P2>
__device__ __forceinline__ Frac split(float p, float h, float inv_h)
{
    float q = P2 ? p * inv_h : div_h(p, h, inv_h);
    float fl = floorf(q);
    Frac r;
    r.i = (int)fl;
    r.f = q - fl;
    r.omf = 1.0f - r.f;
    return r;
}

__device__ __forceinline__ float lerp32(float a, float b, float f, float omf)
{
    return fmaf(omf, a, f * b);
}

__device__ __forceinline__ float clampf(float a, float lo, float hi)
{
    return fminf(fmaxf(lo, a), hi);
}

// Two independent lerps in one instruction pair: sm_100a packed fp32 (FMUL2 + FFMA2).  Each
// lane is the IEEE operation of lerp32, so results are bit-identical; only issue slots halve.
__device__ __forceinline__ float2 lerp32x2(float2 a, float2 b, float2 f, float2 omf)
{
#ifdef BMQ_NO_PACKED_FP32
    return make_float2(lerp32(a.x, b.x, f.x, omf.x), lerp32(a.y, b.y, f.y, omf.y));
#else
    return __ffma2_rn(omf, a, __fmul2_rn(f, b));
#endif
}

// Trilinear sample of a dense x-fastest array; p points at node (x.i, y.i, z.i).
// The x- and y-lerps of the two z-planes are paired ((z0,z1) lanes), so the nest costs
// 2+2+... = 4 packed pairs + 1 scalar lerp instead of 7 scalar lerps; order x -> y -> z kept.
__device__ __forceinline__ float tri8(const float *__restrict__ p, int sy, int sz, const Frac &x,
                                      const Frac &y, const Frac &z)
{
    const float2 P0 = make_float2(__ldg(p), __ldg(p + sz));
    const float2 P1 = make_float2(__ldg(p + 1), __ldg(p + sz + 1));
    const float2 Q0 = make_float2(__ldg(p + sy), __ldg(p + sz + sy));
    const float2 Q1 = make_float2(__ldg(p + sy + 1), __ldg(p + sz + sy + 1));
    const float2 fx = make_float2(x.f, x.f), ox = make_float2(x.omf, x.omf);
    const float2 fy = make_float2(y.f, y.f), oy = make_float2(y.omf, y.omf);
    const float2 A = lerp32x2(P0, P1, fx, ox);   // (y0 row of z0, y0 row of z1)
    const float2 B = lerp32x2(Q0, Q1, fx, ox);   // (y1 row of z0, y1 row of z1)
    const float2 C = lerp32x2(A, B, fy, oy);     // (z0, z1)
    return lerp32(C.x, C.y, z.f, z.omf);
}

// sample_buffer (GPU_kernel.cu:43-62) for a field of x-extent nx, y-extent ny whose origin is
// (ox,oy,oz) (= -0.5h on a staggered axis): p is the WORLD position.
template <bool P2>
__device__ __forceinline__ float sample(const float *__restrict__ b, int nx, int ny, const Grid3 &g,
                                        float ox, float oy, float oz, float px, float py, float pz)
{
    Frac x = split<P2>(px - ox, g.h, g.inv_h);
    Frac y = split<P2>(py - oy, g.h, g.inv_h);
    Frac z = split<P2>(pz - oz, g.h, g.inv_h);
    int sy = nx, sz = nx * ny;
    return tri8(b + (x.i + sy * y.i + sz * z.i), sy, sz, x, y, z);
}

struct Vel3 {
    const float *__restrict__ u;
    const float *__restrict__ v;
    const float *__restrict__ w;
};

// getVelocity (GPU_kernel.cu:64-72): the three staggered samples share the six distinct
// (axis, offset) splits instead of computing nine.
template <bool P2>
__device__ __forceinline__ float3 get_velocity(const Vel3 &vel, const Grid3 &g, float px, float py,
                                               float pz)
{
    const float hh = 0.5f * g.h;   // pos - (-0.5h)
    Frac x0 = split<P2>(px, g.h, g.inv_h), x5 = split<P2>(px + hh, g.h, g.inv_h);
    Frac y0 = split<P2>(py, g.h, g.inv_h), y5 = split<P2>(py + hh, g.h, g.inv_h);
    Frac z0 = split<P2>(pz, g.h, g.inv_h), z5 = split<P2>(pz + hh, g.h, g.inv_h);
    float3 r;
    {
        int sy = g.ni + 1, sz = (g.ni + 1) * g.nj;
        r.x = tri8(vel.u + (x5.i + sy * y0.i + sz * z0.i), sy, sz, x5, y0, z0);
    }
    {
        int sy = g.ni, sz = g.ni * (g.nj + 1);
        r.y = tri8(vel.v + (x0.i + sy * y5.i + sz * z0.i), sy, sz, x0, y5, z0);
    }
    {
        int sy = g.ni, sz = g.ni * g.nj;
        r.z = tri8(vel.w + (x0.i + sy * y0.i + sz * z5.i), sy, sz, x0, y0, z5);
    }
    return r;
}

// getVelocity at the grid node (i,j,k) when h is a power of two: every fraction is exactly 0
// except 1/2 along the sampled component's own axis.
__device__ __forceinline__ float3 velocity_at_node(const Vel3 &vel, const Grid3 &g, int i, int j, int k)
{
    float3 r;
    {
        const float *p = vel.u + (i + (g.ni + 1) * (j + g.nj * k));
        r.x = lerp32(__ldg(p), __ldg(p + 1), 0.5f, 0.5f);
    }
    {
        const float *p = vel.v + (i + g.ni * (j + (g.nj + 1) * k));
        r.y = lerp32(__ldg(p), __ldg(p + g.ni), 0.5f, 0.5f);
    }
    {
        const float *p = vel.w + (i + g.ni * (j + g.nj * k));
        r.z = lerp32(__ldg(p), __ldg(p + g.ni * g.nj), 0.5f, 0.5f);
    }
    return r;
}

// ---- bit-exact variant of the reference sampler, used ONLY where the reference's own formula
// amplifies last-ulp differences: the DMC update computes 1 - exp(-a*s) in fp32
// (GPU_kernel.cu:194-196), so one ulp of the velocity samples that form `a` can move the
// back-traced point by up to ~6e-8/(a*s) of a cell.  To stay within tolerance of the reference
// there, the two velocity evaluations of the DMC kernel reproduce the reference lerp exactly
// as nvcc compiles it: (double)(1.0 - c) * a + (double)(float)(c*b), one double fma, rounded
// to float.  (When h is a power of two the fractions at grid points are exactly 0 or 1/2 and
// the fp32 lerp is already bit-identical, so the P2 path does not need this.)
__device__ __forceinline__ float lerp_ref(float a, float b, float c)
{
    return __double2float_rn(__fma_rn(__dsub_rn(1.0, (double)c), (double)a, (double)__fmul_rn(c, b)));
}

__device__ __forceinline__ float tri8_ref(const float *__restrict__ p, int sy, int sz, const Frac &x,
                                          const Frac &y, const Frac &z)
{
    float v000 = __ldg(p), v001 = __ldg(p + 1);
    float v010 = __ldg(p + sy), v011 = __ldg(p + sy + 1);
    float v100 = __ldg(p + sz), v101 = __ldg(p + sz + 1);
    float v110 = __ldg(p + sz + sy), v111 = __ldg(p + sz + sy + 1);
    float a0 = lerp_ref(v000, v001, x.f), a1 = lerp_ref(v010, v011, x.f);
    float a2 = lerp_ref(v100, v101, x.f), a3 = lerp_ref(v110, v111, x.f);
    return lerp_ref(lerp_ref(a0, a1, y.f), lerp_ref(a2, a3, y.f), z.f);
}

__device__ __forceinline__ float3 get_velocity_ref(const Vel3 &vel, const Grid3 &g, float px, float py, float pz)
{
    const float hh = 0.5f * g.h;
    Frac x0 = split<false>(px, g.h, g.inv_h), x5 = split<false>(px + hh, g.h, g.inv_h);
    Frac y0 = split<false>(py, g.h, g.inv_h), y5 = split<false>(py + hh, g.h, g.inv_h);
    Frac z0 = split<false>(pz, g.h, g.inv_h), z5 = split<false>(pz + hh, g.h, g.inv_h);
    float3 r;
    {
        int sy = g.ni + 1, sz = (g.ni + 1) * g.nj;
        r.x = tri8_ref(vel.u + (x5.i + sy * y0.i + sz * z0.i), sy, sz, x5, y0, z0);
    }
    {
        int sy = g.ni, sz = g.ni * (g.nj + 1);
        r.y = tri8_ref(vel.v + (x0.i + sy * y5.i + sz * z0.i), sy, sz, x0, y5, z0);
    }
    {
        int sy = g.ni, sz = g.ni * g.nj;
        r.z = tri8_ref(vel.w + (x0.i + sy * y0.i + sz * z5.i), sy, sz, x0, y0, z5);
    }
    return r;
}

// traceRK3 (GPU_kernel.cu:74-90): Ralston RK3 with the reference's clamp band [h,(n-1)h].
// c1,c2,c3 are rounded from double like the reference.  The reference forms the midpoints in
// double, input + (0.5*dt)*v1 and input + (0.75*dt)*v2, and rounds once to float: 0.5*dt is
// exact in fp32, so one fmaf gives the identical result; 0.75*dt is not, so that midpoint is
// formed with one double fma (3 per sub-step; everything else stays fp32).
template <bool P2>
__device__ __forceinline__ float3 trace_rk3(const Vel3 &vel, const Grid3 &g, float dt, float3 p)
{
    const float c1 = (float)(2.0 / 9.0 * (double)dt);
    const float c2 = (float)(3.0 / 9.0 * (double)dt);
    const float c3 = (float)(4.0 / 9.0 * (double)dt);
    const float hd = 0.5f * dt;
    const double qd = 0.75 * (double)dt;
    float3 v1 = get_velocity<P2>(vel, g, p.x, p.y, p.z);
    float3 v2 = get_velocity<P2>(vel, g, fmaf(hd, v1.x, p.x), fmaf(hd, v1.y, p.y), fmaf(hd, v1.z, p.z));
    float3 v3 = get_velocity<P2>(vel, g, __double2float_rn(__fma_rn(qd, (double)v2.x, (double)p.x)),
                                 __double2float_rn(__fma_rn(qd, (double)v2.y, (double)p.y)),
                                 __double2float_rn(__fma_rn(qd, (double)v2.z, (double)p.z)));
    float3 o;
    o.x = fmaf(c3, v3.x, fmaf(c2, v2.x, fmaf(c1, v1.x, p.x)));
    o.y = fmaf(c3, v3.y, fmaf(c2, v2.y, fmaf(c1, v1.y, p.y)));
    o.z = fmaf(c3, v3.z, fmaf(c2, v2.z, fmaf(c1, v1.z, p.z)));
    o.x = clampf(o.x, g.h, (float)g.ni * g.h - g.h);
    o.y = clampf(o.y, g.h, (float)g.nj * g.h - g.h);
    o.z = clampf(o.z, g.h, (float)g.nk * g.h - g.h);
    return o;
}

// trace (GPU_kernel.cu:92-125): sub-steps of cfldt until |dt| is covered; the float loop is the
// reference's, so every thread takes the same sub-step sequence as the reference does.
template <bool P2>
__device__ __forceinline__ float3 trace(const Vel3 &vel, const Grid3 &g, float cfldt, float dt, float3 p)
{
    const float sgn = dt > 0.f ? 1.0f : -1.0f;
    const float T = fabsf(dt);
    float t = 0.f, sub = cfldt;
    while (t < T) {
        if (t + sub > T) sub = T - t;
        p = trace_rk3<P2>(vel, g, sgn * sub, p);
        t += sub;
    }
    return p;
}

// N independent particles traced in lock-step: the sub-step sequence depends only on (cfldt, dt),
// so the RK3 stages of the particles are interleaved inside one loop, which keeps N times as many
// velocity gathers in flight as tracing them one after the other (same arithmetic per particle).
template <bool P2, int N>
__device__ __forceinline__ void trace_multi(const Vel3 &vel, const Grid3 &g, float cfldt, float dt, float3 (&p)[N])
{
    const float sgn = dt > 0.f ? 1.0f : -1.0f;
    const float T = fabsf(dt);
    float t = 0.f, sub = cfldt;
    while (t < T) {
        if (t + sub > T) sub = T - t;
        const float h_ = sgn * sub;
        const float c1 = (float)(2.0 / 9.0 * (double)h_), c2 = (float)(3.0 / 9.0 * (double)h_), c3 = (float)(4.0 / 9.0 * (double)h_);
        const float hd = 0.5f * h_;
        const double qd = 0.75 * (double)h_;
        float3 v1[N], v2[N], v3[N];
#pragma unroll
        for (int m = 0; m < N; ++m) v1[m] = get_velocity<P2>(vel, g, p[m].x, p[m].y, p[m].z);
#pragma unroll
        for (int m = 0; m < N; ++m)
            v2[m] = get_velocity<P2>(vel, g, fmaf(hd, v1[m].x, p[m].x), fmaf(hd, v1[m].y, p[m].y), fmaf(hd, v1[m].z, p[m].z));
#pragma unroll
        for (int m = 0; m < N; ++m)
            v3[m] = get_velocity<P2>(vel, g, __double2float_rn(__fma_rn(qd, (double)v2[m].x, (double)p[m].x)),
                                     __double2float_rn(__fma_rn(qd, (double)v2[m].y, (double)p[m].y)),
                                     __double2float_rn(__fma_rn(qd, (double)v2[m].z, (double)p[m].z)));
#pragma unroll
        for (int m = 0; m < N; ++m) {
            float3 o;
            o.x = fmaf(c3, v3[m].x, fmaf(c2, v2[m].x, fmaf(c1, v1[m].x, p[m].x)));
            o.y = fmaf(c3, v3[m].y, fmaf(c2, v2[m].y, fmaf(c1, v1[m].y, p[m].y)));
            o.z = fmaf(c3, v3[m].z, fmaf(c2, v2[m].z, fmaf(c1, v1[m].z, p[m].z)));
            p[m].x = clampf(o.x, g.h, (float)g.ni * g.h - g.h);
            p[m].y = clampf(o.y, g.h, (float)g.nj * g.h - g.h);
            p[m].z = clampf(o.z, g.h, (float)g.nk * g.h - g.h);
        }
        t += sub;
    }
}

// Three co-located map components sampled at one world position (the maps are cell-centred
// arrays of the global grid with origin 0): one split, 24 loads.
struct Map3 {
    const float *__restrict__ x;
    const float *__restrict__ y;
    const float *__restrict__ z;
};

template <bool P2>
__device__ __forceinline__ float3 sample_map(const Map3 &m, const Grid3 &g, float px, float py, float pz)
{
    Frac x = split<P2>(px, g.h, g.inv_h);
    Frac y = split<P2>(py, g.h, g.inv_h);
    Frac z = split<P2>(pz, g.h, g.inv_h);
    int sy = g.ni, sz = g.ni * g.nj;
    int o = x.i + sy * y.i + sz * z.i;
    float3 r;
    r.x = tri8(m.x + o, sy, sz, x, y, z);
    r.y = tri8(m.y + o, sy, sz, x, y, z);
    r.z = tri8(m.z + o, sy, sz, x, y, z);
    return r;
}

// NF co-located fields (same staggering) sampled at one world position: one split.
template <bool P2, int NF>
__device__ __forceinline__ void sample_fields(const float *const (&src)[NF], int nx, int ny,
                                              const Grid3 &g, float ox, float oy, float oz, float px,
                                              float py, float pz, float (&out)[NF])
{
    Frac x = split<P2>(px - ox, g.h, g.inv_h);
    Frac y = split<P2>(py - oy, g.h, g.inv_h);
    Frac z = split<P2>(pz - oz, g.h, g.inv_h);
    int sy = nx, sz = nx * ny;
    int o = x.i + sy * y.i + sz * z.i;
#pragma unroll
    for (int f = 0; f < NF; ++f) out[f] = tri8(src[f] + o, sy, sz, x, y, z);
}

// The quadrature shared by advect / compensate / cumulate (GPU_kernel.cu:317-371): eight
// sub-cell points at +-h/4 (or one point when is_point) plus the centre.  For every point:
// sample the map, clamp to [lo,hi], sample NS co-located source fields there.  Returns
//   sum[f]   = sum_ii wgt[f] * src_f(map(c + off_ii))   accumulated in the reference's order
//   value[f] = src_f(map(c))
template <bool P2, int NS>
__device__ __forceinline__ void quad_gather(const Map3 &m, const Grid3 &g, float cx, float cy, float cz,
                                            float lo, float hix, float hiy, float hiz,
                                            const float *const (&src)[NS], int fnx, int fny, float ox,
                                            float oy, float oz, bool is_point, const float (&wgt)[NS],
                                            float (&sum)[NS], float (&value)[NS])
{
#pragma unroll
    for (int f = 0; f < NS; ++f) sum[f] = 0.f;
    const float q = 0.25f * g.h;
    const int ev = is_point ? 1 : 8;
    for (int ii = 0; ii < ev; ++ii) {
        // offsets in the reference's order: x sign = bit 2, y sign = bit 1, z sign = bit 0
        float dx = is_point ? 0.f : ((ii & 4) ? -q : q);
        float dy = is_point ? 0.f : ((ii & 2) ? -q : q);
        float dz = is_point ? 0.f : ((ii & 1) ? -q : q);
        float3 mp = sample_map<P2>(m, g, cx + dx, cy + dy, cz + dz);
        mp.x = clampf(mp.x, lo, hix);
        mp.y = clampf(mp.y, lo, hiy);
        mp.z = clampf(mp.z, lo, hiz);
        float s[NS];
        sample_fields<P2, NS>(src, fnx, fny, g, ox, oy, oz, mp.x, mp.y, mp.z, s);
#pragma unroll
        for (int f = 0; f < NS; ++f) sum[f] = fmaf(wgt[f], s[f], sum[f]);
    }
    float3 mp = sample_map<P2>(m, g, cx, cy, cz);
    mp.x = clampf(mp.x, lo, hix);
    mp.y = clampf(mp.y, lo, hiy);
    mp.z = clampf(mp.z, lo, hiz);
    sample_fields<P2, NS>(src, fnx, fny, g, ox, oy, oz, mp.x, mp.y, mp.z, value);
}

// ---- windowed quadrature ---------------------------------------------------------------------
// The eight sub-cell points of an output cell sit at +-h/4 around a grid-aligned position, so
// on every axis their map samples use one of two (cell, fraction) pairs and all of them read
// from one 3x3x3 (2 on the staggered axis) window of map nodes around the cell.  The reference
// fetches 9 x 8 nodes per map component (GPU_kernel.cu:350-352); here the window is loaded once
// (27 or 18 loads) and the 8 trilinear samples are evaluated separably -- 18 x-lerps, 12
// y-lerps, 8 z-lerps, each the SAME fmaf(1-f, a, f*b) on the same operands in the same x->y->z
// order as eight independent trilerps, so the results are bit-identical to them.
//   axis geometry (cell index c, window base c-1):   minus point        plus point       centre
//     unstaggered axis (pos = c h):                 cell c-1, f~3/4    cell c,   f~1/4   cell c,   f~0
//     staggered axis   (pos = (c-1/2) h):           cell c-1, f~1/4    cell c-1, f~3/4   cell c-1, f~1/2
// The fractions are computed with the reference's expression (position/h - floor) unless h is
// a power of two, where they are exactly the constants above.  floor() of the +-1/4 points is
// structurally safe (they are 1/4 cell away from an integer); the centre point's floor is not
// (c h / h may round below c), so without P2 the centre sample uses the generic 8-node path.
template <bool P2, bool STAGGERED>
struct AxisW {
    float fm, om, fp, op;   // minus / plus fractions and their complements
    __device__ __forceinline__ void init(float c_pos, float h, float inv_h)
    {
        if (P2) {
            fm = STAGGERED ? 0.25f : 0.75f;
            fp = STAGGERED ? 0.75f : 0.25f;
        } else {
            const float q = 0.25f * h;
            const float qm = div_h(c_pos - q, h, inv_h), qp = div_h(c_pos + q, h, inv_h);
            fm = qm - floorf(qm);
            fp = qp - floorf(qp);
        }
        om = 1.0f - fm;
        op = 1.0f - fp;
    }
};

// 8 corner samples (index = reference order ii: x sign bit 2, y sign bit 1, z sign bit 0; 0 = plus)
// and, when P2, the centre sample of ONE map component from its node window.
// `p` points at node (i-1, j-1, k-1) of the component array.
template <bool P2, int STAG>
__device__ __forceinline__ void window_samples(const float *__restrict__ p, int sy, int sz,
                                               const AxisW<P2, STAG == 1> &ax, const AxisW<P2, STAG == 2> &ay,
                                               const AxisW<P2, STAG == 3> &az, float (&out)[8], float &centre)
{
    constexpr int NX = STAG == 1 ? 2 : 3, NY = STAG == 2 ? 2 : 3, NZ = STAG == 3 ? 2 : 3;
    // every value is a pair over the x sign: .x = plus point, .y = minus point (packed fp32)
    const float2 fx = make_float2(ax.fp, ax.fm), ox = make_float2(ax.op, ax.om);
    const float2 fym = make_float2(ay.fm, ay.fm), oym = make_float2(ay.om, ay.om);
    const float2 fyp = make_float2(ay.fp, ay.fp), oyp = make_float2(ay.op, ay.op);
    const float2 fzm = make_float2(az.fm, az.fm), ozm = make_float2(az.om, az.om);
    const float2 fzp = make_float2(az.fp, az.fp), ozp = make_float2(az.op, az.op);
    float2 Y[NZ][2];     // [z node][y sign] (0 = plus, 1 = minus)
    float cyz[NZ];       // centre-line values per z node (P2 only)
#pragma unroll
    for (int z = 0; z < NZ; ++z) {
        float2 X[NY];
        float xc[NY];
#pragma unroll
        for (int y = 0; y < NY; ++y) {
            const float *r = p + y * sy + z * sz;
            const float n0 = __ldg(r), n1 = __ldg(r + 1);
            const float n2 = NX == 3 ? __ldg(r + 2) : 0.f;
            // plus point: cells (n1,n2) unstaggered / (n0,n1) staggered; minus point: (n0,n1)
            X[y] = NX == 3 ? lerp32x2(make_float2(n1, n0), make_float2(n2, n1), fx, ox)
                           : lerp32x2(make_float2(n0, n0), make_float2(n1, n1), fx, ox);
            if (P2) xc[y] = NX == 3 ? n1 : lerp32(n0, n1, 0.5f, 0.5f);
        }
        Y[z][1] = lerp32x2(X[0], X[1], fym, oym);
        Y[z][0] = NY == 3 ? lerp32x2(X[1], X[2], fyp, oyp) : lerp32x2(X[0], X[1], fyp, oyp);
        if (P2) cyz[z] = NY == 3 ? xc[1] : lerp32(xc[0], xc[1], 0.5f, 0.5f);
    }
#pragma unroll
    for (int sy_ = 0; sy_ < 2; ++sy_) {
        const float2 zm = lerp32x2(Y[0][sy_], Y[1][sy_], fzm, ozm);
        const float2 zp = NZ == 3 ? lerp32x2(Y[1][sy_], Y[2][sy_], fzp, ozp) : lerp32x2(Y[0][sy_], Y[1][sy_], fzp, ozp);
        out[0 * 4 + sy_ * 2 + 1] = zm.x;   // x plus,  z minus
        out[1 * 4 + sy_ * 2 + 1] = zm.y;   // x minus, z minus
        out[0 * 4 + sy_ * 2 + 0] = zp.x;
        out[1 * 4 + sy_ * 2 + 0] = zp.y;
    }
    if (P2) centre = NZ == 3 ? cyz[1] : lerp32(cyz[0], cyz[1], 0.5f, 0.5f);
}

// Windowed version of quad_gather for the 8-point quadrature (is_point == false).
// (i,j,k) is the output cell in its own (staggered) array; (cx,cy,cz) its world position.
template <bool P2, int STAG, int NS>
__device__ __forceinline__ void quad_gather_win(const Map3 &m, const Grid3 &g, int i, int j, int k, float cx,
                                                float cy, float cz, float lo, float hix, float hiy, float hiz,
                                                const float *const (&src)[NS], int fnx, int fny, float ox,
                                                float oy, float oz, const float (&wgt)[NS], float (&sum)[NS],
                                                float (&value)[NS])
{
    AxisW<P2, STAG == 1> ax;
    AxisW<P2, STAG == 2> ay;
    AxisW<P2, STAG == 3> az;
    ax.init(cx, g.h, g.inv_h);
    ay.init(cy, g.h, g.inv_h);
    az.init(cz, g.h, g.inv_h);
    const int sy = g.ni, sz = g.ni * g.nj;
    const int base = (i - 1) + sy * (j - 1) + sz * (k - 1);
    float px[8], py[8], pz[8], ccx, ccy, ccz;
    window_samples<P2, STAG>(m.x + base, sy, sz, ax, ay, az, px, ccx);
    window_samples<P2, STAG>(m.y + base, sy, sz, ax, ay, az, py, ccy);
    window_samples<P2, STAG>(m.z + base, sy, sz, ax, ay, az, pz, ccz);
    if (!P2) {
        float3 c = sample_map<false>(m, g, cx, cy, cz);
        ccx = c.x; ccy = c.y; ccz = c.z;
    }
#pragma unroll
    for (int f = 0; f < NS; ++f) sum[f] = 0.f;
#pragma unroll
    for (int ii = 0; ii < 8; ++ii) {
        float s[NS];
        sample_fields<P2, NS>(src, fnx, fny, g, ox, oy, oz, clampf(px[ii], lo, hix), clampf(py[ii], lo, hiy),
                              clampf(pz[ii], lo, hiz), s);
#pragma unroll
        for (int f = 0; f < NS; ++f) sum[f] = fmaf(wgt[f], s[f], sum[f]);
    }
    sample_fields<P2, NS>(src, fnx, fny, g, ox, oy, oz, clampf(ccx, lo, hix), clampf(ccy, lo, hiy),
                          clampf(ccz, lo, hiz), value);
}

// ---- reductions -------------------------------------------------------------------------
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide max of non-negative floats, then one atomicMax per block on the int pattern
// (valid because the values are >= 0).
__device__ __forceinline__ void block_atomic_max(float v, float *dst)
{
    __shared__ float smax[32];
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nthreads = blockDim.x * blockDim.y * blockDim.z;
    v = warp_max(v);
    if ((tid & 31) == 0) smax[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
        float r = tid < (nthreads + 31) / 32 ? smax[tid] : 0.f;
        r = warp_max(r);
        if (tid == 0 && r > 0.f) atomicMax(reinterpret_cast<int *>(dst), __float_as_int(r));
    }
}

}  // namespace bmq
