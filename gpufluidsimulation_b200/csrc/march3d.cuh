// z-marching gather kernels: advect / time-0 error / accumulate / apply-and-clamp.
//
// The windowed kernels of kernels3d.cu load, for every output cell, the 3x3x3 (2 along a staggered axis)
// window of map nodes around it: 81 loads per cell and map, of which 54 are the two z-planes the cell
// below it has just loaded.  Here a thread owns a COLUMN of cells (i, j, k0..k1): it keeps the x- and
// y-interpolated values of the last two map planes in registers and loads one new plane per cell
// (27 loads instead of 81; 12 instead of 38 packed lerps per map component).
//
// What bounds these kernels on B200 is neither HBM nor the load count but the FP32 pipe: FFMA / FMUL /
// FADD / IMAD share one pipe that takes a warp instruction every second cycle per scheduler
// (B300_MICROARCH.md "Pipe rates"), and a trilinear field sample is ~26 such instructions when written
// naively.  The gather core below therefore
//   * works in GRID units when h is a power of two: the scaling by 1/h is folded into the (constant)
//     z-weights of the map interpolation, which is exact (power-of-two scaling commutes with rounding);
//   * evaluates the eight corner samples as four PAIRS (x plus, x minus) in packed fp32 (FMUL2 / FFMA2):
//     the two lanes are the two samples, each lane the scalar operation of the reference nest
//     x -> y -> z, so every sample is bit-identical to an independent trilerp;
//   * accumulates the quadrature sum in the reference's order ii = 0..7 (GPU_kernel.cu:350-357).
// Results are bit-identical to the windowed kernels and, through them, to the reference
// (GPU_kernel.cu:312-499); tests/test_march_gpu.py compares the two variants.
//
// The apply kernel's 27-neighbour extrema clamp (clampExtrema_kernel, GPU_kernel.cu:146-167) marches the
// same way: min/max of the 3x3 neighbourhood per plane are carried, 9 loads per cell instead of 27.
#pragma once
#include "launch3d.h"
#include "device3d.cuh"

namespace bmq {

// GM_APPLY_NC: the apply kernel without its extrema clamp (the clamp then runs as a separate shared-memory tiled
// stencil kernel, clamp27.cu: bmq_set_gather_variant(2))
enum { GM_ADVECT = 0, GM_ERROR = 1, GM_CUMULATE = 2, GM_APPLY = 3, GM_APPLY_NC = 4 };

#ifndef BMQ_MARCH_BY
#define BMQ_MARCH_BY 4          // CTA = 32 x BMQ_MARCH_BY columns
#endif
#ifndef BMQ_MARCH_MINBLOCKS
#define BMQ_MARCH_MINBLOCKS 6   // x 128 threads: 24 warps per SM at <= 80 registers (sweep: profiles/r2_march_variants.md)
#endif
#ifndef BMQ_MARCH_PREFETCH
#define BMQ_MARCH_PREFETCH 1    // 1 / 2: prefetch the next cell's new map plane and field plane into L1 / L2
#endif
#ifndef BMQ_MARCH_PF_STREAM
#define BMQ_MARCH_PF_STREAM 1   // prefetch the next cell's streamed (non-gathered) operands (error / accumulate kernels)
#endif
#ifndef BMQ_MARCH_EARLY_CENTRE
#define BMQ_MARCH_EARLY_CENTRE 0
#endif
#ifndef BMQ_MARCH_PF_ROWS
#define BMQ_MARCH_PF_ROWS 2     // field rows prefetched around the centre sample's base node (2: j, j+1; 3: also j-1)
#endif
#ifndef BMQ_MARCH_PF_DIST
#define BMQ_MARCH_PF_DIST 2     // field plane prefetched, relative to the centre sample's base plane
#endif

__device__ __forceinline__ void prefetch_line(const float *p)
{
#if BMQ_MARCH_PREFETCH == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#elif BMQ_MARCH_PREFETCH == 2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// Timing experiments only: the marginal cost, inside the real kernels, of one more load per gathered node.  The
// extra load's bits are AND-ed with a run-time zero and OR-ed into the value, so results stay exact and the
// compiler cannot drop it.  3: the same (misaligned) global load again; 4: a shared-memory load; 5: a global load
// aligned to its 128-byte line.  profiles/r2_march_variants.md
#ifndef BMQ_HACK_GATHER
#define BMQ_HACK_GATHER 0
#endif
#if BMQ_HACK_GATHER == 4
extern __shared__ float bmq_hack_smem[];
#endif
__device__ __forceinline__ float gld(const float *p, unsigned zero)
{
    const float v = __ldg(p);
#if BMQ_HACK_GATHER == 3
    float e;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(e) : "l"(p));
#elif BMQ_HACK_GATHER == 4
    const float e = bmq_hack_smem[(reinterpret_cast<size_t>(p) >> 2) & 1023];
#elif BMQ_HACK_GATHER == 5
    float e;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(e) : "l"(reinterpret_cast<const float *>(reinterpret_cast<size_t>(p) & ~size_t(127)) + (threadIdx.x & 31)));
#endif
#if BMQ_HACK_GATHER >= 3
    return __int_as_float(__float_as_int(v) | (__float_as_int(e) & (int)zero));
#else
    (void)zero;
    return v;
#endif
}

template <int NF, int NS> struct MarchArgs {
    float *out[NF];          // advect: f_adv | error: e0 | cumulate: target (read-modify-write) | apply: f
    const float *src[NS];    // gathered through the map: init | f_adv | change sets [set][field] | e0
    const float *aux[NF];    // error: init | apply: f_adv | otherwise unused
    float coeff[NS / NF];    // cumulate: one coefficient per change set
};

// x- and y-interpolated values of ONE z-node plane of ONE map component: the part of window_samples
// (device3d.cuh) that does not depend on z.  yp / ym: y plus / minus point, each a pair over the x sign
// (.x = plus, .y = minus); c: centre line (power-of-two h only).
struct PlaneXY {
    float2 yp, ym;
    float c;
};

template <bool P2, int STAG>
__device__ __forceinline__ PlaneXY plane_xy(const float *__restrict__ p, int sy, const AxisW<P2, STAG == 1> &ax,
                                            const AxisW<P2, STAG == 2> &ay)
{
    constexpr int NX = STAG == 1 ? 2 : 3, NY = STAG == 2 ? 2 : 3;
    const float2 fx = make_float2(ax.fp, ax.fm), ox = make_float2(ax.op, ax.om);
    const float2 fym = make_float2(ay.fm, ay.fm), oym = make_float2(ay.om, ay.om);
    const float2 fyp = make_float2(ay.fp, ay.fp), oyp = make_float2(ay.op, ay.op);
    float2 X[NY];
    float xc[NY];
#pragma unroll
    for (int y = 0; y < NY; ++y) {
        const float *r = p + y * sy;
        const float n0 = __ldg(r), n1 = __ldg(r + 1);
        const float n2 = NX == 3 ? __ldg(r + 2) : 0.f;
        X[y] = NX == 3 ? lerp32x2(make_float2(n1, n0), make_float2(n2, n1), fx, ox)
                       : lerp32x2(make_float2(n0, n0), make_float2(n1, n1), fx, ox);
        if (P2) xc[y] = NX == 3 ? n1 : lerp32(n0, n1, 0.5f, 0.5f);
    }
    PlaneXY o;
    o.ym = lerp32x2(X[0], X[1], fym, oym);
    o.yp = NY == 3 ? lerp32x2(X[1], X[2], fyp, oyp) : lerp32x2(X[0], X[1], fyp, oyp);
    o.c = 0.f;
    if (P2) o.c = NY == 3 ? xc[1] : lerp32(xc[0], xc[1], 0.5f, 0.5f);
    return o;
}

// The z part of window_samples.  pos[ii], ii = 2*(y minus) + (z minus), is the PAIR (x plus, x minus) of
// corner samples ii and ii + 4 of the reference's order (x sign bit 2, y sign bit 1, z sign bit 0; 0 = plus).
// `scale` multiplies the z weights: 1 (world units) or 1/h, a power of two (grid units; exact).
template <bool P2, int STAG>
__device__ __forceinline__ void z_combine(const PlaneXY &Pa, const PlaneXY &Pb, const PlaneXY &Pc, const AxisW<P2, STAG == 3> &az,
                                          float scale, float2 (&pos)[4], float &centre)
{
    constexpr int NZ = STAG == 3 ? 2 : 3;
    const float fm = P2 ? az.fm * scale : az.fm, om = P2 ? az.om * scale : az.om;
    const float fp = P2 ? az.fp * scale : az.fp, op = P2 ? az.op * scale : az.op;
    const float2 fzm = make_float2(fm, fm), ozm = make_float2(om, om);
    const float2 fzp = make_float2(fp, fp), ozp = make_float2(op, op);
#pragma unroll
    for (int sy_ = 0; sy_ < 2; ++sy_) {
        const float2 y0 = sy_ ? Pa.ym : Pa.yp, y1 = sy_ ? Pb.ym : Pb.yp, y2 = sy_ ? Pc.ym : Pc.yp;
        pos[sy_ * 2 + 1] = lerp32x2(y0, y1, fzm, ozm);                                                   // z minus
        pos[sy_ * 2 + 0] = NZ == 3 ? lerp32x2(y1, y2, fzp, ozp) : lerp32x2(y0, y1, fzp, ozp);            // z plus
    }
    if (P2) centre = NZ == 3 ? Pb.c * scale : lerp32(Pa.c, Pb.c, 0.5f * scale, 0.5f * scale);
}

// ---- the gather core -------------------------------------------------------------------------------------
// Positions reach it clamped: in cells when GRID (power-of-two h), as world positions otherwise.  Returns
// (position - field origin) / h, the reference's `pos / h` of sample_buffer (GPU_kernel.cu:46-51).
template <bool GRID>
__device__ __forceinline__ float to_cells(float p, float off_world, float off_grid, float h, float inv_h)
{
    // GRID: p is already in cells; the staggering offset (+1/2) is the only arithmetic left
    if (GRID) return off_grid != 0.f ? p + off_grid : p;
    return div_h(off_world != 0.f ? p - off_world : p, h, inv_h);
}

// The same for a pair of positions.  General h: the verified division sequence as three PACKED instructions for
// both lanes behind ONE range test (the scalar form tests each lane: 2 x (2 compares + 3) instructions); a lane that
// is exactly 0 -- a position clamped to the lower wall -- or tiny sends the pair down the scalar path.
template <bool GRID>
__device__ __forceinline__ float2 to_cells2(float2 p, float off_world, float off_grid, float h, float inv_h)
{
    if (GRID) return make_float2(to_cells<true>(p.x, off_world, off_grid, h, inv_h), to_cells<true>(p.y, off_world, off_grid, h, inv_h));
#if defined(BMQ_NO_PACKED_FP32) || defined(BMQ_DIV_SCALAR)
    return make_float2(to_cells<false>(p.x, off_world, off_grid, h, inv_h), to_cells<false>(p.y, off_world, off_grid, h, inv_h));
#else
    if (off_world != 0.f) p = make_float2(p.x - off_world, p.y - off_world);
    if (inv_h > 0.f && fminf(fabsf(p.x), fabsf(p.y)) >= BMQ_DIV_TINY) {
        const float2 y = make_float2(inv_h, inv_h), mh = make_float2(-h, -h);
        const float2 q0 = __fmul2_rn(p, y);
        const float2 r = __ffma2_rn(mh, q0, p);
        return __ffma2_rn(r, y, q0);
    }
    return make_float2(div_h(p.x, h, inv_h), div_h(p.y, h, inv_h));
#endif
}

struct Split2 {
    float2 f, omf;
    int i0, i1;
};
// (cell, fraction, 1 - fraction) of two coordinates at once: q - floor(q) and 1 - f as exact FFMA2s with -1.
// BMQ_SPLIT_MODE selects how floor and the integer cell index are obtained (0 <= q < 2^22 here: positions are
// clamped to the domain).  All three give the same bits; they load different pipes:
//   0: FRND.FLOOR + F2I            (two conversions on the narrow XU pipe per coordinate)
//   1: F2I.FLOOR, floor as float from the integer by exponent splicing: (2^23 | i) - 2^23   (one XU op)
//   2: no conversion at all: RN(q) = (q + 2^23) - 2^23, floor = RN(q) - (RN(q) > q), index from the mantissa bits
// measured per 512^3 step: 0 -> 79.5 ms, 1 -> 78.6 ms, 2 -> 83.0 ms (profiles/r2_march_variants.md)
#ifndef BMQ_SPLIT_MODE
#define BMQ_SPLIT_MODE 1
#endif
__device__ __forceinline__ void floor_index(float q, float &fl, int &i)
{
#if BMQ_SPLIT_MODE == 1
    i = __float2int_rd(q);
    fl = __fadd_rn(__int_as_float(0x4B000000 | i), -8388608.0f);
#elif BMQ_SPLIT_MODE == 3
    i = __float2int_rd(q);              // F2I.FLOOR + I2FP: the integer-to-float conversion is exact here (|i| < 2^22)
    fl = (float)i;
#elif BMQ_SPLIT_MODE == 2
    const float t = __fadd_rn(q, 8388608.0f);            // integer-valued float in [2^23, 2^24): RN(q) + 2^23
    const float r = __fadd_rn(t, -8388608.0f);           // RN(q), exact
    const bool up = r > q;                                 // rounded up: floor is one less
    i = (__float_as_int(t) - 0x4B000000) - (up ? 1 : 0);
    fl = up ? __fadd_rn(r, -1.0f) : r;
#else
    fl = floorf(q);
    i = (int)fl;
#endif
}

__device__ __forceinline__ Split2 split2(float2 q)
{
    float2 fl;
    Split2 s;
    floor_index(q.x, fl.x, s.i0);
    floor_index(q.y, fl.y, s.i1);
    const float2 m1 = make_float2(-1.0f, -1.0f), one = make_float2(1.0f, 1.0f);
#if defined(BMQ_NO_PACKED_FP32) || defined(BMQ_SPLIT2_SCALAR)
    s.f = make_float2(q.x - fl.x, q.y - fl.y);
    s.omf = make_float2(1.0f - s.f.x, 1.0f - s.f.y);
#else
    s.f = __ffma2_rn(fl, m1, q);         // q - floor(q): exact
    s.omf = __ffma2_rn(s.f, m1, one);    // 1 - f: the same single rounding as 1.0f - f
#endif
    return s;
}

// two trilinear samples (the two lanes) of NS co-located fields; lane arithmetic = tri8 (device3d.cuh)
template <int NS>
__device__ __forceinline__ void gather_pair(const float *const (&src)[NS], int sy, int sz, const Split2 &x, const Split2 &y,
                                            const Split2 &z, float (&lane0)[NS], float (&lane1)[NS], unsigned zero = 0)
{
    const int o0 = x.i0 + sy * y.i0 + sz * z.i0, o1 = x.i1 + sy * y.i1 + sz * z.i1;
#pragma unroll
    for (int f = 0; f < NS; ++f) {
        const float *p0 = src[f] + o0, *p1 = src[f] + o1;
        const float2 n000 = make_float2(gld(p0, zero), gld(p1, zero)), n001 = make_float2(gld(p0 + 1, zero), gld(p1 + 1, zero));
        const float2 n010 = make_float2(gld(p0 + sy, zero), gld(p1 + sy, zero)), n011 = make_float2(gld(p0 + sy + 1, zero), gld(p1 + sy + 1, zero));
        const float2 n100 = make_float2(gld(p0 + sz, zero), gld(p1 + sz, zero)), n101 = make_float2(gld(p0 + sz + 1, zero), gld(p1 + sz + 1, zero));
        const float2 n110 = make_float2(gld(p0 + sz + sy, zero), gld(p1 + sz + sy, zero));
        const float2 n111 = make_float2(gld(p0 + sz + sy + 1, zero), gld(p1 + sz + sy + 1, zero));
        const float2 a00 = lerp32x2(n000, n001, x.f, x.omf), a01 = lerp32x2(n010, n011, x.f, x.omf);
        const float2 a10 = lerp32x2(n100, n101, x.f, x.omf), a11 = lerp32x2(n110, n111, x.f, x.omf);
        const float2 b0 = lerp32x2(a00, a01, y.f, y.omf), b1 = lerp32x2(a10, a11, y.f, y.omf);
        const float2 c = lerp32x2(b0, b1, z.f, z.omf);
        lane0[f] = c.x;
        lane1[f] = c.y;
    }
}

// one trilinear sample of NS co-located fields from cell coordinates; returns the offset of its base node
template <int NS>
__device__ __forceinline__ int gather_one(const float *const (&src)[NS], int sy, int sz, float qx, float qy, float qz,
                                          float (&out)[NS], int &zi)
{
    Frac x, y, z;
    float fl;
    floor_index(qx, fl, x.i); x.f = qx - fl; x.omf = 1.0f - x.f;
    floor_index(qy, fl, y.i); y.f = qy - fl; y.omf = 1.0f - y.f;
    floor_index(qz, fl, z.i); z.f = qz - fl; z.omf = 1.0f - z.f;
    const int o = x.i + sy * y.i + sz * z.i;
#pragma unroll
    for (int f = 0; f < NS; ++f) out[f] = tri8(src[f] + o, sy, sz, x, y, z);
    zi = z.i;
    return o;
}

// min / max of the 3x3 (x, y) neighbourhood of one plane, and the centre value
struct Plane9 {
    float mn, mx, c;
};
__device__ __forceinline__ Plane9 plane9(const float *__restrict__ p, int sy)
{
    Plane9 o;
    o.c = __ldg(p);
    o.mn = o.mx = o.c;
#pragma unroll
    for (int jj = -1; jj <= 1; ++jj)
#pragma unroll
        for (int ii = -1; ii <= 1; ++ii) {
            if (ii == 0 && jj == 0) continue;
            const float v = __ldg(p + ii + jj * sy);
            o.mx = fmaxf(o.mx, v);
            o.mn = fminf(o.mn, v);
        }
    return o;
}

template <int MODE, bool P2, int STAG, int NF, int NCH, int FIX>
__global__ void __launch_bounds__(32 * BMQ_MARCH_BY, BMQ_MARCH_MINBLOCKS)
k_march(Grid3 g_, int kbeg, int kend, int kchunk, MarchArgs<NF, NF * NCH> a, Map3 m)
{
    const Grid3 g = fix_grid<FIX>(g_);
    const unsigned hack_zero = BMQ_HACK_GATHER ? (unsigned)g_.nk >> 30 : 0u;     // 0 at run time, unknown at compile time
    constexpr int DX = STAG == 1, DY = STAG == 2, DZ = STAG == 3, NZ = STAG == 3 ? 2 : 3, NS = NF * NCH;
    constexpr int GL = MODE == GM_ADVECT ? 2 : 1;     // interior guard GL + D < idx < f - GL - 1 (reference :341/:405/:467)
    const int fi = g.ni + DX, fj = g.nj + DY, fk = g.nk + DZ;
    const int i = blockIdx.x * 32 + threadIdx.x;
    const int j = blockIdx.y * BMQ_MARCH_BY + threadIdx.y;
    if (i >= fi || j >= fj) return;
    const int kc0 = kbeg + blockIdx.z * kchunk;
    const int kc1 = min(kc0 + kchunk, kend);
    const bool gij = GL + DX < i && i < fi - GL - 1 && GL + DY < j && j < fj - GL - 1;
    const int ka = max(kc0, GL + DZ + 1), kb = min(kc1, fk - GL - 1);   // cells [ka, kb) of this column gather

    const float h = g.h;
    // field origin: -h/2 on the staggered axis (GPU_kernel.cu:212,259,332), exactly 0 elsewhere
    const float ox = DX ? -0.5f * h : 0.f, oy = DY ? -0.5f * h : 0.f, oz = DZ ? -0.5f * h : 0.f;
    // explicit roundings: a plain h * i could be contracted into the subtractions of AxisW::init (general h) and
    // change the map fractions by an ulp
    const float cx = DX ? fmaf(h, (float)i, ox) : __fmul_rn(h, (float)i), cy = DY ? fmaf(h, (float)j, oy) : __fmul_rn(h, (float)j);
    // clamp band of the mapped positions: [h, (n-1)h] (advect, :356) or [0, n h] (:419, :479); in cells when P2
    const float ps = P2 ? g.inv_h : 1.0f;
    const float lo = (MODE == GM_ADVECT ? h : 0.f) * ps;
    const float hix = (MODE == GM_ADVECT ? h * (float)g.ni - h : h * (float)g.ni) * ps;
    const float hiy = (MODE == GM_ADVECT ? h * (float)g.nj - h : h * (float)g.nj) * ps;
    const float hiz = (MODE == GM_ADVECT ? h * (float)g.nk - h : h * (float)g.nk) * ps;
    const int col = i + fi * j, fplane = fi * fj;

    constexpr bool APPLYISH = MODE == GM_APPLY || MODE == GM_APPLY_NC;
    if (!APPLYISH) {
        if (!gij || ka >= kb) return;
    } else if (!(i > 0 && i < fi - 1 && j > 0 && j < fj - 1)) {
        // rim columns of the apply kernel: f = f_adv (the reference leaves f_adv untouched there)
        for (int k = kc0; k < kc1; ++k)
#pragma unroll
            for (int f = 0; f < NF; ++f) a.out[f][col + fplane * k] = __ldg(a.aux[f] + col + fplane * k);
        return;
    }

    AxisW<P2, STAG == 1> ax;
    AxisW<P2, STAG == 2> ay;
    AxisW<P2, STAG == 3> az;
    ax.init(cx, h, g.inv_h);
    ay.init(cy, h, g.inv_h);
    az.init(0.f, h, g.inv_h);     // power-of-two h: constants; otherwise re-initialised per cell
    const int sy = g.ni, sz = g.ni * g.nj;
    const int mbase = (i - 1) + sy * (j - 1);
    PlaneXY Px[3], Py[3], Pz[3];  // map planes of the current window (z nodes 0..NZ-1)
    Plane9 M[NF][3];              // apply: ring of 3x3 extrema of f_adv planes k-1, k, k+1
    const bool gathers = gij && ka < kb;
    const int kfirst = APPLYISH ? kc0 : ka, klast = APPLYISH ? kc1 : kb;

    // One cell: z node n of its window is Px/Py/Pz[n]; the newest plane is loaded here, the others are carried.
    // (Unrolling the k loop by the ring period instead of moving the planes down was measured: 3-6x the code,
    // 10 % slower -- profiles/r2_march_variants.md.)
    auto body = [&](int k) {
        constexpr int S0 = 0, S1 = 1, S2 = NZ == 3 ? 2 : 0;
        constexpr int SN = NZ - 1;                                           // slot the new plane goes to
        const int idx = col + fplane * k;
        const bool gk = !APPLYISH || (gathers && k >= ka && k < kb);
        float sum[NS], val[NS];
#pragma unroll
        for (int f = 0; f < NS; ++f) sum[f] = val[f] = 0.f;
#if BMQ_MARCH_PF_STREAM
        // the operands that are streamed rather than gathered (the accumulate target, the error kernel's init) are
        // compulsory DRAM misses at the end of the cell: ask for the next cell's lines now (ncu: the consumer of the
        // centre sample shared their scoreboard and waited ~670 cycles; error -5 %, accumulate -1..4 %).  The same
        // for the apply kernel's next f_adv plane (BMQ_MARCH_PF_STREAM == 2) made that kernel 3-6 % slower.
        if (k + 1 < klast) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                if (MODE == GM_CUMULATE) prefetch_line(a.out[f] + idx + fplane);
                if (MODE == GM_ERROR) prefetch_line(a.aux[f] + idx + fplane);
                if (BMQ_MARCH_PF_STREAM == 2 && MODE == GM_APPLY && k + 2 < fk) {
                    prefetch_line(a.aux[f] + idx + 2 * fplane);
                    if (threadIdx.y == 0) prefetch_line(a.aux[f] + idx + 2 * fplane - fi);
                    if (threadIdx.y == BMQ_MARCH_BY - 1) prefetch_line(a.aux[f] + idx + 2 * fplane + fi);
                }
            }
        }
#endif
        if (gk) {
            // BMQ_MARCH_EARLY_CENTRE: with a power-of-two h and three window planes the centre sample's position is the
            // centre-line value of the CARRIED middle plane, so its gathers can be issued before this cell's new map
            // plane is even requested and travel while it loads
            constexpr bool EARLY = BMQ_MARCH_EARLY_CENTRE && P2 && NZ == 3;
            int zi = 0, oc = 0;
#if BMQ_MARCH_EARLY_CENTRE
            if (EARLY) {
                const float qcx = to_cells<P2>(clampf(Px[1].c * ps, lo, hix), ox, DX * 0.5f, h, g.inv_h);
                const float qcy = to_cells<P2>(clampf(Py[1].c * ps, lo, hiy), oy, DY * 0.5f, h, g.inv_h);
                const float qcz = to_cells<P2>(clampf(Pz[1].c * ps, lo, hiz), oz, DZ * 0.5f, h, g.inv_h);
                oc = gather_one<NS>(a.src, fi, fplane, qcx, qcy, qcz, val, zi);
            }
#endif
            const int onew = mbase + sz * (NZ == 3 ? k + 1 : k);
            Px[SN] = plane_xy<P2, STAG>(m.x + onew, sy, ax, ay);
            Py[SN] = plane_xy<P2, STAG>(m.y + onew, sy, ax, ay);
            Pz[SN] = plane_xy<P2, STAG>(m.z + onew, sy, ax, ay);
            if (BMQ_MARCH_PREFETCH && k + 1 < kb) {
                // the plane the NEXT cell of this column will load: rows j-1..j+1 (lanes cover i-1..i+30, +1 the rest)
#pragma unroll
                for (int y = 0; y < (STAG == 2 ? 2 : 3); ++y) {
                    prefetch_line(m.x + onew + sz + y * sy + 1);
                    prefetch_line(m.y + onew + sz + y * sy + 1);
                    prefetch_line(m.z + onew + sz + y * sy + 1);
                }
            }
            const float cz = DZ ? fmaf(h, (float)k, oz) : __fmul_rn(h, (float)k);
            if (!P2) az.init(cz, h, g.inv_h);
            float2 px[4], py[4], pz[4];
            float ccx = 0.f, ccy = 0.f, ccz = 0.f;
            z_combine<P2, STAG>(Px[S0], Px[S1], Px[S2], az, ps, px, ccx);
            z_combine<P2, STAG>(Py[S0], Py[S1], Py[S2], az, ps, py, ccy);
            z_combine<P2, STAG>(Pz[S0], Pz[S1], Pz[S2], az, ps, pz, ccz);
            if (!P2) {
                const float3 c = sample_map<false>(m, g, cx, cy, cz);
                ccx = c.x; ccy = c.y; ccz = c.z;
            }
            float wgt[NS];
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int f = 0; f < NF; ++f)
                    wgt[c * NF + f] = MODE == GM_CUMULATE ? 0.125f * a.coeff[c] : APPLYISH ? 0.125f * -0.5f : 0.125f;
            // corner samples ii and ii + 4 as one packed pair; the x-minus halves wait for their turn in the sum
            float late[4][NS];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                float2 qx, qy, qz;
                qx = to_cells2<P2>(make_float2(clampf(px[ii].x, lo, hix), clampf(px[ii].y, lo, hix)), ox, DX * 0.5f, h, g.inv_h);
                qy = to_cells2<P2>(make_float2(clampf(py[ii].x, lo, hiy), clampf(py[ii].y, lo, hiy)), oy, DY * 0.5f, h, g.inv_h);
                qz = to_cells2<P2>(make_float2(clampf(pz[ii].x, lo, hiz), clampf(pz[ii].y, lo, hiz)), oz, DZ * 0.5f, h, g.inv_h);
                const Split2 spx = split2(qx), spy = split2(qy), spz = split2(qz);
                float s0[NS];
                gather_pair<NS>(a.src, fi, fplane, spx, spy, spz, s0, late[ii], hack_zero);
#pragma unroll
                for (int f = 0; f < NS; ++f) sum[f] = fmaf(wgt[f], s0[f], sum[f]);
            }
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                for (int f = 0; f < NS; ++f) sum[f] = fmaf(wgt[f], late[ii][f], sum[f]);
            if (!EARLY) {
                const float qcx = to_cells<P2>(clampf(ccx, lo, hix), ox, DX * 0.5f, h, g.inv_h);
                const float qcy = to_cells<P2>(clampf(ccy, lo, hiy), oy, DY * 0.5f, h, g.inv_h);
                const float qcz = to_cells<P2>(clampf(ccz, lo, hiz), oz, DZ * 0.5f, h, g.inv_h);
                oc = gather_one<NS>(a.src, fi, fplane, qcx, qcy, qcz, val, zi);
            }
            if (BMQ_MARCH_PREFETCH && k + 1 < kb && zi + BMQ_MARCH_PF_DIST < fk) {
                // the field plane the next cell's samples will newly touch: two planes above the centre sample's cell
#pragma unroll
                for (int f = 0; f < NS; ++f) {
                    prefetch_line(a.src[f] + oc + BMQ_MARCH_PF_DIST * fplane);
                    prefetch_line(a.src[f] + oc + BMQ_MARCH_PF_DIST * fplane + fi);
                    if (BMQ_MARCH_PF_ROWS == 3) prefetch_line(a.src[f] + oc + BMQ_MARCH_PF_DIST * fplane - fi);
                }
            }
        }
        if (MODE == GM_ADVECT) {
#pragma unroll
            for (int f = 0; f < NF; ++f) a.out[f][idx] = fmaf(0.5f, sum[f], 0.5f * val[f]);
        } else if (MODE == GM_ERROR) {
#pragma unroll
            for (int f = 0; f < NF; ++f) a.out[f][idx] = fmaf(0.5f, sum[f], 0.5f * val[f]) - __ldg(a.aux[f] + idx);
        } else if (MODE == GM_CUMULATE) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                float t = a.out[f][idx];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const float v = a.coeff[c] * val[c * NF + f];
                    t += fmaf(0.5f, sum[c * NF + f], 0.5f * v);
                }
                a.out[f][idx] = t;
            }
        } else if (MODE == GM_APPLY_NC) {
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                float r = __ldg(a.aux[f] + idx);
                if (gk) r += fmaf(0.5f, sum[f], 0.5f * (-0.5f * val[f]));
                a.out[f][idx] = r;
            }
        } else {
            constexpr int Q0 = 0, Q1 = 1, Q2 = 2;     // extrema of planes k-1, k, k+1
            const bool clamps = k > 0 && k < fk - 1;
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                M[f][Q2] = k + 1 < fk ? plane9(a.aux[f] + idx + fplane, fi) : M[f][Q1];
                float r = M[f][Q1].c;
                if (gk) r += fmaf(0.5f, sum[f], 0.5f * (-0.5f * val[f]));
                if (clamps) {
                    const float mx = fmaxf(fmaxf(M[f][Q0].mx, M[f][Q1].mx), M[f][Q2].mx);
                    const float mn = fminf(fminf(M[f][Q0].mn, M[f][Q1].mn), M[f][Q2].mn);
                    r = fminf(fmaxf(mn, r), mx);
                }
                a.out[f][idx] = r;
            }
        }
    };

    // carried planes at loop entry: the window of the first gathering cell minus its newest plane
    if (gathers) {
        const int o0 = mbase + sz * (ka - 1);
#pragma unroll
        for (int n = 0; n < NZ - 1; ++n) {
            Px[n] = plane_xy<P2, STAG>(m.x + o0 + n * sz, sy, ax, ay);
            Py[n] = plane_xy<P2, STAG>(m.y + o0 + n * sz, sy, ax, ay);
            Pz[n] = plane_xy<P2, STAG>(m.z + o0 + n * sz, sy, ax, ay);
        }
    }
    if (MODE == GM_APPLY) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            M[f][1] = plane9(a.aux[f] + col + fplane * kc0, fi);
            M[f][0] = kc0 > 0 ? plane9(a.aux[f] + col + fplane * (kc0 - 1), fi) : M[f][1];
        }
    }
#pragma unroll 1
    for (int k = kfirst; k < klast; ++k) {
        body(k);
        if (!APPLYISH || (gathers && k >= ka && k < kb)) {
            Px[0] = Px[1]; Py[0] = Py[1]; Pz[0] = Pz[1];
            if (NZ == 3) { Px[1] = Px[2]; Py[1] = Py[2]; Pz[1] = Pz[2]; }
        }
        if (MODE == GM_APPLY) {
#pragma unroll
            for (int f = 0; f < NF; ++f) { M[f][0] = M[f][1]; M[f][1] = M[f][2]; }
        }
    }
}

// ---- host side: grid shape and dispatch over (power-of-two h, staggering, pitch specialisation)
int march_fix_of(const Grid3 &g);         // kernels3d.cu: 0 or the compile-time plane extent
bool march_is_pow2_h(const Grid3 &g);

static inline int march_chunk(int fi, int fj, int nplanes)
{
    // planes per CTA: long columns amortise the two-plane prologue, but the grid must still fill the
    // machine (148 SMs x BMQ_MARCH_MINBLOCKS CTAs) a few times over
    const long long cols = (long long)((fi + 31) / 32) * ((fj + BMQ_MARCH_BY - 1) / BMQ_MARCH_BY);
    int kc = 32;
    while (kc > 8 && cols * ((nplanes + kc - 1) / kc) < 148ll * BMQ_MARCH_MINBLOCKS * 4) kc /= 2;
    return kc;
}

#define BMQ_MARCH_CASE(MODE, P2V, STAGV, NFV, NCHV, FIXV)                                                     \
    k_march<MODE, P2V, STAGV, NFV, NCHV, FIXV><<<gr, bl, BMQ_HACK_GATHER == 4 ? 4096 : 0, s>>>(g, r.kbeg, r.kend, kc, a, m)
#define BMQ_MARCH_FIX(MODE, P2V, STAGV, NFV, NCHV)                                                            \
    switch (fix) {                                                                                            \
    case 512: BMQ_MARCH_CASE(MODE, P2V, STAGV, NFV, NCHV, 512); break;                                        \
    case 256: BMQ_MARCH_CASE(MODE, P2V, STAGV, NFV, NCHV, 256); break;                                        \
    case 128: BMQ_MARCH_CASE(MODE, P2V, STAGV, NFV, NCHV, 128); break;                                        \
    default: BMQ_MARCH_CASE(MODE, P2V, STAGV, NFV, NCHV, 0); break;                                           \
    }
#define BMQ_MARCH_P2(MODE, STAGV, NFV, NCHV)                                                                  \
    if (p2) { BMQ_MARCH_FIX(MODE, true, STAGV, NFV, NCHV) } else { BMQ_MARCH_FIX(MODE, false, STAGV, NFV, NCHV) }

// one (NF, NCH) configuration over all staggerings; NF == 2 exists for centred fields only
template <int MODE, int NF, int NCH>
cudaError_t launch_march(cudaStream_t s, const Grid3 &g, KRange r, int stag, const MarchArgs<NF, NF * NCH> &a, const Map3 &m)
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    const int fi = g.ni + (stag == 1), fj = g.nj + (stag == 2);
    const int kc = march_chunk(fi, fj, r.kend - r.kbeg);
    const dim3 bl(32, BMQ_MARCH_BY, 1);
    const dim3 gr((fi + 31) / 32, (fj + BMQ_MARCH_BY - 1) / BMQ_MARCH_BY, (r.kend - r.kbeg + kc - 1) / kc);
    const int fix = march_fix_of(g);
    const bool p2 = march_is_pow2_h(g);
    if (NF == 2 && stag != 0) return cudaErrorInvalidValue;
    if (stag == 0) {
        BMQ_MARCH_P2(MODE, 0, NF, NCH)
    } else if constexpr (NF == 1) {
        if (stag == 1) { BMQ_MARCH_P2(MODE, 1, NF, NCH) }
        else if (stag == 2) { BMQ_MARCH_P2(MODE, 2, NF, NCH) }
        else { BMQ_MARCH_P2(MODE, 3, NF, NCH) }
    }
    return cudaGetLastError();
}

}  // namespace bmq
