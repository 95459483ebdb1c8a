/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the 3D BiMocq^2 advection hot path.
 *
 * This file is a plain-C restatement of the reference's CUDA kernels
 * (/root/reference/src/bimocq3D/GPU_kernel.cu:9-734).  It exists so that the hand-written
 * sm_100a kernels in gpufluidsimulation_b200/csrc can be checked on the same seeded inputs.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it; the product path never does.
 *
 * Pinning: the reference ships no tests or golden vectors, so this
 * restatement is pinned against the reference's own kernels executed on a B200
 * (oracle/_ref/libref3d.so, built unmodified from GPU_kernel.cu by oracle/Makefile;
 * tests/test_kernels_gpu.py) and against golden fixtures generated from that
 * run (tests/golden/).
 *
 * Arithmetic notes (all verified in the PTX nvcc 12.9 emits for the reference file):
 *  - lerp is written (1.0-c)*a + c*b (GPU_kernel.cu:22-25): c*b is a float product, (1.0-c)*a
 *    is a double product, the sum is a double fma rounded to float on return.
 *  - cell positions float(i)*h + origin are contracted to one fmaf by nvcc.
 *  - exp() on a float argument resolves to the float overload (expf).
 * Every function takes a global k-range [kbeg,kend): a z-slab rank passes virtual base
 * pointers (local pointer minus kz0*nx*ny) so that indices stay global.  The full domain is
 * kbeg=0, kend=nk+dimz.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

typedef struct { float x, y, z; } f3;

static inline f3 mk3(float x, float y, float z) { f3 r = {x, y, z}; return r; }

/* GPU_kernel.cu:9-12 */
static inline float clampf(float a, float lo, float hi) { return fminf(fmaxf(lo, a), hi); }

/* GPU_kernel.cu:14-20 */
static inline f3 clampv3(f3 in, f3 lo, f3 hi)
{
    return mk3(clampf(in.x, lo.x, hi.x), clampf(in.y, lo.y, hi.y), clampf(in.z, lo.z, hi.z));
}

/* GPU_kernel.cu:22-25; see arithmetic notes above */
static inline float lerp_ref(float a, float b, float c)
{
    float cb = c * b;
    return (float)fma(1.0 - (double)c, (double)a, (double)cb);
}

/* GPU_kernel.cu:27-41 */
static inline float trilerp_ref(float v000, float v001, float v010, float v011, float v100,
                                float v101, float v110, float v111, float a, float b, float c)
{
    return lerp_ref(lerp_ref(lerp_ref(v000, v001, a), lerp_ref(v010, v011, a), b),
                    lerp_ref(lerp_ref(v100, v101, a), lerp_ref(v110, v111, a), b), c);
}

/* GPU_kernel.cu:43-62 -- no bounds check, exactly like the reference */
static inline float sample_buffer(const float *b, int nx, int ny, int nz, float h, f3 off, f3 pos)
{
    (void)nz;
    float sx = pos.x - off.x, sy = pos.y - off.y, sz = pos.z - off.z;
    float qx = sx / h, qy = sy / h, qz = sz / h;
    int i = (int)floorf(qx), j = (int)floorf(qy), k = (int)floorf(qz);
    float fx = qx - (float)i, fy = qy - (float)j, fz = qz - (float)k;
    ptrdiff_t i0 = (ptrdiff_t)(i + nx * j + nx * ny * k);
    const float *p = b + i0;
    ptrdiff_t sy_ = nx, sz_ = (ptrdiff_t)nx * ny;
    return trilerp_ref(p[0], p[1], p[sy_], p[sy_ + 1], p[sz_], p[sz_ + 1], p[sz_ + sy_],
                       p[sz_ + sy_ + 1], fx, fy, fz);
}

/* GPU_kernel.cu:64-72 */
static inline f3 get_velocity(const float *u, const float *v, const float *w, float h, int nx,
                              int ny, int nz, f3 pos)
{
    float mh = (float)(-0.5 * (double)h);
    float _u = sample_buffer(u, nx + 1, ny, nz, h, mk3(mh, 0, 0), pos);
    float _v = sample_buffer(v, nx, ny + 1, nz, h, mk3(0, mh, 0), pos);
    float _w = sample_buffer(w, nx, ny, nz + 1, h, mk3(0, 0, mh), pos);
    return mk3(_u, _v, _w);
}

/* GPU_kernel.cu:74-90.  The float sum input + c1*v1 + c2*v2 + c3*v3 is contracted by nvcc
 * into a chain of three fmaf. */
static inline f3 trace_rk3(const float *u, const float *v, const float *w, float h, int ni, int nj,
                           int nk, float dt, f3 pos)
{
    float c1 = (float)(2.0 / 9.0 * (double)dt), c2 = (float)(3.0 / 9.0 * (double)dt),
          c3 = (float)(4.0 / 9.0 * (double)dt);
    f3 in = pos;
    f3 v1 = get_velocity(u, v, w, h, ni, nj, nk, in);
    double hd = 0.5 * (double)dt;
    f3 m1 = mk3((float)((double)in.x + hd * (double)v1.x), (float)((double)in.y + hd * (double)v1.y),
                (float)((double)in.z + hd * (double)v1.z));
    f3 v2 = get_velocity(u, v, w, h, ni, nj, nk, m1);
    double qd = 0.75 * (double)dt;
    f3 m2 = mk3((float)((double)in.x + qd * (double)v2.x), (float)((double)in.y + qd * (double)v2.y),
                (float)((double)in.z + qd * (double)v2.z));
    f3 v3 = get_velocity(u, v, w, h, ni, nj, nk, m2);
    f3 out = mk3(fmaf(c3, v3.x, fmaf(c2, v2.x, fmaf(c1, v1.x, in.x))),
                 fmaf(c3, v3.y, fmaf(c2, v2.y, fmaf(c1, v1.y, in.y))),
                 fmaf(c3, v3.z, fmaf(c2, v2.z, fmaf(c1, v1.z, in.z))));
    return clampv3(out, mk3(h, h, h),
                   mk3((float)ni * h - h, (float)nj * h - h, (float)nk * h - h));
}

/* GPU_kernel.cu:92-125 */
static inline f3 trace(const float *u, const float *v, const float *w, float h, int ni, int nj,
                       int nk, float cfldt, float dt, f3 pos)
{
    float sign = dt > 0 ? 1.0f : -1.0f;
    float T = dt > 0 ? dt : -dt;
    f3 opos = pos;
    float t = 0;
    float substep = cfldt;
    while (t < T) {
        if (t + substep > T) substep = T - t;
        opos = trace_rk3(u, v, w, h, ni, nj, nk, sign * substep, opos);
        t += substep;
    }
    return opos;
}

#define IDX3(i, j, k, nx, ny) ((ptrdiff_t)(i) + (ptrdiff_t)(nx) * (j) + (ptrdiff_t)(nx) * (ny) * (k))

/* forward_kernel, GPU_kernel.cu:127-144 (in place) */
void o3_forward(const float *u, const float *v, const float *w, float *xf, float *yf, float *zf,
                float h, int ni, int nj, int nk, float cfldt, float dt, int kbeg, int kend)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = kbeg; k < kend; k++)
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++) {
                if (i > 1 && i < ni - 2 && j > 1 && j < nj - 2 && k > 1 && k < nk - 2) {
                    ptrdiff_t idx = IDX3(i, j, k, ni, nj);
                    f3 p = trace(u, v, w, h, ni, nj, nk, cfldt, dt, mk3(xf[idx], yf[idx], zf[idx]));
                    xf[idx] = p.x; yf[idx] = p.y; zf[idx] = p.z;
                }
            }
}

/* clampExtrema_kernel, GPU_kernel.cu:146-167 */
void o3_clamp_extrema(const float *before, float *after, int ni, int nj, int nk, int kbeg, int kend)
{
#pragma omp parallel for schedule(static)
    for (int k = kbeg; k < kend; k++)
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++) {
                if (i > 0 && i < ni - 1 && j > 0 && j < nj - 1 && k > 0 && k < nk - 1) {
                    ptrdiff_t idx = IDX3(i, j, k, ni, nj);
                    float mx = before[idx], mn = before[idx];
                    for (int kk = k - 1; kk <= k + 1; kk++)
                        for (int jj = j - 1; jj <= j + 1; jj++)
                            for (int ii = i - 1; ii <= i + 1; ii++) {
                                float b = before[IDX3(ii, jj, kk, ni, nj)];
                                if (b > mx) mx = b;
                                if (b < mn) mn = b;
                            }
                    after[idx] = fminf(fmaxf(mn, after[idx]), mx);
                }
            }
}

/* one axis of GPU_kernel.cu:194-196; the Euler branch is contracted to fmaf(-v, s, p) */
static inline float dmc_axis(float p, float vel, float a, float s)
{
    if ((double)fabsf(a) > 1e-4) return p - (1 - expf(-a * s)) * vel / a;
    return fmaf(-vel, s, p);
}

/* DMC_backward_kernel, GPU_kernel.cu:169-204 */
void o3_dmc_backward(const float *u, const float *v, const float *w, const float *xin,
                     const float *yin, const float *zin, float *xout, float *yout, float *zout,
                     float h, int ni, int nj, int nk, float substep, int kbeg, int kend)
{
    const f3 zero = {0, 0, 0};
#pragma omp parallel for schedule(static)
    for (int k = kbeg; k < kend; k++)
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++) {
                if (i > 1 && i < ni - 2 && j > 1 && j < nj - 2 && k > 1 && k < nk - 2) {
                    ptrdiff_t idx = IDX3(i, j, k, ni, nj);
                    f3 pt = mk3(h * (float)i, h * (float)j, h * (float)k);
                    f3 vel = get_velocity(u, v, w, h, ni, nj, nk, pt);
                    f3 tp = mk3(vel.x > 0 ? pt.x - h : pt.x + h, vel.y > 0 ? pt.y - h : pt.y + h,
                                vel.z > 0 ? pt.z - h : pt.z + h);
                    f3 tv = get_velocity(u, v, w, h, ni, nj, nk, tp);
                    float ax = (vel.x - tv.x) / (pt.x - tp.x);
                    float ay = (vel.y - tv.y) / (pt.y - tp.y);
                    float az = (vel.z - tv.z) / (pt.z - tp.z);
                    f3 pn = mk3(dmc_axis(pt.x, vel.x, ax, substep), dmc_axis(pt.y, vel.y, ay, substep),
                                dmc_axis(pt.z, vel.z, az, substep));
                    xout[idx] = sample_buffer(xin, ni, nj, nk, h, zero, pn);
                    yout[idx] = sample_buffer(yin, ni, nj, nk, h, zero, pn);
                    zout[idx] = sample_buffer(zin, ni, nj, nk, h, zero, pn);
                }
            }
}

static inline f3 origin_of(float h, int dx, int dy, int dz)
{
    return mk3(-(float)dx * 0.5f * h, -(float)dy * 0.5f * h, -(float)dz * 0.5f * h);
}

/* semilag_kernel, GPU_kernel.cu:206-233 */
void o3_semilag(float *field, const float *src, const float *u, const float *v, const float *w,
                int dx, int dy, int dz, float h, int ni, int nj, int nk, float cfldt, float dt,
                int kbeg, int kend)
{
    f3 org = origin_of(h, dx, dy, dz);
    int fi = ni + dx, fj = nj + dy, fk = nk + dz;
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = kbeg; k < kend; k++)
        for (int j = 0; j < fj; j++)
            for (int i = 0; i < fi; i++) {
                if (i > 1 && i < fi - 2 - dx && j > 1 && j < fj - 2 - dy && k > 1 && k < fk - 2 - dz) {
                    f3 pt = mk3(fmaf(h, (float)i, org.x), fmaf(h, (float)j, org.y), fmaf(h, (float)k, org.z));
                    f3 pn = trace(u, v, w, h, ni, nj, nk, cfldt, dt, pt);
                    field[IDX3(i, j, k, fi, fj)] = sample_buffer(src, fi, fj, fk, h, org, pn);
                }
            }
}

/* the 8 sub-cell offsets of GPU_kernel.cu:317-322 (in units of 0.25*h) */
static const int VOL[8][3] = {{1, 1, 1}, {1, 1, -1}, {1, -1, 1}, {1, -1, -1},
                              {-1, 1, 1}, {-1, 1, -1}, {-1, -1, 1}, {-1, -1, -1}};

static inline f3 vol_off(int ii, float h, int is_point)
{
    if (is_point) return mk3(0, 0, 0);
    float q = 0.25f * h;
    return mk3(VOL[ii][0] > 0 ? q : -q, VOL[ii][1] > 0 ? q : -q, VOL[ii][2] > 0 ? q : -q);
}

static inline f3 map3(const float *mx, const float *my, const float *mz, int ni, int nj, int nk,
                      float h, f3 pos)
{
    const f3 zero = {0, 0, 0};
    return mk3(sample_buffer(mx, ni, nj, nk, h, zero, pos), sample_buffer(my, ni, nj, nk, h, zero, pos),
               sample_buffer(mz, ni, nj, nk, h, zero, pos));
}

/* advect_kernel, GPU_kernel.cu:312-374 */
void o3_advect(float *field, const float *field_init, const float *bx, const float *by,
               const float *bz, float h, int ni, int nj, int nk, int dx, int dy, int dz,
               int is_point, int kbeg, int kend)
{
    int ev = is_point ? 1 : 8;
    float weight = (float)(1.0 / (double)(float)ev);
    f3 org = origin_of(h, dx, dy, dz);
    int fi = ni + dx, fj = nj + dy, fk = nk + dz;
    f3 lo = mk3(h, h, h), hi = mk3(h * (float)ni - h, h * (float)nj - h, h * (float)nk - h);
#pragma omp parallel for schedule(static)
    for (int k = kbeg; k < kend; k++)
        for (int j = 0; j < fj; j++)
            for (int i = 0; i < fi; i++) {
                if (2 + dx < i && i < fi - 3 && 2 + dy < j && j < fj - 3 && 2 + dz < k && k < fk - 3) {
                    f3 c = mk3(fmaf(h, (float)i, org.x), fmaf(h, (float)j, org.y), fmaf(h, (float)k, org.z));
                    float sum = 0.0f;
                    for (int ii = 0; ii < ev; ii++) {
                        f3 o = vol_off(ii, h, is_point);
                        f3 pos = mk3(c.x + o.x, c.y + o.y, c.z + o.z);
                        f3 pi = clampv3(map3(bx, by, bz, ni, nj, nk, h, pos), lo, hi);
                        sum = fmaf(weight, sample_buffer(field_init, fi, fj, fk, h, org, pi), sum);
                    }
                    f3 pi = clampv3(map3(bx, by, bz, ni, nj, nk, h, c), lo, hi);
                    float value = sample_buffer(field_init, fi, fj, fk, h, org, pi);
                    field[IDX3(i, j, k, fi, fj)] = fmaf(0.5f, sum, 0.5f * value);
                }
            }
}

/* doubleAdvect_kernel, GPU_kernel.cu:236-310 */
void o3_double_advect(float *field, const float *temp_field, const float *bx, const float *by,
                      const float *bz, const float *bxp, const float *byp, const float *bzp, float h,
                      int ni, int nj, int nk, int dx, int dy, int dz, int is_point, float blend,
                      int kbeg, int kend)
{
    int ev = is_point ? 1 : 8;
    float weight = (float)(1.0 / (double)(float)ev);
    f3 org = origin_of(h, dx, dy, dz);
    int fi = ni + dx, fj = nj + dy, fk = nk + dz;
    f3 lo = mk3(h, h, h), hi = mk3(h * (float)ni - h, h * (float)nj - h, h * (float)nk - h);
#pragma omp parallel for schedule(static)
    for (int k = kbeg; k < kend; k++)
        for (int j = 0; j < fj; j++)
            for (int i = 0; i < fi; i++) {
                if (2 + dx < i && i < fi - 3 && 2 + dy < j && j < fj - 3 && 2 + dz < k && k < fk - 3) {
                    f3 c = mk3(fmaf(h, (float)i, org.x), fmaf(h, (float)j, org.y), fmaf(h, (float)k, org.z));
                    float sum = 0.0f;
                    for (int ii = 0; ii < ev; ii++) {
                        f3 o = vol_off(ii, h, is_point);
                        f3 pos = mk3(c.x + o.x, c.y + o.y, c.z + o.z);
                        f3 mid = clampv3(map3(bx, by, bz, ni, nj, nk, h, pos), lo, hi);
                        f3 fin = clampv3(map3(bxp, byp, bzp, ni, nj, nk, h, mid), lo, hi);
                        sum = fmaf(weight, sample_buffer(temp_field, fi, fj, fk, h, org, fin), sum);
                    }
                    f3 mid = clampv3(map3(bx, by, bz, ni, nj, nk, h, c), lo, hi);
                    f3 fin = clampv3(map3(bxp, byp, bzp, ni, nj, nk, h, mid), lo, hi);
                    float value = sample_buffer(temp_field, fi, fj, fk, h, org, fin);
                    float prev_value = 0.5f * (sum + value);
                    ptrdiff_t idx = IDX3(i, j, k, fi, fj);
                    field[idx] = fmaf(field[idx], blend, (1 - blend) * prev_value);
                }
            }
}

/* cumulate_kernel, GPU_kernel.cu:376-436.  sum = 0.5*sum + 0.5*value is a double expression
 * there (0.5 is a double literal) rounded to float on assignment. */
void o3_cumulate(const float *dfield, float *dfield_init, const float *mx, const float *my,
                 const float *mz, float h, int ni, int nj, int nk, int dx, int dy, int dz,
                 int is_point, float coeff, int kbeg, int kend)
{
    int ev = is_point ? 1 : 8;
    float weight = (float)(1.0 / (double)(float)ev);
    f3 org = origin_of(h, dx, dy, dz);
    int fi = ni + dx, fj = nj + dy, fk = nk + dz;
    f3 lo = mk3(0, 0, 0), hi = mk3(h * (float)ni, h * (float)nj, h * (float)nk);
    float wc = weight * coeff;
#pragma omp parallel for schedule(static)
    for (int k = kbeg; k < kend; k++)
        for (int j = 0; j < fj; j++)
            for (int i = 0; i < fi; i++) {
                if (1 + dx < i && i < fi - 2 && 1 + dy < j && j < fj - 2 && 1 + dz < k && k < fk - 2) {
                    f3 c = mk3(fmaf(h, (float)i, org.x), fmaf(h, (float)j, org.y), fmaf(h, (float)k, org.z));
                    float sum = 0.0f;
                    for (int ii = 0; ii < ev; ii++) {
                        f3 o = vol_off(ii, h, is_point);
                        f3 pos = mk3(c.x + o.x, c.y + o.y, c.z + o.z);
                        f3 mp = clampv3(map3(mx, my, mz, ni, nj, nk, h, pos), lo, hi);
                        sum = fmaf(wc, sample_buffer(dfield, fi, fj, fk, h, org, mp), sum);
                    }
                    f3 mp = clampv3(map3(mx, my, mz, ni, nj, nk, h, c), lo, hi);
                    float value = coeff * sample_buffer(dfield, fi, fj, fk, h, org, mp);
                    sum = (float)(0.5 * (double)sum + 0.5 * (double)value);
                    dfield_init[IDX3(i, j, k, fi, fj)] += sum;
                }
            }
}

/* compensate_kernel, GPU_kernel.cu:438-499 */
void o3_compensate(const float *src, const float *temp, float *test, const float *mx,
                   const float *my, const float *mz, float h, int ni, int nj, int nk, int dx, int dy,
                   int dz, int is_point, int kbeg, int kend)
{
    int ev = is_point ? 1 : 8;
    float weight = (float)(1.0 / (double)(float)ev);
    f3 org = origin_of(h, dx, dy, dz);
    int fi = ni + dx, fj = nj + dy, fk = nk + dz;
    f3 lo = mk3(0, 0, 0), hi = mk3(h * (float)ni, h * (float)nj, h * (float)nk);
#pragma omp parallel for schedule(static)
    for (int k = kbeg; k < kend; k++)
        for (int j = 0; j < fj; j++)
            for (int i = 0; i < fi; i++) {
                if (1 + dx < i && i < fi - 2 && 1 + dy < j && j < fj - 2 && 1 + dz < k && k < fk - 2) {
                    f3 c = mk3(fmaf(h, (float)i, org.x), fmaf(h, (float)j, org.y), fmaf(h, (float)k, org.z));
                    float sum = 0.0f;
                    for (int ii = 0; ii < ev; ii++) {
                        f3 o = vol_off(ii, h, is_point);
                        f3 pos = mk3(c.x + o.x, c.y + o.y, c.z + o.z);
                        f3 mp = clampv3(map3(mx, my, mz, ni, nj, nk, h, pos), lo, hi);
                        sum = fmaf(weight, sample_buffer(src, fi, fj, fk, h, org, mp), sum);
                    }
                    f3 mp = clampv3(map3(mx, my, mz, ni, nj, nk, h, c), lo, hi);
                    float value = sample_buffer(src, fi, fj, fk, h, org, mp);
                    sum = (float)(0.5 * (double)sum + 0.5 * (double)value);
                    ptrdiff_t idx = IDX3(i, j, k, fi, fj);
                    test[idx] = sum - temp[idx];
                }
            }
}

/* estimate_kernel, GPU_kernel.cu:501-537 */
void o3_estimate(float *dist, const float *x1, const float *y1, const float *z1, const float *x2,
                 const float *y2, const float *z2, float h, int ni, int nj, int nk, int kbeg, int kend)
{
#pragma omp parallel for schedule(static)
    for (int k = kbeg; k < kend; k++)
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++) {
                if (i > 1 && i < ni - 2 && j > 1 && j < nj - 2 && k > 1 && k < nk - 2) {
                    f3 pt = mk3(h * (float)i, h * (float)j, h * (float)k);
                    f3 b = map3(x1, y1, z1, ni, nj, nk, h, pt);
                    f3 f = map3(x2, y2, z2, ni, nj, nk, h, b);
                    float dbf = (pt.x - f.x) * (pt.x - f.x) + (pt.y - f.y) * (pt.y - f.y) +
                                (pt.z - f.z) * (pt.z - f.z);
                    f = map3(x2, y2, z2, ni, nj, nk, h, pt);
                    b = map3(x1, y1, z1, ni, nj, nk, h, f);
                    float dfb = (pt.x - b.x) * (pt.x - b.x) + (pt.y - b.y) * (pt.y - b.y) +
                                (pt.z - b.z) * (pt.z - b.z);
                    dist[IDX3(i, j, k, ni, nj)] = fmaxf(dbf, dfb);
                }
            }
}

/* add_kernel, GPU_kernel.cu:560-565 (with the bounds check the reference lacks) */
void o3_add(float *f1, const float *f2, float coeff, long n)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) f1[i] = fmaf(coeff, f2[i], f1[i]);
}

/* add_field_kernel, GPU_kernel.cu:878-883 */
void o3_add_field(float *out, const float *f1, const float *f2, float coeff, long n)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) out[i] = fmaf(coeff, f2[i], f1[i]);
}

/* max |f| over an array: one leg of getCFL, BimocqSolver.cpp:1093-1117 */
float o3_maxabs(const float *f, long n, float start)
{
    float m = start;
#pragma omp parallel for reduction(max : m) schedule(static)
    for (long i = 0; i < n; i++) {
        float a = fabsf(f[i]);
        if (a > m) m = a;
    }
    return m;
}

/* host-side max of the distortion buffer, Mapping.cpp:100-117 (boundary may be NULL = no solids) */
float o3_max_dist(const float *dist, const signed char *boundary, int ni, int nj, int nk)
{
    float m = 0.f;
    long n = (long)ni * nj * nk;
#pragma omp parallel for reduction(max : m) schedule(static)
    for (long i = 0; i < n; i++)
        if ((!boundary || boundary[i] != 2) && dist[i] > m) m = dist[i];
    return sqrtf(m);
}

/* ---- source terms, GPU_kernel.cu:736-876, 952-964 ------------------- */

/* emit_smoke_velocity_kernel, :736-758; ni,nj,nk are the field's dimensions */
void o3_emit_velocity(float *field, float h, int ni, int nj, int nk, float cx, float cy, float cz, float radius,
                      float emiter)
{
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nk; k++)
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++) {
                if (!(i > 1 && i < ni - 2 && j > 1 && j < nj - 2 && k > 1 && k < nk - 2)) continue;
                float dx = (float)(((double)(float)i - 0.5) * (double)h - (double)cx);
                float dy = (float)j * h - cy, dz = (float)k * h - cz;
                float length = sqrtf(dx * dx + dy * dy + dz * dz);
                if (length < radius) {
                    float theta = acosf(dy / hypotf(dy, dz));
                    field[IDX3(i, j, k, ni, nj)] =
                        (float)((double)emiter * 0.06 * (1.0 + 0.01 * (double)cosf(8.0f * theta)));
                }
            }
}

/* emit_smoke_field_kernel, :760-780 */
void o3_emit_field(float *rho, float *T, float h, int ni, int nj, int nk, float cx, float cy, float cz, float radius,
                   float density, float temperature)
{
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nk; k++)
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++) {
                if (!(i > 1 && i < ni - 2 && j > 1 && j < nj - 2 && k > 1 && k < nk - 2)) continue;
                float dx = (float)i * h - cx, dy = (float)j * h - cy, dz = (float)k * h - cz;
                if (sqrtf(dx * dx + dy * dy + dz * dz) < radius) {
                    rho[IDX3(i, j, k, ni, nj)] = density;
                    T[IDX3(i, j, k, ni, nj)] = temperature;
                }
            }
}

/* add_buoyancy_kernel, :804-823, launched with nj+1 rows (:831): density/temperature are indexed
 * with the v-face index, exactly like the reference */
void o3_add_buoyancy(float *field, const float *density, const float *temperature, int ni, int njp1, int nk,
                     float alpha, float beta, float dt)
{
#pragma omp parallel for schedule(static)
    for (int k = 0; k < nk; k++)
        for (int j = 1; j < njp1; j++)
            for (int i = 0; i < ni; i++) {
                ptrdiff_t idx = IDX3(i, j, k, ni, njp1), idx1 = idx - ni;
                float inner = beta * (temperature[idx] + temperature[idx1]) - alpha * (density[idx] + density[idx1]);
                field[idx] += (float)(0.5 * (double)dt * (double)inner);
            }
}

/* diffuse_field_kernel, :834-853: one sweep */
void o3_diffuse_sweep(const float *field, const float *in, float *out, int ni, int nj, int nk, float coef)
{
#pragma omp parallel for schedule(static)
    for (int k = 1; k < nk - 1; k++)
        for (int j = 1; j < nj - 1; j++)
            for (int i = 1; i < ni - 1; i++) {
                ptrdiff_t q = IDX3(i, j, k, ni, nj), sy = ni, sz = (ptrdiff_t)ni * nj;
                float s = in[q - 1] + in[q + 1];
                s += in[q - sy]; s += in[q + sy]; s += in[q - sz]; s += in[q + sz];
                out[q] = fmaf(coef, s, field[q]) / fmaf(coef, 6.0f, 1.0f);
            }
}

/* mad_kernel, :952-957 */
void o3_mad(float *field, const float *f1, const float *f2, float c1, float c2, long n)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; i++) field[i] = fmaf(c1, f1[i], c2 * f2[i]);
}
