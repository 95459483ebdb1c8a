/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product library.
 *
 * Plain-C restatement of the reference's multigrid-corrected CG pressure projection,
 * gpu_multi_grid_conjugate_gradient (/root/reference/src/bimocq3D/GPU_kernel.cu:1784-1828) and
 * everything it launches, sweep by sweep and memset by memset, in the reference's order:
 *   divergence_kernel(double) :984-1001      gradient_kernel(double) :1003-1021
 *   calc_poisson_value :1047-1059            dot_vector :1086-1119 (with its float rounding and
 *   calc_sum :1134-1178                        its sharedMem[+3] / [+7] / [+11] / [+15] slips)
 *   calc_max :1192-1222                      update_residual_kernel :1250-1261
 *   update_x / update_dir / mul / add :1291-1343
 *   smoothing_jacobi(_kernel) :1445-1491     smoothing_conjugate_gradient, updateDir :1493-1513
 *   sample_buffer<T> + triLerp_t -> float lerp :1515-1548, :22-25
 *   restriction :1550-1599, prolongation :1611-1622, V_Cycle (live branch) :1634-1712
 * Compiled with -ffp-contract=off; nvcc's contractions are written out as fma().
 *
 * Pinned on a B200 against the reference's own kernels (oracle/_ref/libref3d.so,
 * tests/test_projection_gpu.py) and against golden outputs of those kernels
 * (tests/golden/ref3d_projection.npz).  One deliberate difference: reads past the end of a
 * level's array (prolongation from a level with (n-1)/2 cells to an even n, :1620) return 0
 * instead of whatever follows the allocation. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int ni, nj, nk, number;
    double alpha, beta;
    double *b, *x, *r;
} lvl_t;

static inline double sum6(const double *x, long q, long sy, long sz)
{
    return ((((x[q - 1] + x[q + 1]) + x[q - sy]) + x[q + sy]) + x[q - sz]) + x[q + sz];
}
static inline double poisson_at(const double *x, long q, long sy, long sz) { return fma(-x[q], 6.0, sum6(x, q, sy, sz)); }

static void residual(double *r, const double *b, const double *x, int ni, int nj, int nk)
{
#pragma omp parallel for schedule(static)
    for (int k = 1; k < nk - 1; k++)
        for (int j = 1; j < nj - 1; j++)
            for (int i = 1; i < ni - 1; i++) {
                long q = i + (long)ni * (j + (long)nj * k);
                r[q] = b[q] - poisson_at(x, q, ni, (long)ni * nj);
            }
}

static void poisson(const double *x, double *out, int ni, int nj, int nk)
{
#pragma omp parallel for schedule(static)
    for (int k = 1; k < nk - 1; k++)
        for (int j = 1; j < nj - 1; j++)
            for (int i = 1; i < ni - 1; i++) {
                long q = i + (long)ni * (j + (long)nj * k);
                out[q] = poisson_at(x, q, ni, (long)ni * nj);
            }
}

static double calc_max(const double *v, long n)
{
    double m = 0.0;
    for (long e = 0; e < n; e++)
        if (v[e] > m) m = v[e];
    return m;
}

/* dot_vector<double>: one partial per 256 elements */
static void dot_partials(const double *v0, const double *v1, double *out, long count)
{
    long nblk = (count + 255) / 256;
#pragma omp parallel for schedule(static)
    for (long blk = 0; blk < nblk; blk++) {
        double s[272];
        for (int t = 0; t < 256; t++) {
            long e = blk * 256 + t;
            s[t] = e < count ? v0[e] * v1[e] : 0.0;
        }
        for (int t = 0; t < 16; t++) {
            double a = s[t * 16];
            for (int m = 1; m < 16; m++) a += s[t * 16 + m];
            s[256 + t] = (double)(float)a;                 /* `float sum0` */
        }
        double a = 0.0;
        for (int m = 0; m < 16; m++) {
            double term = (m & 3) == 3 ? s[m] : s[256 + m];   /* sharedMem[+3], [+7], [+11], [+15] */
            a = m == 0 ? term : a + term;
        }
        out[blk] = (double)(float)a;                       /* `float sum` */
    }
}

/* calc_sum<double> over `count` partials, 256 chains of `cpt` */
static double calc_sum(const double *v, long count, long cpt)
{
    double s[272];
    for (int t = 0; t < 256; t++) {
        double a = 0.0;
        for (long q = 0; q < cpt; q++)
            if (t * cpt + q < count) a += v[t * cpt + q];
        s[t] = a;
    }
    for (int t = 0; t < 16; t++) {
        double a = s[t * 16];
        for (int m = 1; m < 16; m++) a += s[t * 16 + m];
        s[256 + t] = a;
    }
    double a = s[256];
    for (int m = 1; m < 16; m++) a += s[256 + m];
    return a;
}

static double dot(const double *v0, const double *v1, double *partials, long number)
{
    long nblk = (number + 255) / 256;
    dot_partials(v0, v1, partials, number);
    return calc_sum(partials, nblk, (nblk + 255) / 256);
}

/* smoothing_jacobi<double>: iter (made even) sweeps ping-ponging x <-> tmp, result in x */
static void jacobi(double *x, const double *b, double *tmp, double alpha, double beta, int ni, int nj, int nk, int iter)
{
    if (iter % 2 == 1) iter += 1;
    double *in = x, *out = tmp;
    for (int it = 0; it < iter; it++) {
#pragma omp parallel for schedule(static)
        for (int k = 1; k < nk - 1; k++)
            for (int j = 1; j < nj - 1; j++)
                for (int i = 1; i < ni - 1; i++) {
                    long q = i + (long)ni * (j + (long)nj * k);
                    out[q] = fma(alpha, b[q], sum6(in, q, ni, (long)ni * nj)) * beta;
                }
        double *t = in; in = out; out = t;
    }
}

/* the float lerp of GPU_kernel.cu:22-25 */
static inline float lerp_f(float a, float b, float c) { return (float)fma(1.0 - (double)c, (double)a, (double)(c * b)); }

static inline float tri_f(const double *b, int nx, int ny, long number, int i, int j, int k, float fx, float fy, float fz)
{
    long o = i + (long)nx * j + (long)nx * ny * k, sy = nx, sz = (long)nx * ny;
#define AT(q) ((q) < number ? (float)b[(q)] : 0.f)
    float v000 = AT(o), v001 = AT(o + 1), v010 = AT(o + sy), v011 = AT(o + sy + 1);
    float v100 = AT(o + sz), v101 = AT(o + sz + 1), v110 = AT(o + sz + sy), v111 = AT(o + sz + sy + 1);
#undef AT
    return lerp_f(lerp_f(lerp_f(v000, v001, fx), lerp_f(v010, v011, fx), fy),
                  lerp_f(lerp_f(v100, v101, fx), lerp_f(v110, v111, fx), fy), fz);
}

static void restriction(const double *r, double *coarse, int ni, int nj, int nk, int ci, int cj, int ck)
{
    long number = (long)ni * nj * nk;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < ck; k++)
        for (int j = 0; j < cj; j++)
            for (int i = 0; i < ci; i++) {
                double acc = 0.0;
                for (int m = 0; m < 8; m++) {
                    int ox = (m >> 2) & 1, oy = (m >> 1) & 1, oz = m & 1;
                    double v = (double)tri_f(r, ni, nj, number, 2 * i + ox, 2 * j + oy, 2 * k + oz, 0.5f, 0.5f, 0.5f);
                    acc = m == 0 ? v : acc + v;
                }
                coarse[i + (long)ci * (j + (long)cj * k)] = acc / 8.0;
            }
}

static void prolongation(double *x, const double *coarse, int ni, int nj, int nk, int ci, int cj, int ck)
{
#pragma omp parallel for schedule(static)
    for (int k = 1; k < nk - 1; k++)
        for (int j = 1; j < nj - 1; j++)
            for (int i = 1; i < ni - 1; i++) {
                float px = (float)((double)((float)i / 2.f) - 0.5), py = (float)((double)((float)j / 2.f) - 0.5),
                      pz = (float)((double)((float)k / 2.f) - 0.5);
                int c_i = (int)floorf(px), c_j = (int)floorf(py), c_k = (int)floorf(pz);
                float fx = (float)((double)px - (double)(float)c_i), fy = (float)((double)py - (double)(float)c_j),
                      fz = (float)((double)pz - (double)(float)c_k);
                x[i + (long)ni * (j + (long)nj * k)] += (double)tri_f(coarse, ci, cj, (long)ci * cj * ck, c_i, c_j, c_k, fx, fy, fz);
            }
}

static void v_cycle(const double *b, double *x, double *res, lvl_t *L, double *temp0, int nlev)
{
    double scale[16];
    for (int i = 0; i < 16; i++) scale[i] = 1.0;
    scale[1] = 8.0;
    size_t n0 = sizeof(double) * (size_t)L[0].number;
    memcpy(L[0].b, res, n0);
    for (int i = 0; i < nlev - 1; i++) {
        memset(temp0, 0, n0);
        memset(L[i].x, 0, sizeof(double) * (size_t)L[i].number);
        jacobi(L[i].x, L[i].b, temp0, L[i].alpha * scale[i], L[i].beta, L[i].ni, L[i].nj, L[i].nk, 32);
        residual(L[i].r, L[i].b, L[i].x, L[i].ni, L[i].nj, L[i].nk);
        restriction(L[i].r, L[i + 1].b, L[i].ni, L[i].nj, L[i].nk, L[i + 1].ni, L[i + 1].nj, L[i + 1].nk);
    }
    int c = nlev - 1;
    memset(temp0, 0, n0);
    memset(L[c].x, 0, sizeof(double) * (size_t)L[c].number);
    jacobi(L[c].x, L[c].b, temp0, L[c].alpha * scale[c], L[c].beta, L[c].ni, L[c].nj, L[c].nk, 32);
    for (int i = nlev - 2; i >= 0; --i) {
        prolongation(L[i].x, L[i + 1].x, L[i].ni, L[i].nj, L[i].nk, L[i + 1].ni, L[i + 1].nj, L[i + 1].nk);
        memset(temp0, 0, n0);
        jacobi(L[i].x, L[i].b, temp0, L[i].alpha * scale[i], L[i].beta, L[i].ni, L[i].nj, L[i].nk, 4);
    }
    for (long e = 0; e < L[0].number; e++) x[e] = fma(L[0].x[e], 1.0, x[e]);
    residual(res, b, x, L[0].ni, L[0].nj, L[0].nk);
}

/* u, v, w: face fields, updated in place.  div, p, dir, res: ni*nj*nk doubles (div/res/dir rings
 * keep their incoming values, like the reference's buffers).  result: 4096 doubles.
 * Returns 0, or -1 when a level would be smaller than 3 cells. */
int o3_mgpcg(float *u, float *v, float *w, double *div, double *p, double *dir, double *res, double *result, int ni,
             int nj, int nk, int nlev, int iter, double halfrdx)
{
    lvl_t L[16];
    if (nlev < 1 || nlev > 16) return -1;
    long number = (long)ni * nj * nk;
    for (int i = 0; i < nlev; i++) {
        L[i].ni = i ? (L[i - 1].ni - 1) / 2 : ni;
        L[i].nj = i ? (L[i - 1].nj - 1) / 2 : nj;
        L[i].nk = i ? (L[i - 1].nk - 1) / 2 : nk;
        if (L[i].ni < 3 || L[i].nj < 3 || L[i].nk < 3) return -1;
        L[i].number = L[i].ni * L[i].nj * L[i].nk;
        L[i].alpha = -1.0;
        L[i].beta = 1.0 / 6.0;
        L[i].b = calloc(L[i].number, sizeof(double));
        L[i].x = calloc(L[i].number, sizeof(double));
        L[i].r = calloc(L[i].number, sizeof(double));
    }
    double *temp0 = calloc(number, sizeof(double)), *temp1 = calloc(number, sizeof(double));

#pragma omp parallel for schedule(static)
    for (int k = 0; k < nk; k++)
        for (int j = 0; j < nj; j++)
            for (int i = 0; i < ni; i++) {
                double ul = u[(long)k * (ni + 1) * nj + (long)j * (ni + 1) + i], ur = u[(long)k * (ni + 1) * nj + (long)j * (ni + 1) + i + 1];
                double vf = v[(long)k * ni * (nj + 1) + (long)j * ni + i], vb = v[(long)k * ni * (nj + 1) + (long)(j + 1) * ni + i];
                double wd = w[(long)k * ni * nj + (long)j * ni + i], wu = w[(long)(k + 1) * ni * nj + (long)j * ni + i];
                div[i + (long)ni * (j + (long)nj * k)] = halfrdx * (((ur - ul) + (vb - vf)) + (wu - wd));
            }
    memset(p, 0, sizeof(double) * number);
    residual(res, div, p, ni, nj, nk);
    for (long e = 0; e < number; e++) dir[e] = res[e] * 1.0;
    result[2000] = calc_max(res, number);
    result[0] = dot(res, res, temp0, number);
    for (int it = 0; it < iter; it++) {
        int off = it * 2;
        poisson(dir, temp0, ni, nj, nk);
        result[off + 1] = dot(dir, temp0, temp1, number);
        for (long e = 0; e < number; e++) p[e] += dir[e] * result[off] / result[off + 1];
        residual(res, div, p, ni, nj, nk);
        v_cycle(div, p, res, L, temp0, nlev);
        result[2001 + it] = calc_max(res, number);
        result[off + 2] = dot(res, res, temp0, number);
        for (long e = 0; e < number; e++) dir[e] = res[e] + dir[e] * result[off + 2] / result[off];
    }
    /* gradient_kernel x3 */
    for (int c = 0; c < 3; c++) {
        int dx = c == 0, dy = c == 1, dz = c == 2;
        int fi = ni + dx, fj = nj + dy, fk = nk + dz;
        float *f = c == 0 ? u : c == 1 ? v : w;
#pragma omp parallel for schedule(static)
        for (int k = 2; k < nk; k++)
            for (int j = 2; j < nj; j++)
                for (int i = 2; i < ni; i++) {
                    double p0 = p[(long)k * nj * ni + (long)j * ni + i], p1 = p[(long)(k - dz) * nj * ni + (long)(j - dy) * ni + i - dx];
                    long q = i + (long)fi * (j + (long)fj * k);
                    f[q] = f[q] - (float)(halfrdx * (p0 - p1));
                }
        (void)fk;
    }
    for (int i = 0; i < nlev; i++) { free(L[i].b); free(L[i].x); free(L[i].r); }
    free(temp0); free(temp1);
    return 0;
}
