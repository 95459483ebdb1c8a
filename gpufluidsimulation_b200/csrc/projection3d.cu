// Pressure projection: the multigrid-preconditioned conjugate
// gradient solver the reference's device-resident solver calls right after advection
// (BimocqGPUSolver.cpp:406-467 -> gpuMapper::projectionMultiGrid, GPU_Advection.h:622-626 ->
// gpu_multi_grid_conjugate_gradient, GPU_kernel.cu:1784-1828), with the same prototype and the
// same results bit for bit, plus a handle API (bmq_mgpcg_*) that owns the fp64 work buffers.
//
// What has to be reproduced exactly (all fp64, deterministic):
//   * the iteration itself (GPU_kernel.cu:1784-1828): one CG step on (dir, residual), one
//     V-cycle correction, a new direction; the scalars live in tempResult[2*i .. 2*i+2],
//     residual maxima in tempResult[2000 + i];
//   * the reference's reduction trees, including their quirks: dot_vector (:1088-1119) rounds
//     its two partial stages to float and reads four raw products instead of four group sums
//     (sharedMem[+3], [+7], [+11], [+15]); calc_sum (:1134-1178) sums 256 sequential chains;
//   * the V-cycle (:1634-1712): 32 Jacobi sweeps going down, 4 going up, alpha scaled by 8 on
//     level 1 only, restriction / prolongation through the *float* lerp (:22-25, every nested
//     lerp rounds to float).
//
// B200 design: the solve is HBM-bound fp64 stencil work and the Jacobi smoother is ~2/3 of
// its traffic.  k_jacobi_tb applies K = 4 (or 2) sweeps per pass: a CTA owns an x-y tile,
// streams along z, keeps K time levels in flight (xy-neighbours through double-buffered shared
// memory planes, z-neighbours in registers), so each pass reads x and b once and writes once
// instead of K times.  Every cell value is produced by the same fp64 expression from the same
// operands as in the reference's sweep-by-sweep kernels, hence bit-identical.  The reference's
// memsets (two level-0-sized ones per level) disappear: the first pass of a smoothing run knows
// x = 0, and every pass writes the zero Dirichlet ring itself.  (Tried and rejected on B200: a
// 2 x 2-cells-per-thread variant with parity-split shared planes -- 24 B instead of 40 B of shared
// traffic per update but 128 registers, 16 warps/SM and 42 % issue utilisation: 35 % slower; two
// cells per thread at 1024 threads: register spills, 34 % slower.  ncu: the pass as it stands keeps
// the shared-memory pipe 85 % busy (profiles/r1_k_jacobi_tb_ncu_details.txt).)  Single-CTA reductions of the
// reference (calc_max over all cells) become grid-wide ones (max is exact in any order).
#include <algorithm>
#include <cstdint>
#include <vector>

#include "common.h"
#include <mutex>

namespace bmq {
void count_launches(unsigned n);   // kernels3d.cu
}

namespace {

using bmq::count_launches;

// same layout as the reference's SCoarseLevelInfo (GPU_Advection.h:13-24) == bmq_coarse_level
typedef bmq_coarse_level Lvl;

__device__ __forceinline__ double sum6(double xl, double xr, double xf, double xb, double xd, double xu)
{
    return __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(xl, xr), xf), xb), xd), xu);
}
// calc_poisson_value, GPU_kernel.cu:1047-1059: (sum of the six neighbours) - 6 x_c, contracted by
// nvcc into one fma
__device__ __forceinline__ double poisson_at(const double *x, int idx, int sy, int sz)
{
    const double s = sum6(x[idx - 1], x[idx + 1], x[idx - sy], x[idx + sy], x[idx - sz], x[idx + sz]);
    return __fma_rn(-x[idx], 6.0, s);
}
// one Jacobi update, smoothing_jacobi_kernel(double), GPU_kernel.cu:1445-1463
__device__ __forceinline__ double jacobi_value(double s, double alpha, double b, double beta)
{
    return __dmul_rn(__fma_rn(alpha, b, s), beta);
}
// the reference's float lerp (GPU_kernel.cu:22-25), `(1.0-c)*a + c*b` = float product c*b, double
// fma, rounding to float.  For the fractions that occur here (0 and 1/2) 1-c is exact in fp32, and
// rounding the exact (1-c)*a + (c*b) once (fmaf) equals rounding it to 53 and then to 24 bits
// (innocuous double rounding, 53 >= 2*24+2) -- so plain fp32 gives the same bits without the
// quarter-rate double<->float conversions.
__device__ __forceinline__ float lerp_f(float a, float b, float c)
{
    return __fmaf_rn(__fsub_rn(1.0f, c), a, __fmul_rn(c, b));
}

#define P3_IJK(fi, fj, fk)                                        \
    const int i = blockIdx.x * 32 + threadIdx.x;                  \
    const int j = blockIdx.y * 4 + threadIdx.y;                   \
    const int k = blockIdx.z;                                     \
    if (i >= (fi) || j >= (fj) || k >= (fk)) return;              \
    const int index = i + (fi) * (j + (fj) * k);

dim3 pblk() { return dim3(32, 4, 1); }
dim3 pgrd(int fi, int fj, int fk) { return dim3((fi + 31) / 32, (fj + 3) / 4, fk); }
int sblocks(size_t n, int per_thread = 1)
{
    size_t b = (n / per_thread + 255) / 256;
    return (int)std::max<size_t>(1, std::min<size_t>(b, 148 * 16));
}

// divergence_kernel(double), GPU_kernel.cu:984-1001
__global__ void __launch_bounds__(128)
k_divergence(const float *__restrict__ u, const float *__restrict__ v, const float *__restrict__ w, double *div, int ni,
             int nj, int nk, double halfrdx)
{
    P3_IJK(ni, nj, nk)
    const double ul = u[k * (ni + 1) * nj + j * (ni + 1) + i], ur = u[k * (ni + 1) * nj + j * (ni + 1) + i + 1];
    const double vf = v[k * ni * (nj + 1) + j * ni + i], vb = v[k * ni * (nj + 1) + (j + 1) * ni + i];
    const double wd = w[k * ni * nj + j * ni + i], wu = w[(k + 1) * ni * nj + j * ni + i];
    div[index] = __dmul_rn(halfrdx, __dadd_rn(__dadd_rn(__dsub_rn(ur, ul), __dsub_rn(vb, vf)), __dsub_rn(wu, wd)));
}

// gradient_kernel(double p), GPU_kernel.cu:1003-1021 (fi,fj,fk = the face field's dimensions)
__global__ void __launch_bounds__(128)
k_gradient(float *field, const double *__restrict__ p, int fi, int fj, int fk, int dimx, int dimy, int dimz, double halfrdx)
{
    P3_IJK(fi, fj, fk)
    const int pi = fi - dimx, pj = fj - dimy, pk = fk - dimz;
    if (!(i > 1 && i < pi && j > 1 && j < pj && k > 1 && k < pk)) return;
    const double p0 = p[k * pj * pi + j * pi + i];
    const double p1 = p[(k - dimz) * pj * pi + (j - dimy) * pi + i - dimx];
    field[index] = __fsub_rn(field[index], __double2float_rn(__dmul_rn(halfrdx, __dsub_rn(p0, p1))));
}

// update_residual_kernel(double), GPU_kernel.cu:1250-1261: r = b - A x on interior cells.  With
// WITH_MAX the kernel also folds max(r) over ALL cells (ring cells keep and contribute their old
// value) into *maxbits, replacing the reference's single-CTA calc_max (:1224-1235) -- exact in
// any order.
template <bool WITH_MAX>
__global__ void __launch_bounds__(128)
k_residual(double *r, const double *__restrict__ b, const double *__restrict__ x, int ni, int nj, int nk,
           unsigned long long *maxbits)
{
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    double val = 0.0;
    if (i < ni && j < nj && k < nk) {
        const int index = i + ni * (j + nj * k);
        if (i > 0 && i < ni - 1 && j > 0 && j < nj - 1 && k > 0 && k < nk - 1) {
            val = __dsub_rn(b[index], poisson_at(x, index, ni, ni * nj));
            r[index] = val;
        } else if (WITH_MAX) {
            val = r[index];
        }
    }
    if (WITH_MAX) {
        double m = val > 0.0 ? val : 0.0;   // calc_max starts from 0 and skips NaN (fmax semantics)
        for (int o = 16; o; o >>= 1) {
            const double other = __shfl_xor_sync(0xffffffffu, m, o);
            m = other > m ? other : m;
        }
        __shared__ double wm[4];
        if (threadIdx.x == 0) wm[threadIdx.y] = m;
        __syncthreads();
        if (threadIdx.x == 0 && threadIdx.y == 0) {
            m = fmax(fmax(wm[0], wm[1]), fmax(wm[2], wm[3]));
            // m >= 0: bit order == value order.  Most CTAs lose against the current maximum: test first
            if (m > 0.0 && m > __longlong_as_double((long long)*(volatile unsigned long long *)maxbits))
                atomicMax(maxbits, (unsigned long long)__double_as_longlong(m));
        }
    }
}

// calc_poisson_kernel(double), GPU_kernel.cu:1074-1084: out = A x on interior cells
__global__ void __launch_bounds__(128)
k_poisson(const double *__restrict__ x, double *out, int ni, int nj, int nk)
{
    P3_IJK(ni, nj, nk)
    if (!(i > 0 && i < ni - 1 && j > 0 && j < nj - 1 && k > 0 && k < nk - 1)) return;
    out[index] = poisson_at(x, index, ni, ni * nj);
}

// dot_vector<double>, GPU_kernel.cu:1086-1119: one partial per 256 consecutive elements.  One warp
// per reference block: products go to padded shared memory, lanes 0..15 form the sixteen group
// sums (left to right, in double, then rounded to float as the reference's `float sum0` does),
// lane 0 adds them in the reference's order -- which takes elements 3, 7, 11 and 15 of the block
// (raw products) where group sums 3, 7, 11, 15 were meant -- and rounds to float again.
__global__ void __launch_bounds__(256)
k_dot(const double *__restrict__ v0, const double *__restrict__ v1, double *out, int count, int nref)
{
    __shared__ double prod[8][272 + 16];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *s = prod[wid];
    for (int blk = blockIdx.x * 8 + wid; blk < nref; blk += gridDim.x * 8) {
        const int base = blk * 256;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int e = lane + 32 * m, idx = base + e;
            s[e + (e >> 4)] = idx < count ? __dmul_rn(v0[idx], v1[idx]) : 0.0;
        }
        __syncwarp();
        float sum0 = 0.f;
        if (lane < 16) {
            const double *g = s + 17 * lane;
            double a = g[0];
#pragma unroll
            for (int q = 1; q < 16; ++q) a = __dadd_rn(a, g[q]);
            sum0 = __double2float_rn(a);
        }
        // gather: lane 0 needs group sums 0,1,2,4,5,6,8,9,10,12,13,14 and products 3,7,11,15
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float gq = __shfl_sync(0xffffffffu, sum0, q);
            const double term = (q & 3) == 3 ? s[q] : (double)gq;   // s[q]: q < 16, no padding shift
            acc = q == 0 ? term : __dadd_rn(acc, term);
        }
        if (lane == 0) out[blk] = (double)__double2float_rn(acc);
        __syncwarp();
    }
}

// calc_sum<double>, GPU_kernel.cu:1134-1178: 256 sequential chains of `cpt` partials, then a
// 16 x 16 tree in double.  The reference runs it in ONE CTA with strided, uncoalesced reads; here
// every chain gets its own warp (coalesced 256-byte reads, the 32 values broadcast in order by
// shuffles so that the adds happen in the reference's order), and a second tiny kernel does the tree.
__global__ void __launch_bounds__(128) k_sum_chains(const double *__restrict__ v, double *chain_sums, int count, int cpt)
{
    const int lane = threadIdx.x & 31, chain = blockIdx.x * 4 + (threadIdx.x >> 5);   // 64 CTAs x 4 warps = 256 chains
    const long long beg = (long long)chain * cpt;
    double a = 0.0;
    double nxt = (lane < cpt && beg + lane < count) ? v[beg + lane] : 0.0;
    for (int q = 0; q < cpt; q += 32) {
        const double cur = nxt;
        const int qn = q + 32 + lane;
        nxt = (qn < cpt && beg + qn < count) ? v[beg + qn] : 0.0;
        const int lim = min(32, min(cpt - q, (int)max(0LL, min((long long)count - beg - q, 32LL))));
        if (lim == 32) {
#pragma unroll
            for (int m = 0; m < 32; ++m) a = __dadd_rn(a, __shfl_sync(0xffffffffu, cur, m));
        } else {
            for (int m = 0; m < lim; ++m) a = __dadd_rn(a, __shfl_sync(0xffffffffu, cur, m));
        }
    }
    if (lane == 0) chain_sums[chain] = a;
}
__global__ void __launch_bounds__(32) k_sum_tree(const double *__restrict__ chain_sums, double *out, int slot)
{
    const int t = threadIdx.x;
    double g = 0.0;
    if (t < 16) {
        g = chain_sums[t * 16];
#pragma unroll
        for (int m = 1; m < 16; ++m) g = __dadd_rn(g, chain_sums[t * 16 + m]);
    }
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const double gm = __shfl_sync(0xffffffffu, g, m);
        acc = m == 0 ? gm : __dadd_rn(acc, gm);
    }
    if (t == 0) out[slot] = acc;
}

// update_x_kernel(double), GPU_kernel.cu:1291-1298:  x += dir * alpha[r] / alpha[d]
// and update_residual-free variants below; scalars are read from the device result array.
__global__ void __launch_bounds__(256)
k_update_x(double *x, const double *__restrict__ dir, const double *__restrict__ res, size_t n, int rI, int dI)
{
    const double ar = res[rI], ad = res[dI];
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
        x[e] = __dadd_rn(x[e], __ddiv_rn(__dmul_rn(dir[e], ar), ad));
}
// update_dir_kernel(double), GPU_kernel.cu:1309-1316:  dir = residual + dir * beta[r+] / beta[r]
__global__ void __launch_bounds__(256)
k_update_dir(double *dir, const double *__restrict__ r, const double *__restrict__ res, size_t n, int rI, int rPlusI)
{
    const double bp = res[rPlusI], br = res[rI];
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
        dir[e] = __dadd_rn(r[e], __ddiv_rn(__dmul_rn(dir[e], bp), br));
}
// add_kernel(double), GPU_kernel.cu:1336-1343 with coef = 1:  a += b
__global__ void __launch_bounds__(256) k_add(double *a, const double *__restrict__ b, size_t n)
{
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x)
        a[e] = __dadd_rn(a[e], b[e]);
}
__global__ void k_store_max(double *res, int slot, const unsigned long long *maxbits)
{
    res[slot] = __longlong_as_double((long long)*maxbits);
}

// ---- transfer operators ---------------------------------------------------------------------
// trilinear sample of a double array through the float lerp, fractions fx,fy,fz in {0, 0.5}
// (sample_buffer<T> + triLerp_t, GPU_kernel.cu:1529-1548: triLerp_t calls the float `lerp`).
// Reads beyond `number` return 0 (the reference reads whatever follows the allocation).
__device__ __forceinline__ float tri_f(const double *__restrict__ b, int nx, int ny, int number, int i, int j, int k,
                                       float fx, float fy, float fz)
{
    const int o = i + nx * j + nx * ny * k;
    auto at = [&](int idx) -> float { return idx < number ? __double2float_rn(b[idx]) : 0.f; };
    const float v000 = at(o), v001 = at(o + 1), v010 = at(o + nx), v011 = at(o + nx + 1);
    const float v100 = at(o + nx * ny), v101 = at(o + nx * ny + 1), v110 = at(o + nx * ny + nx), v111 = at(o + nx * ny + nx + 1);
    return lerp_f(lerp_f(lerp_f(v000, v001, fx), lerp_f(v010, v011, fx), fy),
                  lerp_f(lerp_f(v100, v101, fx), lerp_f(v110, v111, fx), fy), fz);
}

// restriction<double>, GPU_kernel.cu:1550-1599: average of the eight samples at 2c + {0.5, 1.5};
// writes ALL coarse cells.
__global__ void __launch_bounds__(128)
k_restrict(const double *__restrict__ r, double *coarse, int ni, int nj, int nk, int ci, int cj, int ck)
{
    P3_IJK(ci, cj, ck)
    // the eight samples sit at the centres of the eight 2x2x2 sub-blocks of the fine 3x3x3 block at
    // (2i, 2j, 2k); all fractions are 1/2.  Load the 27 values once (rounded to float as the
    // reference's float lerp does on entry); indices stay inside the fine array (2c+2 <= n-1).
    float f[3][3][3];
    const double *base = r + (2 * i + (size_t)ni * (2 * j + (size_t)nj * (2 * k)));
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int b_ = 0; b_ < 3; ++b_)
#pragma unroll
            for (int a = 0; a < 3; ++a) f[c][b_][a] = __double2float_rn(base[a + (size_t)ni * (b_ + (size_t)nj * c)]);
    double acc = 0.0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        // value0..7: z varies fastest, then y, then x (:1568-1590)
        const int ox = (m >> 2) & 1, oy = (m >> 1) & 1, oz = m & 1;
        const float v = lerp_f(lerp_f(lerp_f(f[oz][oy][ox], f[oz][oy][ox + 1], 0.5f), lerp_f(f[oz][oy + 1][ox], f[oz][oy + 1][ox + 1], 0.5f), 0.5f),
                               lerp_f(lerp_f(f[oz + 1][oy][ox], f[oz + 1][oy][ox + 1], 0.5f), lerp_f(f[oz + 1][oy + 1][ox], f[oz + 1][oy + 1][ox + 1], 0.5f), 0.5f),
                               0.5f);
        acc = m == 0 ? (double)v : __dadd_rn(acc, (double)v);
    }
    coarse[index] = __ddiv_rn(acc, 8.0);
}

// prolongation_kernel(double), GPU_kernel.cu:1611-1622: x += sample(coarse, (i/2 - 0.5, ...)) on
// interior cells
__global__ void __launch_bounds__(128)
k_prolong(double *x, const double *__restrict__ coarse, int ni, int nj, int nk, int ci, int cj, int ck)
{
    P3_IJK(ni, nj, nk)
    if (!(i > 0 && i < ni - 1 && j > 0 && j < nj - 1 && k > 0 && k < nk - 1)) return;
    // i odd: position (i-1)/2 exactly, fraction 0; i even: (i-2)/2 + 0.5
    const int c_i = (i - 1) >> 1, c_j = (j - 1) >> 1, c_k = (k - 1) >> 1;
    const float fx = (i & 1) ? 0.f : 0.5f, fy = (j & 1) ? 0.f : 0.5f, fz = (k & 1) ? 0.f : 0.5f;
    const float v = tri_f(coarse, ci, cj, ci * cj * ck, c_i, c_j, c_k, fx, fy, fz);
    x[index] = __dadd_rn(x[index], (double)v);
}

// ---- temporally blocked Jacobi smoother -------------------------------------------------------
// K sweeps of  x <- ((sum of six neighbours) + alpha b) beta  on interior cells (ring = 0) in one
// pass.  CTA = TX x TY threads on an x-y tile with a K-cell apron, marching along z; time level
// t+1 lags level t by one plane.  Per thread: the last two planes of every level (its own
// column) and the b values of the last K planes in registers; per level one double-buffered
// shared-memory plane for the four x-y neighbours.  One __syncthreads per plane.
template <int K, int TX, int TY, bool ZERO_IN>
__global__ void __launch_bounds__(TX *TY, 1)
k_jacobi_tb(const double *__restrict__ xin, const double *__restrict__ b, double *__restrict__ xout, double alpha,
            double beta, int ni, int nj, int nk, int zchunk)
{
    constexpr int OX = TX - 2 * K, OY = TY - 2 * K, PX = TX + 2, PLANE = (TX + 2) * (TY + 2);
    extern __shared__ double sm[];   // [K][2][TY+2][TX+2], borders stay zero
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int gx = blockIdx.x * OX - K + tx, gy = blockIdx.y * OY - K + ty;
    const bool in_dom = gx >= 0 && gx < ni && gy >= 0 && gy < nj;
    const bool in_xy = gx > 0 && gx < ni - 1 && gy > 0 && gy < nj - 1;
    const bool owner = in_dom && tx >= K && tx < TX - K && ty >= K && ty < TY - K;
    const int zc0 = blockIdx.z * zchunk, zc1 = min(zc0 + zchunk, nk);
    const int zs = max(zc0 - K, 0), ze = zc1 - 1 + K;
    const size_t plane = (size_t)ni * nj;
    const size_t col = in_dom ? (size_t)gx + (size_t)ni * gy : 0;

    for (int e = ty * TX + tx; e < K * 2 * PLANE; e += TX * TY) sm[e] = 0.0;
    __syncthreads();

    // lv[slot][t]: level t on the last planes; the three slots change roles with the iteration
    // (the z loop is unrolled 3x so that this is register renaming, not moves).  bq[t]: b on plane z-1-t.
    double lv[3][K], bq[K];
#pragma unroll
    for (int t = 0; t < K; ++t) lv[0][t] = lv[1][t] = lv[2][t] = bq[t] = 0.0;

    double nx = 0.0, nb = 0.0;
    if (in_dom) {
        if (!ZERO_IN) nx = xin[col + plane * zs];
        nb = b[col + plane * zs];
    }
    double *my = sm + (ty + 1) * PX + (tx + 1);
    int wro = 0, rdo = PLANE;   // offsets of the write / read halves of the double buffer
    // iterations past ze (at most two) only produce planes nobody stores
    for (int zb = zs; zb <= ze; zb += 3) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int z = zb + r;
            const int NEWS = r, OLD = (r + 1) % 3;   // slot that receives plane z-t; slot holding plane z-2-t
            double up = nx;   // level t at plane z - t
            const double curb = nb;
            nx = 0.0;
            nb = 0.0;
            if (in_dom && z + 1 < nk) {
                if (!ZERO_IN) nx = xin[col + plane * (z + 1)];
                nb = b[col + plane * (z + 1)];
            }
#pragma unroll
            for (int t = 0; t < K; ++t) {
                my[t * 2 * PLANE + wro] = up;
                const double *s = my + t * 2 * PLANE + rdo;   // level t, plane z-1-t
                const double sum = sum6(s[-1], s[1], s[-PX], s[PX], lv[OLD][t], up);
                const int p = z - 1 - t;
                const double val = (in_xy && p > 0 && p < nk - 1) ? jacobi_value(sum, alpha, bq[t], beta) : 0.0;
                lv[NEWS][t] = up;
                up = val;
            }
#pragma unroll
            for (int t = K - 1; t > 0; --t) bq[t] = bq[t - 1];
            bq[0] = curb;
            const int po = z - K;
            if (owner && po >= zc0 && po < zc1) xout[col + plane * po] = up;
            const int tmp = wro; wro = rdo; rdo = tmp;
            __syncthreads();
        }
    }
}

// ---- host side --------------------------------------------------------------------------------
struct Mg {
    cudaStream_t st;
    double *scratch;               // device: [0] = bits of the running max, [1..256] = calc_sum chain sums
    unsigned long long *maxbits;   // == scratch[0]
    int status = BMQ_OK;

    void ck(cudaError_t e, const char *what)
    {
        if (status == BMQ_OK && e != cudaSuccess) status = bmq::check_cuda(e, what, __FILE__, __LINE__);
    }
    void post(const char *what, unsigned n = 1)
    {
        count_launches(n);
        ck(cudaGetLastError(), what);
    }

    template <int K, int TX, int TY>
    void jacobi_pass(const double *in, const double *b, double *out, double alpha, double beta, int ni, int nj, int nk, bool zero_in)
    {
        constexpr int OX = TX - 2 * K, OY = TY - 2 * K;
        const int tiles = ((ni + OX - 1) / OX) * ((nj + OY - 1) / OY);
        // z-chunks: one CTA per SM is resident (148 SMs); pick the chunk count that wastes least
        // between the last partial wave and the K apron planes every chunk re-computes
        int nz = 1;
        double best = 0.0;
        for (int c = 1; c <= std::max(1, nk / (4 * K)); ++c) {
            const int zc = (nk + c - 1) / c, ctas = tiles * ((nk + zc - 1) / zc);
            const double eff = (double)ctas / (148.0 * ((ctas + 147) / 148)) * zc / (zc + 2.0 * K);
            if (eff > best + 1e-9) { best = eff; nz = c; }
        }
        const int zchunk = (nk + nz - 1) / nz;
        nz = (nk + zchunk - 1) / zchunk;
        dim3 grid((ni + OX - 1) / OX, (nj + OY - 1) / OY, nz), block(TX, TY);
        const size_t smem = sizeof(double) * K * 2 * (TX + 2) * (TY + 2);
        if (zero_in) {
            static bool once = (cudaFuncSetAttribute(k_jacobi_tb<K, TX, TY, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
            (void)once;
            k_jacobi_tb<K, TX, TY, true><<<grid, block, smem, st>>>(in, b, out, alpha, beta, ni, nj, nk, zchunk);
        } else {
            static bool once = (cudaFuncSetAttribute(k_jacobi_tb<K, TX, TY, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
            (void)once;
            k_jacobi_tb<K, TX, TY, false><<<grid, block, smem, st>>>(in, b, out, alpha, beta, ni, nj, nk, zchunk);
        }
        post("k_jacobi_tb");
    }

    // `sweeps` (made even, smoothing_jacobi :1472-1473) Jacobi sweeps starting from x (taken as 0
    // when x_is_zero); ping-pongs between x and tmp; returns the buffer holding the result
    double *smooth(double *x, double *tmp, const double *b, double alpha, double beta, int ni, int nj, int nk, int sweeps, bool x_is_zero)
    {
        if (sweeps & 1) ++sweeps;
        double *cur = x, *oth = tmp;
        while (sweeps > 0) {
            if (sweeps >= 4) {
                jacobi_pass<4, 32, 32>(cur, b, oth, alpha, beta, ni, nj, nk, x_is_zero);
                sweeps -= 4;
            } else {
                jacobi_pass<2, 32, 16>(cur, b, oth, alpha, beta, ni, nj, nk, x_is_zero);
                sweeps -= 2;
            }
            x_is_zero = false;
            std::swap(cur, oth);
        }
        return cur;
    }

    void residual(double *r, const double *b, const double *x, int ni, int nj, int nk, bool with_max)
    {
        if (with_max) {
            ck(cudaMemsetAsync(maxbits, 0, sizeof(*maxbits), st), "memset maxbits");
            k_residual<true><<<pgrd(ni, nj, nk), pblk(), 0, st>>>(r, b, x, ni, nj, nk, maxbits);
        } else {
            k_residual<false><<<pgrd(ni, nj, nk), pblk(), 0, st>>>(r, b, x, ni, nj, nk, nullptr);
        }
        post("k_residual");
    }
    void store_max(double *res, int slot)
    {
        k_store_max<<<1, 1, 0, st>>>(res, slot, maxbits);
        post("k_store_max");
    }
    // dot_vector_kernel + calc_sum_kernel: res[slot] = "v0 . v1"
    void dot(const double *v0, const double *v1, double *partials, double *res, int slot, int number)
    {
        const int nref = (number + 255) / 256, cpt = (nref + 255) / 256;
        k_dot<<<std::min((nref + 7) / 8, 148 * 8), 256, 0, st>>>(v0, v1, partials, number, nref);
        k_sum_chains<<<64, 128, 0, st>>>(partials, scratch + 1, nref, cpt);
        k_sum_tree<<<1, 32, 0, st>>>(scratch + 1, res, slot);
        post("k_dot/k_sum", 3);
    }

    // V_Cycle(b, x, residual, levels, temp0, tempResult, levelnum, offset), GPU_kernel.cu:1634-1712
    // (live branch).  Scratch (levels[].b/x/r, temp0) is not part of the result contract.
    void v_cycle(const double *b, double *x, double *residual_, const Lvl *L, double *temp0, int nlev, bool want_max)
    {
        std::vector<double *> xs(nlev);
        // the reference copies residual into levels[0].b (:1663); residual is only rewritten by the
        // last kernel of the cycle, so level 0 smooths against it in place
        auto rhs = [&](int i) -> const double * { return i == 0 ? residual_ : L[i].b; };
        auto scaled_alpha = [&](int i) { return i == 1 ? L[i].alpha * 8.0 : L[i].alpha * 1.0; };
        for (int i = 0; i < nlev - 1; ++i) {
            xs[i] = smooth(L[i].x, temp0, rhs(i), scaled_alpha(i), L[i].beta, L[i].ni, L[i].nj, L[i].nk, 32, true);
            if (xs[i] != L[i].x) {   // keep the level's x out of the shared scratch while coarser levels run
                ck(cudaMemcpyAsync(L[i].x, xs[i], sizeof(double) * L[i].number, cudaMemcpyDeviceToDevice, st), "copy x");
                xs[i] = L[i].x;
            }
            residual(L[i].r, rhs(i), xs[i], L[i].ni, L[i].nj, L[i].nk, false);
            k_restrict<<<pgrd(L[i + 1].ni, L[i + 1].nj, L[i + 1].nk), pblk(), 0, st>>>(L[i].r, L[i + 1].b, L[i].ni, L[i].nj, L[i].nk,
                                                                                     L[i + 1].ni, L[i + 1].nj, L[i + 1].nk);
            post("k_restrict");
        }
        const int c = nlev - 1;
        xs[c] = smooth(L[c].x, temp0, rhs(c), scaled_alpha(c), L[c].beta, L[c].ni, L[c].nj, L[c].nk, 32, true);
        for (int i = nlev - 2; i >= 0; --i) {
            k_prolong<<<pgrd(L[i].ni, L[i].nj, L[i].nk), pblk(), 0, st>>>(xs[i], xs[i + 1], L[i].ni, L[i].nj, L[i].nk, L[i + 1].ni,
                                                                         L[i + 1].nj, L[i + 1].nk);
            post("k_prolong");
            // the coarser result may sit in temp0: it has been consumed, temp0 is free again
            xs[i] = smooth(xs[i], xs[i] == temp0 ? L[i].x : temp0, rhs(i), scaled_alpha(i), L[i].beta, L[i].ni, L[i].nj, L[i].nk, 4, false);
        }
        k_add<<<sblocks(L[0].number), 256, 0, st>>>(x, xs[0], (size_t)L[0].number);
        post("k_add");
        residual(residual_, b, x, L[0].ni, L[0].nj, L[0].nk, want_max);
    }

    // gpu_multi_grid_conjugate_gradient, GPU_kernel.cu:1784-1828
    void solve(float *u, float *v, float *w, double *div, double *p, double *dir, double *residual_, double *temp0, double *temp1,
               double *res, const Lvl *L, int nlev, int iter, double halfrdx)
    {
        const int ni = L[0].ni, nj = L[0].nj, nk = L[0].nk, number = L[0].number;
        k_divergence<<<pgrd(ni, nj, nk), pblk(), 0, st>>>(u, v, w, div, ni, nj, nk, halfrdx);
        post("k_divergence");
        ck(cudaMemsetAsync(p, 0, sizeof(double) * number, st), "memset p");
        residual(residual_, div, p, ni, nj, nk, true);
        store_max(res, 2000);
        ck(cudaMemcpyAsync(dir, residual_, sizeof(double) * number, cudaMemcpyDeviceToDevice, st), "dir = r");   // mul_kernel(.., 1)
        dot(residual_, residual_, temp0, res, 0, number);
        for (int it = 0; it < iter && status == BMQ_OK; ++it) {
            const int off = it * 2;
            // smoothing_conjugate_gradient, :1494-1504
            k_poisson<<<pgrd(ni, nj, nk), pblk(), 0, st>>>(dir, temp0, ni, nj, nk);
            post("k_poisson");
            dot(dir, temp0, temp1, res, off + 1, number);
            k_update_x<<<sblocks(number), 256, 0, st>>>(p, dir, res, (size_t)number, off, off + 1);
            post("k_update_x");
            residual(residual_, div, p, ni, nj, nk, false);
            v_cycle(div, p, residual_, L, temp0, nlev, true);
            store_max(res, 2001 + it);
            // updateDir, :1506-1513
            dot(residual_, residual_, temp0, res, off + 2, number);
            k_update_dir<<<sblocks(number), 256, 0, st>>>(dir, residual_, res, (size_t)number, off, off + 2);
            post("k_update_dir");
        }
        k_gradient<<<pgrd(ni + 1, nj, nk), pblk(), 0, st>>>(u, p, ni + 1, nj, nk, 1, 0, 0, halfrdx);
        k_gradient<<<pgrd(ni, nj + 1, nk), pblk(), 0, st>>>(v, p, ni, nj + 1, nk, 0, 1, 0, halfrdx);
        k_gradient<<<pgrd(ni, nj, nk + 1), pblk(), 0, st>>>(w, p, ni, nj, nk + 1, 0, 0, 1, halfrdx);
        post("k_gradient", 3);
    }
};

bool levels_ok(const Lvl *L, int nlev)
{
    if (!L || nlev < 1 || nlev > 16) return false;
    for (int i = 0; i < nlev; ++i)
        if (L[i].ni < 3 || L[i].nj < 3 || L[i].nk < 3 || L[i].number != L[i].ni * L[i].nj * L[i].nk || !L[i].b || !L[i].x || !L[i].r) return false;
    return true;
}

}  // namespace

struct bmq_mgpcg {
    int ni, nj, nk, nlev;
    std::vector<Lvl> levels;
    std::vector<void *> allocs;
    double *div = nullptr, *p = nullptr, *dir = nullptr, *residual = nullptr, *temp0 = nullptr, *temp1 = nullptr, *result = nullptr;
    double *scratch = nullptr;
    cudaStream_t stream = 0;
};

extern "C" {

void gpu_multi_grid_conjugate_gradient(float *u, float *v, float *w, double *div, double *p, double *dir, double *residual,
                                       double *temp0, double *temp1, double *tempResult, bmq_coarse_level *levels, int levelNum,
                                       int iter, double halfrdx)
{
    if (!bmq::require_device()) return;
    if (!levels_ok(levels, levelNum)) {
        bmq::set_error(BMQ_ERR_ARG, "gpu_multi_grid_conjugate_gradient: bad level table");
        return;
    }
    // 257 doubles of device scratch per device, allocated once (legacy entry points: one host thread per device)
    static std::mutex mu;
    static double *per_device[64] = {};
    int dev = 0;
    BMQ_CKV(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { bmq::set_error(BMQ_ERR_ARG, "gpu_multi_grid_conjugate_gradient: device ordinal out of range"); return; }
    double *scratch = nullptr;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!per_device[dev]) BMQ_CKV(cudaMalloc(&per_device[dev], 257 * sizeof(double)));
        scratch = per_device[dev];
    }
    Mg mg{0, scratch, reinterpret_cast<unsigned long long *>(scratch)};
    mg.solve(u, v, w, div, p, dir, residual, temp0, temp1, tempResult, levels, levelNum, iter, halfrdx);
}

int bmq_mgpcg_create(int ni, int nj, int nk, int levels, bmq_mgpcg **out)
{
    if (!out) return bmq::set_error(BMQ_ERR_ARG, "bmq_mgpcg_create: null out");
    *out = nullptr;
    if (!bmq::require_device()) return BMQ_ERR_NODEVICE;
    if (ni < 3 || nj < 3 || nk < 3 || levels < 1 || levels > 16 || (double)ni * nj * nk > 2.0e9)
        return bmq::set_error(BMQ_ERR_ARG, "bmq_mgpcg_create: bad dimensions %dx%dx%d, %d levels", ni, nj, nk, levels);
    bmq_mgpcg *m = new bmq_mgpcg;
    m->ni = ni; m->nj = nj; m->nk = nk; m->nlev = levels;
    auto alloc = [&](size_t bytes, void **ptr) -> int {
        BMQ_CK(cudaMalloc(ptr, bytes));
        m->allocs.push_back(*ptr);
        BMQ_CK(cudaMemset(*ptr, 0, bytes));
        return BMQ_OK;
    };
    int st = BMQ_OK;
    // level table as BimocqGPUSolver's constructor builds it (BimocqGPUSolver.cpp:68-90)
    m->levels.resize(levels);
    for (int i = 0; i < levels && st == BMQ_OK; ++i) {
        Lvl &l = m->levels[i];
        l.ni = i ? (m->levels[i - 1].ni - 1) / 2 : ni;
        l.nj = i ? (m->levels[i - 1].nj - 1) / 2 : nj;
        l.nk = i ? (m->levels[i - 1].nk - 1) / 2 : nk;
        if (l.ni < 3 || l.nj < 3 || l.nk < 3) {
            st = bmq::set_error(BMQ_ERR_ARG, "bmq_mgpcg_create: level %d would be %dx%dx%d", i, l.ni, l.nj, l.nk);
            break;
        }
        l.number = l.ni * l.nj * l.nk;
        l.alpha = -1.0;
        l.beta = 1.0 / 6.0;
        const size_t bytes = sizeof(double) * l.number;
        if ((st = alloc(bytes, (void **)&l.b)) != BMQ_OK) break;
        if ((st = alloc(bytes, (void **)&l.x)) != BMQ_OK) break;
        if ((st = alloc(bytes, (void **)&l.r)) != BMQ_OK) break;
    }
    const size_t nb = sizeof(double) * (size_t)ni * nj * nk;
    double **bufs[] = {&m->div, &m->p, &m->dir, &m->residual, &m->temp0, &m->temp1};
    for (double **b : bufs)
        if (st == BMQ_OK) st = alloc(nb, (void **)b);
    if (st == BMQ_OK) st = alloc(sizeof(double) * 4096, (void **)&m->result);
    if (st == BMQ_OK) st = alloc(257 * sizeof(double), (void **)&m->scratch);
    if (st != BMQ_OK) {
        for (void *a : m->allocs) cudaFree(a);
        delete m;
        return st;
    }
    *out = m;
    return BMQ_OK;
}

void bmq_mgpcg_destroy(bmq_mgpcg *m)
{
    if (!m) return;
    for (void *a : m->allocs) cudaFree(a);
    delete m;
}

int bmq_mgpcg_set_stream(bmq_mgpcg *m, void *stream)
{
    if (!m) return bmq::set_error(BMQ_ERR_ARG, "bmq_mgpcg_set_stream: null handle");
    m->stream = (cudaStream_t)stream;
    return BMQ_OK;
}

int bmq_mgpcg_solve(bmq_mgpcg *m, float *u, float *v, float *w, int iter, double halfrdx)
{
    if (!m || !u || !v || !w) return bmq::set_error(BMQ_ERR_ARG, "bmq_mgpcg_solve: null argument");
    if (iter < 0 || iter > 1000) return bmq::set_error(BMQ_ERR_ARG, "bmq_mgpcg_solve: iter %d outside [0, 1000]", iter);
    Mg mg{m->stream, m->scratch, reinterpret_cast<unsigned long long *>(m->scratch)};
    mg.solve(u, v, w, m->div, m->p, m->dir, m->residual, m->temp0, m->temp1, m->result, m->levels.data(), m->nlev, iter, halfrdx);
    return mg.status;
}

int bmq_mgpcg_buffer(bmq_mgpcg *m, int which, double **ptr, long long *count)
{
    if (!m || !ptr) return bmq::set_error(BMQ_ERR_ARG, "bmq_mgpcg_buffer: null argument");
    const long long n = (long long)m->ni * m->nj * m->nk;
    switch (which) {
        case BMQ_MG_DIV: *ptr = m->div; break;
        case BMQ_MG_P: *ptr = m->p; break;
        case BMQ_MG_DIR: *ptr = m->dir; break;
        case BMQ_MG_RESIDUAL: *ptr = m->residual; break;
        case BMQ_MG_RESULT: *ptr = m->result; if (count) *count = 4096; return BMQ_OK;
        default: return bmq::set_error(BMQ_ERR_ARG, "bmq_mgpcg_buffer: unknown buffer %d", which);
    }
    if (count) *count = n;
    return BMQ_OK;
}

int bmq_mgpcg_levels(bmq_mgpcg *m, bmq_coarse_level *out, int capacity)
{
    if (!m || !out) return bmq::set_error(BMQ_ERR_ARG, "bmq_mgpcg_levels: null argument");
    if (capacity < m->nlev) return bmq::set_error(BMQ_ERR_ARG, "bmq_mgpcg_levels: capacity %d < %d", capacity, m->nlev);
    for (int i = 0; i < m->nlev; ++i) out[i] = m->levels[i];
    return m->nlev;
}

}  // extern "C"
