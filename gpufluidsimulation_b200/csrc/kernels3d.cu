// 3D BiMocq^2 advection kernels for sm_100a and their host launchers.
//
// Every kernel works on GLOBAL indices: field pointers are "virtual bases" (pointer to plane 0
// of the global grid, which for a z-slab rank is its local pointer minus first_plane*nx*ny) and
// the launch covers global planes [kbeg,kend).  Interior guards are the reference's, in global
// indices, so a slab rank computes exactly the values a single GPU would.
//
// Thread mapping: blockDim = (32, 4, 1); x -> i (coalesced rows), y -> j, blockIdx.z -> k.
// No integer div/mod per thread (the reference decodes a flat index with two of each).
#include "launch3d.h"
#include "device3d.cuh"

#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

namespace bmq {

static std::atomic<unsigned long long> g_launches{0};
static std::atomic<bool> g_fast_division{true};   // testing knob, see bmq_set_fast_division
static std::atomic<bool> g_tolerance{false};      // opt-in, see bmq_set_tolerance_mode
unsigned long long kernel_launch_count() { return g_launches.load(); }
static inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launches(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static inline bool exact_pow2_h(int ni, int nj, int nk, float h, bool tolerance)
{
    int e;
    float m = frexpf(h, &e);
    // i*h, i*h +- h/4, +- h/2 must all be exact: (4*n+4) < 2^24
    int nmax = ni > nj ? ni : nj;
    nmax = nmax > nk ? nmax : nk;
    // tolerance mode: every cell size takes the kernels written for a power of two (positions in grid units by one
    // multiplication with RN(1/h), node shortcuts, constant window fractions); exact only when h IS a power of two
    return (m == 0.5f || tolerance) && (long long)nmax * 4 + 8 < (1ll << 24);
}
static inline bool is_pow2_h(const Grid3 &g) { return g.p2 != 0; }     // decided when the grid was made (make_grid)

// ---- exhaustive check of div_h (device3d.cuh) for one divisor: every float p in [0, p_max]
// (div_h takes the three-instruction path for p == 0 and p >= 2^-100, which is what gets verified)
__global__ void __launch_bounds__(256) k_verify_div(float h, float inv_h, unsigned last_bits, unsigned *mismatches)
{
    unsigned bad = 0;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= last_bits;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float p = __uint_as_float((unsigned)b);       // tiny p: div_h itself falls back to IEEE division there
        bad += __float_as_uint(div_h(p, h, inv_h)) != __float_as_uint(__fdiv_rn(p, h));
    }
    if (bad) atomicAdd(mismatches, bad);
}

// true when div_h reproduces IEEE division by h for every float in [0, p_max] (the sequence is odd in p, so
// negative positions are covered).  Checked once per (h, range) on the current device: ~1e9 quotients, a few ms.
bool division_verified(float h, float p_max)
{
    static std::mutex mu;
    static std::map<unsigned, std::pair<float, bool>> cache;    // bits of h -> (verified range, outcome)
    if (!g_fast_division.load(std::memory_order_relaxed)) return false;      // testing knob: IEEE division everywhere
    unsigned key;
    memcpy(&key, &h, sizeof key);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end() && (it->second.first >= p_max || !it->second.second)) return it->second.second;
    bool ok = false;
    unsigned *d_bad = nullptr, h_bad = 1, last;
    memcpy(&last, &p_max, sizeof last);
    if (cudaMalloc(&d_bad, sizeof(unsigned)) == cudaSuccess) {
        if (cudaMemset(d_bad, 0, sizeof(unsigned)) == cudaSuccess) {
            k_verify_div<<<148 * 8, 256>>>(h, 1.0f / h, last, d_bad);
            if (cudaMemcpy(&h_bad, d_bad, sizeof(unsigned), cudaMemcpyDeviceToHost) == cudaSuccess) ok = h_bad == 0;
        }
        cudaFree(d_bad);
    }
    if (cudaGetLastError() != cudaSuccess) ok = false;      // no device, launch failure: IEEE division, loudly elsewhere
    else cache[key] = std::make_pair(p_max, ok);
    return ok;
}

Grid3 make_grid(int ni, int nj, int nk, float h)
{
    Grid3 g;
    g.ni = ni; g.nj = nj; g.nk = nk; g.h = h; g.inv_h = 1.0f / h;
    g.p2 = exact_pow2_h(ni, nj, nk, h, g_tolerance.load(std::memory_order_relaxed)) ? 1 : 0;
    if (!is_pow2_h(g)) {      // (tolerance mode: p2 is set and inv_h stays the plain reciprocal)
        // positions stay inside the clamped domain plus one DMC reach; verify four times the domain
        int nmax = ni > nj ? ni : nj;
        nmax = nmax > nk ? nmax : nk;
        if (!division_verified(h, 4.0f * (float)(nmax + 8) * h)) g.inv_h = -g.inv_h;
    }
    return g;
}
void set_fast_division(bool on) { g_fast_division.store(on); }
void set_tolerance_mode(bool on) { g_tolerance.store(on); }
bool tolerance_mode() { return g_tolerance.load(); }
bool division_is_fast(float h, int nmax) { return make_grid(nmax, nmax, nmax, h).inv_h > 0.f; }

// CTA shape (32, BMQ_BY, BMQ_BZ); default 32x4x1 = 128 threads.  A CTA that spans several z-planes shares the
// k-1/k/k+1 planes of the map windows and of the near-identity field gathers in L1.
#ifndef BMQ_BY
#define BMQ_BY 4
#endif
#ifndef BMQ_BZ
#define BMQ_BZ 1
#endif
static inline dim3 block3() { return dim3(32, BMQ_BY, BMQ_BZ); }
static inline dim3 grid3(int fi, int fj, KRange r)
{
    return dim3((fi + 31) / 32, (fj + BMQ_BY - 1) / BMQ_BY, (r.kend - r.kbeg + BMQ_BZ - 1) / BMQ_BZ);
}
// planes per CTA of the column kernels (BMQ_COLUMN): long enough for L1 reuse along z, short enough that the grid
// still fills 148 SMs x 8 CTAs a few times over; BMQ_COLUMN_CHUNK overrides (1 = one cell per thread)
static int column_chunk(int fi, int fj, int nplanes)
{
    static const int forced = [] { const char *e = getenv("BMQ_COLUMN_CHUNK"); return e ? atoi(e) : 0; }();
    if (forced > 0) return forced;
    const long long cols = (long long)((fi + 31) / 32) * ((fj + BMQ_BY - 1) / BMQ_BY);
    int kc = 16;
    while (kc > 1 && cols * ((nplanes + kc - 1) / kc) < 148ll * 8 * 4) kc /= 2;
    return kc;
}
static inline dim3 grid3_columns(int fi, int fj, KRange r, int kc)
{
    return dim3((fi + 31) / 32, (fj + BMQ_BY - 1) / BMQ_BY, (r.kend - r.kbeg + kc - 1) / kc);
}

#define BMQ_IJK(fi, fj)                                          \
    const int i = blockIdx.x * 32 + threadIdx.x;                 \
    const int j = blockIdx.y * BMQ_BY + threadIdx.y;             \
    const int k = kbeg + blockIdx.z * BMQ_BZ + threadIdx.z;      \
    if (i >= (fi) || j >= (fj) || k >= kend_) return;

// Column form of the same index space: a thread owns cells (i, j, kc0..kc1) and walks them in z, so that the planes
// k-1, k, k+1 its gathers touch are still in L1 when the next cell needs them (a CTA per plane re-fetched them from
// L2 for every plane).  The map-update and distortion kernels use it; kchunk = 1 is the old one-cell-per-thread form.
// launch bounds of the three column kernels (CTAs of 128 threads): BMQ_MAPK_MINBLOCKS caps the registers per thread
#ifndef BMQ_MAPK_MINBLOCKS
#define BMQ_MAPK_BOUNDS __launch_bounds__(256)
#else
#define BMQ_MAPK_BOUNDS __launch_bounds__(128, BMQ_MAPK_MINBLOCKS)
#endif
#define BMQ_COLUMN(fi, fj)                                       \
    const int i = blockIdx.x * 32 + threadIdx.x;                 \
    const int j = blockIdx.y * BMQ_BY + threadIdx.y;             \
    const int kc0 = kbeg + blockIdx.z * kchunk;                  \
    const int kc1 = min(kc0 + kchunk, kend_);

// ------------------------------------------------------------------ forward map (a4)
// FIX != 0: the x and y extents are the compile-time constant FIX, so that row and plane pitches of
// every gather become immediates of the load instructions instead of 64-bit address arithmetic.

// forward_kernel, GPU_kernel.cu:127-144: psi <- trace(psi, +dt) in place, NMAP mappers at once
// (each mapper's particle is independent; tracing two per thread doubles the loads in flight).
template <bool P2, int NMAP, int FIX = 0>
__global__ void BMQ_MAPK_BOUNDS
k_forward(Grid3 g_, int kbeg, int kend_, int kchunk, Vel3 vel, MapSetRW<NMAP> maps, float cfldt, float dt)
{
    const Grid3 g = fix_grid<FIX>(g_);
    BMQ_COLUMN(g.ni, g.nj)
    if (!(i > 1 && i < g.ni - 2 && j > 1 && j < g.nj - 2)) return;
#pragma unroll 1
    for (int k = max(kc0, 2); k < min(kc1, g.nk - 2); ++k) {
        const int idx = i + g.ni * (j + g.nj * k);
        float3 p[NMAP];
#pragma unroll
        for (int m = 0; m < NMAP; ++m) p[m] = make_float3(maps.x[m][idx], maps.y[m][idx], maps.z[m][idx]);
        trace_multi<P2, NMAP>(vel, g, cfldt, dt, p);
#pragma unroll
        for (int m = 0; m < NMAP; ++m) {
            maps.x[m][idx] = p[m].x;
            maps.y[m][idx] = p[m].y;
            maps.z[m][idx] = p[m].z;
        }
    }
}

// ------------------------------------------------------------------ DMC backward sub-step (a5)
// DMC_backward_kernel, GPU_kernel.cu:169-204.  The back-traced point depends only on the cell
// and the velocity, so both mappers (velocity + scalar) are updated from one evaluation.
__device__ __forceinline__ float dmc_axis(float p, float v, float a, float s)
{
    // The reference's expression, literally (GPU_kernel.cu:194-196: float-vs-double comparison, `exp`
    // resolving to expf, int literal 1), so that nvcc/ptxas make the same contraction choices as in
    // the reference binary.  A hand-contracted form (explicit fmaf in the Euler branch) differed by
    // one ulp in rare cells with |v| ~ 1e-6 (found by tests/test_gpusolver_frame_gpu.py).
    return (fabs(a) > 1e-4) ? p - (1 - exp(-a * s)) * v / a : p - v * s;
}

template <bool P2, int NMAP, int FIX = 0>
__global__ void BMQ_MAPK_BOUNDS
k_dmc(Grid3 g_, int kbeg, int kend_, int kchunk, Vel3 vel, MapSetRO<NMAP> in, MapSetRW<NMAP> out, float substep)
{
    const Grid3 g = fix_grid<FIX>(g_);
    BMQ_COLUMN(g.ni, g.nj)
    if (!(i > 1 && i < g.ni - 2 && j > 1 && j < g.nj - 2)) return;
    const float h = g.h;
    const float px = h * (float)i, py = h * (float)j;
#pragma unroll 1
    for (int k = max(kc0, 2); k < min(kc1, g.nk - 2); ++k) {
        const int idx = i + g.ni * (j + g.nj * k);
        const float pz = h * (float)k;
        // Two velocity samples.  General h: bit-exact reference arithmetic (lerp_ref in device3d.cuh).
        // Power-of-two h: both points are grid nodes, where the staggered component has fraction 1/2
        // and the other two axes fraction 0, so each component is the mean of two faces (2 loads
        // instead of 8; the dropped terms have weight exactly 0).
        float3 v0, v1;
        float tx, ty, tz;
        if (P2) {
            v0 = velocity_at_node(vel, g, i, j, k);
            const int it = v0.x > 0.f ? i - 1 : i + 1, jt = v0.y > 0.f ? j - 1 : j + 1, kt = v0.z > 0.f ? k - 1 : k + 1;
            tx = v0.x > 0.f ? px - h : px + h;
            ty = v0.y > 0.f ? py - h : py + h;
            tz = v0.z > 0.f ? pz - h : pz + h;
            v1 = velocity_at_node(vel, g, it, jt, kt);
        } else {
            v0 = get_velocity_ref(vel, g, px, py, pz);
            tx = v0.x > 0.f ? px - h : px + h;
            ty = v0.y > 0.f ? py - h : py + h;
            tz = v0.z > 0.f ? pz - h : pz + h;
            v1 = get_velocity_ref(vel, g, tx, ty, tz);
        }
        // power-of-two h: px - tx is exactly +-h (both are exact multiples of h), and dividing by +-h is the exact
        // multiplication by +-1/h -- the same bits as the reference's division without three IEEE divisions
        const float ax = P2 ? (v0.x - v1.x) * (v0.x > 0.f ? g.inv_h : -g.inv_h) : (v0.x - v1.x) / (px - tx);
        const float ay = P2 ? (v0.y - v1.y) * (v0.y > 0.f ? g.inv_h : -g.inv_h) : (v0.y - v1.y) / (py - ty);
        const float az = P2 ? (v0.z - v1.z) * (v0.z > 0.f ? g.inv_h : -g.inv_h) : (v0.z - v1.z) / (pz - tz);
        const float nx = dmc_axis(px, v0.x, ax, substep);
        const float ny = dmc_axis(py, v0.y, ay, substep);
        const float nz = dmc_axis(pz, v0.z, az, substep);
        Frac fx = split<P2>(nx, g.h, g.inv_h), fy = split<P2>(ny, g.h, g.inv_h), fz = split<P2>(nz, g.h, g.inv_h);
        const int sy = g.ni, sz = g.ni * g.nj;
        const int o = fx.i + sy * fy.i + sz * fz.i;
    #pragma unroll
        for (int m = 0; m < NMAP; ++m) {
            out.x[m][idx] = tri8(in.x[m] + o, sy, sz, fx, fy, fz);
            out.y[m][idx] = tri8(in.y[m] + o, sy, sz, fx, fy, fz);
            out.z[m][idx] = tri8(in.z[m] + o, sy, sz, fx, fy, fz);
        }
    }
}

// ------------------------------------------------------------------ semi-Lagrangian (a12)
// semilag_kernel, GPU_kernel.cu:206-233; NF co-located fields share one back-trace.
template <bool P2, int NF>
__global__ void __launch_bounds__(256)
k_semilag(Grid3 g, int kbeg, int kend_, Vel3 vel, Stag st, FieldSetRW<NF> out, FieldSetRO<NF> src, float cfldt, float dt)
{
    const int fi = g.ni + st.dx, fj = g.nj + st.dy, fk = g.nk + st.dz;
    BMQ_IJK(fi, fj)
    if (!(i > 1 && i < fi - 2 - st.dx && j > 1 && j < fj - 2 - st.dy && k > 1 && k < fk - 2 - st.dz)) return;
    const float ox = -(float)st.dx * 0.5f * g.h, oy = -(float)st.dy * 0.5f * g.h, oz = -(float)st.dz * 0.5f * g.h;
    float3 p = make_float3(fmaf(g.h, (float)i, ox), fmaf(g.h, (float)j, oy), fmaf(g.h, (float)k, oz));
    p = trace<P2>(vel, g, cfldt, dt, p);
    float s[NF];
    sample_fields<P2, NF>(src.p, fi, fj, g, ox, oy, oz, p.x, p.y, p.z, s);
    const int idx = i + fi * (j + fj * k);
#pragma unroll
    for (int f = 0; f < NF; ++f) out.p[f][idx] = s[f];
}

// ------------------------------------------------------------------ advect (a6)
// advect_kernel, GPU_kernel.cu:312-374: f = 1/2 * sum_8 1/8 f0(clamp(chi(x+d))) + 1/2 f0(clamp(chi(x)))
template <bool P2, int NF>
__global__ void __launch_bounds__(256)
k_advect(Grid3 g, int kbeg, int kend_, Stag st, bool is_point, FieldSetRW<NF> out, FieldSetRO<NF> init, Map3 chi)
{
    const int fi = g.ni + st.dx, fj = g.nj + st.dy, fk = g.nk + st.dz;
    BMQ_IJK(fi, fj)
    if (!(2 + st.dx < i && i < fi - 3 && 2 + st.dy < j && j < fj - 3 && 2 + st.dz < k && k < fk - 3)) return;
    const float h = g.h;
    const float ox = -(float)st.dx * 0.5f * h, oy = -(float)st.dy * 0.5f * h, oz = -(float)st.dz * 0.5f * h;
    const float cx = fmaf(h, (float)i, ox), cy = fmaf(h, (float)j, oy), cz = fmaf(h, (float)k, oz);
    float sum[NF], val[NF], wgt[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) wgt[f] = is_point ? 1.0f : 0.125f;
    quad_gather<P2, NF>(chi, g, cx, cy, cz, h, h * (float)g.ni - h, h * (float)g.nj - h, h * (float)g.nk - h,
                        init.p, fi, fj, ox, oy, oz, is_point, wgt, sum, val);
    const int idx = i + fi * (j + fj * k);
#pragma unroll
    for (int f = 0; f < NF; ++f) out.p[f][idx] = fmaf(0.5f, sum[f], 0.5f * val[f]);
}

// ------------------------------------------------------------------ time-0 error (a7, first kernel)
// compensate_kernel, GPU_kernel.cu:438-499: e0 = quad9[f(clamp0(psi(x+d)))] - f_init
template <bool P2, int NF>
__global__ void __launch_bounds__(256)
k_error(Grid3 g, int kbeg, int kend_, Stag st, bool is_point, FieldSetRW<NF> e0, FieldSetRO<NF> src, FieldSetRO<NF> init, Map3 psi)
{
    const int fi = g.ni + st.dx, fj = g.nj + st.dy, fk = g.nk + st.dz;
    BMQ_IJK(fi, fj)
    if (!(1 + st.dx < i && i < fi - 2 && 1 + st.dy < j && j < fj - 2 && 1 + st.dz < k && k < fk - 2)) return;
    const float h = g.h;
    const float ox = -(float)st.dx * 0.5f * h, oy = -(float)st.dy * 0.5f * h, oz = -(float)st.dz * 0.5f * h;
    const float cx = fmaf(h, (float)i, ox), cy = fmaf(h, (float)j, oy), cz = fmaf(h, (float)k, oz);
    float sum[NF], val[NF], wgt[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) wgt[f] = is_point ? 1.0f : 0.125f;
    quad_gather<P2, NF>(psi, g, cx, cy, cz, 0.f, h * (float)g.ni, h * (float)g.nj, h * (float)g.nk,
                        src.p, fi, fj, ox, oy, oz, is_point, wgt, sum, val);
    const int idx = i + fi * (j + fj * k);
#pragma unroll
    for (int f = 0; f < NF; ++f) e0.p[f][idx] = fmaf(0.5f, sum[f], 0.5f * val[f]) - __ldg(init.p[f] + idx);
}

// ------------------------------------------------------------------ accumulate (a9)
// cumulate_kernel, GPU_kernel.cu:376-436: target += coeff * quad9[d(clamp0(map(x+d)))].
// NCH change sets (external forces, projection) are gathered through the same map in one pass
// and added in the reference's call order, (target + s0) + s1, so the result is the one two
// consecutive launches give.
template <bool P2, int NF, int NCH>
__global__ void __launch_bounds__(256)
k_cumulate(Grid3 g, int kbeg, int kend_, Stag st, bool is_point, FieldSetRW<NF> target, FieldSetRO<NF * NCH> change,
           Coeffs<NCH> coeff, Map3 map)
{
    const int fi = g.ni + st.dx, fj = g.nj + st.dy, fk = g.nk + st.dz;
    BMQ_IJK(fi, fj)
    if (!(1 + st.dx < i && i < fi - 2 && 1 + st.dy < j && j < fj - 2 && 1 + st.dz < k && k < fk - 2)) return;
    const float h = g.h;
    const float ox = -(float)st.dx * 0.5f * h, oy = -(float)st.dy * 0.5f * h, oz = -(float)st.dz * 0.5f * h;
    const float cx = fmaf(h, (float)i, ox), cy = fmaf(h, (float)j, oy), cz = fmaf(h, (float)k, oz);
    float sum[NF * NCH], val[NF * NCH], wgt[NF * NCH];
    // the reference accumulates weight*coeff*sample (GPU_kernel.cu:420): one weight per change set
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int f = 0; f < NF; ++f) wgt[c * NF + f] = (is_point ? 1.0f : 0.125f) * coeff.c[c];
    quad_gather<P2, NF * NCH>(map, g, cx, cy, cz, 0.f, h * (float)g.ni, h * (float)g.nj, h * (float)g.nk,
                              change.p, fi, fj, ox, oy, oz, is_point, wgt, sum, val);
    const int idx = i + fi * (j + fj * k);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        float t = target.p[f][idx];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const float v = coeff.c[c] * val[c * NF + f];
            t += fmaf(0.5f, sum[c * NF + f], 0.5f * v);
        }
        target.p[f][idx] = t;
    }
}

// ------------------------------------------------------------------ apply correction + clamp (a7)
// cumulate_kernel(coeff=-0.5) followed by clampExtrema_kernel (GPU_kernel.cu:659-665, 146-167)
// in one pass: out = clamp(f_adv - 1/2 quad9[e0(clamp0(chi(x+d)))], min27(f_adv), max27(f_adv)).
// Writes every cell of the plane (out = f_adv where the reference leaves f_adv untouched), so
// out may be a different buffer than f_adv and no device-to-device copy is needed.
template <bool P2, int NF>
__global__ void __launch_bounds__(256)
k_apply_clamp(Grid3 g, int kbeg, int kend_, Stag st, bool is_point, FieldSetRW<NF> out, FieldSetRO<NF> fadv,
              FieldSetRO<NF> e0, Map3 chi)
{
    const int fi = g.ni + st.dx, fj = g.nj + st.dy, fk = g.nk + st.dz;
    BMQ_IJK(fi, fj)
    const int idx = i + fi * (j + fj * k);
    float r[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) r[f] = __ldg(fadv.p[f] + idx);
    if (1 + st.dx < i && i < fi - 2 && 1 + st.dy < j && j < fj - 2 && 1 + st.dz < k && k < fk - 2) {
        const float h = g.h;
        const float ox = -(float)st.dx * 0.5f * h, oy = -(float)st.dy * 0.5f * h, oz = -(float)st.dz * 0.5f * h;
        const float cx = fmaf(h, (float)i, ox), cy = fmaf(h, (float)j, oy), cz = fmaf(h, (float)k, oz);
        float sum[NF], val[NF], wgt[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) wgt[f] = (is_point ? 1.0f : 0.125f) * -0.5f;
        quad_gather<P2, NF>(chi, g, cx, cy, cz, 0.f, h * (float)g.ni, h * (float)g.nj, h * (float)g.nk,
                            e0.p, fi, fj, ox, oy, oz, is_point, wgt, sum, val);
#pragma unroll
        for (int f = 0; f < NF; ++f) r[f] += fmaf(0.5f, sum[f], 0.5f * (-0.5f * val[f]));
    }
    if (i > 0 && i < fi - 1 && j > 0 && j < fj - 1 && k > 0 && k < fk - 1) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const float *b = fadv.p[f];
            float mx = __ldg(b + idx), mn = mx;
#pragma unroll
            for (int kk = -1; kk <= 1; ++kk)
#pragma unroll
                for (int jj = -1; jj <= 1; ++jj)
#pragma unroll
                    for (int ii = -1; ii <= 1; ++ii) {
                        const float v = __ldg(b + idx + ii + fi * (jj + fj * kk));
                        mx = fmaxf(mx, v);
                        mn = fminf(mn, v);
                    }
            r[f] = fminf(fmaxf(mn, r[f]), mx);
        }
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) out.p[f][idx] = r[f];
}

// ------------------------------------------------------------------ windowed 8-point variants
// Same arithmetic as k_advect / k_error / k_cumulate / k_apply_clamp with is_point == false, but
// the map samples of the 8 sub-cell points come from one node window per map component
// (quad_gather_win in device3d.cuh).  STAG: 0 centred, 1/2/3 = u/v/w faces (compile time).
#ifndef BMQ_WIN_MINBLOCKS
#define BMQ_WIN_MINBLOCKS 7   // 128-thread CTAs, >= 7 per SM (<= 73 registers, no spills): best of the A/B sweep in profiles/r1_variants.txt
#endif
#define BMQ_STAG_SETUP                                                                             \
    constexpr int DX = STAG == 1, DY = STAG == 2, DZ = STAG == 3;                                  \
    const int fi = g.ni + DX, fj = g.nj + DY, fk = g.nk + DZ;                                      \
    (void)fk;                                                                                      \
    BMQ_IJK(fi, fj)                                                                                \
    const float h = g.h;                                                                           \
    const float ox = -(float)DX * 0.5f * h, oy = -(float)DY * 0.5f * h, oz = -(float)DZ * 0.5f * h; \
    const float cx = fmaf(h, (float)i, ox), cy = fmaf(h, (float)j, oy), cz = fmaf(h, (float)k, oz); \
    const int idx = i + fi * (j + fj * k);

template <bool P2, int STAG, int NF, int FIX = 0>
__global__ void __launch_bounds__(32 * BMQ_BY * BMQ_BZ, BMQ_WIN_MINBLOCKS)
k_advect_win(Grid3 g_, int kbeg, int kend_, FieldSetRW<NF> out, FieldSetRO<NF> init, Map3 chi)
{
    const Grid3 g = fix_grid<FIX>(g_);
    BMQ_STAG_SETUP
    if (!(2 + DX < i && i < fi - 3 && 2 + DY < j && j < fj - 3 && 2 + DZ < k && k < fk - 3)) return;
    float sum[NF], val[NF], wgt[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) wgt[f] = 0.125f;
    quad_gather_win<P2, STAG, NF>(chi, g, i, j, k, cx, cy, cz, h, h * (float)g.ni - h, h * (float)g.nj - h,
                                  h * (float)g.nk - h, init.p, fi, fj, ox, oy, oz, wgt, sum, val);
#pragma unroll
    for (int f = 0; f < NF; ++f) out.p[f][idx] = fmaf(0.5f, sum[f], 0.5f * val[f]);
}

template <bool P2, int STAG, int NF, int FIX = 0>
__global__ void __launch_bounds__(32 * BMQ_BY * BMQ_BZ, BMQ_WIN_MINBLOCKS)
k_error_win(Grid3 g_, int kbeg, int kend_, FieldSetRW<NF> e0, FieldSetRO<NF> src, FieldSetRO<NF> init, Map3 psi)
{
    const Grid3 g = fix_grid<FIX>(g_);
    BMQ_STAG_SETUP
    if (!(1 + DX < i && i < fi - 2 && 1 + DY < j && j < fj - 2 && 1 + DZ < k && k < fk - 2)) return;
    float sum[NF], val[NF], wgt[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) wgt[f] = 0.125f;
    quad_gather_win<P2, STAG, NF>(psi, g, i, j, k, cx, cy, cz, 0.f, h * (float)g.ni, h * (float)g.nj, h * (float)g.nk,
                                  src.p, fi, fj, ox, oy, oz, wgt, sum, val);
#pragma unroll
    for (int f = 0; f < NF; ++f) e0.p[f][idx] = fmaf(0.5f, sum[f], 0.5f * val[f]) - __ldg(init.p[f] + idx);
}

template <bool P2, int STAG, int NF, int NCH, int FIX = 0>
__global__ void __launch_bounds__(32 * BMQ_BY * BMQ_BZ, BMQ_WIN_MINBLOCKS)
k_cumulate_win(Grid3 g_, int kbeg, int kend_, FieldSetRW<NF> target, FieldSetRO<NF * NCH> change, Coeffs<NCH> coeff, Map3 map)
{
    const Grid3 g = fix_grid<FIX>(g_);
    BMQ_STAG_SETUP
    if (!(1 + DX < i && i < fi - 2 && 1 + DY < j && j < fj - 2 && 1 + DZ < k && k < fk - 2)) return;
    float sum[NF * NCH], val[NF * NCH], wgt[NF * NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c)
#pragma unroll
        for (int f = 0; f < NF; ++f) wgt[c * NF + f] = 0.125f * coeff.c[c];
    quad_gather_win<P2, STAG, NF * NCH>(map, g, i, j, k, cx, cy, cz, 0.f, h * (float)g.ni, h * (float)g.nj,
                                        h * (float)g.nk, change.p, fi, fj, ox, oy, oz, wgt, sum, val);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        float t = target.p[f][idx];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const float v = coeff.c[c] * val[c * NF + f];
            t += fmaf(0.5f, sum[c * NF + f], 0.5f * v);
        }
        target.p[f][idx] = t;
    }
}

template <bool P2, int STAG, int NF, int FIX = 0>
__global__ void __launch_bounds__(32 * BMQ_BY * BMQ_BZ, BMQ_WIN_MINBLOCKS)
k_apply_clamp_win(Grid3 g_, int kbeg, int kend_, FieldSetRW<NF> out, FieldSetRO<NF> fadv, FieldSetRO<NF> e0, Map3 chi)
{
    const Grid3 g = fix_grid<FIX>(g_);
    BMQ_STAG_SETUP
    float r[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) r[f] = __ldg(fadv.p[f] + idx);
    if (1 + DX < i && i < fi - 2 && 1 + DY < j && j < fj - 2 && 1 + DZ < k && k < fk - 2) {
        float sum[NF], val[NF], wgt[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) wgt[f] = 0.125f * -0.5f;
        quad_gather_win<P2, STAG, NF>(chi, g, i, j, k, cx, cy, cz, 0.f, h * (float)g.ni, h * (float)g.nj,
                                      h * (float)g.nk, e0.p, fi, fj, ox, oy, oz, wgt, sum, val);
#pragma unroll
        for (int f = 0; f < NF; ++f) r[f] += fmaf(0.5f, sum[f], 0.5f * (-0.5f * val[f]));
    }
    if (i > 0 && i < fi - 1 && j > 0 && j < fj - 1 && k > 0 && k < fk - 1) {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const float *b = fadv.p[f];
            float mx = __ldg(b + idx), mn = mx;
#pragma unroll
            for (int kk = -1; kk <= 1; ++kk)
#pragma unroll
                for (int jj = -1; jj <= 1; ++jj)
#pragma unroll
                    for (int ii = -1; ii <= 1; ++ii) {
                        const float v = __ldg(b + idx + ii + fi * (jj + fj * kk));
                        mx = fmaxf(mx, v);
                        mn = fminf(mn, v);
                    }
            r[f] = fminf(fmaxf(mn, r[f]), mx);
        }
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) out.p[f][idx] = r[f];
}

// clampExtrema_kernel alone (legacy gpu_compensate_* keeps the reference's buffer contract)
__global__ void __launch_bounds__(256)
k_clamp_extrema(int fi, int fj, int fk, int kbeg, int kend_, const float *__restrict__ before, float *after)
{
    BMQ_IJK(fi, fj)
    if (!(i > 0 && i < fi - 1 && j > 0 && j < fj - 1 && k > 0 && k < fk - 1)) return;
    const int idx = i + fi * (j + fj * k);
    float mx = __ldg(before + idx), mn = mx;
#pragma unroll
    for (int kk = -1; kk <= 1; ++kk)
#pragma unroll
        for (int jj = -1; jj <= 1; ++jj)
#pragma unroll
            for (int ii = -1; ii <= 1; ++ii) {
                const float v = __ldg(before + idx + ii + fi * (jj + fj * kk));
                mx = fmaxf(mx, v);
                mn = fminf(mn, v);
            }
    after[idx] = fminf(fmaxf(mn, after[idx]), mx);
}

// ------------------------------------------------------------------ two-level blend (a8)
// doubleAdvect_kernel, GPU_kernel.cu:236-310
template <bool P2, int NF>
__global__ void __launch_bounds__(256)
k_double_advect(Grid3 g, int kbeg, int kend_, Stag st, bool is_point, FieldSetRW<NF> field, FieldSetRO<NF> prev,
                Map3 chi, Map3 chip, float blend)
{
    const int fi = g.ni + st.dx, fj = g.nj + st.dy, fk = g.nk + st.dz;
    BMQ_IJK(fi, fj)
    if (!(2 + st.dx < i && i < fi - 3 && 2 + st.dy < j && j < fj - 3 && 2 + st.dz < k && k < fk - 3)) return;
    const float h = g.h;
    const float ox = -(float)st.dx * 0.5f * h, oy = -(float)st.dy * 0.5f * h, oz = -(float)st.dz * 0.5f * h;
    const float cx = fmaf(h, (float)i, ox), cy = fmaf(h, (float)j, oy), cz = fmaf(h, (float)k, oz);
    const float hix = h * (float)g.ni - h, hiy = h * (float)g.nj - h, hiz = h * (float)g.nk - h;
    const float q = 0.25f * h;
    const int ev = is_point ? 1 : 8;
    const float wgt = is_point ? 1.0f : 0.125f;
    float sum[NF], val[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) sum[f] = 0.f;
    for (int ii = 0; ii <= ev; ++ii) {
        const bool centre = ii == ev;
        const float dx = (centre || is_point) ? 0.f : ((ii & 4) ? -q : q);
        const float dy = (centre || is_point) ? 0.f : ((ii & 2) ? -q : q);
        const float dz = (centre || is_point) ? 0.f : ((ii & 1) ? -q : q);
        float3 mid = sample_map<P2>(chi, g, cx + dx, cy + dy, cz + dz);
        mid.x = clampf(mid.x, h, hix); mid.y = clampf(mid.y, h, hiy); mid.z = clampf(mid.z, h, hiz);
        float3 fin = sample_map<P2>(chip, g, mid.x, mid.y, mid.z);
        fin.x = clampf(fin.x, h, hix); fin.y = clampf(fin.y, h, hiy); fin.z = clampf(fin.z, h, hiz);
        float s[NF];
        sample_fields<P2, NF>(prev.p, fi, fj, g, ox, oy, oz, fin.x, fin.y, fin.z, s);
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            if (centre) val[f] = s[f];
            else sum[f] = fmaf(wgt, s[f], sum[f]);
        }
    }
    const int idx = i + fi * (j + fj * k);
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        const float pv = 0.5f * (sum[f] + val[f]);
        field.p[f][idx] = fmaf(field.p[f][idx], blend, (1.f - blend) * pv);
    }
}

// ------------------------------------------------------------------ distortion (a10) + reductions
// estimate_kernel, GPU_kernel.cu:501-537, for NMAP mappers, with the max-reduction the
// reference does on the host (Mapping.cpp:100-117) fused in: warp shuffle -> block -> one
// atomicMax per block.  Also reduces max |map_z - z| (in world units) per mapper for halo sizing.
template <bool P2, int NMAP, int FIX = 0>
// power-of-two h: 80 registers (six CTAs of 128 threads per SM) let the four independent map samples of a cell keep
// their 96 loads in flight: 4.35 -> 2.78 ms at 512^3; with a general h the same budget is slower (4.8 -> 5.3 ms), and
// so are the DMC and forward kernels with any larger budget (profiles/r2_march_variants.md)
#ifndef BMQ_MAPK_MINBLOCKS
__global__ void __launch_bounds__(P2 ? 128 : 256, P2 ? 6 : 1)
#else
__global__ void BMQ_MAPK_BOUNDS
#endif
k_estimate(Grid3 g_, int kbeg, int kend_, int kchunk, MapSetRO<NMAP> bwd, MapSetRO<NMAP> fwd, DistOut<NMAP> outp,
           const signed char *__restrict__ boundary)
{
    const Grid3 g = fix_grid<FIX>(g_);
    BMQ_COLUMN(g.ni, g.nj)
    float d2[NMAP], dispz[NMAP];      // maxima over the column (max is exact and order-free: same bits as one cell per thread)
#pragma unroll
    for (int m = 0; m < NMAP; ++m) d2[m] = dispz[m] = 0.f;
    const bool column = i > 1 && i < g.ni - 2 && j > 1 && j < g.nj - 2;
#pragma unroll 1
    for (int k = max(kc0, 2); column && k < min(kc1, g.nk - 2); ++k) {
        const int idx = i + g.ni * (j + g.nj * k);
        const float px = g.h * (float)i, py = g.h * (float)j, pz = g.h * (float)k;
        const bool counted = boundary == nullptr || boundary[idx] != 2;
#pragma unroll
        for (int m = 0; m < NMAP; ++m) {
            Map3 B{bwd.x[m], bwd.y[m], bwd.z[m]}, F{fwd.x[m], fwd.y[m], fwd.z[m]};
            // the first sample of each round trip is taken AT the cell centre, a map node: with
            // power-of-two h its fractions are exactly 0 and the sample is the node value
            float3 b = P2 ? make_float3(__ldg(B.x + idx), __ldg(B.y + idx), __ldg(B.z + idx)) : sample_map<P2>(B, g, px, py, pz);
            float3 f = sample_map<P2>(F, g, b.x, b.y, b.z);
            const float dbf = (px - f.x) * (px - f.x) + (py - f.y) * (py - f.y) + (pz - f.z) * (pz - f.z);
            float3 f2 = P2 ? make_float3(__ldg(F.x + idx), __ldg(F.y + idx), __ldg(F.z + idx)) : sample_map<P2>(F, g, px, py, pz);
            float3 b2 = sample_map<P2>(B, g, f2.x, f2.y, f2.z);
            const float dfb = (px - b2.x) * (px - b2.x) + (py - b2.y) * (py - b2.y) + (pz - b2.z) * (pz - b2.z);
            const float d = fmaxf(dbf, dfb);
            if (outp.dist[m]) outp.dist[m][idx] = d;
            if (counted) d2[m] = fmaxf(d2[m], d);
            dispz[m] = fmaxf(dispz[m], fmaxf(fabsf(b.z - pz), fabsf(f2.z - pz)));
        }
    }
#pragma unroll
    for (int m = 0; m < NMAP; ++m) {
        if (outp.d2max[m]) block_atomic_max(d2[m], outp.d2max[m]);
        __syncthreads();
    }
#pragma unroll
    for (int m = 0; m < NMAP; ++m) {
        if (outp.dispz[m]) block_atomic_max(dispz[m], outp.dispz[m]);
        __syncthreads();
    }
}

// max |x| over up to three arrays (getCFL, BimocqSolver.cpp:1093-1117), grid-stride, float4 loads
__global__ void __launch_bounds__(256)
k_maxabs3(const float *__restrict__ a, size_t na, const float *__restrict__ b, size_t nb,
          const float *__restrict__ c, size_t nc, float *out)
{
    float m = 0.f;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float *arr[3] = {a, b, c};
    const size_t len[3] = {na, nb, nc};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const float *p = arr[q];
        if (!p) continue;
        const size_t n = len[q];
        // head to 16-byte alignment, vector body, tail
        size_t head = ((16 - ((uintptr_t)p & 15)) & 15) / 4;
        if (head > n) head = n;
        for (size_t e = tid; e < head; e += stride) m = fmaxf(m, fabsf(__ldg(p + e)));
        const float4 *p4 = reinterpret_cast<const float4 *>(p + head);
        const size_t n4 = (n - head) / 4;
        for (size_t e = tid; e < n4; e += stride) {
            const float4 v = __ldg(p4 + e);
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
        for (size_t e = head + n4 * 4 + tid; e < n; e += stride) m = fmaxf(m, fabsf(__ldg(p + e)));
    }
    block_atomic_max(m, out);
}

// a[i] += c*b[i]  (add_kernel, GPU_kernel.cu:560-565, with the missing bounds check) and
// out[i] = a[i] + c*b[i] (add_field_kernel, :878-883).  Streaming: 128-bit loads/stores when the
// three pointers are 16-byte aligned (always true for whole fields), scalar tail.
__device__ __forceinline__ bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__global__ void __launch_bounds__(256)
k_add_field(float *out, const float *a, const float *b, float c, size_t n)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if (aligned16(out) && aligned16(a) && aligned16(b)) {
        const size_t n4 = n / 4;
        const float4 *a4 = reinterpret_cast<const float4 *>(a), *b4 = reinterpret_cast<const float4 *>(b);
        float4 *o4 = reinterpret_cast<float4 *>(out);
        for (size_t e = tid; e < n4; e += stride) {
            const float4 x = a4[e], y = b4[e];
            o4[e] = make_float4(fmaf(c, y.x, x.x), fmaf(c, y.y, x.y), fmaf(c, y.z, x.z), fmaf(c, y.w, x.w));
        }
        done = n4 * 4;
    }
    for (size_t e = done + tid; e < n; e += stride) out[e] = fmaf(c, b[e], a[e]);
}
__global__ void __launch_bounds__(256) k_axpy(float *a, const float *b, float c, size_t n)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if (aligned16(a) && aligned16(b)) {
        const size_t n4 = n / 4;
        const float4 *b4 = reinterpret_cast<const float4 *>(b);
        float4 *a4 = reinterpret_cast<float4 *>(a);
        for (size_t e = tid; e < n4; e += stride) {
            const float4 x = a4[e], y = b4[e];
            a4[e] = make_float4(fmaf(c, y.x, x.x), fmaf(c, y.y, x.y), fmaf(c, y.z, x.z), fmaf(c, y.w, x.w));
        }
        done = n4 * 4;
    }
    for (size_t e = done + tid; e < n; e += stride) a[e] = fmaf(c, b[e], a[e]);
}

// identity maps x = i*h (Mapping.cpp:310-324) for up to two mappers x (psi, chi)
__global__ void __launch_bounds__(256) k_identity(Grid3 g, int kbeg, int kend_, IdentityOut o)
{
    BMQ_IJK(g.ni, g.nj)
    const int idx = i + g.ni * (j + g.nj * k);
    const float x = (float)i * g.h, y = (float)j * g.h, z = (float)k * g.h;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        if (o.x[m]) { o.x[m][idx] = x; o.y[m][idx] = y; o.z[m][idx] = z; }
    }
}

// ================================================================== host launchers
// Pitch specialisation: cubic-plane grids (ni == nj in {128, 256, 512}, the BASELINE sizes) run
// kernels whose x and y extents are compile-time constants (fix_grid).
static std::atomic<bool> g_pitch_spec{true};
void set_pitch_specialisation(bool on) { g_pitch_spec.store(on); }
static inline int fix_of(const Grid3 &g)
{
    if (!g_pitch_spec.load(std::memory_order_relaxed) || g.ni != g.nj) return 0;
    return (g.ni == 512 || g.ni == 256 || g.ni == 128) ? g.ni : 0;
}
int march_fix_of(const Grid3 &g) { return fix_of(g); }
bool march_is_pow2_h(const Grid3 &g) { return is_pow2_h(g); }
// Gather-kernel variant: 1 (default) = z-marching columns (march3d.cuh), 0 = one windowed cell per thread.
// Same arithmetic, bit-identical results; the knob lets tests and benchmarks compare the two.
static std::atomic<int> g_gather_variant{1};
void set_gather_variant(int v) { g_gather_variant.store(v); }
// 0: windowed kernels (round 1); 1: z-marching kernels, clamp fused into the apply kernel; 2: z-marching kernels, the
// clamp as its own shared-memory tiled kernel (clamp27.cu)
static inline bool use_march() { return g_gather_variant.load(std::memory_order_relaxed) >= 1; }
static inline bool split_clamp() { return g_gather_variant.load(std::memory_order_relaxed) == 2; }
#define DISPATCH_P2_FIX(g, KERNEL, NM, ...)                                                     \
    do {                                                                                        \
        switch (fix_of(g) * 2 + (is_pow2_h(g) ? 1 : 0)) {                                       \
        case 512 * 2 + 1: KERNEL<true, NM, 512><<<gr, bl, 0, s>>>(__VA_ARGS__); break;               \
        case 512 * 2: KERNEL<false, NM, 512><<<gr, bl, 0, s>>>(__VA_ARGS__); break;                  \
        case 256 * 2 + 1: KERNEL<true, NM, 256><<<gr, bl, 0, s>>>(__VA_ARGS__); break;               \
        case 256 * 2: KERNEL<false, NM, 256><<<gr, bl, 0, s>>>(__VA_ARGS__); break;                  \
        case 128 * 2 + 1: KERNEL<true, NM, 128><<<gr, bl, 0, s>>>(__VA_ARGS__); break;               \
        case 128 * 2: KERNEL<false, NM, 128><<<gr, bl, 0, s>>>(__VA_ARGS__); break;                  \
        case 1: KERNEL<true, NM><<<gr, bl, 0, s>>>(__VA_ARGS__); break;                         \
        default: KERNEL<false, NM><<<gr, bl, 0, s>>>(__VA_ARGS__); break;                       \
        }                                                                                       \
    } while (0)

#define DISPATCH_P2(g, CALL_T, CALL_F) do { if (is_pow2_h(g)) { CALL_T; } else { CALL_F; } } while (0)

cudaError_t launch_forward(cudaStream_t s, const Grid3 &g, KRange r, const float *u, const float *v,
                           const float *w, int nmap, float *const maps[][3], float cfldt, float dt)
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    Vel3 vel{u, v, w};
    const int kc = column_chunk(g.ni, g.nj, r.kend - r.kbeg);
    dim3 gr = grid3_columns(g.ni, g.nj, r, kc), bl = block3();
    if (nmap == 1) {
        MapSetRW<1> m; m.x[0] = maps[0][0]; m.y[0] = maps[0][1]; m.z[0] = maps[0][2];
        DISPATCH_P2_FIX(g, k_forward, 1, g, r.kbeg, r.kend, kc, vel, m, cfldt, dt);
    } else {
        MapSetRW<2> m;
        for (int q = 0; q < 2; ++q) { m.x[q] = maps[q][0]; m.y[q] = maps[q][1]; m.z[q] = maps[q][2]; }
        DISPATCH_P2_FIX(g, k_forward, 2, g, r.kbeg, r.kend, kc, vel, m, cfldt, dt);
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_dmc(cudaStream_t s, const Grid3 &g, KRange r, const float *u, const float *v,
                       const float *w, int nmap, const float *const in[][3], float *const out[][3],
                       float substep)
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    Vel3 vel{u, v, w};
    const int kc = column_chunk(g.ni, g.nj, r.kend - r.kbeg);
    dim3 gr = grid3_columns(g.ni, g.nj, r, kc), bl = block3();
    if (nmap == 1) {
        MapSetRO<1> a; MapSetRW<1> b;
        a.x[0] = in[0][0]; a.y[0] = in[0][1]; a.z[0] = in[0][2];
        b.x[0] = out[0][0]; b.y[0] = out[0][1]; b.z[0] = out[0][2];
        DISPATCH_P2_FIX(g, k_dmc, 1, g, r.kbeg, r.kend, kc, vel, a, b, substep);
    } else {
        MapSetRO<2> a; MapSetRW<2> b;
        for (int q = 0; q < 2; ++q) {
            a.x[q] = in[q][0]; a.y[q] = in[q][1]; a.z[q] = in[q][2];
            b.x[q] = out[q][0]; b.y[q] = out[q][1]; b.z[q] = out[q][2];
        }
        DISPATCH_P2_FIX(g, k_dmc, 2, g, r.kbeg, r.kend, kc, vel, a, b, substep);
    }
    count_launch();
    return cudaGetLastError();
}


// dispatch helpers: STAG (compile time) from the runtime staggering
static inline int stag_id(Stag st) { return st.dx ? 1 : st.dy ? 2 : st.dz ? 3 : 0; }
#define DISPATCH_STAG_P2(g, st, KERNEL, NFLIST, ...)                                            \
    do {                                                                                        \
        const int key_ = (is_pow2_h(g) ? 4 : 0) + stag_id(st);                                  \
        switch (fix_of(g) * 8 + key_) {                                                         \
        case 512 * 8 + 4 + 0: KERNEL<true, 0, NFLIST, 512><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 512 * 8 + 0: KERNEL<false, 0, NFLIST, 512><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 256 * 8 + 4 + 0: KERNEL<true, 0, NFLIST, 256><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 256 * 8 + 0: KERNEL<false, 0, NFLIST, 256><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 128 * 8 + 4 + 0: KERNEL<true, 0, NFLIST, 128><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 128 * 8 + 0: KERNEL<false, 0, NFLIST, 128><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 4 + 0: KERNEL<true, 0, NFLIST><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 0: KERNEL<false, 0, NFLIST><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 512 * 8 + 4 + 1: KERNEL<true, 1, NFLIST, 512><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 512 * 8 + 1: KERNEL<false, 1, NFLIST, 512><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 256 * 8 + 4 + 1: KERNEL<true, 1, NFLIST, 256><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 256 * 8 + 1: KERNEL<false, 1, NFLIST, 256><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 128 * 8 + 4 + 1: KERNEL<true, 1, NFLIST, 128><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 128 * 8 + 1: KERNEL<false, 1, NFLIST, 128><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 4 + 1: KERNEL<true, 1, NFLIST><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 1: KERNEL<false, 1, NFLIST><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 512 * 8 + 4 + 2: KERNEL<true, 2, NFLIST, 512><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 512 * 8 + 2: KERNEL<false, 2, NFLIST, 512><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 256 * 8 + 4 + 2: KERNEL<true, 2, NFLIST, 256><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 256 * 8 + 2: KERNEL<false, 2, NFLIST, 256><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 128 * 8 + 4 + 2: KERNEL<true, 2, NFLIST, 128><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 128 * 8 + 2: KERNEL<false, 2, NFLIST, 128><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 4 + 2: KERNEL<true, 2, NFLIST><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 2: KERNEL<false, 2, NFLIST><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 512 * 8 + 4 + 3: KERNEL<true, 3, NFLIST, 512><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 512 * 8 + 3: KERNEL<false, 3, NFLIST, 512><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 256 * 8 + 4 + 3: KERNEL<true, 3, NFLIST, 256><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 256 * 8 + 3: KERNEL<false, 3, NFLIST, 256><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 128 * 8 + 4 + 3: KERNEL<true, 3, NFLIST, 128><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 128 * 8 + 3: KERNEL<false, 3, NFLIST, 128><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 4 + 3: KERNEL<true, 3, NFLIST><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        case 3: KERNEL<false, 3, NFLIST><<<gr, bl, 0, s>>>(__VA_ARGS__); break; \
        default: break;                                                                         \
        }                                                                                       \
    } while (0)
#define COMMA ,

template <int NF> static FieldSetRO<NF> ro(const float *const *p) { FieldSetRO<NF> f; for (int q = 0; q < NF; ++q) f.p[q] = p[q]; return f; }
template <int NF> static FieldSetRW<NF> rw(float *const *p) { FieldSetRW<NF> f; for (int q = 0; q < NF; ++q) f.p[q] = p[q]; return f; }

cudaError_t launch_semilag(cudaStream_t s, const Grid3 &g, KRange r, Stag st, const float *u,
                           const float *v, const float *w, int nf, float *const *out,
                           const float *const *src, float cfldt, float dt)
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    Vel3 vel{u, v, w};
    dim3 gr = grid3(g.ni + st.dx, g.nj + st.dy, r), bl = block3();
    if (nf == 1)
        DISPATCH_P2(g, (k_semilag<true, 1><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, vel, st, rw<1>(out), ro<1>(src), cfldt, dt)),
                    (k_semilag<false, 1><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, vel, st, rw<1>(out), ro<1>(src), cfldt, dt)));
    else
        DISPATCH_P2(g, (k_semilag<true, 2><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, vel, st, rw<2>(out), ro<2>(src), cfldt, dt)),
                    (k_semilag<false, 2><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, vel, st, rw<2>(out), ro<2>(src), cfldt, dt)));
    count_launch();
    return cudaGetLastError();
}

static void k_dispatch_centred2_advect(cudaStream_t s, const Grid3 &g, KRange r, dim3 gr, dim3 bl, float *const *out,
                                       const float *const *init, Map3 m)
{
    DISPATCH_STAG_P2(g, (Stag{0, 0, 0}), k_advect_win, 2, g, r.kbeg, r.kend, rw<2>(out), ro<2>(init), m);
}

cudaError_t launch_advect(cudaStream_t s, const Grid3 &g, KRange r, Stag st, bool is_point, int nf,
                          float *const *out, const float *const *init, const float *const chi[3])
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    Map3 m{chi[0], chi[1], chi[2]};
    dim3 gr = grid3(g.ni + st.dx, g.nj + st.dy, r), bl = block3();
    if (!is_point && (nf == 1 || stag_id(st) == 0) && use_march()) {
        count_launch();
        return launch_advect_march(s, g, r, st, nf, out, init, chi);
    }
    if (!is_point && (nf == 1 || stag_id(st) == 0)) {
        if (nf == 1) DISPATCH_STAG_P2(g, st, k_advect_win, 1, g, r.kbeg, r.kend, rw<1>(out), ro<1>(init), m);
        else k_dispatch_centred2_advect(s, g, r, gr, bl, out, init, m);
        count_launch();
        return cudaGetLastError();
    }
    if (nf == 1)
        DISPATCH_P2(g, (k_advect<true, 1><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<1>(out), ro<1>(init), m)),
                    (k_advect<false, 1><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<1>(out), ro<1>(init), m)));
    else
        DISPATCH_P2(g, (k_advect<true, 2><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<2>(out), ro<2>(init), m)),
                    (k_advect<false, 2><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<2>(out), ro<2>(init), m)));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_error(cudaStream_t s, const Grid3 &g, KRange r, Stag st, bool is_point, int nf,
                         float *const *e0, const float *const *src, const float *const *init,
                         const float *const psi[3])
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    Map3 m{psi[0], psi[1], psi[2]};
    dim3 gr = grid3(g.ni + st.dx, g.nj + st.dy, r), bl = block3();
    if (!is_point && (nf == 1 || stag_id(st) == 0) && use_march()) {
        count_launch();
        return launch_error_march(s, g, r, st, nf, e0, src, init, psi);
    }
    if (!is_point && (nf == 1 || stag_id(st) == 0)) {
        if (nf == 1) DISPATCH_STAG_P2(g, st, k_error_win, 1, g, r.kbeg, r.kend, rw<1>(e0), ro<1>(src), ro<1>(init), m);
        else DISPATCH_STAG_P2(g, (Stag{0, 0, 0}), k_error_win, 2, g, r.kbeg, r.kend, rw<2>(e0), ro<2>(src), ro<2>(init), m);
        count_launch();
        return cudaGetLastError();
    }
    if (nf == 1)
        DISPATCH_P2(g, (k_error<true, 1><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<1>(e0), ro<1>(src), ro<1>(init), m)),
                    (k_error<false, 1><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<1>(e0), ro<1>(src), ro<1>(init), m)));
    else
        DISPATCH_P2(g, (k_error<true, 2><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<2>(e0), ro<2>(src), ro<2>(init), m)),
                    (k_error<false, 2><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<2>(e0), ro<2>(src), ro<2>(init), m)));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_cumulate(cudaStream_t s, const Grid3 &g, KRange r, Stag st, bool is_point, int nf,
                            int nch, float *const *target, const float *const *change,
                            const float *coeff, const float *const map[3])
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    Map3 m{map[0], map[1], map[2]};
    dim3 gr = grid3(g.ni + st.dx, g.nj + st.dy, r), bl = block3();
    if (!is_point && (nf == 1 || stag_id(st) == 0) && use_march() && ((nf == 1 && nch <= 2) || (nf == 2 && nch == 1))) {
        count_launch();
        return launch_cumulate_march(s, g, r, st, nf, nch, target, change, coeff, map);
    }
    if (!is_point && (nf == 1 || stag_id(st) == 0)) {
        if (nf == 1 && nch == 1) {
            Coeffs<1> c; c.c[0] = coeff[0];
            DISPATCH_STAG_P2(g, st, k_cumulate_win, 1 COMMA 1, g, r.kbeg, r.kend, rw<1>(target), ro<1>(change), c, m);
        } else if (nf == 1 && nch == 2) {
            Coeffs<2> c; c.c[0] = coeff[0]; c.c[1] = coeff[1];
            DISPATCH_STAG_P2(g, st, k_cumulate_win, 1 COMMA 2, g, r.kbeg, r.kend, rw<1>(target), ro<2>(change), c, m);
        } else if (nf == 2 && nch == 1) {
            Coeffs<1> c; c.c[0] = coeff[0];
            DISPATCH_STAG_P2(g, (Stag{0, 0, 0}), k_cumulate_win, 2 COMMA 1, g, r.kbeg, r.kend, rw<2>(target), ro<2>(change), c, m);
        } else return cudaErrorInvalidValue;
        count_launch();
        return cudaGetLastError();
    }
#define CUM(NF, NCH)                                                                                   \
    {                                                                                                  \
        Coeffs<NCH> c;                                                                                 \
        for (int q = 0; q < NCH; ++q) c.c[q] = coeff[q];                                               \
        DISPATCH_P2(g, (k_cumulate<true, NF, NCH><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<NF>(target), ro<NF * NCH>(change), c, m)), \
                    (k_cumulate<false, NF, NCH><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<NF>(target), ro<NF * NCH>(change), c, m))); \
    }
    if (nf == 1 && nch == 1) CUM(1, 1)
    else if (nf == 1 && nch == 2) CUM(1, 2)
    else if (nf == 2 && nch == 1) CUM(2, 1)
    else return cudaErrorInvalidValue;
#undef CUM
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_apply_clamp(cudaStream_t s, const Grid3 &g, KRange r, Stag st, bool is_point, int nf,
                               float *const *out, const float *const *fadv, const float *const *e0,
                               const float *const chi[3])
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    Map3 m{chi[0], chi[1], chi[2]};
    dim3 gr = grid3(g.ni + st.dx, g.nj + st.dy, r), bl = block3();
    if (!is_point && (nf == 1 || stag_id(st) == 0) && use_march()) {
        count_launch();
        if (split_clamp()) {
            count_launch();
            return launch_apply_march_split(s, g, r, st, nf, out, fadv, e0, chi);
        }
        return launch_apply_march(s, g, r, st, nf, out, fadv, e0, chi);
    }
    if (!is_point && (nf == 1 || stag_id(st) == 0)) {
        if (nf == 1) DISPATCH_STAG_P2(g, st, k_apply_clamp_win, 1, g, r.kbeg, r.kend, rw<1>(out), ro<1>(fadv), ro<1>(e0), m);
        else DISPATCH_STAG_P2(g, (Stag{0, 0, 0}), k_apply_clamp_win, 2, g, r.kbeg, r.kend, rw<2>(out), ro<2>(fadv), ro<2>(e0), m);
        count_launch();
        return cudaGetLastError();
    }
    if (nf == 1)
        DISPATCH_P2(g, (k_apply_clamp<true, 1><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<1>(out), ro<1>(fadv), ro<1>(e0), m)),
                    (k_apply_clamp<false, 1><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<1>(out), ro<1>(fadv), ro<1>(e0), m)));
    else
        DISPATCH_P2(g, (k_apply_clamp<true, 2><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<2>(out), ro<2>(fadv), ro<2>(e0), m)),
                    (k_apply_clamp<false, 2><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<2>(out), ro<2>(fadv), ro<2>(e0), m)));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_clamp_extrema(cudaStream_t s, int fi, int fj, int fk, KRange r, const float *before,
                                 float *after)
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    k_clamp_extrema<<<grid3(fi, fj, r), block3(), 0, s>>>(fi, fj, fk, r.kbeg, r.kend, before, after);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_double_advect(cudaStream_t s, const Grid3 &g, KRange r, Stag st, bool is_point, int nf,
                                 float *const *field, const float *const *prev, const float *const chi[3],
                                 const float *const chip[3], float blend)
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    Map3 m{chi[0], chi[1], chi[2]}, mp{chip[0], chip[1], chip[2]};
    dim3 gr = grid3(g.ni + st.dx, g.nj + st.dy, r), bl = block3();
    if (nf == 1)
        DISPATCH_P2(g, (k_double_advect<true, 1><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<1>(field), ro<1>(prev), m, mp, blend)),
                    (k_double_advect<false, 1><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<1>(field), ro<1>(prev), m, mp, blend)));
    else
        DISPATCH_P2(g, (k_double_advect<true, 2><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<2>(field), ro<2>(prev), m, mp, blend)),
                    (k_double_advect<false, 2><<<gr, bl, 0, s>>>(g, r.kbeg, r.kend, st, is_point, rw<2>(field), ro<2>(prev), m, mp, blend)));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_estimate(cudaStream_t s, const Grid3 &g, KRange r, int nmap, const float *const bwd[][3],
                            const float *const fwd[][3], float *const *dist, float *const *d2max,
                            float *const *dispz, const signed char *boundary)
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    const int kc = column_chunk(g.ni, g.nj, r.kend - r.kbeg);
    dim3 gr = grid3_columns(g.ni, g.nj, r, kc), bl = block3();
    if (nmap == 1) {
        MapSetRO<1> b, f; DistOut<1> o;
        b.x[0] = bwd[0][0]; b.y[0] = bwd[0][1]; b.z[0] = bwd[0][2];
        f.x[0] = fwd[0][0]; f.y[0] = fwd[0][1]; f.z[0] = fwd[0][2];
        o.dist[0] = dist ? dist[0] : nullptr; o.d2max[0] = d2max ? d2max[0] : nullptr; o.dispz[0] = dispz ? dispz[0] : nullptr;
        DISPATCH_P2_FIX(g, k_estimate, 1, g, r.kbeg, r.kend, kc, b, f, o, boundary);
    } else {
        MapSetRO<2> b, f; DistOut<2> o;
        for (int q = 0; q < 2; ++q) {
            b.x[q] = bwd[q][0]; b.y[q] = bwd[q][1]; b.z[q] = bwd[q][2];
            f.x[q] = fwd[q][0]; f.y[q] = fwd[q][1]; f.z[q] = fwd[q][2];
            o.dist[q] = dist ? dist[q] : nullptr; o.d2max[q] = d2max ? d2max[q] : nullptr;
            o.dispz[q] = dispz ? dispz[q] : nullptr;
        }
        DISPATCH_P2_FIX(g, k_estimate, 2, g, r.kbeg, r.kend, kc, b, f, o, boundary);
    }
    count_launch();
    return cudaGetLastError();
}

static int reduce_blocks(size_t n)
{
    size_t b = (n / 4 + 255) / 256;
    const size_t cap = 148 * 8;   // 148 SMs x 8 resident 256-thread CTAs
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

cudaError_t launch_maxabs3(cudaStream_t s, const float *a, size_t na, const float *b, size_t nb,
                           const float *c, size_t nc, float *out_dev)
{
    size_t n = na > nb ? na : nb;
    n = n > nc ? n : nc;
    k_maxabs3<<<reduce_blocks(n), 256, 0, s>>>(a, na, b, nb, c, nc, out_dev);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_axpy(cudaStream_t s, float *a, const float *b, float c, size_t n)
{
    if (n == 0) return cudaSuccess;
    k_axpy<<<reduce_blocks(n * 4), 256, 0, s>>>(a, b, c, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_add_field(cudaStream_t s, float *out, const float *a, const float *b, float c, size_t n)
{
    if (n == 0) return cudaSuccess;
    k_add_field<<<reduce_blocks(n * 4), 256, 0, s>>>(out, a, b, c, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_identity(cudaStream_t s, const Grid3 &g, KRange r, int nsets, float *const sets[][3])
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    IdentityOut o;
    for (int m = 0; m < 4; ++m) {
        o.x[m] = m < nsets ? sets[m][0] : nullptr;
        o.y[m] = m < nsets ? sets[m][1] : nullptr;
        o.z[m] = m < nsets ? sets[m][2] : nullptr;
    }
    k_identity<<<grid3(g.ni, g.nj, r), block3(), 0, s>>>(g, r.kbeg, r.kend, o);
    count_launch();
    return cudaGetLastError();
}

}  // namespace bmq
