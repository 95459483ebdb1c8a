// Source terms around the advection path: smoke emitter, buoyancy,
// Jacobi diffusion and the mad helper -- the reference's `extern "C"` gpu_emit_smoke,
// gpu_add_buoyancy, gpu_diffuse_field, gpu_mad (bimocq3D/GPU_Advection.h:88-95,103; definitions
// bimocq3D/GPU_kernel.cu:736-876, 952-964) with the same prototypes, on the legacy default stream.
//
// All four are streaming / 7-point-stencil kernels bounded by HBM bandwidth.  The arithmetic
// expressions are the reference's (same float/double promotions, same CUDA math functions), so
// the toolchain makes the same contraction choices and the results are bit-identical.
#include "common.h"
#include "launch3d.h"

namespace bmq {
void count_launches(unsigned n);   // kernels3d.cu
}

namespace {

#define SRC_IJK(fi, fj, fk)                                       \
    const int i = blockIdx.x * 32 + threadIdx.x;                  \
    const int j = blockIdx.y * 4 + threadIdx.y;                   \
    const int k = blockIdx.z;                                     \
    if (i >= (fi) || j >= (fj) || k >= (fk)) return;              \
    const int index = i + (fi) * (j + (fj) * k);

dim3 sblk() { return dim3(32, 4, 1); }
dim3 sgrd(int fi, int fj, int fk) { return dim3((fi + 31) / 32, (fj + 3) / 4, fk); }

// emit_smoke_velocity_kernel, GPU_kernel.cu:736-758 (ni,nj,nk are the FIELD's dimensions)
__global__ void __launch_bounds__(128)
k_emit_velocity(float *field, float h, int ni, int nj, int nk, float cx, float cy, float cz, float radius, float emiter)
{
    SRC_IJK(ni, nj, nk)
    if (!(i > 1 && i < ni - 2 && j > 1 && j < nj - 2 && k > 1 && k < nk - 2)) return;
    // written like the reference source (mixed double/float literals included) so that nvcc and
    // ptxas make the same promotion and contraction choices: the result is bit-identical
    const float dx = ((float)i - 0.5) * h - cx;
    const float dy = j * h - cy;
    const float dz = k * h - cz;
    const float length = norm3df(dx, dy, dz);
    if (length < radius) {
        const float theta = acosf(dy / hypotf(dy, dz));
        const float vel_x = emiter * 0.06 * (1.0 + 0.01 * cosf(8.0 * theta));
        field[index] = vel_x;
    }
}

// emit_smoke_field_kernel, GPU_kernel.cu:760-780
__global__ void __launch_bounds__(128)
k_emit_field(float *rho, float *T, float h, int ni, int nj, int nk, float cx, float cy, float cz, float radius,
             float density, float temperature)
{
    SRC_IJK(ni, nj, nk)
    if (!(i > 1 && i < ni - 2 && j > 1 && j < nj - 2 && k > 1 && k < nk - 2)) return;
    const float dx = i * h - cx, dy = j * h - cy, dz = k * h - cz;
    if (norm3df(dx, dy, dz) < radius) {
        rho[index] = density;
        T[index] = temperature;
    }
}

// add_buoyancy_kernel, GPU_kernel.cu:804-823: v(i,j,k) += 0.5 dt (beta (T_j + T_{j-1}) - alpha (rho_j + rho_{j-1})).
// The reference indexes density/temperature with the v-face index (nj+1 rows per plane), i.e. with
// the v-field's row pitch; that addressing is reproduced (it is what the reference computes).  Through it
// the last ni*nk face indices point past the end of the cell-centred arrays (the reference reads whatever
// lies behind them); here such reads return 0 (n_scalar = number of floats density / temperature hold).
__global__ void __launch_bounds__(128)
k_add_buoyancy(float *field, const float *__restrict__ density, const float *__restrict__ temperature, int ni, int nj,
               int nk, float alpha, float beta, float dt, long long n_scalar)
{
    SRC_IJK(ni, nj, nk)
    if (!(j > 0)) return;
    const int index1 = index - ni;
    const bool in0 = index < n_scalar, in1 = index1 < n_scalar;
    const float d0 = in0 ? density[index] : 0.f, T0 = in0 ? temperature[index] : 0.f;
    const float d1 = in1 ? density[index1] : 0.f, T1 = in1 ? temperature[index1] : 0.f;
    const float f = 0.5 * dt * (beta * (T0 + T1) - alpha * (d0 + d1));   // the reference's expression, same contraction
    field[index] += f;
}

// the same with four cells per thread (128-bit accesses; needs ni % 4 == 0 and 16-byte aligned bases)
__global__ void __launch_bounds__(128)
k_add_buoyancy4(float *field, const float *__restrict__ density, const float *__restrict__ temperature, int ni, int nj,
                int nk, float alpha, float beta, float dt, long long n_scalar)
{
    const int i4 = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    if (i4 * 4 >= ni || j >= nj || k >= nk || !(j > 0)) return;
    const size_t index = (size_t)i4 * 4 + (size_t)ni * (j + (size_t)nj * k), index1 = index - ni;
    // n_scalar and both indices are multiples of 4 (ni % 4 == 0): a quad is wholly inside or wholly outside
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool in0 = (long long)index < n_scalar, in1 = (long long)index1 < n_scalar;
    const float4 d0 = in0 ? *reinterpret_cast<const float4 *>(density + index) : zero, T0 = in0 ? *reinterpret_cast<const float4 *>(temperature + index) : zero;
    const float4 d1 = in1 ? *reinterpret_cast<const float4 *>(density + index1) : zero, T1 = in1 ? *reinterpret_cast<const float4 *>(temperature + index1) : zero;
    float4 f = *reinterpret_cast<float4 *>(field + index);
    // the reference's expression per lane, same contraction
    const float ix = 0.5 * dt * (beta * (T0.x + T1.x) - alpha * (d0.x + d1.x));
    const float iy = 0.5 * dt * (beta * (T0.y + T1.y) - alpha * (d0.y + d1.y));
    const float iz = 0.5 * dt * (beta * (T0.z + T1.z) - alpha * (d0.z + d1.z));
    const float iw = 0.5 * dt * (beta * (T0.w + T1.w) - alpha * (d0.w + d1.w));
    f.x += ix; f.y += iy; f.z += iz; f.w += iw;
    *reinterpret_cast<float4 *>(field + index) = f;
}

// diffuse_field_kernel, GPU_kernel.cu:834-853: one Jacobi sweep of (I - coef Lap) x = field
__global__ void __launch_bounds__(128)
k_diffuse(const float *__restrict__ field, const float *__restrict__ in, float *out, int ni, int nj, int nk, float coef)
{
    SRC_IJK(ni, nj, nk)
    if (!(i > 0 && i < ni - 1 && j > 0 && j < nj - 1 && k > 0 && k < nk - 1)) return;
    const int sy = ni, sz = ni * nj;
    float s = __fadd_rn(__ldg(in + index - 1), __ldg(in + index + 1));
    s = __fadd_rn(s, __ldg(in + index - sy));
    s = __fadd_rn(s, __ldg(in + index + sy));
    s = __fadd_rn(s, __ldg(in + index - sz));
    s = __fadd_rn(s, __ldg(in + index + sz));
    out[index] = __fdiv_rn(__fmaf_rn(coef, s, __ldg(field + index)), __fmaf_rn(coef, 6.0f, 1.0f));
}

// mad_kernel, GPU_kernel.cu:952-957 (with the bounds check the reference lacks); 128-bit accesses
// when the pointers allow it
__global__ void __launch_bounds__(256)
k_mad(float *field, const float *f1, const float *f2, float c1, float c2, size_t n)
{
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    size_t done = 0;
    if (((reinterpret_cast<uintptr_t>(field) | reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(f2)) & 15) == 0) {
        const size_t n4 = n / 4;
        const float4 *a4 = reinterpret_cast<const float4 *>(f1), *b4 = reinterpret_cast<const float4 *>(f2);
        float4 *o4 = reinterpret_cast<float4 *>(field);
        for (size_t e = tid; e < n4; e += stride) {
            const float4 x = a4[e], y = b4[e];
            o4[e] = make_float4(__fmaf_rn(c1, x.x, __fmul_rn(c2, y.x)), __fmaf_rn(c1, x.y, __fmul_rn(c2, y.y)),
                                __fmaf_rn(c1, x.z, __fmul_rn(c2, y.z)), __fmaf_rn(c1, x.w, __fmul_rn(c2, y.w)));
        }
        done = n4 * 4;
    }
    for (size_t e = done + tid; e < n; e += stride) field[e] = __fmaf_rn(c1, f1[e], __fmul_rn(c2, f2[e]));
}

// clamp_extrema_kernel, GPU_kernel.cu:892-941 (MacCormack / Reflection schemes,
// BimocqGPUSolver.cpp:232-338): every field cell is traced back with a midpoint step and the value
// of fieldTemp AT THE BACK-TRACED CELL is replaced by the trilinear sample of `field` when it lies
// outside the range of the eight surrounding values.  Reproduced literally, including
//   * the grid index taken as floor(position) without dividing by h (:913-915), so for the shipped
//     scenes (L = 0.2) every thread addresses cell (0,0,0) or (-1,..);
//   * the scatter: several threads may write the same fieldTemp cell (a race in the reference; the
//     result there is the value of whichever thread wins -- same here).
// Only difference: reads outside `field` / u / v / w return 0 and writes outside fieldTemp are dropped
// (the reference accesses whatever memory is there).
// get_velocity_ref / tri8_ref (device3d.cuh) with every node read guarded: the reference samples the velocity half a
// cell past the last face of a staggered field (point = h * (index + 0.5), :900) and wherever a wild back-trace lands;
// nodes outside the arrays read as 0 here instead of whatever memory is there.  In-range arithmetic is unchanged.
__device__ __forceinline__ float tri8_ref_guarded(const float *__restrict__ base, long long n, long long sy, long long sz,
                                                  const bmq::Frac &x, const bmq::Frac &y, const bmq::Frac &z)
{
    const long long o = (long long)x.i + sy * y.i + sz * z.i;
    auto at = [&](long long q) -> float { return q >= 0 && q < n ? __ldg(base + q) : 0.f; };
    const float v000 = at(o), v001 = at(o + 1), v010 = at(o + sy), v011 = at(o + sy + 1);
    const float v100 = at(o + sz), v101 = at(o + sz + 1), v110 = at(o + sz + sy), v111 = at(o + sz + sy + 1);
    const float a0 = bmq::lerp_ref(v000, v001, x.f), a1 = bmq::lerp_ref(v010, v011, x.f);
    const float a2 = bmq::lerp_ref(v100, v101, x.f), a3 = bmq::lerp_ref(v110, v111, x.f);
    return bmq::lerp_ref(bmq::lerp_ref(a0, a1, y.f), bmq::lerp_ref(a2, a3, y.f), z.f);
}
__device__ __forceinline__ float3 get_velocity_ref_guarded(const bmq::Vel3 &vel, const bmq::Grid3 &g, float px, float py, float pz)
{
    const float hh = 0.5f * g.h;
    const bmq::Frac x0 = bmq::split<false>(px, g.h, g.inv_h), x5 = bmq::split<false>(px + hh, g.h, g.inv_h);
    const bmq::Frac y0 = bmq::split<false>(py, g.h, g.inv_h), y5 = bmq::split<false>(py + hh, g.h, g.inv_h);
    const bmq::Frac z0 = bmq::split<false>(pz, g.h, g.inv_h), z5 = bmq::split<false>(pz + hh, g.h, g.inv_h);
    const long long ni = g.ni, nj = g.nj, nk = g.nk;
    float3 r;
    r.x = tri8_ref_guarded(vel.u, (ni + 1) * nj * nk, ni + 1, (ni + 1) * nj, x5, y0, z0);
    r.y = tri8_ref_guarded(vel.v, ni * (nj + 1) * nk, ni, ni * (nj + 1), x0, y5, z0);
    r.z = tri8_ref_guarded(vel.w, ni * nj * (nk + 1), ni, ni * nj, x0, y0, z5);
    return r;
}

__global__ void __launch_bounds__(128)
k_clamp_extrema_mc(const float *__restrict__ field, float *fieldTemp, bmq::Vel3 vel, int ni, int nj, int nk, int dimx,
                   int dimy, int dimz, float ox, float oy, float oz, float h, float dt)
{
    SRC_IJK(ni, nj, nk)
    (void)index;
    bmq::Grid3 g;
    g.ni = ni - dimx; g.nj = nj - dimy; g.nk = nk - dimz; g.h = h; g.inv_h = -(1.0f / h); g.p2 = 0;   // negative: IEEE division (device3d.cuh)
    const float ptx = h * (float(i) + ox), pty = h * (float(j) + oy), ptz = h * (float(k) + oz);
    float3 v = get_velocity_ref_guarded(vel, g, ptx, pty, ptz);
    const float halfdt = 0.5f * dt;
    float pxx = ptx - v.x * halfdt, pxy = pty - v.y * halfdt, pxz = ptz - v.z * halfdt;
    v = get_velocity_ref_guarded(vel, g, pxx, pxy, pxz);
    pxx = ptx - v.x * dt; pxy = pty - v.y * dt; pxz = ptz - v.z * dt;
    const int gi = (int)floor(pxx), gj = (int)floor(pxy), gk = (int)floor(pxz);
    const float cx = pxx - (float)gi, cy = pxy - (float)gj, cz = pxz - (float)gk;
    const long long ol = (long long)gk * nj * ni + (long long)gj * ni + gi;
    const long long total = (long long)ni * nj * nk;
    if (ol < 0 || ol >= total) return;                      // the reference would write outside the array
    auto at = [&](long long q) -> float { return q < total ? field[q] : 0.f; };   // reads past the end: zero padding
    const long long sy = ni, sz = (long long)nj * ni;
    const int o = (int)ol;
    const float v0 = at(ol), v1 = at(ol + 1), v2 = at(ol + sy), v3 = at(ol + sy + 1);
    const float v4 = at(ol + sz), v5 = at(ol + sz + 1), v6 = at(ol + sz + sy), v7 = at(ol + sz + sy + 1);
    const float mn = min(v0, min(v1, min(v2, min(v3, min(v4, min(v5, min(v6, v7)))))));
    const float mx = max(v0, max(v1, max(v2, max(v3, max(v4, max(v5, max(v6, v7)))))));
    const float temp = fieldTemp[o];
    if (temp < mn || temp > mx) {
        const float iv1 = bmq::lerp_ref(bmq::lerp_ref(v0, v1, cx), bmq::lerp_ref(v2, v3, cx), cy);
        const float iv2 = bmq::lerp_ref(bmq::lerp_ref(v4, v5, cx), bmq::lerp_ref(v6, v7, cx), cy);
        fieldTemp[o] = bmq::lerp_ref(iv1, iv2, cz);
    }
}

// ---- 8^3-blocked host container layout ------------------------------------
// buffer3Df stores cell (i,j,k) at ((K*bx*by + J*bx + I) << 9) + (kk << 6) + (jj << 3) + ii with
// I = i>>3, ii = i&7, ... and bx = ceil(nx/8) (include/fluid_buffer3D.h:173-189).  A warp moves 128
// contiguous bytes of a dense row = the 32-byte rows of four neighbouring blocks: whole sectors on
// both sides.
template <bool TO_LINEAR>
__global__ void __launch_bounds__(256)
k_relayout(const float *__restrict__ src, float *__restrict__ dst, int nx, int ny, int nz, int bx, int by)
{
    // CTA (32, 8): lanes run along i (128 contiguous bytes of a dense row = the 32-byte rows of four
    // blocks), threadIdx.y = jj, blockIdx.y = J, blockIdx.z = K, loop over kk.  No divisions.
    const int i = blockIdx.x * 32 + threadIdx.x, I = i >> 3, ii = i & 7;
    if (I >= bx) return;
    const int J = blockIdx.y, K = blockIdx.z, jj = threadIdx.y, j = J * 8 + jj;
    const size_t blk = ((size_t)(K * by + J) * bx + I) << 9;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        const int k = K * 8 + kk;
        const size_t e = blk + (kk << 6) + (jj << 3) + ii;
        const bool in = i < nx && j < ny && k < nz;
        const size_t lin = (size_t)i + (size_t)nx * ((size_t)j + (size_t)ny * k);
        if (TO_LINEAR) {
            if (in) dst[lin] = src[e];
        } else {
            dst[e] = in ? src[lin] : 0.f;   // padding cells: zero, as Buffer3D::init leaves them
        }
    }
}

}  // namespace

namespace bmq {
size_t blocked_elems(int nx, int ny, int nz) { return (size_t)((nx + 7) / 8) * ((ny + 7) / 8) * ((nz + 7) / 8) * 512; }
cudaError_t launch_relayout(cudaStream_t s, bool to_linear, const float *src, float *dst, int nx, int ny, int nz)
{
    const int bx = (nx + 7) / 8, by = (ny + 7) / 8, bz = (nz + 7) / 8;
    dim3 grid((bx * 8 + 31) / 32, by, bz), block(32, 8);
    if (to_linear) k_relayout<true><<<grid, block, 0, s>>>(src, dst, nx, ny, nz, bx, by);
    else k_relayout<false><<<grid, block, 0, s>>>(src, dst, nx, ny, nz, bx, by);
    count_launches(1);
    return cudaGetLastError();
}
}  // namespace bmq

extern "C" {

long long bmq_blocked_elems(int nx, int ny, int nz)
{
    if (nx <= 0 || ny <= 0 || nz <= 0) return 0;
    return (long long)bmq::blocked_elems(nx, ny, nz);
}

int bmq_blocked_to_linear(const float *blocked_dev, float *linear_dev, int nx, int ny, int nz, void *stream)
{
    if (!blocked_dev || !linear_dev || nx <= 0 || ny <= 0 || nz <= 0) return bmq::set_error(BMQ_ERR_ARG, "bmq_blocked_to_linear: bad argument");
    if (!bmq::require_device()) return BMQ_ERR_NODEVICE;
    BMQ_CK(bmq::launch_relayout((cudaStream_t)stream, true, blocked_dev, linear_dev, nx, ny, nz));
    return BMQ_OK;
}

int bmq_linear_to_blocked(const float *linear_dev, float *blocked_dev, int nx, int ny, int nz, void *stream)
{
    if (!blocked_dev || !linear_dev || nx <= 0 || ny <= 0 || nz <= 0) return bmq::set_error(BMQ_ERR_ARG, "bmq_linear_to_blocked: bad argument");
    if (!bmq::require_device()) return BMQ_ERR_NODEVICE;
    BMQ_CK(bmq::launch_relayout((cudaStream_t)stream, false, linear_dev, blocked_dev, nx, ny, nz));
    return BMQ_OK;
}

void gpu_clamp_extrema(float *field, float *fieldTemp, float *u, float *v, float *w, int ni, int nj, int nk, int dimx, int dimy,
                       int dimz, float ox, float oy, float oz, float h, float dt)
{
    if (!bmq::require_device()) return;
    bmq::Vel3 vel{u, v, w};
    k_clamp_extrema_mc<<<sgrd(ni, nj, nk), sblk()>>>(field, fieldTemp, vel, ni, nj, nk, dimx, dimy, dimz, ox, oy, oz, h, dt);
    bmq::count_launches(1);
    BMQ_CKV(cudaGetLastError());
}

void gpu_emit_smoke(float *u, float *v, float *w, float *rho, float *T, float h, int ni, int nj, int nk, float centerX,
                    float centerY, float centerZ, float radius, float density, float temperature, float emiter)
{
    if (!bmq::require_device()) return;
    k_emit_velocity<<<sgrd(ni + 1, nj, nk), sblk()>>>(u, h, ni + 1, nj, nk, centerX, centerY, centerZ, radius, emiter);
    k_emit_velocity<<<sgrd(ni, nj + 1, nk), sblk()>>>(v, h, ni, nj + 1, nk, centerX, centerY, centerZ, radius, 0.f);
    k_emit_velocity<<<sgrd(ni, nj, nk + 1), sblk()>>>(w, h, ni, nj, nk + 1, centerX, centerY, centerZ, radius, 0.f);
    k_emit_field<<<sgrd(ni, nj, nk), sblk()>>>(rho, T, h, ni, nj, nk, centerX, centerY, centerZ, radius, density, temperature);
    BMQ_CKV(cudaGetLastError());
}

void gpu_add_buoyancy(float *field, float *density, float *temperature, int ni, int nj, int nk, float alpha, float beta, float dt)
{
    if (!bmq::require_device()) return;
    const bool vec = ni % 4 == 0 && ((reinterpret_cast<uintptr_t>(field) | reinterpret_cast<uintptr_t>(density) | reinterpret_cast<uintptr_t>(temperature)) & 15) == 0;
    const long long n_scalar = (long long)ni * nj * nk;
    if (vec) k_add_buoyancy4<<<sgrd(ni / 4, nj + 1, nk), sblk()>>>(field, density, temperature, ni, nj + 1, nk, alpha, beta, dt, n_scalar);
    else k_add_buoyancy<<<sgrd(ni, nj + 1, nk), sblk()>>>(field, density, temperature, ni, nj + 1, nk, alpha, beta, dt, n_scalar);
    BMQ_CKV(cudaGetLastError());
}

void gpu_diffuse_field(float *field, float *fieldTemp0, float *fieldTemp1, int ni, int nj, int nk, int iter, float coef)
{
    if (!bmq::require_device()) return;
    const size_t bytes = sizeof(float) * (size_t)ni * nj * nk;
    float *in = fieldTemp0, *out = fieldTemp1;
    BMQ_CKV(cudaMemcpyAsync(in, field, bytes, cudaMemcpyDeviceToDevice, 0));
    for (int it = 0; it < iter; ++it) {
        k_diffuse<<<sgrd(ni, nj, nk), sblk()>>>(field, in, out, ni, nj, nk, coef);
        float *t = out; out = in; in = t;
    }
    BMQ_CKV(cudaGetLastError());
    BMQ_CKV(cudaMemcpyAsync(field, out, bytes, cudaMemcpyDeviceToDevice, 0));
}

void gpu_mad(float *field, float *field1, float *field2, float coeff1, float coeff2, int number)
{
    if (!bmq::require_device() || number <= 0) return;
    size_t blocks = ((size_t)number / 4 + 255) / 256 + 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_mad<<<(unsigned)blocks, 256>>>(field, field1, field2, coeff1, coeff2, (size_t)number);
    BMQ_CKV(cudaGetLastError());
}

}  // extern "C"
