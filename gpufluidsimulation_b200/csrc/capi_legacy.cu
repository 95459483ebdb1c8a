// The reference's 14 hot-path `extern "C" gpu_*` entry points (bimocq3D/GPU_Advection.h:26-86,97;
// definitions bimocq3D/GPU_kernel.cu:567-734, 885-890) on top of the sm_100a kernels.
// Same prototypes, same buffer/aliasing contract, legacy default stream, void returns with the
// error latched for bmq_last_error().
#include "common.h"
#include <mutex>
#include "launch3d.h"

using namespace bmq;

static inline KRange full(int n) { return KRange{0, n}; }
static const cudaStream_t kLegacy = 0;

extern "C" {

void gpu_solve_forward(float *u, float *v, float *w, float *x_fwd, float *y_fwd, float *z_fwd, float h,
                       int ni, int nj, int nk, float cfldt, float dt)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    float *const maps[1][3] = {{x_fwd, y_fwd, z_fwd}};
    BMQ_CKV(launch_forward(kLegacy, g, full(nk), u, v, w, 1, maps, cfldt, dt));
}

void gpu_solve_backwardDMC(float *u, float *v, float *w, float *x_in, float *y_in, float *z_in, float *x_out,
                           float *y_out, float *z_out, float h, int ni, int nj, int nk, float substep)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    const float *const in[1][3] = {{x_in, y_in, z_in}};
    float *const out[1][3] = {{x_out, y_out, z_out}};
    BMQ_CKV(launch_dmc(kLegacy, g, full(nk), u, v, w, 1, in, out, substep));
}

void gpu_advect_velocity(float *u, float *v, float *w, float *u_init, float *v_init, float *w_init,
                         float *backward_x, float *backward_y, float *backward_z, float h, int ni, int nj,
                         int nk, bool is_point)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    const float *const chi[3] = {backward_x, backward_y, backward_z};
    float *outs[3] = {u, v, w};
    const float *inits[3] = {u_init, v_init, w_init};
    const Stag st[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int c = 0; c < 3; ++c)
        BMQ_CKV(launch_advect(kLegacy, g, full(nk + st[c].dz), st[c], is_point, 1, &outs[c], &inits[c], chi));
}

void gpu_advect_vel_double(float *u, float *v, float *w, float *utemp, float *vtemp, float *wtemp,
                           float *backward_x, float *backward_y, float *backward_z, float *backward_xprev,
                           float *backward_yprev, float *backward_zprev, float h, int ni, int nj, int nk,
                           bool is_point, float blend_coeff)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    const float *const chi[3] = {backward_x, backward_y, backward_z};
    const float *const chip[3] = {backward_xprev, backward_yprev, backward_zprev};
    float *f[3] = {u, v, w};
    const float *p[3] = {utemp, vtemp, wtemp};
    const Stag st[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int c = 0; c < 3; ++c)
        BMQ_CKV(launch_double_advect(kLegacy, g, full(nk + st[c].dz), st[c], is_point, 1, &f[c], &p[c], chi, chip,
                                     blend_coeff));
}

void gpu_advect_field(float *field, float *field_init, float *backward_x, float *backward_y, float *backward_z,
                      float h, int ni, int nj, int nk, bool is_point)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    const float *const chi[3] = {backward_x, backward_y, backward_z};
    const float *init = field_init;
    BMQ_CKV(launch_advect(kLegacy, g, full(nk), Stag{0, 0, 0}, is_point, 1, &field, &init, chi));
}

void gpu_advect_field_double(float *field, float *field_prev, float *backward_x, float *backward_y,
                             float *backward_z, float *backward_xprev, float *backward_yprev,
                             float *backward_zprev, float h, int ni, int nj, int nk, bool is_point,
                             float blend_coeff)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    const float *const chi[3] = {backward_x, backward_y, backward_z};
    const float *const chip[3] = {backward_xprev, backward_yprev, backward_zprev};
    const float *p = field_prev;
    BMQ_CKV(launch_double_advect(kLegacy, g, full(nk), Stag{0, 0, 0}, is_point, 1, &field, &p, chi, chip,
                                 blend_coeff));
}

void gpu_accumulate_velocity(float *u_change, float *v_change, float *w_change, float *du_init, float *dv_init,
                             float *dw_init, float *forward_x, float *forward_y, float *forward_z, float h,
                             int ni, int nj, int nk, bool is_point, float coeff)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    const float *const psi[3] = {forward_x, forward_y, forward_z};
    float *t[3] = {du_init, dv_init, dw_init};
    const float *c[3] = {u_change, v_change, w_change};
    const Stag st[3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int q = 0; q < 3; ++q)
        BMQ_CKV(launch_cumulate(kLegacy, g, full(nk + st[q].dz), st[q], is_point, 1, 1, &t[q], &c[q], &coeff, psi));
}

void gpu_accumulate_field(float *field_change, float *dfield_init, float *forward_x, float *forward_y,
                          float *forward_z, float h, int ni, int nj, int nk, bool is_point, float coeff)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    const float *const psi[3] = {forward_x, forward_y, forward_z};
    const float *c = field_change;
    BMQ_CKV(launch_cumulate(kLegacy, g, full(nk), Stag{0, 0, 0}, is_point, 1, 1, &dfield_init, &c, &coeff, psi));
}

void gpu_estimate_distortion(float *du, float *x_back, float *y_back, float *z_back, float *x_fwd, float *y_fwd,
                             float *z_fwd, float h, int ni, int nj, int nk)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    const float *const b[1][3] = {{x_back, y_back, z_back}};
    const float *const f[1][3] = {{x_fwd, y_fwd, z_fwd}};
    float *dist[1] = {du};
    BMQ_CKV(launch_estimate(kLegacy, g, full(nk), 1, b, f, dist, nullptr, nullptr, nullptr));
}

void gpu_add(float *field1, float *field2, float coeff, int number)
{
    if (!require_device()) return;
    BMQ_CKV(launch_axpy(kLegacy, field1, field2, coeff, (size_t)(number > 0 ? number : 0)));
}

void gpu_add_field(float *out, float *field1, float *field2, float coeff, int number)
{
    if (!require_device()) return;
    BMQ_CKV(launch_add_field(kLegacy, out, field1, field2, coeff, (size_t)(number > 0 ? number : 0)));
}

// One component of gpu_compensate_* (GPU_kernel.cu:652-665): error via psi, copy, apply via chi, clamp.
static int compensate_one(const Grid3 &g, Stag st, bool is_point, float *f, float *df, float *f_src,
                          const float *const psi[3], const float *const chi[3])
{
    const int fi = g.ni + st.dx, fj = g.nj + st.dy, fk = g.nk + st.dz;
    KRange r = full(fk);
    const float *src = f, *init = df;
    BMQ_CK(launch_error(kLegacy, g, r, st, is_point, 1, &f_src, &src, &init, psi));
    BMQ_CK(cudaMemcpyAsync(df, f, sizeof(float) * (size_t)fi * fj * fk, cudaMemcpyDeviceToDevice, kLegacy));
    const float *e0 = f_src;
    const float mhalf = -0.5f;
    BMQ_CK(launch_cumulate(kLegacy, g, r, st, is_point, 1, 1, &f, &e0, &mhalf, chi));
    BMQ_CK(launch_clamp_extrema(kLegacy, fi, fj, fk, r, df, f));
    return BMQ_OK;
}

void gpu_compensate_velocity(float *u, float *v, float *w, float *du, float *dv, float *dw, float *u_src,
                             float *v_src, float *w_src, float *forward_x, float *forward_y, float *forward_z,
                             float *backward_x, float *backward_y, float *backward_z, float h, int ni, int nj,
                             int nk, bool is_point)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    const float *const psi[3] = {forward_x, forward_y, forward_z};
    const float *const chi[3] = {backward_x, backward_y, backward_z};
    if (compensate_one(g, Stag{1, 0, 0}, is_point, u, du, u_src, psi, chi) != BMQ_OK) return;
    if (compensate_one(g, Stag{0, 1, 0}, is_point, v, dv, v_src, psi, chi) != BMQ_OK) return;
    compensate_one(g, Stag{0, 0, 1}, is_point, w, dw, w_src, psi, chi);
}

void gpu_compensate_field(float *u, float *du, float *u_src, float *forward_x, float *forward_y,
                          float *forward_z, float *backward_x, float *backward_y, float *backward_z, float h,
                          int ni, int nj, int nk, bool is_point)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    const float *const psi[3] = {forward_x, forward_y, forward_z};
    const float *const chi[3] = {backward_x, backward_y, backward_z};
    compensate_one(g, Stag{0, 0, 0}, is_point, u, du, u_src, psi, chi);
}

void gpu_semilag(float *field, float *field_src, float *u, float *v, float *w, int dim_x, int dim_y, int dim_z,
                 float h, int ni, int nj, int nk, float cfldt, float dt)
{
    if (!require_device()) return;
    Grid3 g = make_grid(ni, nj, nk, h);
    const float *src = field_src;
    BMQ_CKV(launch_semilag(kLegacy, g, full(nk + dim_z), Stag{dim_x, dim_y, dim_z}, u, v, w, 1, &field, &src, cfldt, dt));
}

// max(|u|, |v|, |w|) on the device, legacy default stream, result copied to *host_out.  Replaces the
// serial host loops of getCFL (BimocqGPUSolver.cpp:348-373, BimocqSolver.cpp:1067-1118), which walk
// host copies of the three face fields; the caller applies the 1e-4 floor and h / max itself.
int bmq_max_abs3(const float *u, long long nu, const float *v, long long nv, const float *w, long long nw, float *host_out)
{
    if (!host_out || nu < 0 || nv < 0 || nw < 0) return set_error(BMQ_ERR_ARG, "bmq_max_abs3: bad argument");
    if (!require_device()) return BMQ_ERR_NODEVICE;
    // 4-byte device scalar per device, allocated once; the legacy entry points are single-threaded per device
    // (they share the legacy default stream and, like the reference's gpuMapper, caller-owned scratch)
    static std::mutex mu;
    static float *scratch[64] = {};
    int dev = 0;
    BMQ_CK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return set_error(BMQ_ERR_ARG, "bmq_max_abs3: device ordinal %d out of range", dev);
    float *d_red = nullptr;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!scratch[dev]) BMQ_CK(cudaMalloc(&scratch[dev], sizeof(float)));
        d_red = scratch[dev];
    }
    BMQ_CK(cudaMemsetAsync(d_red, 0, sizeof(float), kLegacy));
    BMQ_CK(launch_maxabs3(kLegacy, u, (size_t)nu, v, (size_t)nv, w, (size_t)nw, d_red));
    BMQ_CK(cudaMemcpy(host_out, d_red, sizeof(float), cudaMemcpyDeviceToHost));
    return BMQ_OK;
}

}  // extern "C"
