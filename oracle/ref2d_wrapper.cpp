// TEST INFRASTRUCTURE ONLY.  C wrapper around the REFERENCE'S OWN 2D solver class
// (/root/reference/src/bimocq2D/BimocqSolver2D.{h,cpp}), compiled unmodified next to this file
// by oracle/Makefile into oracle/_ref/libref2d.so.  Nothing from the reference is copied here:
// this file only calls the reference's public methods and exposes its public Array2f members.
//
// ref2d_phase_a / ref2d_phase_b restate the ORCHESTRATION lines of advanceBIMOCQ
// (BimocqSolver2D.cpp:390-508) around the reference's own hot-path methods, with the two
// non-advection calls in the middle (applyBuoyancyForce :447, projection :454) replaced by
// caller-supplied "velocity after forces" and "final" fields, so that the advection path can be
// driven with any forcing.
#include "BimocqSolver2D.h"

#include <cstring>
#include <map>
#include <string>

namespace {
struct Ref2D {
    BimocqSolver2D *s;
    Array2f u_presave, v_presave, u_save, v_save, rho_save, T_save;
    float cond_vel = 0, cond_rho = 0, maxvel = 0;
    int vel_remap = 0, rho_remap = 0;
};

Array2f *member(BimocqSolver2D *s, const std::string &n)
{
#define M(x) if (n == #x) return &s->x;
    M(u) M(v) M(u_temp) M(v_temp) M(rho) M(temperature)
    M(forward_x) M(forward_y) M(forward_scalar_x) M(forward_scalar_y)
    M(backward_x) M(backward_y) M(backward_xprev) M(backward_yprev)
    M(backward_scalar_x) M(backward_scalar_y) M(backward_scalar_xprev) M(backward_scalar_yprev)
    M(u_init) M(v_init) M(u_origin) M(v_origin) M(du) M(dv) M(du_prev) M(dv_prev)
    M(du_temp) M(dv_temp) M(du_proj) M(dv_proj) M(drho) M(drho_temp) M(drho_prev) M(dT) M(dT_temp) M(dT_prev)
    M(rho_init) M(rho_orig) M(T_init) M(T_orig)
#undef M
    return nullptr;
}
}  // namespace

extern "C" {

void *ref2d_create(int nx, int ny, float L, float blend)
{
    Ref2D *r = new Ref2D;
    r->s = new BimocqSolver2D(nx, ny, L, blend, /*N particles per cell edge (unused by BIMOCQ; 0 divides by zero in seedParticles)*/ 1, /*neumann*/ true, BIMOCQ);
    r->s->alpha = 0.f;
    r->s->beta = 0.f;
    return r;
}

void ref2d_destroy(void *p)
{
    Ref2D *r = (Ref2D *)p;
    delete r->s;
    delete r;
}

// pointer to a member's storage (row-major a[i + ni*j]); dims returned
float *ref2d_field(void *p, const char *name, int *ni, int *nj)
{
    Array2f *a = member(((Ref2D *)p)->s, name);
    if (!a) return nullptr;
    if (ni) *ni = a->ni;
    if (nj) *nj = a->nj;
    return a->a.data;
}

float ref2d_h(void *p) { return ((Ref2D *)p)->s->h; }
void ref2d_set_levelset(void *p, int on) { ((Ref2D *)p)->s->advect_levelset = on != 0; }
void ref2d_set_counters(void *p, int lastremeshing, int rho_lastremeshing)
{
    ((Ref2D *)p)->s->lastremeshing = lastremeshing;
    ((Ref2D *)p)->s->rho_lastremeshing = rho_lastremeshing;
}
void ref2d_get_counters(void *p, int *out)
{
    BimocqSolver2D *s = ((Ref2D *)p)->s;
    out[0] = s->lastremeshing; out[1] = s->rho_lastremeshing;
    out[2] = s->total_resampleCount; out[3] = s->total_scalar_resample;
    out[4] = ((Ref2D *)p)->vel_remap; out[5] = ((Ref2D *)p)->rho_remap;
}
void ref2d_get_scalars(void *p, float *out)
{
    Ref2D *r = (Ref2D *)p;
    out[0] = r->s->_cfl; out[1] = r->cond_vel; out[2] = r->cond_rho; out[3] = r->maxvel;
}

// ---- single reference methods, for kernel-level tests
float ref2d_max_vel(void *p) { return ((Ref2D *)p)->s->maxVel(); }
void ref2d_get_cfl(void *p) { ((Ref2D *)p)->s->getCFL(); }
void ref2d_update_forward(void *p, float dt, int scalar)
{
    BimocqSolver2D *s = ((Ref2D *)p)->s;
    if (scalar) s->updateForward(dt, s->forward_scalar_x, s->forward_scalar_y);
    else s->updateForward(dt, s->forward_x, s->forward_y);
}
void ref2d_update_backward(void *p, float dt, int scalar)
{
    BimocqSolver2D *s = ((Ref2D *)p)->s;
    if (scalar) s->updateBackward(dt, s->backward_scalar_x, s->backward_scalar_y);
    else s->updateBackward(dt, s->backward_x, s->backward_y);
}
// dst (caller array of the field's size) = semiLagAdvect(member `name`)
void ref2d_semilag(void *p, const char *name, float dt, float *dst)
{
    BimocqSolver2D *s = ((Ref2D *)p)->s;
    Array2f *src = member(s, name);
    Array2f out;
    out.resize(src->ni, src->nj, 0.0f);
    float ox = (src->ni == s->ni + 1) ? 0.0f : 0.5f, oy = (src->nj == s->nj + 1) ? 0.0f : 0.5f;
    s->semiLagAdvect(*src, out, dt, src->ni, src->nj, ox, oy);
    std::memcpy(dst, out.a.data, sizeof(float) * src->ni * src->nj);
}
float ref2d_estimate_distortion(void *p, int scalar)
{
    BimocqSolver2D *s = ((Ref2D *)p)->s;
    return scalar ? s->estimateDistortion(s->backward_scalar_x, s->backward_scalar_y, s->forward_scalar_x, s->forward_scalar_y)
                  : s->estimateDistortion(s->backward_x, s->backward_y, s->forward_x, s->forward_y);
}

// ---- advanceBIMOCQ, lines 394-445 (everything before applyBuoyancyForce)
void ref2d_phase_a(void *p, float dt, int currentframe)
{
    Ref2D *r = (Ref2D *)p;
    BimocqSolver2D *s = r->s;
    const int ni = s->ni, nj = s->nj;
    s->getCFL();
    if (currentframe != 0 && !s->advect_levelset) {
        s->u = s->u_temp;
        s->v = s->v_temp;
    }
    s->frameCount++;
    s->resampleCount++;
    if (!s->advect_levelset) {
        s->updateForward(dt, s->forward_x, s->forward_y);
        s->updateBackward(dt, s->backward_x, s->backward_y);
    }
    s->updateForward(dt, s->forward_scalar_x, s->forward_scalar_y);
    s->updateBackward(dt, s->backward_scalar_x, s->backward_scalar_y);
    Array2f semi_u, semi_v, semi_rho, semi_T;
    semi_u.resize(ni + 1, nj, 0.0);
    semi_v.resize(ni, nj + 1, 0.0);
    semi_rho.resize(ni, nj, 0.0);
    semi_T.resize(ni, nj, 0.0);
    s->semiLagAdvect(s->rho, semi_rho, dt, ni, nj, 0.5, 0.5);
    s->semiLagAdvect(s->temperature, semi_T, dt, ni, nj, 0.5, 0.5);
    s->semiLagAdvect(s->u, semi_u, dt, ni + 1, nj, 0.0, 0.5);
    s->semiLagAdvect(s->v, semi_v, dt, ni, nj + 1, 0.5, 0.0);
    r->u_presave = s->u;
    r->v_presave = s->v;
    if (!s->advect_levelset) {
        s->advectVelocity(semi_u, semi_v);
        s->correctVelocity(semi_u, semi_v);
    }
    s->advectScalars(semi_rho, semi_T);
    if (!s->advect_levelset) s->correctScalars(semi_rho, semi_T);
    r->u_save = s->u;
    r->v_save = s->v;
    r->rho_save = s->rho;
    r->T_save = s->temperature;
}

// ---- advanceBIMOCQ, lines 447-507 with forces + projection supplied by the caller:
// (u,v)_forced = velocity after external forces, *_final = fields after projection.
void ref2d_phase_b(void *p, float dt, int currentframe, const float *u_forced, const float *v_forced,
                   const float *u_final, const float *v_final, const float *rho_final, const float *T_final)
{
    Ref2D *r = (Ref2D *)p;
    BimocqSolver2D *s = r->s;
    const int ni = s->ni, nj = s->nj;
    float proj_coeff = 2.0;
    std::memcpy(s->u.a.data, u_forced, sizeof(float) * (ni + 1) * nj);
    std::memcpy(s->v.a.data, v_forced, sizeof(float) * ni * (nj + 1));
    s->du_temp = s->u; s->du_temp -= r->u_save;
    s->dv_temp = s->v; s->dv_temp -= r->v_save;
    r->u_save = s->u;
    r->v_save = s->v;
    std::memcpy(s->u.a.data, u_final, sizeof(float) * (ni + 1) * nj);
    std::memcpy(s->v.a.data, v_final, sizeof(float) * ni * (nj + 1));
    std::memcpy(s->rho.a.data, rho_final, sizeof(float) * ni * nj);
    std::memcpy(s->temperature.a.data, T_final, sizeof(float) * ni * nj);
    float d_vel = s->estimateDistortion(s->backward_x, s->backward_y, s->forward_x, s->forward_y);
    float d_scalar = s->estimateDistortion(s->backward_scalar_x, s->backward_scalar_y, s->forward_scalar_x, s->forward_scalar_y);
    float vel = s->maxVel();
    r->maxvel = vel;
    r->cond_vel = d_vel / (vel * dt);
    r->cond_rho = d_scalar / (vel * dt);
    bool vel_remapping = ((d_vel / (vel * dt)) > 1.0 || currentframe - s->lastremeshing >= 8);
    bool rho_remapping = ((d_scalar / (vel * dt)) > 1.0 || currentframe - s->rho_lastremeshing >= 20);
    if (vel_remapping) proj_coeff = 1.0;
    if (!s->advect_levelset) {
        s->du_proj = s->u; s->du_proj -= r->u_save;
        s->dv_proj = s->v; s->dv_proj -= r->v_save;
        s->drho_temp = s->rho; s->drho_temp -= r->rho_save;
        s->dT_temp = s->temperature; s->dT_temp -= r->T_save;
        s->accumulateVelocity(s->du_temp, s->dv_temp, 1.0, false);
        s->accumulateVelocity(s->du_proj, s->dv_proj, proj_coeff, false);
        s->accumulateScalars(s->drho_temp, s->dT_temp, false);
    }
    r->vel_remap = r->rho_remap = 0;
    if (vel_remapping && !s->advect_levelset) {
        s->lastremeshing = currentframe;
        s->resampleVelBuffer(dt);
        s->accumulateVelocity(s->du_proj, s->dv_proj, proj_coeff, false);
        r->vel_remap = 1;
    }
    if (rho_remapping) {
        s->rho_lastremeshing = currentframe;
        s->resampleRhoBuffer(dt);
        r->rho_remap = 1;
    }
    s->u_temp = s->u;
    s->v_temp = s->v;
    if (currentframe != 0) {
        for (int t = 0; t < (ni + 1) * nj; ++t) s->u.a[t] = 0.5 * (r->u_presave.a[t] + s->u.a[t]);
        for (int t = 0; t < ni * (nj + 1); ++t) s->v.a[t] = 0.5 * (r->v_presave.a[t] + s->v.a[t]);
    }
}

}  // extern "C"
