// Handle API (bmq3d_*): device-resident BiMocq^2 advection state, the fused stage sequence and
// the reinitialisation scheduler of BimocqSolver::advanceBimocq (bimocq3D/BimocqSolver.cpp:88-230)
// with MapperBase semantics (bimocq3D/Mapping.cpp:7-271).
//
// Memory: every field is one cudaMalloc of its stored planes plus one plane + one row of zero
// padding (the reference's trilinear sampler reads one node past the clamp bound with weight 0,
// GPU_kernel.cu:53-61 with the clamps at :356/:419; the padding keeps those reads in bounds and
// finite).  A z-slab rank stores planes [k_own0-halo, k_own1+halo+1) clipped to the field and
// hands the kernels virtual base pointers, so all indices are global.
//
// Re-initialisation is pointer rotation: chi_prev <-> chi, f_prev <-> f_init, one identity fill
// and one copy current -> init per field, instead of the reference's 15 full-field copies
// (Mapping.cpp:430-447, BimocqSolver.cpp:1433-1451).
#include "common.h"
#include "launch3d.h"
#include "mg_signal.h"

#include <cmath>
#include <cstring>
#include <unistd.h>
#include <new>
#include <utility>
#include <vector>

using namespace bmq;

namespace {

enum { N_SCRATCH = 10, SCRATCH_BASE = 64, N_TMPMAP = 6 };

struct Field {
    int alloc_id = -1;        // identity of the allocation: stays with the buffer when fields rotate (z-slab peers index by it)
    float *alloc = nullptr;   // first stored plane
    int nx = 0, ny = 0, nz = 0;   // full (global) dims of the field
    int p0 = 0, p1 = 0;       // stored global planes [p0, p1)
    size_t plane() const { return (size_t)nx * ny; }
    size_t stored_elems() const { return plane() * (size_t)(p1 - p0); }
    float *vbase() const { return alloc ? alloc - (ptrdiff_t)plane() * p0 : nullptr; }
};

}  // namespace

struct bmq3d_solver {
    int ni, nj, nk;
    float h, blend;
    int k0, k1, halo;          // owned planes [k0,k1), halo width
    int n_allocs = 0;          // allocation ids handed out so far (Field::alloc_id)
    cudaStream_t stream = 0;
    Grid3 g;
    Field f[BMQ_F_COUNT];
    Field scratch[N_SCRATCH];  // adv[5], err[5]
    Field tmpmap[N_TMPMAP];    // DMC ping-pong: velocity chi xyz, scalar chi xyz
    float *d_red = nullptr;    // device reduction scalars [8]
    float *h_red = nullptr;    // pinned mirror
    // scheduler state (BimocqSolver.h:142-143, Mapping.h:36)
    int vel_last_reinit = 0, scalar_last_reinit = 0;
    int vel_reinit_count = 0, scalar_reinit_count = 0;
    float max_v = 0.f, cfldt = 0.f, proj_coeff = 2.f;
    bool vel_reinit = false, scalar_reinit = false;
    bmq3d_stats stats;
    bool semi_alloc = false;
    // copy streams + events of the host-buffer path (bmq3d_*_host): transfers overlap the stages
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    // host buffers in the reference's 8^3-blocked buffer3Df layout: raw transfers + relayout on the device
    int host_layout = BMQ_LAYOUT_LINEAR;
    float *stage_up = nullptr, *stage_down = nullptr;   // blocked staging, one per copy direction
    cudaEvent_t ev_copy[16] = {};
    // optional per-stage CUDA-event timing (bmq3d_timing_*): pairs recorded on `stream`
    bool timing = false;
    struct Span { int slot; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
};

namespace {

Stag stag_of(int id)
{
    switch (id) {
    case BMQ_F_U: case BMQ_F_U_INIT: case BMQ_F_U_PREV: case BMQ_F_DU_EXT: case BMQ_F_DU_PROJ: case BMQ_F_U_SEMI:
    case BMQ_F_U_ADV: case BMQ_F_U_ERR:
        return Stag{1, 0, 0};
    case BMQ_F_V: case BMQ_F_V_INIT: case BMQ_F_V_PREV: case BMQ_F_DV_EXT: case BMQ_F_DV_PROJ: case BMQ_F_V_SEMI:
    case BMQ_F_V_ADV: case BMQ_F_V_ERR:
        return Stag{0, 1, 0};
    case BMQ_F_W: case BMQ_F_W_INIT: case BMQ_F_W_PREV: case BMQ_F_DW_EXT: case BMQ_F_DW_PROJ: case BMQ_F_W_SEMI:
    case BMQ_F_W_ADV: case BMQ_F_W_ERR:
        return Stag{0, 0, 1};
    default:
        return Stag{0, 0, 0};
    }
}

int alloc_field(bmq3d_solver *s, Field &fd, Stag st)
{
    fd.nx = s->ni + st.dx;
    fd.ny = s->nj + st.dy;
    fd.nz = s->nk + st.dz;
    fd.p0 = s->k0 - s->halo > 0 ? s->k0 - s->halo : 0;
    fd.p1 = s->k1 + s->halo + 1 < fd.nz ? s->k1 + s->halo + 1 : fd.nz;
    const size_t n = fd.stored_elems() + fd.plane() + fd.nx + 2;   // zero padding, see file header
    if (fd.alloc_id < 0) fd.alloc_id = s->n_allocs++;
    BMQ_CK(cudaMalloc(&fd.alloc, n * sizeof(float)));
    BMQ_CK(cudaMemsetAsync(fd.alloc, 0, n * sizeof(float), s->stream));
    return BMQ_OK;
}

Field *field_of(bmq3d_solver *s, int id)
{
    if (id >= 0 && id < BMQ_F_COUNT) return &s->f[id];
    if (id >= SCRATCH_BASE && id < SCRATCH_BASE + N_SCRATCH) return &s->scratch[id - SCRATCH_BASE];
    if (id >= BMQ_F_TMPMAP0 && id <= BMQ_F_TMPMAP5) return &s->tmpmap[id - BMQ_F_TMPMAP0];
    return nullptr;
}

// owned compute range of a field with staggering dz: [k0,k1) plus the top face on the last slab
KRange own(const bmq3d_solver *s, int dz) { return KRange{s->k0, s->k1 + ((dz && s->k1 == s->nk) ? 1 : 0)}; }
// planes a whole-field fill should touch (everything stored)
KRange stored(const Field &fd) { return KRange{fd.p0, fd.p1}; }

int ensure_semi(bmq3d_solver *s)
{
    if (s->semi_alloc) return BMQ_OK;
    for (int id = BMQ_F_U_SEMI; id <= BMQ_F_T_SEMI; ++id) {
        int st = alloc_field(s, s->f[id], stag_of(id));
        if (st != BMQ_OK) return st;
    }
    s->semi_alloc = true;
    return BMQ_OK;
}

int identity_fill(bmq3d_solver *s, int first_id)
{
    float *const sets[1][3] = {{s->f[first_id].vbase(), s->f[first_id + 1].vbase(), s->f[first_id + 2].vbase()}};
    BMQ_CK(launch_identity(s->stream, s->g, stored(s->f[first_id]), 1, sets));
    return BMQ_OK;
}
int identity_fill_tmp(bmq3d_solver *s, int first)
{
    float *const sets[1][3] = {{s->tmpmap[first].vbase(), s->tmpmap[first + 1].vbase(), s->tmpmap[first + 2].vbase()}};
    BMQ_CK(launch_identity(s->stream, s->g, stored(s->tmpmap[first]), 1, sets));
    return BMQ_OK;
}

int copy_field(bmq3d_solver *s, Field &dst, const Field &src)
{
    BMQ_CK(cudaMemcpyAsync(dst.alloc, src.alloc, src.stored_elems() * sizeof(float), cudaMemcpyDeviceToDevice,
                           s->stream));
    return BMQ_OK;
}

// ---- per-stage timing ----------------------------------------------------------------------
const char *const kSlotNames[BMQ_T_COUNT] = {
    "maxvel", "dmc_backward", "forward", "semilag", "advect_velocity", "error_velocity", "apply_velocity",
    "blend_velocity", "advect_scalars", "error_scalars", "apply_scalars", "blend_scalars", "distortion",
    "accumulate_velocity", "accumulate_scalars", "reinit"};

cudaEvent_t take_event(bmq3d_solver *s)
{
    cudaEvent_t e = nullptr;
    if (!s->event_pool.empty()) { e = s->event_pool.back(); s->event_pool.pop_back(); }
    else cudaEventCreate(&e);
    return e;
}

struct StageTimer {
    bmq3d_solver *s; int idx = -1;
    StageTimer(bmq3d_solver *s_, int slot) : s(s_)
    {
        if (!s->timing) return;
        bmq3d_solver::Span sp{slot, take_event(s), take_event(s)};
        cudaEventRecord(sp.a, s->stream);
        s->spans.push_back(sp);
        idx = (int)s->spans.size() - 1;
    }
    ~StageTimer() { if (idx >= 0) cudaEventRecord(s->spans[idx].b, s->stream); }
};

// ---- stages ------------------------------------------------------------------------------

// out == nullptr: the result stays in s->d_red[0] (the z-slab driver reduces it over the ranks on the device)
int stage_maxvel(bmq3d_solver *s, float *out)
{
    StageTimer _t(s, BMQ_T_MAXVEL);
    BMQ_CK(cudaMemsetAsync(s->d_red, 0, sizeof(float), s->stream));
    // owned planes only, so a slab's halo copies are not double counted (harmless for a max anyway)
    const Field &u = s->f[BMQ_F_U], &v = s->f[BMQ_F_V], &w = s->f[BMQ_F_W];
    KRange ru = own(s, 0), rw = own(s, 1);
    BMQ_CK(launch_maxabs3(s->stream, u.vbase() + u.plane() * ru.kbeg, u.plane() * (ru.kend - ru.kbeg),
                          v.vbase() + v.plane() * ru.kbeg, v.plane() * (ru.kend - ru.kbeg),
                          w.vbase() + w.plane() * rw.kbeg, w.plane() * (rw.kend - rw.kbeg), s->d_red));
    if (!out) return BMQ_OK;
    BMQ_CK(cudaMemcpyAsync(s->h_red, s->d_red, sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    BMQ_CK(cudaStreamSynchronize(s->stream));
    *out = s->h_red[0];
    return BMQ_OK;
}

// getCFL (BimocqSolver.cpp:1093-1117) + the frame-0 rule (:94)
void set_cfl(bmq3d_solver *s, int framenum, float max_abs)
{
    float mv = 1e-4f;
    if (max_abs > mv) mv = max_abs;
    s->cfldt = s->h / mv;
    s->max_v = framenum == 0 ? s->h : mv;
    s->stats.max_v = s->max_v;
    s->stats.cfldt = s->cfldt;
}

int stage_dmc(bmq3d_solver *s, float substep)
{
    StageTimer _t(s, BMQ_T_DMC);
    const float *const in[2][3] = {
        {s->f[BMQ_F_VBWD_X].vbase(), s->f[BMQ_F_VBWD_Y].vbase(), s->f[BMQ_F_VBWD_Z].vbase()},
        {s->f[BMQ_F_SBWD_X].vbase(), s->f[BMQ_F_SBWD_Y].vbase(), s->f[BMQ_F_SBWD_Z].vbase()}};
    float *const out[2][3] = {{s->tmpmap[0].vbase(), s->tmpmap[1].vbase(), s->tmpmap[2].vbase()},
                              {s->tmpmap[3].vbase(), s->tmpmap[4].vbase(), s->tmpmap[5].vbase()}};
    BMQ_CK(launch_dmc(s->stream, s->g, own(s, 0), s->f[BMQ_F_U].vbase(), s->f[BMQ_F_V].vbase(),
                      s->f[BMQ_F_W].vbase(), 2, in, out, substep));
    // ping-pong instead of the reference's three device-to-device copies (GPU_Advection.h:466-468)
    for (int c = 0; c < 3; ++c) {
        std::swap(s->f[BMQ_F_VBWD_X + c], s->tmpmap[c]);
        std::swap(s->f[BMQ_F_SBWD_X + c], s->tmpmap[3 + c]);
    }
    return BMQ_OK;
}

int stage_forward(bmq3d_solver *s, float dt)
{
    StageTimer _t(s, BMQ_T_FORWARD);
    float *const maps[2][3] = {
        {s->f[BMQ_F_VFWD_X].vbase(), s->f[BMQ_F_VFWD_Y].vbase(), s->f[BMQ_F_VFWD_Z].vbase()},
        {s->f[BMQ_F_SFWD_X].vbase(), s->f[BMQ_F_SFWD_Y].vbase(), s->f[BMQ_F_SFWD_Z].vbase()}};
    BMQ_CK(launch_forward(s->stream, s->g, own(s, 0), s->f[BMQ_F_U].vbase(), s->f[BMQ_F_V].vbase(),
                          s->f[BMQ_F_W].vbase(), 2, maps, s->cfldt, dt));
    return BMQ_OK;
}

// BimocqSolver::semilagAdvect (BimocqSolver.cpp:645-668): traces back with -dt
int stage_semilag(bmq3d_solver *s, float dt)
{
    StageTimer _t(s, BMQ_T_SEMILAG);
    int st = ensure_semi(s);
    if (st != BMQ_OK) return st;
    const float *u = s->f[BMQ_F_U].vbase(), *v = s->f[BMQ_F_V].vbase(), *w = s->f[BMQ_F_W].vbase();
    for (int c = 0; c < 3; ++c) {
        float *o = s->f[BMQ_F_U_SEMI + c].vbase();
        const float *src = s->f[BMQ_F_U + c].vbase();
        Stag sg = stag_of(BMQ_F_U + c);
        BMQ_CK(launch_semilag(s->stream, s->g, own(s, sg.dz), sg, u, v, w, 1, &o, &src, s->cfldt, -dt));
    }
    float *o2[2] = {s->f[BMQ_F_RHO_SEMI].vbase(), s->f[BMQ_F_T_SEMI].vbase()};
    const float *s2[2] = {s->f[BMQ_F_RHO].vbase(), s->f[BMQ_F_T].vbase()};
    BMQ_CK(launch_semilag(s->stream, s->g, own(s, 0), Stag{0, 0, 0}, u, v, w, 2, o2, s2, s->cfldt, -dt));
    return BMQ_OK;
}

void map_ptrs(bmq3d_solver *s, int first, const float *out[3])
{
    for (int c = 0; c < 3; ++c) out[c] = s->f[first + c].vbase();
}

// which: 0 velocity (three staggered components, one launch each), 1 scalars (rho+T in one launch)
int stage_advect(bmq3d_solver *s, int which)
{
    StageTimer _t(s, which == 0 ? BMQ_T_ADVECT_V : BMQ_T_ADVECT_S);
    const float *chi[3];
    if (which == 0) {
        map_ptrs(s, BMQ_F_VBWD_X, chi);
        for (int c = 0; c < 3; ++c) {
            float *o = s->scratch[c].vbase();
            const float *init = s->f[BMQ_F_U_INIT + c].vbase();
            Stag sg = stag_of(BMQ_F_U + c);
            BMQ_CK(launch_advect(s->stream, s->g, own(s, sg.dz), sg, false, 1, &o, &init, chi));
        }
    } else {
        map_ptrs(s, BMQ_F_SBWD_X, chi);
        float *o[2] = {s->scratch[3].vbase(), s->scratch[4].vbase()};
        const float *init[2] = {s->f[BMQ_F_RHO_INIT].vbase(), s->f[BMQ_F_T_INIT].vbase()};
        BMQ_CK(launch_advect(s->stream, s->g, own(s, 0), Stag{0, 0, 0}, false, 2, o, init, chi));
    }
    return BMQ_OK;
}

int stage_error(bmq3d_solver *s, int which)
{
    StageTimer _t(s, which == 0 ? BMQ_T_ERROR_V : BMQ_T_ERROR_S);
    const float *psi[3];
    if (which == 0) {
        map_ptrs(s, BMQ_F_VFWD_X, psi);
        for (int c = 0; c < 3; ++c) {
            float *e = s->scratch[5 + c].vbase();
            const float *src = s->scratch[c].vbase();
            const float *init = s->f[BMQ_F_U_INIT + c].vbase();
            Stag sg = stag_of(BMQ_F_U + c);
            BMQ_CK(launch_error(s->stream, s->g, own(s, sg.dz), sg, false, 1, &e, &src, &init, psi));
        }
    } else {
        map_ptrs(s, BMQ_F_SFWD_X, psi);
        float *e[2] = {s->scratch[8].vbase(), s->scratch[9].vbase()};
        const float *src[2] = {s->scratch[3].vbase(), s->scratch[4].vbase()};
        const float *init[2] = {s->f[BMQ_F_RHO_INIT].vbase(), s->f[BMQ_F_T_INIT].vbase()};
        BMQ_CK(launch_error(s->stream, s->g, own(s, 0), Stag{0, 0, 0}, false, 2, e, src, init, psi));
    }
    return BMQ_OK;
}

// comp >= 0 (velocity only): that component alone, so that the host-buffer path can send it home
// while the next one is computed
int stage_apply(bmq3d_solver *s, int which, int comp);
int stage_apply(bmq3d_solver *s, int which) { return stage_apply(s, which, -1); }
int stage_apply(bmq3d_solver *s, int which, int comp)
{
    StageTimer _t(s, which == 0 ? BMQ_T_APPLY_V : BMQ_T_APPLY_S);
    const float *chi[3];
    if (which == 0) {
        map_ptrs(s, BMQ_F_VBWD_X, chi);
        for (int c = 0; c < 3; ++c) {
            if (comp >= 0 && c != comp) continue;
            float *o = s->f[BMQ_F_U + c].vbase();
            const float *adv = s->scratch[c].vbase();
            const float *e = s->scratch[5 + c].vbase();
            Stag sg = stag_of(BMQ_F_U + c);
            BMQ_CK(launch_apply_clamp(s->stream, s->g, own(s, sg.dz), sg, false, 1, &o, &adv, &e, chi));
        }
    } else {
        map_ptrs(s, BMQ_F_SBWD_X, chi);
        float *o[2] = {s->f[BMQ_F_RHO].vbase(), s->f[BMQ_F_T].vbase()};
        const float *adv[2] = {s->scratch[3].vbase(), s->scratch[4].vbase()};
        const float *e[2] = {s->scratch[8].vbase(), s->scratch[9].vbase()};
        BMQ_CK(launch_apply_clamp(s->stream, s->g, own(s, 0), Stag{0, 0, 0}, false, 2, o, adv, e, chi));
    }
    return BMQ_OK;
}

// Two-level blend (Mapping.cpp:196-201, 228-233).  The reference launches doubleAdvect_kernel
// even with blend 1, where it computes f*1 + 0*p = f; that launch is skipped here.
int stage_blend(bmq3d_solver *s, int which)
{
    StageTimer _t(s, which == 0 ? BMQ_T_BLEND_V : BMQ_T_BLEND_S);
    const int count = which == 0 ? s->vel_reinit_count : s->scalar_reinit_count;
    if (count == 0 || s->blend == 1.0f) return BMQ_OK;
    const float *chi[3], *chip[3];
    if (which == 0) {
        map_ptrs(s, BMQ_F_VBWD_X, chi);
        map_ptrs(s, BMQ_F_VBWDP_X, chip);
        for (int c = 0; c < 3; ++c) {
            float *fl = s->f[BMQ_F_U + c].vbase();
            const float *pv = s->f[BMQ_F_U_PREV + c].vbase();
            Stag sg = stag_of(BMQ_F_U + c);
            BMQ_CK(launch_double_advect(s->stream, s->g, own(s, sg.dz), sg, false, 1, &fl, &pv, chi, chip, s->blend));
        }
    } else {
        map_ptrs(s, BMQ_F_SBWD_X, chi);
        map_ptrs(s, BMQ_F_SBWDP_X, chip);
        float *fl[2] = {s->f[BMQ_F_RHO].vbase(), s->f[BMQ_F_T].vbase()};
        const float *pv[2] = {s->f[BMQ_F_RHO_PREV].vbase(), s->f[BMQ_F_T_PREV].vbase()};
        BMQ_CK(launch_double_advect(s->stream, s->g, own(s, 0), Stag{0, 0, 0}, false, 2, fl, pv, chi, chip, s->blend));
    }
    return BMQ_OK;
}

// estimateDistortion for both mappers (Mapping.cpp:91-118) with the max on the device
// dispz[0], dispz[1]: max |map_z - z| in cells of the velocity / scalar mapper (z-slab halo sizing)
// vel_d2 == nullptr: the four results stay in s->d_red[1..4] (squared distortions, z displacements in world units)
int stage_distortion(bmq3d_solver *s, float *vel_d2, float *sca_d2, float *dispz)
{
    StageTimer _t(s, BMQ_T_DISTORTION);
    BMQ_CK(cudaMemsetAsync(s->d_red, 0, 5 * sizeof(float), s->stream));
    const float *const b[2][3] = {
        {s->f[BMQ_F_VBWD_X].vbase(), s->f[BMQ_F_VBWD_Y].vbase(), s->f[BMQ_F_VBWD_Z].vbase()},
        {s->f[BMQ_F_SBWD_X].vbase(), s->f[BMQ_F_SBWD_Y].vbase(), s->f[BMQ_F_SBWD_Z].vbase()}};
    const float *const f[2][3] = {
        {s->f[BMQ_F_VFWD_X].vbase(), s->f[BMQ_F_VFWD_Y].vbase(), s->f[BMQ_F_VFWD_Z].vbase()},
        {s->f[BMQ_F_SFWD_X].vbase(), s->f[BMQ_F_SFWD_Y].vbase(), s->f[BMQ_F_SFWD_Z].vbase()}};
    float *d2[2] = {s->d_red + 1, s->d_red + 2};
    float *dz[2] = {s->d_red + 3, s->d_red + 4};
    BMQ_CK(launch_estimate(s->stream, s->g, own(s, 0), 2, b, f, nullptr, d2, dz, nullptr));
    if (!vel_d2) return BMQ_OK;
    BMQ_CK(cudaMemcpyAsync(s->h_red, s->d_red, 5 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    BMQ_CK(cudaStreamSynchronize(s->stream));
    *vel_d2 = s->h_red[1];
    *sca_d2 = s->h_red[2];
    dispz[0] = s->h_red[3] / s->h;
    dispz[1] = s->h_red[4] / s->h;
    s->stats.max_disp_z = dispz[0] > dispz[1] ? dispz[0] : dispz[1];
    s->stats.max_disp_z_vel = dispz[0];
    s->stats.max_disp_z_scalar = dispz[1];
    return BMQ_OK;
}

// reinit decision, BimocqSolver.cpp:165-185
void decide(bmq3d_solver *s, int framenum, float dt, float vel_d2, float sca_d2)
{
    s->proj_coeff = 2.f;
    s->vel_reinit = s->scalar_reinit = false;
    const float vd = sqrtf(vel_d2) / (s->max_v * dt);
    const float sd = sqrtf(sca_d2) / (s->max_v * dt);
    s->stats.vel_distortion = vd;
    s->stats.scalar_distortion = sd;
    if (vd > 1.f || framenum - s->vel_last_reinit > 10) {
        s->vel_reinit = true;
        s->vel_last_reinit = framenum;
        s->proj_coeff = 1.f;
    }
    if (sd > 5.f || framenum - s->scalar_last_reinit > 30) {
        s->scalar_reinit = true;
        s->scalar_last_reinit = framenum;
    }
    s->stats.vel_reinit = s->vel_reinit;
    s->stats.scalar_reinit = s->scalar_reinit;
}

// accumulate: init += 1*quad9[d_ext o psi] then += proj_coeff*quad9[d_proj o psi]  (BimocqSolver.cpp:193-196)
int stage_accumulate(bmq3d_solver *s, int which, int comp);
int stage_accumulate(bmq3d_solver *s, int which) { return stage_accumulate(s, which, -1); }
int stage_accumulate(bmq3d_solver *s, int which, int comp)
{
    StageTimer _t(s, which == 0 ? BMQ_T_ACCUM_V : BMQ_T_ACCUM_S);
    const float *psi[3];
    if (which == 0) {
        map_ptrs(s, BMQ_F_VFWD_X, psi);
        const float coeff[2] = {1.f, s->proj_coeff};
        for (int c = 0; c < 3; ++c) {
            if (comp >= 0 && c != comp) continue;
            float *t = s->f[BMQ_F_U_INIT + c].vbase();
            const float *ch[2] = {s->f[BMQ_F_DU_EXT + c].vbase(), s->f[BMQ_F_DU_PROJ + c].vbase()};
            Stag sg = stag_of(BMQ_F_U + c);
            BMQ_CK(launch_cumulate(s->stream, s->g, own(s, sg.dz), sg, false, 1, 2, &t, ch, coeff, psi));
        }
    } else {
        map_ptrs(s, BMQ_F_SFWD_X, psi);
        const float one = 1.f;
        float *t[2] = {s->f[BMQ_F_RHO_INIT].vbase(), s->f[BMQ_F_T_INIT].vbase()};
        const float *ch[2] = {s->f[BMQ_F_DRHO_EXT].vbase(), s->f[BMQ_F_DT_EXT].vbase()};
        BMQ_CK(launch_cumulate(s->stream, s->g, own(s, 0), Stag{0, 0, 0}, false, 2, 1, t, ch, &one, psi));
    }
    return BMQ_OK;
}

// phase 0: reinitializeMapping + velocityReinitialize/scalarReinitialize by rotation
// phase 1 (velocity only): the extra accumulate(duproj, 1.0) of BimocqSolver.cpp:214
int stage_reinit(bmq3d_solver *s, int which, int phase)
{
    StageTimer _t(s, BMQ_T_REINIT);
    if (which == 0) {
        if (phase == 0) {
            s->vel_reinit_count++;
            for (int c = 0; c < 3; ++c) {
                std::swap(s->f[BMQ_F_VBWDP_X + c], s->f[BMQ_F_VBWD_X + c]);   // chi_prev <- chi
                std::swap(s->f[BMQ_F_U_PREV + c], s->f[BMQ_F_U_INIT + c]);    // prev <- init
            }
            int st = identity_fill(s, BMQ_F_VBWD_X);
            if (st == BMQ_OK) st = identity_fill(s, BMQ_F_VFWD_X);
            for (int c = 0; c < 3 && st == BMQ_OK; ++c) st = copy_field(s, s->f[BMQ_F_U_INIT + c], s->f[BMQ_F_U + c]);
            return st;
        }
        const float *psi[3];
        map_ptrs(s, BMQ_F_VFWD_X, psi);
        const float one = 1.f;
        for (int c = 0; c < 3; ++c) {
            float *t = s->f[BMQ_F_U_INIT + c].vbase();
            const float *ch = s->f[BMQ_F_DU_PROJ + c].vbase();
            Stag sg = stag_of(BMQ_F_U + c);
            BMQ_CK(launch_cumulate(s->stream, s->g, own(s, sg.dz), sg, false, 1, 1, &t, &ch, &one, psi));
        }
        return BMQ_OK;
    }
    if (phase != 0) return BMQ_OK;
    s->scalar_reinit_count++;
    for (int c = 0; c < 3; ++c) std::swap(s->f[BMQ_F_SBWDP_X + c], s->f[BMQ_F_SBWD_X + c]);
    std::swap(s->f[BMQ_F_RHO_PREV], s->f[BMQ_F_RHO_INIT]);
    std::swap(s->f[BMQ_F_T_PREV], s->f[BMQ_F_T_INIT]);
    int st = identity_fill(s, BMQ_F_SBWD_X);
    if (st == BMQ_OK) st = identity_fill(s, BMQ_F_SFWD_X);
    if (st == BMQ_OK) st = copy_field(s, s->f[BMQ_F_RHO_INIT], s->f[BMQ_F_RHO]);
    if (st == BMQ_OK) st = copy_field(s, s->f[BMQ_F_T_INIT], s->f[BMQ_F_T]);
    return st;
}

#define RET_IF(x) do { int _s = (x); if (_s != BMQ_OK) return _s; } while (0)
#define NEED(s) do { if (!(s)) return set_error(BMQ_ERR_ARG, "%s: null solver handle", __func__); } while (0)

}  // namespace

extern "C" {

int bmq3d_create_slab(int ni, int nj, int nk, float h, float blend_coeff, int k_own0, int k_own1, int halo,
                      bmq3d_solver **out)
{
    if (!out) return set_error(BMQ_ERR_ARG, "bmq3d_create: out is null");
    *out = nullptr;
    if (!require_device()) return BMQ_ERR_NODEVICE;
    if (ni < 8 || nj < 8 || nk < 8 || !(h > 0.f) || k_own0 < 0 || k_own1 > nk || k_own0 >= k_own1 || halo < 0)
        return set_error(BMQ_ERR_ARG, "bmq3d_create: bad grid %dx%dx%d h=%g slab [%d,%d) halo %d", ni, nj, nk,
                         (double)h, k_own0, k_own1, halo);
    if ((long long)(ni + 1) * (nj + 1) * (long long)(nk + 2) >= (1ll << 31))
        return set_error(BMQ_ERR_ARG, "bmq3d_create: grid too large for 32-bit indexing");
    bmq3d_solver *s = new (std::nothrow) bmq3d_solver();
    if (!s) return set_error(BMQ_ERR_ARG, "bmq3d_create: out of host memory");
    s->ni = ni; s->nj = nj; s->nk = nk; s->h = h; s->blend = blend_coeff;
    s->k0 = k_own0; s->k1 = k_own1; s->halo = halo;
    s->g = make_grid(ni, nj, nk, h);
    memset(&s->stats, 0, sizeof s->stats);
    int st = BMQ_OK;
    for (int id = 0; id < BMQ_F_U_SEMI && st == BMQ_OK; ++id) st = alloc_field(s, s->f[id], stag_of(id));
    for (int q = 0; q < N_SCRATCH && st == BMQ_OK; ++q) st = alloc_field(s, s->scratch[q], stag_of(SCRATCH_BASE + q));
    for (int q = 0; q < N_TMPMAP && st == BMQ_OK; ++q) st = alloc_field(s, s->tmpmap[q], Stag{0, 0, 0});
    if (st == BMQ_OK) st = check_cuda(cudaMalloc(&s->d_red, 8 * sizeof(float)), "cudaMalloc", __FILE__, __LINE__);
    if (st == BMQ_OK) st = check_cuda(cudaMallocHost(&s->h_red, 8 * sizeof(float)), "cudaMallocHost", __FILE__, __LINE__);
    if (st == BMQ_OK) st = bmq3d_reset(s);
    if (st != BMQ_OK) { bmq3d_destroy(s); return st; }
    *out = s;
    return BMQ_OK;
}

int bmq3d_create(int ni, int nj, int nk, float h, float blend_coeff, bmq3d_solver **out)
{
    return bmq3d_create_slab(ni, nj, nk, h, blend_coeff, 0, nk, 0, out);
}

int bmq3d_destroy(bmq3d_solver *s)
{
    if (!s) return BMQ_OK;
    for (auto &fd : s->f) if (fd.alloc) cudaFree(fd.alloc);
    for (auto &fd : s->scratch) if (fd.alloc) cudaFree(fd.alloc);
    for (auto &fd : s->tmpmap) if (fd.alloc) cudaFree(fd.alloc);
    if (s->s_h2d) cudaStreamDestroy(s->s_h2d);
    if (s->s_d2h) cudaStreamDestroy(s->s_d2h);
    for (auto e : s->ev_copy) if (e) cudaEventDestroy(e);
    for (auto &sp : s->spans) { cudaEventDestroy(sp.a); cudaEventDestroy(sp.b); }
    for (auto e : s->event_pool) cudaEventDestroy(e);
    if (s->stage_up) cudaFree(s->stage_up);
    if (s->stage_down) cudaFree(s->stage_down);
    if (s->d_red) cudaFree(s->d_red);
    if (s->h_red) cudaFreeHost(s->h_red);
    delete s;
    return BMQ_OK;
}

// Re-allocates every field with a wider halo, keeping the stored planes where they are in the
// global grid (new planes are zero; the next halo exchange fills them).
int bmq3d_grow_halo(bmq3d_solver *s, int new_halo)
{
    NEED(s);
    if (new_halo <= s->halo) return BMQ_OK;
    BMQ_CK(cudaStreamSynchronize(s->stream));
    const int old_halo = s->halo;
    s->halo = new_halo;
    auto regrow = [&](Field &fd, Stag st) -> int {
        if (!fd.alloc) return BMQ_OK;
        Field old = fd;
        int rc = alloc_field(s, fd, st);
        if (rc != BMQ_OK) { fd = old; return rc; }
        BMQ_CK(cudaMemcpyAsync(fd.alloc + fd.plane() * (size_t)(old.p0 - fd.p0), old.alloc, old.stored_elems() * sizeof(float),
                               cudaMemcpyDeviceToDevice, s->stream));
        BMQ_CK(cudaStreamSynchronize(s->stream));
        BMQ_CK(cudaFree(old.alloc));
        return BMQ_OK;
    };
    int rc = BMQ_OK;
    for (int id = 0; id < BMQ_F_COUNT && rc == BMQ_OK; ++id) rc = regrow(s->f[id], stag_of(id));
    for (int q = 0; q < N_SCRATCH && rc == BMQ_OK; ++q) rc = regrow(s->scratch[q], stag_of(SCRATCH_BASE + q));
    for (int q = 0; q < N_TMPMAP && rc == BMQ_OK; ++q) rc = regrow(s->tmpmap[q], Stag{0, 0, 0});
    if (rc != BMQ_OK) { s->halo = old_halo; return rc; }
    // the identity maps must hold the identity in the new halo planes as well (a reinitialised map is
    // the identity everywhere it is stored); everything else is refreshed by its exchange
    return BMQ_OK;
}

int bmq3d_set_stream(bmq3d_solver *s, void *stream)
{
    NEED(s);
    s->stream = (cudaStream_t)stream;
    return BMQ_OK;
}

int bmq3d_field_ptr(bmq3d_solver *s, int field_id, float **dev_ptr, int *first_plane, int *n_planes, int *nx,
                    int *ny)
{
    NEED(s);
    if (field_id >= BMQ_F_U_SEMI && field_id <= BMQ_F_T_SEMI) RET_IF(ensure_semi(s));
    Field *fd = field_of(s, field_id);
    if (!fd || !fd->alloc) return set_error(BMQ_ERR_ARG, "bmq3d_field_ptr: bad field id %d", field_id);
    if (dev_ptr) *dev_ptr = fd->alloc;
    if (first_plane) *first_plane = fd->p0;
    if (n_planes) *n_planes = fd->p1 - fd->p0;
    if (nx) *nx = fd->nx;
    if (ny) *ny = fd->ny;
    return BMQ_OK;
}

// Host <-> device field transfer in the handle's host layout, enqueued on `st`.  Blocked layout
// (gpuMapper::copyHostToDevice / copyDeviceToHost, GPU_Advection.h:249-299, minus their host-side
// relayout loops): the raw blocked buffer crosses PCIe and a device kernel permutes it.
static int h2d_field(bmq3d_solver *s, const Field &fd, const float *host, cudaStream_t st)
{
    if (s->host_layout == BMQ_LAYOUT_LINEAR) {
        BMQ_CK(cudaMemcpyAsync(fd.alloc, host, fd.stored_elems() * sizeof(float), cudaMemcpyHostToDevice, st));
        return BMQ_OK;
    }
    BMQ_CK(cudaMemcpyAsync(s->stage_up, host, blocked_elems(fd.nx, fd.ny, fd.nz) * sizeof(float), cudaMemcpyHostToDevice, st));
    BMQ_CK(launch_relayout(st, true, s->stage_up, fd.alloc, fd.nx, fd.ny, fd.nz));
    return BMQ_OK;
}
static int d2h_field(bmq3d_solver *s, const Field &fd, float *host, cudaStream_t st)
{
    if (s->host_layout == BMQ_LAYOUT_LINEAR) {
        BMQ_CK(cudaMemcpyAsync(host, fd.alloc, fd.stored_elems() * sizeof(float), cudaMemcpyDeviceToHost, st));
        return BMQ_OK;
    }
    BMQ_CK(launch_relayout(st, false, fd.alloc, s->stage_down, fd.nx, fd.ny, fd.nz));
    BMQ_CK(cudaMemcpyAsync(host, s->stage_down, blocked_elems(fd.nx, fd.ny, fd.nz) * sizeof(float), cudaMemcpyDeviceToHost, st));
    return BMQ_OK;
}

int bmq3d_set_host_layout(bmq3d_solver *s, int layout)
{
    NEED(s);
    if (layout != BMQ_LAYOUT_LINEAR && layout != BMQ_LAYOUT_BLOCKED8) return set_error(BMQ_ERR_ARG, "bmq3d_set_host_layout: unknown layout %d", layout);
    if (layout == BMQ_LAYOUT_BLOCKED8) {
        if (s->k0 != 0 || s->k1 != s->nk) return set_error(BMQ_ERR_ARG, "bmq3d_set_host_layout: blocked host buffers need a full-domain handle");
        if (!s->stage_up) {
            const size_t n = blocked_elems(s->ni + 1, s->nj + 1, s->nk + 1) * sizeof(float);   // covers every field shape
            BMQ_CK(cudaMalloc(&s->stage_up, n));
            BMQ_CK(cudaMalloc(&s->stage_down, n));
        }
    }
    s->host_layout = layout;
    return BMQ_OK;
}

int bmq3d_upload(bmq3d_solver *s, int field_id, const float *host)
{
    NEED(s);
    if (field_id >= BMQ_F_U_SEMI && field_id <= BMQ_F_T_SEMI) RET_IF(ensure_semi(s));
    Field *fd = field_of(s, field_id);
    if (!fd || !fd->alloc || !host) return set_error(BMQ_ERR_ARG, "bmq3d_upload: bad field id %d or null host", field_id);
    RET_IF(h2d_field(s, *fd, host, s->stream));
    BMQ_CK(cudaStreamSynchronize(s->stream));
    return BMQ_OK;
}

int bmq3d_download(bmq3d_solver *s, int field_id, float *host)
{
    NEED(s);
    if (field_id >= BMQ_F_U_SEMI && field_id <= BMQ_F_T_SEMI) RET_IF(ensure_semi(s));
    Field *fd = field_of(s, field_id);
    if (!fd || !fd->alloc || !host) return set_error(BMQ_ERR_ARG, "bmq3d_download: bad field id %d or null host", field_id);
    RET_IF(d2h_field(s, *fd, host, s->stream));
    BMQ_CK(cudaStreamSynchronize(s->stream));
    return BMQ_OK;
}

int bmq3d_reset(bmq3d_solver *s)
{
    NEED(s);
    const int maps[6] = {BMQ_F_VFWD_X, BMQ_F_VBWD_X, BMQ_F_VBWDP_X, BMQ_F_SFWD_X, BMQ_F_SBWD_X, BMQ_F_SBWDP_X};
    for (int m : maps) RET_IF(identity_fill(s, m));
    RET_IF(identity_fill_tmp(s, 0));
    RET_IF(identity_fill_tmp(s, 3));
    for (int c = 0; c < 5; ++c) {
        RET_IF(copy_field(s, s->f[BMQ_F_U_INIT + c], s->f[BMQ_F_U + c]));
        RET_IF(copy_field(s, s->f[BMQ_F_U_PREV + c], s->f[BMQ_F_U + c]));
    }
    s->vel_last_reinit = s->scalar_last_reinit = 0;
    s->vel_reinit_count = s->scalar_reinit_count = 0;
    s->proj_coeff = 2.f;
    BMQ_CK(cudaStreamSynchronize(s->stream));
    return BMQ_OK;
}

int bmq3d_stage_maxvel(bmq3d_solver *s, float *max_abs_out)
{
    NEED(s);
    float m = 0.f;
    RET_IF(stage_maxvel(s, &m));
    if (max_abs_out) *max_abs_out = m;
    return BMQ_OK;
}
int bmq3d_stage_set_cfl(bmq3d_solver *s, int framenum, float global_max_abs) { NEED(s); set_cfl(s, framenum, global_max_abs); return BMQ_OK; }
int bmq3d_stage_dmc_substep(bmq3d_solver *s, float substep) { NEED(s); return stage_dmc(s, substep); }
int bmq3d_stage_forward(bmq3d_solver *s, float dt) { NEED(s); return stage_forward(s, dt); }
int bmq3d_stage_semilag(bmq3d_solver *s, float dt) { NEED(s); return stage_semilag(s, dt); }
static int for_which(bmq3d_solver *s, int which, int (*fn)(bmq3d_solver *, int))
{
    if (which < 0 || which > 2) return set_error(BMQ_ERR_ARG, "which must be 0 (velocity), 1 (scalars) or 2 (both)");
    if (which != 1) RET_IF(fn(s, 0));
    if (which != 0) RET_IF(fn(s, 1));
    return BMQ_OK;
}
int bmq3d_stage_advect(bmq3d_solver *s, int which) { NEED(s); return for_which(s, which, stage_advect); }
int bmq3d_stage_error(bmq3d_solver *s, int which) { NEED(s); return for_which(s, which, stage_error); }
int bmq3d_stage_apply(bmq3d_solver *s, int which) { NEED(s); return for_which(s, which, static_cast<int (*)(bmq3d_solver *, int)>(stage_apply)); }
int bmq3d_stage_blend(bmq3d_solver *s, int which) { NEED(s); return for_which(s, which, stage_blend); }
int bmq3d_stage_accumulate(bmq3d_solver *s, int which) { NEED(s); return for_which(s, which, static_cast<int (*)(bmq3d_solver *, int)>(stage_accumulate)); }
int bmq3d_stage_distortion(bmq3d_solver *s, float *vel_d2, float *scalar_d2, float *max_disp_z)
{
    NEED(s);
    float a = 0, b = 0, c[2] = {0, 0};
    RET_IF(stage_distortion(s, &a, &b, c));
    if (vel_d2) *vel_d2 = a;
    if (scalar_d2) *scalar_d2 = b;
    if (max_disp_z) *max_disp_z = s->stats.max_disp_z;
    return BMQ_OK;
}
int bmq3d_stage_distortion2(bmq3d_solver *s, float *vel_d2, float *scalar_d2, float *disp_z_vel, float *disp_z_scalar)
{
    NEED(s);
    float a = 0, b = 0, c[2] = {0, 0};
    RET_IF(stage_distortion(s, &a, &b, c));
    if (vel_d2) *vel_d2 = a;
    if (scalar_d2) *scalar_d2 = b;
    if (disp_z_vel) *disp_z_vel = c[0];
    if (disp_z_scalar) *disp_z_scalar = c[1];
    return BMQ_OK;
}
int bmq3d_stage_decide(bmq3d_solver *s, int framenum, float dt, float vel_d2, float scalar_d2)
{
    NEED(s);
    decide(s, framenum, dt, vel_d2, scalar_d2);
    return BMQ_OK;
}
int bmq3d_stage_reinit(bmq3d_solver *s, int which, int phase)
{
    NEED(s);
    if (which != 0 && which != 1) return set_error(BMQ_ERR_ARG, "bmq3d_stage_reinit: which must be 0 or 1");
    int st = stage_reinit(s, which, phase);
    s->stats.vel_reinit_count = s->vel_reinit_count;
    s->stats.scalar_reinit_count = s->scalar_reinit_count;
    return st;
}

// Phase A, BimocqSolver.cpp:90-126
int bmq3d_advect(bmq3d_solver *s, int framenum, float dt, int with_semilag)
{
    NEED(s);
    if (!(dt > 0.f)) return set_error(BMQ_ERR_ARG, "bmq3d_advect: dt must be positive");
    float mabs = 0.f;
    RET_IF(stage_maxvel(s, &mabs));
    set_cfl(s, framenum, mabs);
    // updateBackward, Mapping.cpp:7-24 / 354-368: the float loop that fixes the sub-step sequence
    float T = 0.f, substep = s->cfldt;
    int n = 0;
    while (T < dt) {
        if (T + substep > dt) substep = dt - T;
        RET_IF(stage_dmc(s, substep));
        T += substep;
        if (++n > 4096) return set_error(BMQ_ERR_ARG, "bmq3d_advect: more than 4096 CFL sub-steps (dt/cfldt too large)");
    }
    s->stats.n_substeps = n;
    RET_IF(stage_forward(s, dt));
    if (with_semilag) RET_IF(stage_semilag(s, dt));
    for (int which = 0; which < 2; ++which) {
        RET_IF(stage_advect(s, which));
        RET_IF(stage_error(s, which));
        RET_IF(stage_apply(s, which));
        RET_IF(stage_blend(s, which));
    }
    return BMQ_OK;
}

// Phase B, BimocqSolver.cpp:164-229
int bmq3d_accumulate(bmq3d_solver *s, int framenum, float dt)
{
    NEED(s);
    float vd2 = 0, sd2 = 0, dz[2] = {0, 0};
    RET_IF(stage_distortion(s, &vd2, &sd2, dz));
    decide(s, framenum, dt, vd2, sd2);
    RET_IF(stage_accumulate(s, 0));
    RET_IF(stage_accumulate(s, 1));
    if (s->vel_reinit) {
        RET_IF(stage_reinit(s, 0, 0));
        RET_IF(stage_reinit(s, 0, 1));
    }
    if (s->scalar_reinit) RET_IF(stage_reinit(s, 1, 0));
    s->stats.vel_reinit_count = s->vel_reinit_count;
    s->stats.scalar_reinit_count = s->scalar_reinit_count;
    return BMQ_OK;
}

int bmq3d_timing_enable(bmq3d_solver *s, int on)
{
    NEED(s);
    s->timing = on != 0;
    return BMQ_OK;
}

const char *bmq3d_timing_slot_name(int slot) { return slot >= 0 && slot < BMQ_T_COUNT ? kSlotNames[slot] : ""; }

// Sums the recorded spans per slot (synchronises the stream), returns them and clears the record.
int bmq3d_timing_read(bmq3d_solver *s, float *ms_out, int *spans_out, int n_slots)
{
    NEED(s);
    if (!ms_out || n_slots < BMQ_T_COUNT) return set_error(BMQ_ERR_ARG, "bmq3d_timing_read: need %d slots", (int)BMQ_T_COUNT);
    BMQ_CK(cudaStreamSynchronize(s->stream));
    for (int q = 0; q < n_slots; ++q) { ms_out[q] = 0.f; if (spans_out) spans_out[q] = 0; }
    for (auto &sp : s->spans) {
        float ms = 0.f;
        BMQ_CK(cudaEventElapsedTime(&ms, sp.a, sp.b));
        ms_out[sp.slot] += ms;
        if (spans_out) spans_out[sp.slot]++;
        s->event_pool.push_back(sp.a);
        s->event_pool.push_back(sp.b);
    }
    s->spans.clear();
    return BMQ_OK;
}

// As bmq3d_timing_read, plus per slot the time the stream spent between the end of the previous recorded stage and
// the start of this one: what a stage waited for (a halo exchange, a reduction with its host round trip, the
// caller's own kernels) -- the fixed cost of the z-slab step, named by the stage that had to wait.
int bmq3d_timing_read_gaps(bmq3d_solver *s, float *ms_out, int *spans_out, float *gap_ms_out, int n_slots)
{
    NEED(s);
    if (!ms_out || !gap_ms_out || n_slots < BMQ_T_COUNT) return set_error(BMQ_ERR_ARG, "bmq3d_timing_read_gaps: need %d slots", (int)BMQ_T_COUNT);
    BMQ_CK(cudaStreamSynchronize(s->stream));
    for (int q = 0; q < n_slots; ++q) gap_ms_out[q] = 0.f;
    for (size_t i = 1; i < s->spans.size(); ++i) {
        float ms = 0.f;
        BMQ_CK(cudaEventElapsedTime(&ms, s->spans[i - 1].b, s->spans[i].a));
        gap_ms_out[s->spans[i].slot] += ms;
    }
    return bmq3d_timing_read(s, ms_out, spans_out, n_slots);
}

int bmq3d_get_stats(bmq3d_solver *s, bmq3d_stats *out)
{
    NEED(s);
    if (!out) return set_error(BMQ_ERR_ARG, "bmq3d_get_stats: out is null");
    *out = s->stats;
    return BMQ_OK;
}

static int ensure_copy_streams(bmq3d_solver *s)
{
    if (s->s_h2d) return BMQ_OK;
    BMQ_CK(cudaStreamCreateWithFlags(&s->s_h2d, cudaStreamNonBlocking));
    BMQ_CK(cudaStreamCreateWithFlags(&s->s_d2h, cudaStreamNonBlocking));
    for (auto &e : s->ev_copy) BMQ_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return BMQ_OK;
}

static int upload_async(bmq3d_solver *s, Field &fd, const float *host, cudaEvent_t done)
{
    RET_IF(h2d_field(s, fd, host, s->s_h2d));
    BMQ_CK(cudaEventRecord(done, s->s_h2d));
    return BMQ_OK;
}

// Phase A through host buffers.  Transfers run on their own streams and overlap the stages:
// only u,v,w are uploaded (phase A never reads the current density / temperature: they are pure
// outputs of MapperBase::advectField, Mapping.cpp:208-236); u,v,w travel back while the scalar
// stages run.
// Error exits of the *_host calls must not leave asynchronous copies touching the caller's buffers in flight.
struct HostCallGuard {
    bmq3d_solver *s;
    ~HostCallGuard()
    {
        if (s->s_h2d) cudaStreamSynchronize(s->s_h2d);
        if (s->s_d2h) cudaStreamSynchronize(s->s_d2h);
        cudaStreamSynchronize(s->stream);
    }
};

static int advect_host_impl(bmq3d_solver *s, int framenum, float dt, float *u, float *v, float *w, float *rho, float *T);
int bmq3d_advect_host(bmq3d_solver *s, int framenum, float dt, float *u, float *v, float *w, float *rho, float *T)
{
    NEED(s);
    HostCallGuard guard{s};      // runs after the body, on every path
    return advect_host_impl(s, framenum, dt, u, v, w, rho, T);
}
static int advect_host_impl(bmq3d_solver *s, int framenum, float dt, float *u, float *v, float *w, float *rho, float *T)
{
    if (!(dt > 0.f)) return set_error(BMQ_ERR_ARG, "bmq3d_advect_host: dt must be positive");
    float *host[5] = {u, v, w, rho, T};
    for (int c = 0; c < 5; ++c)
        if (!host[c]) return set_error(BMQ_ERR_ARG, "bmq3d_advect_host: null field pointer %d", c);
    RET_IF(ensure_copy_streams(s));
    cudaEvent_t *ev = s->ev_copy;
    BMQ_CK(cudaEventRecord(ev[0], s->stream));                  // everything queued so far
    BMQ_CK(cudaStreamWaitEvent(s->s_h2d, ev[0], 0));
    for (int c = 0; c < 3; ++c) RET_IF(upload_async(s, s->f[BMQ_F_U + c], host[c], ev[1]));
    BMQ_CK(cudaStreamWaitEvent(s->stream, ev[1], 0));
    float mabs = 0.f;
    RET_IF(stage_maxvel(s, &mabs));
    set_cfl(s, framenum, mabs);
    float Tt = 0.f, substep = s->cfldt;
    int n = 0;
    while (Tt < dt) {
        if (Tt + substep > dt) substep = dt - Tt;
        RET_IF(stage_dmc(s, substep));
        Tt += substep;
        if (++n > 4096) return set_error(BMQ_ERR_ARG, "bmq3d_advect_host: more than 4096 CFL sub-steps");
    }
    s->stats.n_substeps = n;
    RET_IF(stage_forward(s, dt));
    // The two mappers' stages are independent: scalars first (their download overlaps the longer
    // velocity stages), then the velocity, each component going home as soon as it is final.
    RET_IF(stage_advect(s, 1));
    RET_IF(stage_error(s, 1));
    RET_IF(stage_apply(s, 1));
    RET_IF(stage_blend(s, 1));
    BMQ_CK(cudaEventRecord(ev[2], s->stream));
    BMQ_CK(cudaStreamWaitEvent(s->s_d2h, ev[2], 0));
    for (int c = 3; c < 5; ++c) RET_IF(d2h_field(s, s->f[BMQ_F_U + c], host[c], s->s_d2h));
    RET_IF(stage_advect(s, 0));
    RET_IF(stage_error(s, 0));
    const bool blend_active = s->vel_reinit_count > 0 && s->blend != 1.0f;   // stage_blend works on all three
    for (int c = 0; c < 3; ++c) {
        RET_IF(stage_apply(s, 0, c));
        if (blend_active) {
            if (c < 2) continue;
            RET_IF(stage_blend(s, 0));
        }
        BMQ_CK(cudaEventRecord(ev[12 + c], s->stream));
        BMQ_CK(cudaStreamWaitEvent(s->s_d2h, ev[12 + c], 0));
        for (int q = blend_active ? 0 : c; q <= c; ++q) RET_IF(d2h_field(s, s->f[BMQ_F_U + q], host[q], s->s_d2h));
    }
    BMQ_CK(cudaStreamSynchronize(s->s_d2h));
    BMQ_CK(cudaStreamSynchronize(s->stream));
    return BMQ_OK;
}

static int accumulate_host_impl(bmq3d_solver *s, int framenum, float dt, const float *u_forced, const float *v_forced,
                                const float *w_forced, const float *u_final, const float *v_final, const float *w_final,
                                const float *rho_final, const float *T_final);
int bmq3d_accumulate_host(bmq3d_solver *s, int framenum, float dt, const float *u_forced, const float *v_forced,
                          const float *w_forced, const float *u_final, const float *v_final, const float *w_final,
                          const float *rho_final, const float *T_final)
{
    NEED(s);
    HostCallGuard guard{s};
    return accumulate_host_impl(s, framenum, dt, u_forced, v_forced, w_forced, u_final, v_final, w_final, rho_final, T_final);
}
static int accumulate_host_impl(bmq3d_solver *s, int framenum, float dt, const float *u_forced, const float *v_forced,
                                const float *w_forced, const float *u_final, const float *v_final, const float *w_final,
                                const float *rho_final, const float *T_final)
{
    const float *forced[3] = {u_forced, v_forced, w_forced};
    const float *fin[5] = {u_final, v_final, w_final, rho_final, T_final};
    for (int c = 0; c < 5; ++c)
        if (!fin[c] || (c < 3 && !forced[c])) return set_error(BMQ_ERR_ARG, "bmq3d_accumulate_host: null field %d", c);
    RET_IF(ensure_copy_streams(s));
    cudaEvent_t *ev = s->ev_copy;
    // The change fields are formed on the device exactly as the reference forms them on the host
    // (BimocqSolver.cpp:149-162): d_ext = forced - advected, d_proj = final - forced,
    // d_scalar = final - advected; the device copy of the current fields becomes `final`.
    // Uploads go to staging buffers on the copy stream (forced -> D*_PROJ, final -> the advect
    // scratch, which is free in phase B) and overlap the distortion kernel and each other's
    // consumers; the scratch is cleared afterwards because it must keep its zero ring.
    BMQ_CK(cudaEventRecord(ev[0], s->stream));
    BMQ_CK(cudaStreamWaitEvent(s->s_h2d, ev[0], 0));
    for (int c = 0; c < 3; ++c) {
        RET_IF(upload_async(s, s->f[BMQ_F_DU_PROJ + c], forced[c], ev[4 + 2 * c]));
        RET_IF(upload_async(s, s->scratch[c], fin[c], ev[5 + 2 * c]));
    }
    for (int c = 3; c < 5; ++c) RET_IF(upload_async(s, s->scratch[c], fin[c], ev[7 + c]));
    float vd2 = 0, sd2 = 0, dz[2] = {0, 0};
    RET_IF(stage_distortion(s, &vd2, &sd2, dz));
    decide(s, framenum, dt, vd2, sd2);
    for (int c = 0; c < 3; ++c) {
        Field &cur = s->f[BMQ_F_U + c], &ext = s->f[BMQ_F_DU_EXT + c], &proj = s->f[BMQ_F_DU_PROJ + c], &stg = s->scratch[c];
        const size_t nel = cur.stored_elems();
        BMQ_CK(cudaStreamWaitEvent(s->stream, ev[4 + 2 * c], 0));
        BMQ_CK(launch_add_field(s->stream, ext.alloc, proj.alloc, cur.alloc, -1.f, nel));     // forced - advected
        BMQ_CK(cudaStreamWaitEvent(s->stream, ev[5 + 2 * c], 0));
        BMQ_CK(launch_add_field(s->stream, proj.alloc, stg.alloc, proj.alloc, -1.f, nel));    // final - forced
        BMQ_CK(cudaMemcpyAsync(cur.alloc, stg.alloc, nel * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
        BMQ_CK(cudaMemsetAsync(stg.alloc, 0, nel * sizeof(float), s->stream));
        RET_IF(stage_accumulate(s, 0, c));   // overlaps the uploads of the next components
    }
    for (int c = 3; c < 5; ++c) {
        Field &cur = s->f[BMQ_F_U + c], &ext = s->f[BMQ_F_DU_EXT + c], &stg = s->scratch[c];
        const size_t nel = cur.stored_elems();
        BMQ_CK(cudaStreamWaitEvent(s->stream, ev[7 + c], 0));
        BMQ_CK(launch_add_field(s->stream, ext.alloc, stg.alloc, cur.alloc, -1.f, nel));      // final - advected
        BMQ_CK(cudaMemcpyAsync(cur.alloc, stg.alloc, nel * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
        BMQ_CK(cudaMemsetAsync(stg.alloc, 0, nel * sizeof(float), s->stream));
    }
    RET_IF(stage_accumulate(s, 1));
    if (s->vel_reinit) {
        RET_IF(stage_reinit(s, 0, 0));
        RET_IF(stage_reinit(s, 0, 1));
    }
    if (s->scalar_reinit) RET_IF(stage_reinit(s, 1, 0));
    s->stats.vel_reinit_count = s->vel_reinit_count;
    s->stats.scalar_reinit_count = s->scalar_reinit_count;
    BMQ_CK(cudaStreamSynchronize(s->stream));
    return BMQ_OK;
}


}  // extern "C"

// ==================================================================================================
// bmq3d_mg_*: the z-slab decomposition as a C API.  One bmq3d_mg per rank / GPU /
// process; it owns the rank's slab handle, maps every other rank's field allocations (CUDA IPC, or raw
// pointers when several ranks live in one process) and runs the advection phases with the halo
// exchanges of gpufluidsimulation_b200/zslab.py:ZSlabStepper -- same schedule, same widths -- as
// stream-ordered peer copies over NVLink.  What it cannot do itself, because the host program owns the
// communicator, comes in as two callbacks: a max all-reduce of a few floats and a barrier ordered in a
// CUDA stream.
// ==================================================================================================
namespace {

enum { MG_NARROW = 5, MG_EVENTS = 32 };
// posted exchanges whose completion is waited for later in the step, each with its own event
enum MgDone { MG_D_BLOCKING = 0, MG_D_VEL, MG_D_INIT, MG_D_BWD, MG_D_FWD, MG_D_ADV_V, MG_D_ADV_S, MG_D_ERR_V, MG_D_ERR_S, MG_D_CH_V, MG_D_CH_S,
              MG_DONE_EVENTS };

struct MgGroup { const int *ids; int n; int width; };

struct MgBlobHeader { long long pid; int n_allocs; int halo; int rank; int pad; unsigned char sig_ipc[64]; void *sig_raw; };
struct MgBlobEntry { unsigned char ipc[64]; void *raw; };

int mg_slab_k0(int nk, int world, int r) { const int base = nk / world, rem = nk % world; return r * base + (r < rem ? r : rem); }

}  // namespace

struct bmq3d_mg {
    bmq3d_solver *s = nullptr;
    int rank = 0, world = 1;
    cudaStream_t copy = nullptr, copy2 = nullptr;   // copy: barrier + pulls from the lower side; copy2: pulls from the upper side
    cudaStream_t red = nullptr;                     // the distortion reduction, beside the scalar accumulation
    cudaEvent_t ev[MG_EVENTS] = {};                 // transient events (waited for right after they are recorded)
    int ev_next = 0;
    cudaEvent_t done_ev[MG_DONE_EVENTS] = {};       // completion of the posted exchanges, one per purpose (see MgDone)
    std::vector<std::vector<void *>> peer;      // [rank][alloc_id]: base of that rank's allocation, mapped here
    std::vector<char> peer_is_ipc;              // [rank]: mapped with cudaIpcOpenMemHandle (must be closed)
    bmq_allreduce_max_fn allreduce = nullptr;
    bmq_stream_barrier_fn barrier = nullptr;
    void *ctx = nullptr;
    float disp[2] = {0.f, 0.f}, disp_prev[2] = {0.f, 0.f};
    int reinit_count[2] = {0, 0};
    int wv = 3, ws = 3;
    int chi_valid[2] = {0, 0};                  // halo planes of the backward maps that are up to date, per mapper
    bmq3d_mg_stats stats;
    // device-side signalling (mg_signal.h): this rank's block, the peers' blocks mapped here, counters
    int signal_request = BMQ_MG_SIGNAL_AUTO, signal_mode = BMQ_MG_SIGNAL_HOST;
    MgSignal *sig = nullptr;
    std::vector<MgSignal *> peer_sig;           // [rank]
    MgWaitList near_ranks = {}, other_ranks = {};   // ranks within halo reach (exchange barrier) / all other ranks (reductions)
    unsigned epoch = 0, red_seq = 0;
    float *red_host = nullptr;                  // pinned: MG_RED_MAX results + the timed-out flag
};

namespace {

cudaEvent_t mg_event(bmq3d_mg *m) { cudaEvent_t e = m->ev[m->ev_next]; m->ev_next = (m->ev_next + 1) % MG_EVENTS; return e; }

// global planes [a, b) of a field that rank r owns (w faces: the top face belongs to the last rank)
void mg_owned(const bmq3d_mg *m, int dz, int r, int &a, int &b)
{
    const int nk = m->s->nk;
    a = mg_slab_k0(nk, m->world, r);
    b = mg_slab_k0(nk, m->world, r + 1) + ((dz && r == m->world - 1) ? 1 : 0);
}

// Posts one exchange on the copy stream: wait for this rank's producers, barrier, pull every halo segment
// out of its owner's memory.  Returns the event the consumers wait for.
// Two implementations of barrier + pull (bmq3d_mg_set_signalling):
//   host callbacks:  the host's stream barrier (an all-rank collective), then one cudaMemcpyAsync per segment;
//   device flags:    ONE kernel that publishes this rank's arrival, waits for the ranks within halo reach and
//                    copies all segments with 128-bit loads over NVLink (mg_signal.cu); BMQ_MG_SIGNAL_DEVICE_CE
//                    keeps the flag barrier but leaves the copies to the copy engines.
int mg_post(bmq3d_mg *m, const MgGroup *groups, int ngroups, MgDone which)
{
    bmq3d_solver *s = m->s;
    cudaEvent_t produced = mg_event(m);
    BMQ_CK(cudaEventRecord(produced, s->stream));
    BMQ_CK(cudaStreamWaitEvent(m->copy, produced, 0));
    const bool flags = m->world > 1 && m->signal_mode != BMQ_MG_SIGNAL_HOST;
    const bool by_kernel = m->signal_mode == BMQ_MG_SIGNAL_DEVICE;
    if (m->world > 1 && !flags) {
        if (!m->barrier) return set_error(BMQ_ERR_ARG, "bmq3d_mg: no stream barrier callback set");
        if (m->barrier((void *)m->copy, m->ctx) != 0) return set_error(BMQ_ERR_CUDA, "bmq3d_mg: the stream barrier callback failed");
    }
    MgSegList segs;
    segs.n = 0;
    unsigned epoch = flags ? ++m->epoch : 0u;
    if (flags && !by_kernel) {
        segs.n = 0;
        BMQ_CK(launch_mg_pull(m->copy, m->sig, m->near_ranks, epoch, segs));      // barrier only
        epoch = 0;
    }
    // copy engines: the two sides of the slab are pulled by two streams, so two engines (and both directions of the
    // links) work at once; the upper side starts after the barrier and joins before the exchange is complete
    const bool two_streams = m->world > 1 && !by_kernel;
    if (two_streams) {
        cudaEvent_t open = mg_event(m);
        BMQ_CK(cudaEventRecord(open, m->copy));
        BMQ_CK(cudaStreamWaitEvent(m->copy2, open, 0));
    }
    for (int g = 0; g < ngroups; ++g) {
        for (int q = 0; q < groups[g].n; ++q) {
            Field *fd = field_of(s, groups[g].ids[q]);
            if (!fd || !fd->alloc) return set_error(BMQ_ERR_ARG, "bmq3d_mg: bad field id %d in an exchange", groups[g].ids[q]);
            const int dz = fd->nz - s->nk, w = groups[g].width;
            int a0, b0;
            mg_owned(m, dz, m->rank, a0, b0);
            const int want[2][2] = {{m->rank > 0 ? (a0 - w > 0 ? a0 - w : 0) : 0, m->rank > 0 ? a0 : 0},
                                    {m->rank < m->world - 1 ? b0 : 0, m->rank < m->world - 1 ? (b0 + w + 1 < fd->nz ? b0 + w + 1 : fd->nz) : 0}};
            for (int side = 0; side < 2; ++side) {
                for (int r = 0; r < m->world; ++r) {
                    if (r == m->rank) continue;
                    int qa, qb;
                    mg_owned(m, dz, r, qa, qb);
                    const int a = want[side][0] > qa ? want[side][0] : qa, b = want[side][1] < qb ? want[side][1] : qb;
                    if (a >= b) continue;
                    if (a < fd->p0 || b > fd->p1) return set_error(BMQ_ERR_HALO, "bmq3d_mg: halo planes [%d,%d) are not stored (allocated halo %d)", a, b, s->halo);
                    const int k0r = mg_slab_k0(s->nk, m->world, r);
                    const int q0 = k0r - s->halo > 0 ? k0r - s->halo : 0;     // first plane rank r stores (same halo everywhere)
                    const char *src = (const char *)m->peer[r][fd->alloc_id] + sizeof(float) * fd->plane() * (size_t)(a - q0);
                    char *dst = (char *)fd->alloc + sizeof(float) * fd->plane() * (size_t)(a - fd->p0);
                    const size_t bytes = sizeof(float) * fd->plane() * (size_t)(b - a);
                    if (by_kernel) {
                        if (segs.n == MG_MAX_SEG) {
                            BMQ_CK(launch_mg_pull(m->copy, m->sig, m->near_ranks, epoch, segs));
                            epoch = 0;
                            segs.n = 0;
                        }
                        segs.seg[segs.n++] = MgSeg{src, dst, (unsigned long long)bytes};
                    } else {
                        BMQ_CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, side && two_streams ? m->copy2 : m->copy));
                    }
                    m->stats.bytes_exchanged += (long long)bytes;
                }
            }
        }
    }
    if (by_kernel && (segs.n > 0 || epoch)) BMQ_CK(launch_mg_pull(m->copy, m->sig, m->near_ranks, epoch, segs));
    if (two_streams) {
        cudaEvent_t upper = mg_event(m);
        BMQ_CK(cudaEventRecord(upper, m->copy2));
        BMQ_CK(cudaStreamWaitEvent(m->copy, upper, 0));
    }
    BMQ_CK(cudaEventRecord(m->done_ev[which], m->copy));
    m->stats.exchanges++;
    return BMQ_OK;
}
int mg_wait(bmq3d_mg *m, MgDone which) { BMQ_CK(cudaStreamWaitEvent(m->s->stream, m->done_ev[which], 0)); return BMQ_OK; }
int mg_exchange(bmq3d_mg *m, const MgGroup *groups, int ngroups)
{
    RET_IF(mg_post(m, groups, ngroups, MG_D_BLOCKING));
    return mg_wait(m, MG_D_BLOCKING);
}

// Device-flag signalling only: the reduction as two halves, so that work that does not depend on the result can be
// queued in between.  begin: the one-warp mailbox kernel on the side stream `red`, after everything queued on the
// compute stream so far (dev_vals = result of a kernel there).  end: wait for it, hand out the maxima.
int mg_reduce_begin(bmq3d_mg *m, int n, const float *dev_vals)
{
    if (n > MG_RED_MAX) return set_error(BMQ_ERR_ARG, "bmq3d_mg: reduction of %d values", n);
    cudaEvent_t produced = mg_event(m);
    BMQ_CK(cudaEventRecord(produced, m->s->stream));
    BMQ_CK(cudaStreamWaitEvent(m->red, produced, 0));
    BMQ_CK(launch_mg_allreduce_max(m->red, m->sig, m->other_ranks, nullptr, dev_vals, n, ++m->red_seq, m->red_host));
    return BMQ_OK;
}
int mg_reduce_end(bmq3d_mg *m, float *vals, int n)
{
    BMQ_CK(cudaStreamSynchronize(m->red));
    if (m->red_host[MG_RED_MAX] != 0.f) return set_error(BMQ_ERR_CUDA, "bmq3d_mg: a peer did not answer within 30 s (rank %d)", m->rank);
    for (int q = 0; q < n; ++q) vals[q] = m->red_host[q];
    return BMQ_OK;
}

// dev_vals != nullptr (device-flag signalling only): this rank's contribution is still on the device, `vals` receives the maxima
int mg_reduce(bmq3d_mg *m, float *vals, int n, const float *dev_vals = nullptr)
{
    if (m->world == 1) return BMQ_OK;
    if (m->signal_mode != BMQ_MG_SIGNAL_HOST) {
        // mailbox reduction over peer memory (mg_signal.cu): one one-warp kernel and one stream synchronisation
        if (n > MG_RED_MAX) return set_error(BMQ_ERR_ARG, "bmq3d_mg: reduction of %d values", n);
        BMQ_CK(launch_mg_allreduce_max(m->s->stream, m->sig, m->other_ranks, vals, dev_vals, n, ++m->red_seq, m->red_host));
        BMQ_CK(cudaStreamSynchronize(m->s->stream));
        if (m->red_host[MG_RED_MAX] != 0.f) return set_error(BMQ_ERR_CUDA, "bmq3d_mg: a peer did not answer within 30 s (rank %d)", m->rank);
        for (int q = 0; q < n; ++q) vals[q] = m->red_host[q];
        return BMQ_OK;
    }
    if (!m->allreduce) return set_error(BMQ_ERR_ARG, "bmq3d_mg: no all-reduce callback set");
    if (m->allreduce(vals, n, m->ctx) != 0) return set_error(BMQ_ERR_CUDA, "bmq3d_mg: the all-reduce callback failed");
    return BMQ_OK;
}

const int kVel[3] = {BMQ_F_U, BMQ_F_V, BMQ_F_W};
const int kInitV[3] = {BMQ_F_U_INIT, BMQ_F_V_INIT, BMQ_F_W_INIT}, kInitS[2] = {BMQ_F_RHO_INIT, BMQ_F_T_INIT};
const int kPrevV[3] = {BMQ_F_U_PREV, BMQ_F_V_PREV, BMQ_F_W_PREV}, kPrevS[2] = {BMQ_F_RHO_PREV, BMQ_F_T_PREV};
const int kAdvV[3] = {BMQ_F_U_ADV, BMQ_F_V_ADV, BMQ_F_W_ADV}, kAdvS[2] = {BMQ_F_RHO_ADV, BMQ_F_T_ADV};
const int kErrV[3] = {BMQ_F_U_ERR, BMQ_F_V_ERR, BMQ_F_W_ERR}, kErrS[2] = {BMQ_F_RHO_ERR, BMQ_F_T_ERR};
const int kChV[6] = {BMQ_F_DU_EXT, BMQ_F_DV_EXT, BMQ_F_DW_EXT, BMQ_F_DU_PROJ, BMQ_F_DV_PROJ, BMQ_F_DW_PROJ};
const int kChS[2] = {BMQ_F_DRHO_EXT, BMQ_F_DT_EXT};
const int kBwdV[3] = {BMQ_F_VBWD_X, BMQ_F_VBWD_Y, BMQ_F_VBWD_Z}, kBwdS[3] = {BMQ_F_SBWD_X, BMQ_F_SBWD_Y, BMQ_F_SBWD_Z};
const int kFwdV[3] = {BMQ_F_VFWD_X, BMQ_F_VFWD_Y, BMQ_F_VFWD_Z}, kFwdS[3] = {BMQ_F_SFWD_X, BMQ_F_SFWD_Y, BMQ_F_SFWD_Z};
const int kBwdpV[3] = {BMQ_F_VBWDP_X, BMQ_F_VBWDP_Y, BMQ_F_VBWDP_Z}, kBwdpS[3] = {BMQ_F_SBWDP_X, BMQ_F_SBWDP_Y, BMQ_F_SBWDP_Z};

void mg_unmap(bmq3d_mg *m)
{
    for (int r = 0; r < (int)m->peer.size(); ++r) {
        if (r < (int)m->peer_is_ipc.size() && m->peer_is_ipc[r]) {
            for (void *p : m->peer[r]) if (p) cudaIpcCloseMemHandle(p);
            if (r < (int)m->peer_sig.size() && m->peer_sig[r]) cudaIpcCloseMemHandle(m->peer_sig[r]);
        }
        m->peer[r].clear();
    }
    m->peer.clear();
    m->peer_sig.clear();
    m->peer_is_ipc.clear();
    m->near_ranks.n = m->other_ranks.n = 0;
    m->signal_mode = BMQ_MG_SIGNAL_HOST;
}

}  // namespace

extern "C" {

int bmq3d_mg_create(int ni, int nj, int nk, float h, float blend_coeff, int rank, int world, int halo, bmq3d_mg **out)
{
    if (!out) return set_error(BMQ_ERR_ARG, "bmq3d_mg_create: out is null");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world || nk / world < 1)
        return set_error(BMQ_ERR_ARG, "bmq3d_mg_create: bad rank %d of %d for %d planes", rank, world, nk);
    bmq3d_mg *m = new (std::nothrow) bmq3d_mg();
    if (!m) return set_error(BMQ_ERR_ARG, "bmq3d_mg_create: out of host memory");
    m->rank = rank; m->world = world;
    memset(&m->stats, 0, sizeof m->stats);
    if (halo > nk) halo = nk;
    int st = bmq3d_create_slab(ni, nj, nk, h, blend_coeff, mg_slab_k0(nk, world, rank), mg_slab_k0(nk, world, rank + 1), world > 1 ? halo : 0, &m->s);
    if (st == BMQ_OK) {
        // the exchange kernels go first whenever an SM has room: a posted halo must not queue behind the
        // thousands of CTAs of the gather kernel it overlaps
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        st = check_cuda(cudaStreamCreateWithPriority(&m->copy, cudaStreamNonBlocking, hi), "cudaStreamCreate", __FILE__, __LINE__);
        if (st == BMQ_OK) st = check_cuda(cudaStreamCreateWithPriority(&m->copy2, cudaStreamNonBlocking, hi), "cudaStreamCreate", __FILE__, __LINE__);
        if (st == BMQ_OK) st = check_cuda(cudaStreamCreateWithPriority(&m->red, cudaStreamNonBlocking, hi), "cudaStreamCreate", __FILE__, __LINE__);
    }
    if (st == BMQ_OK && world > MG_MAX_WORLD) st = set_error(BMQ_ERR_ARG, "bmq3d_mg_create: at most %d ranks", (int)MG_MAX_WORLD);
    // its own 2 MB allocation: a CUDA IPC handle names a whole allocation, and a block shared with other small
    // buffers could not be opened a second time by a peer
    if (st == BMQ_OK) st = check_cuda(cudaMalloc(&m->sig, 2u << 20), "cudaMalloc", __FILE__, __LINE__);
    if (st == BMQ_OK) st = check_cuda(cudaMemset(m->sig, 0, 2u << 20), "cudaMemset", __FILE__, __LINE__);
    if (st == BMQ_OK) st = check_cuda(cudaHostAlloc(&m->red_host, (MG_RED_MAX + 4) * sizeof(float), cudaHostAllocDefault), "cudaHostAlloc", __FILE__, __LINE__);
    for (int e = 0; e < MG_EVENTS && st == BMQ_OK; ++e)
        st = check_cuda(cudaEventCreateWithFlags(&m->ev[e], cudaEventDisableTiming), "cudaEventCreate", __FILE__, __LINE__);
    for (int e = 0; e < MG_DONE_EVENTS && st == BMQ_OK; ++e)
        st = check_cuda(cudaEventCreateWithFlags(&m->done_ev[e], cudaEventDisableTiming), "cudaEventCreate", __FILE__, __LINE__);
    if (st != BMQ_OK) { bmq3d_mg_destroy(m); return st; }
    m->stats.halo_allocated = m->s->halo;
    *out = m;
    return BMQ_OK;
}

int bmq3d_mg_destroy(bmq3d_mg *m)
{
    if (!m) return BMQ_OK;
    if (m->copy) cudaStreamSynchronize(m->copy);
    if (m->copy2) cudaStreamSynchronize(m->copy2);
    mg_unmap(m);
    for (auto e : m->ev) if (e) cudaEventDestroy(e);
    for (auto e : m->done_ev) if (e) cudaEventDestroy(e);
    if (m->copy) cudaStreamDestroy(m->copy);
    if (m->copy2) cudaStreamDestroy(m->copy2);
    if (m->red) cudaStreamDestroy(m->red);
    if (m->sig) cudaFree(m->sig);
    if (m->red_host) cudaFreeHost(m->red_host);
    if (m->s) bmq3d_destroy(m->s);
    delete m;
    return BMQ_OK;
}

int bmq3d_mg_solver(bmq3d_mg *m, bmq3d_solver **out)
{
    if (!m || !out) return set_error(BMQ_ERR_ARG, "bmq3d_mg_solver: null argument");
    *out = m->s;
    return BMQ_OK;
}

int bmq3d_mg_set_collectives(bmq3d_mg *m, bmq_allreduce_max_fn allreduce_max, bmq_stream_barrier_fn stream_barrier, void *ctx)
{
    if (!m) return set_error(BMQ_ERR_ARG, "bmq3d_mg_set_collectives: null handle");
    m->allreduce = allreduce_max; m->barrier = stream_barrier; m->ctx = ctx;
    return BMQ_OK;
}

int bmq3d_mg_export_size(bmq3d_mg *m, size_t *bytes)
{
    if (!m || !bytes) return set_error(BMQ_ERR_ARG, "bmq3d_mg_export_size: null argument");
    *bytes = sizeof(MgBlobHeader) + sizeof(MgBlobEntry) * (size_t)m->s->n_allocs;
    return BMQ_OK;
}

int bmq3d_mg_export(bmq3d_mg *m, void *blob)
{
    if (!m || !blob) return set_error(BMQ_ERR_ARG, "bmq3d_mg_export: null argument");
    bmq3d_solver *s = m->s;
    MgBlobHeader hd{(long long)getpid(), s->n_allocs, s->halo, m->rank, 0, {0}, m->sig};
    {
        cudaIpcMemHandle_t hnd;
        BMQ_CK(cudaIpcGetMemHandle(&hnd, m->sig));
        memcpy(hd.sig_ipc, &hnd, 64);
    }
    memcpy(blob, &hd, sizeof hd);
    MgBlobEntry *ent = reinterpret_cast<MgBlobEntry *>((char *)blob + sizeof hd);
    memset(ent, 0, sizeof(MgBlobEntry) * (size_t)s->n_allocs);
    auto put = [&](const Field &fd) -> int {
        if (!fd.alloc || fd.alloc_id < 0) return BMQ_OK;
        cudaIpcMemHandle_t hnd;
        BMQ_CK(cudaIpcGetMemHandle(&hnd, fd.alloc));
        memcpy(ent[fd.alloc_id].ipc, &hnd, 64);
        ent[fd.alloc_id].raw = fd.alloc;
        return BMQ_OK;
    };
    for (auto &fd : s->f) RET_IF(put(fd));
    for (auto &fd : s->scratch) RET_IF(put(fd));
    for (auto &fd : s->tmpmap) RET_IF(put(fd));
    return BMQ_OK;
}

int bmq3d_mg_disconnect(bmq3d_mg *m)
{
    if (!m) return set_error(BMQ_ERR_ARG, "bmq3d_mg_disconnect: null handle");
    BMQ_CK(cudaStreamSynchronize(m->copy));
    BMQ_CK(cudaStreamSynchronize(m->copy2));
    BMQ_CK(cudaStreamSynchronize(m->s->stream));
    mg_unmap(m);
    return BMQ_OK;
}

int bmq3d_mg_connect(bmq3d_mg *m, const void *all_blobs)
{
    if (!m || !all_blobs) return set_error(BMQ_ERR_ARG, "bmq3d_mg_connect: null argument");
    size_t one = 0;
    RET_IF(bmq3d_mg_export_size(m, &one));
    mg_unmap(m);
    m->peer.assign(m->world, std::vector<void *>());
    m->peer_sig.assign(m->world, nullptr);
    m->peer_is_ipc.assign(m->world, 0);
    bool any_same_process = false;
    for (int r = 0; r < m->world; ++r) {
        if (r == m->rank) continue;
        const char *blob = (const char *)all_blobs + one * (size_t)r;
        MgBlobHeader hd;
        memcpy(&hd, blob, sizeof hd);
        if (hd.rank != r || hd.n_allocs != m->s->n_allocs || hd.halo != m->s->halo)
            return set_error(BMQ_ERR_ARG, "bmq3d_mg_connect: blob %d does not match (rank %d, %d allocations, halo %d)", r, hd.rank, hd.n_allocs, hd.halo);
        const MgBlobEntry *ent = reinterpret_cast<const MgBlobEntry *>(blob + sizeof hd);
        m->peer[r].assign(hd.n_allocs, nullptr);
        const bool same_process = hd.pid == (long long)getpid();
        m->peer_is_ipc[r] = same_process ? 0 : 1;
        any_same_process |= same_process;
        if (same_process) m->peer_sig[r] = static_cast<MgSignal *>(hd.sig_raw);
        else {
            cudaIpcMemHandle_t hnd;
            memcpy(&hnd, hd.sig_ipc, 64);
            void *p = nullptr;
            BMQ_CK(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
            m->peer_sig[r] = static_cast<MgSignal *>(p);
        }
        for (int a = 0; a < hd.n_allocs; ++a) {
            if (!ent[a].raw) continue;
            if (same_process) { m->peer[r][a] = ent[a].raw; continue; }
            cudaIpcMemHandle_t hnd;
            memcpy(&hnd, ent[a].ipc, 64);
            BMQ_CK(cudaIpcOpenMemHandle(&m->peer[r][a], hnd, cudaIpcMemLazyEnablePeerAccess));
        }
    }
    // who takes part in this rank's exchange barrier: every rank whose slab lies within the allocated halo of
    // this one (either may then pull from the other, whatever width an exchange uses); reductions involve everybody
    m->near_ranks.n = m->other_ranks.n = 0;
    for (int r = 0; r < m->world; ++r) {
        if (r == m->rank) continue;
        m->other_ranks.sig[m->other_ranks.n++] = m->peer_sig[r];
        const int gap = r > m->rank ? mg_slab_k0(m->s->nk, m->world, r) - m->s->k1 : m->s->k0 - mg_slab_k0(m->s->nk, m->world, r + 1);
        if (gap < m->s->halo + 2) m->near_ranks.sig[m->near_ranks.n++] = m->peer_sig[r];
    }
    // device-side signalling spins in kernels: the default for one process per GPU; ranks that share a process
    // (and possibly a GPU, where a spinning kernel could keep its peer's kernel from running) stay with the callbacks
    int mode = m->signal_request;
    if (mode == BMQ_MG_SIGNAL_AUTO) {
        // measured at 8 GPUs, 512^3: flags + copy engines 11.97 ms per step, flags + pull kernel 12.60 (its CTAs take SM
        // time from the stages they overlap), host collectives 12.40 (profiles/r2_scaling.md)
        mode = any_same_process ? BMQ_MG_SIGNAL_HOST : BMQ_MG_SIGNAL_DEVICE_CE;
        if (const char *e = getenv("BMQ_MG_SIGNAL")) mode = atoi(e);
    }
    if (mode != BMQ_MG_SIGNAL_HOST && mode != BMQ_MG_SIGNAL_DEVICE && mode != BMQ_MG_SIGNAL_DEVICE_CE)
        return set_error(BMQ_ERR_ARG, "bmq3d_mg_connect: unknown signalling mode %d", mode);
    m->signal_mode = mode;
    m->stats.signalling = mode;
    return BMQ_OK;
}

int bmq3d_mg_set_signalling(bmq3d_mg *m, int mode)
{
    if (!m) return set_error(BMQ_ERR_ARG, "bmq3d_mg_set_signalling: null handle");
    if (mode < BMQ_MG_SIGNAL_AUTO || mode > BMQ_MG_SIGNAL_DEVICE_CE) return set_error(BMQ_ERR_ARG, "bmq3d_mg_set_signalling: unknown mode %d", mode);
    if (!m->peer.empty()) return set_error(BMQ_ERR_ARG, "bmq3d_mg_set_signalling: call it before bmq3d_mg_connect (every rank the same mode)");
    m->signal_request = mode;
    return BMQ_OK;
}

int bmq3d_mg_grow_halo(bmq3d_mg *m, int new_halo)
{
    if (!m) return set_error(BMQ_ERR_ARG, "bmq3d_mg_grow_halo: null handle");
    if (!m->peer.empty()) return set_error(BMQ_ERR_ARG, "bmq3d_mg_grow_halo: disconnect first (peers still map the old buffers)");
    if (new_halo > m->s->nk) new_halo = m->s->nk;
    RET_IF(bmq3d_grow_halo(m->s, new_halo));
    m->stats.halo_allocated = m->s->halo;
    m->stats.halo_grown++;
    return BMQ_OK;
}

// Phase A over z-slabs (zslab.py:ZSlabStepper.advect).  BMQ_ERR_HALO = the allocated halo is narrower than
// stats.halo_needed: nothing has been modified yet; disconnect, grow, export / connect again and call again.
int bmq3d_mg_advect(bmq3d_mg *m, int framenum, float dt)
{
    if (!m) return set_error(BMQ_ERR_ARG, "bmq3d_mg_advect: null handle");
    if (!(dt > 0.f)) return set_error(BMQ_ERR_ARG, "bmq3d_mg_advect: dt must be positive");
    bmq3d_solver *s = m->s;
    if (m->world > 1 && m->peer.empty()) return set_error(BMQ_ERR_ARG, "bmq3d_mg_advect: not connected");
    float gmax = 0.f;
    const bool on_device = m->world > 1 && m->signal_mode != BMQ_MG_SIGNAL_HOST;   // reductions without a host round trip in between
    // The velocity halo (read by the DMC and the forward trace) does not have to wait for the max-velocity reduction
    // that sizes it: post it now with last step's width plus two planes, top it up afterwards in the rare case that
    // the true width is larger.  It then travels while the reduction makes its host round trip.
    int w_guess = 0;
    if (m->world > 1) {
        w_guess = (m->wv > m->ws ? m->wv : m->ws) + 2;
        if (w_guess > s->halo) w_guess = s->halo;
        const MgGroup g[1] = {{kVel, 3, w_guess}};
        RET_IF(mg_post(m, g, 1, MG_D_VEL));
    }
    if (on_device) {
        RET_IF(stage_maxvel(s, nullptr));
        RET_IF(mg_reduce(m, &gmax, 1, s->d_red));
    } else {
        RET_IF(stage_maxvel(s, &gmax));
        RET_IF(mg_reduce(m, &gmax, 1));
    }
    set_cfl(s, framenum, gmax);
    const float cfl_frame = dt * (gmax > 1e-4f ? gmax : 1e-4f) / s->h;
    int need[2], need_b[2] = {0, 0}, widest = MG_NARROW;
    bool blend_on[2];
    for (int q = 0; q < 2; ++q) {
        need[q] = (int)ceilf(m->disp[q] + cfl_frame) + 3;
        blend_on[q] = s->blend != 1.0f && m->reinit_count[q] > 0;
        if (blend_on[q]) need_b[q] = (int)ceilf(m->disp_prev[q] + m->disp[q] + cfl_frame) + 3;
        widest = need[q] > widest ? need[q] : widest;
        widest = need_b[q] > widest ? need_b[q] : widest;
    }
    m->stats.halo_needed = widest;
    if (m->world > 1 && widest > s->halo)
        return set_error(BMQ_ERR_HALO, "bmq3d_mg_advect: need %d halo planes, %d allocated: grow and call again", widest, s->halo);
    const int wv = m->wv = need[0], ws = m->ws = need[1], wmax = wv > ws ? wv : ws;
    m->stats.halo_vel = wv; m->stats.halo_scalar = ws;
    if (m->world == 1) {         // a single rank: the plain phase A
        return bmq3d_advect(s, framenum, dt, 0);
    }
    RET_IF(mg_wait(m, MG_D_VEL));                                                                   // consumers: DMC, forward
    if (wmax > w_guess) { const MgGroup g[1] = {{kVel, 3, wmax}}; RET_IF(mg_exchange(m, g, 1)); }
    {   // the DMC sub-step reads 5 planes of chi past the slab; they are still there from the last step's wide exchange
        // (or from the identity fill of a re-initialisation) unless that step needed fewer than 5
        MgGroup g[2];
        int ng = 0;
        if (m->chi_valid[0] < MG_NARROW) g[ng++] = MgGroup{kBwdV, 3, MG_NARROW};
        if (m->chi_valid[1] < MG_NARROW) g[ng++] = MgGroup{kBwdS, 3, MG_NARROW};
        if (ng) RET_IF(mg_exchange(m, g, ng));
    }
    { const MgGroup g[2] = {{kInitV, 3, wv}, {kInitS, 2, ws}}; RET_IF(mg_post(m, g, 2, MG_D_INIT)); }   // overlaps DMC + forward
    float T = 0.f, substep = s->cfldt;
    int n = 0;
    while (T < dt) {
        if (T + substep > dt) substep = dt - T;
        RET_IF(stage_dmc(s, substep));
        T += substep;
        if (++n > 4096) return set_error(BMQ_ERR_ARG, "bmq3d_mg_advect: more than 4096 CFL sub-steps");
        if (T < dt) { const MgGroup g[2] = {{kBwdV, 3, MG_NARROW}, {kBwdS, 3, MG_NARROW}}; RET_IF(mg_exchange(m, g, 2)); }
        else {
            const MgGroup g[2] = {{kBwdV, 3, wv}, {kBwdS, 3, ws}};
            RET_IF(mg_post(m, g, 2, MG_D_BWD));                                                       // overlaps forward
            m->chi_valid[0] = wv; m->chi_valid[1] = ws;
        }
    }
    s->stats.n_substeps = n;
    RET_IF(stage_forward(s, dt));
    { const MgGroup g[2] = {{kFwdV, 3, wv}, {kFwdS, 3, ws}}; RET_IF(mg_post(m, g, 2, MG_D_FWD)); }      // overlaps advect
    RET_IF(mg_wait(m, MG_D_BWD));
    RET_IF(mg_wait(m, MG_D_INIT));
    RET_IF(stage_advect(s, 0));
    { const MgGroup g[1] = {{kAdvV, 3, wv}}; RET_IF(mg_post(m, g, 1, MG_D_ADV_V)); }
    RET_IF(stage_advect(s, 1));
    { const MgGroup g[1] = {{kAdvS, 2, ws}}; RET_IF(mg_post(m, g, 1, MG_D_ADV_S)); }
    RET_IF(mg_wait(m, MG_D_FWD));
    RET_IF(mg_wait(m, MG_D_ADV_V));
    RET_IF(stage_error(s, 0));
    { const MgGroup g[1] = {{kErrV, 3, wv}}; RET_IF(mg_post(m, g, 1, MG_D_ERR_V)); }
    RET_IF(mg_wait(m, MG_D_ADV_S));
    RET_IF(stage_error(s, 1));
    { const MgGroup g[1] = {{kErrS, 2, ws}}; RET_IF(mg_post(m, g, 1, MG_D_ERR_S)); }
    RET_IF(mg_wait(m, MG_D_ERR_V));
    RET_IF(stage_apply(s, 0));
    RET_IF(mg_wait(m, MG_D_ERR_S));
    RET_IF(stage_apply(s, 1));
    for (int which = 0; which < 2; ++which) {
        if (!blend_on[which]) continue;
        const MgGroup g[2] = {{which ? kPrevS : kPrevV, which ? 2 : 3, need_b[which]}, {which ? kBwdpS : kBwdpV, 3, need_b[which]}};
        RET_IF(mg_exchange(m, g, 2));
        RET_IF(stage_blend(s, which));
    }
    return BMQ_OK;
}

// Phase B over z-slabs (zslab.py:ZSlabStepper.accumulate)
int bmq3d_mg_accumulate(bmq3d_mg *m, int framenum, float dt)
{
    if (!m) return set_error(BMQ_ERR_ARG, "bmq3d_mg_accumulate: null handle");
    bmq3d_solver *s = m->s;
    if (m->world > 1) {
        // the scalars' two change fields first, then the velocity's six: the scalar accumulation runs while the larger
        // exchange is still travelling (both overlap the distortion kernel and its reduction)
        const MgGroup gs[1] = {{kChS, 2, m->ws}}, gv[1] = {{kChV, 6, m->wv}};
        RET_IF(mg_post(m, gs, 1, MG_D_CH_S));
        RET_IF(mg_post(m, gv, 1, MG_D_CH_V));
    }
    float red[4] = {0, 0, 0, 0};
    const bool on_device = m->world > 1 && m->signal_mode != BMQ_MG_SIGNAL_HOST;
    if (on_device) {
        // The scalars' accumulation does not depend on the re-initialisation decision (the velocity's does, through the
        // projection coefficient): it is queued BEFORE the host waits for the reduced distortion, and the reduction runs
        // beside it on its own stream.  The wait for the slowest rank and the host round trip hide behind it.
        RET_IF(stage_distortion(s, nullptr, nullptr, nullptr));
        RET_IF(mg_reduce_begin(m, 4, s->d_red + 1));
        RET_IF(mg_wait(m, MG_D_CH_S));
        RET_IF(stage_accumulate(s, 1));
        RET_IF(mg_reduce_end(m, red, 4));
        red[2] /= s->h; red[3] /= s->h;                         // as stage_distortion reports them: in cells
    } else {
        RET_IF(stage_distortion(s, &red[0], &red[1], &red[2]));     // red[2], red[3] = z displacement per mapper
        RET_IF(mg_reduce(m, red, 4));
    }
    m->disp[0] = red[2]; m->disp[1] = red[3];
    s->stats.max_disp_z_vel = red[2]; s->stats.max_disp_z_scalar = red[3];
    s->stats.max_disp_z = red[2] > red[3] ? red[2] : red[3];
    decide(s, framenum, dt, red[0], red[1]);
    if (!on_device) {
        if (m->world > 1) RET_IF(mg_wait(m, MG_D_CH_S));
        RET_IF(stage_accumulate(s, 1));         // independent of the velocity's accumulation: any order gives the same fields
    }
    if (m->world > 1) RET_IF(mg_wait(m, MG_D_CH_V));
    RET_IF(stage_accumulate(s, 0));
    if (s->vel_reinit) {
        RET_IF(stage_reinit(s, 0, 0));
        RET_IF(stage_reinit(s, 0, 1));
        m->reinit_count[0]++;
        m->disp_prev[0] = m->disp[0]; m->disp[0] = 0.f;
        m->chi_valid[0] = s->halo;                               // the identity fill covers every stored plane
    }
    if (s->scalar_reinit) {
        RET_IF(stage_reinit(s, 1, 0));
        m->reinit_count[1]++;
        m->disp_prev[1] = m->disp[1]; m->disp[1] = 0.f;
        m->chi_valid[1] = s->halo;
    }
    s->stats.vel_reinit_count = s->vel_reinit_count;
    s->stats.scalar_reinit_count = s->scalar_reinit_count;
    return BMQ_OK;
}

int bmq3d_mg_get_stats(bmq3d_mg *m, bmq3d_mg_stats *out)
{
    if (!m || !out) return set_error(BMQ_ERR_ARG, "bmq3d_mg_get_stats: null argument");
    *out = m->stats;
    return BMQ_OK;
}

}  // extern "C"
