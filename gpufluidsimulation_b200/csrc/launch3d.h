// Host-side launch interface of the 3D kernels (kernels3d.cu).  All pointers are device
// pointers to VIRTUAL BASES (plane 0 of the global grid); KRange is the global plane range a
// launch covers.  Every launcher returns the launch status and counts one kernel launch.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace bmq {

struct Grid3;                       // device3d.cuh
struct KRange { int kbeg, kend; };  // global planes [kbeg, kend)
struct Stag { int dx, dy, dz; };    // staggering of a field: (1,0,0)=u, (0,1,0)=v, (0,0,1)=w, 0=centred

template <int N> struct MapSetRO { const float *x[N], *y[N], *z[N]; };
template <int N> struct MapSetRW { float *x[N], *y[N], *z[N]; };
template <int N> struct FieldSetRO { const float *p[N]; };
template <int N> struct FieldSetRW { float *p[N]; };
template <int N> struct Coeffs { float c[N]; };
template <int N> struct DistOut { float *dist[N]; float *d2max[N]; float *dispz[N]; };
struct IdentityOut { float *x[4], *y[4], *z[4]; };

}  // namespace bmq

#include "device3d.cuh"

namespace bmq {

Grid3 make_grid(int ni, int nj, int nk, float h);
unsigned long long kernel_launch_count();
void set_pitch_specialisation(bool on);   // testing knob, see bmq_set_pitch_specialisation
void set_gather_variant(int v);           // testing knob, see bmq_set_gather_variant
void set_fast_division(bool on);          // testing knob, see bmq_set_fast_division
void set_tolerance_mode(bool on);         // opt-in, see bmq_set_tolerance_mode
bool tolerance_mode();
bool division_is_fast(float h, int nmax);
// div_h (device3d.cuh) == IEEE division by h for every float in {0} U [2^-100, p_max]?  Checked on the device, cached per h.
bool division_verified(float h, float p_max);
// z-marching gather kernels (march_*.cu); same contracts as launch_advect / _error / _cumulate / _apply_clamp
// with is_point == false
cudaError_t launch_advect_march(cudaStream_t s, const Grid3 &g, KRange r, Stag st, int nf, float *const *out,
                                const float *const *init, const float *const chi[3]);
cudaError_t launch_error_march(cudaStream_t s, const Grid3 &g, KRange r, Stag st, int nf, float *const *e0,
                               const float *const *src, const float *const *init, const float *const psi[3]);
cudaError_t launch_cumulate_march(cudaStream_t s, const Grid3 &g, KRange r, Stag st, int nf, int nch,
                                  float *const *target, const float *const *change, const float *coeff,
                                  const float *const map[3]);
cudaError_t launch_apply_march(cudaStream_t s, const Grid3 &g, KRange r, Stag st, int nf, float *const *out,
                               const float *const *fadv, const float *const *e0, const float *const chi[3]);
cudaError_t launch_apply_march_split(cudaStream_t s, const Grid3 &g, KRange r, Stag st, int nf, float *const *out,
                               const float *const *fadv, const float *const *e0, const float *const chi[3]);
cudaError_t launch_clamp27(cudaStream_t s, int fi, int fj, int fk, KRange r, int nf, const float *const *before, float *const *f);

cudaError_t launch_forward(cudaStream_t s, const Grid3 &g, KRange r, const float *u, const float *v,
                           const float *w, int nmap, float *const maps[][3], float cfldt, float dt);
cudaError_t launch_dmc(cudaStream_t s, const Grid3 &g, KRange r, const float *u, const float *v,
                       const float *w, int nmap, const float *const in[][3], float *const out[][3],
                       float substep);
cudaError_t launch_semilag(cudaStream_t s, const Grid3 &g, KRange r, Stag st, const float *u,
                           const float *v, const float *w, int nf, float *const *out,
                           const float *const *src, float cfldt, float dt);
cudaError_t launch_advect(cudaStream_t s, const Grid3 &g, KRange r, Stag st, bool is_point, int nf,
                          float *const *out, const float *const *init, const float *const chi[3]);
cudaError_t launch_error(cudaStream_t s, const Grid3 &g, KRange r, Stag st, bool is_point, int nf,
                         float *const *e0, const float *const *src, const float *const *init,
                         const float *const psi[3]);
// change is laid out [set][field]: change[c*nf + f]
cudaError_t launch_cumulate(cudaStream_t s, const Grid3 &g, KRange r, Stag st, bool is_point, int nf,
                            int nch, float *const *target, const float *const *change,
                            const float *coeff, const float *const map[3]);
cudaError_t launch_apply_clamp(cudaStream_t s, const Grid3 &g, KRange r, Stag st, bool is_point, int nf,
                               float *const *out, const float *const *fadv, const float *const *e0,
                               const float *const chi[3]);
cudaError_t launch_clamp_extrema(cudaStream_t s, int fi, int fj, int fk, KRange r, const float *before,
                                 float *after);
cudaError_t launch_double_advect(cudaStream_t s, const Grid3 &g, KRange r, Stag st, bool is_point, int nf,
                                 float *const *field, const float *const *prev, const float *const chi[3],
                                 const float *const chip[3], float blend);
// dist / d2max / dispz may be null (or hold null entries).  d2max and dispz (one per mapper) are
// atomically max-ed into, so zero them first.
cudaError_t launch_estimate(cudaStream_t s, const Grid3 &g, KRange r, int nmap, const float *const bwd[][3],
                            const float *const fwd[][3], float *const *dist, float *const *d2max,
                            float *const *dispz, const signed char *boundary);
cudaError_t launch_maxabs3(cudaStream_t s, const float *a, size_t na, const float *b, size_t nb,
                           const float *c, size_t nc, float *out_dev);
cudaError_t launch_axpy(cudaStream_t s, float *a, const float *b, float c, size_t n);
cudaError_t launch_add_field(cudaStream_t s, float *out, const float *a, const float *b, float c, size_t n);
// 8^3-blocked buffer3Df layout <-> linear (sources3d.cu)
size_t blocked_elems(int nx, int ny, int nz);
cudaError_t launch_relayout(cudaStream_t s, bool to_linear, const float *src, float *dst, int nx, int ny, int nz);
cudaError_t launch_identity(cudaStream_t s, const Grid3 &g, KRange r, int nsets, float *const sets[][3]);

}  // namespace bmq
