/*
 * bimocq_b200.h -- C ABI of libbimocq_b200.so: the BiMocq^2 advection hot path for NVIDIA B200
 * (sm_100a).  Plain pointers and sizes only; no C++ or torch types cross this boundary.
 *
 * Two groups of entry points:
 *
 *  1. LEGACY DROP-IN SYMBOLS.  The 14 hot-path `extern "C" void gpu_*` functions of the
 *     reference -- and, widened per SURVEY 8(f), the other 8 (source terms, clamp, the three
 *     pressure solvers): all 22 of GPU_Advection.h:26-108 -- with byte-identical prototypes, so that the reference's gpuMapper /
 *     MapperBase / MapperBaseGPU (bimocq3D/GPU_Advection.h:110-627, bimocq3D/Mapping.cpp) link
 *     against this library unchanged.  Each prototype cites the reference declaration it
 *     replaces.  All pointers are DEVICE pointers to dense x-fastest float arrays
 *     (idx = i + nx*j + nx*ny*k) of the sizes the reference uses; they run on the legacy
 *     default stream, return void, and latch errors for bmq_last_error().  Like the reference's
 *     gpuMapper they are meant for ONE host thread per device (shared default stream, per-device
 *     scratch); the handle APIs below are safe to use from different threads on different handles.
 *
 *  2. HANDLE API (bmq3d_*, bmq2d_*, bmq_mgpcg_*).  Device-resident solver state with the fused kernels and the
 *     reinitialisation scheduler of BimocqSolver::advanceBimocq (bimocq3D/BimocqSolver.cpp:88-230)
 *     inside.  Returns int status codes (0 = BMQ_OK), never exits the process.
 *
 * There is no CPU fallback: every entry point needs a CUDA device and fails loudly without one.
 */
#ifndef BIMOCQ_B200_H
#define BIMOCQ_B200_H

#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status / errors */
enum {
    BMQ_OK = 0,
    BMQ_ERR_CUDA = 1,       /* a CUDA runtime call or kernel launch failed */
    BMQ_ERR_ARG = 2,        /* bad argument (null handle, bad field id, bad size) */
    BMQ_ERR_HALO = 3,       /* z-slab halo narrower than the measured map displacement */
    BMQ_ERR_NODEVICE = 4    /* no CUDA device visible */
};
/* Last error message recorded on the calling thread's process (empty string if none). */
const char *bmq_last_error(void);
/* Clears the latched error; returns the status code that was latched. */
int bmq_clear_error(void);
/* Library version string, e.g. "bimocq_b200 0.1 (sm_100a)". */
const char *bmq_version(void);
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
unsigned long long bmq_kernel_launch_count(void);
/* Testing knob.  Grids with ni == nj in {128, 256, 512} run kernels whose
 * row and plane pitches are compile-time constants (same arithmetic, 17 % fewer instructions); 0
 * switches them off so that tests can compare the two paths bit for bit.  Default: on. */
int bmq_set_pitch_specialisation(int on);
/* Testing knob.  The advect / error / apply / accumulate gathers run as z-marching column kernels
 * (variant 1: a thread keeps the x-y-interpolated map planes of its column in registers, the extrema clamp fused
 * into the apply kernel), the same with the clamp as its own shared-memory tiled stencil kernel (variant 2), or as
 * one windowed cell per thread (variant 0).  Same arithmetic, bit-identical results. */
int bmq_set_gather_variant(int variant);
/* Division by the cell size.  The reference computes pos / h (GPU_kernel.cu:46-51); when h is not a power
 * of two the kernels use a three-instruction sequence (multiply by RN(1/h), exact residual, correction) that
 * the library has checked against IEEE division for EVERY float a position can take, once per h, on the
 * device; if a single quotient differs (or with bmq_set_fast_division(0), a testing knob that applies to
 * handles and grids created afterwards) they use IEEE division.  bmq_division_is_fast reports the outcome for
 * a grid of nmax cells per axis (1 = fast sequence in use; power-of-two h: always 1, no division at all). */
int bmq_set_fast_division(int on);
int bmq_division_is_fast(float h, int nmax);
/* OPT-IN tolerance mode (default off = the bit-exact paths above).  With a cell size that is not a power of two the
 * exact path pays for the reference's own rounding: a correctly rounded p / h per sampled coordinate, the generic
 * 8-node centre sample, double-precision lerps in the DMC update.  In tolerance mode every cell size runs the kernels
 * written for a power of two (one multiplication by RN(1/h), grid-unit positions, node shortcuts).  One step from
 * identical state differs from the reference by rounding noise (<= 5e-6 relative L-inf); over a run the difference
 * grows to ~1e-3, because the reference's DMC formula (1 - exp(-a s) in fp32) amplifies relative differences of its
 * velocity input several hundred times (tests/test_tolerance_mode_gpu.py, tests/test_oracle_cpu.py).  An experiment that prices exactness, not a drop-in mode.
 * Applies to handles and grids created afterwards; the 3D paths only. */
int bmq_set_tolerance_mode(int on);
int bmq_tolerance_mode(void);

/* ---- peer-memory plumbing for the z-slab halo exchange over NVLink (one process per GPU).
 * bmq_ipc_export: CUDA IPC handle (64 bytes) of the allocation `dev_ptr` is the base of;
 * bmq_ipc_open: map a peer process's allocation into this process (lazy peer access);
 * bmq_copy_async: stream-ordered device-to-device copy that accepts peer-mapped pointers. */
int bmq_ipc_export(const void *dev_ptr, unsigned char handle[64]);
int bmq_ipc_open(const unsigned char handle[64], void **dev_ptr);
int bmq_ipc_close(void *dev_ptr);
int bmq_copy_async(void *dst, const void *src, size_t bytes, void *stream);

/* ------------------------------------------------------------------ legacy drop-in symbols */
/* replaces GPU_Advection.h:26-28 (def. GPU_kernel.cu:567-574): psi <- trace(psi, +dt), in place */
void gpu_solve_forward(float *u, float *v, float *w, float *x_fwd, float *y_fwd, float *z_fwd,
                       float h, int ni, int nj, int nk, float cfldt, float dt);
/* replaces GPU_Advection.h:30-33 (def. GPU_kernel.cu:576-584): one DMC sub-step, in -> out */
void gpu_solve_backwardDMC(float *u, float *v, float *w, float *x_in, float *y_in, float *z_in,
                           float *x_out, float *y_out, float *z_out, float h, int ni, int nj,
                           int nk, float substep);
/* replaces GPU_Advection.h:35-38 (def. GPU_kernel.cu:586-598) */
void gpu_advect_velocity(float *u, float *v, float *w, float *u_init, float *v_init,
                         float *w_init, float *backward_x, float *backward_y, float *backward_z,
                         float h, int ni, int nj, int nk, bool is_point);
/* replaces GPU_Advection.h:40-44 (def. GPU_kernel.cu:600-618) */
void gpu_advect_vel_double(float *u, float *v, float *w, float *utemp, float *vtemp, float *wtemp,
                           float *backward_x, float *backward_y, float *backward_z,
                           float *backward_xprev, float *backward_yprev, float *backward_zprev,
                           float h, int ni, int nj, int nk, bool is_point, float blend_coeff);
/* replaces GPU_Advection.h:46-48 (def. GPU_kernel.cu:620-627) */
void gpu_advect_field(float *field, float *field_init, float *backward_x, float *backward_y,
                      float *backward_z, float h, int ni, int nj, int nk, bool is_point);
/* replaces GPU_Advection.h:50-53 (def. GPU_kernel.cu:629-638) */
void gpu_advect_field_double(float *field, float *field_init, float *backward_x,
                             float *backward_y, float *backward_z, float *backward_xprev,
                             float *backward_yprev, float *backward_zprev, float h, int ni, int nj,
                             int nk, bool is_point, float blend_coeff);
/* replaces GPU_Advection.h:55-58 (def. GPU_kernel.cu:684-696) */
void gpu_accumulate_velocity(float *u_change, float *v_change, float *w_change, float *du_init,
                             float *dv_init, float *dw_init, float *forward_x, float *forward_y,
                             float *forward_z, float h, int ni, int nj, int nk, bool is_point,
                             float coeff);
/* replaces GPU_Advection.h:60-62 (def. GPU_kernel.cu:698-705) */
void gpu_accumulate_field(float *field_change, float *dfield_init, float *forward_x,
                          float *forward_y, float *forward_z, float h, int ni, int nj, int nk,
                          bool is_point, float coeff);
/* replaces GPU_Advection.h:64-67 (def. GPU_kernel.cu:707-716): per-cell squared distortion -> du */
void gpu_estimate_distortion(float *du, float *x_init, float *y_init, float *z_init, float *x_fwd,
                             float *y_fwd, float *z_fwd, float h, int ni, int nj, int nk);
/* replaces GPU_Advection.h:69 (def. GPU_kernel.cu:729-734): field1 += coeff*field2 */
void gpu_add(float *field1, float *field2, float coeff, int number);
/* replaces GPU_Advection.h:71-76 (def. GPU_kernel.cu:640-666).  Same aliasing contract as the
 * reference: u_src receives the time-0 error, du/dv/dw are OVERWRITTEN with the pre-correction
 * u/v/w, and u/v/w receive the compensated, extrema-clamped result. */
void gpu_compensate_velocity(float *u, float *v, float *w, float *du, float *dv, float *dw,
                             float *u_src, float *v_src, float *w_src, float *forward_x,
                             float *forward_y, float *forward_z, float *backward_x,
                             float *backward_y, float *backward_z, float h, int ni, int nj, int nk,
                             bool is_point);
/* replaces GPU_Advection.h:78-81 (def. GPU_kernel.cu:668-682) */
void gpu_compensate_field(float *u, float *du, float *u_src, float *forward_x, float *forward_y,
                          float *forward_z, float *backward_x, float *backward_y,
                          float *backward_z, float h, int ni, int nj, int nk, bool is_point);
/* replaces GPU_Advection.h:83-86 (def. GPU_kernel.cu:718-727) */
void gpu_semilag(float *field, float *field_src, float *u, float *v, float *w, int dim_x,
                 int dim_y, int dim_z, float h, int ni, int nj, int nk, float cfldt, float dt);
/* replaces GPU_Advection.h:97 (def. GPU_kernel.cu:885-890): out = field1 + coeff*field2 */
void gpu_add_field(float *out, float *field1, float *field2, float coeff, int number);

/* ---- source terms on either side of the advection path, same prototypes as
 * the reference; the Poisson solvers (gpu_projection_jacobi, gpu_conjugate_gradient,
 * gpu_multi_grid_conjugate_gradient) and the MacCormack clamp (gpu_clamp_extrema) are not provided. */
/* replaces GPU_Advection.h:88-90 (def. GPU_kernel.cu:782-802) */
void gpu_emit_smoke(float *u, float *v, float *w, float *rho, float *T, float h, int ni, int nj, int nk,
                    float centerX, float centerY, float centerZ, float radius, float density,
                    float temperature, float emiter);
/* replaces GPU_Advection.h:92-93 (def. GPU_kernel.cu:825-832).  `field` is the v-face array (ni x (nj+1) x nk),
 * density / temperature are cell-centred (ni x nj x nk) but -- as in the reference -- are indexed with the
 * v-face index, so the last ni*nk face indices point past their end: the reference reads whatever memory
 * follows, this library reads 0 there.  In-range arithmetic is the reference's, bit for bit. */
void gpu_add_buoyancy(float *field, float *density, float *temperature, int ni, int nj, int nk,
                      float alpha, float beta, float dt);
/* replaces GPU_Advection.h:95 (def. GPU_kernel.cu:855-876): `iter` Jacobi sweeps between the two
 * scratch buffers, then the same buffer the reference copies back (:875) is copied into `field` */
void gpu_diffuse_field(float *field, float *fieldTemp0, float *filedTemp1, int ni, int nj, int nk,
                       int iter, float coef);
/* replaces GPU_Advection.h:103 (def. GPU_kernel.cu:959-964): field = coeff1*field1 + coeff2*field2 */
void gpu_mad(float *field, float *field1, float *field2, float coeff1, float coeff2, int number);

/* replaces the host loops of getCFL (BimocqGPUSolver.cpp:348-373, BimocqSolver.cpp:1067-1118):
 * *host_out = max(|u|, |v|, |w|) over device arrays of nu, nv, nw floats (legacy default stream). */
int bmq_max_abs3(const float *u, long long nu, const float *v, long long nv, const float *w, long long nw,
                 float *host_out);

/* replaces GPU_Advection.h:101 (def. GPU_kernel.cu:892-950; MacCormack / Reflection schemes,
 * BimocqGPUSolver.cpp:232-338).  ni,nj,nk are the FIELD's dimensions, dim* its staggering, o* the
 * sample offset in cells.  Literal reproduction, units slip (:913-915) and scatter included; reads
 * past the end of `field` return 0, writes outside `fieldTemp` are dropped. */
void gpu_clamp_extrema(float *field, float *fieldTemp, float *u, float *v, float *w, int ni, int nj, int nk,
                       int dimx, int dimy, int dimz, float ox, float oy, float oz, float h, float dt);

/* ------------------------------------------------------------------ blocked host containers */
/* The reference's host fields are buffer3Df: 8x8x8 blocks, cell (i,j,k) at
 * ((K*bx*by + J*bx + I) << 9) + (kk << 6) + (jj << 3) + ii, bx = ceil(nx/8), padded to whole blocks
 * (include/fluid_buffer3D.h:56-75,173-189).  gpuMapper::copyHostToDevice / copyDeviceToHost
 * (GPU_Advection.h:249-299) relayout them on the HOST before / after every transfer; these do it on
 * the device.  bmq_blocked_elems = the physical element count Buffer3D allocates. */
long long bmq_blocked_elems(int nx, int ny, int nz);
int bmq_blocked_to_linear(const float *blocked_dev, float *linear_dev, int nx, int ny, int nz, void *cuda_stream);
/* padding cells of the blocked buffer are written as 0 (what Buffer3D::init leaves there) */
int bmq_linear_to_blocked(const float *linear_dev, float *blocked_dev, int nx, int ny, int nz, void *cuda_stream);

/* ------------------------------------------------------------------ pressure projection */
/* One multigrid level; identical layout to the reference's SCoarseLevelInfo (GPU_Advection.h:13-24),
 * so a reference-side `SCoarseLevelInfo levels[LEVEL_COUNT]` can be passed as is. */
typedef struct bmq_coarse_level {
    int ni, nj, nk;
    int number;          /* ni*nj*nk */
    double alpha, beta;  /* Jacobi: x <- (sum of six neighbours + alpha*b) * beta */
    double *b, *x, *r;   /* device, `number` doubles each */
} bmq_coarse_level;

/* replaces GPU_Advection.h:107-108 (def. GPU_kernel.cu:1784-1828; called from
 * BimocqGPUSolver::projection, BimocqGPUSolver.cpp:443-445): `iter` iterations of the reference's
 * multigrid-corrected CG on  lap p = halfrdx * div(u,v,w), then u,v,w -= halfrdx * grad p.
 * Outputs bit-identical to the reference: u, v, w, p, div, residual, dir, tempResult[0..2*iter+2]
 * (CG scalars) and tempResult[2000..2000+iter] (max residual).  temp0, temp1 and levels[].b/x/r are
 * scratch.  Prolongation on a level whose size is even reads one plane past levels[].x in the
 * reference (GPU_kernel.cu:1611-1622 with ci = (ni-1)/2); here those reads return 0. */
void gpu_multi_grid_conjugate_gradient(float *u, float *v, float *w, double *div, double *p, double *dir,
                                       double *residual, double *temp0, double *temp1, double *tempResult,
                                       bmq_coarse_level *levels, int levelNum, int iter, double halfrdx);

/* replaces GPU_Advection.h:105 (def. GPU_kernel.cu:1345-1419): fp32 CG on lap p = halfrdx*div(u,v,w)
 * starting from the caller's p, `iter` iterations, then u,v,w -= halfrdx*grad p.  dotResult: >= 4096
 * floats (CG scalars at [0..2*iter+2], max residual at [2000..2000+iter]).  The reference compiles
 * its call site out (BimocqGPUSolver.cpp:423-441) but exports the symbol. */
void gpu_conjugate_gradient(float *u, float *v, float *w, float *div, float *p, float *residual, float *dir,
                            float *dotResult, int ni, int nj, int nk, int iter, float halfrdx);
/* replaces GPU_Advection.h:99 (def. GPU_kernel.cu:1816-1886): `iter` fp32 Jacobi sweeps
 * p <- (sum of six neighbours + alpha*div)*beta between p and p_temp; like the reference, the
 * gradient is taken of the iterate BEFORE the last one (:1866-1884).  debugParam: >= 4096 floats. */
void gpu_projection_jacobi(float *u, float *v, float *w, float *div, float *p, float *p_temp, float *debugParam,
                           int ni, int nj, int nk, int iter, float halfrdx, float alpha, float beta);

/* Handle that owns the fp64 work buffers BimocqGPUSolver's constructor allocates
 * (BimocqGPUSolver.cpp:56-90): div, p, dir, residual, temp0, temp1, tempResult[4096] and `levels`
 * levels with n_{l+1} = (n_l - 1)/2, alpha = -1, beta = 1/6. */
typedef struct bmq_mgpcg bmq_mgpcg;
enum { BMQ_MG_DIV = 0, BMQ_MG_P, BMQ_MG_DIR, BMQ_MG_RESIDUAL, BMQ_MG_RESULT };
int  bmq_mgpcg_create(int ni, int nj, int nk, int levels, bmq_mgpcg **out);
void bmq_mgpcg_destroy(bmq_mgpcg *m);
int  bmq_mgpcg_set_stream(bmq_mgpcg *m, void *cuda_stream);
/* = BimocqGPUSolver::projection (iter = 50, halfrdx = 0.5 there); u, v, w are device face fields */
int  bmq_mgpcg_solve(bmq_mgpcg *m, float *u, float *v, float *w, int iter, double halfrdx);
int  bmq_mgpcg_buffer(bmq_mgpcg *m, int which, double **device_ptr, long long *count);
/* copies the level table (device pointers inside) to `out`; returns the number of levels */
int  bmq_mgpcg_levels(bmq_mgpcg *m, bmq_coarse_level *out, int capacity);

/* ------------------------------------------------------------------ handle API (3D) */
typedef struct bmq3d_solver bmq3d_solver;

/* Layout of the HOST buffers passed to bmq3d_upload / bmq3d_download / bmq3d_advect_host /
 * bmq3d_accumulate_host: dense x-fastest (default), or the reference's 8^3-blocked buffer3Df
 * (`buffer3Df::_data->getPtr()`, bmq_blocked_elems floats per field) -- then the raw blocked buffer
 * is transferred and relaid out on the device, replacing gpuMapper::copyHostToDevice /
 * copyDeviceToHost (GPU_Advection.h:249-299).  Full-domain handles only. */
enum { BMQ_LAYOUT_LINEAR = 0, BMQ_LAYOUT_BLOCKED8 = 1 };
int bmq3d_set_host_layout(bmq3d_solver *s, int layout);

/* Field identifiers for bmq3d_field_ptr / bmq3d_upload / bmq3d_download.  Velocity faces are
 * (ni+1)*nj*nk, ni*(nj+1)*nk, ni*nj*(nk+1); everything else ni*nj*nk.  Dense, x-fastest. */
enum {
    BMQ_F_U = 0, BMQ_F_V, BMQ_F_W, BMQ_F_RHO, BMQ_F_T,                 /* current fields          */
    BMQ_F_U_INIT, BMQ_F_V_INIT, BMQ_F_W_INIT, BMQ_F_RHO_INIT, BMQ_F_T_INIT,
    BMQ_F_U_PREV, BMQ_F_V_PREV, BMQ_F_W_PREV, BMQ_F_RHO_PREV, BMQ_F_T_PREV,
    BMQ_F_DU_EXT, BMQ_F_DV_EXT, BMQ_F_DW_EXT, BMQ_F_DRHO_EXT, BMQ_F_DT_EXT, /* change: forces/sources */
    BMQ_F_DU_PROJ, BMQ_F_DV_PROJ, BMQ_F_DW_PROJ,                       /* change: projection      */
    BMQ_F_VFWD_X, BMQ_F_VFWD_Y, BMQ_F_VFWD_Z,                          /* velocity mapper psi     */
    BMQ_F_VBWD_X, BMQ_F_VBWD_Y, BMQ_F_VBWD_Z,                          /* velocity mapper chi     */
    BMQ_F_VBWDP_X, BMQ_F_VBWDP_Y, BMQ_F_VBWDP_Z,                       /* velocity mapper chi_prev*/
    BMQ_F_SFWD_X, BMQ_F_SFWD_Y, BMQ_F_SFWD_Z,                          /* scalar mapper psi       */
    BMQ_F_SBWD_X, BMQ_F_SBWD_Y, BMQ_F_SBWD_Z,
    BMQ_F_SBWDP_X, BMQ_F_SBWDP_Y, BMQ_F_SBWDP_Z,
    BMQ_F_U_SEMI, BMQ_F_V_SEMI, BMQ_F_W_SEMI, BMQ_F_RHO_SEMI, BMQ_F_T_SEMI, /* semi-Lagrangian fallback */
    BMQ_F_COUNT
};

/* Per-step scalars the reference prints (BimocqSolver.cpp:95,170-173,215). */
typedef struct bmq3d_stats {
    float max_v;              /* getCFL(): max(1e-4, max|u|,|v|,|w|); h on frame 0 (BimocqSolver.cpp:94) */
    float cfldt;              /* h / max|vel|                                                   */
    int n_substeps;           /* DMC / RK3 sub-steps taken                                       */
    float vel_distortion;     /* estimateDistortion / (max_v*dt), velocity mapper                */
    float scalar_distortion;  /* same, scalar mapper                                             */
    int vel_reinit;           /* 1 if the velocity maps were reinitialised this step             */
    int scalar_reinit;
    int vel_reinit_count;     /* MapperBase::total_reinit_count                                  */
    int scalar_reinit_count;
    float max_disp_z;         /* max |map_z - z| over both mappers, in cells (halo sizing)       */
    float max_disp_z_vel;     /* the same per mapper: the velocity mapper is reinitialised at    */
    float max_disp_z_scalar;  /* least every 10 frames, the scalar mapper every 30               */
} bmq3d_stats;

/* Creates a solver for an ni x nj x nk grid of cell size h on the current CUDA device.
 * blend_coeff is MapperBase::blend_coeff (1 = one-level map, as every shipped scene uses).
 * Slab form: the solver owns global planes [k_own0, k_own1) of a global grid of nk planes and
 * stores planes [k_own0-halo, k_own1+halo) clipped to the domain.  k_own0=0,k_own1=nk,halo=0
 * is the single-GPU case (bmq3d_create). */
int bmq3d_create(int ni, int nj, int nk, float h, float blend_coeff, bmq3d_solver **out);
int bmq3d_create_slab(int ni, int nj, int nk, float h, float blend_coeff, int k_own0, int k_own1,
                      int halo, bmq3d_solver **out);
int bmq3d_destroy(bmq3d_solver *s);
/* Use `stream` (a cudaStream_t cast to void*) for all subsequent work; NULL = legacy default. */
int bmq3d_set_stream(bmq3d_solver *s, void *stream);
/* Device pointer of a field's first STORED plane and the global index of that plane. */
int bmq3d_field_ptr(bmq3d_solver *s, int field_id, float **dev_ptr, int *first_plane,
                    int *n_planes, int *nx, int *ny);
/* Host <-> device copies of a whole stored field (pinned or pageable host memory). */
int bmq3d_upload(bmq3d_solver *s, int field_id, const float *host);
int bmq3d_download(bmq3d_solver *s, int field_id, float *host);
/* BimocqSolver::velocityReinitialize / scalarReinitialize semantics for frame 0: init <- current,
 * prev <- init, maps <- identity, counters <- 0.  Call after uploading the initial fields. */
int bmq3d_reset(bmq3d_solver *s);

/* Phase A of BimocqSolver::advanceBimocq (BimocqSolver.cpp:90-126): getCFL, update both map
 * pairs, optional semi-Lagrangian fallback fields, advect + compensate (+ two-level blend)
 * velocity, density and temperature.  Reads and overwrites U,V,W,RHO,T. */
int bmq3d_advect(bmq3d_solver *s, int framenum, float dt, int with_semilag);
/* Phase B (BimocqSolver.cpp:164-229): distortion estimate, reinit decision, accumulate the
 * change fields D*_EXT (coeff 1) and D*_PROJ (coeff proj_coeff) into the init buffers,
 * reinitialise when triggered.  The caller fills the change fields between the phases. */
int bmq3d_accumulate(bmq3d_solver *s, int framenum, float dt);
int bmq3d_get_stats(bmq3d_solver *s, bmq3d_stats *out);

/* Optional per-stage timing with CUDA events recorded on the solver's stream (used by bench.py
 * for the live roofline number).  bmq3d_timing_read synchronises, writes the milliseconds and
 * the number of recorded spans per slot (arrays of BMQ_T_COUNT), and clears the record. */
enum { BMQ_T_MAXVEL = 0, BMQ_T_DMC, BMQ_T_FORWARD, BMQ_T_SEMILAG, BMQ_T_ADVECT_V, BMQ_T_ERROR_V,
       BMQ_T_APPLY_V, BMQ_T_BLEND_V, BMQ_T_ADVECT_S, BMQ_T_ERROR_S, BMQ_T_APPLY_S, BMQ_T_BLEND_S,
       BMQ_T_DISTORTION, BMQ_T_ACCUM_V, BMQ_T_ACCUM_S, BMQ_T_REINIT, BMQ_T_COUNT };
int bmq3d_timing_enable(bmq3d_solver *s, int on);
int bmq3d_timing_read(bmq3d_solver *s, float *ms_out, int *spans_out, int n_slots);
/* the same, plus per slot the idle time of the stream before that stage started (what it waited for) */
int bmq3d_timing_read_gaps(bmq3d_solver *s, float *ms_out, int *spans_out, float *gap_ms_out, int n_slots);
const char *bmq3d_timing_slot_name(int slot);

/* Whole step through HOST buffers (the reference's host-orchestrated solver keeps its fields on
 * the host, Mapping.cpp:7-236).  bmq3d_advect_host uploads u,v,w (phase A never reads the current
 * density / temperature, which are pure outputs of MapperBase::advectField), runs phase A and
 * downloads the advected u,v,w,rho,T into the arrays; transfers overlap the stages.  The caller then applies its forces and
 * its projection on the host and hands bmq3d_accumulate_host the velocity after the external
 * forces (`*_forced`), the final velocity after projection and the final scalars; the change
 * fields are formed on the device exactly as BimocqSolver.cpp:149-162 forms them
 * (d_ext = forced - advected, d_proj = final - forced, d_scalar = final - advected) and phase B
 * runs.  No pointer may be NULL; host arrays are dense, whole fields. */
int bmq3d_advect_host(bmq3d_solver *s, int framenum, float dt, float *u, float *v, float *w,
                      float *rho, float *T);
int bmq3d_accumulate_host(bmq3d_solver *s, int framenum, float dt, const float *u_forced,
                          const float *v_forced, const float *w_forced, const float *u_final,
                          const float *v_final, const float *w_final, const float *rho_final,
                          const float *T_final);

/* ---- fine-grained stages for the z-slab driver (one call = one kernel over owned planes).
 * The multi-GPU driver exchanges halos between these calls.  `which` = 0 velocity mapper,
 * 1 scalar mapper, 2 both. */
int bmq3d_stage_maxvel(bmq3d_solver *s, float *max_abs_out);          /* local max|u|,|v|,|w| */
int bmq3d_stage_set_cfl(bmq3d_solver *s, int framenum, float global_max_abs);
int bmq3d_stage_dmc_substep(bmq3d_solver *s, float substep);          /* both mappers, ping-pong */
int bmq3d_stage_forward(bmq3d_solver *s, float dt);
int bmq3d_stage_semilag(bmq3d_solver *s, float dt);
int bmq3d_stage_advect(bmq3d_solver *s, int which);                   /* f_adv = quad9[init o chi] */
int bmq3d_stage_error(bmq3d_solver *s, int which);                    /* e0 = quad9[f_adv o psi]-init */
int bmq3d_stage_apply(bmq3d_solver *s, int which);                    /* f = clamp(f_adv-.5 quad9[e0 o chi]) */
int bmq3d_stage_blend(bmq3d_solver *s, int which);                    /* two-level blend, if active */
int bmq3d_stage_distortion(bmq3d_solver *s, float *vel_d2, float *scalar_d2, float *max_disp_z);
/* same, with the z-displacement (cells) of each mapper's maps: halo widths are sized per mapper */
int bmq3d_stage_distortion2(bmq3d_solver *s, float *vel_d2, float *scalar_d2, float *disp_z_vel,
                            float *disp_z_scalar);
/* Slab handles: re-allocate every field with `new_halo` (> current) halo planes, keeping the stored
 * planes; device pointers change (re-export IPC handles).  New planes are zero until exchanged. */
int bmq3d_grow_halo(bmq3d_solver *s, int new_halo);
int bmq3d_stage_decide(bmq3d_solver *s, int framenum, float dt, float vel_d2, float scalar_d2);
int bmq3d_stage_accumulate(bmq3d_solver *s, int which);
int bmq3d_stage_reinit(bmq3d_solver *s, int which, int phase);        /* phase 0: rotate+identity; 1: post-accumulate */
/* Scratch fields the slab driver must also exchange (ids continue after BMQ_F_COUNT). */
enum { BMQ_F_U_ADV = 64, BMQ_F_V_ADV, BMQ_F_W_ADV, BMQ_F_RHO_ADV, BMQ_F_T_ADV,
       BMQ_F_U_ERR, BMQ_F_V_ERR, BMQ_F_W_ERR, BMQ_F_RHO_ERR, BMQ_F_T_ERR,
       /* the six DMC ping-pong buffers (they rotate with the backward maps) */
       BMQ_F_TMPMAP0 = 80, BMQ_F_TMPMAP1, BMQ_F_TMPMAP2, BMQ_F_TMPMAP3, BMQ_F_TMPMAP4, BMQ_F_TMPMAP5 };

/* ------------------------------------------------------------------ z-slab decomposition over the GPUs of one box
 * The reference has no multi-GPU path; this is the new domain decomposition of BASELINE.json's
 * north_star behind a C API, so that a C++ host (one process, or one thread, per GPU) can drive it.  Rank r owns
 * the global planes [r*nk/world, (r+1)*nk/world) and stores `halo` more on both sides; kernels work on global
 * indices, so every rank computes bit for bit what a single GPU computes.  Halos are PULLED out of their owners'
 * memory with stream-ordered peer copies over NVLink (CUDA IPC mappings; raw pointers inside one process);
 * widths follow the measured z-displacement of each mapper's maps (velocity mapper: reinitialised at least every
 * 10 frames, scalar mapper every 30) and may reach past the neighbouring slab.
 *
 * Set-up:   bmq3d_mg_create on every rank; fill the fields through bmq3d_mg_solver + bmq3d_upload / bmq3d_field_ptr;
 *           bmq3d_reset; bmq3d_mg_export -> all-gather the blobs with whatever the host has (MPI, NCCL, files) ->
 *           bmq3d_mg_connect(all blobs, rank order); bmq3d_mg_set_collectives.
 * Per step: bmq3d_mg_advect; the caller's forces / projection on its own planes (change fields D*_EXT, D*_PROJ);
 *           bmq3d_mg_accumulate.
 * BMQ_ERR_HALO from bmq3d_mg_advect = the allocated halo is narrower than stats.halo_needed (nothing has been
 *           modified): on EVERY rank (the decision derives from all-reduced numbers) bmq3d_mg_disconnect, a host
 *           barrier, bmq3d_mg_grow_halo, export / all-gather / connect again, call bmq3d_mg_advect again. */
typedef struct bmq3d_mg bmq3d_mg;
/* in-place max all-reduce of n floats over all ranks (blocking); 0 = success */
typedef int (*bmq_allreduce_max_fn)(float *vals, int n, void *ctx);
/* a barrier over all ranks ORDERED IN `cuda_stream`: work queued on that stream afterwards starts only when every
 * rank's work queued on its stream before its own call has completed (e.g. a one-element ncclAllReduce on that
 * stream).  Must not block the host on the device; 0 = success */
typedef int (*bmq_stream_barrier_fn)(void *cuda_stream, void *ctx);
typedef struct bmq3d_mg_stats {
    int halo_allocated, halo_needed, halo_vel, halo_scalar;   /* planes: allocated; widest needed; exchanged last step per mapper */
    int halo_grown;                                           /* number of bmq3d_mg_grow_halo calls */
    long long exchanges, bytes_exchanged;                     /* since creation */
    int signalling;                                           /* BMQ_MG_SIGNAL_* in effect since the last connect */
} bmq3d_mg_stats;
/* How the ranks synchronise a halo exchange and reduce their maxima.
 *   HOST       the two callbacks of bmq3d_mg_set_collectives (an all-rank stream barrier per exchange, a blocking
 *              all-reduce) and one cudaMemcpyAsync per halo segment;
 *   DEVICE     no callbacks: every rank owns a block of flags in device memory that its peers map; an exchange is
 *              ONE kernel that publishes the rank's arrival, waits for the ranks within halo reach and pulls all
 *              segments over NVLink with 128-bit loads; reductions go through per-rank mailboxes (csrc/mg_signal.h);
 *   DEVICE_CE  the flag barrier of DEVICE, copies by the copy engines (cudaMemcpyAsync);
 *   AUTO       (default) DEVICE when every peer is another process (one process per GPU), HOST when ranks share
 *              a process; the environment variable BMQ_MG_SIGNAL = 0 / 1 / 2 overrides AUTO.
 * Every rank must use the same mode; set it before bmq3d_mg_connect. */
enum { BMQ_MG_SIGNAL_AUTO = -1, BMQ_MG_SIGNAL_HOST = 0, BMQ_MG_SIGNAL_DEVICE = 1, BMQ_MG_SIGNAL_DEVICE_CE = 2 };
int bmq3d_mg_create(int ni, int nj, int nk, float h, float blend_coeff, int rank, int world, int halo, bmq3d_mg **out);
int bmq3d_mg_destroy(bmq3d_mg *m);
int bmq3d_mg_solver(bmq3d_mg *m, bmq3d_solver **out);        /* the rank's slab handle (owned by m) */
int bmq3d_mg_set_collectives(bmq3d_mg *m, bmq_allreduce_max_fn allreduce_max, bmq_stream_barrier_fn stream_barrier, void *ctx);
int bmq3d_mg_set_signalling(bmq3d_mg *m, int mode);          /* BMQ_MG_SIGNAL_*; before bmq3d_mg_connect */
int bmq3d_mg_export_size(bmq3d_mg *m, size_t *bytes);
int bmq3d_mg_export(bmq3d_mg *m, void *blob);
int bmq3d_mg_connect(bmq3d_mg *m, const void *all_blobs);    /* world blobs of bmq3d_mg_export_size bytes, in rank order */
int bmq3d_mg_disconnect(bmq3d_mg *m);
int bmq3d_mg_grow_halo(bmq3d_mg *m, int new_halo);
int bmq3d_mg_advect(bmq3d_mg *m, int framenum, float dt);
int bmq3d_mg_accumulate(bmq3d_mg *m, int framenum, float dt);
int bmq3d_mg_get_stats(bmq3d_mg *m, bmq3d_mg_stats *out);

/* ------------------------------------------------------------------ handle API (2D)
 * The 2D reference (src/bimocq2D) is CPU code with no device seam; this API is the seam behind
 * BimocqSolver2D::advanceBIMOCQ (bimocq2D/BimocqSolver2D.cpp:390-508), whose signature
 * advance(float dt, int frame) (BimocqSolver2D.h:149) stays as it is.
 * Fields are row-major a[i + ni*j] like Array2f (include/array2.h:93-103): u is (ni+1) x nj,
 * v is ni x (nj+1), everything else ni x nj. */
typedef struct bmq2d_solver bmq2d_solver;
enum {
    BMQ2_F_U = 0, BMQ2_F_V, BMQ2_F_RHO, BMQ2_F_T,                         /* BimocqSolver2D::u, v, rho, temperature */
    BMQ2_F_U_TEMP, BMQ2_F_V_TEMP,                                         /* u_temp, v_temp (un-averaged velocity)  */
    BMQ2_F_U_INIT, BMQ2_F_V_INIT, BMQ2_F_RHO_INIT, BMQ2_F_T_INIT,
    BMQ2_F_U_ORIG, BMQ2_F_V_ORIG, BMQ2_F_RHO_ORIG, BMQ2_F_T_ORIG,         /* u_origin, v_origin, rho_orig, T_orig   */
    BMQ2_F_DU, BMQ2_F_DV, BMQ2_F_DRHO, BMQ2_F_DT,
    BMQ2_F_DU_PREV, BMQ2_F_DV_PREV, BMQ2_F_DRHO_PREV, BMQ2_F_DT_PREV,
    BMQ2_F_DU_EXT, BMQ2_F_DV_EXT, BMQ2_F_DRHO_EXT, BMQ2_F_DT_EXT,         /* du_temp, dv_temp, drho_temp, dT_temp   */
    BMQ2_F_DU_PROJ, BMQ2_F_DV_PROJ,
    BMQ2_F_U_FORCED, BMQ2_F_V_FORCED,                                     /* velocity after the external forces     */
    BMQ2_F_FWD_X, BMQ2_F_FWD_Y, BMQ2_F_BWD_X, BMQ2_F_BWD_Y, BMQ2_F_BWDP_X, BMQ2_F_BWDP_Y,
    BMQ2_F_SFWD_X, BMQ2_F_SFWD_Y, BMQ2_F_SBWD_X, BMQ2_F_SBWD_Y, BMQ2_F_SBWDP_X, BMQ2_F_SBWDP_Y,
    BMQ2_F_MAP_TMPX, BMQ2_F_MAP_TMPY,
    BMQ2_F_U_PRESAVE, BMQ2_F_V_PRESAVE, BMQ2_F_U_SAVE, BMQ2_F_V_SAVE, BMQ2_F_RHO_SAVE, BMQ2_F_T_SAVE,
    BMQ2_F_U_SEMI, BMQ2_F_V_SEMI, BMQ2_F_RHO_SEMI, BMQ2_F_T_SEMI,
    BMQ2_F_U_SCRATCH, BMQ2_F_U_SCRATCH2, BMQ2_F_V_SCRATCH, BMQ2_F_V_SCRATCH2, BMQ2_F_C_SCRATCH, BMQ2_F_C_SCRATCH2,
    BMQ2_F_COUNT
};
typedef struct bmq2d_stats {
    float cfl;                /* _cfl = h / |maxVel()| used for the DMC sub-steps (BimocqSolver2D.cpp:53-56) */
    float max_vel_pre;        /* maxVel() the CFL was taken from                                             */
    int n_substeps;
    float max_vel;            /* maxVel() of the post-projection velocity (:457)                             */
    float vel_condition;      /* d_vel / (vel*dt), the printed "Velocity remapping condition" (:458)         */
    float scalar_condition;
    int vel_remap, scalar_remap;
    int last_remesh, last_scalar_remesh, total_remesh, total_scalar_remesh;
} bmq2d_stats;
int bmq2d_create(int ni, int nj, float h, float blend_coeff, bmq2d_solver **out);
int bmq2d_destroy(bmq2d_solver *s);
int bmq2d_reset(bmq2d_solver *s);                                    /* constructor state (:156-270) */
int bmq2d_set_levelset(bmq2d_solver *s, int on);                     /* advect_levelset (BimocqSolver2D.h:285) */
int bmq2d_set_counters(bmq2d_solver *s, int lastremeshing, int rho_lastremeshing);
int bmq2d_field_ptr(bmq2d_solver *s, int field_id, float **dev_ptr, int *fni, int *fnj);
int bmq2d_upload(bmq2d_solver *s, int field_id, const float *host);
int bmq2d_download(bmq2d_solver *s, int field_id, float *host);
/* advanceBIMOCQ lines 394-445: CFL, map updates, semi-Lagrangian fields, advect + correct */
int bmq2d_advect(bmq2d_solver *s, int frame, float dt);
/* lines 449-507: on entry U_FORCED/V_FORCED = velocity after forces, U,V,RHO,T = after projection */
int bmq2d_accumulate(bmq2d_solver *s, int frame, float dt);
int bmq2d_get_stats(bmq2d_solver *s, bmq2d_stats *out);
/* diagnostic: cells of the last step whose solveODE went past its first round, per work list (forward maps velocity /
 * scalar, semi-Lagrangian rho, T, u, v); synchronises */
int bmq2d_deferred_counts(bmq2d_solver *s, int *counts6);
/* the same per round: counts36[6 * list + (r - 1)] = cells that entered round r = 1..6 of solveODE */
int bmq2d_deferred_round_counts(bmq2d_solver *s, int *counts36);
int bmq2d_advect_host(bmq2d_solver *s, int frame, float dt, float *u, float *v, float *rho, float *T);
int bmq2d_accumulate_host(bmq2d_solver *s, int frame, float dt, const float *u_forced, const float *v_forced,
                          float *u_final, float *v_final, const float *rho_final, const float *T_final);
unsigned long long bmq2d_kernel_launch_count(bmq2d_solver *s);

#ifdef __cplusplus
}
#endif
#endif /* BIMOCQ_B200_H */
