// z-marching error kernels (march3d.cuh); one translation unit per kernel family so that they build in parallel.
#include "march3d.cuh"

namespace bmq {
static inline int march_stag_id(Stag st) { return st.dx ? 1 : st.dy ? 2 : st.dz ? 3 : 0; }

cudaError_t launch_error_march(cudaStream_t s, const Grid3 &g, KRange r, Stag st, int nf, float *const *e0,
                               const float *const *src, const float *const *init, const float *const psi[3])
{
    const Map3 m{psi[0], psi[1], psi[2]};
    if (nf == 1) {
        MarchArgs<1, 1> a{};
        a.out[0] = e0[0]; a.src[0] = src[0]; a.aux[0] = init[0];
        return launch_march<GM_ERROR, 1, 1>(s, g, r, march_stag_id(st), a, m);
    }
    MarchArgs<2, 2> a{};
    for (int f = 0; f < 2; ++f) { a.out[f] = e0[f]; a.src[f] = src[f]; a.aux[f] = init[f]; }
    return launch_march<GM_ERROR, 2, 1>(s, g, r, march_stag_id(st), a, m);
}

}  // namespace bmq
