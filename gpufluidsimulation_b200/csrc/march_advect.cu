// z-marching advect kernels (march3d.cuh); one translation unit per kernel family so that they build in parallel.
#include "march3d.cuh"

namespace bmq {
static inline int march_stag_id(Stag st) { return st.dx ? 1 : st.dy ? 2 : st.dz ? 3 : 0; }

cudaError_t launch_advect_march(cudaStream_t s, const Grid3 &g, KRange r, Stag st, int nf, float *const *out,
                                const float *const *init, const float *const chi[3])
{
    const Map3 m{chi[0], chi[1], chi[2]};
    if (nf == 1) {
        MarchArgs<1, 1> a{};
        a.out[0] = out[0]; a.src[0] = init[0];
        return launch_march<GM_ADVECT, 1, 1>(s, g, r, march_stag_id(st), a, m);
    }
    MarchArgs<2, 2> a{};
    for (int f = 0; f < 2; ++f) { a.out[f] = out[f]; a.src[f] = init[f]; }
    return launch_march<GM_ADVECT, 2, 1>(s, g, r, march_stag_id(st), a, m);
}

}  // namespace bmq
