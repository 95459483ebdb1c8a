// z-marching apply kernels (march3d.cuh); one translation unit per kernel family so that they build in parallel.
#include "march3d.cuh"

namespace bmq {
static inline int march_stag_id(Stag st) { return st.dx ? 1 : st.dy ? 2 : st.dz ? 3 : 0; }

cudaError_t launch_apply_march(cudaStream_t s, const Grid3 &g, KRange r, Stag st, int nf, float *const *out,
                               const float *const *fadv, const float *const *e0, const float *const chi[3])
{
    const Map3 m{chi[0], chi[1], chi[2]};
    if (nf == 1) {
        MarchArgs<1, 1> a{};
        a.out[0] = out[0]; a.src[0] = e0[0]; a.aux[0] = fadv[0];
        return launch_march<GM_APPLY, 1, 1>(s, g, r, march_stag_id(st), a, m);
    }
    MarchArgs<2, 2> a{};
    for (int f = 0; f < 2; ++f) { a.out[f] = out[f]; a.src[f] = e0[f]; a.aux[f] = fadv[f]; }
    return launch_march<GM_APPLY, 2, 1>(s, g, r, march_stag_id(st), a, m);
}

// the same in two kernels: the apply kernel without its clamp, then the shared-memory tiled clamp (clamp27.cu)
cudaError_t launch_apply_march_split(cudaStream_t s, const Grid3 &g, KRange r, Stag st, int nf, float *const *out,
                                     const float *const *fadv, const float *const *e0, const float *const chi[3])
{
    const Map3 m{chi[0], chi[1], chi[2]};
    cudaError_t e;
    if (nf == 1) {
        MarchArgs<1, 1> a{};
        a.out[0] = out[0]; a.src[0] = e0[0]; a.aux[0] = fadv[0];
        e = launch_march<GM_APPLY_NC, 1, 1>(s, g, r, march_stag_id(st), a, m);
    } else {
        MarchArgs<2, 2> a{};
        for (int f = 0; f < 2; ++f) { a.out[f] = out[f]; a.src[f] = e0[f]; a.aux[f] = fadv[f]; }
        e = launch_march<GM_APPLY_NC, 2, 1>(s, g, r, march_stag_id(st), a, m);
    }
    if (e != cudaSuccess) return e;
    return launch_clamp27(s, g.ni + st.dx, g.nj + st.dy, g.nk + st.dz, r, nf, fadv, out);
}

}  // namespace bmq
