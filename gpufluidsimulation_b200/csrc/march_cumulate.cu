// z-marching cumulate kernels (march3d.cuh); one translation unit per kernel family so that they build in parallel.
#include "march3d.cuh"

namespace bmq {
static inline int march_stag_id(Stag st) { return st.dx ? 1 : st.dy ? 2 : st.dz ? 3 : 0; }

// change is laid out [set][field]: change[c*nf + f]
cudaError_t launch_cumulate_march(cudaStream_t s, const Grid3 &g, KRange r, Stag st, int nf, int nch,
                                  float *const *target, const float *const *change, const float *coeff,
                                  const float *const map[3])
{
    const Map3 m{map[0], map[1], map[2]};
    if (nf == 1 && nch == 1) {
        MarchArgs<1, 1> a{};
        a.out[0] = target[0]; a.src[0] = change[0]; a.coeff[0] = coeff[0];
        return launch_march<GM_CUMULATE, 1, 1>(s, g, r, march_stag_id(st), a, m);
    }
    if (nf == 1 && nch == 2) {
        MarchArgs<1, 2> a{};
        a.out[0] = target[0];
        for (int c = 0; c < 2; ++c) { a.src[c] = change[c]; a.coeff[c] = coeff[c]; }
        return launch_march<GM_CUMULATE, 1, 2>(s, g, r, march_stag_id(st), a, m);
    }
    if (nf == 2 && nch == 1) {
        MarchArgs<2, 2> a{};
        for (int f = 0; f < 2; ++f) { a.out[f] = target[f]; a.src[f] = change[f]; }
        a.coeff[0] = coeff[0];
        return launch_march<GM_CUMULATE, 2, 1>(s, g, r, march_stag_id(st), a, m);
    }
    return cudaErrorInvalidValue;
}

}  // namespace bmq
