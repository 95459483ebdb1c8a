// The 27-neighbour extrema clamp of the BFECC correction (clampExtrema_kernel, GPU_kernel.cu:146-167;
// Mapping.cpp:375-407) as its own kernel: a regular 3 x 3 x 3 min / max stencil, the one place on the advection path
// where a shared-memory tile pays (the gathers do not).  A CTA owns a 32 x 8 tile of columns
// and walks it in z: every plane of `before` is staged once in shared memory (34 x 10 values, double buffered, one
// barrier per plane), each thread takes the 3 x 3 extrema of its cell from the tile and carries the extrema of the
// two planes below in registers.  Per cell: ~1.3 global loads of `before`, one of `f`, nine shared-memory loads --
// against nine global loads per plane inside the apply kernel, which is bound by instruction issue.
// Result bits are those of the fused form (min / max are exact and order-free).
#include "launch3d.h"
#include "common.h"

namespace bmq {

namespace {

constexpr int TX = 32, TY = 8, SX = TX + 2, SY = TY + 2;

struct Ext { float mn, mx; };

template <int NF>
__global__ void __launch_bounds__(TX * TY)
k_clamp27(FieldSetRO<NF> before, FieldSetRW<NF> f, int fi, int fj, int fk, int kbeg, int kend, int kchunk)
{
    __shared__ float tile[2][NF][SY][SX];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int kc0 = kbeg + blockIdx.z * kchunk, kc1 = min(kc0 + kchunk, kend);
    const size_t plane = (size_t)fi * fj;
    const bool inside = i < fi && j < fj;
    const bool interior = i > 0 && i < fi - 1 && j > 0 && j < fj - 1;
    const size_t col = (size_t)i + (size_t)fi * j;

    // Staging: every thread brings its own cell, the first 84 threads one halo entry each (top and bottom row of 34,
    // left and right column of 8); which one is fixed per thread, so the per-plane work is one or two loads and
    // stores with no index arithmetic.  Entries outside the field are never read by an interior cell.
    int hx = -1, hy = -1;
    if (tid < SX) { hx = tid; hy = 0; }
    else if (tid < 2 * SX) { hx = tid - SX; hy = SY - 1; }
    else if (tid < 2 * SX + TY) { hx = 0; hy = tid - 2 * SX + 1; }
    else if (tid < 2 * SX + 2 * TY) { hx = SX - 1; hy = tid - 2 * SX - TY + 1; }
    const int gi = i0 - 1 + hx, gj = j0 - 1 + hy;
    const bool halo_ok = hx >= 0 && gi >= 0 && gi < fi && gj >= 0 && gj < fj;
    const size_t hcol = halo_ok ? (size_t)gi + (size_t)fi * gj : 0;
    auto stage = [&](int p, int buf) {
        const bool pok = p >= 0 && p < fk;
#pragma unroll
        for (int q = 0; q < NF; ++q) {
            tile[buf][q][ty + 1][tx + 1] = (inside && pok) ? __ldg(before.p[q] + col + plane * p) : 0.f;
            if (hx >= 0) tile[buf][q][hy][hx] = (halo_ok && pok) ? __ldg(before.p[q] + hcol + plane * p) : 0.f;
        }
    };
    auto extrema = [&](int buf, int q) {
        Ext e;
        e.mn = e.mx = tile[buf][q][ty + 1][tx + 1];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float v = tile[buf][q][ty + dy][tx + dx];
                e.mn = fminf(e.mn, v);
                e.mx = fmaxf(e.mx, v);
            }
        return e;
    };

    Ext lo[NF], mid[NF];          // extrema of planes k-1 and k
    stage(kc0 - 1, 0);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NF; ++q) lo[q] = extrema(0, q);
    stage(kc0, 1);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NF; ++q) mid[q] = extrema(1, q);
#pragma unroll 1
    for (int k = kc0; k < kc1; ++k) {
        const int buf = (k - kc0) & 1;          // plane k + 1 goes where plane k - 1 was
        stage(k + 1, buf);
        __syncthreads();
        const bool clamps = inside && interior && k > 0 && k < fk - 1;
#pragma unroll
        for (int q = 0; q < NF; ++q) {
            const Ext hi = extrema(buf, q);
            if (clamps) {
                const float mx = fmaxf(fmaxf(lo[q].mx, mid[q].mx), hi.mx);
                const float mn = fminf(fminf(lo[q].mn, mid[q].mn), hi.mn);
                float *pf = f.p[q] + col + plane * k;
                *pf = fminf(fmaxf(mn, *pf), mx);
            }
            lo[q] = mid[q];
            mid[q] = hi;
        }
    }
}

}  // namespace

// f <- clamp(f, min27(before), max27(before)) on planes [r.kbeg, r.kend) of nf fields of extents fi x fj x fk
cudaError_t launch_clamp27(cudaStream_t s, int fi, int fj, int fk, KRange r, int nf, const float *const *before, float *const *f)
{
    if (r.kend <= r.kbeg) return cudaSuccess;
    const int planes = r.kend - r.kbeg;
    const long long cols = (long long)((fi + TX - 1) / TX) * ((fj + TY - 1) / TY);
    int kc = 64;
    while (kc > 8 && cols * ((planes + kc - 1) / kc) < 148ll * 8 * 2) kc /= 2;
    const dim3 gr((fi + TX - 1) / TX, (fj + TY - 1) / TY, (planes + kc - 1) / kc), bl(TX, TY, 1);
    if (nf == 1) {
        FieldSetRO<1> b{{before[0]}};
        FieldSetRW<1> o{{f[0]}};
        k_clamp27<1><<<gr, bl, 0, s>>>(b, o, fi, fj, fk, r.kbeg, r.kend, kc);
    } else if (nf == 2) {
        FieldSetRO<2> b{{before[0], before[1]}};
        FieldSetRW<2> o{{f[0], f[1]}};
        k_clamp27<2><<<gr, bl, 0, s>>>(b, o, fi, fj, fk, r.kbeg, r.kend, kc);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace bmq
