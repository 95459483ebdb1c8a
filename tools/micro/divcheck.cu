// For which floats p does the three-instruction division by a constant h differ from IEEE division?
// Prints mismatch counts per binade of p.   usage: divcheck <n>  (h = 0.2f / n)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
__device__ __forceinline__ float div3(float p, float h, float y)
{
    const float q0 = __fmul_rn(p, y);
    const float r = __fmaf_rn(-h, q0, p);
    return __fmaf_rn(r, y, q0);
}
__global__ void k(float h, float y, unsigned last, unsigned *per_exp, unsigned *examples)
{
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b <= last; b += (unsigned long long)gridDim.x * blockDim.x) {
        const float p = __uint_as_float((unsigned)b);
        const float a = div3(p, h, y), c = __fdiv_rn(p, h);
        if (__float_as_uint(a) != __float_as_uint(c)) {
            const unsigned slot = atomicAdd(&per_exp[(unsigned)b >> 23], 1u);
            if (((unsigned)b >> 23) > 40 && slot < 4) examples[((unsigned)b >> 23) * 4 + slot] = (unsigned)b;
        }
    }
}
int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 37;
    const float h = 0.2f / (float)n, y = 1.0f / h;
    const float pmax = 4.0f * (n + 8) * h;
    unsigned last; memcpy(&last, &pmax, 4);
    unsigned *d, *e; cudaMalloc(&d, 256 * 4); cudaMemset(d, 0, 256 * 4); cudaMalloc(&e, 1024 * 4); cudaMemset(e, 0, 1024 * 4);
    k<<<148 * 8, 256>>>(h, y, last, d, e);
    unsigned hcnt[256], hex[1024]; cudaMemcpy(hcnt, d, sizeof hcnt, cudaMemcpyDeviceToHost); cudaMemcpy(hex, e, sizeof hex, cudaMemcpyDeviceToHost);
    printf("h = %.9g (0x%08x), y = %.9g, pmax = %g: %s\n", h, *(unsigned *)&h, y, pmax, cudaGetErrorString(cudaGetLastError()));
    for (int x = 0; x < 256; ++x) if (hcnt[x]) {
        printf("  biased exponent %3d (p ~ 2^%d): %u mismatches", x, x - 127, hcnt[x]);
        for (int s = 0; s < 4; ++s) if (hex[x * 4 + s]) { float p; memcpy(&p, &hex[x * 4 + s], 4); printf("  p=%.9g", p); }
        printf("\n");
    }
    return 0;
}
