// What does a misaligned gather cost per warp-level instruction on B200?  One warp reads 32 consecutive
// elements starting at a misaligned offset (the trilinear-gather pattern of the advection kernels):
//   (a) 2 x LDG.32 from two arrays            (two co-sampled fields, separate arrays)
//   (b) 1 x LDG.64 from an interleaved array  (the same two fields as float2)
//   (c) 4 x LDG.32 from four arrays, (d) 1 x LDG.128 interleaved float4
//   (e) LDS.32 x 2 from shared memory (staged tile), (f) LDS.64 x 1
// Reports SM cycles per warp-level "pair fetch".  Arrays are small enough to sit in L1/L2 (the kernels are
// LSU-issue bound, not DRAM bound).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(const float *a, const float *b, const float *c, const float *d, float *out, int iters, int span)
{
    __shared__ float s[4096];
    for (int e = threadIdx.x; e < 4096; e += blockDim.x) s[e] = a[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float acc = 0.f;
    int off = 1 + warp * 37;                       // misaligned start, different per warp
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {              // 8 independent fetches in flight per warp
            const int p = (off + lane + j * 1031) & (span - 1);
            if (MODE == 0) acc += __ldg(a + p) + __ldg(b + p);
            if (MODE == 1) { const float2 v = __ldg(reinterpret_cast<const float2 *>(a) + p); acc += v.x + v.y; }
            if (MODE == 2) acc += __ldg(a + p) + __ldg(b + p) + __ldg(c + p) + __ldg(d + p);
            if (MODE == 3) { const float4 v = __ldg(reinterpret_cast<const float4 *>(a) + p); acc += v.x + v.y + v.z + v.w; }
            if (MODE == 4) acc += s[p & 2047] + s[2048 + (p & 2047)];
            if (MODE == 5) { const float2 v = reinterpret_cast<const float2 *>(s)[p & 2047]; acc += v.x + v.y; }
        }
        off += 33 + (int)(acc * 1e-30f);          // data dependence keeps the loads in the loop
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main(int argc, char **argv)
{
    // default 1 << 16: 256 KB per array as float (float4 view = 1 MB): L2 resident, L1 misses.
    // 1 << 12 (16 KB per array, 64 KB as float4): everything hits L1 -- the regime of the gather kernels (L1 hit
    // rate 88-94 %), where a staged shared-memory tile would have to beat the LDG
    const int span = argc > 1 ? 1 << atoi(argv[1]) : 1 << 16;
    printf("span = %d floats per array\n", span);
    float *buf, *out;
    cudaMalloc(&buf, sizeof(float) * span * 4 * 4);
    cudaMemset(buf, 0, sizeof(float) * span * 4 * 4);
    cudaMalloc(&out, sizeof(float) * 148 * 8 * 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 1000, blocks = 148 * 4;      // 4 CTAs x 8 warps = 32 warps per SM
    const char *names[6] = {"2 x LDG.32 (two arrays)", "1 x LDG.64 (float2)", "4 x LDG.32 (four arrays)", "1 x LDG.128 (float4)",
                            "2 x LDS.32", "1 x LDS.64"};
    for (int mode = 0; mode < 6; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            const float *a = buf, *b = buf + span, *c = buf + 2 * span, *d = buf + 3 * span;
            switch (mode) {
            case 0: k<0><<<blocks, 256>>>(a, b, c, d, out, iters, span); break;
            case 1: k<1><<<blocks, 256>>>(a, b, c, d, out, iters, span); break;
            case 2: k<2><<<blocks, 256>>>(a, b, c, d, out, iters, span); break;
            case 3: k<3><<<blocks, 256>>>(a, b, c, d, out, iters, span); break;
            case 4: k<4><<<blocks, 256>>>(a, b, c, d, out, iters, span); break;
            case 5: k<5><<<blocks, 256>>>(a, b, c, d, out, iters, span); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        // per SM: 32 warps x iters warp-level fetches
        printf("%-28s %.3f ms = %.2f SM-cycles per warp-level fetch (1.965 GHz, 32 warps/SM)\n", names[mode], best,
               best * 1e-3 * 1.965e9 / (32.0 * iters * 8));
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
