// Do SHFL and LDS share one pipe on B200?  Times per iteration of (a) 4 LDS.64, (b) 2 LDS.64 + 4 SHFL.32,
// (c) 4 SHFL.32, (d) 2 LDS.64, with 32 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(1024) k(double *out, int iters)
{
    __shared__ double s[34 * 34];
    for (int e = threadIdx.x; e < 34 * 34; e += 1024) s[e] = e;
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const double *my = s + (ty + 1) * 34 + tx + 1;
    double acc = 0, v = tx;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) acc += my[-1] + my[1] + my[-34] + my[34];
        if (MODE == 1) {
            acc += my[-34] + my[34];
            acc += __shfl_up_sync(0xffffffffu, v, 1) + __shfl_down_sync(0xffffffffu, v, 1);
        }
        if (MODE == 2) acc += __shfl_up_sync(0xffffffffu, v, 1) + __shfl_down_sync(0xffffffffu, v, 1);
        if (MODE == 3) acc += my[-34] + my[34];
        v += acc * 1e-30;
        my = s + (ty + 1) * 34 + ((tx + i) & 31) + 1;
    }
    out[blockIdx.x * 1024 + threadIdx.x] = acc + v;
}
int main()
{
    double *out; cudaMalloc(&out, sizeof(double) * 148 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    const char *names[4] = {"4 LDS.64", "2 LDS.64 + 2 SHFL.64", "2 SHFL.64", "2 LDS.64"};
    for (int mode = 0; mode < 4; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, 1024>>>(out, iters);
            if (mode == 1) k<1><<<148, 1024>>>(out, iters);
            if (mode == 2) k<2><<<148, 1024>>>(out, iters);
            if (mode == 3) k<3><<<148, 1024>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-24s %.3f ms  = %.1f cycles per iteration per SM (32 warps) at 1.9 GHz\n", names[mode], best, best * 1e-3 * 1.9e9 / iters);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
