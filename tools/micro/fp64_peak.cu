// Measures vector FP64 (DFMA / DADD) and shared-memory LDS.64 throughput of one GPU.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
            x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
        } else {
            x0 = __dadd_rn(x0, a); x1 = __dadd_rn(x1, a); x2 = __dadd_rn(x2, a); x3 = __dadd_rn(x3, a);
            x4 = __dadd_rn(x4, a); x5 = __dadd_rn(x5, a); x6 = __dadd_rn(x6, a); x7 = __dadd_rn(x7, a);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void klds(double *out, int iters)
{
    __shared__ double s[1024 + 64];
    s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    double acc = 0;
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int q = 0; q < 8; ++q) acc += s[(idx + q * 3) & 1023];
        idx = (idx + 1) & 1023;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main()
{
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 2, 1024>>>(out, iters, 1.0000001, 1e-9);
            else if (mode == 1) k<1><<<148 * 2, 1024>>>(out, iters, 1e-9, 0);
            else klds<<<148 * 2, 1024>>>(out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        double ops = 148.0 * 2 * 1024 * (double)iters * 8;
        if (mode == 0) printf("DFMA: %.1f G instr/s  = %.2f TFLOP/s  (%.1f thread-ops/clk/SM at 1.9 GHz)\n", ops / best / 1e6, 2 * ops / best / 1e9, ops / best / 1e6 / 148 / 1.9);
        if (mode == 1) printf("DADD: %.1f G instr/s  (%.1f thread-ops/clk/SM at 1.9 GHz)\n", ops / best / 1e6, ops / best / 1e6 / 148 / 1.9);
        if (mode == 2) printf("LDS.64: %.1f G loads/s = %.1f B/clk/SM at 1.9 GHz\n", ops / best / 1e6, ops * 8 / best / 1e6 / 148 / 1.9);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
