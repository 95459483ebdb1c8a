// NOT reference code. Stand-in for cuda-samples' helper_cuda.h, which the reference includes
// (bimocq3D/GPU_Advection.h:9) only for findCudaDevice (GPU_Advection.h:219).
#pragma once
#include <cuda_runtime.h>
static inline int findCudaDevice(int, const char **) { int d = -1; return cudaGetDevice(&d) == cudaSuccess ? d : -1; }
