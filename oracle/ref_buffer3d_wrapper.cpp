// TEST INFRASTRUCTURE ONLY.  C wrapper around the REFERENCE'S OWN host container
// Buffer3D<float> (/root/reference/src/include/fluid_buffer3D.h, header-only), compiled next to this
// file by oracle/Makefile into oracle/_ref/libref_buffer3d.so.  Nothing from the reference is copied
// here: the wrapper only calls init() and operator()(i,j,k) and exposes the raw block storage.
#include <cstring>
typedef unsigned int uint;
#include <tbb/tbb.h>   // oracle/shim: the reference header uses tbb::parallel_for without including it
#include "fluid_buffer3D.h"

extern "C" {

long ref_b3d_physical_n(int nx, int ny, int nz)
{
    Buffer3D<float> b;
    b.init(nx, ny, nz, 1.0, 0.0, 0.0, 0.0);
    return (long)b._physical_n;
}

// blocked_out (ref_b3d_physical_n floats) = the container's storage after b(i,j,k) = linear[i + nx*(j + ny*k)]
void ref_b3d_from_linear(float *blocked_out, const float *linear, int nx, int ny, int nz)
{
    Buffer3D<float> b;
    b.init(nx, ny, nz, 1.0, 0.0, 0.0, 0.0);
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) b(i, j, k) = linear[i + nx * (j + ny * k)];
    std::memcpy(blocked_out, b._data->getPtr(), sizeof(float) * b._physical_n);
}

// linear_out[i + nx*(j + ny*k)] = b(i,j,k) for a container whose storage is `blocked`
void ref_b3d_to_linear(const float *blocked, float *linear_out, int nx, int ny, int nz)
{
    Buffer3D<float> b;
    b.init(nx, ny, nz, 1.0, 0.0, 0.0, 0.0);
    std::memcpy(b._data->getPtr(), blocked, sizeof(float) * b._physical_n);
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) linear_out[i + nx * (j + ny * k)] = b(i, j, k);
}

}  // extern "C"
