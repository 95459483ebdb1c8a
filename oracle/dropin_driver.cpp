// TEST INFRASTRUCTURE ONLY.  Literal drop-in proof: this driver uses the REFERENCE'S OWN
// gpuMapper (bimocq3D/GPU_Advection.h:110-627) and MapperBaseGPU (bimocq3D/Mapping.{h,cpp}),
// compiled unmodified, and is linked twice by oracle/Makefile:
//   _ref/dropin_ref   against the reference's kernels   (GPU_kernel.cu -> _ref/libref3d.so)
//   _ref/dropin_ours  against libbimocq_b200.so          (the 14 legacy extern "C" gpu_* symbols)
// Both run the same frames of mapper calls in BimocqGPUSolver::advanceBimocq's order
// (bimocq3D/BimocqGPUSolver.cpp:129-230) and dump the fields; tests/test_dropin_gpu.py compares
// the dumps byte for byte.  Nothing here is part of the product.
#include "Mapping.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

static float *dev_field(gpuMapper *g, size_t n, const std::vector<float> *init = nullptr)
{
    float *p = nullptr;
    // one plane + row of slack: the reference sampler reads one node past its clamp bound (weight 0)
    g->allocGPUBuffer((void **)&p, (n + n / 8 + 4096) * sizeof(float));
    if (init) cudaMemcpy(p, init->data(), n * sizeof(float), cudaMemcpyHostToDevice);
    return p;
}

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: %s ni nj nk L frames out.bin\n", argv[0]); return 2; }
    const int ni = atoi(argv[1]), nj = atoi(argv[2]), nk = atoi(argv[3]);
    const float L = (float)atof(argv[4]);
    const int frames = atoi(argv[5]);
    const float h = L / ni, dt = 0.02f * L, blend = 0.5f;
    const size_t nu = (size_t)(ni + 1) * nj * nk, nv = (size_t)ni * (nj + 1) * nk, nw = (size_t)ni * nj * (nk + 1), nc = (size_t)ni * nj * nk;

    // smooth analytic initial fields
    std::vector<float> hu(nu), hv(nv), hw(nw), hr(nc), hT(nc), hdu(nu), hdv(nv), hdw(nw), hdr(nc);
    auto wave = [&](float x, float y, float z, float a, float b, float c) {
        return std::sin(6.2831853f * (a * x / L + 0.13f)) * std::cos(6.2831853f * (b * y / (h * nj) + 0.29f)) * std::sin(6.2831853f * (c * z / (h * nk) + 0.41f));
    };
    const float vmax = 1.5f * h / dt;
    for (int k = 0; k < nk; ++k) for (int j = 0; j < nj; ++j) for (int i = 0; i <= ni; ++i)
        hu[i + (size_t)(ni + 1) * (j + (size_t)nj * k)] = vmax * wave((i - 0.5f) * h, j * h, k * h, 1, 2, 1);
    for (int k = 0; k < nk; ++k) for (int j = 0; j <= nj; ++j) for (int i = 0; i < ni; ++i)
        hv[i + (size_t)ni * (j + (size_t)(nj + 1) * k)] = vmax * wave(i * h, (j - 0.5f) * h, k * h, 2, 1, 1);
    for (int k = 0; k <= nk; ++k) for (int j = 0; j < nj; ++j) for (int i = 0; i < ni; ++i)
        hw[i + (size_t)ni * (j + (size_t)nj * k)] = vmax * wave(i * h, j * h, (k - 0.5f) * h, 1, 1, 2);
    for (int k = 0; k < nk; ++k) for (int j = 0; j < nj; ++j) for (int i = 0; i < ni; ++i) {
        const size_t q = i + (size_t)ni * (j + (size_t)nj * k);
        hr[q] = 0.5f + 0.5f * wave(i * h, j * h, k * h, 2, 2, 1);
        hT[q] = 0.5f + 0.5f * wave(i * h, j * h, k * h, 1, 3, 2);
        hdr[q] = 0.01f * wave(i * h, j * h, k * h, 3, 1, 1);
    }
    for (size_t q = 0; q < nu; ++q) hdu[q] = 0.01f * hu[q];
    for (size_t q = 0; q < nv; ++q) hdv[q] = -0.02f * hv[q];
    for (size_t q = 0; q < nw; ++q) hdw[q] = 0.015f * hw[q];

    gpuMapper *g = new gpuMapper(ni, nj, nk, h);
    MapperBaseGPU vel, sca;
    vel.init(ni, nj, nk, h, blend, g);
    sca.init(ni, nj, nk, h, blend, g);
    float *u = dev_field(g, nu, &hu), *v = dev_field(g, nv, &hv), *w = dev_field(g, nw, &hw);
    float *ui = dev_field(g, nu, &hu), *vi = dev_field(g, nv, &hv), *wi = dev_field(g, nw, &hw);
    float *up = dev_field(g, nu, &hu), *vp = dev_field(g, nv, &hv), *wp = dev_field(g, nw, &hw);
    float *rho = dev_field(g, nc, &hr), *rhoi = dev_field(g, nc, &hr), *rhop = dev_field(g, nc, &hr);
    float *T = dev_field(g, nc, &hT), *Ti = dev_field(g, nc, &hT), *Tp = dev_field(g, nc, &hT);
    float *du = dev_field(g, nu, &hdu), *dv = dev_field(g, nv, &hdv), *dw = dev_field(g, nw, &hdw), *dr = dev_field(g, nc, &hdr);

    float maxv = 1e-4f;
    for (float x : hu) maxv = std::fmax(maxv, std::fabs(x));
    for (float x : hv) maxv = std::fmax(maxv, std::fabs(x));
    for (float x : hw) maxv = std::fmax(maxv, std::fabs(x));
    const float cfldt = h / maxv;

    for (int frame = 0; frame < frames; ++frame) {
        vel.updateMapping(u, v, w, cfldt, dt);
        sca.updateMapping(u, v, w, cfldt, dt);
        vel.advectVelocity(u, v, w, ui, vi, wi, up, vp, wp);
        sca.advectField(rho, rhoi, rhop);
        sca.advectField(T, Ti, Tp);
        vel.accumulateVelocity(ui, vi, wi, du, dv, dw, 1.f);
        vel.accumulateVelocity(ui, vi, wi, du, dv, dw, 2.f);
        sca.accumulateField(rhoi, dr);
        if (frame % 3 == 2) {          // BimocqGPUSolver::velocityReinitialize / scalarReinitialize (:503-527)
            vel.reinitializeMapping();
            cudaMemcpy(up, ui, nu * sizeof(float), cudaMemcpyDeviceToDevice); cudaMemcpy(ui, u, nu * sizeof(float), cudaMemcpyDeviceToDevice);
            cudaMemcpy(vp, vi, nv * sizeof(float), cudaMemcpyDeviceToDevice); cudaMemcpy(vi, v, nv * sizeof(float), cudaMemcpyDeviceToDevice);
            cudaMemcpy(wp, wi, nw * sizeof(float), cudaMemcpyDeviceToDevice); cudaMemcpy(wi, w, nw * sizeof(float), cudaMemcpyDeviceToDevice);
            sca.reinitializeMapping();
            cudaMemcpy(rhop, rhoi, nc * sizeof(float), cudaMemcpyDeviceToDevice); cudaMemcpy(rhoi, rho, nc * sizeof(float), cudaMemcpyDeviceToDevice);
            cudaMemcpy(Tp, Ti, nc * sizeof(float), cudaMemcpyDeviceToDevice); cudaMemcpy(Ti, T, nc * sizeof(float), cudaMemcpyDeviceToDevice);
        }
    }
    cudaDeviceSynchronize();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }

    FILE *f = fopen(argv[6], "wb");
    if (!f) return 3;
    struct Out { float *p; size_t n; } outs[] = {{u, nu}, {v, nv}, {w, nw}, {rho, nc}, {T, nc}, {ui, nu}, {rhoi, nc},
                                                 {vel.ForwardX, nc}, {vel.BackwardZ, nc}, {sca.BackwardX, nc}, {sca.BackwardXPrev, nc}};
    std::vector<float> buf;
    for (auto &o : outs) {
        buf.resize(o.n);
        cudaMemcpy(buf.data(), o.p, o.n * sizeof(float), cudaMemcpyDeviceToHost);
        fwrite(buf.data(), sizeof(float), o.n, f);
    }
    fclose(f);
    printf("dropin driver: %d frames on %dx%dx%d done\n", frames, ni, nj, nk);
    return 0;
}
