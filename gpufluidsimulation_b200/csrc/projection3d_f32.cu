// The reference's two fp32 pressure solvers, gpu_conjugate_gradient (GPU_kernel.cu:1345-1419) and
// gpu_projection_jacobi (:1816-1886), with the same prototypes.  The reference compiles their call
// sites out (`#if 0`, BimocqGPUSolver.cpp:408-441) but still exports the symbols
// (GPU_Advection.h:99,105); with them libbimocq_b200.so covers every `extern "C"` function of
// GPU_Advection.h and GPU_kernel.cu need not be built at all.
//
// Same structure as the fp64 solver in projection3d.cu (warp-per-block dot partials with the
// reference's float rounding and its sharedMem[+3] slips, one warp per calc_sum chain, grid-wide
// max), in float.  Arithmetic expressions are the reference's own so that nvcc contracts them the
// same way.  One behavioural difference: the scratch the reference cudaMallocs per call and reads
// uninitialised on ring cells (dotResidual after calc_poisson, :1383) is zero-filled here.
#include <algorithm>

#include "common.h"

namespace bmq {
void count_launches(unsigned n);   // kernels3d.cu
}

namespace {

#define F3_IJK(fi, fj, fk)                                        \
    const int i = blockIdx.x * 32 + threadIdx.x;                  \
    const int j = blockIdx.y * 4 + threadIdx.y;                   \
    const int k = blockIdx.z;                                     \
    if (i >= (fi) || j >= (fj) || k >= (fk)) return;              \
    const int index = i + (fi) * (j + (fj) * k);

dim3 fblk() { return dim3(32, 4, 1); }
dim3 fgrd(int fi, int fj, int fk) { return dim3((fi + 31) / 32, (fj + 3) / 4, fk); }
int fblocks(size_t n) { return (int)std::max<size_t>(1, std::min<size_t>((n + 255) / 256, 148 * 16)); }

// calc_poisson_value<float>, GPU_kernel.cu:1047-1059
__device__ __forceinline__ float poisson_f(const float *x, int i, int j, int k, int ni, int nj)
{
    const float x_center = x[k * ni * nj + j * ni + i];
    const float x_left = x[k * ni * nj + j * ni + i - 1], x_right = x[k * ni * nj + j * ni + i + 1];
    const float x_front = x[k * ni * nj + (j - 1) * ni + i], x_back = x[k * ni * nj + (j + 1) * ni + i];
    const float x_down = x[(k - 1) * ni * nj + j * ni + i], x_up = x[(k + 1) * ni * nj + j * ni + i];
    return (x_left + x_right + x_front + x_back + x_down + x_up) - x_center * 6;
}

// divergence_kernel(float), :966-983
__global__ void __launch_bounds__(128)
kf_divergence(const float *__restrict__ u, const float *__restrict__ v, const float *__restrict__ w, float *div, int ni,
              int nj, int nk, float halfrdx)
{
    F3_IJK(ni, nj, nk)
    const float u_left = u[k * (ni + 1) * nj + j * (ni + 1) + i], u_right = u[k * (ni + 1) * nj + j * (ni + 1) + i + 1];
    const float v_front = v[k * ni * (nj + 1) + j * ni + i], v_back = v[k * ni * (nj + 1) + (j + 1) * ni + i];
    const float w_down = w[k * ni * nj + j * ni + i], w_up = w[(k + 1) * ni * nj + j * ni + i];
    div[index] = halfrdx * ((u_right - u_left) + (v_back - v_front) + (w_up - w_down));
}

// gradient_kernel(float p), :1023-1040
__global__ void __launch_bounds__(128)
kf_gradient(float *field, const float *__restrict__ p, int fi, int fj, int fk, int dimx, int dimy, int dimz, float halfrdx)
{
    F3_IJK(fi, fj, fk)
    const int pi = fi - dimx, pj = fj - dimy, pk = fk - dimz;
    if (!(i > 1 && i < pi && j > 1 && j < pj && k > 1 && k < pk)) return;
    const float p0 = p[k * pj * pi + j * pi + i];
    const float p1 = p[(k - dimz) * pj * pi + (j - dimy) * pi + i - dimx];
    field[index] -= halfrdx * (p0 - p1);
}

// update_residual_kernel(float), :1237-1248, with calc_max<float> (:1192-1222) folded in (see
// k_residual in projection3d.cu)
__global__ void __launch_bounds__(128)
kf_residual(float *r, const float *__restrict__ b, const float *__restrict__ x, int ni, int nj, int nk, unsigned *maxbits)
{
    const int i = blockIdx.x * 32 + threadIdx.x, j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.z;
    float val = 0.f;
    if (i < ni && j < nj && k < nk) {
        const int index = i + ni * (j + nj * k);
        if (i > 0 && i < ni - 1 && j > 0 && j < nj - 1 && k > 0 && k < nk - 1) {
            val = b[index] - poisson_f(x, i, j, k, ni, nj);
            r[index] = val;
        } else {
            val = r[index];
        }
    }
    float m = val > 0.f ? val : 0.f;
    for (int o = 16; o; o >>= 1) {
        const float other = __shfl_xor_sync(0xffffffffu, m, o);
        m = other > m ? other : m;
    }
    __shared__ float wm[4];
    if (threadIdx.x == 0) wm[threadIdx.y] = m;
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        m = fmaxf(fmaxf(wm[0], wm[1]), fmaxf(wm[2], wm[3]));
        if (m > 0.f && m > __uint_as_float(*(volatile unsigned *)maxbits)) atomicMax(maxbits, __float_as_uint(m));
    }
}

// calc_poisson_kernel(float), :1062-1072
__global__ void __launch_bounds__(128) kf_poisson(const float *__restrict__ x, float *out, int ni, int nj, int nk)
{
    F3_IJK(ni, nj, nk)
    if (!(i > 0 && i < ni - 1 && j > 0 && j < nj - 1 && k > 0 && k < nk - 1)) return;
    out[index] = poisson_f(x, i, j, k, ni, nj);
}

// jacobi_kernel, :1816-1833: one sweep, interior cells
__global__ void __launch_bounds__(128)
kf_jacobi(const float *__restrict__ p, const float *__restrict__ div, float *outP, int ni, int nj, int nk, float alpha, float beta)
{
    F3_IJK(ni, nj, nk)
    if (!(i > 0 && i < ni - 1 && j > 0 && j < nj - 1 && k > 0 && k < nk - 1)) return;
    const float p_left = p[k * nj * ni + j * ni + i - 1], p_right = p[k * nj * ni + j * ni + i + 1];
    const float p_front = p[k * nj * ni + (j - 1) * ni + i], p_back = p[k * nj * ni + (j + 1) * ni + i];
    const float p_down = p[(k - 1) * nj * ni + j * ni + i], p_up = p[(k + 1) * nj * ni + j * ni + i];
    outP[index] = (p_left + p_right + p_front + p_back + p_down + p_up + alpha * div[index]) * beta;
}

// dot_vector<float>, :1086-1119: one partial per 256 elements, one warp per partial
__global__ void __launch_bounds__(256)
kf_dot(const float *__restrict__ v0, const float *__restrict__ v1, float *out, int count, int nref)
{
    __shared__ float prod[8][272 + 16];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float *s = prod[wid];
    for (int blk = blockIdx.x * 8 + wid; blk < nref; blk += gridDim.x * 8) {
        const int base = blk * 256;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const int e = lane + 32 * m, idx = base + e;
            s[e + (e >> 4)] = idx < count ? __fmul_rn(v0[idx], v1[idx]) : 0.f;
        }
        __syncwarp();
        float sum0 = 0.f;
        if (lane < 16) {
            const float *g = s + 17 * lane;
            sum0 = g[0];
#pragma unroll
            for (int q = 1; q < 16; ++q) sum0 = __fadd_rn(sum0, g[q]);
        }
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float gq = __shfl_sync(0xffffffffu, sum0, q);
            const float term = (q & 3) == 3 ? s[q] : gq;   // sharedMem[+3], [+7], [+11], [+15]
            acc = q == 0 ? term : __fadd_rn(acc, term);
        }
        if (lane == 0) out[blk] = acc;
        __syncwarp();
    }
}

// calc_sum<float>, :1134-1178 (see k_sum_chains / k_sum_tree in projection3d.cu)
__global__ void __launch_bounds__(128) kf_sum_chains(const float *__restrict__ v, float *chain_sums, int count, int cpt)
{
    const int lane = threadIdx.x & 31, chain = blockIdx.x * 4 + (threadIdx.x >> 5);
    const long long beg = (long long)chain * cpt;
    float a = 0.f;
    float nxt = (lane < cpt && beg + lane < count) ? v[beg + lane] : 0.f;
    for (int q = 0; q < cpt; q += 32) {
        const float cur = nxt;
        const int qn = q + 32 + lane;
        nxt = (qn < cpt && beg + qn < count) ? v[beg + qn] : 0.f;
        const int lim = min(32, min(cpt - q, (int)max(0LL, min((long long)count - beg - q, 32LL))));
        for (int m = 0; m < lim; ++m) a = __fadd_rn(a, __shfl_sync(0xffffffffu, cur, m));
    }
    if (lane == 0) chain_sums[chain] = a;
}
__global__ void __launch_bounds__(32) kf_sum_tree(const float *__restrict__ chain_sums, float *out, int slot)
{
    const int t = threadIdx.x;
    float g = 0.f;
    if (t < 16) {
        g = chain_sums[t * 16];
#pragma unroll
        for (int m = 1; m < 16; ++m) g = __fadd_rn(g, chain_sums[t * 16 + m]);
    }
    float acc = 0.f;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const float gm = __shfl_sync(0xffffffffu, g, m);
        acc = m == 0 ? gm : __fadd_rn(acc, gm);
    }
    if (t == 0) out[slot] = acc;
}

// update_x_kernel(float) :1284-1290, update_dir_kernel(float) :1300-1307
__global__ void __launch_bounds__(256)
kf_update_x(float *x, const float *__restrict__ dir, const float *__restrict__ alpha, int count, int rIndex, int dIndex)
{
    for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < (size_t)count; index += (size_t)gridDim.x * blockDim.x)
        x[index] += dir[index] * alpha[rIndex] / alpha[dIndex];
}
__global__ void __launch_bounds__(256)
kf_update_dir(float *dir, const float *__restrict__ residual, const float *__restrict__ beta, int count, int rIndex, int rPlusIndex)
{
    for (size_t index = (size_t)blockIdx.x * blockDim.x + threadIdx.x; index < (size_t)count; index += (size_t)gridDim.x * blockDim.x)
        dir[index] = residual[index] + dir[index] * beta[rPlusIndex] / beta[rIndex];
}
__global__ void kf_store_max(float *res, int slot, const unsigned *maxbits) { res[slot] = __uint_as_float(*maxbits); }

struct F32 {
    float *scratch;   // device: [0] running max bits, [1..256] chain sums
    int status = BMQ_OK;
    void ck(cudaError_t e, const char *what)
    {
        if (status == BMQ_OK && e != cudaSuccess) status = bmq::check_cuda(e, what, __FILE__, __LINE__);
    }
    void post(const char *what, unsigned n = 1)
    {
        bmq::count_launches(n);
        ck(cudaGetLastError(), what);
    }
    // update_residual + calc_max -> res[max_slot]
    void residual_max(float *r, const float *b, const float *x, int ni, int nj, int nk, float *res, int max_slot)
    {
        ck(cudaMemsetAsync(scratch, 0, sizeof(float), 0), "memset max");
        kf_residual<<<fgrd(ni, nj, nk), fblk()>>>(r, b, x, ni, nj, nk, reinterpret_cast<unsigned *>(scratch));
        kf_store_max<<<1, 1>>>(res, max_slot, reinterpret_cast<const unsigned *>(scratch));
        post("kf_residual", 2);
    }
    void dot(const float *v0, const float *v1, float *partials, float *res, int slot, int number)
    {
        const int nref = (number + 255) / 256, cpt = (nref + 255) / 256;
        kf_dot<<<std::min((nref + 7) / 8, 148 * 8), 256>>>(v0, v1, partials, number, nref);
        kf_sum_chains<<<64, 128>>>(partials, scratch + 1, nref, cpt);
        kf_sum_tree<<<1, 32>>>(scratch + 1, res, slot);
        post("kf_dot", 3);
    }
    void gradient(float *u, float *v, float *w, const float *p, int ni, int nj, int nk, float halfrdx)
    {
        kf_gradient<<<fgrd(ni + 1, nj, nk), fblk()>>>(u, p, ni + 1, nj, nk, 1, 0, 0, halfrdx);
        kf_gradient<<<fgrd(ni, nj + 1, nk), fblk()>>>(v, p, ni, nj + 1, nk, 0, 1, 0, halfrdx);
        kf_gradient<<<fgrd(ni, nj, nk + 1), fblk()>>>(w, p, ni, nj, nk + 1, 0, 0, 1, halfrdx);
        post("kf_gradient", 3);
    }
};

float *f32_scratch()
{
    static float *s = nullptr;   // 257 floats, allocated once per process
    if (!s && cudaMalloc(&s, 257 * sizeof(float)) != cudaSuccess) s = nullptr;
    return s;
}

bool dims_ok(int ni, int nj, int nk, int iter) { return ni >= 3 && nj >= 3 && nk >= 3 && iter >= 0 && iter <= 1000 && (double)ni * nj * nk < 2.0e9; }

}  // namespace

extern "C" {

void gpu_conjugate_gradient(float *u, float *v, float *w, float *div, float *p, float *residual, float *dir, float *dotResult,
                            int ni, int nj, int nk, int iter, float halfrdx)
{
    if (!bmq::require_device()) return;
    if (!dims_ok(ni, nj, nk, iter)) { bmq::set_error(BMQ_ERR_ARG, "gpu_conjugate_gradient: bad dimensions / iteration count"); return; }
    F32 f{f32_scratch()};
    if (!f.scratch) { bmq::set_error(BMQ_ERR_CUDA, "gpu_conjugate_gradient: scratch allocation failed"); return; }
    const int number = ni * nj * nk;
    kf_divergence<<<fgrd(ni, nj, nk), fblk()>>>(u, v, w, div, ni, nj, nk, halfrdx);
    f.post("kf_divergence");
    // x starts from the caller's p (the reference does not clear it, :1353-1355)
    f.residual_max(residual, div, p, ni, nj, nk, dotResult, 2000);
    f.ck(cudaMemcpyAsync(dir, residual, sizeof(float) * number, cudaMemcpyDeviceToDevice, 0), "dir = r");   // mul_kernel(.., 1)
    float *dotResidual = nullptr, *dotDir = nullptr;
    BMQ_CKV(cudaMalloc(&dotResidual, sizeof(float) * number));
    if (bmq::check_cuda(cudaMalloc(&dotDir, sizeof(float) * number), "cudaMalloc", __FILE__, __LINE__) != BMQ_OK) { cudaFree(dotResidual); return; }
    f.ck(cudaMemsetAsync(dotResidual, 0, sizeof(float) * number, 0), "memset");
    f.ck(cudaMemsetAsync(dotDir, 0, sizeof(float) * number, 0), "memset");
    f.dot(residual, residual, dotResidual, dotResult, 0, number);
    for (int i = 0; i < iter && f.status == BMQ_OK; ++i) {
        kf_poisson<<<fgrd(ni, nj, nk), fblk()>>>(dir, dotResidual, ni, nj, nk);
        f.post("kf_poisson");
        f.dot(dir, dotResidual, dotDir, dotResult, i * 2 + 1, number);
        kf_update_x<<<fblocks(number), 256>>>(p, dir, dotResult, number, i * 2, i * 2 + 1);
        f.post("kf_update_x");
        f.residual_max(residual, div, p, ni, nj, nk, dotResult, 2001 + i);
        f.dot(residual, residual, dotResidual, dotResult, (i + 1) * 2, number);
        kf_update_dir<<<fblocks(number), 256>>>(dir, residual, dotResult, number, i * 2, (i + 1) * 2);
        f.post("kf_update_dir");
    }
    f.gradient(u, v, w, p, ni, nj, nk, halfrdx);
    cudaFree(dotResidual);
    cudaFree(dotDir);
}

void gpu_projection_jacobi(float *u, float *v, float *w, float *div, float *p, float *p_temp, float *debugParam, int ni, int nj,
                           int nk, int iter, float halfrdx, float alpha, float beta)
{
    if (!bmq::require_device()) return;
    if (!dims_ok(ni, nj, nk, iter)) { bmq::set_error(BMQ_ERR_ARG, "gpu_projection_jacobi: bad dimensions / iteration count"); return; }
    F32 f{f32_scratch()};
    if (!f.scratch) { bmq::set_error(BMQ_ERR_CUDA, "gpu_projection_jacobi: scratch allocation failed"); return; }
    const int number = ni * nj * nk;
    kf_divergence<<<fgrd(ni, nj, nk), fblk()>>>(u, v, w, div, ni, nj, nk, halfrdx);
    f.post("kf_divergence");
    float *residual = nullptr, *dotResidual = nullptr;
    BMQ_CKV(cudaMalloc(&residual, sizeof(float) * number));
    if (bmq::check_cuda(cudaMalloc(&dotResidual, sizeof(float) * number), "cudaMalloc", __FILE__, __LINE__) != BMQ_OK) { cudaFree(residual); return; }
    f.ck(cudaMemsetAsync(residual, 0, sizeof(float) * number, 0), "memset");
    // diagnostics exactly as the reference keeps them: debugParam[i] = "r.r", debugParam[2000+i] = max r
    f.residual_max(residual, div, p, ni, nj, nk, debugParam, 2000);
    f.dot(residual, residual, dotResidual, debugParam, 0, number);
    float *p_in = p, *p_out = p_temp;
    for (int i = 0; i < iter && f.status == BMQ_OK; ++i) {
        kf_jacobi<<<fgrd(ni, nj, nk), fblk()>>>(p_in, div, p_out, ni, nj, nk, alpha, beta);
        f.post("kf_jacobi");
        f.residual_max(residual, div, p_out, ni, nj, nk, debugParam, 2001 + i);
        f.dot(residual, residual, dotResidual, debugParam, i + 1, number);
        std::swap(p_in, p_out);
    }
    // :1866-1869 -- after the last swap p_out is the iterate BEFORE the last one; the reference copies
    // it over p when it sits in p_temp and takes the gradient of p_out either way
    if (p_out == p_temp) f.ck(cudaMemcpyAsync(p, p_temp, sizeof(float) * number, cudaMemcpyDeviceToDevice, 0), "p <- p_temp");
    f.gradient(u, v, w, p_out, ni, nj, nk, halfrdx);
    cudaFree(residual);
    cudaFree(dotResidual);
}

}  // extern "C"
