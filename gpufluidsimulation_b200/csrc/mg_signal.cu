// Kernels of mg_signal.h: neighbour barrier + halo pull over peer memory, and a max all-reduce through
// per-rank mailboxes.  See the header for the protocol.
#include "mg_signal.h"

namespace bmq {

namespace {

constexpr unsigned long long MG_TIMEOUT_NS = 30ull * 1000ull * 1000ull * 1000ull;
constexpr int PULL_THREADS = 128;

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys(const float *p)
{
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long now_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// waits until *p has reached `want` (counters only grow); false after MG_TIMEOUT_NS
__device__ bool wait_reached(const unsigned *p, unsigned want)
{
    const unsigned long long t0 = now_ns();
    while ((int)(ld_acquire_sys(p) - want) < 0) {
        __nanosleep(100);
        if (now_ns() - t0 > MG_TIMEOUT_NS) return false;
    }
    return true;
}

// 128 threads and <= 32 registers: such a CTA fits beside six resident CTAs of the gather kernels
// (6 x 128 x 80 registers leave 4096), so a posted exchange starts at once instead of waiting for a slot.
__global__ void __launch_bounds__(PULL_THREADS, 16) k_mg_pull(MgSignal *mine, MgWaitList wait, unsigned epoch, MgSegList segs)
{
    if (epoch) {
        // every CTA publishes (the same value) and polls: whichever CTA is scheduled first opens the barrier
        if (threadIdx.x == 0) st_release_sys(&mine->arrive, epoch);
        if (threadIdx.x < wait.n && !wait_reached(&wait.sig[threadIdx.x]->arrive, epoch)) atomicExch(&mine->timed_out, 1u);
        __syncthreads();
    }
    const size_t tid = (size_t)blockIdx.x * PULL_THREADS + threadIdx.x, stride = (size_t)gridDim.x * PULL_THREADS;
    for (int q = 0; q < segs.n; ++q) {
        const MgSeg sg = segs.seg[q];
        if ((((size_t)sg.src | (size_t)sg.dst | (size_t)sg.bytes) & 15) == 0) {
            const float4 *src = static_cast<const float4 *>(sg.src);
            float4 *dst = static_cast<float4 *>(sg.dst);
            const size_t n = sg.bytes / 16;
            size_t i = tid;
            for (; i + 3 * stride < n; i += 4 * stride) {      // four independent 16-byte loads in flight per thread
                const float4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
                dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
            }
            for (; i < n; i += stride) dst[i] = src[i];
        } else {
            const float *src = static_cast<const float *>(sg.src);
            float *dst = static_cast<float *>(sg.dst);
            const size_t n = sg.bytes / 4;
            size_t i = tid;
            for (; i + 3 * stride < n; i += 4 * stride) {
                const float a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
                dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
            }
            for (; i < n; i += stride) dst[i] = src[i];
        }
    }
}

// One warp.  Slot parity: a rank can only start reduction q + 2 after every rank has finished reading
// reduction q (finishing q + 1 needs everybody's contribution to q + 1, which they write after reading q).
__global__ void k_mg_allreduce_max(MgSignal *mine, MgWaitList others, float4 v, const float *dev_vals, int n, unsigned seq, float *host_out)
{
    const int par = seq & 1, t = threadIdx.x;
    if (dev_vals) {     // the contribution is the result of a kernel earlier in this stream (no host round trip)
        v.x = dev_vals[0];
        if (n > 1) v.y = dev_vals[1];
        if (n > 2) v.z = dev_vals[2];
        if (n > 3) v.w = dev_vals[3];
    }
    if (t == 0) {
        mine->red[par][0] = v.x; mine->red[par][1] = v.y; mine->red[par][2] = v.z; mine->red[par][3] = v.w;
        __threadfence_system();
        st_release_sys(&mine->red_seq[par], seq);
    }
    float m[MG_RED_MAX] = {v.x, v.y, v.z, v.w};
    bool ok = true;
    if (t < others.n) {
        const MgSignal *p = others.sig[t];
        ok = wait_reached(&p->red_seq[par], seq);
#pragma unroll
        for (int q = 0; q < MG_RED_MAX; ++q) m[q] = ld_relaxed_sys(&p->red[par][q]);
    }
    if (!ok) atomicExch(&mine->timed_out, 1u);
#pragma unroll
    for (int q = 0; q < MG_RED_MAX; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m[q] = fmaxf(m[q], __shfl_xor_sync(0xffffffffu, m[q], o));
    const unsigned bad = __any_sync(0xffffffffu, !ok) ? 1u : 0u;
    if (t == 0) {
        for (int q = 0; q < MG_RED_MAX; ++q) host_out[q] = q < n ? m[q] : 0.f;
        host_out[MG_RED_MAX] = (bad || ld_acquire_sys(&mine->timed_out)) ? 1.f : 0.f;
    }
}

}  // namespace

cudaError_t launch_mg_pull(cudaStream_t s, MgSignal *mine, const MgWaitList &wait, unsigned epoch, const MgSegList &segs)
{
    // one CTA per SM: ~1.2 MB of loads in flight, enough to cover the NVLink round trip at full rate; a
    // barrier-only launch needs a single CTA
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = segs.n > 0 ? sms : 1;
    k_mg_pull<<<blocks, PULL_THREADS, 0, s>>>(mine, wait, epoch, segs);
    return cudaGetLastError();
}

cudaError_t launch_mg_allreduce_max(cudaStream_t s, MgSignal *mine, const MgWaitList &others, const float *vals, const float *dev_vals,
                                    int n, unsigned seq, float *host_out)
{
    if (n < 1 || n > MG_RED_MAX || (!vals && !dev_vals)) return cudaErrorInvalidValue;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!dev_vals) v = make_float4(vals[0], n > 1 ? vals[1] : 0.f, n > 2 ? vals[2] : 0.f, n > 3 ? vals[3] : 0.f);
    k_mg_allreduce_max<<<1, 32, 0, s>>>(mine, others, v, dev_vals, n, seq, host_out);
    return cudaGetLastError();
}

}  // namespace bmq
