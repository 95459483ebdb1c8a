// Error latch, version and launch counter of libbimocq_b200.so.
#include "common.h"
#include "launch3d.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>

namespace bmq {

static std::mutex g_err_mutex;
static std::string g_err_msg;
static int g_err_code = BMQ_OK;

int set_error(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    std::lock_guard<std::mutex> lk(g_err_mutex);
    g_err_msg = buf;
    g_err_code = code;
    return code;
}

int check_cuda(cudaError_t e, const char *what, const char *file, int line)
{
    if (e == cudaSuccess) return BMQ_OK;
    return set_error(BMQ_ERR_CUDA, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file,
                     line, what);
}

bool require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        set_error(BMQ_ERR_NODEVICE,
                  "libbimocq_b200: no CUDA device visible (%s); there is no CPU fallback",
                  e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return false;
    }
    return true;
}

}  // namespace bmq

extern "C" {

const char *bmq_last_error(void)
{
    // the returned pointer stays valid until the next error is latched
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(bmq::g_err_mutex);
    copy = bmq::g_err_msg;
    return copy.c_str();
}

int bmq_clear_error(void)
{
    std::lock_guard<std::mutex> lk(bmq::g_err_mutex);
    int c = bmq::g_err_code;
    bmq::g_err_code = BMQ_OK;
    bmq::g_err_msg.clear();
    return c;
}

int bmq_ipc_export(const void *dev_ptr, unsigned char handle[64])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    if (!dev_ptr || !handle) return bmq::set_error(BMQ_ERR_ARG, "bmq_ipc_export: null argument");
    cudaIpcMemHandle_t h;
    BMQ_CK(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    memcpy(handle, &h, 64);
    return BMQ_OK;
}

int bmq_ipc_open(const unsigned char handle[64], void **dev_ptr)
{
    if (!handle || !dev_ptr) return bmq::set_error(BMQ_ERR_ARG, "bmq_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    BMQ_CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return BMQ_OK;
}

int bmq_ipc_close(void *dev_ptr)
{
    if (!dev_ptr) return BMQ_OK;
    BMQ_CK(cudaIpcCloseMemHandle(dev_ptr));
    return BMQ_OK;
}

int bmq_copy_async(void *dst, const void *src, size_t bytes, void *stream)
{
    if (bytes == 0) return BMQ_OK;
    if (!dst || !src) return bmq::set_error(BMQ_ERR_ARG, "bmq_copy_async: null pointer");
    BMQ_CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return BMQ_OK;
}

const char *bmq_version(void) { return "bimocq_b200 0.1 (sm_100a)"; }

unsigned long long bmq_kernel_launch_count(void) { return bmq::kernel_launch_count(); }
int bmq_set_pitch_specialisation(int on) { bmq::set_pitch_specialisation(on != 0); return BMQ_OK; }
int bmq_set_gather_variant(int variant) { bmq::set_gather_variant(variant); return BMQ_OK; }
int bmq_set_fast_division(int on) { bmq::set_fast_division(on != 0); return BMQ_OK; }
int bmq_division_is_fast(float h, int nmax) { return bmq::division_is_fast(h, nmax) ? 1 : 0; }
int bmq_set_tolerance_mode(int on) { bmq::set_tolerance_mode(on != 0); return BMQ_OK; }
int bmq_tolerance_mode(void) { return bmq::tolerance_mode() ? 1 : 0; }

}  // extern "C"
