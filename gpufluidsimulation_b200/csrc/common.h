// Error latch and small host utilities shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include "../../include/bimocq_b200.h"

namespace bmq {
// Records `msg` (printf-style) as the library's last error and returns `code`.
int set_error(int code, const char *fmt, ...);
// Latches a CUDA error (if any) with the call site; returns BMQ_OK or BMQ_ERR_CUDA.
int check_cuda(cudaError_t e, const char *what, const char *file, int line);
// True when a CUDA device is usable; latches BMQ_ERR_NODEVICE otherwise.
bool require_device();
}  // namespace bmq

#define BMQ_CK(call)                                                              \
    do {                                                                          \
        int _st = ::bmq::check_cuda((call), #call, __FILE__, __LINE__);           \
        if (_st != BMQ_OK) return _st;                                            \
    } while (0)
// same, for the legacy void-returning symbols: latch and bail out
#define BMQ_CKV(call)                                                             \
    do {                                                                          \
        if (::bmq::check_cuda((call), #call, __FILE__, __LINE__) != BMQ_OK) return; \
    } while (0)
