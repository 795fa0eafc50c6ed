// scs_internal.h -- shared host-side plumbing of the C-ABI library.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/scsommerclock_b200.h"

struct scs_ctx {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    size_t total_mem;
    int l2_bytes;
    // small device scratch reused by the reduction entry points (no cudaMalloc/cudaFree -- and so no
    // device-wide synchronisation -- on the per-chunk path of StreamingPlacement)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
};

namespace scs {
// Device scratch of at least `bytes` owned by the context (grown, never shrunk); nullptr on failure.
void* ctx_scratch(scs_ctx* ctx, size_t bytes);
}  // namespace scs

namespace scs {
// Records a printf-style message for scs_last_error() and returns `code`.
int set_error(int code, const char* fmt, ...);
}  // namespace scs

#define SCS_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return scs::set_error(SCS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
