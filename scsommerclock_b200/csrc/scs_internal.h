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
    // ring of page-locked staging slots for small host -> device uploads that must not queue behind big copies
    // (scs::upload_small); each slot waits for its previous user before it is rewritten.  Not thread safe: a context is
    // driven by one host thread at a time (the sweep's parser / formatter threads never touch it)
    struct PinnedSlot {
        void* ptr = nullptr;
        size_t bytes = 0;
        cudaEvent_t ev = nullptr;
        bool own = false;           // its own allocation (an upload larger than an arena slot)
    };
    PinnedSlot pin[32];
    int pin_next = 0;
    void* pin_arena = nullptr;      // one page-locked allocation that holds all slots
};

namespace scs {
// Device scratch of at least `bytes` owned by the context (grown, never shrunk); nullptr on failure.
void* ctx_scratch(scs_ctx* ctx, size_t bytes);
}  // namespace scs

namespace scs {
// Small host -> device uploads as a KERNEL that reads page-locked host memory over PCIe instead of a copy-engine transfer:
// the one H2D queue of the device serves all streams, so a 1 MB upload issued while a 100 MB matrix copy of another stream
// is in flight waits for it (measured: the small uploads of a replicate sub-batch took 3.8 ms behind the next sub-batch's
// matrix).  upload_pinned: `h_pinned` is page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) and stays
// unchanged until the kernel has run; upload_small: any host memory, staged through the context's ring of pinned slots
// before the call returns.  Both asynchronous on `st`.
int upload_pinned(void* d_dst, const void* h_pinned, size_t bytes, cudaStream_t st);
int upload_small(scs_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st);
}  // namespace scs

namespace scs {
// Records a printf-style message for scs_last_error() and returns `code`.
int set_error(int code, const char* fmt, ...);
}  // namespace scs

namespace scs {
// Row filter + per-cell missing counts (scs_decode.cu) as an asynchronous launch on caller-owned counters plus a
// fetch: what scs_gt_reduce does in one call, and what scs_placement_prepare falls back to on the GEMM path.
size_t gt_reduce_scratch_bytes(int n_cells);
int gt_reduce_launch(scs_ctx* ctx, const uint8_t* d_gt, int64_t pitch_bytes, int64_t n_snv, int n_cells, const uint8_t* col_active,
                     int filter_rows, uint8_t* d_row_keep, uint8_t* scratch, bool accumulate, cudaStream_t st);
int gt_reduce_fetch(const uint8_t* scratch, int n_cells, int64_t* n_kept, double* missing_frac, cudaStream_t st);
int gt_reduce_groups_launch(scs_ctx* ctx, const uint8_t* d_gt, int64_t pitch_bytes, int n_groups, const long long* d_group_row0,
                            int64_t max_rows, int n_cells, const uint32_t* d_active_mask, int filter_rows, uint8_t* d_row_keep,
                            unsigned long long* d_n_kept, unsigned int* d_missing, cudaStream_t st);
}  // namespace scs

#define SCS_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return scs::set_error(SCS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)
