// scs_api.cu -- context, error reporting and the one-shot host entry points.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "scs_internal.h"

namespace scs {
static thread_local char g_err[1024] = "";
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace scs

namespace scs {
void* ctx_scratch(scs_ctx* ctx, size_t bytes) {
    if (bytes > ctx->scratch_bytes) {
        if (ctx->scratch) cudaFree(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        const size_t want = (bytes + 4095) / 4096 * 4096 * 2;
        if (cudaMalloc(&ctx->scratch, want) != cudaSuccess) return nullptr;
        ctx->scratch_bytes = want;
    }
    return ctx->scratch;
}
}  // namespace scs

namespace scs {
__global__ void scs_upload_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, size_t bytes) {
    const size_t n16 = (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) ? bytes / 16 : 0;
    const size_t stride = size_t(gridDim.x) * blockDim.x, i0 = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    for (size_t i = i0; i < n16; i += stride) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    for (size_t i = n16 * 16 + i0; i < bytes; i += stride) dst[i] = src[i];
}

int upload_pinned(void* d_dst, const void* h_pinned, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return SCS_OK;
    void* dev_view = nullptr;
    if (cudaHostGetDevicePointer(&dev_view, const_cast<void*>(h_pinned), 0) != cudaSuccess) {
        cudaGetLastError();          // not page-locked after all: the copy engine
        SCS_CUDA(cudaMemcpyAsync(d_dst, h_pinned, bytes, cudaMemcpyHostToDevice, st));
        return SCS_OK;
    }
    const int threads = 256;
    const int blocks = int(std::min<size_t>(64, (bytes / 16 + threads - 1) / threads + 1));
    scs_upload_kernel<<<blocks, threads, 0, st>>>(static_cast<const unsigned char*>(dev_view), static_cast<unsigned char*>(d_dst), bytes);
    SCS_CUDA(cudaGetLastError());
    return SCS_OK;
}

int upload_small(scs_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return SCS_OK;
    // cudaHostAlloc / cudaFreeHost synchronise the whole device (measured: a slot that had to grow while another stream's
    // matrix copies were in flight stalled the enqueueing thread for 4 ms; slots allocated one by one put a 0.5 ms call into
    // each of the first dozen steps), so ALL slots are carved out of one arena allocated on first use: 32 x 256 KB, larger
    // than any upload of per-group constants seen so far.  An upload that does not fit gets its own allocation, once.
    constexpr size_t SLOT = size_t(256) << 10;
    constexpr int NSLOT = int(sizeof(ctx->pin) / sizeof(ctx->pin[0]));
    if (!ctx->pin_arena) {
        if (cudaHostAlloc(&ctx->pin_arena, SLOT * NSLOT, cudaHostAllocDefault) != cudaSuccess)
            return set_error(SCS_ERR_OOM, "page-locked staging allocation of %zu bytes failed", SLOT * NSLOT);
        for (int i = 0; i < NSLOT; ++i) {
            ctx->pin[i].ptr = static_cast<char*>(ctx->pin_arena) + size_t(i) * SLOT;
            ctx->pin[i].bytes = SLOT;
            ctx->pin[i].own = false;
            SCS_CUDA(cudaEventCreateWithFlags(&ctx->pin[i].ev, cudaEventDisableTiming));
        }
    }
    scs_ctx::PinnedSlot& s = ctx->pin[ctx->pin_next];
    ctx->pin_next = (ctx->pin_next + 1) % NSLOT;
    SCS_CUDA(cudaEventSynchronize(s.ev));        // the slot's previous upload has been read (never recorded: returns at once)
    if (bytes > s.bytes) {
        if (s.own) cudaFreeHost(s.ptr);
        s.ptr = nullptr;
        s.bytes = 0;
        const size_t want = (bytes + 65535) / 65536 * 65536;
        if (cudaHostAlloc(&s.ptr, want, cudaHostAllocDefault) != cudaSuccess) return set_error(SCS_ERR_OOM, "page-locked staging allocation of %zu bytes failed", want);
        s.bytes = want;
        s.own = true;
    }
    memcpy(s.ptr, h_src, bytes);
    const int rc = upload_pinned(d_dst, s.ptr, bytes, st);
    if (rc != SCS_OK) return rc;
    SCS_CUDA(cudaEventRecord(s.ev, st));
    return SCS_OK;
}
}  // namespace scs

using namespace scs;

extern "C" {

// Host -> device upload of a small page-locked buffer by a kernel instead of the copy engine (see scs_internal.h).
int scs_upload_async(void* d_dst, const void* h_pinned, int64_t bytes, void* stream) {
    if (!d_dst || !h_pinned || bytes < 0) return set_error(SCS_ERR_ARG, "bad argument");
    return upload_pinned(d_dst, h_pinned, size_t(bytes), reinterpret_cast<cudaStream_t>(stream));
}

const char* scs_version(void) { return "scsommerclock_b200 0.1 (sm_100a)"; }
const char* scs_last_error(void) { return g_err; }

int scs_ctx_create(int32_t device, scs_ctx** out) {
    if (!out) return set_error(SCS_ERR_ARG, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return set_error(SCS_ERR_DEVICE, "no CUDA device is visible (%s); the PT hot path has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return set_error(SCS_ERR_ARG, "device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    SCS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error(SCS_ERR_DEVICE, "device %d (%s) is sm_%d%d; this library contains sm_100a code only", device, prop.name, prop.major, prop.minor);
    SCS_CUDA(cudaSetDevice(device));
    scs_ctx* c = new scs_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->total_mem = prop.totalGlobalMem;
    c->l2_bytes = prop.l2CacheSize;
    *out = c;
    return SCS_OK;
}

int scs_ctx_destroy(scs_ctx* ctx) {
    if (ctx && ctx->scratch) cudaFree(ctx->scratch);
    if (ctx) {
        for (auto& s : ctx->pin) {
            if (s.ev) cudaEventDestroy(s.ev);
            if (s.own && s.ptr) cudaFreeHost(s.ptr);
        }
        if (ctx->pin_arena) cudaFreeHost(ctx->pin_arena);
    }
    delete ctx;
    return SCS_OK;
}

int scs_ctx_info(scs_ctx* ctx, int64_t* info, int32_t n) {
    if (!ctx || !info) return set_error(SCS_ERR_ARG, "null argument");
    const int64_t v[6] = {ctx->sm_count, ctx->cc_major, ctx->cc_minor, int64_t(ctx->total_mem >> 20), ctx->l2_bytes, ctx->device};
    for (int i = 0; i < n && i < 6; ++i) info[i] = v[i];
    return SCS_OK;
}

int scs_map_mutations(scs_ctx* ctx, const uint8_t* gt, int64_t pitch_bytes, const uint8_t* row_keep, int64_t n_snv,
                      int32_t n_cells, const uint32_t* clade_bits, int64_t pitch_words, int32_t n_rows, double FP,
                      double FN, double* soft_weights) {
    scs_placement* pl = nullptr;
    int rc = scs_placement_create(ctx, n_snv, n_cells, n_rows, &pl);
    if (rc != SCS_OK) return rc;
    rc = scs_placement_set_clades_host(pl, clade_bits, pitch_words, nullptr);      // first: the masks decide the algorithm
    if (rc == SCS_OK) rc = scs_placement_set_genotypes_host(pl, gt, pitch_bytes, row_keep, nullptr);
    if (rc == SCS_OK) rc = scs_placement_run_host(pl, FP, FN, soft_weights, nullptr);
    scs_placement_destroy(pl);
    return rc;
}

}  // extern "C"
