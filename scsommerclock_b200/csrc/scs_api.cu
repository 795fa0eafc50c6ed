// scs_api.cu -- context, error reporting and the one-shot host entry points.
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

using namespace scs;

extern "C" {

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
