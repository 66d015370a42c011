// Library-wide runtime support: error strings, launch counter, device properties and the TMA
// tensor-map encoder (cuTensorMapEncodeTiled fetched through the runtime so we do not link libcuda).
#include "common.cuh"
#include "runtime.h"
#include "../../include/layoutdetr_sm100.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>

namespace ld {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_last_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return (int)e;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (cached[dev] == 0) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cached[dev] = n > 0 ? n : 148;
    }
    return cached[dev];
}

// Per-stream cap on the number of CTAs a persistent kernel may occupy (ld_set_stream_cta_limit): lets the host run
// a tensor-bound lane (frozen text encoder) next to latency-bound lanes without one kernel holding every SM.
static std::mutex g_lim_mu;
static struct { void* stream; int limit; } g_lims[32];
static int g_nlims = 0;

int cta_limit_for(void* stream) {
    const int sms = sm_count();
    std::lock_guard<std::mutex> lk(g_lim_mu);
    for (int i = 0; i < g_nlims; ++i)
        if (g_lims[i].stream == stream) return g_lims[i].limit < sms ? g_lims[i].limit : sms;
    return sms;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
    });
    return fn;
}

int encode_tmap_bf16_4d(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_bytes[3],
                        uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
    // cuTensorMapEncodeTiled is a driver entry point: it needs the primary context bound to THIS thread (autograd runs
    // backward on worker threads that may not have issued a runtime call yet).
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point unavailable (driver too old / no GPU)"); return LD_ERR_DRIVER; }
    cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t gstr[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t box[4] = {box_inner, box_outer, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed (CUresult %d): ptr=%p dims=[%llu,%llu,%llu,%llu] strides=[%llu,%llu,%llu] box=[%u,%u]",
                       (int)r, ptr, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                       (unsigned long long)dims[3], (unsigned long long)strides_bytes[0], (unsigned long long)strides_bytes[1],
                       (unsigned long long)strides_bytes[2], box_inner, box_outer);
        return LD_ERR_DRIVER;
    }
    return 0;
}

// General form: explicit box and traversal strides per dimension (implicit-GEMM convolution: a box of output pixels over an
// NHWC image, elementStrides = the convolution stride on W / H).
int encode_tmap_bf16_4d_box(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_bytes[3],
                            const uint32_t box[4], const uint32_t estr[4], int swizzle_bytes) {
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { set_last_error("cuTensorMapEncodeTiled entry point unavailable (driver too old / no GPU)"); return LD_ERR_DRIVER; }
    cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t gstr[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t e[4] = {estr[0], estr[1], estr[2], estr[3]};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, b, e,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed (CUresult %d): ptr=%p dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u] estr=[%u,%u,%u,%u]",
                       (int)r, ptr, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                       (unsigned long long)dims[3], box[0], box[1], box[2], box[3], estr[0], estr[1], estr[2], estr[3]);
        return LD_ERR_DRIVER;
    }
    return 0;
}

}  // namespace ld

extern "C" {
const char* ld_last_error(void) { return ld::g_err; }
int ld_version(void) { return 100; }
int64_t ld_launch_count(void) { return ld::g_launches.load(); }
void ld_launch_count_reset(void) { ld::g_launches.store(0); }

int ld_set_stream_cta_limit(void* stream, int limit) {
    std::lock_guard<std::mutex> lk(ld::g_lim_mu);
    int i = 0;
    for (; i < ld::g_nlims; ++i) if (ld::g_lims[i].stream == stream) break;
    if (limit <= 0) {                                   // clear
        if (i < ld::g_nlims) ld::g_lims[i] = ld::g_lims[--ld::g_nlims];
        return 0;
    }
    if (i == ld::g_nlims) {
        if (ld::g_nlims == 32) { ld::set_last_error("ld_set_stream_cta_limit: more than 32 limited streams"); return LD_ERR_INVALID_ARG; }
        ++ld::g_nlims;
    }
    ld::g_lims[i].stream = stream;
    ld::g_lims[i].limit = limit;
    return 0;
}

int ld_get_stream_cta_limit(void* stream) { return ld::cta_limit_for(stream); }
}
