// Row-wise softmax cross-entropy with label smoothing and ignore_index, forward + gradient in one
// pass over the logits (reference: CrossEntropyLoss(label_smoothing=0.1) in training/med.py:917-918,
// F.cross_entropy in training/networks_detr.py:185,344 and training/loss.py:105,178,189).
//   loss_r = (1-eps) * (lse - x[y]) + eps * (lse - mean_c x[c])          (0 for ignored rows)
//   dlogits[r, c] = grad_scale * (softmax_c - (1-eps) * [c == y] - eps / V)   (0 for ignored rows)
// One CTA per row; the row stays in L1/L2 between the three sweeps.
#include "common.cuh"
#include "runtime.h"

namespace {
using namespace ld;
constexpr int CE_THREADS = 256;

template <typename T> __device__ __forceinline__ float ce_ld(const T* p, long i);
template <> __device__ __forceinline__ float ce_ld<float>(const float* p, long i) { return p[i]; }
template <> __device__ __forceinline__ float ce_ld<__nv_bfloat16>(const __nv_bfloat16* p, long i) { return bf16_to_f32(p[i]); }
template <typename T> __device__ __forceinline__ void ce_st(T* p, long i, float v);
template <> __device__ __forceinline__ void ce_st<float>(float* p, long i, float v) { p[i] = v; }
template <> __device__ __forceinline__ void ce_st<__nv_bfloat16>(__nv_bfloat16* p, long i, float v) { p[i] = f32_to_bf16(v); }

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = (lane < CE_THREADS / 32) ? sh[lane] : (is_max ? -INFINITY : 0.f);
    r = is_max ? warp_max(r) : warp_sum(r);
    return r;   // valid in every thread of every warp (each warp reduces the same 8 partials)
}

template <typename T, typename TG>
__global__ void __launch_bounds__(CE_THREADS)
ce_kernel(const T* __restrict__ logits, long ld_, const int64_t* __restrict__ labels, float* __restrict__ loss_rows,
          TG* __restrict__ dlogits, long ldg, int V, float eps, long ignore_index, float grad_scale,
          const float* __restrict__ grad_scale_dev) {
    __shared__ float sh[CE_THREADS / 32];
    if (grad_scale_dev) grad_scale *= __ldg(grad_scale_dev);
    const long r = blockIdx.x;
    const T* x = logits + r * ld_;
    const long y = labels[r];
    if (y == ignore_index) {
        if (threadIdx.x == 0 && loss_rows) loss_rows[r] = 0.f;
        if (dlogits) for (int c = threadIdx.x; c < V; c += CE_THREADS) ce_st<TG>(dlogits, r * ldg + c, 0.f);
        return;
    }
    float mx = -INFINITY;
    for (int c = threadIdx.x; c < V; c += CE_THREADS) mx = fmaxf(mx, ce_ld<T>(x, c));
    mx = block_reduce(mx, true, sh);
    float se = 0.f, sx = 0.f;
    for (int c = threadIdx.x; c < V; c += CE_THREADS) { const float v = ce_ld<T>(x, c); se += __expf(v - mx); sx += v; }
    se = block_reduce(se, false, sh);
    sx = block_reduce(sx, false, sh);
    const float lse = mx + __logf(se);
    if (threadIdx.x == 0 && loss_rows) {
        const float xy = ce_ld<T>(x, y);
        loss_rows[r] = (1.f - eps) * (lse - xy) + eps * (lse - sx / (float)V);
    }
    if (dlogits) {
        const float inv = 1.f / se, u = eps / (float)V;
        for (int c = threadIdx.x; c < V; c += CE_THREADS) {
            float g = __expf(ce_ld<T>(x, c) - mx) * inv - u;
            if (c == y) g -= (1.f - eps);
            ce_st<TG>(dlogits, r * ldg + c, g * grad_scale);
        }
    }
}
}  // namespace

extern "C" int ld_cross_entropy(const void* logits, int dtype, int64_t ld_, const int64_t* labels, float* loss_rows,
                                void* dlogits, int g_dtype, int64_t ldg, int64_t rows, int V, float label_smoothing,
                                int64_t ignore_index, float grad_scale, const float* grad_scale_dev, void* stream) {
    LD_CHECK_ARG(logits && labels && rows > 0 && V > 0, "cross_entropy: bad argument");
    LD_CHECK_ARG(loss_rows || dlogits, "cross_entropy: nothing to compute");
    cudaStream_t st = (cudaStream_t)stream;
#define CE(T, TG) ce_kernel<T, TG><<<(unsigned)rows, CE_THREADS, 0, st>>>((const T*)logits, ld_, labels, loss_rows, (TG*)dlogits, ldg, V, label_smoothing, ignore_index, grad_scale, grad_scale_dev)
    if (dtype == LD_F32 && g_dtype == LD_F32) CE(float, float);
    else if (dtype == LD_F32) CE(float, __nv_bfloat16);
    else if (g_dtype == LD_F32) CE(__nv_bfloat16, float);
    else CE(__nv_bfloat16, __nv_bfloat16);
#undef CE
    ld::count_launch();
    LD_LAUNCH_CHECK("cross_entropy");
    return 0;
}
