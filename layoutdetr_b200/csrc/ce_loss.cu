// Row-wise softmax cross-entropy with label smoothing and ignore_index, forward + gradient in one
// pass over the logits (reference: CrossEntropyLoss(label_smoothing=0.1) in training/med.py:917-918,
// F.cross_entropy in training/networks_detr.py:185,344 and training/loss.py:105,178,189).
//   loss_r = (1-eps) * (lse - x[y]) + eps * (lse - mean_c x[c])          (0 for ignored rows)
//   dlogits[r, c] = grad_scale * (softmax_c - (1-eps) * [c == y] - eps / V)   (0 for ignored rows)
// One CTA per row.  Generic kernel: the row stays in L1/L2 between three sweeps; bf16 rows of 1k..32k classes (the
// LM head) take the register-row kernel below: one 128-bit read + one 128-bit write per eight logits.
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

// bf16 logits -> bf16 gradient with the whole row held in registers: ONE 128-bit read and ONE 128-bit write per eight
// logits (the generic kernel above sweeps the row three times with 2-byte accesses: 2.7 TB/s on the LM-head shape).
// Row of up to 32768 classes = 512 threads x 8 groups x 8 values; V need not be a multiple of 8 but the row pitch must
// cover pad8(V) (the pad columns are read, ignored, and written as zeros).  In-place (dlogits == logits) is safe: a
// row is completely read before it is written.
constexpr int CEV_THREADS = 512;
constexpr int CEV_GROUPS = 8;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float block_reduce_v(float v, bool is_max, float* sh) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = (lane < CEV_THREADS / 32) ? sh[lane] : (is_max ? -INFINITY : 0.f);
    r = is_max ? warp_max(r) : warp_sum(r);
    return r;
}

__global__ void __launch_bounds__(CEV_THREADS, 2)
ce_row_regs_kernel(const __nv_bfloat16* logits, long ld_, const int64_t* __restrict__ labels, float* __restrict__ loss_rows,
                   __nv_bfloat16* dlogits, long ldg, int V, float eps, long ignore_index, float grad_scale,
                   const float* __restrict__ grad_scale_dev) {
    __shared__ float sh[CEV_THREADS / 32];
    if (grad_scale_dev) grad_scale *= __ldg(grad_scale_dev);
    const long r = blockIdx.x;
    const int G = (V + 7) >> 3;                                  // 8-value groups of this row
    const long y = labels[r];
    if (y == ignore_index) {
        if (threadIdx.x == 0 && loss_rows) loss_rows[r] = 0.f;
        if (dlogits) {
            uint4* d4 = reinterpret_cast<uint4*>(dlogits + r * ldg);
            for (int g = threadIdx.x; g < G; g += CEV_THREADS) d4[g] = make_uint4(0u, 0u, 0u, 0u);
        }
        return;
    }
    const uint4* x4 = reinterpret_cast<const uint4*>(logits + r * ld_);
    uint4 q[CEV_GROUPS];
#pragma unroll
    for (int j = 0; j < CEV_GROUPS; ++j) {
        const int g = j * CEV_THREADS + threadIdx.x;
        q[j] = (g < G) ? x4[g] : make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);    // bf16 -inf pairs
        if (g == G - 1 && (V & 7)) {                             // pad columns of the last group count as -inf
            uint32_t w[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = g * 8 + 2 * k;
                if (c >= V) w[k] = (w[k] & 0xFFFF0000u) | 0x0000FF80u;
                if (c + 1 >= V) w[k] = (w[k] & 0x0000FFFFu) | 0xFF800000u;
            }
            q[j] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    float mx = -INFINITY, sx = 0.f;
#pragma unroll
    for (int j = 0; j < CEV_GROUPS; ++j) {
        const uint32_t w[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float lo, hi; unpack_bf16x2(w[k], lo, hi);
            mx = fmaxf(mx, fmaxf(lo, hi));
            sx += (lo == -INFINITY ? 0.f : lo) + (hi == -INFINITY ? 0.f : hi);
        }
    }
#pragma unroll
    for (int j = 0; j < CEV_GROUPS; ++j)      // keep the row PACKED between sweeps (otherwise the compiler keeps 64 unpacked floats live)
        asm volatile("" : "+r"(q[j].x), "+r"(q[j].y), "+r"(q[j].z), "+r"(q[j].w));
    mx = block_reduce_v(mx, true, sh);
    constexpr float LOG2E = 1.4426950408889634f;
    const float mxl = mx * LOG2E;
    float se = 0.f;
#pragma unroll
    for (int j = 0; j < CEV_GROUPS; ++j) {
        const uint32_t w[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float lo, hi; unpack_bf16x2(w[k], lo, hi);
            se += ex2_approx(fmaf(lo, LOG2E, -mxl)) + ex2_approx(fmaf(hi, LOG2E, -mxl));       // 2^(-inf) = 0 for the pads
        }
    }
#pragma unroll
    for (int j = 0; j < CEV_GROUPS; ++j)      // keep the row PACKED between sweeps (otherwise the compiler keeps 64 unpacked floats live)
        asm volatile("" : "+r"(q[j].x), "+r"(q[j].y), "+r"(q[j].z), "+r"(q[j].w));
    se = block_reduce_v(se, false, sh);
    sx = block_reduce_v(sx, false, sh);
    const float lse = mx + __logf(se);
    if (threadIdx.x == 0 && loss_rows) {
        const float xy = bf16_to_f32(logits[r * ld_ + y]);
        loss_rows[r] = (1.f - eps) * (lse - xy) + eps * (lse - sx / (float)V);
    }
    if (dlogits) {
        __syncthreads();                                         // in-place: thread 0's read of x[y] precedes every store of the row
        const float inv = grad_scale / se, u = grad_scale * eps / (float)V, hit = grad_scale * (1.f - eps);
        uint4* d4 = reinterpret_cast<uint4*>(dlogits + r * ldg);
        const int gy = (int)(y >> 3), ky = (int)(y & 7);
#pragma unroll
        for (int j = 0; j < CEV_GROUPS; ++j) {
            const int g = j * CEV_THREADS + threadIdx.x;
            if (g < G) {
                const uint32_t w[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
                float o[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float lo, hi; unpack_bf16x2(w[k], lo, hi);
                    o[2 * k] = fmaf(ex2_approx(fmaf(lo, LOG2E, -mxl)), inv, -u);
                    o[2 * k + 1] = fmaf(ex2_approx(fmaf(hi, LOG2E, -mxl)), inv, -u);
                }
                if (g == gy) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (k == ky) o[k] -= hit;
                }
                if (g == G - 1 && (V & 7)) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) if (g * 8 + k >= V) o[k] = 0.f;
                }
                d4[g] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
            }
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
    {   // 128-bit register-row path (LM head: V = 30524, pitch 30528)
        auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
        const int64_t Vp = ((int64_t)V + 7) & ~(int64_t)7;
        if (dtype == LD_BF16 && (!dlogits || g_dtype == LD_BF16) && V >= 1024 && Vp <= (int64_t)CEV_THREADS * CEV_GROUPS * 8 &&
            ld_ % 8 == 0 && ld_ >= Vp && al16(logits) && (!dlogits || (ldg % 8 == 0 && ldg >= Vp && al16(dlogits)))) {
            ce_row_regs_kernel<<<(unsigned)rows, CEV_THREADS, 0, st>>>((const __nv_bfloat16*)logits, ld_, labels, loss_rows,
                                                                        (__nv_bfloat16*)dlogits, ldg, V, label_smoothing, ignore_index,
                                                                        grad_scale, grad_scale_dev);
            ld::count_launch();
            LD_LAUNCH_CHECK("cross_entropy");
            return 0;
        }
    }
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
