// Fused optimizer step over flat parameter storage (reference training/training_loop.py:303-328):
//   g   = nan_to_num(grad * grad_scale, nan=0, +-inf -> +-1e5)          (misc.nan_to_num, :309)
//   m   = b1 * m + (1 - b1) * g          (skipped when b1 == 0: torch.optim.Adam(betas=(0, 0.99)), train.py:210)
//   v   = b2 * v + (1 - b2) * g * g
//   p  -= lr * (m / (1 - b1^t)) / (sqrt(v / (1 - b2^t)) + eps)
//   p16 = bf16(p)                         (tensor-core shadow refreshed in the same pass)
// and the generator EMA:  p_ema = p + beta * (p_ema - p)   (:320-328)
// One pass, 128-bit accesses, grid = multiple of the SM count.  Pure HBM traffic: 22 B/param for Adam with
// b1 = 0 (r p,g,v; w p,v,p16), 12 B/param for the EMA.
#include "common.cuh"
#include "runtime.h"
#include <algorithm>

namespace {
using namespace ld;

__device__ __forceinline__ float sanitize(float g) {
    if (g != g) return 0.f;
    return fminf(fmaxf(g, -1e5f), 1e5f);
}

__global__ void __launch_bounds__(256)
adam_flat_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                 uint2* __restrict__ p16, long n4, float lr, float b1, float b2, float eps, float bc1, float bc2, float grad_scale,
                 const float* __restrict__ hyper_dev) {
    if (hyper_dev) { lr = __ldg(hyper_dev); bc1 = __ldg(hyper_dev + 1); bc2 = __ldg(hyper_dev + 2); }
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 pv = p[i];
        const float4 gv4 = g[i];
        float4 vv = v[i];
        float gg[4] = {sanitize(gv4.x * grad_scale), sanitize(gv4.y * grad_scale), sanitize(gv4.z * grad_scale), sanitize(gv4.w * grad_scale)};
        float mm[4] = {gg[0], gg[1], gg[2], gg[3]};
        if (b1 != 0.f) {
            float4 mv = m[i];
            mm[0] = b1 * mv.x + (1.f - b1) * gg[0]; mm[1] = b1 * mv.y + (1.f - b1) * gg[1];
            mm[2] = b1 * mv.z + (1.f - b1) * gg[2]; mm[3] = b1 * mv.w + (1.f - b1) * gg[3];
            m[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        }
        float pp[4] = {pv.x, pv.y, pv.z, pv.w};
        float v4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            v4[k] = b2 * v4[k] + (1.f - b2) * gg[k] * gg[k];
            const float denom = sqrtf(v4[k]) / sqrtf(bc2) + eps;
            pp[k] -= (lr / bc1) * (mm[k] / denom);
        }
        p[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        v[i] = make_float4(v4[0], v4[1], v4[2], v4[3]);
        if (p16) { uint2 o; o.x = pack_bf16x2(pp[0], pp[1]); o.y = pack_bf16x2(pp[2], pp[3]); p16[i] = o; }
    }
}

__global__ void __launch_bounds__(256)
ema_flat_kernel(float4* __restrict__ p_ema, const float4* __restrict__ p, uint2* __restrict__ p16, long n4, float beta) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 a = p[i]; float4 e = p_ema[i];
        e.x = a.x + beta * (e.x - a.x); e.y = a.y + beta * (e.y - a.y);
        e.z = a.z + beta * (e.z - a.z); e.w = a.w + beta * (e.w - a.w);
        p_ema[i] = e;
        if (p16) { uint2 o; o.x = pack_bf16x2(e.x, e.y); o.y = pack_bf16x2(e.z, e.w); p16[i] = o; }
    }
}
}  // namespace

extern "C" {
// n must be a multiple of 4 (flat storage is padded); all pointers 16-byte aligned.
int ld_adam_flat(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1, float beta2,
                 float eps, int step, float grad_scale, const float* hyper_dev, void* stream) {
    LD_CHECK_ARG(p && g && v && n > 0 && n % 4 == 0 && step >= 1, "adam_flat: bad argument");
    LD_CHECK_ARG(beta1 == 0.f || m != nullptr, "adam_flat: beta1 != 0 needs the first-moment buffer");
    const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    const long n4 = n / 4;
    const int grid = (int)std::min<long>((n4 + 255) / 256, (long)ld::sm_count() * 8);
    adam_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, (uint2*)p_bf16,
                                                             n4, lr, beta1, beta2, eps, bc1, bc2, grad_scale, hyper_dev);
    ld::count_launch();
    LD_LAUNCH_CHECK("adam_flat");
    return 0;
}

int ld_ema_flat(float* p_ema, const float* p, void* p_ema_bf16, int64_t n, float beta, void* stream) {
    LD_CHECK_ARG(p_ema && p && n > 0 && n % 4 == 0, "ema_flat: bad argument");
    const long n4 = n / 4;
    const int grid = (int)std::min<long>((n4 + 255) / 256, (long)ld::sm_count() * 8);
    ema_flat_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)p_ema, (const float4*)p, (uint2*)p_ema_bf16, n4, beta);
    ld::count_launch();
    LD_LAUNCH_CHECK("ema_flat");
    return 0;
}
}
