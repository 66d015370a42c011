// Dropout support shared by the training-mode kernels (reference: nn.Dropout sites of training/med.py:96,213,240,318 and
// training/detr_transformer.py:185-194,210-214,270-285; the reference trains with G / D in .train(), training_loop.py:133-134).
//   ld_rng_advance  bumps the step word of the device-resident generator state (one launch per training iteration, inside the
//                   iteration's CUDA graph, so every replay draws fresh masks without host involvement);
//   ld_dropout      y = keep ? x / (1 - p) : 0 over a contiguous tensor, eight elements per Philox group (group = index / 8) —
//                   the standalone form (embedding / FFN-inner dropout) and the mask re-application of every backward pass.
// The fused forms live next to their kernels: attention probabilities (attention_fwd_sm100.cu / attention_bwd_sm100.cu) and
// LayerNorm(dropout(x) + residual) (norm.cu).
#include "common.cuh"
#include "runtime.h"
#include <algorithm>

namespace {
using namespace ld;

__global__ void rng_advance_kernel(uint32_t* state) { state[2] += 1u; }

template <typename TI, typename TO> struct Io8;
template <> struct Io8<__nv_bfloat16, __nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 a = *reinterpret_cast<const uint4*>(p);
        unpack_bf16x2(a.x, v[0], v[1]); unpack_bf16x2(a.y, v[2], v[3]); unpack_bf16x2(a.z, v[4], v[5]); unpack_bf16x2(a.w, v[6], v[7]);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p) = o;
    }
};
template <> struct Io8<float, float> {
    static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
};

template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long n8, const uint32_t* __restrict__ rng_state, uint32_t site,
               uint32_t thresh16, float scale) {
    DropoutRng rng;
    rng.init(rng_state, site, thresh16, scale);
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < n8; g += (long)gridDim.x * blockDim.x) {
        float v[8];
        Io8<T, T>::load(x + g * 8, v);
        const uint32_t keep = rng.keep8((uint64_t)g);
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ((keep >> k) & 1u) ? v[k] * scale : 0.0f;
        Io8<T, T>::store(y + g * 8, v);
    }
}
}  // namespace

extern "C" int ld_rng_advance(uint32_t* rng_state, void* stream) {
    LD_CHECK_ARG(rng_state != nullptr, "rng_advance: null state");
    rng_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(rng_state);
    ld::count_launch();
    LD_LAUNCH_CHECK("rng_advance");
    return 0;
}

extern "C" int ld_dropout(const void* x, void* y, int dtype, int64_t n, float dropout_p, const uint32_t* rng_state,
                          uint32_t rng_site, void* stream) {
    using namespace ld;
    LD_CHECK_ARG(x && y && n > 0 && n % 8 == 0, "dropout: n (%lld) must be a positive multiple of 8", (long long)n);
    LD_CHECK_ARG(dropout_p > 0.0f && dropout_p < 1.0f && rng_state, "dropout: p must be in (0, 1) with an rng_state");
    LD_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "dropout: 16-byte alignment");
    const uint32_t thresh16 = (uint32_t)(dropout_p * 65536.0f + 0.5f);
    const float scale = 65536.0f / (65536.0f - (float)thresh16);
    const long n8 = n / 8;
    const int grid = (int)std::max<long>(1, std::min<long>((n8 + 255) / 256, (long)sm_count() * 8));
    if (dtype == LD_BF16)
        dropout_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n8, rng_state, rng_site, thresh16, scale);
    else if (dtype == LD_F32)
        dropout_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, n8, rng_state, rng_site, thresh16, scale);
    else { set_last_error("dropout: bad dtype %d", dtype); return LD_ERR_INVALID_ARG; }
    count_launch();
    LD_LAUNCH_CHECK("dropout");
    return 0;
}
